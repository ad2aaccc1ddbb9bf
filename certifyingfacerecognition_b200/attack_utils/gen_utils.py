"""The reference's attack_utils/gen_utils.py as far as the B200 path goes: what certify.py uses (:17-41 constants,
get_latent_codes :44-46, get_all_matrices :607-631) and the linear-combination PGD attack driver of SURVEY.md section
8f-4 (get_optim :88-96, get_dists :99-105, compute_loss :160-223, init_deltas :226-245, get_dists_and_logits :248-256,
check_deltas :319-332, find_adversaries_pgd :335-393).

The reference back-propagates through StyleGAN + the FRM to get d loss / d delta.  The attack searches a FIVE-dimensional
attribute space, so here the gradient is a central difference through the forward-only CUDA engine: 10 extra forward
samples per identity and iteration (at ~7 600 forward samples/s that is ~700 PGD iterations/s per GPU) and no backward
kernels.  Success is always decided by a real forward classification, exactly as in the reference; trajectories are not
bit-identical to autograd's (tests/test_attack_gpu.py checks the estimated gradients against the unmodified reference's
autograd gradients).  Out of scope: the 512-D `--no-lin-comb` variant, AutoAttack (third-party APGD / FAB), result files
and plots (main_attack.py)."""
from __future__ import annotations

import os.path as osp

import numpy as np
import torch

import torch.nn.functional as F

from .proj_utils import get_projection_matrices, in_ellps, proj2region, sample_ellipsoid, sq_distance

INP_RESOLS = {"insightface": 112, "facenet": 160, "facenet-vggface2": 160}
STD = 0.5
MEAN = 0.5
EMB_SIZE = 512
LAT_SPACE = "w"
DATASET = "ffhq"
GAN_NAME = "stylegan"
FRS_METHODS = ["insightface", "facenet", "facenet-vggface2"]
ORIG_DATA_PATH = f"data/{GAN_NAME}_{DATASET}_1M"
LAT_CODES_PATH = osp.join(ORIG_DATA_PATH, f"{LAT_SPACE}.npy")
WEIGHTS_PATH = "weights/ms1mv3_arcface_r50/backbone.pth"          # main_attack.py:43
STYLEGAN_PATH = "models/pretrain/stylegan_ffhq.pth"                # models/model_settings.py:50


def _device():
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def get_latent_codes(generator=None) -> torch.Tensor:
    """gen_utils.py:44-46 (+ preprocess, mod_stylegan_generator.py:160-190, W branch): [N,512] float32."""
    return torch.from_numpy(np.load(LAT_CODES_PATH).reshape(-1, EMB_SIZE).astype(np.float32))


def get_all_matrices(attrs2drop=None, scale_factor=1.0, device=None):
    """gen_utils.py:607-631 -> (proj_mat, ellipse_mat, ellipse_mat_inv, dirs, dirs_inv, red_ellipse_mat,
    red_ellipse_mat_inv) as float32 tensors on the device."""
    dev = device or _device()
    proj_mat, ellipse_mat, dirs, red_ellipse_mat, _ = get_projection_matrices(
        dataset=DATASET, gan_name=GAN_NAME, attrs2drop=list(attrs2drop or []), scale_factor=scale_factor)
    dirs = torch.tensor(dirs, dtype=torch.float32, device=dev)
    proj_mat = torch.tensor(proj_mat, dtype=torch.float32, device=dev)
    ellipse_mat = torch.tensor(ellipse_mat, dtype=torch.float32, device=dev)
    assert proj_mat.size(0) == proj_mat.size(1) == EMB_SIZE
    red_ellipse_mat = torch.tensor(red_ellipse_mat, dtype=torch.float32, device=dev)
    assert red_ellipse_mat.size(0) == dirs.size(1)
    red_ellipse_mat_inv = 1 / red_ellipse_mat
    dirs_inv = torch.linalg.pinv(dirs)
    ellipse_mat_inv = torch.linalg.inv(ellipse_mat)
    return proj_mat, ellipse_mat, ellipse_mat_inv, dirs, dirs_inv, red_ellipse_mat, red_ellipse_mat_inv


# ------------------------------------------------------------------------------------------------------------------
# Attack driver, linear-combination variant
# ------------------------------------------------------------------------------------------------------------------
LOSS_TYPES = ["away", "nearest", "diff", "xent", "dlr"]


def get_optim(deltas, optim_name="SGD", lr=0.001, momentum=0.9):
    """gen_utils.py:88-96."""
    if optim_name == "SGD":
        return torch.optim.SGD([deltas], lr=lr, momentum=momentum)
    if optim_name == "Adam":
        return torch.optim.Adam([deltas], lr=lr)
    if optim_name == "RMSProp":
        return torch.optim.RMSprop([deltas], lr=lr)
    raise ValueError(f"unknown optimiser '{optim_name}'")


def get_dists(embs1, embs2, method="insightface"):
    """gen_utils.py:99-105: Euclidean distances for ArcFace, 1 - cosine for the (already normalised) FaceNet embeddings."""
    if method == "insightface":
        return torch.cdist(embs1, embs2, compute_mode="donot_use_mm_for_euclid_dist")
    return 1 - embs1 @ embs2.T


def per_sample_loss(all_dists, labels, loss_type="away", use_probs=True, scale_dists=True):
    """The per-identity terms whose mean is `compute_loss` (gen_utils.py:160-223) -> [b]."""
    lab = labels.view(-1, 1)
    vals = F.softmax(-(all_dists / np.sqrt(EMB_SIZE) if scale_dists else all_dists), dim=1) if use_probs else all_dists
    target = torch.gather(vals, 1, lab).squeeze(1)
    masked = torch.scatter(vals, 1, lab, -1.0 if use_probs else float("inf"))
    nearest = masked.max(1).values if use_probs else masked.min(1).values
    sign = 1.0 if use_probs else -1.0                # probabilities of the label are minimised, distances maximised
    if loss_type == "away":
        return sign * target
    if loss_type == "nearest":
        return -sign * nearest
    if loss_type == "diff":
        return sign * (target - nearest)
    if loss_type == "xent":
        assert use_probs, "xent loss should be used together with probs"
        # (the reference has already divided `all_dists` by sqrt(512) for the softmax above when it divides by sqrt(512)
        #  again here, gen_utils.py:163-165,208-209: the cross-entropy scores are -d / 512; reproduced literally)
        scores = -all_dists / EMB_SIZE if scale_dists else -all_dists
        return -F.cross_entropy(scores, labels, reduction="none")
    if loss_type == "dlr":
        assert not use_probs, "dlr loss works in terms of logits"
        top = torch.topk(-all_dists, k=3, dim=1, largest=True, sorted=True).values
        return -(target - nearest) / (top[:, 0] - top[:, 2])             # difference-of-logits ratio
    raise ValueError(f"unknown loss '{loss_type}'")


def compute_loss(all_dists, labels, loss_type="away", use_probs=True, scale_dists=True):
    """gen_utils.py:160-223."""
    return per_sample_loss(all_dists, labels, loss_type, use_probs, scale_dists).mean()


def init_deltas(random_init, lin_comb, n_vecs, on_surface, ellipse_mat, proj_mat, dirs):
    """gen_utils.py:226-245 (lin_comb): uniform in the low-dimensional ellipsoid, optionally pushed onto its surface.
    `ellipse_mat` is the diagonal as a vector (red_ellipse_mat)."""
    if not lin_comb:
        raise NotImplementedError("the 512-D `--no-lin-comb` attack is not built (SURVEY.md section 8f-4)")
    if not random_init:
        return torch.zeros(n_vecs, ellipse_mat.shape[0], device=ellipse_mat.device)
    ell_mat = torch.diag(ellipse_mat)
    deltas = sample_ellipsoid(ell_mat, n_vecs=n_vecs)
    if on_surface:
        deltas, _ = proj2region(deltas, proj_mat=None, ellipse_mat=ell_mat, check=True, to_subs=False, dirs=None,
                                on_surface=True)
    return deltas.clone().detach()


def _embedder(generator):
    """The forward map latents [b,512] -> embeddings [b,512]: an Engine, a WrappedModel (its engine) or anything with
    `embed_latents` (lat2embs gen_utils.py:108-139 incl. the truncation trick)."""
    eng = getattr(generator, "engine", generator)
    if not hasattr(eng, "embed_latents"):
        raise TypeError("generator must be an Engine / WrappedModel of this package (or provide embed_latents)")
    return eng.embed_latents


def get_dists_and_logits(generator, net, codes, transform, orig_embs, frs_method):
    """gen_utils.py:248-256.  `net` / `transform` are folded into the engine and ignored."""
    embs = _embedder(generator)(codes)
    all_dists = get_dists(embs, orig_embs.to(embs.device), method=frs_method)
    return all_dists, -1.0 * all_dists


def check_deltas(deltas, lin_comb, red_ellipse_mat, ellipse_mat, proj_mat, check=True):
    """gen_utils.py:319-332 (lin_comb): squared ellipsoid norm of every delta (1 = on the budget surface)."""
    if not lin_comb:
        raise NotImplementedError("the 512-D `--no-lin-comb` attack is not built")
    ell = torch.diag(red_ellipse_mat)
    magnitudes = sq_distance(ell, deltas.unsqueeze(2))
    if check:
        assert in_ellps(deltas.T, ell, atol=1e-3)
    return magnitudes


def loss_gradient_fd(generator, lat_codes, deltas, labels, orig_embs, dirs, red_ellipse_mat, frs_method, loss_type,
                     fd_step=0.1):
    """d mean_i loss_i / d delta  [b,n_dirs] by central differences: delta_i +- h_k e_k with h_k = fd_step * eps_k (a
    tenth of the attribute's budget), all 2 * n_dirs * b perturbed latents in one engine call."""
    b, k = deltas.shape
    h = fd_step / torch.sqrt(red_ellipse_mat)                                        # [k]
    steps = torch.cat([torch.diag(h), -torch.diag(h)])                                # [2k,k]
    pert = (deltas.unsqueeze(1) + steps.unsqueeze(0)).reshape(b * 2 * k, k)           # [b*2k,k]
    codes = lat_codes.repeat_interleave(2 * k, dim=0) + pert @ dirs.T
    d, _ = get_dists_and_logits(generator, None, codes, None, orig_embs, frs_method)
    ls = per_sample_loss(d, labels.repeat_interleave(2 * k), loss_type, use_probs=loss_type != "dlr").view(b, 2, k)
    return (ls[:, 0] - ls[:, 1]) / (2 * h) / b


def find_adversaries_pgd(generator, net, lat_codes, labels, orig_embs, opt_name, lr, iters, momentum, frs_method,
                         loss_type, transform, ellipse_mat, proj_mat, dirs, dirs_inv, red_ellipse_mat, random_init=True,
                         rand_init_on_surf=True, lin_comb=True, restarts=5, fd_step=0.1):
    """gen_utils.py:335-393: projected gradient descent on the attribute offsets of every identity in the batch, with
    restarts; an identity's first successful delta is kept.  dirs: [512, n_dirs]; red_ellipse_mat: [n_dirs] (1 / eps^2).
    -> (best_deltas [b,n_dirs] cpu, found_adv [b] bool, squared ellipsoid norms [b])."""
    if not lin_comb:
        raise NotImplementedError("the 512-D `--no-lin-comb` attack needs d loss / d latent (512 coordinates); only the "
                                  "5-D linear-combination attack is built on the forward-only engine")
    dev = lat_codes.device
    dirs, red_ellipse_mat, labels = dirs.to(dev), red_ellipse_mat.to(dev), labels.to(dev)
    orig_embs = orig_embs.to(dev)
    b = lat_codes.size(0)
    best_deltas = torch.zeros(b, dirs.size(1), device=dev)
    found_adv = torch.zeros(b, dtype=torch.bool, device=dev)
    for _ in range(restarts):
        deltas = init_deltas(random_init, lin_comb, b, rand_init_on_surf, red_ellipse_mat, proj_mat, dirs)
        deltas = deltas.clone().detach().requires_grad_(True)
        optim = get_optim(deltas, optim_name=opt_name, lr=lr, momentum=momentum)
        for _ in range(iters):
            with torch.no_grad():
                all_dists, _ = get_dists_and_logits(generator, net, lat_codes + deltas @ dirs.T, transform, orig_embs,
                                                    frs_method)
                success = all_dists.argmin(1) != labels
                new = success & ~found_adv
                best_deltas[new] = deltas.detach()[new]
                found_adv |= success
                if bool(found_adv.all()):
                    break
                grad = loss_gradient_fd(generator, lat_codes, deltas.detach(), labels, orig_embs, dirs, red_ellipse_mat,
                                        frs_method, loss_type, fd_step)
            optim.zero_grad()
            deltas.grad = grad
            optim.step()
            with torch.no_grad():
                proj, _ = proj2region(deltas.detach(), proj_mat=None, ellipse_mat=red_ellipse_mat, to_subs=False,
                                      check=True, on_surface=False, diag_ellipse_mat=True)
                deltas[:] = proj
        if bool(found_adv.all()):
            break
    magnitudes = check_deltas(best_deltas, lin_comb, red_ellipse_mat, ellipse_mat, proj_mat)
    return best_deltas.detach().cpu(), found_adv, magnitudes
