"""Generate tests/golden/attack_vectors.npz: the UNMODIFIED reference's attack path on the
fixture -- `get_dists_and_logits` (gen_utils.py:248-256) with autograd through its own StyleGAN + iresnet50 modules,
`compute_loss` (:160-223) for every loss type, the gradient of each loss with respect to the 5-D attribute offsets
(what `loss.backward()` gives `find_adversaries_pgd`, :378-382), the host helpers `init_deltas`, `proj2region`,
`sample_ellipsoid`, `check_deltas` on seeded inputs, and one short `find_adversaries_pgd` run.

    python -m oracle.make_golden_attack

TEST INFRASTRUCTURE ONLY (build container: needs /root/reference).  ~2 minutes of CPU.
"""
from __future__ import annotations

import os
import sys
import tempfile
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import fixtures, reference_shims as RS  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
N_IDS, N_GALLERY, B = 8, 5000, 3
LOSSES = ("xent", "away", "diff", "nearest", "dlr")


def main() -> None:
    assert RS.available(), "needs /root/reference"
    torch.set_num_threads(os.cpu_count())
    gold = np.load(os.path.join(GOLDEN, "reference_vectors.npz"))
    g_sd, f_sd = fixtures.build_models(cache_dir=os.path.join(ROOT, ".fixture_cache"))
    w_all = fixtures.latents(N_IDS)
    dir_mat = torch.from_numpy(gold["dirs"])                       # [5,512] (WrappedModel layout)
    dirs = dir_mat.T.contiguous()                                  # [512,5] (attack layout, gen_utils.py:607-631)
    red_ellipse_mat = torch.from_numpy(1.0 / gold["red_ellipse_mat_inv"]).float()     # diag of A: 1 / eps^2

    scratch = tempfile.mkdtemp(prefix="cfr_ref_attack_")
    RS.make_scratch(scratch, w_all, torch.zeros(N_IDS, 512), f_sd)
    os.chdir(scratch)
    ref = RS.import_reference("cpu")
    import attack_utils.proj_utils as ref_proj
    gu = ref.gen_utils
    model = ref.WrappedModel(dir_mat, "insightface", n_embs=N_IDS, load_embs=True)
    RS.load_stylegan_into(model.generator.model, g_sd)
    model.generator.model.eval()
    model.eval()
    generator, net, transform = model.generator, model.face_reco, model.transform
    rows = torch.from_numpy(np.load(os.path.join(GOLDEN, "votes_gallery.npz"))["rows"])
    gallery = fixtures.synthetic_gallery(rows, N_GALLERY)
    lat = torch.from_numpy(w_all[:B])
    labels = torch.arange(B)
    out = {}

    # ---- host helpers on seeded inputs ---------------------------------------------------------------------------
    torch.manual_seed(77)
    d_surf = gu.init_deltas(True, True, B, True, red_ellipse_mat, None, None)          # on the surface
    torch.manual_seed(78)
    d_in = gu.init_deltas(True, True, 16, False, red_ellipse_mat, None, None)          # inside
    out["init_surface"], out["init_inside"] = d_surf.numpy(), d_in.numpy()
    g = torch.Generator().manual_seed(5)
    pts = torch.randn(24, 5, generator=g) * torch.tensor([0.9, 0.3, 0.4, 1.2, 0.1])   # some inside, most outside
    proj, _ = ref_proj.proj2region(pts.clone(), proj_mat=None, ellipse_mat=red_ellipse_mat, to_subs=False, check=True,
                                   on_surface=False, diag_ellipse_mat=True)
    out["proj_in"], out["proj_out"] = pts.numpy(), proj.numpy()
    out["proj_mag"] = gu.check_deltas(proj, True, red_ellipse_mat, None, None).numpy()
    out["red_ellipse_mat"] = red_ellipse_mat.numpy()

    # ---- distances, losses and their gradients with respect to the attribute offsets ----------------------------
    t0 = time.time()
    deltas = d_surf.clone().detach().requires_grad_(True)
    pert = (dirs @ deltas.T).T
    all_dists, logits = gu.get_dists_and_logits(generator, net, lat + pert, transform, gallery, "insightface")
    out["deltas"], out["all_dists"] = deltas.detach().numpy(), all_dists.detach().numpy()
    for lt in LOSSES:
        loss = gu.compute_loss(all_dists, labels, loss_type=lt, use_probs=lt != "dlr")
        (grad,) = torch.autograd.grad(loss, deltas, retain_graph=True)
        out[f"loss_{lt}"], out[f"grad_{lt}"] = np.float64(loss.item()), grad.numpy()
        print(f"{lt}: loss {loss.item():.6f} |grad| {grad.norm(dim=1).tolist()}", flush=True)
    print(f"forward + 5 backward passes: {time.time() - t0:.0f} s", flush=True)

    # ---- one short PGD run of the reference itself (loose comparison: our driver estimates the gradient by central
    #      differences through the forward-only engine, so the trajectories are not bit-identical) --------------------
    t0 = time.time()
    ref_proj.set_seed(ref.device, seed=123)
    best, found, mags = gu.find_adversaries_pgd(
        generator, net, lat, labels, gallery, opt_name="SGD", lr=1e2, iters=4, momentum=0.9, frs_method="insightface",
        loss_type="xent", transform=transform, ellipse_mat=None, proj_mat=None, dirs=dirs, dirs_inv=None,
        red_ellipse_mat=red_ellipse_mat, random_init=True, rand_init_on_surf=True, lin_comb=True, restarts=2)
    out["pgd_best"], out["pgd_found"], out["pgd_mags"] = best.numpy(), found.numpy(), mags.numpy()
    print(f"reference find_adversaries_pgd: found {found.tolist()} magnitudes {mags.tolist()} ({time.time() - t0:.0f} s)")
    np.savez_compressed(os.path.join(GOLDEN, "attack_vectors.npz"), **out)
    print("wrote", os.path.join(GOLDEN, "attack_vectors.npz"))


if __name__ == "__main__":
    main()
