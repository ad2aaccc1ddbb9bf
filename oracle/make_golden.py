"""Generate tests/golden/*.npz by running the UNMODIFIED reference (from /root/reference) on the
seeded fixtures.  Run once in the build container:   python -m oracle.make_golden

TEST INFRASTRUCTURE ONLY.  The reference cannot travel to the GPU box, so its outputs are committed
as small fixtures and ``tests/test_oracle_golden.py`` pins ``oracle/mc_path.py`` against them.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import fixtures, reference_shims as RS  # noqa: E402
from oracle import mc_path as M  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
N_IDS = 8            # identities in the tiny reference-side gallery
SIGMA = 0.1


def stats(t: torch.Tensor) -> np.ndarray:
    """per-sample (mean, std, absmax)"""
    f = t.detach().float().flatten(1)
    return torch.stack([f.mean(1), f.std(1), f.abs().amax(1)], dim=1).numpy()


def main() -> None:
    assert RS.available(), "needs /root/reference"
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLDEN, exist_ok=True)
    cache = os.path.join(ROOT, ".fixture_cache")
    g_sd, f_sd = fixtures.build_models(cache_dir=cache)
    w_all = fixtures.latents(N_IDS)

    scratch = tempfile.mkdtemp(prefix="cfr_ref_")
    RS.make_scratch(scratch, w_all, torch.zeros(N_IDS, 512), f_sd)   # gallery filled in below
    os.chdir(scratch)
    ref = RS.import_reference("cpu")

    # ---- geometry (host setup) -------------------------------------------------------------
    mats = ref.gen_utils.get_all_matrices()
    dirs = mats[3].T.contiguous().cpu().numpy().astype(np.float32)          # certify.py:71
    rem_inv = mats[6].cpu().numpy().astype(np.float32)                       # certify.py:88
    rem = mats[5].cpu().numpy().astype(np.float32)
    np.save(os.path.join(GOLDEN, "dirs.npy"), dirs)

    # ---- the reference model objects, fixture weights loaded ------------------------------
    model = ref.WrappedModel(torch.from_numpy(dirs), "insightface", n_embs=N_IDS, load_embs=True)
    RS.load_stylegan_into(model.generator.model, g_sd)
    model.generator.model.eval()
    model.eval()

    # ---- stage-by-stage vectors for 4 perturbed latents -----------------------------------
    g = torch.Generator().manual_seed(77)
    p = torch.randn(4, 1, 1, 5, generator=g) * SIGMA
    z = torch.from_numpy(w_all[0:1])
    w_pert = z + p.squeeze(2).squeeze(1) @ torch.from_numpy(dirs)
    w_in = torch.cat([w_pert[:2], torch.from_numpy(w_all[1:3])])             # 2 perturbed + 2 other ids
    layer_stats, block_stats = {}, {}
    hooks = []
    syn = model.generator.model.synthesis
    for li in range(M.NUM_LAYERS):
        hooks.append(getattr(syn, f"layer{li}").register_forward_hook(
            lambda m, i, o, li=li: layer_stats.__setitem__(li, stats(o))))
    net = model.face_reco
    hooks.append(net.prelu.register_forward_hook(lambda m, i, o: block_stats.__setitem__("stem", stats(o))))
    for li in range(1, 5):
        for bi, blk in enumerate(getattr(net, f"layer{li}")):
            hooks.append(blk.register_forward_hook(
                lambda m, i, o, k=f"layer{li}.{bi}": block_stats.__setitem__(k, stats(o))))
    with torch.no_grad():
        out = model.generator.easy_synthesize(w_in, latent_space_type="w")
        img = out["image"]
        img112 = model.transform(img)
        emb = net(img112)
    for h in hooks:
        h.remove()
    block_keys = ["stem"] + [f"layer{li}.{bi}" for li, n in enumerate(M.IRESNET50_LAYERS, 1) for bi in range(n)]

    # gallery for the end-to-end run: true embeddings of the 8 ids, via the reference's lat2embs
    with torch.no_grad():
        gal, _ = ref.gen_utils.lat2embs(model.generator, net, torch.from_numpy(w_all), model.transform, few=False)
    model.orig_embs = gal.clone()

    # ---- end-to-end Smooth.certify / predict through the reference ------------------------
    cert = ref.L2Certificate(1, device=ref.device)
    x = torch.zeros(1, 5)
    e2e = {}
    for tag, sigma in (("iso", torch.tensor([SIGMA])), ("aniso", 2.0 * torch.from_numpy(rem_inv))):
        smooth = ref.Smooth(model, N_IDS, sigma, cert)
        torch.manual_seed(1234)
        counts0 = smooth._sample_noise(z, x, 4, 4, device=ref.device)
        torch.manual_seed(1234)
        pred, gap = smooth.certify(z, x, torch.tensor([0]), 4, 12, 0.001, 4, device=ref.device)
        e2e[tag + "_counts0"] = counts0
        e2e[tag + "_pred"] = np.int64(pred)
        e2e[tag + "_gap"] = np.float64(gap)
        e2e[tag + "_sigma"] = sigma.numpy()
    # a mislabelled identity takes the early-exit branch (smooth.py:66-68)
    smooth = ref.Smooth(model, N_IDS, torch.tensor([SIGMA]), cert)
    torch.manual_seed(99)
    pred_wrong, gap_wrong = smooth.certify(z, x, torch.tensor([3]), 4, 8, 0.001, 4, device=ref.device)

    # ---- closed-form KATs through the (shimmed) reference methods -------------------------
    kat_in = np.array([[100, 100], [990, 1000], [1000, 1000], [9990, 10000], [60, 100], [0, 10], [7, 8]])
    kat_p = np.array([smooth._lower_confidence_bound(int(a), int(n), 0.001) for a, n in kat_in])
    kat_gap = np.array([cert.compute_gap(float(v)) for v in kat_p])

    sl = slice(0, 1024, 64)
    np.savez_compressed(
        os.path.join(GOLDEN, "reference_vectors.npz"),
        dirs=dirs, red_ellipse_mat=rem, red_ellipse_mat_inv=rem_inv,
        w_all=w_all, p=p.numpy(), w_in=w_in.numpy(),
        wp=model.generator.model.truncation(w_in).numpy(),
        layer_stats=np.stack([layer_stats[i] for i in range(M.NUM_LAYERS)]),
        image_stats=stats(img), image_sub=img[:, :, sl, sl].numpy(),
        img112_0=img112[0].numpy(), img112_stats=stats(img112),
        block_stats=np.stack([block_stats[k] for k in block_keys]),
        emb=emb.numpy(), gallery=gal.numpy(),
        pred_wrong=np.int64(pred_wrong), gap_wrong=np.float64(gap_wrong),
        kat_in=kat_in, kat_p=kat_p, kat_gap=kat_gap, **e2e)
    print("golden written:", os.listdir(GOLDEN))
    print("emb norms", emb.norm(dim=1), "gallery pdist", torch.cdist(gal, gal))
    print("e2e", {k: v for k, v in e2e.items()}, pred_wrong, gap_wrong)


if __name__ == "__main__":
    main()
