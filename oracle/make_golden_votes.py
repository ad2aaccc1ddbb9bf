"""Generate tests/golden/votes_<regime>.npz: >= 1000 Monte-Carlo samples per regime classified by the UNMODIFIED
reference (``/root/reference``: ``Smooth._sample_noise`` -> ``WrappedModel.forward`` -> ``lat2embs`` ->
``compute_probs`` -> ``argmax``), on the SURVEY.md section 8d fixture with its 64 decoy gallery rows.

    python -m oracle.make_golden_votes [--n 1000] [--regimes iso aniso] [--batch 100]

TEST INFRASTRUCTURE ONLY.  About 1 s of CPU per sample on 8 cores.  The reference's own code path runs; the only
additions are *recording* wrappers around ``certificate.sample_noise`` (smooth.py:134) and
``WrappedModel.compute_probs`` (smoothing_model.py:56-61) that keep what went through them.

Per regime the file holds
  noise  [N0+n,5]  f32   the tensors ``sample_noise`` returned (torch.randn under ``torch.manual_seed(seed)`` *
                         sigma) while ``Smooth.certify(z, x, 0, N0=100, n, alpha=0.001, batch)`` ran: the first N0 rows
                         are the selection pass, the rest the estimation pass
  emb    [N0+n,512] f32  the embeddings ``lat2embs`` produced for them
  pred   [N0+n]    i32   ``probs.argmax(1)``
  d1,d2  [N0+n]    f32   the reference's own smallest / second-smallest gallery distance (margin diagnostics)
  counts [N]    f64      estimation-pass vote counts
  cert_pred, cert_gap, cert_radius   what ``Smooth.certify`` returned (radius = sigma.min() * gap, certify.py:141)
  sigma, seed, and (votes_gallery.npz) the gallery itself: 8 true rows (tests/golden/reference_vectors.npz) +
  64 decoys = reference embeddings of w_0 + offsets @ dirs + Gaussian rows (oracle.fixtures.synthetic_gallery).
"""
from __future__ import annotations

import argparse
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
N_IDS = 8
N_GALLERY = 5000
SIGMA = 0.1
N0, ALPHA = 100, 0.001            # certify.py defaults (--N0 100 --alpha 0.001)


def decoy_offsets() -> np.ndarray:
    """[64,5] attribute-space offsets of the decoy rows around the certified identity (decoys =
    reference embeddings of w_0 + delta @ dirs).  On this fixture sigma = 0.1 moves the embedding by ~4 (|d emb /
    d attr| ~ 40), so rows 1-2 sigma away would leave the label ~1 % of the votes and certify() would always take
    the early exit; 30 axis rows at +-4, 5, 7 sigma plus 34 seeded random directions at 5 sigma give the label
    60-75 % with ~25 decoys drawing votes, i.e. the estimation pass, the abstain test and near-tie flips all fire."""
    rows = []
    for k in range(5):
        for a in (4.0, 5.0, 7.0):
            for s in (1.0, -1.0):
                v = np.zeros(5)
                v[k] = s * a * SIGMA
                rows.append(v)
    u = np.random.RandomState(11).randn(34, 5)
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    return np.vstack([np.array(rows), 5.0 * SIGMA * u]).astype(np.float32)


REGIMES = {
    # certify.py:85-95 -- scalar sigma ("isotropic") or sigma * red_ellipse_mat_inv ("anisotropic": 0.3 * eps^2 =
    # [.075, .075, .012, .075, .192], one direction far below and one well above the isotropic 0.1)
    "iso": (lambda rem_inv: torch.tensor([SIGMA]), 4321),
    "aniso": (lambda rem_inv: 0.3 * torch.from_numpy(rem_inv).float(), 4322),
}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=100)
    ap.add_argument("--regimes", nargs="+", default=list(REGIMES))
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    args = ap.parse_args()
    assert RS.available(), "needs /root/reference"
    torch.set_num_threads(args.threads)
    gold = np.load(os.path.join(GOLDEN, "reference_vectors.npz"))
    g_sd, f_sd = fixtures.build_models(cache_dir=os.path.join(ROOT, ".fixture_cache"))
    w_all = fixtures.latents(N_IDS)
    assert np.array_equal(w_all, gold["w_all"])
    dirs = torch.from_numpy(gold["dirs"])
    rem_inv = gold["red_ellipse_mat_inv"]

    scratch = tempfile.mkdtemp(prefix="cfr_ref_votes_")
    RS.make_scratch(scratch, w_all, torch.zeros(N_IDS, 512), f_sd)
    os.chdir(scratch)
    ref = RS.import_reference("cpu")
    model = ref.WrappedModel(dirs, "insightface", n_embs=N_IDS, load_embs=True)
    RS.load_stylegan_into(model.generator.model, g_sd)
    model.generator.model.eval()
    model.eval()

    # ---- gallery: true rows + decoys through the reference's lat2embs, + synthetic rows ---------------------------
    gal_path = os.path.join(GOLDEN, "votes_gallery.npz")
    if os.path.isfile(gal_path):
        gallery = torch.from_numpy(np.load(gal_path)["rows"])
        gallery = fixtures.synthetic_gallery(gallery, N_GALLERY)
    else:
        t0 = time.time()
        offs = torch.from_numpy(decoy_offsets())
        wd = torch.from_numpy(w_all[0:1]) + offs @ dirs
        with torch.no_grad():
            dec, _ = ref.gen_utils.lat2embs(model.generator, model.face_reco, wd, model.transform, few=False)
        rows = torch.cat([torch.from_numpy(gold["gallery"]), dec])
        np.savez_compressed(gal_path, rows=rows.numpy(), decoy_offsets=offs.numpy())
        gallery = fixtures.synthetic_gallery(rows, N_GALLERY)
        print(f"gallery: {rows.shape[0]} reference rows in {time.time() - t0:.0f} s", flush=True)
    model.orig_embs = gallery.clone()

    # ---- recording wrappers (the wrapped callables are the reference's own) --------------------------------------
    cert = ref.L2Certificate(1, device=ref.device)
    rec = {"noise": [], "emb": [], "pred": [], "d1": [], "d2": []}
    cert_sample = cert.sample_noise
    compute_probs = model.compute_probs

    def sample_noise(batch, theta):
        out = cert_sample(batch, theta)
        rec["noise"].append(out.detach().reshape(-1, 5).clone())
        return out

    def probs_rec(embedding):
        probs = compute_probs(embedding)
        rec["emb"].append(embedding.detach().cpu().clone())
        rec["pred"].append(probs.argmax(1).cpu())
        d = torch.cdist(embedding.cpu(), model.orig_embs, compute_mode="donot_use_mm_for_euclid_dist")
        two = d.topk(2, dim=1, largest=False).values
        rec["d1"].append(two[:, 0].clone())
        rec["d2"].append(two[:, 1].clone())
        print(f"  {sum(t.shape[0] for t in rec['emb'])} samples, {time.time() - t_start:.0f} s", flush=True)
        return probs

    cert.sample_noise = sample_noise
    model.compute_probs = probs_rec

    z = torch.from_numpy(w_all[0:1])
    x = torch.zeros(1, 5)
    for tag in args.regimes:
        mk_sigma, seed = REGIMES[tag]
        sigma = mk_sigma(rem_inv)
        for v in rec.values():
            v.clear()
        smooth = ref.Smooth(model, N_GALLERY, sigma, cert)
        t_start = time.time()
        torch.manual_seed(seed)
        # Smooth.certify (smooth.py:39-77): N0 selection samples, then (label == cAHat) n estimation samples
        cert_pred, cert_gap = smooth.certify(z, x, torch.tensor([0]), N0, args.n, ALPHA, args.batch, device=ref.device)
        out = {k: torch.cat(v).numpy() for k, v in rec.items()}
        total = out["noise"].shape[0]
        assert total in (N0, N0 + args.n) and out["emb"].shape == (total, 512)
        counts = np.bincount(out["pred"][N0:], minlength=N_GALLERY).astype(np.float64)
        np.savez_compressed(os.path.join(GOLDEN, f"votes_{tag}.npz"), noise=out["noise"].astype(np.float32),
                            emb=out["emb"].astype(np.float32), pred=out["pred"].astype(np.int32),
                            d1=out["d1"].astype(np.float32), d2=out["d2"].astype(np.float32), counts=counts,
                            sigma=sigma.numpy(), seed=np.int64(seed), identity=np.int64(0), n0=np.int64(N0),
                            alpha=np.float64(ALPHA), cert_pred=np.int64(cert_pred), cert_gap=np.float64(cert_gap),
                            cert_radius=np.float64(float(sigma.min()) * cert_gap))
        nz = np.nonzero(counts)[0]
        print(f"{tag}: certify -> ({cert_pred}, {cert_gap:.4f}); votes {dict(zip(nz.tolist(), counts[nz].astype(int).tolist()))} "
              f"median margin {np.median(out['d2'] - out['d1']):.4f}  ({time.time() - t_start:.0f} s)", flush=True)


if __name__ == "__main__":
    main()
