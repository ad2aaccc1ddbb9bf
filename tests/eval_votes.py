"""Precision / throughput trade-off of the split-precision prefix (GPU box): strict top-1 agreement with the reference's
golden votes (tests/golden/votes_*.npz) for several Engine(hp_layers=...) settings.

    python tests/eval_votes.py --hp 0 4 6 8 10 [--out gpurun_out/eval_votes.txt]

TEST TOOLING (loads the oracle fixtures as the checker)."""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hp", type=int, nargs="+", default=[-1])
    ap.add_argument("--chunk", type=int, default=125)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from certifyingfacerecognition_b200.engine import Engine
    from certifyingfacerecognition_b200.smoothing.smooth import lower_confidence_bound
    from oracle import fixtures
    from scipy.stats import norm
    g_sd, f_sd = fixtures.build_models(cache_dir=os.path.join(ROOT, ".fixture_cache"))
    gold = np.load(os.path.join(GOLDEN, "reference_vectors.npz"))
    dirs = torch.from_numpy(gold["dirs"])
    z = torch.from_numpy(gold["w_all"][0:1])
    rows = torch.from_numpy(np.load(os.path.join(GOLDEN, "votes_gallery.npz"))["rows"])
    gallery = fixtures.synthetic_gallery(rows, 5000)
    lines = []

    def say(s):
        print(s, flush=True)
        lines.append(s)

    for hp in args.hp:
        eng = Engine(g_sd, f_sd, dirs, gallery, chunk=args.chunk, frm_group=2 if args.chunk >= 100 else 1,
                     hp_layers=None if hp < 0 else hp)
        for tag in ("iso", "aniso"):
            if not os.path.isfile(os.path.join(GOLDEN, f"votes_{tag}.npz")):
                continue
            v = np.load(os.path.join(GOLDEN, f"votes_{tag}.npz"))
            noise = torch.from_numpy(v["noise"])
            n, n0 = noise.shape[0], int(v["n0"])
            torch.cuda.synchronize()
            t0 = time.time()
            counts, ex = eng.sample_votes(z, torch.zeros(1, 5), torch.from_numpy(v["sigma"]), n, noise=noise,
                                          want_pred=True, want_emb=True)
            torch.cuda.synchronize()
            dt = time.time() - t0
            pred, emb = ex["pred"].cpu().long(), ex["emb"].cpu()
            pref, eref = torch.from_numpy(v["pred"]).long(), torch.from_numpy(v["emb"])
            agree = pred == pref
            err = (emb - eref).norm(dim=1)
            na = int((pred[n0:] == 0).sum())
            pbar = lower_confidence_bound(na, n - n0, float(v["alpha"]))
            radius = float(v["sigma"].min()) * float(norm.ppf(pbar)) if pbar >= 0.5 else 0.0
            margins = sorted(round(float(m), 4) for m in (v["d2"] - v["d1"])[(~agree).numpy()])
            say(f"hp={eng.hp_layers} {tag}: n={n} agreement {agree.float().mean().item():.4%} flips {int((~agree).sum())} "
                f"cos_min {F.cosine_similarity(emb, eref).min().item():.6f} embL2 median {err.median().item():.4f} "
                f"max {err.max().item():.4f} | nA {na} vs ref {int(v['counts'][0])} radius {radius:.5f} vs ref "
                f"{float(v['cert_radius']):.5f} ({abs(radius / float(v['cert_radius']) - 1):.3%}) | {n / dt:.0f} samples/s "
                f"| flipped margins {margins}")
        del eng
        torch.cuda.empty_cache()
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as fh:
            fh.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
