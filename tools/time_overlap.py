"""A/B of the sampler's two-stream overlap: one call of `--num` MC samples (several groups), overlap on / off.

    python tools/time_overlap.py --num 1000 [--reps 5]
    CFR_DEBUG_KNOBS=1 CFR_SPLIT_SMS=120,28 python tools/time_overlap.py    # persistent-grid caps synthesis / FRM
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num", type=int, default=1000)
    ap.add_argument("--chunk", type=int, default=125)
    ap.add_argument("--frm-group", type=int, default=2)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    from certifyingfacerecognition_b200 import synthetic as fixtures
    from certifyingfacerecognition_b200.engine import Engine
    g_sd, f_sd = fixtures.build_models()
    dirs = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "dirs.npy")))
    lat = torch.from_numpy(fixtures.latents(4))
    eng = Engine(g_sd, f_sd, dirs, torch.zeros(8, 512), chunk=args.chunk, frm_group=args.frm_group)
    eng.set_gallery(fixtures.synthetic_gallery(eng.embed_latents(lat).cpu(), 5000))
    x, sigma = torch.zeros(1, 5), torch.tensor([0.1])
    res = {}
    for on in (True, False, True, False):
        eng.set_overlap(on)
        eng.sample_votes(lat[0:1], x, sigma, args.num, seed=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(args.reps):
            eng.sample_votes(lat[r % 4:r % 4 + 1], x, sigma, args.num, seed=r)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        res.setdefault(on, []).append(ms)
        print(f"overlap={on}: {ms:.2f} ms per call of {args.num} samples -> {args.num / ms * 1e3:.0f} samples/s", flush=True)
    print({k: min(v) for k, v in res.items()}, "split", os.environ.get("CFR_SPLIT_SMS"))


if __name__ == "__main__":
    main()
