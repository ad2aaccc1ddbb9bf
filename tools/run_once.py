"""Tiny driver for ncu captures: builds the engine at a small chunk and runs the synthesis + ArcFace programs
`--reps` times (first repetition = warm-up)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunk", type=int, default=8)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    from certifyingfacerecognition_b200 import synthetic as fixtures
    from certifyingfacerecognition_b200.engine import Engine
    g_sd, f_sd = fixtures.build_models()
    dirs = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "dirs.npy")))
    eng = Engine(g_sd, f_sd, dirs, torch.zeros(8, 512), chunk=args.chunk)
    w = torch.from_numpy(fixtures.latents(args.chunk))
    for _ in range(args.reps):
        e = eng.embed_latents(w)
    torch.cuda.synchronize()
    print("ok", float(e.abs().mean()))


if __name__ == "__main__":
    main()
