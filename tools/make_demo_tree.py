"""Create, under a scratch directory, the cwd-relative file layout the reference's certify.py reads (boundaries/*.npy,
data/stylegan_ffhq_1M/w.npy, embeddings/embs_insightface.pth, weights/ms1mv3_arcface_r50/backbone.pth,
models/pretrain/stylegan_ffhq.pth) from the seeded synthetic weights, so the CLI can be tried without any checkpoint:

  python tools/make_demo_tree.py /tmp/demo --identities 8
  cd /tmp/demo && python /root/repo/certify.py --face-recog-model insightface --outfile out/c.tsv --sigma 0.1 --N 200
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("dest")
    ap.add_argument("--identities", type=int, default=8)
    ap.add_argument("--chunk", type=int, default=8)
    args = ap.parse_args()
    from certifyingfacerecognition_b200 import synthetic
    from certifyingfacerecognition_b200.attack_utils import proj_utils
    from certifyingfacerecognition_b200.engine import Engine
    d = args.dest
    g_sd, f_sd = synthetic.build_models()
    dirs = np.load(os.path.join(ROOT, "tests", "golden", "dirs.npy"))
    for sub in ("boundaries", "data/stylegan_ffhq_1M", "embeddings", "weights/ms1mv3_arcface_r50", "models/pretrain"):
        os.makedirs(os.path.join(d, sub), exist_ok=True)
    for k, attr in enumerate(proj_utils.ATTRS):
        np.save(os.path.join(d, "boundaries", f"stylegan_ffhq_{attr}_w_boundary.npy"), dirs[k:k + 1].astype(np.float64))
    w = synthetic.latents(args.identities)
    np.save(os.path.join(d, "data", "stylegan_ffhq_1M", "w.npy"), w)
    torch.save(f_sd, os.path.join(d, "weights", "ms1mv3_arcface_r50", "backbone.pth"))
    torch.save(g_sd, os.path.join(d, "models", "pretrain", "stylegan_ffhq.pth"))
    eng = Engine(g_sd, f_sd, torch.from_numpy(dirs), torch.zeros(1, 512), chunk=args.chunk)
    torch.save(eng.embed_latents(torch.from_numpy(w)).cpu(), os.path.join(d, "embeddings", "embs_insightface.pth"))
    print("wrote", d)


if __name__ == "__main__":
    main()
