"""Gallery builder: W latents -> StyleGAN-1024 -> ArcFace embeddings -> `embs_<frm>.pth`.

The reference builds its gallery with `lat2embs` over every identity and `torch.save`s the [N,512] tensor
(main_attack.py:210-216; smoothing_model.py:48-53 is the in-memory variant).  This is the same job on the CUDA
engine: identities are sharded over ranks (torchrun, no collective on the data path; rank 0 gathers the rows),
the file it writes is what `WrappedModel(..., load_embs=True, embs_file=...)` / the reference read back.

  python tools/build_gallery.py --latents data/stylegan_ffhq_1M/w.npy --out embeddings/embs_insightface.pth
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/build_gallery.py ...

Without --latents / --generator / --frm it runs on the seeded synthetic fixtures (certifyingfacerecognition_b200/synthetic.py) -- that mode
is the throughput measurement of SURVEY.md section 8d config 5(i) (identities/s).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--latents", default=None, help="w.npy ([N,512] or reshapable); default: synthetic fixture")
    ap.add_argument("--generator", default=None, help="stylegan_ffhq.pth state dict; default: synthetic fixture")
    ap.add_argument("--frm", default=None, help="iresnet50 backbone.pth state dict; default: synthetic fixture")
    ap.add_argument("--identities", dest="num", type=int, default=512, help="identities (synthetic mode / truncation of --latents)")
    ap.add_argument("--chunk", type=int, default=128)
    ap.add_argument("--out", default=None, help="embs .pth to write (rank 0)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("build_gallery needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")

    from certifyingfacerecognition_b200.engine import Engine
    if args.generator is None or args.frm is None:
        from certifyingfacerecognition_b200 import synthetic as fixtures          # synthetic weights only (test / measurement mode)
        g_sd, f_sd = fixtures.build_models()
    if args.generator is not None:
        g_sd = torch.load(args.generator, map_location="cpu")
    if args.frm is not None:
        f_sd = torch.load(args.frm, map_location="cpu")
    if args.latents is not None:
        w = torch.from_numpy(np.load(args.latents).reshape(-1, 512).astype(np.float32))[: args.num if args.num > 0 else None]
    else:
        from certifyingfacerecognition_b200 import synthetic as fixtures
        w = torch.from_numpy(fixtures.latents(args.num))
    n = w.shape[0]
    dirs = torch.zeros(5, 512)               # the perturbation directions play no role when embedding plain latents
    eng = Engine(g_sd, f_sd, dirs, torch.zeros(8, 512), chunk=args.chunk, device=f"cuda:{local}")

    lo, hi = rank * n // world, (rank + 1) * n // world          # contiguous identity shard of this rank
    eng.embed_latents(w[lo:min(hi, lo + args.chunk)])             # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    emb = eng.embed_latents(w[lo:hi])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        parts = [torch.empty((r + 1) * n // world - r * n // world, 512, device=emb.device) for r in range(world)]
        dist.all_gather(parts, emb)
        emb = torch.cat(parts, 0)
    if rank == 0:
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            torch.save(emb.cpu(), args.out)
        print(json.dumps({"metric": "gallery identities embedded / s", "value": n / dt, "unit": "identities/s",
                          "n_gpus": world, "identities": n, "chunk": args.chunk, "seconds": dt,
                          "out": args.out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
