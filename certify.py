"""Certification CLI on the B200 engine.  Command line, defaults and the TSV it writes follow the reference's
certify.py (:16-157) so that downstream scripts keep working:

    idx <TAB> label <TAB> predict <TAB> correct <TAB> gap <TAB> radius <TAB> time

One GPU:   python certify.py --face-recog-model insightface --outfile out/cert.tsv --sigma 0.1
N GPUs:    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 certify.py ...
           every identity's Monte-Carlo samples are split over the ranks by global sample index, one NCCL int64
           all-reduce per vote tally, rank 0 owns the output file.
"""
from __future__ import annotations

import argparse
import datetime
import os
import time
from dataclasses import dataclass
from typing import Iterator, Optional

import torch

from certifyingfacerecognition_b200.attack_utils.gen_utils import FRS_METHODS, get_all_matrices
from certifyingfacerecognition_b200.models.smoothing_model import WrappedModel
from certifyingfacerecognition_b200.smoothing.certificate import L2Certificate
from certifyingfacerecognition_b200.smoothing.smooth import Smooth

TSV_HEADER = "\t".join(("idx", "label", "predict", "correct", "gap", "radius", "time"))
TSV_ROW = "{}\t{}\t{}\t{}\t{:.3}\t{:.3}\t{}"          # reference certify.py:146-157 (3 significant digits for gap / radius)

# (flag, kwargs) in the reference's order; defaults are the reference's
_REFERENCE_FLAGS = (
    ("--face-recog-model", dict(required=True, type=str, choices=FRS_METHODS, help="face recognition network")),
    ("--outfile", dict(required=True, type=str, help="TSV to write")),
    ("--sigma", dict(required=True, type=float, help="smoothing noise scale")),
    ("--anisotropic-sigma", dict(action="store_true", default=False, help="scale sigma per attribute direction")),
    ("--skip", dict(type=int, default=1, help="certify every skip-th identity")),
    ("--max", dict(type=int, default=-1, help="stop when this identity count is reached")),
    ("--batch-sz", dict(type=int, default=100, help="Monte-Carlo samples per batch")),
    ("--N0", dict(type=int, default=100, help="samples of the selection pass")),
    ("--N", dict(type=int, default=100000, help="samples of the estimation pass")),
    ("--alpha", dict(type=float, default=0.001, help="failure probability of the certificate")),
    ("--load-n-embs", dict(type=int, default=1_000_000, help="gallery rows to load")),
)
_EXTRA_FLAGS = (
    ("--seed", dict(type=int, default=1234, help="key of the device-side Philox noise stream")),
    ("--chunk", dict(type=int, default=32, help="samples per synthesis + recognition program run")),
    ("--frm-weights", dict(type=str, default=None, help="state dict of the recognition network (needed for the FaceNet "
                                                         "variants: the reference downloads those weights at run time)")),
)


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description="Randomized-smoothing certification of a face recognition pipeline")
    for flag, kw in _REFERENCE_FLAGS + _EXTRA_FLAGS:
        ap.add_argument(flag, **kw)
    return ap


@dataclass
class Ranks:
    rank: int = 0
    world: int = 1
    group: Optional[object] = None

    @classmethod
    def from_env(cls) -> "Ranks":
        r = cls(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        if r.world > 1:
            import torch.distributed as dist
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            dist.init_process_group("nccl")
            r.group = dist.group.WORLD
        return r

    def close(self) -> None:
        if self.world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


def selected_identities(n: int, skip: int, stop: int) -> Iterator[int]:
    """The reference's loop filter (certify.py:122-127): 1-based position divisible by `skip`, stop AT position `stop`."""
    for i in range(n):
        pos = i + 1
        if pos % skip:
            continue
        if pos == stop:
            return
        yield i


def smoothing_scale(args, device) -> torch.Tensor:
    if not args.anisotropic_sigma:
        return torch.tensor([args.sigma], device=device)
    print("Using anisotropic sigma")
    return args.sigma * get_all_matrices(device=device)[6].to(device)     # diag of the inverse ellipse matrix (:88-93)


def main(argv=None) -> None:
    args = build_parser().parse_args(argv)
    ranks = Ranks.from_env()
    device = torch.device("cuda", torch.cuda.current_device())

    directions = get_all_matrices(device=device)[3].T.contiguous()       # [5, 512] attribute directions (:71)
    frm_state = torch.load(args.frm_weights, map_location="cpu") if args.frm_weights else None
    model = WrappedModel(directions, args.face_recog_model, n_embs=args.load_n_embs, load_embs=True, embs_file=None,
                         chunk=args.chunk, frm_state=frm_state)
    latents = model.latents.to(device)
    sigma = smoothing_scale(args, device)
    n_ids, n_dirs = latents.shape[0], directions.shape[0]
    print(f"Found {n_ids} classes")
    print(f"Found {n_dirs} directions")

    if ranks.rank == 0:
        folder = os.path.dirname(args.outfile)
        if folder:
            os.makedirs(folder, exist_ok=True)
        with open(args.outfile, "w") as fh:
            fh.write(TSV_HEADER + "\n")

    smoother = Smooth(model, n_ids, sigma, L2Certificate(1, device=device), seed=args.seed, process_group=ranks.group)
    origin = torch.zeros((1, n_dirs), device=device)
    try:
        from tqdm import tqdm
        todo = tqdm(list(selected_identities(n_ids, args.skip, args.max)))
    except ImportError:                                  # pragma: no cover
        todo = selected_identities(n_ids, args.skip, args.max)
    for i in todo:
        label = torch.tensor([i], device=device)
        t0 = time.time()
        predicted, gap = smoother.certify(latents[i:i + 1], origin, label, args.N0, args.N, args.alpha, args.batch_sz,
                                          device=device)
        elapsed = datetime.timedelta(seconds=time.time() - t0)
        if ranks.rank == 0:
            row = TSV_ROW.format(i, i, predicted, int(predicted == i), gap, sigma.min().item() * gap, str(elapsed))
            with open(args.outfile, "a") as fh:
                fh.write(row + "\n")
    ranks.close()


if __name__ == "__main__":
    main()
