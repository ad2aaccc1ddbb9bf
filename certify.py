"""Drop-in for the reference's certify.py (:16-157): same flags / defaults, same TSV
(``idx label predict correct gap radius time``), backed by the B200 engine.

Single GPU:   python certify.py --face-recog-model insightface --outfile out/cert.tsv --sigma 0.1
8 GPUs:       python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 certify.py ...
              (the MC samples of every identity are split across the ranks, the int64 vote counts are summed
              with one NCCL all-reduce per _sample_noise; rank 0 writes the TSV)
"""
import argparse
import datetime
import os
import os.path as osp
from time import time

import torch

from certifyingfacerecognition_b200.attack_utils.gen_utils import FRS_METHODS, get_all_matrices
from certifyingfacerecognition_b200.models.smoothing_model import WrappedModel
from certifyingfacerecognition_b200.smoothing.certificate import L2Certificate
from certifyingfacerecognition_b200.smoothing.smooth import Smooth

try:
    from tqdm import tqdm
except ImportError:                                      # pragma: no cover
    tqdm = lambda it: it


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="Certify face recognition examples")
    parser.add_argument("--face-recog-model", required=True, choices=FRS_METHODS, type=str,
                        help="type of model to load for face recognition")
    parser.add_argument("--outfile", required=True, type=str, help="output csv file")
    parser.add_argument("--sigma", type=float, required=True,
                        help="noise hyperparameter, required for initialization in isotropic_dd and ancer")
    parser.add_argument("--anisotropic-sigma", action="store_true", default=False,
                        help="Whether to use Anisotropic Sigma for certification")
    parser.add_argument("--skip", type=int, default=1, help="skip examples in the dataset")
    parser.add_argument("--max", type=int, default=-1, help="stop after a certain number of examples")
    parser.add_argument("--batch-sz", type=int, default=100, help="certification batch size")
    parser.add_argument("--N0", type=int, default=100)
    parser.add_argument("--N", type=int, default=100000, help="number of samples to use")
    parser.add_argument("--alpha", type=float, default=0.001, help="failure probability")
    parser.add_argument("--load-n-embs", type=int, default=1_000_000,
                        help="num of embs. Default is all of them (1M)")
    # additions (not in the reference)
    parser.add_argument("--seed", type=int, default=1234, help="Philox key of the device noise stream")
    parser.add_argument("--chunk", type=int, default=32, help="samples per GAN+FRM program run")
    return parser


def main(argv=None) -> None:
    args = build_parser().parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    group = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
        group = dist.group.WORLD
    device = torch.device("cuda", torch.cuda.current_device())

    dirs = get_all_matrices(device=device)[3].T.contiguous()            # certify.py:71
    model = WrappedModel(dirs, args.face_recog_model, n_embs=args.load_n_embs, load_embs=True, embs_file=None,
                         chunk=args.chunk)
    dataset = model.latents.to(device)
    certificate = L2Certificate(1, device=device)
    if args.anisotropic_sigma:
        print("Using anisotropic sigma")
        sigma = args.sigma * get_all_matrices(device=device)[6].to(device)   # certify.py:88-93
    else:
        sigma = torch.tensor([args.sigma], device=device)

    if rank == 0:
        parent_dir = osp.dirname(args.outfile)
        if parent_dir and not osp.exists(parent_dir):
            os.makedirs(parent_dir, exist_ok=True)
        with open(args.outfile, "w+") as f:
            print("idx\tlabel\tpredict\tcorrect\tgap\tradius\ttime", file=f, flush=True)

    num_classes = dataset.shape[0]
    print(f"Found {num_classes} classes")
    num_dirs = dirs.shape[0]
    print(f"Found {num_dirs} directions")
    x = torch.zeros((1, num_dirs), device=device)
    smoothed_classifier = Smooth(model, num_classes, sigma, certificate, seed=args.seed, process_group=group)

    for i in tqdm(range(num_classes)):
        if (i + 1) % args.skip != 0:
            continue
        if (i + 1) == args.max:
            break
        z, label = dataset[i].to(device), torch.tensor([i], device=device)
        before_time = time()
        prediction, gap = smoothed_classifier.certify(z.unsqueeze(0), x, label, args.N0, args.N, args.alpha,
                                                      args.batch_sz, device=device)
        after_time = time()
        correct = int(prediction == label)
        radius = sigma.min().item() * gap
        time_elapsed = str(datetime.timedelta(seconds=(after_time - before_time)))
        if rank == 0:
            with open(args.outfile, "a") as f:
                print("{}\t{}\t{}\t{}\t{:.3}\t{:.3}\t{}".format(i, label.item(), prediction, correct, gap, radius,
                                                              time_elapsed), file=f, flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
