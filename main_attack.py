"""Semantic adversarial attacks on the B200 engine -- the reference's `main_attack.py` for `--attack-type manual`
(projected gradient descent on the attribute offsets, `attack_utils/gen_utils.py:335-393`), same flags
(`attack_utils/opts.py`), same cwd-relative inputs (`boundaries/`, `data/stylegan_ffhq_1M/w.npy`,
`models/pretrain/stylegan_ffhq.pth`, `weights/ms1mv3_arcface_r50/backbone.pth`, `embeddings/embs_1M_<frm>.pth`) and the
same outputs under `exp_results/<output-dir>/`:

    results/results_chunk<i>of<chunks>.pth   {'deltas', 'successes', 'magnitudes'} of the identities that were broken
    logs/results_chunk<i>of<chunks>.txt      successes:<n> / instances:<n> / avg_mags:<mean sqrt ellipsoid norm>
    results.txt                              totals over the chunk logs (main_attack.py --eval-files, without the figure)

    python main_attack.py --output-dir demo --load-embs --embs-file embeddings/embs_insightface.pth --chunks 2 --iters 10

The gradient of the loss with respect to the (five) attribute offsets is a central difference through the forward-only
engine instead of a backward pass (see attack_utils/gen_utils.py).  Not built: `--no-lin-comb` (512-D PGD) and the
AutoAttack variants (`fab-t`, `fab`, `apgd-*`): they exit with an explanation.
"""
from __future__ import annotations

import argparse
import glob
import os
import os.path as osp
import sys
import time

import numpy as np
import torch

ROOT = osp.dirname(osp.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from certifyingfacerecognition_b200.attack_utils import gen_utils as G                      # noqa: E402
from certifyingfacerecognition_b200.attack_utils.proj_utils import ATTRS, set_seed          # noqa: E402

ATTACKS = ["fab-t", "fab", "apgd-ce", "apgd-dlr", "apgd-t", "manual"]
OPTIMS = ["SGD", "Adam", "RMSProp"]


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Compute semantic adversaries")
    p.add_argument("--lr", type=float, default=1e+2, help="Learning rate")
    p.add_argument("--momentum", type=float, default=0.9, help="Momentum for SGD")
    p.add_argument("--loss", type=str, default="xent", choices=G.LOSS_TYPES, help="Loss to optimize")
    p.add_argument("--optim", type=str, default="SGD", choices=OPTIMS, help="Optimizer to use")
    p.add_argument("--no-lin-comb", action="store_true", default=False)
    p.add_argument("--attack-type", type=str, default="manual", choices=ATTACKS, help="Attack to perform")
    p.add_argument("--iters", type=int, default=10, help="Optimization iterations per instance")
    p.add_argument("--restarts", type=int, default=10, help="random restarts per instance")
    p.add_argument("--n-target-classes", type=int, default=10)
    p.add_argument("--attrs2drop", nargs="+", default=[], choices=list(ATTRS.keys()))
    p.add_argument("--scale-factor", type=float, default=1.0, help="Factor for scaling Sigma")
    p.add_argument("--not-on-surf", action="store_true", default=False)
    p.add_argument("--output-dir", type=str, required=True, help="Directory (under exp_results/) for the results")
    p.add_argument("--face-recog-method", type=str, default="insightface", choices=G.FRS_METHODS)
    p.add_argument("--chunks", type=int, default=50_000, help="num of chunks in which to break the dataset")
    p.add_argument("--num-chunk", type=int, default=None, help="index of chunk to evaluate on")
    p.add_argument("--eval-files", action="store_true", default=False, help="only aggregate the chunk logs")
    p.add_argument("--load-embs", action="store_true", default=False, help="Whether to load embs from file")
    p.add_argument("--load-n-embs", type=int, default=1_000_000)
    p.add_argument("--embs-file", type=str, default=None)
    p.add_argument("--seed", type=int, default=0, help="for deterministic behavior")
    # this implementation only
    p.add_argument("--batch-size", type=int, default=4,
                   help="identities attacked together (the reference's generator batch size is 4); every identity costs "
                        "11 forward samples per iteration")
    p.add_argument("--chunk", type=int, default=64, help="forward samples per engine program run")
    p.add_argument("--fd-step", type=float, default=0.1, help="central-difference step as a fraction of each attribute's budget")
    p.add_argument("--frm-weights", type=str, default=None, help="FRM state dict (default: the reference's ArcFace path)")
    args = p.parse_args(argv)
    args.output_dir = osp.join("exp_results", args.output_dir)
    args.lin_comb = not args.no_lin_comb
    args.results_dir = osp.join(args.output_dir, "results")
    args.logs_dir = osp.join(args.output_dir, "logs")
    for d in (args.output_dir, args.results_dir, args.logs_dir):
        os.makedirs(d, exist_ok=True)
    args.final_results = osp.join(args.output_dir, "results.txt")
    return args


def log(args, text: str) -> None:
    print(text, flush=True)
    with open(osp.join(args.output_dir, "log.txt"), "a") as fh:
        fh.write(text + "\n")


def save_results(results, deltas, successes, magnitudes, num_chunk, args) -> str:
    """gen_utils.py:413-437: same file names and dictionary keys."""
    filename = f"results_chunk{num_chunk}of{args.chunks}"
    data_file = osp.join(args.results_dir, f"{filename}.pth")
    if int(successes.sum()) != 0:
        torch.save({"deltas": deltas[successes].detach(), "successes": torch.nonzero(successes).detach(),
                    "magnitudes": magnitudes[successes].detach()}, data_file)
    log_file = osp.join(args.logs_dir, f"{filename}.txt")
    with open(log_file, "w") as fh:
        fh.write("\n".join(f"{k}:{v}" for k, v in results.items()) + "\n")
    return log_file


def eval_chunk(model, lat_codes, embs, num_chunk, mats, args) -> str:
    """gen_utils.py:634-752 without the image dumps: attack every identity of the chunk in batches, re-verify the reported
    adversaries with an independent forward pass, write the chunk's result / log files."""
    proj_mat, ellipse_mat, _, dirs, dirs_inv, red_ellipse_mat, _ = mats
    dev = embs.device
    chunk_length = len(lat_codes) / args.chunks
    assert float(chunk_length).is_integer(), "Partition of set should be exact"
    chunk_length = int(chunk_length)
    start = num_chunk * chunk_length
    lats = lat_codes[start:start + chunk_length]
    deltas, successes, magnitudes, all_labels = [], [], [], []
    t0 = time.time()
    for idx, b0 in enumerate(range(0, chunk_length, args.batch_size)):
        codes = lats[b0:b0 + args.batch_size].to(dev)
        set_seed(dev, seed=args.seed + num_chunk * chunk_length + idx)
        labels = torch.arange(b0, b0 + codes.size(0), device=dev) + start
        d, succ, mags = G.find_adversaries_pgd(
            model, None, codes, labels, embs, opt_name=args.optim, lr=args.lr, iters=args.iters, momentum=args.momentum,
            frs_method=args.face_recog_method, loss_type=args.loss, transform=None, ellipse_mat=ellipse_mat,
            proj_mat=proj_mat, dirs=dirs, dirs_inv=dirs_inv, red_ellipse_mat=red_ellipse_mat, random_init=True,
            rand_init_on_surf=not args.not_on_surf, lin_comb=True, restarts=args.restarts, fd_step=args.fd_step)
        # check_advs (gen_utils.py:396-410): a reported adversary must still be one under a fresh forward pass
        if bool(succ.any()):
            dist, _ = G.get_dists_and_logits(model, None, codes[succ] + d.to(dev)[succ] @ dirs.T, None, embs,
                                             args.face_recog_method)
            still = dist.argmin(1) != labels[succ]
            succ = succ.clone()
            succ[succ.clone()] = still
        deltas.append(d)
        successes.append(succ.cpu())
        magnitudes.append(mags.cpu())
        all_labels.append(labels.cpu())
    deltas, successes, magnitudes = torch.cat(deltas), torch.cat(successes), torch.cat(magnitudes)
    n_succ = int(successes.sum())
    avg = float(magnitudes[successes].sqrt().mean()) if n_succ else 0
    log(args, f"chunk {num_chunk}/{args.chunks}: {n_succ} advs for {len(successes)} IDs -> avg. pert.: {avg:3.4f} "
              f"({time.time() - t0:.1f} s)")
    return save_results({"successes": n_succ, "instances": len(successes), "avg_mags": avg}, deltas, successes,
                        magnitudes, num_chunk, args)


def eval_files(log_files, args) -> None:
    """main_attack.py --eval-files: totals over the chunk logs -> results.txt (the accuracy-vs-perturbation figure of
    gen_utils.py:440-604 is not produced)."""
    tot = {"successes": 0, "instances": 0}
    weighted = 0.0
    for lf in sorted(log_files):
        vals = dict(line.strip().split(":", 1) for line in open(lf) if ":" in line)
        tot["successes"] += int(vals["successes"])
        tot["instances"] += int(vals["instances"])
        weighted += float(vals["avg_mags"]) * int(vals["successes"])
    avg = weighted / tot["successes"] if tot["successes"] else 0.0
    rate = tot["successes"] / max(1, tot["instances"])
    text = (f"successes:{tot['successes']}\ninstances:{tot['instances']}\nsuccess_rate:{rate:.6f}\n"
            f"avg_mags:{avg:.6f}\nchunks:{len(log_files)}\n")
    with open(args.final_results, "w") as fh:
        fh.write(text)
    log(args, f"{tot['successes']} adversaries for {tot['instances']} identities ({100 * rate:.2f} %), avg. pert. {avg:.4f} "
              f"-> {args.final_results}")


def main(argv=None) -> None:
    args = parse_args(argv)
    if args.eval_files:
        eval_files(glob.glob(osp.join(args.logs_dir, "results_chunk*of*.txt")), args)
        return
    if args.attack_type != "manual":
        sys.exit(f"--attack-type {args.attack_type}: the AutoAttack variants (third-party APGD / FAB, backward passes) are not "
                 "built on the forward-only engine; use --attack-type manual")
    if not args.lin_comb:
        sys.exit("--no-lin-comb (PGD on 512 latent coordinates) needs d loss / d latent, i.e. a backward pass through "
                 "StyleGAN + the FRM; only the 5-D linear-combination attack is built")
    if not torch.cuda.is_available():
        sys.exit("main_attack.py needs a CUDA device (there is no CPU fallback)")
    from certifyingfacerecognition_b200.models.smoothing_model import WrappedModel
    dev = torch.device("cuda")
    t0 = time.time()
    mats = G.get_all_matrices(list(args.attrs2drop), scale_factor=args.scale_factor, device=dev)
    # the engine is built on the full five-direction matrix (its Monte-Carlo entry points are fixed to it); the attack
    # itself only uses `embed_latents` and the (possibly reduced) direction set of `mats`
    dirs = G.get_all_matrices([], device=dev)[3] if args.attrs2drop else mats[3]      # [512, 5]
    lat_codes = G.get_latent_codes()
    embs_file = args.embs_file or osp.join("embeddings", f"embs_1M_{args.face_recog_method}.pth")
    frm_state = torch.load(args.frm_weights, map_location="cpu") if args.frm_weights else None
    model = WrappedModel(dirs.T.contiguous(), args.face_recog_method, n_embs=args.load_n_embs, load_embs=args.load_embs,
                         embs_file=embs_file, frm_state=frm_state, chunk=args.chunk)
    if not args.load_embs and args.embs_file is not None:     # main_attack.py:211-216: generated gallery is saved
        torch.save(model.orig_embs.cpu(), args.embs_file)
    embs = model.orig_embs[:args.load_n_embs].to(dev)
    log(args, f"Loaded {embs.size(0)} embeddings, {len(lat_codes)} latent codes")
    chunks = range(args.chunks) if args.num_chunk is None else [args.num_chunk]
    logs = [eval_chunk(model, lat_codes, embs, c, mats, args) for c in chunks]
    if args.num_chunk is None:
        eval_files(logs, args)
    log(args, f"Finished. Total time spent: {time.time() - t0}s")


if __name__ == "__main__":
    main()
