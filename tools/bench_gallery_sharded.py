"""Config 5(ii) with the gallery sharded over ranks (partition C): N synthetic gallery rows split
contiguously over the ranks, b replicated queries per call, per-rank tensor-core match -> 8-byte keys -> all-gather ->
unsigned min -> votes.  Queries are planted next to known rows spread over every shard, so the result is checked.

  python tools/bench_gallery_sharded.py --rows 1000000                                   # one GPU, one shard
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_gallery_sharded.py --rows 1000000
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--queries", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        group = dist.group.WORLD
    from certifyingfacerecognition_b200.gallery_shard import ShardedGallery, shard_bounds
    lo, hi = shard_bounds(args.rows, world, rank)
    g = torch.Generator(device="cuda").manual_seed(1000 + rank)
    rows = torch.randn(hi - lo, 512, generator=g, device="cuda")
    shard = ShardedGallery(rows, lo, args.rows, max_b=args.queries, process_group=group)
    # planted queries: query i sits next to global row t_i; the rank that owns the row contributes it, others zeros
    b = args.queries
    targets = (torch.arange(b, dtype=torch.int64) * (args.rows // b) + 17) % args.rows
    q = torch.zeros(b, 512, device="cuda")
    mine = (targets >= lo) & (targets < hi)
    q[mine.cuda()] = rows[(targets[mine] - lo).cuda()]
    if world > 1:
        dist.all_reduce(q, group=group)                       # every rank now holds all the planted rows
    q += 0.05 * torch.randn(b, 512, generator=torch.Generator(device="cuda").manual_seed(7), device="cuda")
    counts = torch.zeros(args.rows, dtype=torch.int64, device="cuda")
    pred = shard.match_vote(q, counts, want_pred=True)        # warm-up + correctness
    torch.cuda.synchronize()
    ok = bool(torch.equal(pred.cpu().long(), targets))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    e0.record()
    for _ in range(args.reps):
        shard.match_vote(q, counts)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "gallery match queries / s (gallery sharded over ranks)", "value": args.reps * b / (ms.item() * 1e-3),
                          "unit": "queries/s", "n_gpus": world, "rows": args.rows, "rows_per_rank": hi - lo, "queries": b,
                          "ms_per_call": ms.item() / args.reps, "planted_rows_found": ok,
                          "matcher": "tensor-core" if shard.matcher is not None else "exact fp32"}))
    if not ok:
        raise SystemExit("sharded match returned wrong rows")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
