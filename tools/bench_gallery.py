"""BASELINE config 5 (ii): gallery-match throughput at N identities, exact SIMT kernel vs tensor-core matcher.

    python tools/bench_gallery.py --rows 1000000 --b 250
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--b", type=int, default=250)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    from certifyingfacerecognition_b200 import _lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(0)
    gal = torch.randn(args.rows, 512, generator=g).cuda()
    emb = (gal[torch.randint(0, args.rows, (args.b,), generator=g).cuda()] + 0.3 * torch.randn(args.b, 512, generator=g).cuda()).contiguous()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    pred1 = torch.zeros(args.b, dtype=torch.int32, device="cuda")
    pred2 = torch.zeros_like(pred1)
    counts = torch.zeros(args.rows, dtype=torch.int64, device="cuda")
    keys = torch.full((args.b,), -1, dtype=torch.int64, device="cuda")
    m = C.c_void_p()
    L.check(lib.cfr_matcher_create(L.ptr(gal), args.rows, 256, st, C.byref(m)))

    def timed(fn):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.iters

    t_simt = timed(lambda: L.check(lib.cfr_match_vote(L.ptr(emb), args.b, L.ptr(gal), args.rows, L.ptr(keys), L.ptr(pred1), L.ptr(counts), st)))
    t_tc = timed(lambda: L.check(lib.cfr_matcher_run(m, L.ptr(emb), args.b, L.ptr(pred2), L.ptr(counts), st)))
    agree = float((pred1 == pred2).float().mean())
    flops = 2.0 * args.b * args.rows * 1536
    print(json.dumps({"rows": args.rows, "b": args.b, "simt_ms": t_simt, "tc_ms": t_tc, "agree": agree,
                      "simt_queries_per_s": args.b / t_simt * 1e3, "tc_queries_per_s": args.b / t_tc * 1e3,
                      "tc_tflops_split_gemm": flops / (t_tc * 1e-3) / 1e12,
                      "tc_gallery_gbs": args.rows * 1536 * 2 / (t_tc * 1e-3) / 1e9}))
    lib.cfr_matcher_destroy(m)


if __name__ == "__main__":
    main()
