"""Per-op device-time breakdown of the synthesis and ArcFace programs (CUDA events between ops, warm caches).

    python tools/profile_program.py --chunk 50 [--out profiles/ops_rNN.tsv]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(prog, reps=3):
    from certifyingfacerecognition_b200 import _lib as L
    n = prog.num_launches
    buf = (C.c_float * n)()
    best = np.full(n, 1e30)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(reps):
        L.check(prog.lib.cfr_program_run_timed(prog.handle, st, buf, n))
        best = np.minimum(best, np.array(buf[:]))
    labels = [prog.lib.cfr_program_op_label(prog.handle, i).decode() for i in range(n)]
    flops = [prog.lib.cfr_program_op_flops(prog.handle, i) for i in range(n)]
    return labels, best, flops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunk", type=int, default=50)
    ap.add_argument("--out", default=None)
    ap.add_argument("--hp", type=int, default=-1, help="split-precision prefix (Engine hp_layers); -1 = default")
    ap.add_argument("--frm", default="insightface", choices=["insightface", "facenet"])
    args = ap.parse_args()
    from certifyingfacerecognition_b200 import synthetic as fixtures
    from certifyingfacerecognition_b200.engine import Engine
    g_sd, f_sd = fixtures.build_models()
    dirs = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "dirs.npy")))
    if args.frm == "facenet":
        f_sd = fixtures.facenet_weights()
    eng = Engine(g_sd, f_sd, dirs, torch.zeros(8, 512), chunk=args.chunk,
                 frm="insightface" if args.frm == "insightface" else "facenet-vggface2",
                 hp_layers=None if args.hp < 0 else args.hp)
    eng.embed_latents(torch.from_numpy(fixtures.latents(args.chunk)))
    torch.cuda.synchronize()
    lines = []
    total = 0.0
    for name, prog in (("synth", eng.synth), ("frm", eng.frm)):
        labels, ms, flops = timed(prog)
        for i, (l, t, f) in enumerate(zip(labels, ms, flops)):
            tf = f / (t * 1e-3) / 1e12 if f > 0 else 0.0
            lines.append(f"{name}\t{i}\t{t * 1e3 / args.chunk:9.2f}\t{tf:8.1f}\t{l}")
        sub = ms.sum()
        total += sub
        lines.append(f"{name}\tTOTAL\t{sub * 1e3 / args.chunk:9.2f}\t\tus/sample ({sub:.3f} ms per chunk of {args.chunk})")
    lines.append(f"all\tTOTAL\t{total * 1e3 / args.chunk:9.2f}\t\tus/sample -> {args.chunk / (total * 1e-3):.0f} samples/s")
    text = "prog\top\tus/sample\tTFLOP/s\tlabel\n" + "\n".join(lines)
    print(text)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
