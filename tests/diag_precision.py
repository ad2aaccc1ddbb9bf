"""Per-layer precision diagnostic (GPU box): where does the CUDA pipeline's embedding error come from?

    python tests/diag_precision.py [--n 2] [--out gpurun_out/diag_precision.txt]

TEST TOOLING (imports oracle/ as the checker, like the parity tests).  For n fixture latents it replays the synthesis
program layer by layer (cfr_program_run_range) and prints, per StyleGAN layer, the relative L2 error of the
normalised activation x_l = y_l * A + B against the fp32 oracle's layer output; then splits the embedding error
into its StyleGAN and ArcFace parts by crossing the two pipelines (our image -> oracle ArcFace, oracle image -> our ArcFace).
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2)
    ap.add_argument("--out", default=None)
    ap.add_argument("--hp", type=int, nargs="+", default=[-1],
                    help="split-precision prefix lengths to try (Engine hp_layers); -1 = the engine's default")
    args = ap.parse_args()
    from certifyingfacerecognition_b200.engine import Engine
    from oracle import fixtures
    from oracle import mc_path as M
    lines = []

    def say(*a):
        s = " ".join(str(v) for v in a)
        print(s, flush=True)
        lines.append(s)

    torch.set_num_threads(os.cpu_count())
    g_sd, f_sd = fixtures.build_models(cache_dir=os.path.join(ROOT, ".fixture_cache"))
    dirs = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "dirs.npy")))
    n = args.n
    w = torch.from_numpy(fixtures.latents(64)[32:32 + n])
    # ---- oracle, every layer kept
    ref_layers = {}
    with torch.no_grad():
        wp = M.truncation(w, g_sd)
        raw = M.synthesis(wp, g_sd, literal=False, tap=lambda name, t: ref_layers.__setitem__(name, t.clone()))
        ref_img = M.transform(M.postprocess(raw))
        ref_emb = M.iresnet50(ref_img, f_sd)

    for hp in args.hp:
        eng = Engine(g_sd, f_sd, dirs, torch.zeros(4, 512), chunk=n, keep_planar=True, hp_layers=None if hp < 0 else hp,
                     sparse_last=False)
        say(f"==== hp_layers = {eng.hp_layers}")
        one_engine(eng, n, w, ref_layers, ref_img, ref_emb, f_sd, say)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as fh:
            fh.write("\n".join(lines) + "\n")


def one_engine(eng, n, w, ref_layers, ref_img, ref_emb, f_sd, say):
    # ---- engine, layer by layer
    from certifyingfacerecognition_b200 import _lib as L
    from certifyingfacerecognition_b200.engine import PSI
    from oracle import mc_path as M
    syn = eng.synth
    syn.out_slot.zero_()
    L.check(eng.lib.cfr_truncate(L.ptr(w.cuda()), L.ptr(syn.w_avg), PSI, n, L.ptr(syn.wp2), eng._stream()))
    done = 0
    say("layer  res   C   rel_err(x)    max|d|   |  emulated fp16 storage (oracle/precision_study) for comparison: "
        "3.7e-4 at L1 growing to 4.2e-3 at L17")
    for (l, end_op, buf, res, c, pend) in syn.layer_marks:
        syn.run_range(done, end_op)
        done = end_op
        torch.cuda.synchronize()
        y = buf[:n * res * res * c].view(n, res, res, c).float()
        if pend is not None:
            A = pend[0][:n * c].view(n, 1, 1, c)
            B = pend[1][:n * c].view(n, 1, 1, c)
            y = y * A + B
        x = y.permute(0, 3, 1, 2).cpu()
        r = ref_layers[f"layer{l}"]
        rel = ((x - r).norm() / r.norm()).item()
        say(f"L{l:<2d}  {res:5d} {c:4d}   {rel:.3e}   {(x - r).abs().max().item():.3e}")
    syn.run_range(done, syn.num_launches)
    torch.cuda.synchronize()
    img = syn.img_planar[:n].cpu()
    say(f"img112: mean|d| {(img - ref_img).abs().mean().item():.3e}  max|d| {(img - ref_img).abs().max().item():.3e}"
        f"  (emulated: 1.7e-3)")
    eng.pipes[0].img_frm[:n].copy_(syn.img[:n])
    eng.frm.run()
    torch.cuda.synchronize()
    emb = eng.frm.emb[:n].cpu()

    def rep(tag, e):
        say(f"{tag:34s} L2 {[round(v, 4) for v in (e - ref_emb).norm(dim=1).tolist()]}  1-cos "
            f"{['%.1e' % (1 - v) for v in F.cosine_similarity(e, ref_emb).tolist()]}")
    rep("ours (GAN + ArcFace)", emb)
    with torch.no_grad():
        rep("our image -> oracle ArcFace", M.iresnet50(img, f_sd))
    # oracle image -> our ArcFace: write the oracle's 112^2 image into the program's NHWC fp16 input
    fimg = eng.pipes[0].img_frm                      # the FRM programs read their own copy of the images
    fimg[:n].zero_()
    fimg[:n, :, :, :3] = ref_img.permute(0, 2, 3, 1).to(fimg.dtype).cuda()
    eng.frm.run()
    torch.cuda.synchronize()
    rep("oracle image -> our ArcFace", eng.frm.emb[:n].cpu())
    say("ref emb norms", [round(v, 2) for v in ref_emb.norm(dim=1).tolist()])


if __name__ == "__main__":
    main()
