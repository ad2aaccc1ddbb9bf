"""Emulate reduced-precision *storage* (fp32 math, rounded operands/activations) of the planned CUDA
pipeline on the CPU, to choose the operand dtype before writing kernels.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.precision_study [fp16|bf16] [n_samples]

Rounding points mirror DESIGN.md "data layout": conv operands (activations + pre-scaled weights)
are 16-bit, accumulation and every epilogue are fp32, each tensor written to HBM is 16-bit.
"""
from __future__ import annotations

import math
import os
import sys

import torch
import torch.nn.functional as F

from . import fixtures
from . import mc_path as M


def make_q(kind: str):
    dt = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[kind]
    return lambda t: t.to(dt).to(torch.float32)


def synthesis_q(wp, sd, q, fold_norm: bool = False):
    b = wp.shape[0]

    def epilogue_y(x, layer):
        p = f"synthesis.layer{layer}.epilogue."
        x = x + sd[p + "apply_noise.noise"] * sd[p + "apply_noise.weight"].view(1, -1, 1, 1)
        x = x + sd[p + "bias"].view(1, -1, 1, 1)
        return F.leaky_relu(x, 0.2)

    def norm_style(y32, yq, layer):
        """A, B from fp32 statistics; applied to the stored (rounded) y."""
        p = f"synthesis.layer{layer}.epilogue."
        c = y32.shape[1]
        mean = y32.mean(dim=[2, 3], keepdim=True)
        var = (y32 * y32).mean(dim=[2, 3], keepdim=True) - mean * mean
        rstd = 1.0 / torch.sqrt(var + 1e-8)
        style = F.linear(wp[:, layer], sd[p + "style_mod.dense.linear.weight"]) / math.sqrt(512) \
            + sd[p + "style_mod.dense.wscale.bias"].view(1, -1)
        style = style.view(-1, 2, c, 1, 1)
        a = rstd * (style[:, 0] + 1)
        bb = style[:, 1] - mean * a
        return q(yq * a + bb)

    y = epilogue_y(sd["synthesis.layer0.first_layer"].repeat(b, 1, 1, 1), 0)
    x = norm_style(y, q(y), 0)
    for layer in range(1, M.NUM_LAYERS):
        if layer % 2 == 0:
            weq, _ = M.upconv_equiv_weight(sd, layer)
            # 4-phase sub-pixel form: summed taps are rounded once (as the packed weights will be)
            up = F.interpolate(x, scale_factor=2, mode="nearest")
            raw = torch.zeros(b, weq.shape[0], up.shape[2], up.shape[3])
            full = F.conv2d(up, weq, padding=1)          # exact math reference for shape
            # emulate per-phase rounded weights
            xin = x
            H = xin.shape[2]
            xp = F.pad(xin, (1, 1, 1, 1))
            for a in (0, 1):
                for bcol in (0, 1):
                    rows = [(0, [0]), (1, [1, 2])] if a == 0 else [(1, [0, 1]), (2, [2])]
                    cols = [(0, [0]), (1, [1, 2])] if bcol == 0 else [(1, [0, 1]), (2, [2])]
                    acc = 0
                    for ro, rt in rows:
                        for co, ct in cols:
                            wsum = sum(weq[:, :, i, j] for i in rt for j in ct)
                            wsum = q(wsum)
                            patch = xp[:, :, ro:ro + H, co:co + H]
                            acc = acc + torch.einsum("oc,bchw->bohw", wsum, patch)
                    raw[:, :, a::2, bcol::2] = acc
            assert (raw - full).abs().max() < 0.05 * full.abs().max() + 1e-3
            y32 = epilogue_y(M._blur(q(raw)), layer)
        else:
            w = sd[f"synthesis.layer{layer}.conv.weight"]
            wq = q(w * (math.sqrt(2.0) / math.sqrt(w.shape[1] * 9)))
            y32 = epilogue_y(F.conv2d(x, wq, padding=1), layer)
        x = norm_style(y32, q(y32), layer)
    w = sd["synthesis.output8.conv.weight"]
    return F.conv2d(x, w) * (1.0 / math.sqrt(w.shape[1])) + sd["synthesis.output8.bias"].view(1, -1, 1, 1)


def iresnet50_q(x, sd, q, fp32_residual: bool = False):
    def fold(conv_w, bn_p):
        scale = sd[bn_p + ".weight"] / torch.sqrt(sd[bn_p + ".running_var"] + 1e-5)
        shift = sd[bn_p + ".bias"] - sd[bn_p + ".running_mean"] * scale
        return conv_w * scale.view(-1, 1, 1, 1), shift

    def pre_affine(bn_p):
        scale = sd[bn_p + ".weight"] / torch.sqrt(sd[bn_p + ".running_var"] + 1e-5)
        shift = sd[bn_p + ".bias"] - sd[bn_p + ".running_mean"] * scale
        return scale, shift

    x = q(x)
    w, bsh = fold(sd["conv1.weight"], "bn1")
    x = M._prelu(F.conv2d(x, q(w), padding=1) + bsh.view(1, -1, 1, 1), sd["prelu.weight"])
    xs = x if fp32_residual else q(x)          # residual stream
    for li, nblocks in enumerate(M.IRESNET50_LAYERS, start=1):
        for bi in range(nblocks):
            p = f"layer{li}.{bi}."
            stride = 2 if bi == 0 else 1
            s1, t1 = pre_affine(p + "bn1")
            w1, b1 = fold(sd[p + "conv1.weight"], p + "bn2")
            # bn1 scale folded into conv1's input channels, shift via border-exact bias image
            w1 = w1 * s1.view(1, -1, 1, 1)
            xin = q(xs)
            ones = torch.ones(1, xin.shape[1], xin.shape[2], xin.shape[3]) * t1.view(1, -1, 1, 1)
            wq = q(w1)
            tbias = F.conv2d(ones, fold(sd[p + "conv1.weight"], p + "bn2")[0], padding=1)
            h = F.conv2d(xin, wq, padding=1) + tbias + b1.view(1, -1, 1, 1)
            h = q(M._prelu(h, sd[p + "prelu.weight"]))
            w2, b2 = fold(sd[p + "conv2.weight"], p + "bn3")
            out = F.conv2d(h, q(w2), stride=stride, padding=1) + b2.view(1, -1, 1, 1)
            if bi == 0:
                wd, bd = fold(sd[p + "downsample.0.weight"], p + "downsample.1")
                identity = F.conv2d(xin, q(wd), stride=stride) + bd.view(1, -1, 1, 1)
            else:
                identity = xs
            xs = out + identity
            if not fp32_residual:
                xs = q(xs)
    s, t = pre_affine("bn2")
    x = q(xs) * s.view(1, -1, 1, 1) + t.view(1, -1, 1, 1)
    x = torch.flatten(q(x), 1)
    x = F.linear(x, q(sd["fc.weight"]), sd["fc.bias"])
    return M._bn(x, sd, "features")


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "fp16"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    torch.set_num_threads(os.cpu_count())
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g_sd, f_sd = fixtures.build_models(cache_dir=os.path.join(root, ".fixture_cache"))
    q = make_q(kind)
    w = torch.from_numpy(fixtures.latents(64)[32:32 + n])
    with torch.no_grad():
        wp = M.truncation(w, g_sd)
        ref_img = M.transform(M.postprocess(M.synthesis(wp, g_sd, literal=False)))
        ref_emb = M.iresnet50(ref_img, f_sd)
        img = M.transform(M.postprocess(synthesis_q(wp, g_sd, q)))
        print(f"[{kind}] img112 max|d| {(img - ref_img).abs().max():.2e} mean|d| {(img - ref_img).abs().mean():.2e}"
              f"  inter-identity mean|d| {(ref_img[0] - ref_img[1]).abs().mean():.2e}")
        for tag, im in (("gan-q + frm-fp32", img),):
            e = M.iresnet50(im, f_sd)
            print(tag, "cos", F.cosine_similarity(e, ref_emb).tolist())
        for fr in (False, True):
            e = iresnet50_q(ref_img, f_sd, q, fp32_residual=fr)
            print(f"gan-fp32 + frm-q(fp32_res={fr}) cos", F.cosine_similarity(e, ref_emb).tolist())
            e = iresnet50_q(img, f_sd, q, fp32_residual=fr)
            print(f"gan-q + frm-q(fp32_res={fr}) cos", F.cosine_similarity(e, ref_emb).tolist(),
                  "L2", (e - ref_emb).norm(dim=1).tolist(), "norm", ref_emb.norm(dim=1).tolist())


if __name__ == "__main__":
    main()
