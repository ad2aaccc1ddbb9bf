"""Per-kernel parity on the GPU, through the C ABI.  Floating-point kernels are compared against a plain torch
fp32 reference of the same op computed from the same fp16-rounded operands (tolerances stated per test)."""
import ctypes as C
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from certifyingfacerecognition_b200 import engine
    return engine


def _nhwc16(x):
    return x.permute(0, 2, 3, 1).contiguous().half()


def _from_nhwc(y, n, h, w, c):
    return y.view(n, h, w, c).permute(0, 3, 1, 2).float()


def _sync():
    torch.cuda.synchronize()


def _stats(n, c):
    """per-(n,c) sum / sum-of-squares accumulators: Q43.20 fixed point in int64 (integer atomics are reproducible)"""
    return torch.zeros(n, c, dtype=torch.int64, device="cuda"), torch.zeros(n, c, dtype=torch.int64, device="cuda")


def _fx(t):
    return t.double() / 2 ** 20


CONV_CASES = [
    # n, cin, cout, res, stride, ksize
    (2, 64, 64, 16, 1, 3),
    (4, 512, 512, 8, 1, 3),
    (8, 512, 512, 4, 1, 3),
    (2, 128, 256, 32, 1, 3),
    (1, 32, 32, 64, 1, 3),
    (1, 16, 16, 64, 1, 3),
    (2, 16, 64, 112, 1, 3),
    (2, 64, 64, 112, 2, 3),
    (2, 64, 128, 56, 2, 1),
    (8, 128, 128, 28, 1, 3),
    (32, 256, 256, 14, 1, 3),
    (3, 256, 512, 14, 2, 3),
    (5, 512, 512, 7, 1, 3),
    (2, 64, 320, 16, 1, 3),
    (297, 16, 32, 8, 1, 3),      # 149 M tiles (odd): paired-tile (MT=2) path with a dummy last tile
    (150, 64, 128, 16, 1, 3),    # 300 M tiles: paired-tile path, two accumulator sets in flight
    (160, 256, 256, 16, 1, 3),   # 160 tile pairs, Cout 256, K 2304: CTA-pair kernel (cta_group::2, N tile 256 over two SMs)
    (75, 256, 256, 14, 1, 3),    # CTA pairs, 115 M tiles (odd): the peer CTA's last tile is a dummy; tiles span images
    (40, 256, 512, 16, 1, 3),    # CTA pairs, two 256-wide N tiles
    (64, 256, 256, 28, 2, 3),    # CTA pairs, stride 2
    (3, 128, 128, 128, 1, 3),    # 128-wide grid: ROWS kernel (row-shared A operand), 3x3, two K chunks
    (2, 64, 256, 128, 1, 3),     # ROWS kernel, two N tiles
]


@pytest.mark.parametrize("n,cin,cout,res,stride,ks", CONV_CASES)
def test_conv_matches_torch(E, n, cin, cout, res, stride, ks):
    L = E.L
    g = torch.Generator().manual_seed(n * 1000 + cin + cout + res)
    x = torch.randn(n, cin, res, res, generator=g).cuda().half().float()
    w = (torch.randn(cout, cin, ks, ks, generator=g) / math.sqrt(cin * ks * ks)).cuda().half().float()
    bias = torch.randn(cout, generator=g).cuda()
    ores = (res + 2 * (ks // 2) - ks) // stride + 1
    ref = F.conv2d(x, w, bias, stride=stride, padding=ks // 2)
    prog = E.Program()
    out = torch.full((n * ores * ores * cout,), float("nan"), dtype=torch.float16, device="cuda")
    taps = [E.TAPS3] if ks == 3 else [[(0, 0)]]
    prog.conv(inp=_nhwc16(x), n=n, hin=res, win=res, cin=cin, w=E.pack_conv_weight(w.cpu()).cuda().half(), cout=cout,
              hout=ores, wout=ores, tile=E.tile_for(ores, n), out=out, out_hwc=(ores, ores, cout), taps=taps,
              stride=stride, bias=bias)
    prog.run()
    _sync()
    got = _from_nhwc(out, n, ores, ores, cout)
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), err      # fp16 output rounding: 2^-11 relative


def test_conv_epilogue_noise_lrelu_stats(E):
    n, c, res = 3, 64, 32
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, c, res, res, generator=g).cuda().half().float()
    w = (torch.randn(c, c, 3, 3, generator=g) / math.sqrt(c * 9)).cuda().half().float()
    bias, nw = torch.randn(c, generator=g).cuda(), torch.randn(c, generator=g).cuda()
    noise = torch.randn(res, res, generator=g).cuda()
    ref = F.leaky_relu(F.conv2d(x, w, padding=1) + noise.view(1, 1, res, res) * nw.view(1, -1, 1, 1)
                       + bias.view(1, -1, 1, 1), 0.2)
    out = torch.zeros(n * res * res * c, dtype=torch.float16, device="cuda")
    ssum, ssq = _stats(n, c)
    prog = E.Program()
    prog.conv(inp=_nhwc16(x), n=n, hin=res, win=res, cin=c, w=E.pack_conv_weight(w.cpu()).cuda().half(), cout=c,
              hout=res, wout=res, tile=E.tile_for(res), out=out, out_hwc=(res, res, c), taps=[E.TAPS3], bias=bias,
              noise=noise.reshape(-1).contiguous(), noise_w=nw, act=E.L.ACT_LRELU, slope=0.2, stat_sum=ssum, stat_sq=ssq)
    prog.run()
    _sync()
    got = _from_nhwc(out, n, res, res, c)
    assert (got - ref).abs().max().item() < 2e-2
    assert torch.allclose(_fx(ssum), ref.double().sum(dim=[2, 3]), rtol=1e-4, atol=1e-2)
    assert torch.allclose(_fx(ssq), (ref * ref).double().sum(dim=[2, 3]), rtol=1e-4, atol=1e-2)


def test_conv_128_wide_epilogue_noise_lrelu_stats(E, monkeypatch):
    """StyleGAN layer 11's shape (128-pixel-wide grid: the ROWS kernel) with the StyleGAN epilogue and fused statistics, on
    3 CTAs so that every CTA walks many row pairs / ring wraps; replayed for bit identity."""
    monkeypatch.setenv("CFR_MAX_CTAS", "3")
    n, cin, c, res = 2, 128, 128, 128
    g = torch.Generator().manual_seed(56)
    x = torch.randn(n, cin, res, res, generator=g).cuda().half().float()
    w = (torch.randn(c, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).cuda().half().float()
    bias, nw = torch.randn(c, generator=g).cuda(), torch.randn(c, generator=g).cuda()
    noise = torch.randn(res, res, generator=g).cuda()
    ref = F.leaky_relu(F.conv2d(x, w, padding=1) + noise.view(1, 1, res, res) * nw.view(1, -1, 1, 1)
                       + bias.view(1, -1, 1, 1), 0.2)
    outs = []
    for _ in range(2):
        out = torch.zeros(n * res * res * c, dtype=torch.float16, device="cuda")
        ssum, ssq = _stats(n, c)
        prog = E.Program()
        prog.conv(inp=_nhwc16(x), n=n, hin=res, win=res, cin=cin, w=E.pack_conv_weight(w.cpu()).cuda().half(), cout=c,
                  hout=res, wout=res, tile=E.tile_for(res), out=out, out_hwc=(res, res, c), taps=[E.TAPS3], bias=bias,
                  noise=noise.reshape(-1).contiguous(), noise_w=nw, act=E.L.ACT_LRELU, slope=0.2, stat_sum=ssum,
                  stat_sq=ssq)
        prog.run()
        _sync()
        outs.append((out, ssum.clone(), ssq.clone()))
    got = _from_nhwc(outs[0][0], n, res, res, c)
    assert (got - ref).abs().max().item() < 2e-2
    assert torch.allclose(_fx(outs[0][1]), ref.double().sum(dim=[2, 3]), rtol=1e-4, atol=5e-2)
    assert torch.allclose(_fx(outs[0][2]), (ref * ref).double().sum(dim=[2, 3]), rtol=1e-4, atol=5e-2)
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


def test_conv_cta_pair_epilogue_noise_lrelu_stats(E):
    """The CTA-pair kernel (Cout % 256 == 0, >= 37 tile pairs) with the StyleGAN epilogue: noise, bias, LeakyReLU and the
    fused per-(n,c) statistics, each CTA of a pair flushing the sums of its own M tile; replayed for determinism."""
    n, cin, c, res = 6, 256, 256, 64
    g = torch.Generator().manual_seed(55)
    x = torch.randn(n, cin, res, res, generator=g).cuda().half().float()
    w = (torch.randn(c, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).cuda().half().float()
    bias, nw = torch.randn(c, generator=g).cuda(), torch.randn(c, generator=g).cuda()
    noise = torch.randn(res, res, generator=g).cuda()
    ref = F.leaky_relu(F.conv2d(x, w, padding=1) + noise.view(1, 1, res, res) * nw.view(1, -1, 1, 1)
                       + bias.view(1, -1, 1, 1), 0.2)
    outs = []
    for _ in range(2):
        out = torch.zeros(n * res * res * c, dtype=torch.float16, device="cuda")
        ssum, ssq = _stats(n, c)
        prog = E.Program()
        prog.conv(inp=_nhwc16(x), n=n, hin=res, win=res, cin=cin, w=E.pack_conv_weight(w.cpu()).cuda().half(), cout=c,
                  hout=res, wout=res, tile=E.tile_for(res), out=out, out_hwc=(res, res, c), taps=[E.TAPS3], bias=bias,
                  noise=noise.reshape(-1).contiguous(), noise_w=nw, act=E.L.ACT_LRELU, slope=0.2, stat_sum=ssum,
                  stat_sq=ssq)
        assert "BN256" in prog.lib.cfr_program_op_label(prog.handle, 0).decode()
        prog.run()
        _sync()
        outs.append((out, ssum.clone(), ssq.clone()))
    got = _from_nhwc(outs[0][0], n, res, res, c)
    assert (got - ref).abs().max().item() < 2e-2
    assert torch.allclose(_fx(outs[0][1]), ref.double().sum(dim=[2, 3]), rtol=1e-4, atol=2e-2)
    assert torch.allclose(_fx(outs[0][2]), (ref * ref).double().sum(dim=[2, 3]), rtol=1e-4, atol=2e-2)
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


def test_conv_prelu_residual_classbias(E):
    """iresnet block conv1 form: pre-conv affine (scale folded, shift via border-class bias) + PReLU, and conv2
    form with residual add."""
    n, c, res = 2, 64, 28
    g = torch.Generator().manual_seed(6)
    x = torch.randn(n, c, res, res, generator=g).cuda().half().float()
    w = (torch.randn(c, c, 3, 3, generator=g) / math.sqrt(c * 9))
    s1, t1 = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    alpha = torch.rand(c, generator=g).cuda()
    resid = torch.randn(n, c, res, res, generator=g).cuda().half().float()
    wq = (w * s1.view(1, -1, 1, 1)).half().float()
    # reference: conv of (s1*x + t1) zero-padded AFTER the affine, using the same rounded scaled weights
    xin = x + (t1 / s1).cuda().view(1, -1, 1, 1)
    ref = F.conv2d(xin, wq.cuda(), padding=1)
    ref = torch.where(ref >= 0, ref, ref * alpha.view(1, -1, 1, 1)) + resid
    tb = torch.einsum("oikl,i->okl", wq, t1 / s1).reshape(c, 9)
    cb = torch.zeros(9, c)
    for rc in range(3):
        for cc in range(3):
            valid = [k for k, (dy, dx) in enumerate(E.TAPS3) if not (rc == 0 and dy < 0) and not (rc == 2 and dy > 0)
                     and not (cc == 0 and dx < 0) and not (cc == 2 and dx > 0)]
            cb[rc * 3 + cc] = tb[:, valid].sum(dim=1)
    out = torch.zeros(n * res * res * c, dtype=torch.float16, device="cuda")
    prog = E.Program()
    prog.conv(inp=_nhwc16(x), n=n, hin=res, win=res, cin=c, w=E.pack_conv_weight(wq).cuda().half(), cout=c, hout=res,
              wout=res, tile=E.tile_for(res, n), out=out, out_hwc=(res, res, c), taps=[E.TAPS3], cbias=cb.cuda(),
              act=E.L.ACT_PRELU, alpha=alpha, resid=_nhwc16(resid), resid_c=c)
    prog.run()
    _sync()
    got = _from_nhwc(out, n, res, res, c)
    assert (got - ref).abs().max().item() < 3e-2


@pytest.mark.parametrize("cin,cout,lo,fused", [(512, 512, 4, False), (64, 32, 16, True), (32, 16, 32, True),
                                                 (128, 64, 8, False),
                                                 # 128-wide grids (StyleGAN layer 12's shape): the ROWS kernel -- one box of
                                                 # input rows per K chunk serves all taps of two rows and both column phases
                                                 (128, 64, 128, False), (64, 128, 128, False)])
def test_upconv_phases_match_upsample_conv(E, cin, cout, lo, fused):
    n = 2
    g = torch.Generator().manual_seed(cin + lo)
    x = torch.randn(n, cin, lo, lo, generator=g).cuda().half().float()
    weq = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
    wp, taps = E.pack_upconv_phases(weq)
    wp = wp.half()
    # reference from the same rounded phase weights is awkward; compare against exact math with a tolerance
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), weq.cuda(), padding=1)
    res = 2 * lo
    out = torch.full((n * res * res * cout,), float("nan"), dtype=torch.float16, device="cuda")
    prog = E.Program()
    prog.conv(inp=_nhwc16(x), n=n, hin=lo, win=lo, cin=cin, w=wp.cuda(), cout=cout, hout=lo, wout=lo,
              tile=E.tile_for(lo), out=out, out_hwc=(res, res, cout), taps=taps, oscale=2,
              ooff=[(0, 0), (0, 1), (1, 0), (1, 1)], w_rows_per_phase=cout)
    prog.run()
    _sync()
    got = _from_nhwc(out, n, res, res, cout)
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() < 3e-2


def test_fc_as_conv(E):
    n, k, cout = 5, 49 * 512, 512
    g = torch.Generator().manual_seed(9)
    x = torch.randn(n, k, generator=g).cuda().half().float()
    w = (torch.randn(cout, k, generator=g) / math.sqrt(k)).cuda().half().float()
    bias = torch.randn(cout, generator=g).cuda()
    out = torch.zeros(n, cout, device="cuda")
    prog = E.Program()
    prog.conv(inp=x.half().contiguous(), n=n, hin=1, win=1, cin=k, w=w.half().contiguous(), cout=cout, hout=1, wout=1,
              tile=(1, 1, 128), out=out, out_hwc=(1, 1, cout), taps=[[(0, 0)]], bias=bias)
    prog.run()
    _sync()
    ref = x @ w.t() + bias
    assert (out - ref).abs().max().item() < 2e-3
    # split-K form used by ArcFaceProgram: 4 K-slices as conv phases + a fixed-order sum of the partials
    P, kp = 4, k // 4
    wsplit = w.reshape(cout, P, kp).permute(1, 0, 2).reshape(P * cout, kp).half().contiguous()
    partial = torch.full((n * P * cout,), float("nan"), device="cuda")
    out2 = torch.zeros(n, cout, device="cuda")
    p2 = E.Program()
    p2.conv(inp=x.half().contiguous(), n=n, hin=1, win=P, cin=kp, w=wsplit, cout=cout, hout=1, wout=1, tile=(1, 1, 128),
            out=partial, out_hwc=(1, P, cout), taps=[[(0, q)] for q in range(P)], ooff=[(0, q) for q in range(P)],
            w_rows_per_phase=cout)
    E.L.check(p2.lib.cfr_program_add_sum_partials(p2.handle, E.L.ptr(partial), E.L.ptr(bias), n, P, cout, E.L.ptr(out2)))
    p2.run()
    _sync()
    assert (out2 - ref).abs().max().item() < 2e-3
    assert (out2 - out).abs().max().item() < 1e-4          # same products, another summation order


@pytest.mark.parametrize("n,c,res", [(2, 32, 64), (1, 64, 128), (2, 128, 40), (1, 256, 16), (1, 512, 8), (1, 32, 256)])
def test_blur_act_stats(E, n, c, res):
    """BlurLayer + noise/bias/LeakyReLU + IN sums; C <= 256 takes the cp.async ring kernel (strips of 32 rows, so
    res = 40 / 128 / 256 cross strip boundaries), C = 512 the register sliding-window kernel."""
    L = E.L
    lib = L.load()
    g = torch.Generator().manual_seed(11)
    raw = torch.randn(n, c, res, res, generator=g).cuda().half().float()
    noise = torch.randn(res * res, generator=g).cuda()
    nw, bias = torch.randn(c, generator=g).cuda(), torch.randn(c, generator=g).cuda()
    k = torch.tensor([1.0, 2.0, 1.0])
    k = ((k[:, None] * k[None, :]) / 16).view(1, 1, 3, 3).repeat(c, 1, 1, 1).cuda()
    ref = F.conv2d(raw, k, padding=1, groups=c) + noise.view(1, 1, res, res) * nw.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    ref = F.leaky_relu(ref, 0.2)
    y = torch.zeros(n * res * res * c, dtype=torch.float16, device="cuda")
    ssum, ssq = _stats(n, c)
    raw_keep = _nhwc16(raw)
    prog2 = E.Program()
    L.check(lib.cfr_program_add_blur_act_stats(prog2.handle, L.ptr(raw_keep), L.ptr(y), n, res, res, c, L.ptr(noise),
                                               L.ptr(nw), L.ptr(bias), L.ptr(ssum), L.ptr(ssq), 0))
    prog2.run()
    _sync()
    got = _from_nhwc(y, n, res, res, c)
    assert (got - ref).abs().max().item() < 1e-2
    assert torch.allclose(_fx(ssum), ref.double().sum(dim=[2, 3]), rtol=1e-4, atol=1e-2)
    assert torch.allclose(_fx(ssq), (ref * ref).double().sum(dim=[2, 3]), rtol=1e-4, atol=1e-2)
    # statistics-only mode on the produced tensor
    s2, q2 = _stats(n, c)
    prog3 = E.Program()
    L.check(lib.cfr_program_add_blur_act_stats(prog3.handle, L.ptr(y), None, n, res, res, c, None, None, None,
                                               L.ptr(s2), L.ptr(q2), 1))
    prog3.run()
    _sync()
    assert torch.allclose(_fx(s2), got.double().sum(dim=[2, 3]), rtol=1e-4, atol=1e-2)


def test_finalize_and_affine(E):
    L = E.L
    lib = L.load()
    n, c, hw = 3, 64, 256
    g = torch.Generator().manual_seed(12)
    y = (torch.randn(n, hw, c, generator=g) * 2 + 1).cuda().half()
    yf = y.float()
    styles = torch.randn(n, 2 * c + 10, generator=g).cuda()
    ssum = (yf.double().sum(1) * 2 ** 20).round().long().contiguous()
    ssq = ((yf.double() * yf.double()).sum(1) * 2 ** 20).round().long().contiguous()
    A, B = torch.zeros(n, c, device="cuda"), torch.zeros(n, c, device="cuda")
    x = torch.zeros_like(y)
    prog = E.Program()
    L.check(lib.cfr_program_add_finalize_stats(prog.handle, L.ptr(ssum), L.ptr(ssq), L.ptr(styles), 2 * c + 10, 10, n, c,
                                               1.0 / hw, L.ptr(A), L.ptr(B)))
    L.check(lib.cfr_program_add_affine(prog.handle, L.ptr(y), L.ptr(A), L.ptr(B), n, hw, c, L.ptr(x)))
    prog.run()
    _sync()
    xc = yf - yf.mean(1, keepdim=True)
    xn = xc / torch.sqrt((xc * xc).mean(1, keepdim=True) + 1e-8)
    ref = xn * (styles[:, 10:10 + c].unsqueeze(1) + 1) + styles[:, 10 + c:10 + 2 * c].unsqueeze(1)
    assert (x.float() - ref).abs().max().item() < 2e-2


def test_torgb_resize_matches_torch(E):
    L = E.L
    lib = L.load()
    n, c, hin, rout = 2, 16, 256, 112
    g = torch.Generator().manual_seed(13)
    x = torch.randn(n, c, hin, hin, generator=g).cuda().half().float()
    wr, br = torch.randn(3, c, generator=g).cuda() * 0.3, torch.randn(3, generator=g).cuda() * 0.1
    img = F.conv2d(x, wr.view(3, c, 1, 1)) + br.view(1, -1, 1, 1)
    img = torch.clamp((img + 1) / 2 + 0.5 / 255, 0, 1)
    ref = (F.interpolate(img, size=(rout, rout), mode="bilinear", align_corners=False) - 0.5) / 0.5
    out = torch.zeros(n, rout, rout, 16, dtype=torch.float16, device="cuda")
    planar = torch.zeros(n, 3, rout, rout, device="cuda")
    xin = _nhwc16(x)
    prog = E.Program()
    L.check(lib.cfr_program_add_torgb_resize(prog.handle, L.ptr(xin), None, None, n, hin, c, L.ptr(wr.contiguous()),
                                             L.ptr(br), rout, 0.5, 0.5, L.ptr(out), L.ptr(planar), None))
    prog.run()
    _sync()
    assert (planar - ref).abs().max().item() < 1e-5          # fp32 math on identical fp16 inputs
    assert (out[..., :3].permute(0, 3, 1, 2).float() - ref).abs().max().item() < 1e-3
    assert (out[..., 3:] == 0).all()


def test_noise_project_and_truncate(E):
    L = E.L
    lib = L.load()
    g = torch.Generator().manual_seed(14)
    b = 7
    z, x = torch.randn(512, generator=g).cuda(), torch.randn(5, generator=g).cuda() * 0.1
    dirs = torch.randn(5, 512, generator=g).cuda()
    w_avg = torch.randn(512, generator=g).cuda() * 0.1
    noise = torch.randn(b, 5, generator=g).cuda() * 0.3
    wp2 = torch.zeros(b, 2, 512, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(lib.cfr_noise_project(L.ptr(z), L.ptr(x), L.ptr(torch.ones(1, device="cuda")), 1, L.ptr(noise), L.ptr(dirs),
                                  L.ptr(w_avg), 0.7, 0, 0, b, None, L.ptr(wp2), st))
    _sync()
    w = z.view(1, -1) + (x.view(1, 5) + noise) @ dirs
    assert torch.allclose(wp2[:, 1], w_avg + (w - w_avg) * 1.0, atol=1e-5)
    assert torch.allclose(wp2[:, 0], w_avg + (w - w_avg) * 0.7, atol=1e-5)
    # Philox path: deterministic in (seed, offset), N(0, sigma^2) per direction
    nb = 4096
    sig = torch.tensor([0.25, 0.25, 0.04, 0.25, 0.64], device="cuda")
    n1, n2 = torch.zeros(nb, 5, device="cuda"), torch.zeros(nb, 5, device="cuda")
    wp = torch.zeros(nb, 2, 512, device="cuda")
    zero5 = torch.zeros(5, device="cuda")
    L.check(lib.cfr_noise_project(L.ptr(z), L.ptr(zero5), L.ptr(sig), 5, None, L.ptr(dirs), L.ptr(w_avg), 0.7, 1234, 0, nb,
                                  L.ptr(n1), L.ptr(wp), st))
    L.check(lib.cfr_noise_project(L.ptr(z), L.ptr(zero5), L.ptr(sig), 5, None, L.ptr(dirs), L.ptr(w_avg), 0.7, 1234, 100,
                                  nb - 100, L.ptr(n2), L.ptr(wp), st))
    _sync()
    assert torch.equal(n1[100:], n2[:nb - 100])               # counter = global sample index
    assert torch.allclose(n1.std(0), sig, rtol=0.06)
    assert n1.mean(0).abs().max().item() < 0.05
    assert abs(float(((n1 / sig) ** 4).mean()) - 3.0) < 0.3   # Gaussian kurtosis


def test_match_vote_bit_exact(E):
    L = E.L
    lib = L.load()
    g = torch.Generator().manual_seed(15)
    b, n = 37, 5000
    gal = torch.randn(n, 512, generator=g).cuda()
    idx = torch.randint(0, n, (b,), generator=g)
    emb = gal[idx.cuda()] + 0.3 * torch.randn(b, 512, generator=g).cuda()
    gal[77] = gal[4000]                                       # exact duplicate rows: first index must win
    emb[0] = gal[4000]
    keys = torch.full((b,), -1, dtype=torch.int64, device="cuda")
    pred = torch.zeros(b, dtype=torch.int32, device="cuda")
    counts = torch.zeros(n, dtype=torch.int64, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        L.check(lib.cfr_match_vote(L.ptr(emb), b, L.ptr(gal), n, L.ptr(keys), L.ptr(pred), L.ptr(counts), st))
    _sync()
    d = torch.cdist(emb, gal, compute_mode="donot_use_mm_for_euclid_dist")
    ref = F.softmax(-d / np.sqrt(512), dim=1).argmax(1)
    assert pred[0].item() == 77
    assert torch.equal(pred.long(), ref)
    ref_counts = torch.bincount(ref, minlength=n) * 2
    assert torch.equal(counts, ref_counts)


# ------------------------------------------------------------------------------------------------------------
# halo-resident conv kernel (high-resolution StyleGAN layers)
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,cin,cout,h,w,affine,fold", [
    (2, 16, 16, 40, 256, True, False), (1, 32, 32, 24, 128, True, False), (2, 64, 64, 16, 200, True, False),
    (3, 16, 16, 8, 64, False, False), (1, 64, 64, 9, 130, False, False),
    (3, 16, 16, 40, 256, True, True), (2, 32, 32, 24, 200, True, True), (2, 16, 16, 8, 64, False, True),
    (2, 64, 64, 16, 200, True, True), (3, 64, 64, 9, 130, True, True)])
def test_halo_conv_matches_torch(E, n, cin, cout, h, w, affine, fold):
    g = torch.Generator().manual_seed(cin * 7 + h)
    yprev = torch.randn(n, cin, h, w, generator=g).cuda().half().float()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).cuda().half().float()
    bias, nw = torch.randn(cout, generator=g).cuda(), torch.randn(cout, generator=g).cuda()
    noise = torch.randn(h, w, generator=g).cuda()
    A = (torch.rand(n, cin, generator=g) + 0.5).cuda()
    B = torch.randn(n, cin, generator=g).cuda()
    xin = (yprev * A.view(n, cin, 1, 1) + B.view(n, cin, 1, 1)).half().float() if affine else yprev
    ref = F.leaky_relu(F.conv2d(xin, wt, padding=1) + noise.view(1, 1, h, w) * nw.view(1, -1, 1, 1)
                       + bias.view(1, -1, 1, 1), 0.2)
    out = torch.full((n * h * w * cout,), float("nan"), dtype=torch.float16, device="cuda")
    ssum, ssq = _stats(n, cout)
    prog = E.Program()
    wpk = E.pack_halo_weight(wt.cpu()).cuda()
    prog.conv(inp=_nhwc16(yprev), n=n, hin=h, win=w, cin=cin, w=wpk.float().contiguous() if fold else wpk.half(), cout=cout,
              hout=h, wout=w, tile=(16, 8, 1), out=out, out_hwc=(h, w, cout), taps=[E.TAPS3], bias=bias,
              noise=noise.reshape(-1).contiguous(), noise_w=nw, act=E.L.ACT_LRELU, slope=0.2, stat_sum=ssum, stat_sq=ssq,
              halo=True, in_affine=(A.contiguous(), B.contiguous()) if affine else None,
              fold_center_tap=4 if fold else None)
    prog.run()
    prog.run()          # replays must be idempotent for the output (stats accumulate)
    _sync()
    got = _from_nhwc(out, n, h, w, cout)
    assert torch.isfinite(got).all()
    # folded variant: scale rounded into fp16 weights, shift / bias / noise through fp16 aux operands -> looser bound
    tol = (4e-2 if fold else 2e-2) * max(1.0, ref.abs().max().item())
    assert (got - ref).abs().max().item() < tol
    assert (got - ref).abs().mean().item() < 2e-3 * max(1.0, ref.abs().max().item())
    _check_stats(ssum, ssq, got, ref, runs=2, per_pixel=1e-3 if fold else 0.0)


@pytest.mark.parametrize("n,cin,cout,res,rout,ctas", [(2, 16, 16, 256, 28, None), (3, 16, 16, 128, 16, "3"),
                                                      (1, 32, 16, 256, 40, None)])
def test_halo_sparse_store_and_compact_resize(E, monkeypatch, n, cin, cout, res, rout, ctas):
    """Sparse store (cfr_conv_desc.keepMap): only the rows / columns the bilinear resize reads are written, into a compact
    buffer, bit-identical to the dense kernel's values at those pixels; the statistics are those of the dense run; the
    resize reading the compact buffer equals the resize reading the dense one (the last StyleGAN layer + toRGB path)."""
    if ctas:
        monkeypatch.setenv("CFR_MAX_CTAS", ctas)
    L = E.L
    g = torch.Generator().manual_seed(res + rout)
    yprev = torch.randn(n, cin, res, res, generator=g).cuda().half().float()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).cuda().half().float()
    bias, nw = torch.randn(cout, generator=g).cuda(), torch.randn(cout, generator=g).cuda()
    noise = torch.randn(res, res, generator=g).cuda()
    A, B = (torch.rand(n, cin, generator=g) + 0.5).cuda(), torch.randn(n, cin, generator=g).cuda()
    kmap, kdim = E.resize_keep_map(res, rout)
    assert kdim == 2 * rout and int((kmap >= 0).sum()) == kdim
    kmap_d = kmap.cuda()
    wpk = E.pack_halo_weight(wt.cpu()).cuda().float().contiguous()
    outs, stats = [], []
    for sparse in (False, True):
        out = torch.full((n * (kdim * kdim if sparse else res * res) * cout + 64,), float("nan"), dtype=torch.float16,
                         device="cuda")
        ssum, ssq = _stats(n, cout)
        prog = E.Program()
        prog.conv(inp=_nhwc16(yprev), n=n, hin=res, win=res, cin=cin, w=wpk, cout=cout, hout=res, wout=res, tile=(16, 8, 1),
                  out=out, out_hwc=(res, res, cout), taps=[E.TAPS3], bias=bias, noise=noise.reshape(-1).contiguous(),
                  noise_w=nw, act=L.ACT_LRELU, slope=0.2, stat_sum=ssum, stat_sq=ssq, halo=True,
                  in_affine=(A.contiguous(), B.contiguous()), fold_center_tap=4,
                  keep_map=kmap_d if sparse else None, keep_dim=kdim if sparse else 0)
        prog.run()
        _sync()
        outs.append(out)
        stats.append((ssum.clone(), ssq.clone()))
    dense = outs[0][:n * res * res * cout].view(n, res, res, cout)
    compact = outs[1][:n * kdim * kdim * cout].view(n, kdim, kdim, cout)
    assert torch.isnan(outs[1][n * kdim * kdim * cout:]).all()           # nothing written past the compact buffer
    rows = torch.nonzero(kmap_d >= 0).flatten()
    assert torch.equal(compact, dense[:, rows][:, :, rows])              # same values, bit for bit
    assert torch.equal(stats[0][0], stats[1][0]) and torch.equal(stats[0][1], stats[1][1])
    if cout != 16:
        return
    # toRGB + resize from the compact buffer == from the dense buffer
    lib = L.load()
    wr, br = torch.randn(3, cout, generator=g).cuda() * 0.3, torch.randn(3, generator=g).cuda() * 0.1
    A2, B2 = (torch.rand(n, cout, generator=g) + 0.5).cuda(), torch.randn(n, cout, generator=g).cuda() * 0.1
    res_out = []
    for sparse in (False, True):
        planar = torch.zeros(n, 3, rout, rout, device="cuda")
        img = torch.zeros(n, rout, rout, 16, dtype=torch.float16, device="cuda")
        prog = E.Program()
        if sparse:
            L.check(lib.cfr_program_add_torgb_resize_sparse(prog.handle, L.ptr(outs[1]), L.ptr(A2), L.ptr(B2), n, res, cout,
                                                            L.ptr(wr.contiguous()), L.ptr(br), rout, 0.5, 0.5, L.ptr(img),
                                                            L.ptr(planar), None, L.ptr(kmap_d), kdim))
        else:
            L.check(lib.cfr_program_add_torgb_resize(prog.handle, L.ptr(outs[0]), L.ptr(A2), L.ptr(B2), n, res, cout,
                                                     L.ptr(wr.contiguous()), L.ptr(br), rout, 0.5, 0.5, L.ptr(img),
                                                     L.ptr(planar), None))
        prog.run()
        _sync()
        res_out.append((planar, img))
    assert torch.equal(res_out[0][0], res_out[1][0]) and torch.equal(res_out[0][1], res_out[1][1])


def _check_stats(ssum, ssq, got, ref, runs, per_pixel):
    """Fused per-(n,c) sum / sum of squares.  (1) Against the kernel's OWN fp16 output, tightly: the statistics are taken
    on the fp32 values before the store, so the only difference is the fp16 output rounding (random sign) -- a pixel
    counted twice, dropped or attributed to the wrong image shows up here.  (2) Against the fp32 reference, allowing the
    systematic per-pixel offset of the variant (FOLD: bias / shift / noise gain live in fp16 aux weights, |d| <= ~1e-3)."""
    hw = got.shape[2] * got.shape[3]
    own, own2 = runs * got.double().sum(dim=[2, 3]), runs * (got.double() ** 2).sum(dim=[2, 3])
    mag, mag2 = runs * got.double().abs().sum(dim=[2, 3]), own2
    d1, d2 = (_fx(ssum) - own).abs(), (_fx(ssq) - own2).abs()
    assert (d1 <= 0.05 + 1e-4 * mag).all(), f"sum vs own output: max |d| {d1.max().item():.4f}"
    assert (d2 <= 0.05 + 4e-4 * mag2).all(), f"sumsq vs own output: max |d| {d2.max().item():.4f}"
    want, want2 = runs * ref.double().sum(dim=[2, 3]), runs * (ref * ref).double().sum(dim=[2, 3])
    atol = per_pixel * runs * hw + 5e-2
    e1, e2 = (_fx(ssum) - want).abs(), (_fx(ssq) - want2).abs()
    assert (e1 <= atol + 2e-3 * want.abs()).all(), f"sum: max |d| {e1.max().item():.4f} (atol {atol})"
    assert (e2 <= 4 * atol + 2e-3 * want2.abs()).all(), f"sumsq: max |d| {e2.max().item():.4f} (atol {4 * atol})"


@pytest.mark.parametrize("n,cin,cout,lo_h,lo_w,fold", [(2, 64, 32, 12, 128, False), (1, 32, 16, 20, 256, False),
                                                        (2, 64, 32, 5, 96, False), (3, 32, 16, 20, 256, True),
                                                        (2, 64, 32, 12, 128, True)])
def test_halo_upconv_matches_torch(E, n, cin, cout, lo_h, lo_w, fold):
    g = torch.Generator().manual_seed(cin + lo_h)
    yprev = torch.randn(n, cin, lo_h, lo_w, generator=g).cuda().half().float()
    weq = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
    A = (torch.rand(n, cin, generator=g) + 0.5).cuda()
    B = torch.randn(n, cin, generator=g).cuda()
    xin = (yprev * A.view(n, cin, 1, 1) + B.view(n, cin, 1, 1)).half().float()
    ref = F.conv2d(F.interpolate(xin, scale_factor=2, mode="nearest"), weq.cuda(), padding=1)
    wp, taps = E.pack_halo_upconv(weq)
    H, W = 2 * lo_h, 2 * lo_w
    out = torch.full((n * H * W * cout,), float("nan"), dtype=torch.float16, device="cuda")
    prog = E.Program()
    prog.conv(inp=_nhwc16(yprev), n=n, hin=lo_h, win=lo_w, cin=cin, w=wp.cuda().float().contiguous() if fold else wp.cuda().half(),
              cout=cout, hout=lo_h, wout=lo_w, tile=(16, 8, 1), out=out, out_hwc=(H, W, cout), taps=taps, oscale=2,
              ooff=[(0, 0), (0, 1), (1, 0), (1, 1)], halo=True, in_affine=(A.contiguous(), B.contiguous()),
              fold_center_tap=-1 if fold else None)
    prog.run()
    _sync()
    got = _from_nhwc(out, n, H, W, cout)
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() < 3e-2 * max(1.0, ref.abs().max().item())


# ------------------------------------------------------------------------------------------------------------
# tensor-core gallery matcher (BASELINE config 5)
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_gallery,b", [(20000, 37), (1000, 128), (70001, 250)])
def test_tc_matcher_equals_exact_match(E, n_gallery, b):
    L = E.L
    lib = L.load()
    g = torch.Generator().manual_seed(n_gallery)
    gal = (torch.randn(n_gallery, 512, generator=g) * 1.3 + 0.2).cuda()
    idx = torch.randint(0, n_gallery, (b,), generator=g).cuda()
    emb = (gal[idx] + 0.5 * torch.randn(b, 512, generator=g).cuda()).contiguous()
    gal[123] = gal[n_gallery - 7]                               # exact duplicate rows: the first index must win
    emb[0] = gal[n_gallery - 7]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    m = C.c_void_p()
    L.check(lib.cfr_matcher_create(L.ptr(gal), n_gallery, 256, st, C.byref(m)))
    pred = torch.zeros(b, dtype=torch.int32, device="cuda")
    counts = torch.zeros(n_gallery, dtype=torch.int64, device="cuda")
    for _ in range(2):
        L.check(lib.cfr_matcher_run(m, L.ptr(emb), b, L.ptr(pred), L.ptr(counts), st))
    _sync()
    lib.cfr_matcher_destroy(m)
    ref = torch.cdist(emb, gal, compute_mode="donot_use_mm_for_euclid_dist").argmin(1)
    assert pred[0].item() == 123
    assert torch.equal(pred.long(), ref)
    assert torch.equal(counts, torch.bincount(ref, minlength=n_gallery) * 2)


@pytest.mark.parametrize("n,cin,cout,lo_h,lo_w", [(2, 32, 16, 12, 128), (3, 32, 16, 9, 200), (1, 16, 16, 4, 64)])
def test_halo_upconv_blur_composite_matches_torch(E, n, cin, cout, lo_h, lo_w):
    """UpConvBlock incl. BlurLayer and epilogue (up x2 -> conv3x3 -> blur -> +noise*w +bias -> LeakyReLU) as ONE kernel,
    with the previous layer's IN/AdaIN folded in; borders must be exact."""
    g = torch.Generator().manual_seed(cin + lo_h)
    yprev = torch.randn(n, cin, lo_h, lo_w, generator=g).cuda().half().float()
    weq = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
    A = (torch.rand(n, cin, generator=g) + 0.5).cuda()
    B = torch.randn(n, cin, generator=g).cuda()
    H, W = 2 * lo_h, 2 * lo_w
    bias, nw = torch.randn(cout, generator=g).cuda(), torch.randn(cout, generator=g).cuda()
    noise = torch.randn(H, W, generator=g).cuda()
    xin = yprev * A.view(n, cin, 1, 1) + B.view(n, cin, 1, 1)
    raw = F.conv2d(F.interpolate(xin, scale_factor=2, mode="nearest"), weq.cuda(), padding=1)
    k = torch.tensor([1.0, 2.0, 1.0])
    kk = ((k[:, None] * k[None, :]) / 16).view(1, 1, 3, 3).repeat(cout, 1, 1, 1).cuda()
    ref = F.leaky_relu(F.conv2d(raw, kk, padding=1, groups=cout) + noise.view(1, 1, H, W) * nw.view(1, -1, 1, 1)
                       + bias.view(1, -1, 1, 1), 0.2)
    base, corr_d = E.composite_upconv_weights(weq)
    out = torch.full((n * H * W * cout,), float("nan"), dtype=torch.float16, device="cuda")
    ssum, ssq = _stats(n, cout)
    prog = E.Program()
    prog.conv(inp=_nhwc16(yprev), n=n, hin=lo_h, win=lo_w, cin=cin, w=base.cuda(), cout=cout, hout=lo_h, wout=lo_w,
              tile=(16, 8, 1), out=out, out_hwc=(H, W, cout), taps=[E.TAPS3] * 4, oscale=2,
              ooff=[(0, 0), (0, 1), (1, 0), (1, 1)], noise=noise.reshape(-1).contiguous(), noise_w=nw, bias=bias,
              act=E.L.ACT_LRELU, slope=0.2, stat_sum=ssum, stat_sq=ssq, halo=True,
              in_affine=(A.contiguous(), B.contiguous()), fold_center_tap=4, composite_corr=corr_d.cuda())
    prog.run()
    _sync()
    got = _from_nhwc(out, n, H, W, cout)
    assert torch.isfinite(got).all()
    err = (got - ref).abs()
    scale = max(1.0, ref.abs().max().item())
    assert err.max().item() < 4e-2 * scale
    # borders are as accurate as the interior
    border = torch.cat([err[:, :, 0].flatten(), err[:, :, -1].flatten(), err[:, :, :, 0].flatten(), err[:, :, :, -1].flatten()])
    assert border.max().item() < 4e-2 * scale and border.mean().item() < 3e-3 * scale
    assert err.mean().item() < 2e-3 * scale
    _check_stats(ssum, ssq, got, ref, runs=1, per_pixel=1e-3)


# ------------------------------------------------------------------------------------------------------------
# the same kernels with only a few CTAs: every CTA then walks many bands / tile pairs, which is what the full-size
# step does (accumulator ring wrap, mbarrier phase flips, per-sample weight reloads, statistics flushed between
# images, partial accumulator groups) but what a small test on 148 SMs never reaches
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ctas", [1, 3])
@pytest.mark.parametrize("n,cin,cout,h,w,affine,fold", [
    (3, 16, 16, 41, 256, True, True),      # row-stationary per group: bands of 12 rows, last one 5 = groups of 4 + 1
    (2, 16, 16, 10, 300, True, True),      # groups of 4, 4, 2; ragged x band
    (2, 16, 16, 7, 128, False, True),      # groups of 4 + 3
    (3, 32, 32, 24, 200, True, True),      # row-stationary per band: bands of 7, 7, 7, 3 rows, ring wrap inside a band
    (2, 32, 32, 15, 128, True, True),      # 7, 7, 1
    (2, 64, 64, 16, 200, True, False), (2, 32, 32, 24, 128, True, False), (3, 16, 16, 40, 256, True, False),
    (3, 64, 64, 16, 200, True, True)])
def test_halo_conv_few_ctas(E, monkeypatch, ctas, n, cin, cout, h, w, affine, fold):
    monkeypatch.setenv("CFR_MAX_CTAS", str(ctas))
    test_halo_conv_matches_torch(E, n, cin, cout, h, w, affine, fold)


@pytest.mark.parametrize("n,cin,cout,lo_h,lo_w", [(3, 32, 16, 9, 200), (2, 16, 16, 12, 128)])
def test_halo_upconv_blur_composite_few_ctas(E, monkeypatch, n, cin, cout, lo_h, lo_w):
    monkeypatch.setenv("CFR_MAX_CTAS", "2")
    test_halo_upconv_blur_composite_matches_torch(E, n, cin, cout, lo_h, lo_w)


@pytest.mark.parametrize("n,cin,cout,lo_h,lo_w,fold", [(2, 64, 32, 12, 128, False), (2, 32, 16, 20, 256, True)])
def test_halo_upconv_few_ctas(E, monkeypatch, n, cin, cout, lo_h, lo_w, fold):
    monkeypatch.setenv("CFR_MAX_CTAS", "2")
    test_halo_upconv_matches_torch(E, n, cin, cout, lo_h, lo_w, fold)


@pytest.mark.parametrize("n,cin,cout,res,stride,ks", [(8, 128, 128, 28, 1, 3), (297, 16, 32, 8, 1, 3), (75, 256, 256, 14, 1, 3),
                                                       (150, 64, 128, 16, 1, 3), (160, 128, 256, 16, 1, 3),
                                                       (3, 256, 512, 14, 2, 3)])
def test_conv_few_ctas(E, monkeypatch, n, cin, cout, res, stride, ks):
    monkeypatch.setenv("CFR_MAX_CTAS", "3")
    test_conv_matches_torch(E, n, cin, cout, res, stride, ks)


# ---- split-precision (fp16 hi/lo operand pairs) early StyleGAN layers -----------------------------------------------
def _split3_nhwc(x):
    """[n,c,h,w] fp32 -> NHWC fp16 [n,h,w,3c] = [hi | lo | hi]"""
    v = x.permute(0, 2, 3, 1).contiguous()
    hi = v.half()
    lo = (v - hi.float()).half()
    return torch.cat([hi, lo, hi], dim=3).contiguous()


@pytest.mark.parametrize("n,cin,cout,res", [(3, 512, 512, 4), (2, 512, 512, 16), (2, 256, 256, 64), (5, 64, 128, 32)])
def test_split_precision_conv_is_fp32_class(E, n, cin, cout, res):
    """kSplit == 3: un-rounded fp32 operands, x.w = x_hi.w_hi + x_lo.w_hi + x_hi.w_lo in the fp32 accumulator.  Against an
    fp64 torch conv the error must be >= 15x below what fp16-rounded operands give (2^-11 per operand); what is left
    (~1e-5 at K = 4608) is the tensor core's own fp32 accumulation."""
    g = torch.Generator().manual_seed(res * 7 + cin)
    x = torch.randn(n, cin, res, res, generator=g).cuda()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=1)
    out = torch.full((n * res * res * cout,), float("nan"), dtype=torch.float32, device="cuda")
    prog = E.Program()
    prog.conv(inp=_split3_nhwc(x), n=n, hin=res, win=res, cin=3 * cin, w=E.split3_weight(E.pack_conv_weight(w.cpu()), 9, cin).cuda(),
              cout=cout, hout=res, wout=res, tile=E.tile_for(res), out=out, out_hwc=(res, res, cout), taps=[E.TAPS3],
              bias=bias, k_split=3)
    prog.run()
    _sync()
    got = out.view(n, res, res, cout).permute(0, 3, 1, 2).double()
    rel = ((got - ref).norm() / ref.norm()).item()
    rel16 = ((F.conv2d(x.half().double(), w.half().double(), bias.double(), padding=1) - ref).norm() / ref.norm()).item()
    assert rel < 2e-5 and rel < rel16 / 15, (rel, rel16)
    # FLOPs are accounted on the logical Cin, not on the 3x wider operand
    assert prog.lib.cfr_program_op_flops(prog.handle, 0) == pytest.approx(2.0 * n * res * res * 9 * cin * cout)


def test_split_precision_upconv_blur_affine_chain(E):
    """One split-precision up-conv layer end to end: 4-phase conv (fp32 out) -> blur/noise/bias/LeakyReLU/statistics in
    fp32 -> IN + AdaIN coefficients -> x = y*A + B written as [hi|lo|hi]; compared with torch fp64."""
    n, cin, cout, lo = 2, 64, 128, 16
    res = 2 * lo
    g = torch.Generator().manual_seed(31)
    x = torch.randn(n, cin, lo, lo, generator=g).cuda()
    weq = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9))
    noise = torch.randn(res, res, generator=g).cuda()
    nw, bias = torch.randn(cout, generator=g).cuda(), torch.randn(cout, generator=g).cuda()
    styles = torch.randn(n, 2 * cout, generator=g).cuda()
    k = torch.tensor([1.0, 2.0, 1.0], dtype=torch.float64)
    kb = ((k[:, None] * k[None, :]) / 16).view(1, 1, 3, 3).repeat(cout, 1, 1, 1).cuda()
    raw_ref = F.conv2d(F.interpolate(x.double(), scale_factor=2, mode="nearest"), weq.cuda().double(), padding=1)
    y_ref = F.leaky_relu(F.conv2d(raw_ref, kb, padding=1, groups=cout) + noise.double().view(1, 1, res, res)
                         * nw.double().view(1, -1, 1, 1) + bias.double().view(1, -1, 1, 1), 0.2)
    mean = y_ref.mean(dim=[2, 3], keepdim=True)
    xhat = (y_ref - mean) / torch.sqrt(((y_ref - mean) ** 2).mean(dim=[2, 3], keepdim=True) + 1e-8)
    x_ref = xhat * (styles[:, :cout].double().view(n, cout, 1, 1) + 1) + styles[:, cout:].double().view(n, cout, 1, 1)

    wp, taps = E.pack_upconv_phases(weq)
    raw = torch.zeros(n * res * res * cout, device="cuda")
    y = torch.zeros(n * res * res * cout, device="cuda")
    xs = torch.zeros(n * res * res * 3 * cout, dtype=torch.float16, device="cuda")
    ssum, ssq = _stats(n, cout)
    A, B = torch.zeros(n * cout, device="cuda"), torch.zeros(n * cout, device="cuda")
    prog = E.Program()
    L = E.L
    prog.conv(inp=_split3_nhwc(x), n=n, hin=lo, win=lo, cin=3 * cin, w=E.split3_weight(wp, 4, cin).cuda(), cout=cout,
              hout=lo, wout=lo, tile=E.tile_for(lo), out=raw, out_hwc=(res, res, cout), taps=taps, oscale=2,
              ooff=[(0, 0), (0, 1), (1, 0), (1, 1)], w_rows_per_phase=cout, k_split=3)
    L.check(prog.lib.cfr_program_add_blur_act_stats_f32(prog.handle, L.ptr(raw), L.ptr(y), n, res, res, cout,
                                                        L.ptr(noise.reshape(-1).contiguous()), L.ptr(nw), L.ptr(bias),
                                                        L.ptr(ssum), L.ptr(ssq), 0))
    L.check(prog.lib.cfr_program_add_finalize_stats(prog.handle, L.ptr(ssum), L.ptr(ssq), L.ptr(styles), 2 * cout, 0, n,
                                                    cout, 1.0 / (res * res), L.ptr(A), L.ptr(B)))
    L.check(prog.lib.cfr_program_add_affine_f32(prog.handle, L.ptr(y), L.ptr(A), L.ptr(B), n, res * res, cout,
                                                L.ptr(xs), 3))
    prog.keep += [noise, nw, bias, styles, raw, y, xs, ssum, ssq, A, B]
    prog.run()
    _sync()
    y_got = y.view(n, res, res, cout).permute(0, 3, 1, 2).double()
    assert ((y_got - y_ref).norm() / y_ref.norm()).item() < 2e-5
    v = xs.view(n, res, res, 3, cout).double()
    assert torch.equal(v[:, :, :, 0], v[:, :, :, 2])                       # hi stored twice
    x_got = (v[:, :, :, 0] + v[:, :, :, 1]).permute(0, 3, 1, 2)
    assert ((x_got - x_ref).norm() / x_ref.norm()).item() < 5e-5
    # plain fp16 output (the hand-over to the first ordinary layer)
    xp = torch.zeros(n * res * res * cout, dtype=torch.float16, device="cuda")
    p2 = E.Program()
    L.check(p2.lib.cfr_program_add_affine_f32(p2.handle, L.ptr(y), L.ptr(A), L.ptr(B), n, res * res, cout, L.ptr(xp), 1))
    p2.run()
    _sync()
    assert torch.equal(xp.view(n, res, res, cout), xs.view(n, res, res, 3, cout)[:, :, :, 0])


def test_layer0_split_equals_plain_layer0(E):
    n = 3
    g = torch.Generator().manual_seed(8)
    xhat0 = torch.randn(16, 512, generator=g).cuda()
    styles = torch.randn(n, 1024, generator=g).cuda()
    plain = torch.zeros(n * 16 * 512, dtype=torch.float16, device="cuda")
    split = torch.zeros(n * 16 * 1536, dtype=torch.float16, device="cuda")
    L = E.L
    prog = E.Program()
    L.check(prog.lib.cfr_program_add_layer0(prog.handle, L.ptr(xhat0), L.ptr(styles), 1024, 0, n, L.ptr(plain)))
    L.check(prog.lib.cfr_program_add_layer0_split(prog.handle, L.ptr(xhat0), L.ptr(styles), 1024, 0, n, L.ptr(split)))
    prog.run()
    _sync()
    v = split.view(n, 16, 3, 512)
    assert torch.equal(v[:, :, 0], plain.view(n, 16, 512)) and torch.equal(v[:, :, 0], v[:, :, 2])
    ref = xhat0.view(1, 16, 512).double() * (styles[:, :512].double().view(n, 1, 512) + 1) + styles[:, 512:].double().view(n, 1, 512)
    assert ((v[:, :, 0].double() + v[:, :, 1].double() - ref).abs().max() / ref.abs().max()).item() < 1e-6


# ---- stand-ins for compute-sanitizer (closed on this GPU pool, profiles/sanitizer_r02.md) ---------------------------
class _Guarded:
    """Tensors carved out of larger allocations whose guard bands (1 KiB before and after) hold a sentinel."""
    GUARD = 1024

    def __init__(self):
        self.blocks = []

    def make(self, numel, dtype, fill=0):
        es = torch.empty((), dtype=dtype).element_size()
        g = self.GUARD // es
        base = torch.empty(numel + 2 * g, dtype=dtype, device="cuda")
        raw = base.view(torch.uint8)
        raw.fill_(0xA5)
        view = base[g:g + numel]
        view.fill_(fill)
        self.blocks.append((base, g, numel))
        return view

    def check(self):
        for base, g, numel in self.blocks:
            raw = base.view(torch.uint8)
            es = base.element_size()
            assert bool((raw[:g * es] == 0xA5).all()) and bool((raw[(g + numel) * es:] == 0xA5).all()), \
                f"guard band of a {base.dtype} buffer of {numel} elements was written"


def test_kernels_do_not_write_outside_their_buffers(E, monkeypatch):
    """Every output / statistics / per-sample weight buffer sits between sentinel guard bands; after the launches (full grid
    and 3-CTA grid: ring wraps, dummy tiles, partial tiles at image borders) the bands must be untouched."""
    L = E.L
    for ctas in (None, "3"):
        if ctas:
            monkeypatch.setenv("CFR_MAX_CTAS", ctas)
        G = _Guarded()
        g = torch.Generator().manual_seed(17)
        prog = E.Program()
        # (1) igemm: odd tile count (paired-tile path with a dummy tile), fused stats, partial tiles (res 20 on 16x8 boxes)
        n, cin, cout, res = 3, 64, 64, 20
        x = torch.randn(n, cin, res, res, generator=g).cuda()
        w = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
        out = G.make(n * res * res * cout, torch.float16)
        ssum, ssq = G.make(n * cout, torch.int64), G.make(n * cout, torch.int64)
        prog.conv(inp=_nhwc16(x), n=n, hin=res, win=res, cin=cin, w=E.pack_conv_weight(w).cuda().half(), cout=cout,
                  hout=res, wout=res, tile=(16, 8, 1), out=out, out_hwc=(res, res, cout), taps=[E.TAPS3],
                  stat_sum=ssum, stat_sq=ssq)
        # (2) split-precision 4-phase up-conv, fp32 output, then the fp32 blur / affine chain
        lo, ci2, co2 = 6, 64, 64
        x2 = torch.randn(n, ci2, lo, lo, generator=g).cuda()
        wp, taps = E.pack_upconv_phases(torch.randn(co2, ci2, 3, 3, generator=g) / math.sqrt(ci2 * 9))
        raw = G.make(n * 4 * lo * lo * co2, torch.float32)
        y = G.make(n * 4 * lo * lo * co2, torch.float32)
        xs = G.make(n * 4 * lo * lo * 3 * co2, torch.float16)
        s2, q2 = G.make(n * co2, torch.int64), G.make(n * co2, torch.int64)
        A, B = G.make(n * co2, torch.float32, 1.0), G.make(n * co2, torch.float32)
        noise = torch.randn(4 * lo * lo, generator=g).cuda()
        nw, bias = torch.randn(co2, generator=g).cuda(), torch.randn(co2, generator=g).cuda()
        styles = torch.randn(n, 2 * co2, generator=g).cuda()
        prog.conv(inp=_split3_nhwc(x2), n=n, hin=lo, win=lo, cin=3 * ci2, w=E.split3_weight(wp, 4, ci2).cuda(), cout=co2,
                  hout=lo, wout=lo, tile=E.tile_for(lo), out=raw, out_hwc=(2 * lo, 2 * lo, co2), taps=taps, oscale=2,
                  ooff=[(0, 0), (0, 1), (1, 0), (1, 1)], w_rows_per_phase=co2, k_split=3)
        L.check(prog.lib.cfr_program_add_blur_act_stats_f32(prog.handle, L.ptr(raw), L.ptr(y), n, 2 * lo, 2 * lo, co2,
                                                            L.ptr(noise), L.ptr(nw), L.ptr(bias), L.ptr(s2), L.ptr(q2), 0))
        L.check(prog.lib.cfr_program_add_finalize_stats(prog.handle, L.ptr(s2), L.ptr(q2), L.ptr(styles), 2 * co2, 0, n,
                                                        co2, 1.0 / (4 * lo * lo), L.ptr(A), L.ptr(B)))
        L.check(prog.lib.cfr_program_add_affine_f32(prog.handle, L.ptr(y), L.ptr(A), L.ptr(B), n, 4 * lo * lo, co2,
                                                    L.ptr(xs), 3))
        # (3) halo conv, folded (per-sample weights written by the fold kernel), odd width, 9 rows
        h3, w3, c3 = 9, 130, 32
        y3 = torch.randn(n, c3, h3, w3, generator=g).cuda()
        out3 = G.make(n * h3 * w3 * c3, torch.float16)
        s3, q3 = G.make(n * c3, torch.int64), G.make(n * c3, torch.int64)
        A3 = (torch.rand(n, c3, generator=g) + 0.5).cuda().contiguous()
        B3 = torch.randn(n, c3, generator=g).cuda().contiguous()
        nz3 = torch.randn(h3 * w3, generator=g).cuda()
        nw3, b3 = torch.randn(c3, generator=g).cuda(), torch.randn(c3, generator=g).cuda()
        prog.conv(inp=_nhwc16(y3), n=n, hin=h3, win=w3, cin=c3, w=E.pack_halo_weight(
                      torch.randn(c3, c3, 3, 3, generator=g) / math.sqrt(c3 * 9)).cuda().float().contiguous(), cout=c3,
                  hout=h3, wout=w3, tile=(16, 8, 1), out=out3, out_hwc=(h3, w3, c3), taps=[E.TAPS3], bias=b3, noise=nz3,
                  noise_w=nw3, act=L.ACT_LRELU, slope=0.2, stat_sum=s3, stat_sq=q3, halo=True, in_affine=(A3, B3),
                  fold_center_tap=4)
        # (4) toRGB + resize into a guarded image buffer
        img = G.make(n * 16 * 16 * 16, torch.float16)
        x4 = torch.randn(n, 16, 64, 64, generator=g).cuda()
        A4, B4 = torch.ones(n * 16, device="cuda"), torch.zeros(n * 16, device="cuda")
        wr, br = torch.randn(3, 16, generator=g).cuda(), torch.randn(3, generator=g).cuda()
        L.check(prog.lib.cfr_program_add_torgb_resize(prog.handle, L.ptr(_nhwc16(x4)), L.ptr(A4), L.ptr(B4), n, 64, 16,
                                                      L.ptr(wr), L.ptr(br), 16, 0.5, 0.5, L.ptr(img), None, None))
        prog.keep += [noise, nw, bias, styles, nz3, nw3, b3, A4, B4, wr, br, x4]
        prog.run()
        prog.run()
        _sync()
        G.check()
        assert torch.isfinite(out.float()).all() and torch.isfinite(y).all() and torch.isfinite(out3.float()).all()
        monkeypatch.delenv("CFR_MAX_CTAS", raising=False)


@pytest.mark.parametrize("ctas", [None, "3"])
def test_conv_kernels_are_run_to_run_deterministic(E, monkeypatch, ctas):
    """A missing barrier / fence in the TMA -> mbarrier -> tcgen05 -> TMEM -> epilogue pipelines shows up as run-to-run
    differences: 20 replays of the same igemm (paired tiles, fused statistics) and halo (folded) launches must be
    bit-identical in outputs and (integer) statistics."""
    if ctas:
        monkeypatch.setenv("CFR_MAX_CTAS", ctas)
    L = E.L
    g = torch.Generator().manual_seed(23)
    n, c, res = 5, 64, 24
    x = torch.randn(n, c, res, res, generator=g).cuda()
    w = torch.randn(c, c, 3, 3, generator=g) / math.sqrt(c * 9)
    out = torch.zeros(n * res * res * c, dtype=torch.float16, device="cuda")
    ssum, ssq = _stats(n, c)
    h3, w3, c3 = 20, 136, 32
    y3 = torch.randn(n, c3, h3, w3, generator=g).cuda()
    out3 = torch.zeros(n * h3 * w3 * c3, dtype=torch.float16, device="cuda")
    s3, q3 = _stats(n, c3)
    A3 = (torch.rand(n, c3, generator=g) + 0.5).cuda().contiguous()
    B3 = torch.randn(n, c3, generator=g).cuda().contiguous()
    nz3 = torch.randn(h3 * w3, generator=g).cuda()
    nw3, b3 = torch.randn(c3, generator=g).cuda(), torch.randn(c3, generator=g).cuda()
    prog = E.Program()
    prog.memset(ssum), prog.memset(ssq), prog.memset(s3), prog.memset(q3)
    prog.conv(inp=_nhwc16(x), n=n, hin=res, win=res, cin=c, w=E.pack_conv_weight(w).cuda().half(), cout=c, hout=res,
              wout=res, tile=(16, 8, 1), out=out, out_hwc=(res, res, c), taps=[E.TAPS3], stat_sum=ssum, stat_sq=ssq)
    prog.conv(inp=_nhwc16(y3), n=n, hin=h3, win=w3, cin=c3, w=E.pack_halo_weight(
                  torch.randn(c3, c3, 3, 3, generator=g) / math.sqrt(c3 * 9)).cuda().float().contiguous(), cout=c3,
              hout=h3, wout=w3, tile=(16, 8, 1), out=out3, out_hwc=(h3, w3, c3), taps=[E.TAPS3], bias=b3, noise=nz3,
              noise_w=nw3, act=L.ACT_LRELU, slope=0.2, stat_sum=s3, stat_sq=q3, halo=True, in_affine=(A3, B3),
              fold_center_tap=4)
    first = None
    for _ in range(20):
        prog.run()
        _sync()
        snap = [t.clone() for t in (out, ssum, ssq, out3, s3, q3)]
        if first is None:
            first = snap
        else:
            assert all(torch.equal(a, b) for a, b in zip(first, snap))
