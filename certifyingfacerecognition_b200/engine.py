"""Host side of the B200 certification engine: packs the reference's weights into the layouts the CUDA
kernels consume and records the kernel programs (one for StyleGAN synthesis + resize, one for ArcFace).

PyTorch is used for device memory and one-time host-side weight folding only; every launch on the hot
path is one of our own kernels behind the C ABI (``include/cfr_b200.h``).

Weight dicts use the reference's ``state_dict()`` names:
  * StyleGAN  -- models/stylegan_generator_model.py
  * iresnet50 -- models/iresnet.py
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib as L

Tensor = torch.Tensor
NUM_LAYERS = 18
PSI, TRUNC_LAYERS = 0.7, 8          # models/model_settings.py:65-66
N_DIRS = 5                          # attack_utils/proj_utils.py:16-21 (ATTRS): age, eyeglasses, gender, pose, smile


def layer_channels(layer: int) -> int:
    table = [512, 512, 512, 512, 512, 256, 128, 64, 32, 16]      # stylegan_generator_model.py:23-32
    return table[layer // 2 + 1] if layer >= 2 else 512


def layer_res(layer: int) -> int:
    return 2 ** (layer // 2 + 2)


def tile_for(res: int, n: int = 1 << 30) -> Tuple[int, int, int]:
    """(TW, TH, TN) box of 128 output-grid pixels per M tile: the power-of-two box wasting the fewest rows on a
    res x res grid of n images (ties: widest box, i.e. longest contiguous TMA rows)."""
    best, best_key = None, None
    for lw in range(8):
        for lh in range(8 - lw):
            tw, th = 1 << lw, 1 << lh
            tn = 128 // (tw * th)
            if tw > 16 and tw > res:
                continue
            cover = (-(-res // tw) * tw) * (-(-res // th) * th) * (-(-n // tn) * tn if n < (1 << 30) else tn)
            useful = res * res * (n if n < (1 << 30) else tn)
            key = (useful / cover, tw, th)
            if best_key is None or key > best_key:
                best, best_key = (tw, th, tn), key
    return best


def _f16(t: Tensor, dev) -> Tensor:
    return t.to(device=dev, dtype=torch.float16).contiguous()


def _f32(t: Tensor, dev) -> Tensor:
    return t.to(device=dev, dtype=torch.float32).contiguous()


def pack_conv_weight(w: Tensor, cin_pad: Optional[int] = None) -> Tensor:
    """[Cout,Cin,kh,kw] fp32 -> [Cout, Kpad] with K = (tap=(ky,kx), cin), zero-padded to a multiple of 64."""
    cout, cin, kh, kw = w.shape
    cin_pad = cin_pad or cin
    m = torch.zeros(cout, kh * kw, cin_pad, dtype=torch.float32)
    m[:, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
    m = m.reshape(cout, kh * kw * cin_pad)
    kpad = (m.shape[1] + 63) // 64 * 64
    out = torch.zeros(cout, kpad, dtype=torch.float32)
    out[:, :m.shape[1]] = m
    return out


def split3_weight(m: Tensor, ntaps: int, cin: int) -> Tensor:
    """[rows, >= ntaps*cin] fp32 (K = (tap, cin)) -> fp16 [rows, ntaps*3*cin], per tap [w_hi | w_hi | w_lo] with
    hi = fp16(w), lo = fp16(w - hi): the weight side of a split-precision conv (cfr_conv_desc.kSplit == 3), matching
    activations laid out [x_hi | x_lo | x_hi]:  x.w = x_hi.w_hi + x_lo.w_hi + x_hi.w_lo  (+ O(2^-22))."""
    rows = m.shape[0]
    w = m[:, :ntaps * cin].reshape(rows, ntaps, cin).float()
    hi = w.half()
    lo = (w - hi.float()).half()
    out = torch.cat([hi, hi, lo], dim=2).reshape(rows, ntaps * 3 * cin)
    assert out.shape[1] % 64 == 0
    return out.contiguous()


TAPS3 = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]

# Sub-pixel decomposition of nearest-x2 + 3x3 conv (pad 1): output (2i+a, 2j+b) reads the low-res rows
# a=0: {i-1: w[0], i: w[1]+w[2]},  a=1: {i: w[0]+w[1], i+1: w[2]}   (same for columns).
_PHASE_ROWS = {0: [(-1, [0]), (0, [1, 2])], 1: [(0, [0, 1]), (1, [2])]}


def upconv_equiv_weight(sd: Dict[str, Tensor], layer: int) -> Tensor:
    """[Cout,Cin,3,3] (wscale applied) such that UpConvBlock == conv2d(nearest_x2(x), W, pad=1) before the blur.
    Fused layers (res >= 128, stylegan_generator_model.py:667-672) store [kh,kw,Cin,Cout] and are a
    conv_transpose of the box-summed kernel == the spatially flipped 3x3 conv on the upsampled grid."""
    p = f"synthesis.layer{layer}."
    if layer_res(layer) >= 128:
        w = sd[p + "weight"].float()
        scale = math.sqrt(2.0) / math.sqrt(w.shape[2] * 9)
        return torch.flip(w, dims=[0, 1]).permute(3, 2, 0, 1).contiguous() * scale
    w = sd[p + "conv.weight"].float()
    return w * (math.sqrt(2.0) / math.sqrt(w.shape[1] * 9))


def pack_upconv_phases(weq: Tensor) -> Tuple[Tensor, List[List[Tuple[int, int]]]]:
    """-> ([4*Cout, Kpad] rows = (phase, cout), K = (tap, cin)), taps[phase] = [(dy,dx) x4])."""
    cout, cin = weq.shape[:2]
    mats, taps = [], []
    for a in (0, 1):
        for b in (0, 1):
            cols, tp = [], []
            for dy, rs in _PHASE_ROWS[a]:
                for dx, cs in _PHASE_ROWS[b]:
                    cols.append(sum(weq[:, :, i, j] for i in rs for j in cs))     # [Cout,Cin]
                    tp.append((dy, dx))
            mats.append(torch.stack(cols, dim=1).reshape(cout, 4 * cin))
            taps.append(tp)
    m = torch.cat(mats, dim=0)
    kpad = (m.shape[1] + 63) // 64 * 64
    out = torch.zeros(m.shape[0], kpad)
    out[:, :m.shape[1]] = m
    return out, taps


def pack_halo_weight(w: Tensor) -> Tensor:
    """[Cout,Cin,3,3] -> [9*Cout, Cin], rows = (tap=(ky,kx), cout): the halo kernel's per-tap B tiles."""
    cout, cin = w.shape[:2]
    return w.permute(2, 3, 0, 1).reshape(9 * cout, cin).contiguous()


def pack_halo_upconv(weq: Tensor) -> Tuple[Tensor, List[List[Tuple[int, int]]]]:
    """[Cout,Cin,3,3] -> ([4*4*Cout, Cin] rows = (phase, tap, cout), taps[phase])."""
    mats, taps = [], []
    for a in (0, 1):
        for b in (0, 1):
            tp = []
            for dy, rs in _PHASE_ROWS[a]:
                for dx, cs in _PHASE_ROWS[b]:
                    mats.append(sum(weq[:, :, i, j] for i in rs for j in cs))
                    tp.append((dy, dx))
            taps.append(tp)
    return torch.cat(mats, dim=0).contiguous(), taps


# ---- blur o up-conv as ONE convolution on the low-res grid (StyleGAN UpConvBlock :665-676 incl. BlurLayer :463) ----
# out(2i+a, 2j+b) = sum_{u,v} k_u k_v raw(2i+a+u, 2j+b+v) [inside], raw = conv3x3(nearest_x2(x)).  Both the up-sampling
# and the blur are linear, so per phase (a,b) this is a 3x3 conv on x with weights  Wc = R_a (x) C_b (x) Weq,
# R_a[dy][s] = sum_u k_u [floor((a+u+s)/2) == dy].  The blur zero-pads the *raw* image, so the first / last hi-res row
# (column) simply drop u = -1 / u = +1 from that sum: "first"/"last" variants (exact: test_composite_upconv_blur_*).
def _blur_phase_matrix(a: int, cls: str) -> Tensor:
    k = torch.tensor([1.0, 2.0, 1.0], dtype=torch.float64) / 4.0
    m = torch.zeros(3, 3, dtype=torch.float64)
    for u in (-1, 0, 1):
        if (cls == "first" and u == -1) or (cls == "last" and u == 1):
            continue
        for s_ in (-1, 0, 1):
            m[(a + u + s_) // 2 + 1][s_ + 1] += k[u + 1]
    return m


# weight sets of the composite kernel: 0..3 = phases (a,b) on interior rows, 4..5 = first hi-res row (a=0, b=0/1),
# 6..7 = last hi-res row (a=1, b=0/1); columns always use the interior form (border columns get `corr`, below)
COMPOSITE_WSETS = [(0, 0, "int"), (0, 1, "int"), (1, 0, "int"), (1, 1, "int"),
                   (0, 0, "first"), (0, 1, "first"), (1, 0, "last"), (1, 1, "last")]


def composite_upconv_weights(weq: Tensor) -> Tuple[Tensor, Tensor]:
    """weq [Cout,Cin,3,3] (fp32/64) -> (base [8*9*Cout, Cin] rows = (wset, tap=(dy,dx), cout),
    corr_d [2 sides][2 a][3 row classes int/first/last][3 dy][Cout][Cin]).

    corr: the MMA path uses interior column weights for every pixel; for hi-res column 0 (b=0, j=0) / 2W-1 (b=1, j=W-1)
    the exact result differs by  -k_(-/+1) * sum_{dy,s} R_{a,rc}[dy][s] * Weq[s][t=+1/-1] . x(i+dy, 0 / W-1)."""
    w = weq.double()
    sets = []
    for a, b, rc in COMPOSITE_WSETS:
        wc = torch.einsum("ys,xt,oist->yxoi", _blur_phase_matrix(a, rc), _blur_phase_matrix(b, "int"), w)   # [3,3,Co,Ci]
        sets.append(wc.reshape(9 * w.shape[0], w.shape[1]))
    base = torch.cat(sets, dim=0)
    corr = torch.zeros(2, 2, 3, 3, w.shape[0], w.shape[1], dtype=torch.float64)
    for side, t in ((0, 2), (1, 0)):                     # left uses Weq[:, :, s, +1], right Weq[:, :, s, -1]
        for a in (0, 1):
            for rci, rc in enumerate(("int", "first", "last")):
                corr[side, a, rci] = -0.25 * torch.einsum("ys,ois->yoi", _blur_phase_matrix(a, rc), w[:, :, :, t])
    return base.float().contiguous(), corr.float().contiguous()


def resize_keep_map(hin: int, rout: int) -> Tuple[Tensor, int]:
    """Rows (== columns) of an hin x hin image that F.interpolate(bilinear, align_corners=False) to rout x rout reads
    (gen_utils.py:77-85; same arithmetic as k_torgb_resize): -> (int32 [hin] map: compact index or -1, count)."""
    import numpy as np
    scale = np.float32(hin) / np.float32(rout)
    keep = set()
    for o in range(rout):
        sy = max(np.float32(scale * np.float32(o + 0.5)) - np.float32(0.5), np.float32(0.0))
        for y0 in {int(sy), int(max(sy - 1e-3, 0.0)), int(sy + 1e-3)}:     # (fused vs separate multiply-add: keep both
            keep.add(min(y0, hin - 1))                                      #  candidates if sy sits on an integer)
            keep.add(min(y0 + 1, hin - 1))
    rows = sorted(keep)
    m = torch.full((hin,), -1, dtype=torch.int32)
    m[torch.tensor(rows)] = torch.arange(len(rows), dtype=torch.int32)
    return m, len(rows)


class Program:
    """Owns a cfr_program handle plus every tensor its launches reference."""

    def __init__(self):
        self.lib = L.load()
        h = C.c_void_p()
        L.check(self.lib.cfr_program_create(C.byref(h)))
        self.handle = h
        self.keep: List[Tensor] = []

    def hold(self, t: Tensor) -> Tensor:
        self.keep.append(t)
        return t

    def run(self, stream: Optional[int] = None) -> None:
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        L.check(self.lib.cfr_program_run(self.handle, C.c_void_p(stream)))

    def run_range(self, first: int, last: int, stream: Optional[int] = None) -> None:
        """Replay ops [first, last) only (diagnostics)."""
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        L.check(self.lib.cfr_program_run_range(self.handle, first, last, C.c_void_p(stream)))

    @property
    def num_launches(self) -> int:
        return self.lib.cfr_program_num_launches(self.handle)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.cfr_program_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- op recorders --------------------------------------------------------------------------
    def conv(self, *, inp: Tensor, n: int, hin: int, win: int, cin: int, w: Tensor, cout: int, hout: int, wout: int,
             tile: Tuple[int, int, int], out: Tensor, out_hwc: Tuple[int, int, int], taps, stride: int = 1,
             oscale: int = 1, ooff=((0, 0),), w_rows_per_sample: int = 0, w_rows_per_phase: int = 0,
             bias: Optional[Tensor] = None, cbias: Optional[Tensor] = None, cbias_per_sample: bool = False,
             noise: Optional[Tensor] = None, noise_w: Optional[Tensor] = None, act: int = L.ACT_NONE,
             slope: float = 0.2, alpha: Optional[Tensor] = None, resid: Optional[Tensor] = None, resid_c: int = 0,
             stat_sum: Optional[Tensor] = None, stat_sq: Optional[Tensor] = None, halo: bool = False,
             in_affine: Optional[Tuple[Tensor, Tensor]] = None, fold_center_tap: Optional[int] = None,
             composite_corr: Optional[Tensor] = None, k_split: int = 0,
             keep_map: Optional[Tensor] = None, keep_dim: int = 0) -> None:
        """Record one convolution.  ``halo``: use the halo-resident kernel (Cin, Cout <= 64); ``in_affine`` (A, B):
        apply x = y*A + B on load; ``fold_center_tap`` (halo, Cin <= 32): fold A into per-sample weights and carry
        B / bias / noise on the auxiliary band -- ``w`` is then the fp32 base weight [phases*taps*Cout, Cin]."""
        d = L.ConvDesc()
        d.inp, d.N, d.Hin, d.Win, d.Cin = L.ptr(inp), n, hin, win, cin
        d.w, d.wRows, d.Kpad = L.ptr(w), w.shape[0], w.shape[1]
        d.Cout, d.Hout, d.Wout = cout, hout, wout
        d.TW, d.TH, d.TN = tile
        d.stride, d.numPhases = stride, len(taps)
        d.ntaps = len(taps[0])
        for ph, tp in enumerate(taps):
            assert len(tp) == d.ntaps
            for k, (dy, dx) in enumerate(tp):
                d.tap_dy[ph][k] = dy
                d.tap_dx[ph][k] = dx
        d.wRowsPerSample, d.wRowsPerPhase = w_rows_per_sample, w_rows_per_phase
        d.out, d.outIsF32 = L.ptr(out), int(out.dtype == torch.float32)
        d.outH, d.outW, d.outC = out_hwc
        d.oscale = oscale
        for ph, (oy, ox) in enumerate(ooff):
            d.ooff_y[ph] = oy
            d.ooff_x[ph] = ox
        d.bias, d.cbias, d.cbiasPerSample = L.ptr(bias), L.ptr(cbias), int(cbias_per_sample)
        d.noise, d.noise_w = L.ptr(noise), L.ptr(noise_w)
        d.act, d.slope, d.alpha = act, slope, L.ptr(alpha)
        d.resid, d.residC = L.ptr(resid), resid_c
        d.stat_sum, d.stat_sq = L.ptr(stat_sum), L.ptr(stat_sq)
        d.kSplit = k_split
        d.keepMap, d.keepDim = L.ptr(keep_map), keep_dim
        if keep_map is not None:
            assert halo, "sparse store is a halo-kernel feature"
            self.keep.append(keep_map)
        assert k_split in (0, 1) or not halo, "split-precision convs run on the implicit-GEMM kernel"
        for t in (inp, w, out, bias, cbias, noise, noise_w, alpha, resid, stat_sum, stat_sq):
            if t is not None:
                self.keep.append(t)
        if halo:
            a, b = in_affine if in_affine is not None else (None, None)
            for t in (a, b):
                if t is not None:
                    self.keep.append(t)
            if fold_center_tap is not None:
                assert w.dtype == torch.float32
                rows = n * w.shape[0]
                w_main = self.hold(torch.zeros(rows, cin, dtype=torch.float16, device=w.device))
                if composite_corr is not None:          # blur o up-conv: 8 weight sets, border-column correction
                    self.keep.append(composite_corr)
                    w_aux = self.hold(torch.zeros(n * 8 * cout, 16, dtype=torch.float16, device=w.device))
                    corr = self.hold(torch.zeros(n * 2 * out_hwc[0] * cout, device=w.device))
                    L.check(self.lib.cfr_program_add_upconv_blur_folded(self.handle, C.byref(d), L.ptr(w),
                                                                        L.ptr(composite_corr), L.ptr(a), L.ptr(b),
                                                                        L.ptr(w_main), L.ptr(w_aux), L.ptr(corr)))
                else:
                    w_aux = self.hold(torch.zeros(n * len(taps) * cout, 16, dtype=torch.float16, device=w.device))
                    L.check(self.lib.cfr_program_add_conv_halo_folded(self.handle, C.byref(d), L.ptr(w), L.ptr(a),
                                                                      L.ptr(b), L.ptr(w_main), L.ptr(w_aux)))
            else:
                L.check(self.lib.cfr_program_add_conv_halo(self.handle, C.byref(d), L.ptr(a), L.ptr(b)))
        else:
            assert in_affine is None, "affine-on-load is a halo-kernel feature"
            L.check(self.lib.cfr_program_add_conv(self.handle, C.byref(d)))

    def memset(self, t: Tensor, value: int = 0) -> None:
        self.keep.append(t)
        L.check(self.lib.cfr_program_add_memset(self.handle, L.ptr(t), value, t.numel() * t.element_size()))


# ------------------------------------------------------------------------------------------------------
# StyleGAN-FFHQ-1024 synthesis + toRGB + resize  (wp2 [chunk,2,512] -> img [chunk,R,R,16] fp16 NHWC)
# ------------------------------------------------------------------------------------------------------
class SynthesisProgram(Program):
    def __init__(self, g_sd: Dict[str, Tensor], chunk: int, out_res: int = 112, device="cuda",
                 keep_planar: bool = False, mean: float = 0.5, std: float = 0.5, halo: bool = True,
                 fold_small: bool = True, groups: int = 1, blur_on_tensor_cores: bool = True,
                 fused_upblur: bool = True, nhwc_out: bool = True, hp_layers: int = 0, sparse_last: bool = True):
        super().__init__()
        dev = torch.device(device)
        self.chunk, self.out_res = chunk, out_res
        import os as _os
        fold_max_cin = 64                                            # widest layer input that is folded
        if _os.environ.get("CFR_DEBUG_KNOBS") == "1":                # A/B knobs for profiling runs only: they change
            if _os.environ.get("CFR_FUSED_UPBLUR") is not None:      # numerics, so a stray variable must not reach the
                fused_upblur = _os.environ["CFR_FUSED_UPBLUR"] != "0"   # product path
            fold_max_cin = int(_os.environ.get("CFR_FOLD_MAX_CIN", "64"))
        sd = {k: v.detach().float().cpu() for k, v in g_sd.items()}
        lib, h = self.lib, self.handle
        # Split-precision prefix: layers 1..hp_layers take fp16 hi/lo operand pairs (3 MMAs per product, ~2^-21) and keep
        # fp32 activations between kernels.  Rounding errors of the first (4x4 .. 64x64) layers are spatially coarse, get
        # amplified by every later InstanceNorm and dominate the embedding error (tests/diag_precision.py).
        if not 0 <= hp_layers <= 11:
            raise ValueError("hp_layers must be in 0..11 (layer 12 feeds the halo kernel, which reads fp16)")
        self.hp_layers = hp_layers
        # Sparse store of the last layer: bilinear-resizing 1024^2 to 112^2 (160^2) reads only 224 (320) distinct rows and
        # columns, so layer 17 writes just those pixels into a compact buffer (its statistics still cover every pixel)
        kmap, kdim = resize_keep_map(1024, out_res)
        self.sparse_last = bool(sparse_last and halo and fold_small and kdim <= 512)
        self.keep_map = self.hold(kmap.to(dev)) if self.sparse_last else None
        self.keep_dim = kdim if self.sparse_last else 0

        # ---- inputs / small tensors
        self.wp2 = self.hold(torch.zeros(chunk, 2, 512, device=dev))
        self.w_avg = self.hold(_f32(sd["truncation.w_avg"], dev))
        ws, bs, self.style_off = [], [], []
        off = 0
        for l in range(NUM_LAYERS):
            p = f"synthesis.layer{l}.epilogue.style_mod.dense."
            ws.append(sd[p + "linear.weight"])
            bs.append(sd[p + "wscale.bias"])
            self.style_off.append(off)
            off += 2 * layer_channels(l)
        self.style_rows = off
        rows_trunc = self.style_off[TRUNC_LAYERS]
        w_style = self.w_style = self.hold(_f32(torch.cat(ws), dev))      # (kept: StyleGANGenerator's WP path fills `styles` itself)
        b_style = self.b_style = self.hold(_f32(torch.cat(bs), dev))
        self.styles = self.hold(torch.zeros(chunk, off, device=dev))
        L.check(lib.cfr_program_add_styles(h, L.ptr(self.wp2), L.ptr(w_style), L.ptr(b_style), off, rows_trunc,
                                           chunk, L.ptr(self.styles)))

        # ---- statistics / affine buffers, zeroed at the start of every run
        total_c = sum(layer_channels(l) for l in range(NUM_LAYERS))
        self.stats = self.hold(torch.zeros(2, chunk * total_c, dtype=torch.int64, device=dev))   # Q43.20 fixed point
        self.memset(self.stats)
        self.AB = [(self.hold(torch.zeros(chunk * 512, device=dev)), self.hold(torch.zeros(chunk * 512, device=dev)))
                   for _ in range(2)]

        # ---- activation buffers (NHWC fp16), sized for the largest layer
        max_elems = chunk * 1024 * 1024 * 16
        bufs = [self.hold(torch.empty(max_elems, dtype=torch.float16, device=dev)) for _ in range(3)]
        x, y, raw = bufs

        # ---- layer 0: sample-independent normalised const (FirstConvBlock :581-584 + epilogue :559-564)
        p0 = "synthesis.layer0.epilogue."
        c0 = sd["synthesis.layer0.first_layer"][0]                                   # [512,4,4]
        t = c0 + sd[p0 + "apply_noise.noise"][0] * sd[p0 + "apply_noise.weight"].view(-1, 1, 1) \
            + sd[p0 + "bias"].view(-1, 1, 1)
        t = torch.nn.functional.leaky_relu(t, 0.2)
        t = t - t.mean(dim=[1, 2], keepdim=True)
        t = t / torch.sqrt((t * t).mean(dim=[1, 2], keepdim=True) + 1e-8)
        xhat0 = self.hold(_f32(t.permute(1, 2, 0).reshape(16, 512), dev))
        add_l0 = lib.cfr_program_add_layer0_split if hp_layers >= 1 else lib.cfr_program_add_layer0
        L.check(add_l0(h, L.ptr(xhat0), L.ptr(self.styles), off, self.style_off[0], chunk, L.ptr(x)))

        # diagnostics: (layer, ops recorded when its output is complete, buffer, res, C, (A, B) still to apply or None)
        self.layer_marks = [(0, self.num_launches, x, 4, 512, None)]
        stat_off = chunk * 512          # layer 0 uses no statistics slot
        pending = None                  # (A, B) of the previous layer when its IN/AdaIN has not been applied yet
        use_halo = lambda l: halo and layer_channels(l - 1) <= 64 and layer_channels(l) <= 64
        for l in range(1, NUM_LAYERS):
            cin, cout, res = layer_channels(l - 1), layer_channels(l), layer_res(l)
            pe = f"synthesis.layer{l}.epilogue."
            noise = self.hold(_f32(sd[pe + "apply_noise.noise"].reshape(-1), dev))
            noise_w = self.hold(_f32(sd[pe + "apply_noise.weight"], dev))
            bias = self.hold(_f32(sd[pe + "bias"], dev))
            ssum = self.stats[0, stat_off:stat_off + chunk * cout]
            ssq = self.stats[1, stat_off:stat_off + chunk * cout]
            stat_off += chunk * cout
            hk = use_halo(l)
            fold = hk and fold_small and cin <= fold_max_cin
            assert pending is None or hk
            if l <= hp_layers:
                # ---- split-precision layer: x [.., 3*cin] = [hi|lo|hi] fp16 -> y fp32 -> x' fp16 (split again or plain)
                y32, raw32 = y.view(torch.float32), raw.view(torch.float32)
                if l % 2 == 1:
                    w = sd[f"synthesis.layer{l}.conv.weight"] * (math.sqrt(2.0) / math.sqrt(cin * 9))
                    fused_stats = res >= 16
                    self.conv(inp=x, n=chunk, hin=res, win=res, cin=3 * cin, w=self.hold(split3_weight(
                                  pack_conv_weight(w), 9, cin).to(dev)), cout=cout, hout=res, wout=res,
                              tile=tile_for(res), out=y32, out_hwc=(res, res, cout), taps=[TAPS3], noise=noise,
                              noise_w=noise_w, bias=bias, act=L.ACT_LRELU, slope=0.2,
                              stat_sum=ssum if fused_stats else None, stat_sq=ssq if fused_stats else None, k_split=3)
                    if not fused_stats:
                        L.check(lib.cfr_program_add_blur_act_stats_f32(h, L.ptr(y32), None, chunk, res, res, cout, None,
                                                                       None, None, L.ptr(ssum), L.ptr(ssq), 1))
                else:
                    lo = res // 2
                    wp, taps = pack_upconv_phases(upconv_equiv_weight(sd, l))
                    self.conv(inp=x, n=chunk, hin=lo, win=lo, cin=3 * cin, w=self.hold(split3_weight(wp, 4, cin).to(dev)),
                              cout=cout, hout=lo, wout=lo, tile=tile_for(lo), out=raw32, out_hwc=(res, res, cout),
                              taps=taps, oscale=2, ooff=[(0, 0), (0, 1), (1, 0), (1, 1)], w_rows_per_phase=cout, k_split=3)
                    L.check(lib.cfr_program_add_blur_act_stats_f32(h, L.ptr(raw32), L.ptr(y32), chunk, res, res, cout,
                                                                   L.ptr(noise), L.ptr(noise_w), L.ptr(bias), L.ptr(ssum),
                                                                   L.ptr(ssq), 0))
                A, B = self.AB[l % 2]
                L.check(lib.cfr_program_add_finalize_stats(h, L.ptr(ssum), L.ptr(ssq), L.ptr(self.styles), off,
                                                           self.style_off[l], chunk, cout, 1.0 / (res * res),
                                                           L.ptr(A), L.ptr(B)))
                # the conv's input buffer is free now: the next layer's operands go there
                L.check(lib.cfr_program_add_affine_f32(h, L.ptr(y32), L.ptr(A), L.ptr(B), chunk, res * res, cout, L.ptr(x),
                                                       3 if l + 1 <= hp_layers else 1))
                self.layer_marks.append((l, self.num_launches, y32, res, cout, (A, B)))
                continue
            if l % 2 == 1:
                w = sd[f"synthesis.layer{l}.conv.weight"] * (math.sqrt(2.0) / math.sqrt(cin * 9))
                fused_stats = res >= 16
                if fold:
                    wp = self.hold(_f32(pack_halo_weight(w), dev))
                else:
                    wp = self.hold(_f16(pack_halo_weight(w) if hk else pack_conv_weight(w), dev))
                self.conv(inp=x, n=chunk, hin=res, win=res, cin=cin, w=wp, cout=cout, hout=res, wout=res,
                          tile=tile_for(res), out=y, out_hwc=(res, res, cout), taps=[TAPS3], noise=noise,
                          noise_w=noise_w, bias=bias, act=L.ACT_LRELU, slope=0.2,
                          stat_sum=ssum if fused_stats else None, stat_sq=ssq if fused_stats else None,
                          halo=hk, in_affine=pending, fold_center_tap=4 if fold else None,
                          keep_map=self.keep_map if (l == NUM_LAYERS - 1 and fold) else None,
                          keep_dim=self.keep_dim if (l == NUM_LAYERS - 1 and fold) else 0)
                if not fused_stats:
                    L.check(lib.cfr_program_add_blur_act_stats(h, L.ptr(y), None, chunk, res, res, cout, None, None,
                                                               None, L.ptr(ssum), L.ptr(ssq), 1))
            else:
                lo = res // 2
                weq = upconv_equiv_weight(sd, l)
                if fold and cout <= 16 and fused_upblur:
                    # UpConvBlock incl. BlurLayer + epilogue as ONE halo conv (composite 3x3 weights per phase)
                    base, corr_d = composite_upconv_weights(weq)
                    self.conv(inp=x, n=chunk, hin=lo, win=lo, cin=cin, w=self.hold(_f32(base, dev)), cout=cout, hout=lo,
                              wout=lo, tile=tile_for(lo), out=y, out_hwc=(res, res, cout), taps=[TAPS3] * 4, oscale=2,
                              ooff=[(0, 0), (0, 1), (1, 0), (1, 1)], noise=noise, noise_w=noise_w, bias=bias,
                              act=L.ACT_LRELU, slope=0.2, stat_sum=ssum, stat_sq=ssq, halo=True, in_affine=pending,
                              fold_center_tap=4, composite_corr=self.hold(_f32(corr_d, dev)))
                else:
                    wp, taps = pack_halo_upconv(weq) if hk else pack_upconv_phases(weq)
                    wp = self.hold(_f32(wp, dev) if fold else _f16(wp, dev))
                    self.conv(inp=x, n=chunk, hin=lo, win=lo, cin=cin, w=wp, cout=cout, hout=lo, wout=lo,
                              tile=tile_for(lo), out=raw, out_hwc=(res, res, cout), taps=taps, oscale=2,
                              ooff=[(0, 0), (0, 1), (1, 0), (1, 1)], w_rows_per_phase=0 if hk else cout,
                              halo=hk, in_affine=pending, fold_center_tap=-1 if fold else None)
                    if fold and cout <= 16 and blur_on_tensor_cores:
                        # BlurLayer :463 as a depthwise conv on the tensor cores: W[tap] = k[tap] * I (1/16, 1/8, 1/4 are
                        # exact in fp16, accumulation is fp32 => same numerics as the CUDA-core blur), + noise/bias/act/stats
                        k1 = torch.tensor([1.0, 2.0, 1.0]) / 4.0
                        wb = torch.einsum("a,b,oi->oiab", k1, k1, torch.eye(cout))
                        self.conv(inp=raw, n=chunk, hin=res, win=res, cin=cout, w=self.hold(_f32(pack_halo_weight(wb), dev)),
                                  cout=cout, hout=res, wout=res, tile=tile_for(res), out=y, out_hwc=(res, res, cout),
                                  taps=[TAPS3], noise=noise, noise_w=noise_w, bias=bias, act=L.ACT_LRELU, slope=0.2,
                                  stat_sum=ssum, stat_sq=ssq, halo=True, fold_center_tap=4)
                    else:
                        L.check(lib.cfr_program_add_blur_act_stats(h, L.ptr(raw), L.ptr(y), chunk, res, res, cout,
                                                                   L.ptr(noise), L.ptr(noise_w), L.ptr(bias), L.ptr(ssum),
                                                                   L.ptr(ssq), 0))
            # A/B double-buffered by layer parity: the consumer of layer l reads them while layer l+1's are written
            A, B = self.AB[l % 2]
            L.check(lib.cfr_program_add_finalize_stats(h, L.ptr(ssum), L.ptr(ssq), L.ptr(self.styles), off,
                                                       self.style_off[l], chunk, cout, 1.0 / (res * res),
                                                       L.ptr(A), L.ptr(B)))
            if l < NUM_LAYERS - 1 and not use_halo(l + 1):
                L.check(lib.cfr_program_add_affine(h, L.ptr(y), L.ptr(A), L.ptr(B), chunk, res * res, cout, L.ptr(y)))
                pending = None
            else:
                pending = (A, B)            # applied on load by the consumer (halo conv / toRGB)
            self.layer_marks.append((l, self.num_launches, y, res, cout, pending))
            if l < NUM_LAYERS - 1:
                x, y = y, x
        # ---- toRGB + postprocess + bilinear + normalise; the last layer's IN/AdaIN is applied on load
        wrgb = sd["synthesis.output8.conv.weight"].reshape(3, 16) * (1.0 / math.sqrt(16))
        w_rgb = self.hold(_f32(wrgb, dev))
        b_rgb = self.hold(_f32(sd["synthesis.output8.bias"], dev))
        # the image buffer holds `groups` chunks; the op writes group *out_slot (set by the sampler between runs)
        self.groups = groups
        self.out_slot = self.hold(torch.zeros(1, dtype=torch.int32, device=dev))
        self.img = (self.hold(torch.zeros(groups * chunk, out_res, out_res, 16, dtype=torch.float16, device=dev))
                    if nhwc_out else None)
        self.img_planar = self.hold(torch.zeros(groups * chunk, 3, out_res, out_res, device=dev)) if keep_planar else None
        if self.sparse_last:
            L.check(lib.cfr_program_add_torgb_resize_sparse(h, L.ptr(y), L.ptr(pending[0]), L.ptr(pending[1]), chunk, 1024, 16,
                                                            L.ptr(w_rgb), L.ptr(b_rgb), out_res, mean, std, L.ptr(self.img),
                                                            L.ptr(self.img_planar), L.ptr(self.out_slot),
                                                            L.ptr(self.keep_map), self.keep_dim))
        else:
            L.check(lib.cfr_program_add_torgb_resize(h, L.ptr(y), L.ptr(pending[0]), L.ptr(pending[1]), chunk, 1024, 16,
                                                     L.ptr(w_rgb), L.ptr(b_rgb), out_res, mean, std, L.ptr(self.img),
                                                     L.ptr(self.img_planar), L.ptr(self.out_slot)))
        self.last_y = y


# ------------------------------------------------------------------------------------------------------
# ArcFace iresnet50  (img [chunk,112,112,16] fp16 -> emb [chunk,512] fp32)
# ------------------------------------------------------------------------------------------------------
def _bn_affine(sd, p, eps=1e-5):
    s = sd[p + ".weight"] / torch.sqrt(sd[p + ".running_var"] + eps)
    return s, sd[p + ".bias"] - sd[p + ".running_mean"] * s


class ArcFaceProgram(Program):
    LAYERS = (3, 4, 14, 3)
    PLANES = (64, 128, 256, 512)
    FC_SPLIT = 4                   # K-slices of the final FC (<= CFR_MAX_PHASES)

    def __init__(self, f_sd: Dict[str, Tensor], chunk: int, img: Tensor, device="cuda"):
        super().__init__()
        dev = torch.device(device)
        sd = {k: v.detach().float().cpu() for k, v in f_sd.items()}
        self.chunk = chunk
        n = chunk
        max_elems = n * 112 * 112 * 64
        xs = self.hold(torch.empty(max_elems, dtype=torch.float16, device=dev))     # residual stream
        xs2 = self.hold(torch.empty(max_elems, dtype=torch.float16, device=dev))
        hbuf = self.hold(torch.empty(max_elems, dtype=torch.float16, device=dev))    # conv1 output
        idb = self.hold(torch.empty(max_elems // 4, dtype=torch.float16, device=dev))  # downsample output

        # stem: conv1 3->64 + bn1 + prelu (iresnet.py:142-144); input channels padded 3 -> 16
        s, t = _bn_affine(sd, "bn1")
        w = pack_conv_weight(sd["conv1.weight"] * s.view(-1, 1, 1, 1), cin_pad=16)
        self.conv(inp=img, n=n, hin=112, win=112, cin=16, w=self.hold(_f16(w, dev)), cout=64, hout=112, wout=112,
                  tile=tile_for(112), out=xs, out_hwc=(112, 112, 64), taps=[TAPS3], bias=self.hold(_f32(t, dev)),
                  act=L.ACT_PRELU, alpha=self.hold(_f32(sd["prelu.weight"], dev)))
        res, inplanes = 112, 64
        for li, (nblocks, planes) in enumerate(zip(self.LAYERS, self.PLANES), start=1):
            for bi in range(nblocks):
                p = f"layer{li}.{bi}."
                stride = 2 if bi == 0 else 1
                ores = res // stride
                s1, t1 = _bn_affine(sd, p + "bn1")
                s2, t2 = _bn_affine(sd, p + "bn2")
                s3, t3 = _bn_affine(sd, p + "bn3")
                # conv1: bn1 (pre, scale folded into input channels; shift -> border-class bias), bn2 (post), PReLU
                w1 = sd[p + "conv1.weight"] * s2.view(-1, 1, 1, 1)
                tb = torch.einsum("oikl,i->okl", w1, t1).reshape(planes, 9)              # per-tap shift term
                cb = torch.zeros(9, planes)
                for rc in range(3):
                    for cc in range(3):
                        valid = [k for k, (dy, dx) in enumerate(TAPS3)
                                 if not (rc == 0 and dy < 0) and not (rc == 2 and dy > 0)
                                 and not (cc == 0 and dx < 0) and not (cc == 2 and dx > 0)]
                        cb[rc * 3 + cc] = tb[:, valid].sum(dim=1)
                w1p = pack_conv_weight(w1 * s1.view(1, -1, 1, 1))
                cb = cb + t2.view(1, -1)                # bn2's shift rides on the class bias: one vector less to fetch
                self.conv(inp=xs, n=n, hin=res, win=res, cin=inplanes, w=self.hold(_f16(w1p, dev)), cout=planes,
                          hout=res, wout=res, tile=tile_for(res, n), out=hbuf, out_hwc=(res, res, planes), taps=[TAPS3],
                          cbias=self.hold(_f32(cb, dev)), act=L.ACT_PRELU,
                          alpha=self.hold(_f32(sd[p + "prelu.weight"], dev)))
                identity = xs
                if bi == 0:
                    sdn, tdn = _bn_affine(sd, p + "downsample.1")
                    wd = pack_conv_weight(sd[p + "downsample.0.weight"] * sdn.view(-1, 1, 1, 1))
                    self.conv(inp=xs, n=n, hin=res, win=res, cin=inplanes, w=self.hold(_f16(wd, dev)), cout=planes,
                              hout=ores, wout=ores, tile=tile_for(ores, n), out=idb, out_hwc=(ores, ores, planes),
                              taps=[[(0, 0)]], stride=stride, bias=self.hold(_f32(tdn, dev)))
                    identity = idb
                w2p = pack_conv_weight(sd[p + "conv2.weight"] * s3.view(-1, 1, 1, 1))
                self.conv(inp=hbuf, n=n, hin=res, win=res, cin=planes, w=self.hold(_f16(w2p, dev)), cout=planes,
                          hout=ores, wout=ores, tile=tile_for(ores, n), out=xs2, out_hwc=(ores, ores, planes),
                          taps=[TAPS3], stride=stride, bias=self.hold(_f32(t3, dev)), resid=identity, resid_c=planes)
                xs, xs2 = xs2, xs
                res, inplanes = ores, planes
        # bn2 -> flatten (NCHW order, iresnet.py:149-150) -> fc -> features BN1d, all folded into one GEMM
        sb, tb2 = _bn_affine(sd, "bn2")
        sf, tf = _bn_affine(sd, "features")
        wfc = sd["fc.weight"].view(512, 512, 7, 7)
        bias = (sd["fc.bias"] + torch.einsum("ochw,c->o", wfc, tb2)) * sf + tf
        wfold = (wfc * sb.view(1, -1, 1, 1) * sf.view(-1, 1, 1, 1)).permute(0, 2, 3, 1).reshape(512, 49 * 512)
        self.emb = self.hold(torch.zeros(n, 512, device=dev))
        # split-K: the 25088-long contraction as FC_SPLIT "phases" of one conv -- phase p reads K-slice p (the input viewed
        # as [n, 1, FC_SPLIT, 25088 / FC_SPLIT], tap dx = p) against weight rows [p*512, (p+1)*512) and writes its fp32
        # partial sums to column p of [n, FC_SPLIT, 512]; k_sum_partials adds them in a fixed order (+ bias).  One K-slice
        # per work item: 4x the CTAs of the 8 (M tile, N tile) pairs a plain GEMM of this shape has.
        P = self.FC_SPLIT
        kp = 49 * 512 // P
        assert 49 * 512 % P == 0 and kp % 64 == 0
        wsplit = wfold.reshape(512, P, kp).permute(1, 0, 2).reshape(P * 512, kp).contiguous()
        partial = self.hold(torch.zeros(n * P * 512, device=dev))
        self.conv(inp=xs, n=n, hin=1, win=P, cin=kp, w=self.hold(_f16(wsplit, dev)), cout=512, hout=1, wout=1,
                  tile=(1, 1, 128), out=partial, out_hwc=(1, P, 512), taps=[[(0, p)] for p in range(P)],
                  ooff=[(0, p) for p in range(P)], w_rows_per_phase=512)
        L.check(self.lib.cfr_program_add_sum_partials(self.handle, L.ptr(partial), L.ptr(self.hold(_f32(bias, dev))), n, P,
                                                      512, L.ptr(self.emb)))
        self.final_features = xs


# ------------------------------------------------------------------------------------------------------
class _Pipeline:
    """One recorded (synthesis, FRM[, grouped FRM]) program set for a fixed chunk size, plus its C sampler."""

    def __init__(self, g_sd, f_sd, frm_cls, res, chunk, frm_group, device, keep_planar, hp_layers, sparse_last=True):
        self.chunk, self.frm_group = chunk, max(1, int(frm_group))
        import os as _os
        split = _os.environ.get("CFR_SPLIT_SMS") if _os.environ.get("CFR_DEBUG_KNOBS") == "1" else None
        if split:                      # experiment: "S,F" = persistent-grid caps of the synthesis / FRM programs
            _os.environ["CFR_MAX_CTAS"] = split.split(",")[0]
        self.synth = SynthesisProgram(g_sd, chunk, res, device, keep_planar=keep_planar, groups=self.frm_group,
                                      hp_layers=hp_layers, sparse_last=sparse_last)
        # the FRM programs read their own copy of the images: the sampler runs them (+ match + vote) on a second stream
        # while the caller's stream already synthesises the next group into synth.img (cfr_sampler_desc.img_frm)
        self.img_frm = torch.zeros_like(self.synth.img)
        if split:
            _os.environ["CFR_MAX_CTAS"] = split.split(",")[1]
        self.frm = frm_cls(f_sd, chunk, self.img_frm, device)
        self.frm_big = frm_cls(f_sd, chunk * self.frm_group, self.img_frm, device) if self.frm_group > 1 else None
        if split:
            del _os.environ["CFR_MAX_CTAS"]
        self.sampler = None


class Engine:
    """StyleGAN -> resize -> iresnet50 -> gallery vote, for one GPU."""

    TC_MATCH_MIN_ROWS = 32768      # galleries at least this large use the tensor-core matcher (cfr_matcher_*)
    # StyleGAN layers 1..5 (4x4 .. 16x16) run split-precision (SynthesisProgram.hp_layers).  Measured on the reference's
    # golden votes (profiles/votes_r02_hp_sweep.txt, 2 x 1100 samples): strict top-1 agreement iso / aniso 99.00 / 99.55 %
    # at 0, 99.45 / 99.73 % at 4, 99.64 / 99.82 % at 5, 99.64 / 100 % at 6, 99.73 / 99.82 % at 8 (the remaining flips are
    # near-ties of the reference itself, margins 0.004-0.03); 5 is the shortest prefix above the 99.5 % bar in both regimes
    HP_LAYERS = 5

    def __init__(self, g_sd, f_sd, dir_mat: Tensor, gallery: Tensor, chunk: int = 32, device="cuda",
                 keep_planar: bool = False, frm_group: int = 1, tc_match: Optional[bool] = None,
                 frm: str = "insightface", hp_layers: Optional[int] = None, tail_chunks=(), sparse_last: bool = True):
        """``tail_chunks``: extra, smaller chunk sizes recorded as their own program sets (descending, all < chunk); the
        remainder of a ``sample_votes`` call runs on the smallest one that holds it instead of a whole chunk.
        ``sparse_last=False`` keeps the dense 1024^2 output of the last StyleGAN layer (diagnostics only)."""
        if not torch.cuda.is_available():
            raise RuntimeError("certifyingfacerecognition_b200 needs a CUDA device (no CPU fallback)")
        self.lib = L.load()
        self.device = torch.device(device)
        self.chunk = chunk
        self.frm_group = max(1, int(frm_group))
        # FRM: ArcFace iresnet50 at 112^2 (main_attack.py:123-125) or FaceNet InceptionResnetV1 at 160^2 (:126-129);
        # gen_utils.py:17-21 INP_RESOLS gives the resize target, MEAN = STD = 0.5 for both
        self.frm_name = frm
        if frm == "insightface":
            frm_cls, res = ArcFaceProgram, 112
        elif frm in ("facenet", "facenet-vggface2"):
            from .models.facenet import FaceNetProgram, INPUT_RES
            frm_cls, res = FaceNetProgram, INPUT_RES
        else:
            raise ValueError(f"unknown face recognition model '{frm}'")
        if hp_layers is None:
            import os as _os
            hp_layers = self.HP_LAYERS
            if _os.environ.get("CFR_DEBUG_KNOBS") == "1" and _os.environ.get("CFR_HP_LAYERS") is not None:
                hp_layers = int(_os.environ["CFR_HP_LAYERS"])       # A/B runs (precision vs throughput)
        self.hp_layers = hp_layers
        tail_chunks = sorted({int(c) for c in tail_chunks if 0 < int(c) < chunk}, reverse=True)
        self.pipes = [_Pipeline(g_sd, f_sd, frm_cls, res, chunk, self.frm_group, device, keep_planar, hp_layers,
                                sparse_last)]
        self.pipes += [_Pipeline(g_sd, f_sd, frm_cls, res, c, 1, device, False, hp_layers, sparse_last)
                       for c in tail_chunks]
        self.synth, self.frm, self.frm_big = self.pipes[0].synth, self.pipes[0].frm, self.pipes[0].frm_big
        if tuple(dir_mat.shape) != (N_DIRS, 512):
            # the noise kernel (k_noise_project) and the C ABI fix the attribute space at the reference's five
            # InterFaceGAN directions (proj_utils.py:16-21 ATTRS); fewer / more rows would read out of bounds
            raise ValueError(f"direction matrix must be [{N_DIRS}, 512] (got {tuple(dir_mat.shape)})")
        self.dir_mat = _f32(dir_mat, self.device)
        self.tc_match = tc_match
        self.matcher = None
        self.set_gallery(gallery)

    def set_gallery(self, gallery: Tensor) -> None:
        self.gallery = _f32(gallery, self.device)
        if getattr(self, "matcher", None):
            self.lib.cfr_matcher_destroy(self.matcher)
            self.matcher = None
        use_tc = self.tc_match if self.tc_match is not None else self.gallery.shape[0] >= self.TC_MATCH_MIN_ROWS
        if use_tc:
            m = C.c_void_p()
            L.check(self.lib.cfr_matcher_create(L.ptr(self.gallery), self.gallery.shape[0], self.chunk * self.frm_group,
                                                self._stream(), C.byref(m)))
            self.matcher = m
        tail = None
        for pipe in reversed(self.pipes):              # smallest chunk first: each sampler points at the next smaller one
            d = L.SamplerDesc()
            d.synth, d.frm, d.chunk = pipe.synth.handle, pipe.frm.handle, pipe.chunk
            d.wp2, d.emb = L.ptr(pipe.synth.wp2), L.ptr(pipe.frm.emb)
            d.dir_mat, d.w_avg, d.psi = L.ptr(self.dir_mat), L.ptr(pipe.synth.w_avg), PSI
            d.gallery, d.n_gallery = L.ptr(self.gallery), self.gallery.shape[0]
            d.frm_group = pipe.frm_group
            d.frm_big = pipe.frm_big.handle if pipe.frm_big is not None else None
            d.emb_big = L.ptr(pipe.frm_big.emb) if pipe.frm_big is not None else None
            d.out_slot = L.ptr(pipe.synth.out_slot)
            d.matcher = self.matcher
            d.tail = tail
            d.img_src, d.img_frm = L.ptr(pipe.synth.img), L.ptr(pipe.img_frm)
            d.img_chunk_bytes = pipe.synth.img[:pipe.chunk].numel() * pipe.synth.img.element_size()
            if pipe.sampler:
                self.lib.cfr_sampler_destroy(pipe.sampler)
            h = C.c_void_p()
            L.check(self.lib.cfr_sampler_create(C.byref(d), C.byref(h)))
            pipe.sampler = h
            tail = h
        self.sampler = self.pipes[0].sampler

    def set_overlap(self, on: bool) -> None:
        """Two-stream overlap of the FRM side with the next group's synthesis (default on); off = strictly serial launches
        (what per-kernel CUDA-event timing needs)."""
        L.check(self.lib.cfr_sampler_set_overlap(self.sampler, int(bool(on))))

    def __del__(self):
        try:
            for pipe in getattr(self, "pipes", []):
                if pipe.sampler:
                    self.lib.cfr_sampler_destroy(pipe.sampler)
                    pipe.sampler = None
            self.sampler = None
            if getattr(self, "matcher", None):
                self.lib.cfr_matcher_destroy(self.matcher)
                self.matcher = None
        except Exception:
            pass

    @property
    def num_classes(self) -> int:
        return self.gallery.shape[0]

    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def embed_latents(self, w: Tensor) -> Tensor:
        """lat2embs (gen_utils.py:108-139): [n,512] W latents -> [n,512] embeddings (device, fp32)."""
        w = _f32(w, self.device)
        out = torch.empty(w.shape[0], 512, device=self.device)
        for i in range(0, w.shape[0], self.chunk):
            b = min(self.chunk, w.shape[0] - i)
            self.synth.out_slot.zero_()
            L.check(self.lib.cfr_truncate(L.ptr(w[i:i + b]), L.ptr(self.synth.w_avg), PSI, b, L.ptr(self.synth.wp2),
                                          self._stream()))
            self.synth.run()
            self.pipes[0].img_frm[:self.chunk].copy_(self.synth.img[:self.chunk])
            self.frm.run()
            out[i:i + b] = self.frm.emb[:b]
        return out

    def sample_votes_sharded(self, shard, z: Tensor, x: Tensor, sigma: Tensor, num: int, seed: int = 0,
                             sample_offset: int = 0, noise: Optional[Tensor] = None, want_pred: bool = False):
        """Smooth._sample_noise body with the gallery sharded over ranks (partition C): every rank
        synthesises and embeds the SAME samples (same seed / offsets), matches them against its own rows
        (`gallery_shard.ShardedGallery`), the 8-byte keys are all-gathered and merged.  Returns (counts [N] int64, pred)."""
        # the engine's own (placeholder) gallery is not consulted: only the embeddings are taken from this call
        _, ex = self.sample_votes(z, x, sigma, num, seed=seed, sample_offset=sample_offset, noise=noise, want_emb=True,
                                  counts=torch.zeros(self.num_classes, dtype=torch.int64, device=self.device))
        counts = torch.zeros(shard.n_total, dtype=torch.int64, device=self.device)
        pred = shard.match_vote(ex["emb"], counts, want_pred=want_pred)
        return counts, pred

    def sample_votes_multi(self, z: Tensor, x: Tensor, sigma: Tensor, nums, seed: int = 0, sample_offsets=None,
                           counts: Optional[Tensor] = None) -> Tensor:
        """Several identities in one call (cfr_sample_votes_multi): identity g draws nums[g] samples around (z[g], x[g])
        at Philox offsets sample_offsets[g]..; consecutive identities share program runs.  Returns counts [G, N] int64."""
        z = _f32(z.reshape(-1, 512), self.device)
        g = z.shape[0]
        x = _f32(x.reshape(-1, N_DIRS), self.device)
        if x.shape[0] == 1 and g > 1:
            x = x.expand(g, N_DIRS).contiguous()
        sigma = _f32(sigma.reshape(-1), self.device)
        nums = [int(v) for v in nums]
        offs = [int(v) for v in (sample_offsets if sample_offsets is not None else [0] * g)]
        if x.shape[0] != g or len(nums) != g or len(offs) != g or sigma.numel() not in (1, N_DIRS):
            raise ValueError("sample_votes_multi: z [G,512], x [G,5] (or [1,5]), nums / sample_offsets of length G")
        if counts is None:
            counts = torch.zeros(g, self.num_classes, dtype=torch.int64, device=self.device)
        num_arr = (C.c_int64 * g)(*nums)
        off_arr = (C.c_uint64 * g)(*offs)
        L.check(self.lib.cfr_sample_votes_multi(self.sampler, g, L.ptr(z), L.ptr(x), L.ptr(sigma), sigma.numel(), num_arr,
                                                seed, off_arr, L.ptr(counts), self._stream()))
        return counts

    def sample_votes(self, z: Tensor, x: Tensor, sigma: Tensor, num: int, seed: int = 0, sample_offset: int = 0,
                     noise: Optional[Tensor] = None, counts: Optional[Tensor] = None, want_pred: bool = False,
                     want_emb: bool = False, want_noise: bool = False):
        """Smooth._sample_noise body (smooth.py:126-137) on device.  Returns (counts int64 [N], extras dict)."""
        z = _f32(z.reshape(-1), self.device)
        x = _f32(x.reshape(-1), self.device)
        sigma = _f32(sigma.reshape(-1), self.device)
        if z.numel() != 512 or x.numel() != N_DIRS or sigma.numel() not in (1, N_DIRS):
            raise ValueError(f"sample_votes: z must hold ONE latent [1,512], x [1,{N_DIRS}], sigma 1 or {N_DIRS} values "
                             f"(got {z.numel()}, {x.numel()}, {sigma.numel()})")
        if noise is not None and noise.numel() != num * N_DIRS:
            raise ValueError(f"sample_votes: noise must be [num, {N_DIRS}]")
        if counts is None:
            counts = torch.zeros(self.num_classes, dtype=torch.int64, device=self.device)
        noise_d = _f32(noise.reshape(-1, 5), self.device) if noise is not None else None
        pred = torch.empty(num, dtype=torch.int32, device=self.device) if want_pred else None
        emb = torch.empty(num, 512, device=self.device) if want_emb else None
        nz = torch.empty(num, 5, device=self.device) if want_noise else None
        L.check(self.lib.cfr_sample_votes(self.sampler, L.ptr(z), L.ptr(x), L.ptr(sigma), sigma.numel(),
                                          L.ptr(noise_d), num, seed, sample_offset, L.ptr(counts), L.ptr(pred),
                                          L.ptr(emb), L.ptr(nz), self._stream()))
        return counts, {"pred": pred, "emb": emb, "noise": nz}
