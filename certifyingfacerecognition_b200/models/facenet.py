"""FaceNet FRM of BASELINE config 4: facenet_pytorch.InceptionResnetV1 (main_attack.py:21,126-129) as a recorded
program of tcgen05 implicit-GEMM convolutions (every BasicConv2d = conv + folded BatchNorm(eps 1e-3) + ReLU epilogue),
torch.cat realised by writing each branch into its channel slice of one NHWC buffer, the residual blocks'
`relu(x + scale * conv1x1(cat) )` as one conv with folded scale, residual add and post-add ReLU in the epilogue.

PARITY UNPINNED: the package is a third-party dependency that is neither vendored in the reference nor installed
here; the layer table below restates its published architecture and is checked against
`oracle/facenet.py` (an independent torch restatement of the same description), not against the package itself."""
from __future__ import annotations

import math
from typing import Dict, Iterator, List, Tuple

import torch

from .. import _lib as L
from ..engine import Program, _f16, _f32, pack_conv_weight, tile_for

Tensor = torch.Tensor
BN_EPS = 1e-3
INPUT_RES = 160
Spec = Tuple[int, int, Tuple[int, int], int, Tuple[int, int]]      # cin, cout, (kh, kw), stride, (pad_h, pad_w)

STEM: Tuple[Tuple[str, Spec], ...] = (
    ("conv2d_1a.", (3, 32, (3, 3), 2, (0, 0))), ("conv2d_2a.", (32, 32, (3, 3), 1, (0, 0))),
    ("conv2d_2b.", (32, 64, (3, 3), 1, (1, 1))), ("conv2d_3b.", (64, 80, (1, 1), 1, (0, 0))),
    ("conv2d_4a.", (80, 192, (3, 3), 1, (0, 0))), ("conv2d_4b.", (192, 256, (3, 3), 2, (0, 0))))
BLOCK35 = {"branch0": [(256, 32, (1, 1), 1, (0, 0))],
           "branch1": [(256, 32, (1, 1), 1, (0, 0)), (32, 32, (3, 3), 1, (1, 1))],
           "branch2": [(256, 32, (1, 1), 1, (0, 0)), (32, 32, (3, 3), 1, (1, 1)), (32, 32, (3, 3), 1, (1, 1))]}
BLOCK17 = {"branch0": [(896, 128, (1, 1), 1, (0, 0))],
           "branch1": [(896, 128, (1, 1), 1, (0, 0)), (128, 128, (1, 7), 1, (0, 3)), (128, 128, (7, 1), 1, (3, 0))]}
BLOCK8 = {"branch0": [(1792, 192, (1, 1), 1, (0, 0))],
          "branch1": [(1792, 192, (1, 1), 1, (0, 0)), (192, 192, (1, 3), 1, (0, 1)), (192, 192, (3, 1), 1, (1, 0))]}
MIXED_6A = {"branch0": [(256, 384, (3, 3), 2, (0, 0))],
            "branch1": [(256, 192, (1, 1), 1, (0, 0)), (192, 192, (3, 3), 1, (1, 1)), (192, 256, (3, 3), 2, (0, 0))]}
MIXED_7A = {"branch0": [(896, 256, (1, 1), 1, (0, 0)), (256, 384, (3, 3), 2, (0, 0))],
            "branch1": [(896, 256, (1, 1), 1, (0, 0)), (256, 256, (3, 3), 2, (0, 0))],
            "branch2": [(896, 256, (1, 1), 1, (0, 0)), (256, 256, (3, 3), 1, (1, 1)), (256, 256, (3, 3), 2, (0, 0))]}
# (module name, repeats, channels of the concatenated branches, block width, residual scale)
BLOCKS = (("repeat_1", 5, 96, 256, 0.17), ("repeat_2", 10, 256, 896, 0.10), ("repeat_3", 5, 384, 1792, 0.20))
_BRANCHES = {"repeat_1": BLOCK35, "repeat_2": BLOCK17, "repeat_3": BLOCK8, "block8": BLOCK8}


def _branch_convs(branches: dict, prefix: str) -> Iterator[Tuple[str, Spec]]:
    for bname, convs in branches.items():
        for i, spec in enumerate(convs):
            yield (f"{prefix}{bname}." if len(convs) == 1 else f"{prefix}{bname}.{i}."), spec


def all_basic_convs() -> List[Tuple[str, Spec]]:
    """(state-dict prefix, spec) of every conv + BatchNorm + ReLU unit, in module order."""
    out = list(STEM)
    for name, reps, _, _, _ in BLOCKS[:1]:
        for i in range(reps):
            out += list(_branch_convs(_BRANCHES[name], f"{name}.{i}."))
    out += list(_branch_convs(MIXED_6A, "mixed_6a."))
    for i in range(10):
        out += list(_branch_convs(BLOCK17, f"repeat_2.{i}."))
    out += list(_branch_convs(MIXED_7A, "mixed_7a."))
    for i in range(5):
        out += list(_branch_convs(BLOCK8, f"repeat_3.{i}."))
    out += list(_branch_convs(BLOCK8, "block8."))
    return out


def _pad_c(c: int) -> int:
    """Channel count a buffer is allocated with: a multiple of 16, and of 64 above 64 (igemm K blocks)."""
    return (c + 15) // 16 * 16 if c <= 64 else (c + 63) // 64 * 64


class FaceNetProgram(Program):
    """img [chunk,160,160,16] fp16 NHWC (3 live channels, normalised to [-1,1]) -> emb [chunk,512] fp32, L2-normalised."""

    def __init__(self, f_sd: Dict[str, Tensor], chunk: int, img: Tensor, device="cuda"):
        super().__init__()
        self.dev = torch.device(device)
        self.sd = {k: v.detach().float().cpu() for k, v in f_sd.items()}
        self.chunk = n = chunk
        x, h, c = img, INPUT_RES, 16
        for name, spec in STEM[:3]:
            x, h, c = self._basic(x, h, c, name, spec)
        x, h, c = self._maxpool(x, h, c, None, 0)
        for name, spec in STEM[3:]:
            x, h, c = self._basic(x, h, c, name, spec)
        for i in range(5):
            x = self._block(x, h, f"repeat_1.{i}.", BLOCK35, 96, 256, 0.17, True)
        x, h, c = self._mixed(x, h, 256, "mixed_6a.", MIXED_6A, 896)
        for i in range(10):
            x = self._block(x, h, f"repeat_2.{i}.", BLOCK17, 256, 896, 0.10, True)
        x, h, c = self._mixed(x, h, 896, "mixed_7a.", MIXED_7A, 1792)
        for i in range(5):
            x = self._block(x, h, f"repeat_3.{i}.", BLOCK8, 384, 1792, 0.20, True)
        x = self._block(x, h, "block8.", BLOCK8, 384, 1792, 1.0, False)
        # AdaptiveAvgPool2d(1) -> Linear(1792, 512, bias=False) -> BatchNorm1d (folded into the GEMM) -> F.normalize
        pooled = self.hold(torch.zeros(n, 1792, dtype=torch.float16, device=self.dev))
        L.check(self.lib.cfr_program_add_avgpool(self.handle, L.ptr(x), n, h * h, 1792, L.ptr(pooled)))
        sd = self.sd
        s = sd["last_bn.weight"] / torch.sqrt(sd["last_bn.running_var"] + BN_EPS)
        t = sd["last_bn.bias"] - sd["last_bn.running_mean"] * s
        wfc = pack_conv_weight((sd["last_linear.weight"] * s.view(-1, 1)).view(512, 1792, 1, 1))
        raw = self.hold(torch.zeros(n, 512, device=self.dev))
        self.conv(inp=pooled, n=n, hin=1, win=1, cin=1792, w=self.hold(_f16(wfc, self.dev)), cout=512, hout=1, wout=1,
                  tile=(1, 1, 128), out=raw, out_hwc=(1, 1, 512), taps=[[(0, 0)]], bias=self.hold(_f32(t, self.dev)))
        self.emb = self.hold(torch.zeros(n, 512, device=self.dev))
        L.check(self.lib.cfr_program_add_l2norm(self.handle, L.ptr(raw), n, 512, L.ptr(self.emb)))

    # ---- building blocks ----------------------------------------------------------------------------------------
    def _new(self, h: int, c: int) -> Tensor:
        return self.hold(torch.zeros(self.chunk * h * h * c, dtype=torch.float16, device=self.dev))

    def _basic(self, x: Tensor, h: int, c_buf: int, p: str, spec: Spec, out: Tensor = None, out_c: int = 0, c_off: int = 0):
        """conv(no bias) + BatchNorm + ReLU.  `x` has c_buf >= cin channels (extra ones are zero); the result goes to a
        fresh buffer, or into channels [c_off, c_off + cout) of `out` (a torch.cat target with out_c channels)."""
        cin, cout, (kh, kw), stride, (ph, pw) = spec
        sd = self.sd
        s = sd[p + "bn.weight"] / torch.sqrt(sd[p + "bn.running_var"] + BN_EPS)
        t = sd[p + "bn.bias"] - sd[p + "bn.running_mean"] * s
        w = sd[p + "conv.weight"] * s.view(-1, 1, 1, 1)
        ho = (h + 2 * ph - kh) // stride + 1
        assert ho == (h + 2 * pw - kw) // stride + 1
        cout_buf = _pad_c(cout) if out is None else cout
        if cout_buf != cout:                                   # zero rows: padded channels come out as relu(0) = 0
            w = torch.cat([w, torch.zeros(cout_buf - cout, *w.shape[1:])])
            t = torch.cat([t, torch.zeros(cout_buf - cout)])
        taps = [[(ky - ph, kx - pw) for ky in range(kh) for kx in range(kw)]]
        if out is None:
            out, out_c, c_off = self._new(ho, cout_buf), cout_buf, 0
        self.conv(inp=x, n=self.chunk, hin=h, win=h, cin=c_buf, w=self.hold(_f16(pack_conv_weight(w, cin_pad=c_buf), self.dev)),
                  cout=cout_buf, hout=ho, wout=ho, tile=tile_for(ho, self.chunk), out=out[c_off:], out_hwc=(ho, ho, out_c),
                  taps=taps, stride=stride, bias=self.hold(_f32(t, self.dev)), act=L.ACT_LRELU, slope=0.0)
        return out, ho, out_c

    def _maxpool(self, x: Tensor, h: int, c: int, out: Tensor, c_off: int, out_c: int = 0):
        ho = (h - 3) // 2 + 1
        if out is None:
            out, out_c = self._new(ho, c), c
        L.check(self.lib.cfr_program_add_maxpool3s2(self.handle, L.ptr(x), self.chunk, h, h, c, L.ptr(out), out_c, c_off))
        self.keep.append(x)
        return out, ho, out_c

    def _run_branches(self, x: Tensor, h: int, c: int, prefix: str, branches: dict, cat: Tensor, cat_c: int) -> Tuple[int, int]:
        """Every branch's last conv writes its slice of `cat`; returns (output size, channels written)."""
        off, ho = 0, h
        for bname, convs in branches.items():
            y, hy, cy = x, h, c
            for i, spec in enumerate(convs):
                p = f"{prefix}{bname}." if len(convs) == 1 else f"{prefix}{bname}.{i}."
                if i == len(convs) - 1:
                    _, ho, _ = self._basic(y, hy, cy, p, spec, out=cat, out_c=cat_c, c_off=off)
                    off += spec[1]
                else:
                    y, hy, cy = self._basic(y, hy, cy, p, spec)
        return ho, off

    def _block(self, x: Tensor, h: int, prefix: str, branches: dict, ccat: int, width: int, scale: float, relu: bool) -> Tensor:
        """relu(x + scale * (conv1x1(cat(branches)) + bias))  (Block35 / Block17 / Block8)."""
        cat_c = _pad_c(ccat)
        cat = self._new(h, cat_c)                              # padded channels are never written: stay zero
        self._run_branches(x, h, width, prefix, branches, cat, cat_c)
        sd = self.sd
        w = sd[prefix + "conv2d.weight"] * scale
        b = sd[prefix + "conv2d.bias"] * scale
        out = self._new(h, width)
        self.conv(inp=cat, n=self.chunk, hin=h, win=h, cin=cat_c, w=self.hold(_f16(pack_conv_weight(w, cin_pad=cat_c), self.dev)),
                  cout=width, hout=h, wout=h, tile=tile_for(h, self.chunk), out=out, out_hwc=(h, h, width), taps=[[(0, 0)]],
                  bias=self.hold(_f32(b, self.dev)), resid=x, resid_c=width,
                  act=L.ACT_RELU_POST if relu else L.ACT_NONE)
        return out

    def _mixed(self, x: Tensor, h: int, c: int, prefix: str, branches: dict, cout: int):
        """Mixed_6a / Mixed_7a: strided conv branches + MaxPool2d(3, 2) of the input, concatenated."""
        ho = (h - 3) // 2 + 1
        cat = self._new(ho, cout)
        _, off = self._run_branches(x, h, c, prefix, branches, cat, cout)
        assert off + c == cout
        self._maxpool(x, h, c, cat, off, cout)
        return cat, ho, cout
