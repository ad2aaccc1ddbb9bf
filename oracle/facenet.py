"""CPU restatement (torch fp32) of facenet_pytorch.InceptionResnetV1, the FRM of BASELINE config 4.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

PARITY UNPINNED.  The reference imports the network from the third-party package ``facenet_pytorch``
(timesler/facenet-pytorch; unpinned: README.md:54 `pip install facenet_pytorch`; call sites main_attack.py:21,126-129),
which is neither vendored under /root/reference nor installed here, and its pretrained weights are downloaded at run
time.  There is therefore nothing to run this restatement against: it follows the published architecture
(Szegedy et al., Inception-ResNet-v1 as implemented by that package, summarised in SURVEY.md section 8c):

    stem   conv 3x3/2 3->32, conv 3x3 32->32, conv 3x3 p1 32->64, maxpool 3/2, conv 1x1 64->80, conv 3x3 80->192,
           conv 3x3/2 192->256                                                         (160^2 -> 17^2 x 256)
    5  x Block35 (scale 0.17): branches {1x1 32 | 1x1 32, 3x3 32 | 1x1 32, 3x3 32, 3x3 32} -> cat 96 -> 1x1 256 (+bias)
    Mixed_6a: {3x3/2 384 | 1x1 192, 3x3 192, 3x3/2 256 | maxpool 3/2}                    (-> 8^2 x 896)
    10 x Block17 (scale 0.10): {1x1 128 | 1x1 128, 1x7 128, 7x1 128} -> cat 256 -> 1x1 896 (+bias)
    Mixed_7a: {1x1 256, 3x3/2 384 | 1x1 256, 3x3/2 256 | 1x1 256, 3x3 256, 3x3/2 256 | maxpool 3/2}   (-> 3^2 x 1792)
    5  x Block8 (scale 0.20) + 1 x Block8 (scale 1, no ReLU): {1x1 192 | 1x1 192, 1x3 192, 3x1 192} -> cat 384 -> 1x1 1792
    global average pool, Linear(1792, 512, bias=False), BatchNorm1d(eps 1e-3), L2 normalisation.
Every "conv" above except the block-closing 1x1 is conv(no bias) + BatchNorm(eps 1e-3) + ReLU; a block returns
relu(x + scale * closing_conv(cat(branches))).  State-dict keys follow that package's module names."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
BN_EPS = 1e-3
INPUT_RES = 160                       # gen_utils.py:17-21 INP_RESOLS['facenet*']

# (name, cin, cout, (kh, kw), stride, (ph, pw)) of every BasicConv2d, in forward order inside its parent
STEM = (("conv2d_1a", 3, 32, (3, 3), 2, (0, 0)), ("conv2d_2a", 32, 32, (3, 3), 1, (0, 0)),
        ("conv2d_2b", 32, 64, (3, 3), 1, (1, 1)), ("conv2d_3b", 64, 80, (1, 1), 1, (0, 0)),
        ("conv2d_4a", 80, 192, (3, 3), 1, (0, 0)), ("conv2d_4b", 192, 256, (3, 3), 2, (0, 0)))
BLOCK35 = {"branch0": ((256, 32, (1, 1), 1, (0, 0)),),
           "branch1": ((256, 32, (1, 1), 1, (0, 0)), (32, 32, (3, 3), 1, (1, 1))),
           "branch2": ((256, 32, (1, 1), 1, (0, 0)), (32, 32, (3, 3), 1, (1, 1)), (32, 32, (3, 3), 1, (1, 1)))}
BLOCK17 = {"branch0": ((896, 128, (1, 1), 1, (0, 0)),),
           "branch1": ((896, 128, (1, 1), 1, (0, 0)), (128, 128, (1, 7), 1, (0, 3)), (128, 128, (7, 1), 1, (3, 0)))}
BLOCK8 = {"branch0": ((1792, 192, (1, 1), 1, (0, 0)),),
          "branch1": ((1792, 192, (1, 1), 1, (0, 0)), (192, 192, (1, 3), 1, (0, 1)), (192, 192, (3, 1), 1, (1, 0)))}
MIXED_6A = {"branch0": ((256, 384, (3, 3), 2, (0, 0)),),
            "branch1": ((256, 192, (1, 1), 1, (0, 0)), (192, 192, (3, 3), 1, (1, 1)), (192, 256, (3, 3), 2, (0, 0)))}
MIXED_7A = {"branch0": ((896, 256, (1, 1), 1, (0, 0)), (256, 384, (3, 3), 2, (0, 0))),
            "branch1": ((896, 256, (1, 1), 1, (0, 0)), (256, 256, (3, 3), 2, (0, 0))),
            "branch2": ((896, 256, (1, 1), 1, (0, 0)), (256, 256, (3, 3), 1, (1, 1)), (256, 256, (3, 3), 2, (0, 0)))}
BLOCKS = (("repeat_1", 5, BLOCK35, 256, 0.17), ("repeat_2", 10, BLOCK17, 896, 0.10), ("repeat_3", 5, BLOCK8, 1792, 0.20))


def conv_names(branches: dict, prefix: str):
    """(state-dict prefix, spec) of every BasicConv2d of a branch dict; single-conv branches have no index."""
    for bname, convs in branches.items():
        for i, spec in enumerate(convs):
            yield (f"{prefix}{bname}." if len(convs) == 1 else f"{prefix}{bname}.{i}."), spec


class _Calib:
    """Optional BatchNorm calibration: replace the running statistics by those of the batch being pushed through."""

    def __init__(self, sd: SD, on: bool):
        self.sd, self.on = sd, on

    def stats(self, x: torch.Tensor, p: str):
        if self.on:
            dims = [0, 2, 3] if x.dim() == 4 else [0]
            self.sd[p + "running_mean"] = x.mean(dim=dims)
            self.sd[p + "running_var"] = x.var(dim=dims, unbiased=True)
        return self.sd[p + "running_mean"], self.sd[p + "running_var"]


def _basic(x, sd, p, spec, cal: _Calib):
    _, _, _, stride, pad = spec
    x = F.conv2d(x, sd[p + "conv.weight"], stride=stride, padding=pad)
    mean, var = cal.stats(x, p + "bn.")
    x = F.batch_norm(x, mean, var, sd[p + "bn.weight"], sd[p + "bn.bias"], False, 0.0, BN_EPS)
    return F.relu(x)


def _branches(x, sd, prefix, branches, cal):
    outs = []
    for bname, convs in branches.items():
        y = x
        for i, spec in enumerate(convs):
            y = _basic(y, sd, f"{prefix}{bname}." if len(convs) == 1 else f"{prefix}{bname}.{i}.", spec, cal)
        outs.append(y)
    return outs


def _block(x, sd, prefix, branches, scale, relu, cal):
    out = torch.cat(_branches(x, sd, prefix, branches, cal), 1)
    out = F.conv2d(out, sd[prefix + "conv2d.weight"], sd[prefix + "conv2d.bias"])
    out = out * scale + x
    return F.relu(out) if relu else out


def forward(x: torch.Tensor, sd: SD, calibrate: bool = False) -> torch.Tensor:
    """[B,3,160,160] (normalised to [-1,1]) -> [B,512] L2-normalised embeddings.  ``calibrate`` rewrites the BatchNorm
    running statistics in ``sd`` from this batch (used once to give the synthetic fixture sane activations)."""
    cal = _Calib(sd, calibrate)
    for name, *spec in STEM[:3]:
        x = _basic(x, sd, name + ".", tuple(spec), cal)
    x = F.max_pool2d(x, 3, stride=2)
    for name, *spec in STEM[3:]:
        x = _basic(x, sd, name + ".", tuple(spec), cal)
    for i in range(5):
        x = _block(x, sd, f"repeat_1.{i}.", BLOCK35, 0.17, True, cal)
    x = torch.cat(_branches(x, sd, "mixed_6a.", MIXED_6A, cal) + [F.max_pool2d(x, 3, stride=2)], 1)
    for i in range(10):
        x = _block(x, sd, f"repeat_2.{i}.", BLOCK17, 0.10, True, cal)
    x = torch.cat(_branches(x, sd, "mixed_7a.", MIXED_7A, cal) + [F.max_pool2d(x, 3, stride=2)], 1)
    for i in range(5):
        x = _block(x, sd, f"repeat_3.{i}.", BLOCK8, 0.20, True, cal)
    x = _block(x, sd, "block8.", BLOCK8, 1.0, False, cal)
    x = x.mean(dim=(2, 3))
    x = F.linear(x, sd["last_linear.weight"])
    mean, var = cal.stats(x, "last_bn.")
    x = F.batch_norm(x, mean, var, sd["last_bn.weight"], sd["last_bn.bias"], False, 0.0, BN_EPS)
    return F.normalize(x, p=2, dim=1)
