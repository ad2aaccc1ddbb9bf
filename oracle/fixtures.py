"""Seeded synthetic fixtures for the MC-certification path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

No checkpoints are available offline, so weights of the named architectures are drawn from fixed
seeds.  Two things differ from the reference constructors' own init, both deliberately:

* StyleGAN: ``apply_noise.weight``, every bias and ``w_avg`` are zero in the constructor
  (stylegan_generator_model.py:316,474,525,548) which would make those code paths no-ops; they
  are drawn N(0, 0.1^2) here.  Conv / dense weights are N(0,1) as in the TF original (the
  He scale is applied at run time by WScaleLayer :524).
* iresnet50: ``IResNet.__init__`` draws convs N(0, 0.1) (iresnet.py:102-107) which overflows
  fp32 in eval mode with default BN statistics.  Convs are Kaiming-normal fan-in
  here and the BN running statistics are *calibrated* on generated images so embeddings are
  O(1) and discriminative.

The dicts use the reference's ``state_dict()`` key names and shapes, so
they load into the unmodified reference modules.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import mc_path as M

SD = Dict[str, torch.Tensor]
_CACHE_VERSION = 3


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def stylegan_weights(seed: int = 0) -> SD:
    """State dict for StyleGANGeneratorModel(1024,512,'auto',3,0.7,8,False) -- path keys only
    (truncation + synthesis; ``mapping.*`` is not on the certify path)."""
    g = _gen(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    sd: SD = {}
    sd["truncation.w_avg"] = 0.1 * rn(512)
    coefs = torch.ones(1, M.NUM_LAYERS, 1)
    coefs[:, :8] *= 0.7
    sd["truncation.truncation"] = coefs
    sd["synthesis.lod"] = torch.zeros(())
    sd["synthesis.layer0.first_layer"] = 1.0 + 0.5 * rn(1, 512, 4, 4)
    for layer in range(M.NUM_LAYERS):
        c = M.layer_channels(layer)
        res = M.layer_res(layer)
        p = f"synthesis.layer{layer}."
        if layer >= 1:
            cin = M.layer_channels(layer - 1)
            if layer % 2 == 0 and res >= 128:
                sd[p + "weight"] = rn(3, 3, cin, c)
            else:
                sd[p + "conv.weight"] = rn(c, cin, 3, 3)
        if layer % 2 == 0 and layer >= 2:
            k = torch.tensor([1.0, 2.0, 1.0])
            k = (k[:, None] * k[None, :]) / 16.0
            sd[p + "blur.kernel"] = k.view(1, 1, 3, 3).repeat(c, 1, 1, 1)
        sd[p + "epilogue.apply_noise.noise"] = rn(1, 1, res, res)
        sd[p + "epilogue.apply_noise.weight"] = 0.1 * rn(c)
        sd[p + "epilogue.bias"] = 0.1 * rn(c)
        sd[p + "epilogue.style_mod.dense.linear.weight"] = rn(2 * c, 512)
        sd[p + "epilogue.style_mod.dense.wscale.bias"] = 0.1 * rn(2 * c)
    for k in range(9):
        c = M.layer_channels(2 * k + 1)
        sd[f"synthesis.output{k}.conv.weight"] = rn(3, c, 1, 1)
        sd[f"synthesis.output{k}.bias"] = 0.1 * rn(3)
    return sd


def iresnet50_weights(seed: int = 1) -> SD:
    """State dict for iresnet50(False, fp16=False) with *uncalibrated* BN statistics (mean 0, var 1)."""
    g = _gen(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    ru = lambda *s: torch.rand(*s, generator=g)
    sd: SD = {}

    def conv(name, cout, cin, k):
        sd[name + ".weight"] = rn(cout, cin, k, k) * math.sqrt(1.0 / (cin * k * k))

    def bn(name, c):
        sd[name + ".weight"] = 0.8 + 0.4 * ru(c)
        sd[name + ".bias"] = 0.1 * rn(c)
        sd[name + ".running_mean"] = torch.zeros(c)
        sd[name + ".running_var"] = torch.ones(c)
        sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    conv("conv1", 64, 3, 3)
    bn("bn1", 64)
    sd["prelu.weight"] = 0.1 + 0.3 * ru(64)
    inplanes = 64
    for li, (nblocks, planes) in enumerate(zip(M.IRESNET50_LAYERS, M.IRESNET50_PLANES), start=1):
        for bi in range(nblocks):
            p = f"layer{li}.{bi}."
            bn(p + "bn1", inplanes)
            conv(p + "conv1", planes, inplanes, 3)
            bn(p + "bn2", planes)
            sd[p + "prelu.weight"] = 0.1 + 0.3 * ru(planes)
            conv(p + "conv2", planes, planes, 3)
            bn(p + "bn3", planes)
            if bi == 0:
                conv(p + "downsample.0", planes, inplanes, 1)
                bn(p + "downsample.1", planes)
            inplanes = planes
    bn("bn2", 512)
    sd["fc.weight"] = rn(512, 512 * 49) * math.sqrt(1.0 / (512 * 49))
    sd["fc.bias"] = 0.1 * rn(512)
    bn("features", 512)
    sd["features.weight"] = torch.ones(512)          # iresnet.py:99-100 (frozen at 1.0)
    return sd


def calibrate_bn(f_sd: SD, images112: torch.Tensor) -> SD:
    """One train-mode pass (momentum=None => running stats := this batch's mean / unbiased var) over
    ``images112`` [n,3,112,112], visiting the BatchNorms in forward order; returns a new dict."""
    sd = dict(f_sd)

    def bn_cal(x, p):
        dims = [0, 2, 3] if x.dim() == 4 else [0]
        mean = x.mean(dim=dims)
        var_b = x.var(dim=dims, unbiased=False)
        var_u = x.var(dim=dims, unbiased=True)
        sd[p + ".running_mean"] = mean.clone()
        sd[p + ".running_var"] = var_u.clone()
        sd[p + ".num_batches_tracked"] = torch.tensor(1, dtype=torch.long)
        shape = (1, -1, 1, 1) if x.dim() == 4 else (1, -1)
        return (x - mean.view(shape)) / torch.sqrt(var_b.view(shape) + 1e-5) * sd[p + ".weight"].view(shape) \
            + sd[p + ".bias"].view(shape)

    with torch.no_grad():
        x = F.conv2d(images112, sd["conv1.weight"], padding=1)
        x = M._prelu(bn_cal(x, "bn1"), sd["prelu.weight"])
        for li, nblocks in enumerate(M.IRESNET50_LAYERS, start=1):
            for bi in range(nblocks):
                p = f"layer{li}.{bi}."
                stride = 2 if bi == 0 else 1
                identity = x
                out = bn_cal(x, p + "bn1")
                out = F.conv2d(out, sd[p + "conv1.weight"], padding=1)
                out = M._prelu(bn_cal(out, p + "bn2"), sd[p + "prelu.weight"])
                out = F.conv2d(out, sd[p + "conv2.weight"], stride=stride, padding=1)
                out = bn_cal(out, p + "bn3")
                if bi == 0:
                    identity = bn_cal(F.conv2d(x, sd[p + "downsample.0.weight"], stride=stride), p + "downsample.1")
                x = out + identity
        x = bn_cal(x, "bn2")
        x = torch.flatten(x, 1)
        x = F.linear(x, sd["fc.weight"], sd["fc.bias"])
        bn_cal(x, "features")
    return sd


def latents(n: int, seed: int = 2) -> np.ndarray:
    """w.npy stand-in: RandomState(seed).randn(n,512) float32 (gen_utils.py:44-46 loads [N,512])."""
    return np.random.RandomState(seed).randn(n, 512).astype(np.float32)


def synthetic_dirs(seed: int = 5) -> np.ndarray:
    """[5,512] float32 unit-norm rows -- stand-in for the five InterFaceGAN boundaries when the
    real ones (tests/golden/dirs.npy, extracted from the reference's boundaries/*.npy) are not
    wanted."""
    d = np.random.RandomState(seed).randn(5, 512)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return d.astype(np.float32)


def build_models(cache_dir: str | None = None, n_calib: int = 16) -> Tuple[SD, SD]:
    """(stylegan_sd, calibrated iresnet50_sd).  Calibration synthesises ``n_calib`` fixture latents
    on the CPU (about 1 s each); the result is cached as a .pt under ``cache_dir`` when given."""
    g_sd = stylegan_weights()
    path = os.path.join(cache_dir, f"iresnet50_calibrated_v{_CACHE_VERSION}.pt") if cache_dir else None
    if path and os.path.isfile(path):
        return g_sd, torch.load(path)
    f_sd = iresnet50_weights()
    w = torch.from_numpy(latents(n_calib))
    ims = []
    with torch.no_grad():
        for i in range(0, n_calib, M.GAN_CHUNK):
            ims.append(M.transform(M.easy_synthesize(w[i:i + M.GAN_CHUNK], g_sd, literal=False)))
    f_sd = calibrate_bn(f_sd, torch.cat(ims))
    if path:
        os.makedirs(cache_dir, exist_ok=True)
        torch.save(f_sd, path)
    return g_sd, f_sd


def synthetic_gallery(true_rows: torch.Tensor, n: int, seed: int = 3) -> torch.Tensor:
    """Gallery [n,512]: the given true/decoy rows first, remaining rows Gaussian with the per-dimension
    mean/std of the true rows."""
    k = true_rows.shape[0]
    if n <= k:
        return true_rows[:n].clone()
    mu = true_rows.mean(0, keepdim=True)
    sdv = true_rows.std(0, keepdim=True) if k > 1 else torch.ones_like(mu)
    rs = np.random.RandomState(seed)
    rest = torch.from_numpy(rs.randn(n - k, 512).astype(np.float32)) * sdv + mu
    return torch.cat([true_rows, rest], dim=0)
