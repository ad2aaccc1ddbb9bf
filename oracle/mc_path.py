"""CPU restatement (torch fp32) of the reference's MC-certification hot path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Every function cites
the file:line under ``/root/reference`` whose arithmetic it restates.  Weights
are plain ``dict[str, Tensor]`` keyed exactly like the reference's
``state_dict()``, so the very same dict can be loaded
into the unmodified reference modules (``oracle/make_golden.py`` does that to
pin this file).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from scipy.stats import beta as _beta
from scipy.stats import binomtest as _binomtest
from scipy.stats import norm as _norm

Tensor = torch.Tensor
SD = Dict[str, Tensor]

NUM_LAYERS = 18                     # stylegan_generator_model.py:309  (log2(1024)*2-2)
W_DIM = 512
EMB_SIZE = 512                      # gen_utils.py:24
GAN_CHUNK = 4                       # model_settings.py:72  MAX_IMAGES_ON_DEVICE
ABSTAIN = -1                        # smooth.py:19
# proj_utils.py:16-21 (ATTRS, in OrderedDict order)
ATTR_NAMES = ("age", "eyeglasses", "gender", "pose", "smile")
ATTR_EPS = (0.5, 0.5, 0.2, 0.5, 0.8)


def layer_channels(layer: int) -> int:
    """stylegan_generator_model.py:23-32 -- channels of synthesis layer L at 1024."""
    res_channels = [512, 512, 512, 512, 512, 256, 128, 64, 32, 16]
    return res_channels[layer // 2 + 1] if layer >= 2 else 512


def layer_res(layer: int) -> int:
    """stylegan_generator_model.py:472 -- res = 2**(layer_idx//2+2)."""
    return 2 ** (layer // 2 + 2)


# --------------------------------------------------------------------------- #
# StyleGAN v1 mapping network (generate_data.py path)     #
# --------------------------------------------------------------------------- #
def preprocess_z(z: Tensor) -> Tensor:
    """ModStyleGANGenerator.preprocess, latent_space_type 'Z' (mod_stylegan_generator.py:177-180):
    z / ||z|| * sqrt(512)."""
    z = z.reshape(-1, W_DIM)
    return z / z.norm(dim=1, keepdim=True) * math.sqrt(W_DIM)


def mapping(z: Tensor, sd: SD) -> Tensor:
    """MappingModule.forward, stylegan_generator_model.py:265-295: PixelNormLayer (:398-406, eps 1e-8) then 8 x
    DenseBlock (:765-815) = Linear(no bias) -> WScaleLayer (:508-535: x*gain/sqrt(fan_in)*lr_mul + b*lr_mul with
    gain sqrt(2), lr_mul 0.01) -> LeakyReLU(0.2).   [B,512] -> [B,512]."""
    x = z / torch.sqrt(torch.mean(z * z, dim=1, keepdim=True) + 1e-8)
    lr_mul = 0.01
    scale = math.sqrt(2.0) / math.sqrt(W_DIM) * lr_mul
    for i in range(8):
        x = F.linear(x, sd[f"mapping.dense{i}.linear.weight"])
        x = x * scale + sd[f"mapping.dense{i}.wscale.bias"].view(1, -1) * lr_mul
        x = F.leaky_relu(x, 0.2)
    return x


# --------------------------------------------------------------------------- #
# StyleGAN v1 synthesis                                                        #
# --------------------------------------------------------------------------- #
def truncation(w: Tensor, sd: SD, psi: float = 0.7, trunc_layers: int = 8) -> Tensor:
    """TruncationModule.forward, stylegan_generator_model.py:322-328.

    [B,512] -> [B,18,512];  w_avg + (w - w_avg) * coef, coef = psi for layer <
    trunc_layers else 1 (:317-320; psi/layers from model_settings.py:65-66).
    """
    coefs = torch.ones(1, NUM_LAYERS, 1, dtype=w.dtype)
    coefs[:, :trunc_layers] *= psi
    wp = w.view(-1, 1, W_DIM).repeat(1, NUM_LAYERS, 1)
    w_avg = sd["truncation.w_avg"].view(1, 1, W_DIM)
    return w_avg + (wp - w_avg) * coefs


def _blur(x: Tensor) -> Tensor:
    """BlurLayer.forward :441-463 -- depthwise [1,2,1]x[1,2,1]/16, zero pad 1."""
    c = x.shape[1]
    k = torch.tensor([1.0, 2.0, 1.0], dtype=x.dtype)
    k = (k[:, None] * k[None, :])
    k = (k / k.sum()).view(1, 1, 3, 3).repeat(c, 1, 1, 1)
    return F.conv2d(x, k, stride=1, padding=1, groups=c)


def _epilogue(x: Tensor, w_l: Tensor, sd: SD, layer: int) -> Tensor:
    """EpilogueBlock.forward :559-565 (noise :484, bias, lrelu .2, IN :420-422, style :503-505)."""
    p = f"synthesis.layer{layer}.epilogue."
    c = x.shape[1]
    x = x + sd[p + "apply_noise.noise"] * sd[p + "apply_noise.weight"].view(1, -1, 1, 1)
    x = x + sd[p + "bias"].view(1, -1, 1, 1)
    x = F.leaky_relu(x, 0.2)
    x = x - x.mean(dim=[2, 3], keepdim=True)
    x = x / torch.sqrt((x * x).mean(dim=[2, 3], keepdim=True) + 1e-8)
    # DenseBlock :811-815 with WScaleLayer :529-531: gain 1, lr_mult 1, fan_in 512
    style = F.linear(w_l, sd[p + "style_mod.dense.linear.weight"]) * (1.0 / math.sqrt(W_DIM)) \
        + sd[p + "style_mod.dense.wscale.bias"].view(1, -1)
    style = style.view(-1, 2, c, 1, 1)
    return x * (style[:, 0] + 1) + style[:, 1]


def upconv_equiv_weight(sd: SD, layer: int) -> Tuple[Tensor, bool]:
    """Return (weight [Cout,Cin,3,3] already multiplied by wscale, fused?) for an UpConvBlock
    such that   out = conv2d(nearest_x2(x), weight, padding=1).

    Non-fused (res < 128, :673-675): weight = conv.weight * scale.
    Fused (res >= 128, :667-672): conv_transpose2d(stride 2, pad 1) with the 4x4 box-summed kernel
    is exactly nearest-x2 followed by a 3x3 conv with the spatially flipped kernel
    W'[co,ci,a,b] = weight[2-a,2-b,ci,co]*scale (checked in
    tests/test_oracle_golden.py against F.conv_transpose2d).
    """
    res = layer_res(layer)
    p = f"synthesis.layer{layer}."
    if res >= 128:
        w = sd[p + "weight"]                       # [3,3,Cin,Cout]
        cin = w.shape[2]
        scale = math.sqrt(2.0) / math.sqrt(cin * 9)
        return torch.flip(w, dims=[0, 1]).permute(3, 2, 0, 1).contiguous() * scale, True
    w = sd[p + "conv.weight"]                      # [Cout,Cin,3,3]
    scale = math.sqrt(2.0) / math.sqrt(w.shape[1] * 9)
    return w * scale, False


def _upconv(x: Tensor, sd: SD, layer: int, literal: bool) -> Tensor:
    """UpConvBlock.forward :665-676 (up to, not including, the epilogue)."""
    p = f"synthesis.layer{layer}."
    if layer_res(layer) >= 128 and literal:
        w = sd[p + "weight"]
        scale = math.sqrt(2.0) / math.sqrt(w.shape[2] * 9)
        kernel = F.pad(w * scale, (0, 0, 0, 0, 1, 1, 1, 1), "constant", 0.0)
        kernel = kernel[1:, 1:] + kernel[:-1, 1:] + kernel[1:, :-1] + kernel[:-1, :-1]
        x = F.conv_transpose2d(x, kernel.permute(2, 3, 0, 1), stride=2, padding=1)
    else:
        weq, _ = upconv_equiv_weight(sd, layer)
        x = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), weq, padding=1)
    return _blur(x)


def synthesis(wp: Tensor, sd: SD, literal: bool = True,
              tap: Optional[Callable[[str, Tensor], None]] = None) -> Tensor:
    """SynthesisModule.forward :380-395 at lod=0 -> raw image [B,3,1024,1024].

    Only ``output8`` is live at lod 0 (:392 overwrites the other eight), and layer0 being
    evaluated twice (:382,:387) has no effect on the result, so neither is repeated here.
    ``tap(name, tensor)`` observes each layer output (used for golden statistics).
    """
    b = wp.shape[0]
    x = sd["synthesis.layer0.first_layer"].repeat(b, 1, 1, 1)       # FirstConvBlock :581-584
    x = _epilogue(x, wp[:, 0], sd, 0)
    if tap:
        tap("layer0", x)
    for layer in range(1, NUM_LAYERS):
        if layer % 2 == 0:
            x = _upconv(x, sd, layer, literal)
        else:                                                       # ConvBlock :738-741
            w = sd[f"synthesis.layer{layer}.conv.weight"]
            x = F.conv2d(x, w, padding=1) * (math.sqrt(2.0) / math.sqrt(w.shape[1] * 9))
        x = _epilogue(x, wp[:, layer], sd, layer)
        if tap:
            tap(f"layer{layer}", x)
    w = sd["synthesis.output8.conv.weight"]                          # LastConvBlock :759-762
    img = F.conv2d(x, w) * (1.0 / math.sqrt(w.shape[1])) + sd["synthesis.output8.bias"].view(1, -1, 1, 1)
    return img


def postprocess(img: Tensor) -> Tensor:
    """ModStyleGANGenerator.postprocess mod_stylegan_generator.py:303-307 (min -1, max 1)."""
    return torch.clamp((img + 1.0) / 2.0 + 0.5 / 255, 0, 1)


def transform(img: Tensor, size: int = 112, mean: float = 0.5, std: float = 0.5) -> Tensor:
    """get_transform gen_utils.py:77-85 -- bilinear (align_corners=False, no antialias) + Normalize."""
    x = F.interpolate(img, size=(size, size), mode="bilinear", align_corners=False)
    return (x - mean) / std


def easy_synthesize(w: Tensor, sd: SD, literal: bool = True) -> Tensor:
    """ModStyleGANGenerator.easy_synthesize (W branch) mod_stylegan_generator.py:242-255,280-292."""
    return postprocess(synthesis(truncation(w, sd), sd, literal=literal))


# --------------------------------------------------------------------------- #
# ArcFace iresnet50                                                            #
# --------------------------------------------------------------------------- #
IRESNET50_LAYERS = (3, 4, 14, 3)      # iresnet.py:174-176
IRESNET50_PLANES = (64, 128, 256, 512)


def _bn(x: Tensor, sd: SD, p: str, eps: float = 1e-5) -> Tensor:
    """nn.BatchNorm in eval mode (iresnet.py:38,40,43,82,97,98)."""
    shape = (1, -1, 1, 1) if x.dim() == 4 else (1, -1)
    scale = sd[p + ".weight"] / torch.sqrt(sd[p + ".running_var"] + eps)
    return (x - sd[p + ".running_mean"].view(shape)) * scale.view(shape) + sd[p + ".bias"].view(shape)


def _prelu(x: Tensor, a: Tensor) -> Tensor:
    return torch.where(x >= 0, x, x * a.view(1, -1, 1, 1))


def iresnet50(x: Tensor, sd: SD, tap: Optional[Callable[[str, Tensor], None]] = None) -> Tensor:
    """IResNet.forward iresnet.py:140-154 with IBasicBlock.forward :46-57 (eval, fp16=False)."""
    x = F.conv2d(x, sd["conv1.weight"], padding=1)
    x = _prelu(_bn(x, sd, "bn1"), sd["prelu.weight"])
    if tap:
        tap("stem", x)
    for li, nblocks in enumerate(IRESNET50_LAYERS, start=1):
        for bi in range(nblocks):
            p = f"layer{li}.{bi}."
            stride = 2 if bi == 0 else 1
            identity = x
            out = _bn(x, sd, p + "bn1")
            out = F.conv2d(out, sd[p + "conv1.weight"], padding=1)
            out = _prelu(_bn(out, sd, p + "bn2"), sd[p + "prelu.weight"])
            out = F.conv2d(out, sd[p + "conv2.weight"], stride=stride, padding=1)
            out = _bn(out, sd, p + "bn3")
            if bi == 0:
                identity = _bn(F.conv2d(x, sd[p + "downsample.0.weight"], stride=stride), sd, p + "downsample.1")
            x = out + identity
            if tap:
                tap(f"layer{li}.{bi}", x)
    x = _bn(x, sd, "bn2")
    x = torch.flatten(x, 1)
    x = F.linear(x, sd["fc.weight"], sd["fc.bias"])
    return _bn(x, sd, "features")


# --------------------------------------------------------------------------- #
# Base classifier (WrappedModel) and lat2embs                                  #
# --------------------------------------------------------------------------- #
def lat2embs(w: Tensor, g_sd: SD, f_sd: SD, size: int = 112, literal: bool = True,
             faithful_padding: bool = False) -> Tensor:
    """lat2embs gen_utils.py:108-139 (few=False): chunks of 4 -> synth -> transform -> net.

    The reference zero-pads the latent batch by the rule at :112-118 and discards the padded rows
    at :136; with eval-mode BN and per-sample InstanceNorm those rows cannot influence the kept
    ones, so they are only synthesised when ``faithful_padding`` is set.
    """
    n = w.shape[0]
    if faithful_padding:
        to_pad = 0 if GAN_CHUNK % n == 0 else GAN_CHUNK - n % GAN_CHUNK
        w = torch.cat([w, torch.zeros(to_pad, EMB_SIZE)], dim=0)
    embs = []
    with torch.no_grad():
        for i in range(0, w.shape[0], GAN_CHUNK):
            ims = easy_synthesize(w[i:i + GAN_CHUNK], g_sd, literal=literal)
            embs.append(iresnet50(transform(ims, size), f_sd))
    return torch.cat(embs)[:n]


def compute_probs(emb: Tensor, gallery: Tensor) -> Tensor:
    """WrappedModel.compute_probs smoothing_model.py:56-61."""
    d = torch.cdist(emb, gallery, compute_mode="donot_use_mm_for_euclid_dist") / np.sqrt(EMB_SIZE)
    return F.softmax(-d, dim=1)


def perturb_latent(z: Tensor, p: Tensor, dir_mat: Tensor) -> Tensor:
    """WrappedModel.forward smoothing_model.py:63-67:  z[1,512] + p[b,1,1,5].squeeze @ dir_mat[5,512]."""
    return z + p.squeeze(2).squeeze(1) @ dir_mat


def wrapped_forward(z: Tensor, p: Tensor, dir_mat: Tensor, gallery: Tensor, g_sd: SD, f_sd: SD,
                    literal: bool = True) -> Tensor:
    """WrappedModel.forward smoothing_model.py:63-72 -> probs [b,N]."""
    return compute_probs(lat2embs(perturb_latent(z, p, dir_mat), g_sd, f_sd, literal=literal), gallery)


# --------------------------------------------------------------------------- #
# Smooth (MC loop, votes, Clopper-Pearson)                                     #
# --------------------------------------------------------------------------- #
def sample_noise(batch: Tensor, theta: Tensor, generator: Optional[torch.Generator] = None) -> Tensor:
    """L2Certificate.sample_noise certificate.py:64-67 (randn_like * theta)."""
    return torch.randn(batch.shape, dtype=batch.dtype, generator=generator) * theta


def count_arr(preds: Tensor, length: int) -> Tensor:
    """Smooth._count_arr smooth.py:140-146."""
    counts = torch.zeros(length, dtype=torch.long)
    unique, c = preds.unique(sorted=False, return_counts=True)
    counts[unique] = c
    return counts


def lower_confidence_bound(na: int, n: int, alpha: float) -> float:
    """Smooth._lower_confidence_bound smooth.py:148-160.

    ``statsmodels.stats.proportion.proportion_confint(NA, N, alpha=2*alpha, method='beta')[0]``
    (statsmodels is not installed; unpinned in the reference).  Its published definition of the
    Clopper-Pearson lower limit is ``beta.ppf(alpha_/2, count, nobs-count+1)`` with NaN -> 0 when
    count == 0; here alpha_ = 2*alpha.
    """
    if na == 0:
        return 0.0
    return float(_beta.ppf(alpha, na, n - na + 1))


def compute_gap(p_a_bar: float) -> float:
    """L2Certificate.compute_gap certificate.py:69-70."""
    return float(_norm.ppf(p_a_bar))


NoiseFn = Callable[[int], Tensor]   # batch size -> noise [b,1,1,5] already scaled by sigma


def sample_noise_counts(classify: Callable[[Tensor], Tensor], x: Tensor, sigma: Tensor, num: int,
                        batch_size: int, num_classes: int,
                        generator: Optional[torch.Generator] = None,
                        record: Optional[list] = None) -> np.ndarray:
    """Smooth._sample_noise smooth.py:109-138.

    ``classify(p[b,1,1,5]) -> probs [b,N]`` is the base classifier with z bound.  ``record`` (if
    given) collects the noise tensors so a CUDA run can be fed identical noise.
    """
    counts = torch.zeros(num_classes, dtype=torch.float64)
    for _ in range(math.ceil(num / batch_size)):
        b = min(batch_size, num)
        num -= b
        batch = x.repeat((b, 1, 1, 1))
        noise = sample_noise(batch, sigma, generator)
        if record is not None:
            record.append(noise.clone())
        preds = classify(batch + noise).argmax(1)
        counts += count_arr(preds, num_classes)
    return counts.numpy()


def certify(classify: Callable[[Tensor], Tensor], x: Tensor, label: int, sigma: Tensor, n0: int, n: int,
            alpha: float, batch_size: int, num_classes: int,
            generator: Optional[torch.Generator] = None, record: Optional[list] = None):
    """Smooth.certify smooth.py:39-77 -> (prediction, gap)."""
    counts0 = sample_noise_counts(classify, x, sigma, n0, batch_size, num_classes, generator, record)
    c_hat = int(counts0.argmax())
    if c_hat != label:
        return c_hat, 0.0
    counts = sample_noise_counts(classify, x, sigma, n, batch_size, num_classes, generator, record)
    p_a_bar = lower_confidence_bound(int(counts[c_hat]), n, alpha)
    if p_a_bar < 0.5:
        return ABSTAIN, 0.0
    return c_hat, compute_gap(p_a_bar)


def predict(classify: Callable[[Tensor], Tensor], x: Tensor, sigma: Tensor, n: int, alpha: float,
            batch_size: int, num_classes: int, generator: Optional[torch.Generator] = None) -> int:
    """Smooth.predict smooth.py:79-107.  ``scipy.stats.binom_test`` was removed in SciPy 1.12;
    ``binomtest(k, n, p).pvalue`` is its documented replacement (two-sided)."""
    counts = sample_noise_counts(classify, x, sigma, n, batch_size, num_classes, generator)
    top2 = counts.argsort()[::-1][:2]
    c1, c2 = counts[top2[0]], counts[top2[1]]
    if _binomtest(int(c1), int(c1 + c2), 0.5).pvalue > alpha:
        return ABSTAIN
    return int(top2[0])


# --------------------------------------------------------------------------- #
# Geometry setup (host, runs once)                                             #
# --------------------------------------------------------------------------- #
def red_ellipse_mat_inv() -> np.ndarray:
    """get_projection_matrices proj_utils.py:705-712 + get_all_matrices gen_utils.py:628.

    ``mvee`` of the mirrored axis-aligned points {+-eps_k e_k} is the axis-aligned ellipsoid
    diag(1/eps_k^2) (the Khachiyan loop at proj_utils.py:431-459 converges to it; pinned by the
    golden vector ``red_ellipse_mat_inv``), so its inverse -- the anisotropic sigma scaling used
    by certify.py:88-93 -- is eps^2.
    """
    return np.asarray(ATTR_EPS, dtype=np.float64) ** 2
