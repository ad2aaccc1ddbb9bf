"""Image / latent generation front-end: the part of the reference's ModStyleGANGenerator that generate_data.py drives
(models/mod_stylegan_generator.py:142-309, generate_data.py:58-123), backed by the CUDA engine.

Z -> MappingModule (cfr_mapping) -> W -> truncation -> 18-layer synthesis (the same tcgen05 programs as the
certification path) -> toRGB + postprocess at the full 1024^2 resolution.  Method names, argument meaning, returned
dictionary keys ('z', 'w', 'wp', 'styleNN', 'image') and error behaviour follow the reference.  'WP' inputs (18 arbitrary
per-layer latents, no truncation) evaluate the 18 style dense layers on the host side of the API and replay the recorded
program from its second op (the engine's own styles kernel feeds the truncated / un-truncated pair of one W)."""
from __future__ import annotations

import math
from typing import Dict, Iterator, Optional

import numpy as np
import torch

from .. import _lib as L
from ..engine import NUM_LAYERS, PSI, TRUNC_LAYERS, SynthesisProgram, _f32, layer_channels

LATENT_DIM = 512          # model_settings.py:48-72 ('stylegan_ffhq': latent_space_dim = w_space_dim = 512, 18 layers)


def pack_mapping(g_sd: Dict[str, torch.Tensor], device) -> Optional[tuple]:
    """(wt [8,512,512] = W^T * scale, bias [8,512] * lr_mul) for cfr_mapping, or None when the state dict has no
    ``mapping.*`` keys.  scale = sqrt(2)/sqrt(512) * 0.01 (WScaleLayer, stylegan_generator_model.py:519-524,777)."""
    if "mapping.dense0.linear.weight" not in g_sd:
        return None
    lr_mul = 0.01
    scale = math.sqrt(2.0) / math.sqrt(512.0) * lr_mul
    wt = torch.stack([g_sd[f"mapping.dense{i}.linear.weight"].detach().float().t().contiguous() * scale for i in range(8)])
    bias = torch.stack([g_sd[f"mapping.dense{i}.wscale.bias"].detach().float() * lr_mul for i in range(8)])
    return _f32(wt, device), _f32(bias, device)


class StyleGANGenerator:
    """Drop-in for the calls generate_data.py makes on ModStyleGANGenerator('stylegan_ffhq')."""

    gan_type = "stylegan"
    latent_space_dim = LATENT_DIM
    w_space_dim = LATENT_DIM
    num_layers = NUM_LAYERS
    resolution = 1024

    def __init__(self, generator_state: Dict[str, torch.Tensor], batch_size: int = 4, device="cuda") -> None:
        if not torch.cuda.is_available():
            raise RuntimeError("StyleGANGenerator needs a CUDA device (no CPU fallback)")
        self.lib = L.load()
        self.device = torch.device(device)
        self.batch_size = batch_size          # model_settings.MAX_IMAGES_ON_DEVICE in the reference (default 4)
        self.mapping = pack_mapping(generator_state, self.device)
        # full-resolution RGB out: no resize (1024 -> 1024), no normalisation (mean 0, std 1) => postprocess()'d image
        self.synth = SynthesisProgram(generator_state, batch_size, out_res=1024, device=device, keep_planar=True,
                                      mean=0.0, std=1.0, nhwc_out=False)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    # ---- latent handling (mod_stylegan_generator.py:134-192) -------------------------------------------------
    def sample(self, num: int, latent_space_type: str = "Z") -> np.ndarray:
        t = latent_space_type.upper()
        if t == "Z" or t == "W":
            return np.random.randn(num, LATENT_DIM).astype(np.float32)
        if t == "WP":
            return np.random.randn(num, NUM_LAYERS, LATENT_DIM).astype(np.float32)
        raise ValueError(f"Latent space type `{latent_space_type}` is invalid!")

    def preprocess(self, latent_codes: np.ndarray, latent_space_type: str = "Z") -> np.ndarray:
        if not isinstance(latent_codes, np.ndarray):
            raise ValueError("Latent codes should be with type `numpy.ndarray`!")
        t = latent_space_type.upper()
        if t == "Z":
            latent_codes = latent_codes.reshape(-1, LATENT_DIM)
            norm = np.linalg.norm(latent_codes, axis=1, keepdims=True)
            latent_codes = latent_codes / norm * np.sqrt(LATENT_DIM)
        elif t == "W":
            latent_codes = latent_codes.reshape(-1, LATENT_DIM)
        elif t == "WP":
            latent_codes = latent_codes.reshape(-1, NUM_LAYERS, LATENT_DIM)
        else:
            raise ValueError(f"Latent space type `{latent_space_type}` is invalid!")
        return latent_codes.astype(np.float32)

    def easy_sample(self, num: int, latent_space_type: str = "Z") -> np.ndarray:
        return self.preprocess(self.sample(num, latent_space_type), latent_space_type)

    def get_batch_inputs(self, latent_codes: np.ndarray) -> Iterator[np.ndarray]:
        """base_generator.py get_batch_inputs: consecutive slices of at most batch_size rows."""
        for i in range(0, latent_codes.shape[0], self.batch_size):
            yield latent_codes[i:i + self.batch_size]

    # ---- synthesis (mod_stylegan_generator.py:194-309) ---------------------------------------------------------
    def map_latents(self, z: torch.Tensor) -> torch.Tensor:
        """MappingModule.forward on the device: [b,512] Z -> [b,512] W."""
        if self.mapping is None:
            raise ValueError("the generator state dict has no `mapping.*` weights: only W latents can be synthesised")
        z = _f32(z.reshape(-1, LATENT_DIM), self.device)
        w = torch.empty_like(z)
        L.check(self.lib.cfr_mapping(L.ptr(z), L.ptr(self.mapping[0]), L.ptr(self.mapping[1]), z.shape[0], L.ptr(w),
                                     self._stream()))
        return w

    def synthesize(self, latent_codes, latent_space_type: str = "Z", generate_style: bool = False,
                   generate_image: bool = True) -> Dict[str, object]:
        t = latent_space_type.upper()
        lc = latent_codes if isinstance(latent_codes, torch.Tensor) else torch.from_numpy(np.asarray(latent_codes))
        results: Dict[str, object] = {}
        if t in ("Z", "W"):
            if not (lc.dim() == 2 and lc.shape[0] <= self.batch_size and lc.shape[1] == LATENT_DIM):
                raise ValueError("Latent_codes should be with shape [batch_size, latent_space_dim], where `batch_size` no "
                                 f"larger than {self.batch_size}, and `latent_space_dim` equal to {LATENT_DIM}!\n"
                                 f"But {tuple(lc.shape)} received!")
            if t == "Z":
                ws = self.map_latents(lc)
                results["z"] = latent_codes
                results["w"] = ws.cpu().numpy()
            else:
                ws = _f32(lc, self.device)
                results["w"] = latent_codes
        elif t == "WP":
            # per-layer latents go straight to the synthesis network, no truncation (mod_stylegan_generator.py:257-279)
            if not (lc.dim() == 3 and lc.shape[0] <= self.batch_size and lc.shape[1] == NUM_LAYERS and lc.shape[2] == LATENT_DIM):
                raise ValueError("Latent_codes should be with shape [batch_size, num_layers, latent_space_dim], where "
                                 f"`batch_size` no larger than {self.batch_size}, `num_layers` equal to {NUM_LAYERS}, and "
                                 f"`latent_space_dim` equal to {LATENT_DIM}!\nBut {tuple(lc.shape)} received!")
            return self._synthesize_wp(_f32(lc, self.device), latent_codes, generate_style, generate_image)
        else:
            raise ValueError(f"Latent space type `{latent_space_type}` is invalid!")
        b = ws.shape[0]
        w_avg = self.synth.w_avg.view(1, 1, LATENT_DIM)
        coefs = torch.ones(1, NUM_LAYERS, 1, device=self.device)
        coefs[:, :TRUNC_LAYERS] *= PSI
        wps = w_avg + (ws.view(b, 1, LATENT_DIM) - w_avg) * coefs       # TruncationModule :322-328 (tiny, host-side API)
        results["wp"] = wps.cpu().numpy()
        if generate_style or generate_image:
            self.synth.out_slot.zero_()
            L.check(self.lib.cfr_truncate(L.ptr(ws), L.ptr(self.synth.w_avg), PSI, b, L.ptr(self.synth.wp2), self._stream()))
            self.synth.run()
        if generate_style:
            st = self.synth.styles[:b]
            for i in range(NUM_LAYERS):
                off, c = self.synth.style_off[i], layer_channels(i)
                results[f"style{i:02d}"] = st[:, off:off + 2 * c].cpu().numpy()
        if generate_image:
            results["image"] = self.synth.img_planar[:b].clone()          # already postprocess()'d (see easy_synthesize)
        return results

    def _synthesize_wp(self, wps: torch.Tensor, latent_codes, generate_style: bool, generate_image: bool) -> Dict[str, object]:
        """'WP' inputs: every layer has its own latent, so the 18 style dense layers (stylegan_generator_model.py:503) are
        evaluated here -- styles[:, layer l] = wp[:, l] @ W_l^T / sqrt(512) + b_l, a host-side API path -- written into the
        program's style buffer, and the recorded program is replayed from its second op (op 0 is the W-path styles kernel)."""
        results: Dict[str, object] = {"wp": latent_codes}
        b = wps.shape[0]
        syn = self.synth
        if generate_style or generate_image:
            st = syn.styles
            for l in range(NUM_LAYERS):
                off, c = syn.style_off[l], layer_channels(l)
                st[:b, off:off + 2 * c] = (wps[:, l] @ syn.w_style[off:off + 2 * c].T) * (1.0 / math.sqrt(LATENT_DIM)) \
                    + syn.b_style[off:off + 2 * c]
            syn.out_slot.zero_()
            syn.run_range(1, syn.num_launches)
        if generate_style:
            for i in range(NUM_LAYERS):
                off, c = syn.style_off[i], layer_channels(i)
                results[f"style{i:02d}"] = syn.styles[:b, off:off + 2 * c].cpu().numpy()
        if generate_image:
            results["image"] = syn.img_planar[:b].clone()
        return results

    def easy_synthesize(self, latent_codes, **kwargs) -> Dict[str, object]:
        """synthesize() + postprocess() ((x+1)/2 + 0.5/255, clamp [0,1], :300-309).  The toRGB kernel applies the
        postprocess itself, so 'image' is [b,3,1024,1024] fp32 in [0,1] (what the reference's easy_synthesize returns)."""
        return self.synthesize(latent_codes, **kwargs)
