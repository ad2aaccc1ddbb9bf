"""Drop-in for the reference's models/smoothing_model.py (WrappedModel :12-72): the base classifier
latent + perturbation -> StyleGAN -> FRM -> gallery match, backed by the CUDA engine."""
from __future__ import annotations

import logging
import os.path as osp
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..attack_utils.gen_utils import EMB_SIZE, STYLEGAN_PATH, WEIGHTS_PATH, get_latent_codes
from ..engine import Engine

log = logging.getLogger(__name__)


def random_stylegan_state(seed: int = 0) -> Dict[str, torch.Tensor]:
    """What the reference ends up with when models/pretrain/stylegan_ffhq.pth is absent: it logs a warning and
    keeps a random init (base_generator.py:94-99).  Here: N(0,1) conv / dense weights (the run-time wscale does
    the He scaling), zero biases / noise gains / w_avg, ones const input -- the constructor's values."""
    from ..engine import NUM_LAYERS, layer_channels, layer_res
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    sd = {"truncation.w_avg": torch.zeros(512), "synthesis.layer0.first_layer": torch.ones(1, 512, 4, 4)}
    for l in range(NUM_LAYERS):
        c, res, p = layer_channels(l), layer_res(l), f"synthesis.layer{l}."
        if l >= 1:
            cin = layer_channels(l - 1)
            if l % 2 == 0 and res >= 128:
                sd[p + "weight"] = rn(3, 3, cin, c)
            else:
                sd[p + "conv.weight"] = rn(c, cin, 3, 3)
        sd[p + "epilogue.apply_noise.noise"] = rn(1, 1, res, res)
        sd[p + "epilogue.apply_noise.weight"] = torch.zeros(c)
        sd[p + "epilogue.bias"] = torch.zeros(c)
        sd[p + "epilogue.style_mod.dense.linear.weight"] = rn(2 * c, 512)
        sd[p + "epilogue.style_mod.dense.wscale.bias"] = torch.zeros(2 * c)
    sd["synthesis.output8.conv.weight"] = rn(3, 16, 1, 1)
    sd["synthesis.output8.bias"] = torch.zeros(3)
    return sd


class WrappedModel(nn.Module):
    """smoothing_model.py:12-72.  Positional signature unchanged; keyword-only extras let callers hand in
    state dicts / latents directly instead of the cwd-relative files the reference reads."""
    supports_fused_votes = True

    def __init__(self, direction_matrix, face_recog="insightface", n_embs=-1, load_embs=False, embs_file=None, *,
                 generator_state: Optional[Dict[str, torch.Tensor]] = None,
                 frm_state: Optional[Dict[str, torch.Tensor]] = None,
                 latents: Optional[torch.Tensor] = None, orig_embs: Optional[torch.Tensor] = None,
                 chunk: int = 32, frm_group: int = 1, tail_chunks="auto") -> None:
        super().__init__()
        if face_recog not in ("insightface", "facenet", "facenet-vggface2"):
            raise ValueError(f"face_recog='{face_recog}' is not one of the reference's FRS_METHODS (gen_utils.py:31-35)")
        if face_recog != "insightface" and frm_state is None:
            # main_attack.py:126-129 downloads the facenet_pytorch weights at run time; there is no file to read here
            raise ValueError("the FaceNet variants need `frm_state` = an InceptionResnetV1 state dict "
                             "(facenet_pytorch downloads its weights; nothing is stored in the reference tree)")
        self.device = direction_matrix.device
        if self.device.type != "cuda":
            raise RuntimeError("WrappedModel needs its direction matrix on a CUDA device (no CPU fallback)")
        self.face_recog = face_recog
        self.dir_mat = direction_matrix                      # [num_directions, 512]
        if generator_state is None:
            if osp.isfile(STYLEGAN_PATH):
                generator_state = torch.load(STYLEGAN_PATH, map_location="cpu")
            else:
                log.warning("No pre-trained model will be loaded!")      # base_generator.py:99
                generator_state = random_stylegan_state()
        if frm_state is None:
            frm_state = torch.load(WEIGHTS_PATH, map_location="cpu")      # main_attack.py:124
        self.latents = (latents if latents is not None else get_latent_codes()).to(self.device)
        if orig_embs is not None:
            embs = orig_embs
        elif load_embs:
            read_from = osp.join("embeddings", f"embs_{face_recog}.pth") if embs_file is None else embs_file
            print(f'Loading original embeddings from "{read_from}"')
            embs = torch.load(read_from, map_location="cpu")
            n_embs = embs.shape[0] if n_embs == -1 else n_embs
            print(f"Loaded {n_embs} out of {embs.size(0)} embeddings")
            embs = embs[:n_embs]
        else:
            embs = None
        placeholder = embs if embs is not None else torch.zeros(1, EMB_SIZE)
        self.engine = Engine(generator_state, frm_state, self.dir_mat, placeholder, chunk=chunk, frm_group=frm_group,
                             tail_chunks=tuple(c for c in (128, 64, 32, 16) if c < chunk) if tail_chunks == "auto" else tail_chunks,
                             device=self.device,
                             frm=face_recog)
        if embs is None:
            print("Generating original embeddings")
            embs = self.engine.embed_latents(self.latents)
            self.engine.set_gallery(embs)
            print(f"Computed {embs.size(0)} embeddings")
        self.orig_embs = self.engine.gallery

    # ---------------------------------------------------------------------------------------------
    def compute_probs(self, embedding: torch.Tensor) -> torch.Tensor:
        """smoothing_model.py:56-61 (kept on the device; the reference runs it on the CPU)."""
        d = torch.cdist(embedding.to(self.device), self.orig_embs, compute_mode="donot_use_mm_for_euclid_dist")
        return F.softmax(-d / np.sqrt(EMB_SIZE), dim=1)

    def embed(self, x: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
        """Embeddings of x + p @ dir_mat (smoothing_model.py:63-69).  x: [1,512] (what Smooth passes) or [b,512] with
        one latent per row of p, as the reference's broadcast ``x + pert`` allows."""
        p = p.reshape(-1, self.dir_mat.shape[0])
        x = x.reshape(-1, 512)
        if x.shape[0] != 1:
            if x.shape[0] != p.shape[0]:
                raise ValueError(f"WrappedModel: x has {x.shape[0]} latents for {p.shape[0]} perturbations")
            w = x.to(self.device, torch.float32) + p.to(self.device, torch.float32) @ self.dir_mat.float()
            return self.engine.embed_latents(w)                 # lat2embs on the perturbed latents, row by row
        _, extra = self.engine.sample_votes(x, torch.zeros(self.dir_mat.shape[0]), torch.ones(1), p.shape[0],
                                            noise=p, want_emb=True,
                                            counts=torch.zeros(self.engine.num_classes, dtype=torch.int64,
                                                               device=self.device))
        return extra["emb"]

    def forward(self, x, p=0):
        """smoothing_model.py:63-72 -> probs [b, N] on the device.  API-compatibility path: Smooth does not call
        it for this class (it uses ``sample_votes``), since materialising [b, N] probabilities is exactly the
        traffic the fused match+vote kernel avoids."""
        return self.compute_probs(self.embed(x, p))

    def sample_votes(self, z, x, sigma, num: int, seed: int = 0, sample_offset: int = 0,
                     noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Per-identity int64 vote counts of ``num`` MC samples around (z, x) -- the body of
        Smooth._sample_noise (smooth.py:126-137) as one call.  ``noise`` [num, 5]: injected (already scaled) noise
        instead of the Philox stream."""
        counts, _ = self.engine.sample_votes(z, x, sigma, num, seed=seed, sample_offset=sample_offset, noise=noise)
        return counts

    def sample_votes_multi(self, z, x, sigma, nums, seed: int = 0, sample_offsets=None) -> torch.Tensor:
        """``sample_votes`` for several identities at once (z [G,512]); their samples share program runs.  -> [G, N] int64."""
        return self.engine.sample_votes_multi(z, x, sigma, nums, seed=seed, sample_offsets=sample_offsets)
