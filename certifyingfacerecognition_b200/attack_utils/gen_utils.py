"""The five functions / constants of the reference's attack_utils/gen_utils.py that certify.py uses
(:17-41 constants, get_latent_codes :44-46, get_all_matrices :607-631).  Everything else in that file belongs to
the adversarial-attack workload and is out of scope."""
from __future__ import annotations

import os.path as osp

import numpy as np
import torch

from .proj_utils import get_projection_matrices

INP_RESOLS = {"insightface": 112, "facenet": 160, "facenet-vggface2": 160}
STD = 0.5
MEAN = 0.5
EMB_SIZE = 512
LAT_SPACE = "w"
DATASET = "ffhq"
GAN_NAME = "stylegan"
FRS_METHODS = ["insightface", "facenet", "facenet-vggface2"]
ORIG_DATA_PATH = f"data/{GAN_NAME}_{DATASET}_1M"
LAT_CODES_PATH = osp.join(ORIG_DATA_PATH, f"{LAT_SPACE}.npy")
WEIGHTS_PATH = "weights/ms1mv3_arcface_r50/backbone.pth"          # main_attack.py:43
STYLEGAN_PATH = "models/pretrain/stylegan_ffhq.pth"                # models/model_settings.py:50


def _device():
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def get_latent_codes(generator=None) -> torch.Tensor:
    """gen_utils.py:44-46 (+ preprocess, mod_stylegan_generator.py:160-190, W branch): [N,512] float32."""
    return torch.from_numpy(np.load(LAT_CODES_PATH).reshape(-1, EMB_SIZE).astype(np.float32))


def get_all_matrices(attrs2drop=None, scale_factor=1.0, device=None):
    """gen_utils.py:607-631 -> (proj_mat, ellipse_mat, ellipse_mat_inv, dirs, dirs_inv, red_ellipse_mat,
    red_ellipse_mat_inv) as float32 tensors on the device."""
    dev = device or _device()
    proj_mat, ellipse_mat, dirs, red_ellipse_mat, _ = get_projection_matrices(
        dataset=DATASET, gan_name=GAN_NAME, attrs2drop=list(attrs2drop or []), scale_factor=scale_factor)
    dirs = torch.tensor(dirs, dtype=torch.float32, device=dev)
    proj_mat = torch.tensor(proj_mat, dtype=torch.float32, device=dev)
    ellipse_mat = torch.tensor(ellipse_mat, dtype=torch.float32, device=dev)
    assert proj_mat.size(0) == proj_mat.size(1) == EMB_SIZE
    red_ellipse_mat = torch.tensor(red_ellipse_mat, dtype=torch.float32, device=dev)
    assert red_ellipse_mat.size(0) == dirs.size(1)
    red_ellipse_mat_inv = 1 / red_ellipse_mat
    dirs_inv = torch.linalg.pinv(dirs)
    ellipse_mat_inv = torch.linalg.inv(ellipse_mat)
    return proj_mat, ellipse_mat, ellipse_mat_inv, dirs, dirs_inv, red_ellipse_mat, red_ellipse_mat_inv
