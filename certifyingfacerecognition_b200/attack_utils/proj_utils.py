"""Host-side geometry setup used by certify: the part of the reference's attack_utils/proj_utils.py that the
certification path touches (ATTRS :16-21, get_full_points :317-338, mvee :431-459, get_proj_mat :624-627,
get_projection_matrices :661-718, get_ellipse_mat :721-728).  NumPy/SciPy, runs once."""
from __future__ import annotations

import os.path as osp
from collections import OrderedDict
from glob import glob

import numpy as np
import scipy.linalg

BOUNDARIES_DIR = "boundaries"
DATASETS = ["ffhq", "celebahq"]
GAN_NAMES = ["stylegan", "pggan"]
# per-attribute perturbation budgets (proj_utils.py:16-21); users edit these exactly as in the reference
ATTRS = OrderedDict()
ATTRS["age"] = 0.5
ATTRS["eyeglasses"] = 0.5
ATTRS["gender"] = 0.2
ATTRS["pose"] = 0.5
ATTRS["smile"] = 0.8


def set_seed(device, seed=111):
    """proj_utils.py:28-33."""
    import random
    import torch
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if str(device).startswith("cuda"):
        torch.cuda.manual_seed_all(seed)


def get_full_points(points, fill_with_null=False):
    """Mirror the point set through the origin, optionally completing it with a null-space basis first."""
    if fill_with_null:
        null = scipy.linalg.null_space(points.T)
        points = np.concatenate([points, null], axis=1)
        assert points.shape[0] == points.shape[1]
    return np.concatenate((points, -points), axis=1)


def mvee(points, tol=0.001):
    """Khachiyan's minimum-volume enclosing ellipsoid: returns (A, c) with (x-c)^T A (x-c) = 1."""
    points = np.asarray(points, dtype=np.float64)
    n, d = points.shape
    q = np.column_stack((points, np.ones(n))).T
    err = tol + 1.0
    u = np.ones(n) / n
    while err > tol:
        xmat = q @ np.diag(u) @ q.T
        m = np.einsum("ij,ji->i", q.T @ np.linalg.inv(xmat), q)
        j = int(np.argmax(m))
        step = (m[j] - d - 1.0) / ((d + 1) * (m[j] - 1.0))
        new_u = (1 - step) * u
        new_u[j] += step
        err = np.linalg.norm(new_u - u)
        u = new_u
    c = u @ points
    a = np.linalg.inv(points.T @ np.diag(u) @ points - np.outer(c, c)) / d
    return a, c


def get_proj_mat(dirs):
    return dirs @ np.linalg.pinv(dirs)


def get_ellipse_mat(dirs):
    ellipse_mat, c = mvee(get_full_points(dirs, fill_with_null=True).T)
    assert np.allclose(c, 0), "The origin should be the ellipses's center"
    return ellipse_mat


def get_projection_matrices(dataset=DATASETS[0], gan_name=GAN_NAMES[0], attrs2drop=(), scale_factor=1.0):
    """-> (proj_mat[512,512], ellipse_mat[512,512], dirs[512,n_dirs], red_ellipse_mat[n_dirs], files)."""
    template = osp.join(BOUNDARIES_DIR, f"{gan_name}_{dataset}_%s_w_boundary.npy")
    all_bounds = glob(osp.join(BOUNDARIES_DIR, "*.npy"))
    for attr in attrs2drop:
        assert attr in ATTRS.keys(), f"Attribute {attr} is NOT valid"
        ATTRS.pop(attr)
    dirs, files, magns = [], [], []
    for att_name, magn in ATTRS.items():
        this_file = template % att_name
        assert this_file in all_bounds, f'Boundary for attr "{att_name}" not found!'
        dirs.append(np.load(this_file))
        magns.append(magn)
        files.append(this_file)
    dirs = np.concatenate(dirs, axis=0).T
    assert dirs.shape[1] == len(ATTRS)
    proj_mat = get_proj_mat(dirs)
    ellipse_mat = scale_factor * get_ellipse_mat(dirs)
    red_ellipse_mat = scale_factor * get_ellipse_mat(np.diag(np.array(magns)))
    assert np.all(red_ellipse_mat == np.diag(np.diagonal(red_ellipse_mat))), "Matrix should be diagonal"
    return proj_mat, ellipse_mat, dirs, np.diagonal(red_ellipse_mat), files
