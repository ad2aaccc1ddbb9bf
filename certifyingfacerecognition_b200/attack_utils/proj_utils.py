"""Host-side geometry setup used by certify: the part of the reference's attack_utils/proj_utils.py that the
certification path touches (ATTRS :16-21, get_full_points :317-338, mvee :431-459, get_proj_mat :624-627,
get_projection_matrices :661-718, get_ellipse_mat :721-728).  NumPy/SciPy, runs once."""
from __future__ import annotations

import os.path as osp
from collections import OrderedDict

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


def selected_attrs(attrs2drop=()):
    """The (name, budget) pairs that stay after dropping `attrs2drop`, in ATTRS order.  The module-level table is NOT
    edited (the reference pops the dropped names out of its global dict, proj_utils.py:670-675, which is harmless there
    because it builds its matrices once; here main_attack.py asks for the reduced AND the full direction set)."""
    unknown = [a for a in attrs2drop if a not in ATTRS]
    assert not unknown, f"Attribute {unknown[0]} is NOT valid"
    return [(name, eps) for name, eps in ATTRS.items() if name not in set(attrs2drop)]


def get_projection_matrices(dataset=DATASETS[0], gan_name=GAN_NAMES[0], attrs2drop=(), scale_factor=1.0):
    """proj_utils.py:661-718 -> (proj_mat[512,512], ellipse_mat[512,512], dirs[512,n_dirs], red_ellipse_mat[n_dirs], files):
    the projector onto the span of the kept InterFaceGAN boundaries, the minimum-volume ellipsoid through +-eps_k d_k
    (completed with the null space) and its n_dirs-dimensional diagonal counterpart 1 / eps_k^2, both times scale_factor."""
    keep = selected_attrs(attrs2drop)
    files = [osp.join(BOUNDARIES_DIR, f"{gan_name}_{dataset}_{name}_w_boundary.npy") for name, _ in keep]
    for (name, _), path in zip(keep, files):
        assert osp.isfile(path), f'Boundary for attr "{name}" not found!'
    dirs = np.concatenate([np.load(path) for path in files], axis=0).T          # one [1,512] row per file -> [512,n_dirs]
    assert dirs.shape[1] == len(keep)
    budgets = np.diag(np.array([eps for _, eps in keep]))
    ellipse_mat = scale_factor * get_ellipse_mat(dirs)
    reduced = scale_factor * get_ellipse_mat(budgets)
    red_diag = np.diagonal(reduced)
    assert np.all(reduced == np.diag(red_diag)), "Matrix should be diagonal"
    return get_proj_mat(dirs), ellipse_mat, dirs, red_diag, files


# ------------------------------------------------------------------------------------------------------------------
# Ellipsoid helpers of the attack path, torch, for the low-dimensional DIAGONAL ellipsoid the
# linear-combination attack lives in (delta in R^5, sum_k A_k delta_k^2 <= 1 with A_k = 1 / eps_k^2).  The 512-D
# subspace variants of the reference (proj_ellipse_pytorch :134-209, to_subs=True) belong to `--no-lin-comb`, which
# needs gradients with respect to 512 latent coordinates and is not built.
# ------------------------------------------------------------------------------------------------------------------
def sq_distance(A, shifted1, shifted2=None):
    """proj_utils.py:36-48: x1^T A x2 for a batch of column vectors [b,d,1] -> [b]."""
    import torch
    if shifted2 is None:
        shifted2 = shifted1
    return torch.einsum("bi,ij,bj->b", shifted1.squeeze(-1), A.to(shifted1.dtype), shifted2.squeeze(-1))


def _diag_of(mat):
    """The diagonal of `mat` if it is a vector or a diagonal matrix, else None."""
    import torch
    if mat.ndim == 1:
        return mat
    d = torch.diagonal(mat)
    return d if bool(torch.all(mat - torch.diag(d) == 0)) else None


def in_ellps(v, ellipse_mat, atol=1e-4):
    """proj_utils.py:507-510: every column of v [d,n] inside (or within atol of) x^T A x <= 1."""
    import torch
    dists = sq_distance(ellipse_mat, v.T.unsqueeze(2))
    out = dists > 1.0
    return torch.allclose(dists[out], torch.tensor(1.0, dtype=dists.dtype, device=dists.device), atol=atol)


def sample_ellipsoid(ellipsoid_mat, n_vecs=1):
    """proj_utils.py:396-428 (torch branch): uniform in the unit ball (normalised Gaussian directions, radius U^(1/n)),
    mapped onto the ellipsoid by the inverse transposed Cholesky factor.  Same RNG call order as the reference, so a
    seeded CPU generator reproduces its samples."""
    import torch
    n = ellipsoid_mat.shape[0]
    vec = torch.randn(n, n_vecs, device=ellipsoid_mat.device)
    vec = vec / torch.norm(vec, dim=0)
    vec = vec * torch.rand(n_vecs, device=ellipsoid_mat.device) ** (1 / n)
    chol = torch.linalg.cholesky(ellipsoid_mat)
    return (torch.linalg.inv(chol.T) @ vec).T


def proj_ellipse_pytorch_diag(y, vec_A, mu=None, c=1):
    """proj_utils.py:212-285: Euclidean projection of the columns of y [d,n] onto {x : sum_k A_k x_k^2 <= c}.  A point
    outside is x_k = y_k / (1 + t A_k) with t > 0 the root of sum_k A_k y_k^2 / (1 + t A_k)^2 = 1; the reference
    brackets it in [eps, 1e3] and bisects one vector at a time with SciPy -- here all vectors are bisected together in
    float64 (100 halvings).  Points for which the bracket has no sign change (inside, or farther than the bracket
    reaches) are returned unchanged, as in the reference.  -> (projections [d,n], t [m,1], inside mask [n])."""
    import torch
    if mu is not None:
        raise NotImplementedError("proj_ellipse_pytorch_diag: only mu = None is used on the attack path")
    a = (vec_A / c).to(torch.float64)
    yy = y.T.to(torch.float64)                                            # [n,d]

    def phi(t):                                                           # [n] for t [n]
        return (a * yy ** 2 / (1 + t.unsqueeze(1) * a) ** 2).sum(1) - 1

    lo = torch.full((yy.shape[0],), float(np.finfo(float).eps), dtype=torch.float64, device=y.device)
    hi = torch.full_like(lo, 1e3)
    which_out = phi(lo) * phi(hi) < 0
    for _ in range(100):
        mid = 0.5 * (lo + hi)
        pos = phi(mid) > 0                                                # phi decreases in t: root is to the right
        lo = torch.where(pos, mid, lo)
        hi = torch.where(pos, hi, mid)
    t = 0.5 * (lo + hi)
    proj = torch.where(which_out.unsqueeze(1), yy / (1 + t.unsqueeze(1) * a), yy)
    inside = (a * proj ** 2).sum(1) <= 1
    return proj.T.to(y.dtype), t[which_out].reshape(-1, 1), inside


def proj2region(vs, proj_mat, ellipse_mat, check=True, dirs=None, to_subs=True, on_surface=False, max_iters=5,
                diag_ellipse_mat=False):
    """proj_utils.py:513-581 for the low-dimensional ellipsoid (to_subs=False): project the rows of vs [n,d] into (or,
    with on_surface, first radially onto) the ellipsoid.  ellipse_mat: the diagonal as a vector (diag_ellipse_mat=True)
    or a diagonal matrix.  -> (projected [n,d], the vectors before the ellipsoid projection [n,d])."""
    import torch
    if to_subs:
        raise NotImplementedError("proj2region(to_subs=True) is the 512-D `--no-lin-comb` path (not built)")
    vec_A = ellipse_mat if diag_ellipse_mat else _diag_of(ellipse_mat)
    if vec_A is None:
        raise NotImplementedError("proj2region: only diagonal ellipsoid matrices are supported")
    ell_mat = torch.diag(vec_A)

    def proj2surf(v):
        d = sq_distance(ell_mat, v.T.unsqueeze(2))
        return v / (torch.sqrt(d.reshape(1, -1)) + 1e-4)

    v = vs.T
    proj_subs = proj2surf(v) if on_surface else v
    proj_ell, _, _ = proj_ellipse_pytorch_diag(proj_subs, vec_A)
    iters = 0
    while not in_ellps(proj_ell, ell_mat) and iters < max_iters:
        iters += 1
        proj_ell, _, _ = proj_ellipse_pytorch_diag(proj_ell, vec_A)
    if not in_ellps(proj_ell, ell_mat):                                   # still a bit outside: radially onto the surface
        d = sq_distance(ell_mat, proj_ell.T.unsqueeze(2))
        need = (torch.sqrt(d) >= 1)
        proj_ell = torch.where(need.unsqueeze(0), proj2surf(proj_ell), proj_ell)
    if check:
        assert in_ellps(proj_ell, ell_mat), "Some points outside ellipsoid!"
    return proj_ell.T, proj_subs.T
