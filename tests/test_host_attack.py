"""(Named test_host_*: it computes with torch in the pytest process, so it must sort AFTER test_distributed_cpu.py, whose
workers are fork()ed.)  Host logic of the attack driver against vectors produced by the UNMODIFIED reference
(tests/golden/attack_vectors.npz, oracle/make_golden_attack.py): ellipsoid sampling / projection under the reference's own
RNG call order, every loss of `compute_loss` on the reference's distance matrix, `check_deltas`."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "attack_vectors.npz"))


def test_init_deltas_reproduces_the_reference_samples(gold):
    from certifyingfacerecognition_b200.attack_utils import gen_utils as G
    A = torch.from_numpy(gold["red_ellipse_mat"])
    torch.manual_seed(77)
    surf = G.init_deltas(True, True, 3, True, A, None, None)
    torch.manual_seed(78)
    inside = G.init_deltas(True, True, 16, False, A, None, None)
    assert np.allclose(surf.numpy(), gold["init_surface"], atol=1e-6)
    assert np.allclose(inside.numpy(), gold["init_inside"], atol=1e-6)
    mags = G.check_deltas(surf, True, A, None, None)
    assert torch.allclose(mags, torch.ones(3), atol=1e-3)                       # on the surface (1 / (1 + 1e-4)^2)
    assert (G.check_deltas(inside, True, A, None, None) <= 1.0).all()


def test_projection_into_the_ellipsoid_matches_the_reference(gold):
    from certifyingfacerecognition_b200.attack_utils import gen_utils as G, proj_utils as P
    A = torch.from_numpy(gold["red_ellipse_mat"])
    pts = torch.from_numpy(gold["proj_in"])
    proj, before = P.proj2region(pts.clone(), None, A, to_subs=False, check=True, diag_ellipse_mat=True)
    assert torch.equal(before, pts)
    assert np.allclose(proj.numpy(), gold["proj_out"], atol=1e-6)               # SciPy bisection vs batched bisection
    mags = G.check_deltas(proj, True, A, None, None)
    assert np.allclose(mags.numpy(), gold["proj_mag"], atol=1e-5)
    was_inside = (A * pts ** 2).sum(1) <= 1
    assert torch.equal(proj[was_inside], pts[was_inside])                       # interior points are left alone
    assert (mags[~was_inside] > 0.999).all() and (mags <= 1 + 1e-4).all()       # exterior points land on the surface
    # Euclidean projection: no feasible point is closer (spot check against random surface points)
    g = torch.Generator().manual_seed(3)
    u = torch.randn(4000, 5, generator=g)
    u = u / torch.sqrt((A * u ** 2).sum(1, keepdim=True))
    for i in torch.nonzero(~was_inside).flatten()[:6]:
        assert (pts[i] - proj[i]).norm() <= (pts[i] - u).norm(dim=1).min() + 1e-5
    with pytest.raises(NotImplementedError):
        P.proj2region(pts, torch.eye(5), A, to_subs=True)


@pytest.mark.parametrize("loss", ["xent", "away", "diff", "nearest", "dlr"])
def test_losses_on_the_reference_distance_matrix(gold, loss):
    from certifyingfacerecognition_b200.attack_utils import gen_utils as G
    d = torch.from_numpy(gold["all_dists"])
    labels = torch.arange(d.shape[0])
    got = G.compute_loss(d, labels, loss_type=loss, use_probs=loss != "dlr").item()
    assert got == pytest.approx(float(gold[f"loss_{loss}"]), rel=1e-6, abs=1e-9)
    per = G.per_sample_loss(d, labels, loss, use_probs=loss != "dlr")
    assert per.shape == (d.shape[0],) and per.mean().item() == pytest.approx(got, rel=1e-6)


def test_optimisers_and_distances():
    from certifyingfacerecognition_b200.attack_utils import gen_utils as G
    p = torch.zeros(2, 5, requires_grad=True)
    assert isinstance(G.get_optim(p, "SGD", 1e2, 0.9), torch.optim.SGD)
    assert isinstance(G.get_optim(p, "Adam", 1e-2), torch.optim.Adam)
    assert isinstance(G.get_optim(p, "RMSProp", 1e-2), torch.optim.RMSprop)
    with pytest.raises(ValueError):
        G.get_optim(p, "LBFGS")
    a, b = torch.randn(3, 512), torch.randn(7, 512)
    assert torch.allclose(G.get_dists(a, b), torch.cdist(a, b), atol=1e-4)
    an, bn = torch.nn.functional.normalize(a), torch.nn.functional.normalize(b)
    assert torch.allclose(G.get_dists(an, bn, "facenet"), 1 - an @ bn.T)
    with pytest.raises(NotImplementedError):
        G.find_adversaries_pgd(None, None, a, torch.arange(3), b, "SGD", 1.0, 1, 0.9, "insightface", "xent", None, None,
                               None, torch.zeros(512, 5), None, torch.ones(5), lin_comb=False)
