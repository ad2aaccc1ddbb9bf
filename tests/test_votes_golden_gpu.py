"""Top-1 / vote / radius parity of the CUDA path against >= 1000 Monte-Carlo samples per regime classified by the
UNMODIFIED reference (tests/golden/votes_{iso,aniso}.npz, written by oracle/make_golden_votes.py from
/root/reference: Smooth.certify(z, x, 0, N0=100, n=1000, alpha=0.001, batch 100) under a seeded torch RNG, gallery =
8 true rows + 64 decoy rows embedded by the reference + Gaussian rows).

BASELINE.json tolerances, asserted STRICTLY (no near-tie exemptions) on the reference's own noise tensors:
  * embedding cosine >= 0.999 for every sample,
  * top-1 agreement >= 99.5 % per regime,
  * vote counts bit-exact wherever predictions agree (a disagreeing sample moves exactly one vote),
  * Smooth.certify through the drop-in API returns the reference's prediction and its certified radius within 1 %.
Runs at BASELINE config 2's chunking (125 samples per program run, ArcFace once per 2 chunks) and at chunk 8."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

N_GALLERY = 5000
REGIMES = ("iso", "aniso")


def _load(tag):
    path = os.path.join(GOLDEN, f"votes_{tag}.npz")
    if not os.path.isfile(path):
        pytest.fail(f"{path} missing: run `python -m oracle.make_golden_votes` in the build container")
    return np.load(path)


@pytest.fixture(scope="module")
def fixture(golden, models):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from oracle import fixtures
    rows = torch.from_numpy(np.load(os.path.join(GOLDEN, "votes_gallery.npz"))["rows"])
    gallery = fixtures.synthetic_gallery(rows, N_GALLERY)
    g_sd, f_sd = models
    dirs = torch.from_numpy(golden["dirs"])
    z = torch.from_numpy(golden["w_all"][0:1])
    return g_sd, f_sd, dirs, gallery, z


@pytest.fixture(scope="module")
def engine125(fixture):
    from certifyingfacerecognition_b200.engine import Engine
    g_sd, f_sd, dirs, gallery, z = fixture
    return Engine(g_sd, f_sd, dirs, gallery, chunk=125, frm_group=2)


def _run(eng, z, v, lo=0, hi=None):
    noise = torch.from_numpy(v["noise"][lo:hi])
    n = noise.shape[0]
    counts, ex = eng.sample_votes(z, torch.zeros(1, 5), torch.from_numpy(v["sigma"]), n, noise=noise, want_pred=True,
                                  want_emb=True)
    torch.cuda.synchronize()
    return counts.cpu(), ex["pred"].cpu().long(), ex["emb"].cpu()


@pytest.mark.parametrize("tag", REGIMES)
def test_strict_top1_cosine_and_counts_vs_reference(fixture, engine125, tag):
    g_sd, f_sd, dirs, gallery, z = fixture
    v = _load(tag)
    n0 = int(v["n0"])
    total = v["noise"].shape[0]
    assert total >= n0 + 1000
    counts, pred, emb = _run(engine125, z, v)
    pref = torch.from_numpy(v["pred"]).long()
    eref = torch.from_numpy(v["emb"])
    cos = F.cosine_similarity(emb, eref)
    agree = pred == pref
    rate = agree.float().mean().item()
    print(f"[{tag}] n={total} strict top-1 agreement {rate:.4%} ({int((~agree).sum())} flips), "
          f"cosine min {cos.min().item():.6f}, median emb L2 err {(emb - eref).norm(dim=1).median().item():.4f}, "
          f"distinct voted rows {len(torch.unique(pref))}")
    assert cos.min().item() >= 0.999
    assert len(torch.unique(pref)) >= 5                       # the decoys draw votes: not a trivial tally
    assert rate >= 0.995, f"strict top-1 agreement {rate:.4%} < 99.5 %"
    # integer work is exact: our tally is the histogram of our predictions, and it differs from the reference's tally by
    # exactly the disagreeing samples (each moves one vote from the reference's row to ours)
    assert torch.equal(counts, torch.bincount(pred, minlength=N_GALLERY))
    cref = torch.bincount(pref, minlength=N_GALLERY)
    moved = torch.bincount(pred[~agree], minlength=N_GALLERY) - torch.bincount(pref[~agree], minlength=N_GALLERY)
    assert torch.equal(counts - cref, moved)
    assert torch.equal(torch.bincount(pref[n0:], minlength=N_GALLERY).double(), torch.from_numpy(v["counts"]))
    # flips are near-ties of the reference itself (diagnostic, not an exemption): its own margin d2 - d1
    if (~agree).any():
        margin = torch.from_numpy(v["d2"] - v["d1"])[~agree]
        print(f"[{tag}] reference margins of the flipped samples: {sorted(round(m, 4) for m in margin.tolist())}")


@pytest.mark.parametrize("tag", REGIMES)
def test_certify_through_the_api_matches_reference_radius(fixture, tag):
    """Smooth.certify of the drop-in classes on the reference's noise: same prediction, radius within 1 %."""
    from certifyingfacerecognition_b200.models.smoothing_model import WrappedModel
    from certifyingfacerecognition_b200.smoothing import L2Certificate, Smooth
    g_sd, f_sd, dirs, gallery, z = fixture
    v = _load(tag)
    dev = torch.device("cuda")
    n0, n = int(v["n0"]), v["noise"].shape[0] - int(v["n0"])
    model = WrappedModel(dirs.to(dev), "insightface", generator_state=g_sd, frm_state=f_sd,
                         latents=z, orig_embs=gallery, chunk=125)
    sigma = torch.from_numpy(v["sigma"]).to(dev)
    sm = Smooth(model, N_GALLERY, sigma, L2Certificate(1, device=dev))
    sm.inject_noise(torch.from_numpy(v["noise"]))
    pred, gap = sm.certify(z.to(dev), torch.zeros(1, 5, device=dev), torch.tensor([0], device=dev), n0, n,
                           float(v["alpha"]), 100, device=dev)
    radius = float(sigma.min()) * gap                          # certify.py:141
    print(f"[{tag}] certify -> ({pred}, gap {gap:.5f}, radius {radius:.5f}); reference ({int(v['cert_pred'])}, "
          f"{float(v['cert_gap']):.5f}, {float(v['cert_radius']):.5f})")
    assert pred == int(v["cert_pred"])
    assert int(v["cert_pred"]) == 0 and float(v["cert_gap"]) > 0     # the fixture exercises the estimation pass
    assert radius == pytest.approx(float(v["cert_radius"]), rel=1e-2)
    # predict (smooth.py:79-107) on the estimation-pass noise: same decision as the reference's counts give
    from oracle import mc_path as M
    sm.inject_noise(torch.from_numpy(v["noise"][n0:]))
    got = sm.predict(z.to(dev), torch.zeros(1, 5, device=dev), n, float(v["alpha"]), 100, device=dev)
    cref = v["counts"]
    top2 = cref.argsort()[::-1][:2]
    from scipy.stats import binomtest
    want = M.ABSTAIN if binomtest(int(cref[top2[0]]), int(cref[top2[0]] + cref[top2[1]]), 0.5).pvalue > float(v["alpha"]) \
        else int(top2[0])
    assert int(got) == want


def test_small_chunk_engine_agrees_with_reference_too(fixture):
    """Chunking only changes the order of the InstanceNorm partial sums: the chunk-8 engine passes the same bar on the
    first 200 isotropic samples."""
    from certifyingfacerecognition_b200.engine import Engine
    g_sd, f_sd, dirs, gallery, z = fixture
    v = _load("iso")
    eng = Engine(g_sd, f_sd, dirs, gallery, chunk=8)
    counts, pred, emb = _run(eng, z, v, 0, 200)
    pref = torch.from_numpy(v["pred"][:200]).long()
    assert F.cosine_similarity(emb, torch.from_numpy(v["emb"][:200])).min().item() >= 0.999
    assert (pred == pref).float().mean().item() >= 0.995
