"""Pin the CPU restatement (oracle/mc_path.py) against vectors produced by the UNMODIFIED reference
(tests/golden/reference_vectors.npz, written by oracle/make_golden.py)."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT
from oracle import mc_path as M


def _stats(t):
    f = t.float().flatten(1)
    return torch.stack([f.mean(1), f.std(1), f.abs().amax(1)], dim=1).numpy()


def test_geometry(golden):
    assert np.allclose(golden["red_ellipse_mat"], [4, 4, 25, 4, 1.5625], rtol=1e-3)
    assert np.allclose(golden["red_ellipse_mat_inv"], M.red_ellipse_mat_inv(), rtol=1e-3)
    d = golden["dirs"]
    assert d.shape == (5, 512)
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-5)


def test_kats(golden):
    for (na, n), p, gap in zip(golden["kat_in"], golden["kat_p"], golden["kat_gap"]):
        mine = M.lower_confidence_bound(int(na), int(n), 0.001)
        assert mine == pytest.approx(p, rel=1e-12, abs=1e-15)
        if 0 < p < 1:
            assert M.compute_gap(mine) == pytest.approx(gap, rel=1e-12)
    # SURVEY.md section 8c closed-form values
    assert M.lower_confidence_bound(100, 100, 0.001) == pytest.approx(0.9332543008, rel=1e-9)
    assert M.compute_gap(M.lower_confidence_bound(990, 1000, 0.001)) == pytest.approx(1.9780096, rel=1e-6)
    assert M.lower_confidence_bound(60, 100, 0.001) < 0.5


def test_fused_upconv_equivalence(models):
    """conv_transpose2d with the box-summed 4x4 kernel == nearest x2 + flipped 3x3 conv (borders too)."""
    g_sd, _ = models
    x = torch.randn(2, 64, 16, 16, generator=torch.Generator().manual_seed(0)).double()
    sd = {k: v.double() for k, v in g_sd.items() if k.startswith("synthesis.layer14.")}
    a = M._upconv(x, sd, 14, literal=True)
    b = M._upconv(x, sd, 14, literal=False)
    assert (a - b).abs().max().item() < 1e-11


def test_pipeline_against_reference(golden, models):
    g_sd, f_sd = models
    w_in = torch.from_numpy(golden["w_in"])
    wp = M.truncation(w_in, g_sd)
    assert np.allclose(wp.numpy(), golden["wp"], atol=1e-6)
    got = {}
    with torch.no_grad():
        raw = M.synthesis(wp, g_sd, literal=True, tap=lambda k, t: got.__setitem__(k, _stats(t)))
        img = M.postprocess(raw)
        img112 = M.transform(img)
        blocks = {}
        emb = M.iresnet50(img112, f_sd, tap=lambda k, t: blocks.__setitem__(k, _stats(t)))
    ls = np.stack([got[f"layer{i}"] for i in range(M.NUM_LAYERS)])
    assert np.allclose(ls, golden["layer_stats"], rtol=2e-3, atol=2e-4)
    sl = slice(0, 1024, 64)
    assert np.allclose(img[:, :, sl, sl].numpy(), golden["image_sub"], atol=2e-4)
    assert np.allclose(img112[0].numpy(), golden["img112_0"], atol=5e-4)
    ref_emb = torch.from_numpy(golden["emb"])
    cos = F.cosine_similarity(emb, ref_emb).min().item()
    assert cos > 0.99999, cos
    assert (emb - ref_emb).abs().max().item() < 2e-2


def test_certify_against_reference(golden, models):
    g_sd, f_sd = models
    dirs = torch.from_numpy(golden["dirs"])
    gal = torch.from_numpy(golden["gallery"])
    z = torch.from_numpy(golden["w_all"][0:1])
    x = torch.zeros(1, 5)
    classify = lambda p: M.wrapped_forward(z, p, dirs, gal, g_sd, f_sd)
    for tag in ("iso", "aniso"):
        sigma = torch.from_numpy(golden[tag + "_sigma"])
        torch.manual_seed(1234)
        pred, gap = M.certify(classify, x, 0, sigma, 4, 12, 0.001, 4, gal.shape[0])
        assert pred == int(golden[tag + "_pred"])
        assert gap == pytest.approx(float(golden[tag + "_gap"]), rel=1e-9, abs=1e-12)
    torch.manual_seed(99)
    pred, gap = M.certify(classify, x, 3, torch.tensor([0.1]), 4, 12, 0.001, 4, gal.shape[0])
    assert (pred, gap) == (int(golden["pred_wrong"]), float(golden["gap_wrong"]))


def test_count_arr_and_predict():
    preds = torch.tensor([3, 3, 1, 3, 0, 1])
    assert M.count_arr(preds, 5).tolist() == [1, 2, 0, 3, 0]
    probs = torch.zeros(1, 4)

    def classify(p):
        out = torch.zeros(p.shape[0], 4)
        out[:, 2] = 1.0
        return out
    assert M.predict(classify, torch.zeros(1, 5), torch.tensor([0.1]), 40, 0.001, 16, 4) == 2


def test_mapping_network_against_reference():
    """oracle restatement of MappingModule vs the unmodified reference module
    (tests/golden/mapping_vectors.npz, written by oracle/make_golden_mapping.py)."""
    from certifyingfacerecognition_b200 import synthetic
    g = np.load(os.path.join(ROOT, "tests", "golden", "mapping_vectors.npz"))
    sd = synthetic.mapping_weights()
    z = M.preprocess_z(torch.from_numpy(g["z_raw"]))
    assert torch.allclose(z, torch.from_numpy(g["z"]), atol=1e-5)
    w = M.mapping(torch.from_numpy(g["z"]), sd)
    assert torch.allclose(w, torch.from_numpy(g["w"]), atol=2e-5, rtol=1e-5)


def test_port_reproduces_the_reference_vote_fixtures(golden, models):
    """tests/golden/votes_*.npz (2 x 1100 samples classified by the unmodified reference, oracle/make_golden_votes.py):
    the recorded certify results follow from the recorded predictions through the restated statistics, and the port's
    embeddings / predictions on a few of the recorded noise rows equal the reference's."""
    from oracle import fixtures
    g_sd, f_sd = models
    dirs = torch.from_numpy(golden["dirs"])
    z = torch.from_numpy(golden["w_all"][0:1])
    rows = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "votes_gallery.npz"))["rows"])
    assert rows.shape == (72, 512) and np.allclose(rows[:8].numpy(), golden["gallery"])
    gallery = fixtures.synthetic_gallery(rows, 5000)
    for tag in ("iso", "aniso"):
        v = np.load(os.path.join(ROOT, "tests", "golden", f"votes_{tag}.npz"))
        n0, alpha = int(v["n0"]), float(v["alpha"])
        pred = v["pred"]
        assert pred.shape[0] == n0 + 1000 and v["noise"].shape == (n0 + 1000, 5)
        # Smooth.certify (smooth.py:39-77) restated on the recorded predictions
        c0 = np.bincount(pred[:n0], minlength=5000)
        c1 = np.bincount(pred[n0:], minlength=5000)
        assert int(c0.argmax()) == 0 == int(v["cert_pred"])
        assert np.array_equal(c1.astype(np.float64), v["counts"])
        pbar = M.lower_confidence_bound(int(c1[0]), 1000, alpha)
        assert pbar >= 0.5 and M.compute_gap(pbar) == pytest.approx(float(v["cert_gap"]), rel=1e-9)
        assert float(v["cert_radius"]) == pytest.approx(float(v["sigma"].min()) * float(v["cert_gap"]), rel=1e-6)
        assert len(np.nonzero(c1)[0]) >= 10                        # mixed votes
        # the port on three recorded noise rows (the smallest-margin sample among them)
        idx = [0, 1, int(np.argmin(v["d2"] - v["d1"]))]
        p = torch.from_numpy(v["noise"][idx]).view(-1, 1, 1, 5)
        emb = M.lat2embs(M.perturb_latent(z, p, dirs), g_sd, f_sd, literal=True)
        ref = torch.from_numpy(v["emb"][idx])
        assert F.cosine_similarity(emb, ref).min().item() > 0.999999
        assert (emb - ref).norm(dim=1).max().item() < 2e-3
        d = torch.cdist(emb, gallery, compute_mode="donot_use_mm_for_euclid_dist")
        assert np.allclose(d.topk(2, dim=1, largest=False).values[:, 0].numpy(), v["d1"][idx], atol=2e-3)
        assert np.array_equal(M.compute_probs(emb[:2], gallery).argmax(1).numpy(), pred[idx[:2]])
