"""Host-side logic (no GPU): the Smooth mirror on a duck-typed base classifier, Clopper-Pearson KATs, weight packing,
tile selection, geometry setup, CLI surface."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from certifyingfacerecognition_b200 import engine as E
from certifyingfacerecognition_b200.smoothing import L2Certificate, Smooth
from certifyingfacerecognition_b200.smoothing.smooth import lower_confidence_bound
from oracle import mc_path as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CPU = torch.device("cpu")


class FakeClassifier:
    """Votes class 1 when the first perturbation coefficient is > thr, else class 0."""

    def __init__(self, n=4, thr=0.0):
        self.n, self.thr = n, thr

    def eval(self):
        return self

    def __call__(self, z, p):
        out = torch.zeros(p.shape[0], self.n)
        hit = (p[:, 0, 0, 0] > self.thr).long()
        out[torch.arange(p.shape[0]), hit] = 1.0
        return out


def test_kats_match_oracle_and_survey_values(golden):
    for (na, n), p in zip(golden["kat_in"], golden["kat_p"]):
        assert lower_confidence_bound(int(na), int(n), 0.001) == pytest.approx(float(p), rel=1e-12, abs=1e-15)
        assert lower_confidence_bound(int(na), int(n), 0.001) == M.lower_confidence_bound(int(na), int(n), 0.001)
    assert lower_confidence_bound(100, 100, 0.001) == pytest.approx(0.9332543008, rel=1e-9)


def test_generic_loop_certify_predict_abstain():
    cert = L2Certificate(1, device=CPU)
    x = torch.zeros(1, 5)
    z = torch.zeros(1, 512)
    # threshold far above the noise: always class 0 -> certified with the closed-form gap
    s = Smooth(FakeClassifier(thr=10.0), 4, torch.tensor([0.1]), cert)
    pred, gap = s.certify(z, x, torch.tensor([0]), 20, 100, 0.001, 32, device=CPU)
    assert pred == 0 and gap == pytest.approx(M.compute_gap(M.lower_confidence_bound(100, 100, 0.001)))
    assert s.samples_classified == 120
    # wrong label -> early exit after n0 samples (smooth.py:66-68)
    s = Smooth(FakeClassifier(thr=10.0), 4, torch.tensor([0.1]), cert)
    assert s.certify(z, x, torch.tensor([2]), 20, 100, 0.001, 32, device=CPU) == (0, 0.0)
    assert s.samples_classified == 20
    # 50/50 votes -> abstain in certify (pABar < 0.5) and in predict (binomial test)
    torch.manual_seed(0)
    s = Smooth(FakeClassifier(thr=0.0), 4, torch.tensor([0.1]), cert)
    counts = s._sample_noise(z, x, 1000, 100, device=CPU)
    assert counts.dtype == np.float64 and counts.sum() == 1000 and 400 < counts[0] < 600
    label = int(counts.argmax())
    res = s.certify(z, x, torch.tensor([label]), 1000, 1000, 0.001, 100, device=CPU)
    assert res[1] == 0.0
    assert s.predict(z, x, 200, 0.001, 50, device=CPU) == Smooth.ABSTAIN
    s = Smooth(FakeClassifier(thr=10.0), 4, torch.tensor([0.1]), cert)
    assert s.predict(z, x, 64, 0.001, 64, device=CPU) == 0


def test_count_arr_matches_reference_semantics():
    s = Smooth(FakeClassifier(), 5, torch.tensor([0.1]), L2Certificate(1, device=CPU))
    assert s._count_arr(torch.tensor([3, 3, 1, 3, 0, 1]), CPU, 5).tolist() == [1, 2, 0, 3, 0]


def test_anisotropic_sigma_broadcast():
    cert = L2Certificate(1, device=CPU)
    sigma = 0.1 * torch.tensor(M.red_ellipse_mat_inv(), dtype=torch.float32)
    torch.manual_seed(1)
    noise = cert.sample_noise(torch.zeros(4000, 1, 1, 5), sigma)
    assert noise.shape == (4000, 1, 1, 5)
    assert torch.allclose(noise.reshape(-1, 5).std(0), sigma, rtol=0.1)


def test_pack_conv_weight_is_tap_major_k():
    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3)
    p = E.pack_conv_weight(w, cin_pad=16)
    assert p.shape == (2, 192)
    for tap, (ky, kx) in enumerate([(a, b) for a in range(3) for b in range(3)]):
        assert torch.equal(p[:, tap * 16:tap * 16 + 3], w[:, :, ky, kx])
        assert (p[:, tap * 16 + 3:(tap + 1) * 16] == 0).all()
    assert (p[:, 144:] == 0).all()


@pytest.mark.parametrize("packer", ["igemm", "halo"])
def test_upconv_phase_decomposition_is_exact(packer):
    """4 sub-pixel phases x (2x2 taps) on the low-res grid == nearest-x2 upsample + 3x3 conv (pad 1), borders included."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 16, 5, 7, generator=g).double()
    w = torch.randn(32, 16, 3, 3, generator=g).double()
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, padding=1)
    if packer == "igemm":
        wp, taps = E.pack_upconv_phases(w.float())
        wp = wp.double()[:, :4 * 16].reshape(4, 32, 4, 16)            # [phase][cout][tap][cin]
    else:
        wp, taps = E.pack_halo_upconv(w.float())
        wp = wp.double().reshape(4, 4, 32, 16).permute(0, 2, 1, 3)    # [phase][tap][cout][cin] -> [phase][cout][tap][cin]
    out = torch.zeros_like(ref)
    xp = F.pad(x, (1, 1, 1, 1))
    for ph, (a, b) in enumerate([(0, 0), (0, 1), (1, 0), (1, 1)]):
        acc = 0
        for t, (dy, dx) in enumerate(taps[ph]):
            patch = xp[:, :, 1 + dy:1 + dy + 5, 1 + dx:1 + dx + 7]
            acc = acc + torch.einsum("oc,bchw->bohw", wp[ph, :, t], patch)
        out[:, :, a::2, b::2] = acc
    assert (out - ref).abs().max().item() < 1e-5


def test_fused_upconv_equivalent_weight(models):
    g_sd, _ = models
    sd = {k: v.double() for k, v in g_sd.items() if k.startswith("synthesis.layer14.")}
    x = torch.randn(1, 64, 6, 6, generator=torch.Generator().manual_seed(1)).double()
    ref = M._upconv(x, sd, 14, literal=True)           # conv_transpose2d form, then blur
    weq = E.upconv_equiv_weight({k: v.float() for k, v in sd.items()}, 14).double()
    got = M._blur(F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), weq, padding=1))
    assert (got - ref).abs().max().item() < 1e-6         # weq is built in fp32 by the engine


def test_tile_for_covers_128_rows():
    for res in (1, 4, 7, 8, 14, 16, 28, 56, 112, 128, 1024):
        for n in (1, 8, 50, 250):
            tw, th, tn = E.tile_for(res, n)
            assert tw * th * tn == 128
    assert E.tile_for(14, 32) == (2, 2, 32)
    assert E.tile_for(112, 50) == (16, 8, 1)


def test_geometry_setup(tmp_path, monkeypatch):
    from certifyingfacerecognition_b200.attack_utils import gen_utils, proj_utils
    rs = np.random.RandomState(0)
    os.makedirs(tmp_path / "boundaries")
    for attr in proj_utils.ATTRS:
        v = rs.randn(1, 512)
        np.save(tmp_path / "boundaries" / f"stylegan_ffhq_{attr}_w_boundary.npy", v / np.linalg.norm(v))
    monkeypatch.chdir(tmp_path)
    proj, ell, dirs, red, files = proj_utils.get_projection_matrices("ffhq", "stylegan")
    assert dirs.shape == (512, 5) and len(files) == 5
    assert np.allclose(red, [4, 4, 25, 4, 1.5625], rtol=2e-3)            # 1/eps^2
    assert np.allclose(proj @ proj, proj, atol=1e-8) and np.allclose(proj @ dirs, dirs, atol=1e-8)
    mats = gen_utils.get_all_matrices(device=torch.device("cpu"))
    assert len(mats) == 7 and mats[3].shape == (512, 5)
    assert torch.allclose(mats[6], torch.tensor([0.25, 0.25, 0.04, 0.25, 0.64]), rtol=2e-3)
    # dropping attributes gives the reduced set and leaves the module's table alone: main_attack.py --attrs2drop asks
    # for the reduced matrices (the attack's search space) and then for all five directions (the engine's)
    red4 = gen_utils.get_all_matrices(["pose"], device=torch.device("cpu"))
    assert red4[3].shape == (512, 4) and red4[5].shape == (4,)
    assert torch.allclose(red4[6], torch.tensor([0.25, 0.25, 0.04, 0.64]), rtol=2e-3)
    assert torch.equal(red4[3], mats[3][:, [0, 1, 2, 4]])
    assert list(proj_utils.ATTRS) == ["age", "eyeglasses", "gender", "pose", "smile"]
    assert gen_utils.get_all_matrices(device=torch.device("cpu"))[3].shape == (512, 5)
    with pytest.raises(AssertionError):
        proj_utils.get_projection_matrices("ffhq", "stylegan", attrs2drop=["beard"])


def test_cli_surface_matches_reference():
    import certify
    p = certify.build_parser()
    a = p.parse_args(["--face-recog-model", "insightface", "--outfile", "o.tsv", "--sigma", "0.1"])
    assert (a.skip, a.max, a.batch_sz, a.N0, a.N, a.alpha, a.load_n_embs, a.anisotropic_sigma) == \
        (1, -1, 100, 100, 100000, 0.001, 1_000_000, False)
    with pytest.raises(SystemExit):
        p.parse_args(["--face-recog-model", "vgg", "--outfile", "o", "--sigma", "1"])
    row = "{}\t{}\t{}\t{}\t{:.3}\t{:.3}\t{}".format(3, 3, 3, 1, 1.9780096, 0.19780096, "0:00:01.5")
    assert row == "3\t3\t3\t1\t1.98\t0.198\t0:00:01.5"


def test_composite_upconv_blur_weights_are_exact():
    """blur o nearest-x2 o conv3x3 as per-phase 3x3 convs on the low-res grid with first/last-row weight sets and the
    border-column correction == the reference op order (UpConvBlock :665-676), borders included."""
    g = torch.Generator().manual_seed(3)
    H, W, ci, co = 5, 6, 4, 3
    x = torch.randn(2, ci, H, W, generator=g).double()
    weq = torch.randn(co, ci, 3, 3, generator=g).double()
    ref = M._blur(F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), weq, padding=1))
    base, corr = E.composite_upconv_weights(weq)
    base = base.double().reshape(8, 9, co, ci)
    corr = corr.double()
    xp = F.pad(x, (1, 1, 1, 1))
    out = torch.zeros_like(ref)
    for a in (0, 1):
        for b in (0, 1):
            for i in range(H):
                ws = a * 2 + b
                rci = 0
                if i == 0 and a == 0:
                    ws, rci = 4 + b, 1
                if i == H - 1 and a == 1:
                    ws, rci = 6 + b, 2
                wc = base[ws].reshape(3, 3, co, ci)
                for j in range(W):
                    v = torch.einsum("yxoi,niyx->no", wc, xp[:, :, i:i + 3, j:j + 3])
                    if j == 0 and b == 0:
                        v = v + torch.einsum("yoi,niy->no", corr[0, a, rci], xp[:, :, i:i + 3, 1])
                    if j == W - 1 and b == 1:
                        v = v + torch.einsum("yoi,niy->no", corr[1, a, rci], xp[:, :, i:i + 3, W])
                    out[:, :, 2 * i + a, 2 * j + b] = v
    assert (out - ref).abs().max().item() < 1e-5


def test_package_synthetic_weights_equal_the_test_fixtures(models):
    """bench.py / tools draw their random-init weights from certifyingfacerecognition_b200/synthetic.py (the product
    side may not import oracle/); they must be the very tensors the parity tests and golden vectors were made with."""
    from certifyingfacerecognition_b200 import synthetic
    from oracle import fixtures
    g_ref, f_ref = models
    g_sd, f_sd = synthetic.build_models()
    assert set(g_sd) == set(g_ref) and set(f_sd) == set(f_ref)
    for k in g_ref:
        assert torch.equal(g_sd[k], g_ref[k]), k
    for k in f_ref:
        if k.endswith("running_mean") or k.endswith("running_var"):
            assert torch.allclose(f_sd[k], f_ref[k].float(), rtol=0, atol=0), k      # stored as fp32 in the npz
        else:
            assert torch.equal(f_sd[k], f_ref[k]), k
    assert np.array_equal(synthetic.latents(7), fixtures.latents(7))
    rows = torch.randn(3, 512, generator=torch.Generator().manual_seed(0))
    assert torch.equal(synthetic.synthetic_gallery(rows, 10), fixtures.synthetic_gallery(rows, 10))


def test_sharded_gallery_key_merge_is_an_unsigned_min():
    """Partition C host logic: keys are uint64 bit patterns carried in int64; the merge must order them unsigned."""
    from certifyingfacerecognition_b200.gallery_shard import merge_keys_unsigned_min, rows_of_keys, shard_bounds
    g = torch.Generator().manual_seed(5)
    hi = torch.randint(0, 1 << 32, (4, 64), generator=g, dtype=torch.int64)
    lo = torch.randint(0, 1 << 32, (4, 64), generator=g, dtype=torch.int64)
    keys_u = (hi.numpy().astype(np.uint64) << np.uint64(32)) | lo.numpy().astype(np.uint64)
    stack = torch.from_numpy(keys_u.view(np.int64).copy())
    got = merge_keys_unsigned_min(stack).numpy().view(np.uint64)
    assert np.array_equal(got, keys_u.min(axis=0))
    assert np.array_equal(rows_of_keys(torch.from_numpy(got.view(np.int64).copy())).numpy(), (got & np.uint64(0xFFFFFFFF)).astype(np.int64))
    # balanced contiguous shards covering [0, n) exactly, lower ranks lower rows
    for n, w in ((10, 3), (5000, 8), (7, 7), (1_000_000, 8)):
        b = [shard_bounds(n, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_facenet_tables_and_synthetic_weights_are_consistent():
    """Config 4 (parity unpinned): the product-side layer table (models/facenet.py) and the oracle's independent one
    describe the same network, and the synthetic state dict drives the oracle forward to unit-norm embeddings."""
    from certifyingfacerecognition_b200 import synthetic
    from certifyingfacerecognition_b200.models import facenet as P
    from oracle import facenet as O
    assert [s for _, s in P.STEM] == [tuple(s[1:]) for s in O.STEM]
    for a, b in ((P.BLOCK35, O.BLOCK35), (P.BLOCK17, O.BLOCK17), (P.BLOCK8, O.BLOCK8), (P.MIXED_6A, O.MIXED_6A),
                 (P.MIXED_7A, O.MIXED_7A)):
        assert {k: [tuple(c) for c in v] for k, v in a.items()} == {k: [tuple(c) for c in v] for k, v in b.items()}
    sd = synthetic.facenet_weights()
    assert len(P.all_basic_convs()) == 111
    for p, (cin, cout, (kh, kw), _, _) in P.all_basic_convs():
        assert sd[p + "conv.weight"].shape == (cout, cin, kh, kw) and sd[p + "bn.running_var"].shape == (cout,)
        assert float(sd[p + "bn.running_var"].min()) > 0        # calibrated statistics loaded
    with torch.no_grad():
        e = O.forward(torch.randn(2, 3, 160, 160, generator=torch.Generator().manual_seed(0)) * 0.5, sd)
    assert e.shape == (2, 512) and torch.allclose(e.norm(dim=1), torch.ones(2), atol=1e-5)


def test_ncu_launch_summary_is_reproducible():
    """profiles/launches_r01e_summary.tsv is exactly what tools/summarize_ncu.py makes of the committed ncu launch list."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csv_path = os.path.join(root, "profiles", "launches_r01e.csv")
    want = open(os.path.join(root, "profiles", "launches_r01e_summary.tsv")).read().splitlines()
    got = subprocess.run([sys.executable, os.path.join(root, "tools", "summarize_ncu.py"), "launches", csv_path, "x"],
                         capture_output=True, text=True, check=True).stdout.splitlines()
    assert got[1:] == want[1:]                       # (line 0 carries the free-text command)
    shares = {ln.split("\t")[0]: float(ln.split("\t")[3]) for ln in got[2:]}
    halo = sum(v for k, v in shares.items() if "conv_halo_kernel" in k)
    igemm = sum(v for k, v in shares.items() if "conv_igemm_kernel" in k)
    assert 0.3 < halo < 0.5 and 0.3 < igemm < 0.5    # the two conv kernels are ~80 % of the GPU time of the bench process


def test_injected_noise_is_consumed_in_order_by_the_fused_path():
    """Parity runs replay the reference's own noise tensors: Smooth.inject_noise feeds the fused path row by row, first
    the selection pass, then the estimation pass; without injection the Philox offsets keep advancing."""
    class FakeFused:
        supports_fused_votes = True

        def __init__(self):
            self.calls = []

        def eval(self):
            return self

        def sample_votes(self, z, x, sigma, num, seed=0, sample_offset=0, noise=None):
            self.calls.append((num, sample_offset, None if noise is None else noise.clone()))
            c = torch.zeros(3, dtype=torch.int64)
            c[0] = num
            return c

    fake = FakeFused()
    s = Smooth(fake, 3, torch.tensor([0.1]), L2Certificate(1, device=CPU))
    noise = torch.arange(30 * 5, dtype=torch.float32).view(30, 5)
    s.inject_noise(noise)
    pred, gap = s.certify(torch.zeros(1, 512), torch.zeros(1, 5), torch.tensor([0]), 10, 20, 0.001, 8, device=CPU)
    assert pred == 0 and gap > 0
    assert [c[0] for c in fake.calls] == [10, 20]
    assert torch.equal(fake.calls[0][2], noise[:10]) and torch.equal(fake.calls[1][2], noise[10:30])
    with pytest.raises(ValueError):
        s._sample_noise(torch.zeros(1, 512), torch.zeros(1, 5), 1, 1, device=CPU)      # nothing left to replay
    s.inject_noise(None)
    s._sample_noise(torch.zeros(1, 512), torch.zeros(1, 5), 7, 7, device=CPU)
    assert fake.calls[-1][2] is None and fake.calls[-1][1] == 30                       # Philox offset after 30 draws


def test_reference_collection_recipe_lists_only_path_modules():
    """oracle/build_ref.py collects the certify path's modules only (no vendored TF trees) and never writes outside
    oracle/_ref (which is git-ignored)."""
    from oracle import build_ref
    assert build_ref.DST.endswith(os.path.join("oracle", "_ref"))
    assert all("tf_official" not in g and "examples" not in g for g in build_ref.GLOBS)
    ignore = open(os.path.join(ROOT, ".gitignore")).read()
    assert "oracle/_ref/" in ignore
    if os.path.isdir(build_ref.SRC):
        dst = build_ref.build()
        assert os.path.isfile(os.path.join(dst, "smoothing", "smooth.py"))
        assert open(os.path.join(dst, "smoothing", "smooth.py")).read() == \
            open(os.path.join(build_ref.SRC, "smoothing", "smooth.py")).read()


def test_committed_bench_lines_follow_the_contract():
    """The bench lines committed under profiles/ (what the round's numbers are quoted from) carry every key of the bench
    contract, and their derived numbers are consistent with each other."""
    import json
    with open(os.path.join(ROOT, "profiles", "bench_r02_final_1gpu.json")) as fh:
        d = json.loads(fh.read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["warmup"] >= 3
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] == pytest.approx(250 / (d["ms_per_step"] * 1e-3), rel=1e-6)            # 250 samples per step
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    for r in (d["roofline"], d["roofline_second_kernel"]):
        assert r["bound"] in ("hbm", "tensor") and r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
        assert 0 < r["share_of_step"] < 1 and r["launches_timed"] > 0
    assert d["roofline_second_kernel"]["traffic"] is not None and "traffic_r02_halo" in d["roofline_second_kernel"]["traffic_source"]
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and 0 < c["value"] < 100
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    with open(os.path.join(ROOT, "profiles", "bench_r02_reference_arm.json")) as fh:
        r = json.loads(fh.read().strip().splitlines()[-1])
    assert r["impl"] == "reference" and r["metric"] == d["metric"] and r["unit"] == d["unit"]
    assert r["e2e"]["h2d_bytes_per_step"] == 0 and r["cpu_baseline"]["kind"] == "reference"
    # N > 1: the headline is config 3 (samples sharded + NCCL sum), scaling "strong", identities-sharded beside it
    for n in (2, 4, 8):
        with open(os.path.join(ROOT, "profiles", f"cfg_r02_N{n}_bench.json")) as fh:
            m = json.loads(fh.read().strip().splitlines()[-1])
        assert m["n_gpus"] == n and m["scaling"] == "strong" and m["config"]["shard"] == "samples"
        assert m["value"] > 0.9 * n * 7000 and m["identities_sharded"]["value"] > m["value"] * 0.9
