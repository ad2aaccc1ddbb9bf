"""SURVEY.md section 8e, partition C on the GPU: per-shard keys (exact fp32 matcher and tensor-core matcher) merged by
an unsigned minimum equal the unsharded match, including ties across shards; the engine-level sharded vote path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _merged_pred(emb, gallery, cuts, tc):
    from certifyingfacerecognition_b200.gallery_shard import ShardedGallery, merge_keys_unsigned_min, rows_of_keys
    keys = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        sh = ShardedGallery(gallery[lo:hi].cuda(), lo, gallery.shape[0], max_b=64, tc_match=tc)
        keys.append(sh.local_keys(emb.cuda()))
    torch.cuda.synchronize()
    return rows_of_keys(merge_keys_unsigned_min(torch.stack(keys))).cpu()


def test_exact_matcher_shards_equal_unsharded():
    g = torch.Generator().manual_seed(3)
    gallery = torch.randn(5000, 512, generator=g) * 1.5
    gallery[4100] = gallery[37]                          # duplicate rows in different shards
    gallery[2600] = gallery[2599]
    emb = torch.cat([gallery[[37, 4100, 2599, 2600, 4999, 0]], gallery[torch.randint(0, 5000, (90,), generator=g)]
                     + 0.3 * torch.randn(90, 512, generator=g)])
    ref = (-torch.cdist(emb, gallery, compute_mode="donot_use_mm_for_euclid_dist")).argmax(1)
    for cuts in ([0, 5000], [0, 2500, 5000], [0, 1, 2600, 4999, 5000]):
        got = _merged_pred(emb, gallery, cuts, tc=False)
        assert torch.equal(got, ref), cuts
    assert ref[1].item() == 37 and ref[3].item() == 2599


def test_tensor_core_matcher_shards_equal_unsharded():
    import ctypes as C
    from certifyingfacerecognition_b200 import _lib as L
    g = torch.Generator().manual_seed(4)
    n = 70000
    gallery = torch.randn(n, 512, generator=g)
    gallery[66000] = gallery[123]
    emb = torch.cat([gallery[[123, 66000, 69999]], gallery[torch.randint(0, n, (61,), generator=g)]
                     + 0.2 * torch.randn(61, 512, generator=g)])
    lib = L.load()
    m = C.c_void_p()
    gd, ed = gallery.cuda(), emb.cuda()
    stream = torch.cuda.current_stream().cuda_stream
    L.check(lib.cfr_matcher_create(L.ptr(gd), n, 64, stream, C.byref(m)))
    pred = torch.empty(64, dtype=torch.int32, device="cuda")
    counts = torch.zeros(n, dtype=torch.int64, device="cuda")
    L.check(lib.cfr_matcher_run(m, L.ptr(ed), 64, L.ptr(pred), L.ptr(counts), stream))
    torch.cuda.synchronize()
    lib.cfr_matcher_destroy(m)
    for cuts in ([0, 35000, 70000], [0, 33000, 66500, 70000]):
        got = _merged_pred(emb, gallery, cuts, tc=True)
        assert torch.equal(got, pred.cpu().long()), cuts
    assert pred[0].item() == 123 and pred[1].item() == 123 and pred[2].item() == 69999


def test_engine_sharded_votes_equal_unsharded(golden, models):
    from certifyingfacerecognition_b200.engine import Engine
    from certifyingfacerecognition_b200.gallery_shard import ShardedGallery, merge_keys_unsigned_min
    from certifyingfacerecognition_b200 import synthetic, _lib as L
    g_sd, f_sd = models
    dirs = torch.from_numpy(golden["dirs"])
    gallery = synthetic.synthetic_gallery(torch.from_numpy(golden["gallery"]), 999)
    eng = Engine(g_sd, f_sd, dirs, gallery, chunk=8)
    z = torch.from_numpy(golden["w_all"][0:1])
    x, sigma = torch.zeros(1, 5), torch.tensor([0.3])
    full, ex = eng.sample_votes(z, x, sigma, 20, seed=3, want_pred=True, want_emb=True)
    # two "ranks" emulated on one GPU: local keys per shard, merged as the all-gather would
    shards = [ShardedGallery(gallery[lo:hi].cuda(), lo, 999, max_b=8) for lo, hi in ((0, 499), (499, 999))]
    keys = merge_keys_unsigned_min(torch.stack([s.local_keys(ex["emb"]) for s in shards]))
    counts = torch.zeros(999, dtype=torch.int64, device="cuda")
    pred = torch.empty(20, dtype=torch.int32, device="cuda")
    L.check(eng.lib.cfr_vote_keys(L.ptr(keys), 20, L.ptr(pred), L.ptr(counts), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(pred, ex["pred"]) and torch.equal(counts, full)
    # and the one-call path (no process group: a single shard holding everything)
    whole = ShardedGallery(gallery.cuda(), 0, 999, max_b=8)
    c2, p2 = eng.sample_votes_sharded(whole, z, x, sigma, 20, seed=3, want_pred=True)
    torch.cuda.synchronize()
    assert torch.equal(c2, full) and torch.equal(p2, ex["pred"])


def test_matcher_choice_is_rank_independent_at_the_threshold():
    """65 535 rows over 2 ranks = shards of 32 767 and 32 768 rows, one below and one at the tensor-core
    threshold.  The two matchers encode keys differently, so the choice must come from floor(n_total / world) -- the same
    on every rank -- and the merged result must equal the unsharded exact match."""
    from certifyingfacerecognition_b200.gallery_shard import (ShardedGallery, TC_MATCH_MIN_ROWS, merge_keys_unsigned_min,
                                                              rows_of_keys, shard_bounds)
    g = torch.Generator().manual_seed(6)
    n = 2 * TC_MATCH_MIN_ROWS - 1
    gallery = torch.randn(n, 512, generator=g)
    emb = gallery[torch.randint(0, n, (32,), generator=g)] + 0.2 * torch.randn(32, 512, generator=g)
    bounds = [shard_bounds(n, 2, r) for r in range(2)]
    assert sorted(hi - lo for lo, hi in bounds) == [TC_MATCH_MIN_ROWS - 1, TC_MATCH_MIN_ROWS]
    shards = [ShardedGallery(gallery[lo:hi].cuda(), lo, n, max_b=32, world=2) for lo, hi in bounds]
    assert shards[0].use_tc == shards[1].use_tc
    keys = merge_keys_unsigned_min(torch.stack([s.local_keys(emb.cuda()) for s in shards]))
    torch.cuda.synchronize()
    ref = (-torch.cdist(emb, gallery, compute_mode="donot_use_mm_for_euclid_dist")).argmax(1)
    assert torch.equal(rows_of_keys(keys).cpu(), ref)
