"""N > 1 path on CPU (gloo, world_size 2): MC samples of one identity are split across ranks by global sample
index, the int64 counts are summed by one all-reduce, and the result equals the single-rank tally exactly."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from certifyingfacerecognition_b200.smoothing import L2Certificate, Smooth


class FakeFused:
    """Stands in for WrappedModel.sample_votes: the vote of sample i is a pure function of (seed, i)."""
    supports_fused_votes = True

    def __init__(self, n=16):
        self.n = n

    def eval(self):
        return self

    def sample_votes(self, z, x, sigma, num, seed=0, sample_offset=0):
        idx = torch.arange(sample_offset, sample_offset + num, dtype=torch.int64)
        votes = (idx * 2654435761 + seed * 40503) % 7 % self.n
        return torch.bincount(votes, minlength=self.n).to(torch.int64)


class FakeFusedMulti(FakeFused):
    """adds the several-identities entry: identity g's votes depend on its latent's first coordinate"""

    def sample_votes_multi(self, z, x, sigma, nums, seed=0, sample_offsets=None):
        rows = []
        for g, (num, off) in enumerate(zip(nums, sample_offsets)):
            idx = torch.arange(off, off + num, dtype=torch.int64)
            votes = (idx * 2654435761 + seed * 40503 + int(z[g, 0]) * 7) % 5 % self.n
            votes = torch.where(idx % 3 == 0, votes, torch.full_like(votes, int(z[g, 0]) % self.n))   # a clear winner
            rows.append(torch.bincount(votes, minlength=self.n).to(torch.int64))
        return torch.stack(rows)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = Smooth(FakeFused(), 16, torch.tensor([0.1]), L2Certificate(1, device="cpu"), seed=5,
               process_group=dist.group.WORLD)
    c1 = s._sample_noise(None, None, 101, 10, device=torch.device("cpu"))
    c2 = s._sample_noise(None, None, 1000, 10, device=torch.device("cpu"))
    pred = s.certify(None, None, torch.tensor([int(c2.argmax())]), 101, 1000, 0.001, 10, device=torch.device("cpu"))
    if rank == 0:
        np.savez(out, c1=c1, c2=c2, pred=np.array(pred, dtype=np.float64))
    dist.destroy_process_group()


def test_sample_sharding_two_ranks_equals_single_rank(tmp_path):
    out = str(tmp_path / "r.npz")
    mp.start_processes(_worker, args=(2, _free_port(), out), nprocs=2, join=True, start_method="fork")
    got = np.load(out)
    ref = Smooth(FakeFused(), 16, torch.tensor([0.1]), L2Certificate(1, device="cpu"), seed=5)
    c1 = ref._sample_noise(None, None, 101, 10, device=torch.device("cpu"))
    c2 = ref._sample_noise(None, None, 1000, 10, device=torch.device("cpu"))
    assert np.array_equal(got["c1"], c1) and c1.sum() == 101
    assert np.array_equal(got["c2"], c2) and c2.sum() == 1000
    pred = ref.certify(None, None, torch.tensor([int(c2.argmax())]), 101, 1000, 0.001, 10, device=torch.device("cpu"))
    assert np.allclose(got["pred"], np.array(pred, dtype=np.float64))


def _many_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = Smooth(FakeFusedMulti(), 16, torch.tensor([0.1]), L2Certificate(1, device="cpu"), seed=5,
               process_group=dist.group.WORLD)
    z = torch.arange(6, dtype=torch.float32).view(6, 1).repeat(1, 512)
    res = s.certify_many(z, torch.zeros(1, 5), torch.tensor([0, 1, 9, 3, 4, 5]), 101, 1000, 0.001)
    if rank == 0:
        np.save(out, np.array(res, dtype=np.float64))
    dist.destroy_process_group()


def test_certify_many_two_ranks_equals_single_rank(tmp_path):
    """Group certification: selection passes of all identities in one sharded call + one all-reduce, estimation passes
    only for the identities whose selection matched the label (identity 2 is mislabelled: early exit).  (The single-rank
    run happens in a forked worker as well: torch compute in the pytest process would leave threads behind that the next
    test's fork() cannot survive.)"""
    out2, out1 = str(tmp_path / "many2.npy"), str(tmp_path / "many1.npy")
    mp.start_processes(_many_worker, args=(2, _free_port(), out2), nprocs=2, join=True, start_method="fork")
    mp.start_processes(_many_worker, args=(1, _free_port(), out1), nprocs=1, join=True, start_method="fork")
    got, want = np.load(out2), np.load(out1)
    assert np.allclose(got, want)
    assert tuple(want[2]) == (2.0, 0.0) and all(w[0] == g and w[1] > 0 for g, w in enumerate(want) if g != 2)


# ---- partition C: gallery rows sharded over ranks, 8-byte keys all-gathered and merged -------------------------------
def _exact_keys(emb, rows, lo):
    """Pure-torch restatement of cfr_match_keys: (fp32 bits of the squared distance << 32) | global row, minimum over
    the local rows (ties -> lowest row), as int64 bit patterns."""
    d2 = ((emb[:, None, :] - rows[None, :, :]) ** 2).sum(-1).float()
    bits = d2.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    keys = (bits << 32) | (torch.arange(rows.shape[0], dtype=torch.int64)[None, :] + lo)
    return keys.min(dim=1).values           # distances are >= 0, so these keys are positive: signed == unsigned order


def _gallery_fixture():
    g = torch.Generator().manual_seed(9)
    gal = torch.randn(41, 16, generator=g)
    gal[30] = gal[7]                        # duplicate rows in different shards: the lower global index must win
    emb = torch.cat([gal[[7, 30, 12, 40]] + 0.01 * torch.randn(4, 16, generator=g), gal[[7, 30]]])
    return gal, emb


def _shard_worker(rank, world, port, out):
    from certifyingfacerecognition_b200.gallery_shard import allgather_merge, rows_of_keys, shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gal, emb = _gallery_fixture()
    lo, hi = shard_bounds(gal.shape[0], world, rank)
    merged = allgather_merge(_exact_keys(emb, gal[lo:hi], lo), dist.group.WORLD)
    if rank == 1:                           # every rank holds the same merged result; check a non-zero one
        np.save(out, rows_of_keys(merged).numpy())
    dist.destroy_process_group()


def test_gallery_sharding_two_ranks_equals_single_rank(tmp_path):
    out = str(tmp_path / "rows.npy")
    mp.start_processes(_shard_worker, args=(2, _free_port(), out), nprocs=2, join=True, start_method="fork")
    gal, emb = _gallery_fixture()
    d = torch.cdist(emb, gal, compute_mode="donot_use_mm_for_euclid_dist")
    ref = (-d).argmax(dim=1)                # the reference's softmax(-d).argmax(1): first index on ties
    got = np.load(out)
    assert np.array_equal(got, ref.numpy())
    assert got[4] == 7 and got[5] == 7      # exact duplicates resolve to the lower global row
