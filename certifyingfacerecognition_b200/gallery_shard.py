"""Gallery sharded over ranks (partition C; config 5: 1 M identities).

The reference keeps the whole gallery on one device and takes `argmax(softmax(-cdist(e, G)))` (smoothing_model.py:56-61,
smooth.py:135).  Here rank r holds rows [lo, hi) of the N-row gallery; the query embeddings are replicated (every rank
runs the same StyleGAN + ArcFace samples, or receives them by all-gather), each rank reduces its rows to ONE 64-bit
key per query (`cfr_match_keys` / `cfr_matcher_keys`), the keys are all-gathered (8 bytes x b x R) and merged by an
unsigned minimum -- which is the global nearest row with the reference's first-index tie-break, because the low word
of a key is the GLOBAL row index.  No gallery row ever leaves its rank."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib as L

_SIGN = -(1 << 63)          # int64 view of 0x8000000000000000: x ^ _SIGN maps unsigned order onto signed order
TC_MATCH_MIN_ROWS = 32768   # shards at least this large use the tensor-core matcher


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced row range of `rank` (lower ranks hold lower global rows, so ties resolve to the lowest row)."""
    return rank * n // world, (rank + 1) * n // world


def merge_keys_unsigned_min(stack: torch.Tensor) -> torch.Tensor:
    """[R, b] int64 bit patterns of uint64 keys -> [b] unsigned minimum over R (torch has no uint64 compare)."""
    return (stack ^ _SIGN).min(dim=0).values ^ _SIGN


def rows_of_keys(keys: torch.Tensor) -> torch.Tensor:
    """Global gallery row (low 32 bits) of merged keys."""
    return keys & 0xFFFFFFFF


def allgather_merge(keys: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather the per-rank keys (8 bytes per query and rank) and take the unsigned minimum; identity without a group."""
    if group is None:
        return keys
    import torch.distributed as dist
    parts = [torch.empty_like(keys) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, keys, group=group)
    return merge_keys_unsigned_min(torch.stack(parts))


class ShardedGallery:
    def __init__(self, rows: torch.Tensor, lo: int, n_total: int, max_b: int, process_group=None,
                 tc_match: Optional[bool] = None, world: Optional[int] = None) -> None:
        if not rows.is_cuda:
            raise RuntimeError("ShardedGallery needs its rows on a CUDA device (no CPU fallback)")
        if rows.shape[0] == 0:
            raise ValueError("empty gallery shard: use fewer ranks than gallery rows")
        self.lib = L.load()
        self.rows = rows.contiguous().float()
        self.lo, self.n_total, self.max_b = int(lo), int(n_total), int(max_b)
        self.group = process_group
        self.device = rows.device
        # The two matchers encode their keys differently, so EVERY rank must pick the same one: decide from the
        # rank-independent floor(n_total / world) (shard sizes differ by at most one row), never from the local count.
        # (`world`: number of shards when several of them live in one process, as in the single-GPU tests.)
        if process_group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(process_group)
        world = max(1, int(world or 1))
        use_tc = bool(tc_match) if tc_match is not None else (self.n_total // world) >= TC_MATCH_MIN_ROWS
        if process_group is not None:
            import torch.distributed as dist
            flag = torch.tensor([int(use_tc), -int(use_tc)], device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=process_group)
            if int(flag[0]) != -int(flag[1]):
                raise RuntimeError("ShardedGallery: ranks disagree on the matcher (tc_match must be the same everywhere)")
        self.use_tc = use_tc
        self.matcher = None
        if use_tc:
            m = C.c_void_p()
            L.check(self.lib.cfr_matcher_create(L.ptr(self.rows), self.rows.shape[0], self.max_b, self._stream(), C.byref(m)))
            self.matcher = m

    def __del__(self):
        try:
            if getattr(self, "matcher", None):
                self.lib.cfr_matcher_destroy(self.matcher)
                self.matcher = None
        except Exception:
            pass

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def local_keys(self, emb: torch.Tensor) -> torch.Tensor:
        """[b,512] fp32 queries -> [b] int64 (uint64 bit patterns): best local row of every query."""
        emb = emb.contiguous().float()
        b = emb.shape[0]
        keys = torch.empty(b, dtype=torch.int64, device=self.device)
        for i in range(0, b, self.max_b):
            e = emb[i:i + self.max_b]
            k = keys[i:i + self.max_b]
            if self.matcher is not None:
                L.check(self.lib.cfr_matcher_keys(self.matcher, L.ptr(e), e.shape[0], self.lo, L.ptr(k), self._stream()))
            else:
                L.check(self.lib.cfr_match_keys(L.ptr(e), e.shape[0], L.ptr(self.rows), self.rows.shape[0], self.lo,
                                                L.ptr(k), self._stream()))
        return keys

    def merge(self, keys: torch.Tensor) -> torch.Tensor:
        """All-gather the per-rank keys and take the unsigned minimum (identity without a process group)."""
        return allgather_merge(keys, self.group)

    def match_vote(self, emb: torch.Tensor, counts: torch.Tensor, want_pred: bool = False) -> Optional[torch.Tensor]:
        """Global nearest row of every query, tallied into `counts` [N] int64 (identical on every rank)."""
        keys = self.merge(self.local_keys(emb))
        b = keys.shape[0]
        pred = torch.empty(b, dtype=torch.int32, device=self.device) if want_pred else None
        L.check(self.lib.cfr_vote_keys(L.ptr(keys), b, L.ptr(pred), L.ptr(counts), self._stream()))
        return pred
