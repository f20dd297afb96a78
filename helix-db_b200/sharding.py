"""Id-range sharding of one index across the GPUs of a box.

One process per GPU (torch.distributed).  Every shard owns a contiguous id range with its own vectors and
its own HNSW graph; every rank searches every query on its shard; ONE all-gather moves the per-shard
(ids, scores, count) blocks — packed into a single int32 buffer of 12*k+4 bytes per query — over
NCCL/NVLink, and every rank then selects the k smallest by (score, id) with the merge kernel
(hx_merge_topk_device).  The (score, id) rule is the reference's Candidate order (model.rs:41-61), so the
merged answer of exact per-shard scans is independent of the number of shards.

Only tensor plumbing lives here (torch is used for device memory and the collective); the merge itself is a
CUDA kernel behind the C ABI.  The packing helpers are device agnostic, which is what the CPU `gloo` tests
exercise.
"""
from __future__ import annotations

import torch


def shard_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous id range [lo, hi) of `rank` when n ids are split across `world` shards."""
    return rank * n // world, (rank + 1) * n // world


def split_candidates(cand_ids, first_id: int, n: int, world: int, rank: int):
    """Restricted search: the ascending candidate ids that fall into this rank's id range
    (an empty slice is RestrictedVectorCandidates::Empty for this shard)."""
    lo, hi = shard_range(n, world, rank)
    import numpy as np

    a = np.asarray(cand_ids, dtype=np.uint64)
    i0 = int(np.searchsorted(a, np.uint64(first_id + lo), side="left"))
    i1 = int(np.searchsorted(a, np.uint64(first_id + hi), side="left"))
    return a[i0:i1]


def pack_topk(ids: torch.Tensor, scores: torch.Tensor, counts: torch.Tensor, out: torch.Tensor | None = None):
    """[Q,k] int64 ids, [Q,k] float32 scores, [Q] int32 counts -> [Q, 3k+1] int32 (bit-preserving)."""
    Q, k = ids.shape
    if out is None:
        out = torch.empty((Q, 3 * k + 1), dtype=torch.int32, device=ids.device)
    out[:, :2 * k] = ids.contiguous().view(torch.int32).view(Q, 2 * k)
    out[:, 2 * k:3 * k] = scores.contiguous().view(torch.int32)
    out[:, 3 * k] = counts
    return out


def unpack_topk(apack: torch.Tensor, k: int):
    """[S, Q, 3k+1] int32 -> ([S,Q,k] int64, [S,Q,k] float32, [S,Q] int32)."""
    S, Q, _ = apack.shape
    ids = apack[:, :, :2 * k].contiguous().view(torch.int64).view(S, Q, k)
    scores = apack[:, :, 2 * k:3 * k].contiguous().view(torch.float32)
    counts = apack[:, :, 3 * k].contiguous()
    return ids, scores, counts


def all_gather_topk(pack: torch.Tensor, world: int, out: torch.Tensor | None = None):
    """The single collective of the sharded path."""
    import torch.distributed as dist

    if out is None:
        out = torch.empty((world,) + tuple(pack.shape), dtype=pack.dtype, device=pack.device)
    # concatenated layout (world*Q, 3k+1): accepted by both NCCL and gloo
    dist.all_gather_into_tensor(out.view(world * pack.shape[0], pack.shape[1]), pack)
    return out


class ShardedSearcher:
    """Per-rank driver of the sharded HNSW path: local search -> pack -> all-gather -> merge kernel."""

    def __init__(self, hx, index, world: int, rank: int, Q: int, k: int, device):
        self.hx, self.ix, self.world, self.rank, self.Q, self.k = hx, index, world, rank, Q, k
        dev = device
        self.l_ids = torch.zeros((Q, k), dtype=torch.int64, device=dev)
        self.l_sc = torch.zeros((Q, k), dtype=torch.float32, device=dev)
        self.l_cnt = torch.zeros((Q,), dtype=torch.int32, device=dev)
        self.pack = torch.zeros((Q, 3 * k + 1), dtype=torch.int32, device=dev)
        self.apack = torch.zeros((world, Q, 3 * k + 1), dtype=torch.int32, device=dev)
        self.o_ids = torch.zeros((Q, k), dtype=torch.int64, device=dev)
        self.o_sc = torch.zeros((Q, k), dtype=torch.float32, device=dev)
        self.o_cnt = torch.zeros((Q,), dtype=torch.int32, device=dev)
        self.device_index = dev.index if dev.index is not None else 0

    def step(self, d_queries: torch.Tensor, params, stream_ptr: int):
        self.ix.search_device(d_queries.data_ptr(), self.Q, params, self.l_ids.data_ptr(), self.l_sc.data_ptr(),
                              self.l_cnt.data_ptr(), stream_ptr)
        pack_topk(self.l_ids, self.l_sc, self.l_cnt, self.pack)
        all_gather_topk(self.pack, self.world, self.apack)
        a_ids, a_sc, a_cnt = unpack_topk(self.apack, self.k)
        self.hx.merge_topk_device(self.device_index, a_ids.data_ptr(), a_sc.data_ptr(), a_cnt.data_ptr(), self.world,
                                  self.Q, self.k, self.o_ids.data_ptr(), self.o_sc.data_ptr(), self.o_cnt.data_ptr(),
                                  stream_ptr)
        self._keep = (a_ids, a_sc, a_cnt)   # keep the unpacked views alive until the merge has run
        return self.o_ids, self.o_sc, self.o_cnt
