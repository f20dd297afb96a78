"""Id-range sharding of one index across the GPUs of a box — host side of hx_shard_group (C ABI).

One rank per GPU.  Every shard owns a contiguous id range with its own vectors and its own HNSW graph (or just the rows for
the exhaustive paths); every rank searches every query on its shard; the search kernels write ids | scores | counts
straight into the rank's send BLOCK; ONE ncclAllGather moves the blocks over NVLink (issued by the library itself: NCCL is
dlopen'ed by libhelix_b200.so, no torch collective on the data path) and the merge kernel selects the k smallest by
(score, id) from the gathered blocks on every rank.  The (score, id) rule is the reference's Candidate order
(model.rs:41-61), so the merged answer of exact per-shard scans is independent of the number of shards.

This module only (a) splits ids / candidates by range, (b) ships the 128-byte NCCL unique id from rank 0 to the other
ranks through whatever channel the host has (torch.distributed here; a Rust host would use its own RPC), and (c) wraps the
C entry points.  `block_layout` / `block_views` restate the wire format of one block for the CPU tests.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

HNSW, DENSE = 0, 1            # hx_search_sharded `path`
UNIQUE_ID_BYTES = 128


def shard_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous id range [lo, hi) of `rank` when n ids are split across `world` shards."""
    return rank * n // world, (rank + 1) * n // world


def split_candidates(cand_ids, first_id: int, n: int, world: int, rank: int):
    """Restricted search: the ascending candidate ids that fall into this rank's id range
    (an empty slice is RestrictedVectorCandidates::Empty for this shard).  hx_search_restricted_sharded does the same
    slicing inside the library; this is its host-side statement."""
    lo, hi = shard_range(n, world, rank)
    a = np.asarray(cand_ids, dtype=np.uint64)
    i0 = int(np.searchsorted(a, np.uint64(first_id + lo), side="left"))
    i1 = int(np.searchsorted(a, np.uint64(first_id + hi), side="left"))
    return a[i0:i1]


def _al8(x: int) -> int:
    return (x + 7) & ~7


def block_layout(B: int, k: int) -> dict:
    """One rank's block for B queries with k entries each: ids u64[B*k] | scores f32[B*k] | counts u32[B], every field
    padded to 8 bytes (csrc/hx_shard.cu: block_layout)."""
    off_sc = _al8(B * k * 8)
    off_cnt = off_sc + _al8(B * k * 4)
    return {"off_scores": off_sc, "off_counts": off_cnt, "bytes": off_cnt + _al8(B * 4)}


def block_views(buf: np.ndarray, B: int, k: int):
    """(ids[B,k] u64, scores[B,k] f32, counts[B] u32) views of one uint8 block."""
    lay = block_layout(B, k)
    ids = buf[:B * k * 8].view(np.uint64).reshape(B, k)
    sc = buf[lay["off_scores"]:lay["off_scores"] + B * k * 4].view(np.float32).reshape(B, k)
    cnt = buf[lay["off_counts"]:lay["off_counts"] + B * 4].view(np.uint32)
    return ids, sc, cnt


def merge_blocks_reference(blocks, B: int, k_in: int, k_out: int):
    """numpy statement of the merge kernel over gathered blocks (test-side checker): k_out smallest by (score bits, id)."""
    out_ids = np.zeros((B, k_out), dtype=np.uint64)
    out_sc = np.zeros((B, k_out), dtype=np.float32)
    out_cnt = np.zeros(B, dtype=np.uint32)
    views = [block_views(b, B, k_in) for b in blocks]
    for q in range(B):
        items = []
        for ids, sc, cnt in views:
            for j in range(min(int(cnt[q]), k_in)):
                items.append((int(sc[q, j:j + 1].view(np.uint32)[0]), int(ids[q, j])))
        items.sort()
        items = items[:k_out]
        out_cnt[q] = len(items)
        for j, (sb, i) in enumerate(items):
            out_ids[q, j] = i
            out_sc[q, j:j + 1].view(np.uint32)[0] = sb
    return out_ids, out_sc, out_cnt


def exchange_unique_id(rank: int, make_id=None, device=None) -> bytes:
    """Rank 0 obtains the communicator's unique id (hx_shard_unique_id) and every other rank receives it through
    torch.distributed (the process group the launcher already set up; gloo or nccl)."""
    import torch
    import torch.distributed as dist

    t = torch.zeros(UNIQUE_ID_BYTES, dtype=torch.uint8, device=device if device is not None else "cpu")
    if rank == 0:
        if make_id is None:
            from . import load_library, _ck
            buf = (C.c_uint8 * UNIQUE_ID_BYTES)()
            _ck(load_library().hx_shard_unique_id(buf, UNIQUE_ID_BYTES))
            raw = bytes(buf)
        else:
            raw = make_id()
        t.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


class ShardGroup:
    """hx_shard_group: this rank's shard bound to the other ranks by one NCCL communicator."""

    def __init__(self, index, n_shards: int, rank: int, unique_id: bytes | None):
        from . import load_library, _ck
        self.L, self.index, self.n_shards, self.rank = load_library(), index, n_shards, rank
        h = C.c_void_p()
        uid = (C.c_uint8 * UNIQUE_ID_BYTES).from_buffer_copy(unique_id) if unique_id is not None else None
        _ck(self.L.hx_shard_group_create(index.h, n_shards, rank, uid, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.hx_shard_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def search_device(self, path, d_queries_ptr, B, local_params, k_out, d_ids_ptr, d_scores_ptr, d_counts_ptr, stream_ptr=0):
        from . import _ck
        cp = local_params._c()
        _ck(self.L.hx_search_sharded_device(self.h, path, d_queries_ptr, B, C.byref(cp), k_out, d_ids_ptr, d_scores_ptr,
                                            d_counts_ptr, stream_ptr))

    def search(self, path, queries, local_params, k_out):
        from . import _ck
        q = np.ascontiguousarray(queries, dtype=np.float32)
        B = q.shape[0]
        ids = np.zeros((B, k_out), dtype=np.uint64)
        sc = np.zeros((B, k_out), dtype=np.float32)
        cnt = np.zeros(B, dtype=np.uint32)
        cp = local_params._c()
        _ck(self.L.hx_search_sharded(self.h, path, q.ctypes.data_as(C.POINTER(C.c_float)), B, C.byref(cp), k_out,
                                     ids.ctypes.data_as(C.POINTER(C.c_uint64)), sc.ctypes.data_as(C.POINTER(C.c_float)),
                                     cnt.ctypes.data_as(C.POINTER(C.c_uint32))))
        return ids, sc, cnt

    def search_restricted(self, queries, params, cand_ids):
        from . import _ck
        q = np.ascontiguousarray(queries, dtype=np.float32)
        B = q.shape[0]
        cp = params._c()
        k = cp.k
        ca = np.ascontiguousarray(cand_ids, dtype=np.uint64)
        ids = np.zeros((B, k), dtype=np.uint64)
        sc = np.zeros((B, k), dtype=np.float32)
        cnt = np.zeros(B, dtype=np.uint32)
        _ck(self.L.hx_search_restricted_sharded(self.h, q.ctypes.data_as(C.POINTER(C.c_float)), B, C.byref(cp),
                                                ca.ctypes.data_as(C.POINTER(C.c_uint64)), ca.size,
                                                ids.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                sc.ctypes.data_as(C.POINTER(C.c_float)),
                                                cnt.ctypes.data_as(C.POINTER(C.c_uint32))))
        return ids, sc, cnt

    def last_ms(self):
        from . import _ck
        a, b = C.c_float(0), C.c_float(0)
        _ck(self.L.hx_shard_group_last_ms(self.h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)
