"""N > 1 host logic on CPU: world_size = 2 over gloo (no GPU).  Checks the id-range split, the bit-preserving packing
of the single all-gather, and that merging per-shard top-k by (score, id) reproduces the unsharded exact answer."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _merge_reference(a_ids, a_sc, a_cnt, k):
    """numpy statement of hx_merge_topk_device (test-side checker only)."""
    S, Q, _ = a_ids.shape
    out_ids = np.zeros((Q, k), dtype=np.uint64)
    out_sc = np.zeros((Q, k), dtype=np.float32)
    out_cnt = np.zeros(Q, dtype=np.int32)
    for q in range(Q):
        items = []
        for s in range(S):
            for j in range(int(a_cnt[s, q])):
                items.append((np.float32(a_sc[s, q, j]).view(np.uint32).item(), int(a_ids[s, q, j])))
        items.sort()
        items = items[:k]
        out_cnt[q] = len(items)
        for j, (sb, i) in enumerate(items):
            out_ids[q, j] = i
            out_sc[q, j] = np.uint32(sb).view(np.float32)
    return out_ids, out_sc, out_cnt


def _worker(rank, world, port, ret):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import helix_db_b200  # noqa: F401  (registers the package)
    from importlib import import_module
    sh = import_module("helix_db_b200.sharding")
    from oracle import hxo

    n, dim, Q, k = 600, 16, 9, 5
    rng = np.random.default_rng(99)                       # same data on every rank
    rows = rng.integers(-2, 3, size=(n, dim)).astype(np.float32)
    queries = rng.integers(-2, 3, size=(Q, dim)).astype(np.float32)
    ids = np.arange(1000, 1000 + n, dtype=np.uint64)
    lo, hi = sh.shard_range(n, world, rank)
    assert (lo, hi) == (rank * n // world, (rank + 1) * n // world)
    # per-shard exact top-k from the oracle plays the role of the shard's device search
    ora = hxo.Index(hxo.EUCLIDEAN, dim)
    ora.put_vectors(ids[lo:hi], rows[lo:hi])
    ora.set_entry(int(ids[lo]), 0)
    l_ids = np.zeros((Q, k), dtype=np.uint64)
    l_sc = np.zeros((Q, k), dtype=np.float32)
    l_cnt = np.zeros(Q, dtype=np.int32)
    for q in range(Q):
        i, s = ora.search_exact(queries[q], k)
        l_ids[q, :len(i)], l_sc[q, :len(i)], l_cnt[q] = i, s, len(i)
    l_ids[0, 0] = (1 << 63) + 12345 + rank                 # ids above 2^63 must survive the int32 packing
    pack = sh.pack_topk(torch.from_numpy(l_ids.view(np.int64)), torch.from_numpy(l_sc), torch.from_numpy(l_cnt))
    assert pack.shape == (Q, 3 * k + 1) and pack.dtype == torch.int32
    apack = sh.all_gather_topk(pack, world)                 # the ONE collective
    a_ids, a_sc, a_cnt = sh.unpack_topk(apack, k)
    a_ids_n = a_ids.numpy().view(np.uint64)
    assert a_ids_n[rank].tolist() == l_ids.tolist()         # round trip is bit exact
    assert a_sc.numpy()[rank].tobytes() == l_sc.tobytes() and a_cnt.numpy()[rank].tolist() == l_cnt.tolist()
    assert a_ids_n[0, 0, 0] == (1 << 63) + 12345 and a_ids_n[world - 1, 0, 0] == (1 << 63) + 12345 + world - 1
    # restore the sentinel and compare the merge with the unsharded exact answer
    for r in range(world):
        rlo, rhi = sh.shard_range(n, world, r)
        o = hxo.Index(hxo.EUCLIDEAN, dim)
        o.put_vectors(ids[rlo:rhi], rows[rlo:rhi])
        o.set_entry(int(ids[rlo]), 0)
        a_ids_n[r, 0, 0] = o.search_exact(queries[0], k)[0][0]
    m_ids, m_sc, m_cnt = _merge_reference(a_ids_n, a_sc.numpy(), a_cnt.numpy(), k)
    full = hxo.Index(hxo.EUCLIDEAN, dim)
    full.put_vectors(ids, rows)
    full.set_entry(int(ids[0]), 0)
    for q in range(Q):
        ei, es = full.search_exact(queries[q], k)
        assert m_ids[q, :m_cnt[q]].tolist() == ei.tolist() and m_sc[q, :m_cnt[q]].tobytes() == es.tobytes()
    # candidate split for the restricted path
    cand = np.arange(1000, 1000 + n, 7, dtype=np.uint64)
    part = sh.split_candidates(cand, 1000, n, world, rank)
    assert all(1000 + lo <= int(c) < 1000 + hi for c in part)
    tot = torch.tensor([len(part)], dtype=torch.int64)
    dist.all_reduce(tot)
    assert int(tot.item()) == len(cand)
    dist.barrier()
    dist.destroy_process_group()
    ret[rank] = True


def test_sharded_host_logic_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))
