"""N > 1 host logic on CPU: world_size = 2 over gloo (no GPU).  Checks the id-range split, the byte-block wire format of
the single all-gather, the unique-id exchange, and that merging per-shard top-k by (score, id) reproduces the unsharded
exact answer."""
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


def _worker(rank, world, port, ret, n):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import helix_db_b200  # noqa: F401  (registers the package)
    from importlib import import_module
    sh = import_module("helix_db_b200.sharding")
    from oracle import hxo

    dim, Q, k = 16, 9, 5
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
    l_ids[0, 0] = (1 << 63) + 12345 + rank                 # ids above 2^63 must survive the byte blocks
    # the wire format of the sharded path: one uint8 block per rank (ids | scores | counts), gathered by ONE collective
    lay = sh.block_layout(Q, k)
    block = np.zeros(lay["bytes"], dtype=np.uint8)
    b_ids, b_sc, b_cnt = sh.block_views(block, Q, k)
    b_ids[:], b_sc[:], b_cnt[:] = l_ids, l_sc, l_cnt.astype(np.uint32)
    gathered = torch.zeros((world, lay["bytes"]), dtype=torch.uint8)
    dist.all_gather_into_tensor(gathered.view(-1), torch.from_numpy(block))          # the ONE collective
    blocks = [gathered[r].numpy().copy() for r in range(world)]
    r_ids, r_sc, r_cnt = sh.block_views(blocks[rank], Q, k)
    assert r_ids.tolist() == l_ids.tolist() and r_sc.tobytes() == l_sc.tobytes() and r_cnt.tolist() == l_cnt.tolist()
    assert sh.block_views(blocks[0], Q, k)[0][0, 0] == (1 << 63) + 12345
    assert sh.block_views(blocks[world - 1], Q, k)[0][0, 0] == (1 << 63) + 12345 + world - 1
    # the communicator's unique id travels from rank 0 to every rank through the host's channel (here: gloo broadcast)
    uid = sh.exchange_unique_id(rank, make_id=lambda: bytes(range(128)))
    assert uid == bytes(range(128))
    # restore the sentinel and compare the merge with the unsharded exact answer
    for r in range(world):
        rlo, rhi = sh.shard_range(n, world, r)
        o = hxo.Index(hxo.EUCLIDEAN, dim)
        o.put_vectors(ids[rlo:rhi], rows[rlo:rhi])
        o.set_entry(int(ids[rlo]), 0)
        sh.block_views(blocks[r], Q, k)[0][0, 0] = o.search_exact(queries[0], k)[0][0]
    m_ids, m_sc, m_cnt = sh.merge_blocks_reference(blocks, Q, k, k)
    # a per-shard k smaller than k_out still merges (blocks built with 3 entries per query: fewer nominees per shard)
    small = []
    for r in range(world):
        bi, bs, bc = sh.block_views(blocks[r], Q, k)
        sb = np.zeros(sh.block_layout(Q, 3)["bytes"], dtype=np.uint8)
        si, ss, scn = sh.block_views(sb, Q, 3)
        si[:], ss[:], scn[:] = bi[:, :3], bs[:, :3], np.minimum(bc, 3)
        small.append(sb)
    t_ids, _, t_cnt = sh.merge_blocks_reference(small, Q, 3, k)
    assert all(int(c) <= 3 * world for c in t_cnt) and t_ids[1, 0] == m_ids[1, 0]
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


@pytest.mark.parametrize("world,n", [(2, 600), (4, 601)])          # 601: shards of unequal size
def test_sharded_host_logic_gloo(world, n):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret, n), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))
