"""Incremental mirror maintenance: the oracle's insert applied as ROW PATCHES (the dirty rows of the write,
memory_store.rs:105-130) must leave the device mirror answering exactly like a full re-hydration — and like the oracle."""
import struct

import numpy as np
import pytest

import helix_db_b200 as hx
from oracle import hxo
from hx_testutil import mirror_from_oracle
from test_gpu_parity import levels_for

pytestmark = pytest.mark.gpu


def _graph_rows(ora):
    """{(layer, node): [neighbour ids]} and the index state."""
    graph, state = ora.export_graph()
    rows = {}
    for layer, (nodes, offs, nbrs) in graph.items():
        for i, node in enumerate(nodes):
            rows[(int(layer), int(node))] = [int(x) for x in nbrs[offs[i]:offs[i + 1]]]
    return rows, state


def _apply_diff(gpu, before, after, state):
    by_layer = {}
    for key, nb in after.items():
        if before.get(key) != nb:
            by_layer.setdefault(key[0], []).append((key[1], nb))
    for layer, items in sorted(by_layer.items()):
        nodes = np.array([n for n, _ in items], dtype=np.uint64)
        offs = np.zeros(len(items) + 1, dtype=np.uint32)
        offs[1:] = np.cumsum([len(nb) for _, nb in items])
        flat = np.array([x for _, nb in items for x in nb], dtype=np.uint64)
        gpu.upsert_neighbor_rows(layer, nodes, offs, flat)
    if state is not None:
        gpu.set_entry(state[0], state[1])
    return sum(len(v) for v in by_layer.values())


@pytest.mark.parametrize("gm,om", [(hx.Metric.Euclidean, hxo.EUCLIDEAN), (hx.Metric.Cosine, hxo.COSINE)])
def test_insert_as_row_patches_equals_rehydration(gm, om):
    rng = np.random.default_rng(21)
    n0, n1, dim, k, ef = 1500, 300, 48, 10, 60
    rows = rng.standard_normal((n0 + n1, dim)).astype(np.float32)
    ids = np.arange(10, 10 + n0 + n1, dtype=np.uint64)
    lv = levels_for(n0 + n1, 8, 5)
    lv[n0 + 7] = max(lv) + 1                                   # a new node above the old top layer becomes the entry point
    ora = hxo.Index(om, dim, m=8, m0=16, ef_construction=60)
    for i in range(n0):
        ora.insert(int(ids[i]), rows[i], lv[i])
    gpu = hx.VectorIndex(gm, hx.VectorIndexConfig("m", "embedding", dim).with_m(8).with_m0(16).with_ef_construction(60))
    mirror_from_oracle(gpu, ora)
    gpu.set_version(7, 100)
    q = rng.standard_normal((120, dim)).astype(np.float32)
    p = hx.SearchParams.strict(k, ef)
    patched_rows = 0
    for b0 in range(n0, n0 + n1, 50):                          # six committed writes of 50 inserts each
        before, _ = _graph_rows(ora)
        sl = slice(b0, b0 + 50)
        for i in range(b0, b0 + 50):
            ora.insert(int(ids[i]), rows[i], lv[i])
        after, state = _graph_rows(ora)
        gpu.upsert_vectors(ids[sl], rows[sl])
        gpu.set_levels(ids[sl], np.array(lv[sl], dtype=np.uint16))
        patched_rows += _apply_diff(gpu, before, after, state)
        gpu.set_version(7, 100 + b0)
        gi, gs, gc = gpu.search_batch(q, p)
        oi, os_, oc, _, _ = ora.search_batch(q, k, ef, threads=4)
        assert gc.tolist() == oc.tolist() and gi.tolist() == oi.tolist() and gs.tobytes() == os_.tobytes(), b0
    assert gpu.version()[:2] == (7, 100 + n0 + n1 - 50) and gpu.version()[2] > 0
    assert patched_rows > n1                                    # the inserts rewired existing rows, not just their own
    # a full re-hydration of the final state answers identically (large batch: the warp-per-query build)
    fresh = hx.VectorIndex(gm, hx.VectorIndexConfig("f", "embedding", dim).with_m(8).with_m0(16).with_ef_construction(60))
    mirror_from_oracle(fresh, ora)
    big = rng.standard_normal((400, dim)).astype(np.float32)
    a, b = gpu.search_batch(big, p), fresh.search_batch(big, p)
    assert a[0].tolist() == b[0].tolist() and a[1].tobytes() == b[1].tobytes()
    cand = hx.RestrictedVectorCandidates(ids[::3].copy())
    a, b = gpu.search_restricted_batch(big[:40], p, cand), fresh.search_restricted_batch(big[:40], p, cand)
    assert a[0].tolist() == b[0].tolist() and a[1].tobytes() == b[1].tobytes()
    # overwrite in place: same id, new vector
    newv = rng.standard_normal((1, dim)).astype(np.float32)
    gpu.upsert_vectors(ids[5:6], newv)
    assert np.array_equal(gpu.download_vectors(5, 1)[1][0], newv[0])
    # an absent id inside the mirrored range cannot be patched in
    with pytest.raises(hx.HelixDbError) as e:
        gpu.upsert_vectors(np.array([3], dtype=np.uint64), newv)
    assert e.value.variant == "Unsupported"
    gpu.close()
    fresh.close()


def test_delete_and_hot_lane_rows():
    rng = np.random.default_rng(22)
    n, dim, k = 800, 16, 5
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    ids = np.arange(n, dtype=np.uint64)
    ora = hxo.Index(hxo.EUCLIDEAN, dim, m=8, m0=16, ef_construction=60)
    for i, l in zip(range(n), levels_for(n, 8, 9)):
        ora.insert(int(ids[i]), rows[i], l)
    gpu = hx.VectorIndex(hx.Metric.Euclidean, hx.VectorIndexConfig("d", "embedding", dim).with_m(8).with_m0(16)
                         .with_ef_construction(60))
    mirror_from_oracle(gpu, ora)
    _, state = ora.export_graph()
    victim = next(int(i) for i in ids if int(i) != state[0] and ora.node_level(int(i)) == 0)
    before, _ = _graph_rows(ora)
    # the Rust side repairs the victim's neighbours (delete_from_layer); here: drop the victim from every row
    for (layer, node), nb in before.items():
        if node == victim:
            ora.put_neighbors(layer, node, [])
        elif victim in nb:
            ora.put_neighbors(layer, node, [x for x in nb if x != victim])
    after, _ = _graph_rows(ora)
    gpu.delete_vectors([victim])
    _apply_diff(gpu, before, {kk: v for kk, v in after.items() if kk[1] != victim}, None)
    q = rng.standard_normal((60, dim)).astype(np.float32)
    p = hx.SearchParams.strict(k, 40)
    gi, gs, gc = gpu.search_batch(q, p)
    oi, os_, oc, _, _ = ora.search_batch(q, k, 40, threads=4)
    assert gi.tolist() == oi.tolist() and gs.tobytes() == os_.tobytes() and victim not in set(gi.flatten().tolist())
    cand_with = np.array(sorted({victim, 1, 2, 3, 50, 60}), dtype=np.uint64)
    ri, rs, rc = gpu.search_restricted_batch(q[:5], hx.SearchParams.strict(6), hx.RestrictedVectorCandidates(cand_with))
    for b in range(5):
        ei, es = ora.search_restricted(q[b], 6, cand_with[cand_with != victim])
        assert ri[b, :rc[b]].tolist() == ei.tolist() and rs[b, :rc[b]].tobytes() == es.tobytes()
    # [0x13] hot-lane rows: byte-identical copies pass, a stale copy is corruption, a new node is inserted from its row
    def item(v):
        return struct.pack("<f", 0.0) + np.asarray(v, dtype=np.float32).tobytes()   # Euclidean bias header 0.0
    gpu.load_upper_vector_rows([4, 9], item(rows[4]) + item(rows[9]))
    stale = rows[4].copy()
    stale[0] += 1.0
    with pytest.raises(hx.HelixDbError) as e:
        gpu.load_upper_vector_rows([4], item(stale))
    assert e.value.variant == "InvariantViolation"
    fresh_vec = rng.standard_normal(dim).astype(np.float32)
    gpu.load_upper_vector_rows([n + 5], item(fresh_vec))
    assert np.array_equal(gpu.download_vectors(n, 1)[1][0], fresh_vec)
    gpu.close()
