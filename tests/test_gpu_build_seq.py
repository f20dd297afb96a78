"""HX_BUILD_SEQUENTIAL: the device build in one-insert-at-a-time mode must produce the reference's graph — every row of
every layer, the entry point and the top layer equal the oracle's insert_hnsw (mutation.rs:787-895) for the same insertion
order and scripted levels."""
import numpy as np
import pytest

import helix_db_b200 as hx
from oracle import hxo
from test_gpu_parity import levels_for

pytestmark = pytest.mark.gpu


def _rows_of_oracle(ora):
    graph, state = ora.export_graph()
    rows = {}
    for layer, (nodes, offs, nbrs) in graph.items():
        for i, node in enumerate(nodes):
            rows[(int(layer), int(node))] = [int(x) for x in nbrs[offs[i]:offs[i + 1]]]
    return rows, state


def _rows_of_device(gpu, ids):
    g = gpu.download_graph()
    rows = {}
    s0 = g["layer0_stride"]
    for slot in range(g["n"]):
        d = int(g["deg0"][slot])
        rows[(0, int(ids[slot]))] = [int(ids[x]) for x in g["nbr0"][slot * s0: slot * s0 + d]]
    su = g["upper_stride"]
    for r in range(len(g["upper_node"])):
        d = int(g["upper_deg"][r])
        rows[(int(g["upper_layer"][r]), int(ids[g["upper_node"][r]]))] = [int(ids[x]) for x in g["upper_nbr"][r * su: r * su + d]]
    return rows, (g["entry_point"], g["max_layer"])


@pytest.mark.parametrize("gm,om,n,dim,m,m0,efc", [
    (hx.Metric.Euclidean, hxo.EUCLIDEAN, 2500, 32, 8, 16, 40),
    (hx.Metric.Cosine, hxo.COSINE, 1500, 96, 16, 32, 200),        # reference defaults
    (hx.Metric.Euclidean, hxo.EUCLIDEAN, 600, 4, 4, 8, 16),       # integer grid: duplicate vectors, exact score ties
])
def test_sequential_build_equals_the_oracles_insert_hnsw(gm, om, n, dim, m, m0, efc):
    rng = np.random.default_rng(31)
    if dim == 4:
        rows = rng.integers(0, 4, size=(n, dim)).astype(np.float32) + 1.0
    else:
        rows = rng.standard_normal((n, dim)).astype(np.float32)
    ids = np.arange(100, 100 + n, dtype=np.uint64)
    lv = levels_for(n, m, 77)
    ora = hxo.Index(om, dim, m=m, m0=m0, ef_construction=efc)
    for i in range(n):
        ora.insert(int(ids[i]), rows[i], lv[i])
    gpu = hx.VectorIndex(gm, hx.VectorIndexConfig("seq", "embedding", dim).with_m(m).with_m0(m0).with_ef_construction(efc))
    gpu.load_vectors(ids, rows)
    gpu.build(levels=np.array(lv, dtype=np.uint16), sequential=True)
    want, wstate = _rows_of_oracle(ora)
    got, gstate = _rows_of_device(gpu, ids)
    assert (int(gstate[0]), int(gstate[1])) == (int(wstate[0]), int(wstate[1]))
    assert set(got) == set(want)
    diff = [k for k in want if want[k] != got[k]]
    assert not diff, f"{len(diff)} of {len(want)} rows differ, first: {diff[0]} want {want[diff[0]]} got {got[diff[0]]}"
    # and therefore the searches agree bit for bit
    q = rng.standard_normal((50, dim)).astype(np.float32) if dim != 4 else (rng.random((50, dim)) * 4 + 0.5).astype(np.float32)
    gi, gs, gc = gpu.search_batch(q, hx.SearchParams.strict(5, 30))
    oi, os_, oc, _, _ = ora.search_batch(q, 5, 30, threads=4)
    assert gi.tolist() == oi.tolist() and gs.tobytes() == os_.tobytes()
    gpu.close()
