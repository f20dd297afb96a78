"""Parity at a scale where the throughput builds leave their small-graph regime: a 100 000 x 768 (and a 20 000 x 1536)
DEVICE-BUILT graph, strict mode and the production default (`SearchParams::new`), cosine and Euclidean.

The small parity tests run on oracle-built graphs of a few thousand nodes, where the visited hash never spills, the
descent is two hops and every neighbour row sits in L2.  Here the device builds the graph (k_build.cu), the oracle imports
that adjacency and walks it with the reference's algorithm (search.rs:344-589 strict, :595-992 default); ids, order and
score bits must be identical for the warp-per-query build (large batch), the CTA-per-query build (small batch), the
service path (one query per call) and, on the full candidate set, the exact scan against `search_exact`.
"""
import numpy as np
import pytest

import helix_db_b200 as hx
from oracle import hxo

from test_gpu_policy import oracle_cfg

pytestmark = pytest.mark.gpu


def mixture(seed, n, dim, nq, ncent=256, latent=24):
    """Low-rank cluster mixture (the bench's recipe at a smaller size): centres in a `latent`-dimensional subspace."""
    rng = np.random.default_rng(seed)
    proj = (rng.standard_normal((latent, dim)) / np.sqrt(latent)).astype(np.float32)
    cent = rng.standard_normal((ncent, latent)).astype(np.float32)

    def draw(count):
        z = cent[rng.integers(0, ncent, count)] + 0.35 * rng.standard_normal((count, latent)).astype(np.float32)
        x = z @ proj + 0.05 * rng.standard_normal((count, dim)).astype(np.float32)
        return np.ascontiguousarray(x, dtype=np.float32)
    return draw(n), draw(nq), rng


@pytest.mark.parametrize("gm,om,n,dim", [(hx.Metric.Cosine, hxo.COSINE, 100_000, 768),
                                         (hx.Metric.Euclidean, hxo.EUCLIDEAN, 100_000, 768),
                                         (hx.Metric.Euclidean, hxo.EUCLIDEAN, 20_000, 1536)])
def test_device_built_graph_at_scale(gm, om, n, dim):
    rows, queries, rng = mixture(n + dim, n, dim, 256)
    if om == hxo.COSINE:
        rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    ids = np.arange(n, dtype=np.uint64) * 3 + 11                      # sparse external ids
    gpu = hx.VectorIndex(gm, hx.VectorIndexConfig("s", "embedding", dim))          # reference defaults m 16 / m0 32 / efc 200
    gpu.load_vectors(ids, rows)
    gpu.build(seed=7)
    g = gpu.download_graph()
    ora = hxo.Index(om, dim)
    ora.put_vectors(ids, rows)
    ora.import_graph(g["levels"], g["deg0"], g["nbr0"], g["layer0_stride"], g["upper_node"], g["upper_layer"],
                     g["upper_deg"], g["upper_nbr"], g["upper_stride"], g["entry_point"], g["max_layer"])
    k = 10
    strict = hx.SearchParams.strict(k)

    # strict, warp-per-query build (B = 256) vs the oracle's walk of the same graph
    gi, gs, gc = gpu.search_batch(queries, strict)
    nchk, hit = 96, 0
    for q in range(nchk):
        oi, os_ = ora.search(queries[q], k)
        assert gc[q] == len(oi) and gi[q, :gc[q]].tolist() == oi.tolist(), ("strict warp", q)
        assert gs[q, :gc[q]].tobytes() == os_.tobytes(), ("strict warp", q)
        ex, _ = ora.search_exact(queries[q], k)
        hit += len(set(oi.tolist()) & set(ex.tolist()))
    assert hit / float(nchk * k) >= 0.93                              # the walk is worth something on this graph
    # strict, CTA-per-query build (B = 24) and one query per call
    ci, cs, cc = gpu.search_batch(queries[:24], strict)
    assert ci.tolist() == gi[:24].tolist() and cs.tobytes() == gs[:24].tobytes() and cc.tolist() == gc[:24].tolist()
    r = gpu.search(queries[0], strict)
    assert [x.entity_id() for x in r] == gi[0, :gc[0]].tolist()
    # strict through the query service (concurrent one-query submissions)
    with gpu.service(k=k, ef=strict.ef()) as svc:
        tickets = [svc.submit(queries[q]) for q in range(64)]
        for q, t in enumerate(tickets):
            res = svc.wait(t)
            assert [x.entity_id() for x in res] == gi[q, :gc[q]].tolist(), ("service", q)
            assert np.array([x.score() for x in res], dtype=np.float32).tobytes() == gs[q, :gc[q]].tobytes(), ("service", q)

    # the exact scan over the whole id set == the oracle's search_exact (ids, order, score bits)
    allc = hx.RestrictedVectorCandidates(ids)
    ei, es, ec = gpu.search_restricted_batch(queries[:16], strict, allc)
    for q in range(16):
        xi, xs = ora.search_exact(queries[q], k)
        assert ei[q, :ec[q]].tolist() == xi.tolist() and es[q, :ec[q]].tobytes() == xs.tobytes(), ("exact", q)

    # production default: fingerprints projected on the device (bit-equality with the oracle's projection is covered by
    # test_policy_modes_match_the_oracle); the oracle gets the same 64-bit words
    planes = rng.standard_normal((64, dim)).astype(np.float32)
    gpu.set_simhash_planes(planes)
    gpu.compute_simhash()
    bits = gpu.download_simhash(0, n)
    for probe in (0, n // 2, n - 1):
        assert int(bits[probe]) == hxo.simhash_from_planes(planes, rows[probe])
    ora.put_simhash(ids, bits)
    qsim = np.array([hxo.simhash_from_planes(planes, q) for q in queries[:nchk]], dtype=np.uint64)
    gpu.set_simhash_config()
    new = hx.SearchParams.new(k)
    cfg = oracle_cfg(new)
    pi, ps, pc = gpu.search_ex(queries[:nchk], new, query_simhash=qsim)              # CTA build below #SMs ...
    wi, ws, wc = gpu.search_ex(queries, new, query_simhash=np.concatenate(
        [qsim, np.array([hxo.simhash_from_planes(planes, q) for q in queries[nchk:]], dtype=np.uint64)]))   # ... warp build above
    assert wi[:nchk].tolist() == pi.tolist() and ws[:nchk].tobytes() == ps.tobytes() and wc[:nchk].tolist() == pc.tolist()
    for q in range(nchk):
        oi, os_, _, _ = ora.search_policy(queries[q], k, new.ef(), cfg, int(qsim[q]))
        assert pc[q] == len(oi) and pi[q, :pc[q]].tolist() == oi.tolist(), ("default", q)
        assert ps[q, :pc[q]].tobytes() == os_.tobytes(), ("default", q)
