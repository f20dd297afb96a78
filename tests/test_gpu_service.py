"""The reference's calling pattern (one query per call, many concurrent callers: read_index.rs:81-101) through hx_service,
per-query status (hx_search_batch), the unbounded tie stack, and device-buffer calls on concurrent streams.

Everything is compared bit for bit with the CPU oracle (ids, order, score bits).
"""
import threading

import numpy as np
import pytest

import helix_db_b200 as hx
from oracle import hxo
from hx_testutil import mirror_from_oracle
from test_gpu_parity import build_pair

pytestmark = pytest.mark.gpu


def _pairs(results):
    return [(r.entity_id(), np.float32(r.score()).tobytes()) for r in results]


@pytest.mark.parametrize("gm,om,n,dim", [(hx.Metric.Euclidean, hxo.EUCLIDEAN, 3000, 64),
                                         (hx.Metric.Cosine, hxo.COSINE, 1500, 768),
                                         (hx.Metric.Manhattan, hxo.MANHATTAN, 1200, 40)])
@pytest.mark.parametrize("ctas_per_sm", [1, 2, 3])
def test_service_bit_exact_blocking_and_async(gm, om, n, dim, ctas_per_sm):
    rng = np.random.default_rng(7)
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    gpu, ora = build_pair(gm, om, rows, m=8, m0=16, efc=60)
    nq, k, ef = 600, 10, 50
    queries = rng.standard_normal((nq, dim)).astype(np.float32)
    oi, os_, oc, _, _ = ora.search_batch(queries, k, ef, threads=4)
    expect = [list(zip(oi[q, :oc[q]].tolist(), [np.float32(x).tobytes() for x in os_[q, :oc[q]]])) for q in range(nq)]
    with gpu.service(k, ef, capacity=256, max_batch=64, ctas_per_sm=ctas_per_sm) as svc:
        info = svc.stats()
        assert info["ctas_per_sm"] >= 1 and info["rows_in_flight"] >= 1
        # blocking callers: 24 threads, one query per call
        got = [None] * nq
        errs = []

        def worker(t):
            try:
                for q in range(t, nq // 2, 24):
                    got[q] = _pairs(svc.search(queries[q]))
            except Exception as e:   # pragma: no cover
                errs.append(e)

        ths = [threading.Thread(target=worker, args=(t,)) for t in range(24)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        assert not errs, errs
        # async pair: keep up to 200 tickets in flight from one thread (what a tokio worker does with its tasks)
        pending = {}
        nxt = nq // 2
        while nxt < nq or pending:
            while nxt < nq and len(pending) < 200:
                pending[svc.submit(queries[nxt])] = nxt
                nxt += 1
            for t in list(pending):
                r = svc.poll(t)
                if r is not None:
                    got[pending.pop(t)] = _pairs(r)
        for q in range(nq):
            assert got[q] == expect[q], f"query {q} differs from the oracle"
        st = svc.stats()
        assert st["submitted"] == nq and st["completed"] == nq and st["launches"] <= nq
        # a consumed ticket is refused; an invalid query fails at submit, alone
        with pytest.raises(hx.HelixDbError):
            svc.poll(0)
        bad = queries[0].copy()
        bad[3] = np.nan
        with pytest.raises(hx.HelixDbError) as e:
            svc.search(bad)
        assert e.value.variant == "InvalidVectorComponent" and e.value.index == 3
        assert _pairs(svc.search(queries[5])) == expect[5]
        # the C++ host harness (host/hx_callers.cpp over host/vector_index.hpp): 64 concurrent callers, both styles
        from helix_db_b200 import callers
        for mode, threads in (("blocking", 0), ("tasks", 4)):
            rep, ci, cs, cc = callers.run(svc, gpu, queries, k, ef, 64, mode=mode, n_threads=threads, seconds=0.2)
            assert rep["rc"] == 0 and rep["errors"] == 0 and rep["completed"] >= nq
            assert cc.tolist() == oc.tolist() and ci.tolist() == oi.tolist() and cs.tobytes() == os_.tobytes(), mode
    gpu.close()


def _tie_fixture(gm, om, n_dup, n_close, ef):
    """Layer-0 graph in which more than HX_TIE_CAP (32) evicted-unexpanded entries tie with w.max:
    entry 0 and ids 1..n_dup-1 are identical vectors D (they fill the beam with one score), node 1 links to n_close
    distinct closer vectors whose admission evicts the duplicates one by one while w.max still has D's score."""
    dim = 8
    ora = hxo.Index(om, dim, m=16, m0=32, ef_construction=64)
    d = np.zeros(dim, np.float32)
    d[0], d[1] = 3.0, 4.0
    ids = list(range(n_dup + n_close))
    rows = np.zeros((len(ids), dim), np.float32)
    rows[:n_dup] = d
    for j in range(n_close):                       # closer to the query (1,0,...): distinct scores
        rows[n_dup + j] = [1.0 + 0.01 * (j + 1), 0.02 * (j + 1), 0, 0, 0, 0, 0, 0]
    ora.put_vectors(np.array(ids, np.uint64), rows)
    dup_ids = list(range(1, n_dup))
    close_ids = list(range(n_dup, n_dup + n_close))
    ora.put_neighbors(0, 0, dup_ids)
    ora.put_neighbors(0, 1, [0] + close_ids)
    for i in dup_ids[1:]:
        ora.put_neighbors(0, i, [0])
    for c in close_ids:
        ora.put_neighbors(0, c, [1])
    ora.set_entry(0, 0)
    gpu = hx.VectorIndex(gm, hx.VectorIndexConfig("ties", "embedding", dim))
    mirror_from_oracle(gpu, ora)
    q = np.zeros(dim, np.float32)
    q[0] = 1.0
    return gpu, ora, q


@pytest.mark.parametrize("gm,om", [(hx.Metric.Euclidean, hxo.EUCLIDEAN), (hx.Metric.Cosine, hxo.COSINE)])
def test_more_than_32_ties_at_the_beam_boundary(gm, om):
    """the reference's candidates heap is unbounded; the device tie stack used to fail the whole batch past 32."""
    ef, k = 80, 10
    gpu, ora, q = _tie_fixture(gm, om, n_dup=120, n_close=60, ef=ef)
    oi, os_, st = ora.search(q, k, ef=ef, with_stats=True)
    params = hx.SearchParams.strict(k, ef)
    params.collect_stats = True
    for B in (1, 40, 200):                         # CTA-per-query build (register beam), warp-per-query ring build
        qs = np.tile(q, (B, 1))
        gst = hx.SearchStats()
        gi, gs, gc = gpu.search_batch(qs, params, gst)
        for b in (0, B - 1):
            assert gc[b] == len(oi) and gi[b, :gc[b]].tolist() == oi.tolist() and gs[b, :gc[b]].tobytes() == os_.tobytes(), B
        for f in ("expansion_steps", "neighbors_examined", "distance_computations"):
            assert getattr(gst, f) == B * st[f], (f, B)
    with gpu.service(k, ef, capacity=64) as svc:
        r = svc.search(q)
        assert [x.entity_id() for x in r] == oi.tolist()
    gpu.close()


def test_per_query_status_in_a_batch():
    rng = np.random.default_rng(3)
    n, dim = 2000, 32
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    gpu, ora = build_pair(hx.Metric.Cosine, hxo.COSINE, rows, m=8, m0=16, efc=60)
    for B in (6, 40, 300):                         # host validation (B <= 8), CTA build, ring build (fused validation)
        queries = rng.standard_normal((B, dim)).astype(np.float32)
        queries[1, 5] = np.inf
        queries[B - 2] = 0.0
        p = hx.SearchParams.strict(5, 40)
        ids, sc, cnt, status = gpu.search_batch_status(queries, p)
        assert status[1] == hx.HX_ERR_INVALID_VECTOR_COMPONENT and status[B - 2] == hx.HX_ERR_ZERO_NORM_COSINE
        assert cnt[1] == 0 and cnt[B - 2] == 0
        good = [b for b in range(B) if b not in (1, B - 2)]
        assert all(status[b] == hx.HX_OK for b in good)
        for b in good[:20]:
            oi, os_ = ora.search(queries[b], 5, ef=40)
            assert ids[b, :cnt[b]].tolist() == oi.tolist() and sc[b, :cnt[b]].tobytes() == os_.tobytes()
        with pytest.raises(hx.HelixDbError):       # the all-or-nothing entry point still reports the first failure
            gpu.search_batch(queries, p)
    # validation precedes the empty-index answer (search.rs:1101-1128), for every batch size
    empty = hx.VectorIndex(hx.Metric.Cosine, hx.VectorIndexConfig("e", "embedding", dim))
    q = rng.standard_normal((20, dim)).astype(np.float32)
    assert empty.search_batch(q, hx.SearchParams.strict(3))[2].tolist() == [0] * 20
    q[17, 0] = np.nan
    with pytest.raises(hx.HelixDbError) as e:
        empty.search_batch(q, hx.SearchParams.strict(3))
    assert e.value.variant == "InvalidVectorComponent"
    with pytest.raises(hx.HelixDbError) as e:      # default mode on an empty index: the query error, not a config error
        empty.search_batch(q, hx.SearchParams.new(3))
    assert e.value.variant == "InvalidVectorComponent"
    assert empty.search_batch(q[:10], hx.SearchParams.new(3))[2].tolist() == [0] * 10
    with pytest.raises(hx.HelixDbError) as e:      # unsorted candidate ids are refused, like hx_candidates_create
        bad = np.array([5, 3, 9], dtype=np.uint64)
        gpu.search_restricted_batch(q[:1], hx.SearchParams.strict(2), hx.RestrictedVectorCandidates(bad))
    assert e.value.variant == "VectorParameterError"
    gpu.close()
    empty.close()


def test_device_calls_on_concurrent_streams():
    """device-buffer calls on different streams used to share one scratch set."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(11)
    n, dim, B, k = 4000, 64, 400, 10
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    gpu, ora = build_pair(hx.Metric.Euclidean, hxo.EUCLIDEAN, rows, m=8, m0=16, efc=60)
    dev = torch.device("cuda", 0)
    T = 4
    queries = rng.standard_normal((T, B, dim)).astype(np.float32)
    params = hx.SearchParams.strict(k, 50)
    expect = [ora.search_batch(queries[t], k, 50, threads=4) for t in range(T)]
    streams = [torch.cuda.Stream(dev) for _ in range(T)]
    dq = [torch.from_numpy(queries[t]).to(dev) for t in range(T)]
    out = [(torch.zeros((B, k), dtype=torch.int64, device=dev), torch.zeros((B, k), dtype=torch.float32, device=dev),
            torch.zeros((B,), dtype=torch.int32, device=dev)) for _ in range(T)]
    torch.cuda.synchronize(dev)
    errs = []

    def worker(t):
        try:
            for _ in range(20):
                gpu.search_device(dq[t].data_ptr(), B, params, out[t][0].data_ptr(), out[t][1].data_ptr(),
                                  out[t][2].data_ptr(), streams[t].cuda_stream)
            flags, status = gpu.device_flags(streams[t].cuda_stream)
            assert flags == 0 and status == 0
        except Exception as e:   # pragma: no cover
            errs.append(e)

    ths = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert not errs, errs
    for t in range(T):
        ids = out[t][0].cpu().numpy().view(np.uint64)
        assert ids.tolist() == expect[t][0].tolist() and out[t][1].cpu().numpy().tobytes() == expect[t][1].tobytes()
    gpu.close()
