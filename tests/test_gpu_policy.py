"""The production-default search mode on the device: SimHash filtering, sampling and adaptive bypass
(SimHashMode::{Adaptive, Always}, or Off with a pre-sampling override) — k_hnsw_search_policy through hx_search_ex /
hx_search, against the CPU oracle's restatement of search.rs:595-992 + policy.rs on identical rows, graph, fingerprints
and parameters.  Bar: ids, order, score bits, the SearchStats counters and the SimHash counters are all equal.

What the reference pins here is stated in oracle/hx_oracle.h: the policy functions are pinned by the reference's own policy
tests (tests/test_oracle_kat.py); the session RNG's stream values are not pinned by the reference, so device and oracle are
compared with each other (same restated generator), not with a golden value.
"""
import numpy as np
import pytest

import helix_db_b200 as hx
from oracle import hxo

from test_gpu_parity import build_pair

pytestmark = pytest.mark.gpu


def oracle_cfg(params: hx.SearchParams, threshold=43, sampling_ratio=0.8, adaptive_enabled=True, failure=0.1):
    pol = params._policy()
    return hxo.policy_defaults(
        mode=int(params._mode), threshold=threshold,
        sampling_ratio=pol.sampling_ratio_override if pol.sampling_ratio_override >= 0 else sampling_ratio,
        has_pre_override=0 if params._pre_ratio is None else 1,
        pre_override=1.0 if params._pre_ratio is None else params._pre_ratio,
        adaptive_enabled=1 if adaptive_enabled else 0,
        failure_prob=pol.failure_prob_override if pol.failure_prob_override >= 0 else failure,
        bypass_min_frontier=pol.bypass_min_frontier, bypass_window_expansions=pol.bypass_window_expansions,
        bypass_min_filter_rate=pol.bypass_min_filter_rate, read_budget_multiplier=pol.read_budget_multiplier)


def clustered(rng, n, dim, ncent=20, spread=0.4):
    cent = rng.standard_normal((ncent, dim)).astype(np.float32)
    return (cent[rng.integers(0, ncent, n)] + spread * rng.standard_normal((n, dim))).astype(np.float32), cent


def compare(gpu, ora, queries, qsim, params, cfg, what):
    k, ef = params.k(), params.ef()
    st, ps = hx.SearchStats(), hx.PolicyStats()
    params.collect_stats = True
    gi, gs, gc = gpu.search_ex(queries, params, query_simhash=qsim, stats=st, policy_stats=ps)
    tot, ptot = {}, {}
    for q in range(len(queries)):
        oi, os_, ost, ops = ora.search_policy(queries[q], k, ef, cfg, int(qsim[q]))
        assert int(gc[q]) == len(oi), (what, q)
        assert gi[q, :gc[q]].tolist() == oi.tolist(), (what, q)
        assert gs[q, :gc[q]].tobytes() == os_.tobytes(), (what, q)
        for key, val in ost.items():
            tot[key] = tot.get(key, 0) + val
        for key, val in ops.items():
            ptot[key] = ptot.get(key, 0) + val
    for f in ("expansion_steps", "neighbors_examined", "distance_computations", "vectors_loaded"):
        assert getattr(st, f) == tot[f], (what, f)
    assert ps.as_dict() == ptot, what
    # small batches (B < #SMs) run the CTA-per-query build of the same kernel: same bits, same counters
    st2, ps2 = hx.SearchStats(), hx.PolicyStats()
    si, ss, scnt = gpu.search_ex(queries[:40], params, query_simhash=qsim[:40], stats=st2, policy_stats=ps2)
    assert si.tolist() == gi[:40].tolist() and ss.tobytes() == gs[:40].tobytes() and scnt.tolist() == gc[:40].tolist(), what
    st3, ps3 = hx.SearchStats(), hx.PolicyStats()
    gpu.tune(pol_cta=0)                                   # the warp-per-query build for the same 40 queries
    try:
        gpu.search_ex(queries[:40], params, query_simhash=qsim[:40], stats=st3, policy_stats=ps3)
    finally:
        gpu.tune()
    assert ps2.as_dict() == ps3.as_dict() and st2.expansion_steps == st3.expansion_steps, what
    return ptot


@pytest.mark.parametrize("gm,om", [(hx.Metric.Cosine, hxo.COSINE), (hx.Metric.Euclidean, hxo.EUCLIDEAN),
                                   (hx.Metric.Manhattan, hxo.MANHATTAN)])
def test_policy_modes_match_the_oracle(gm, om):
    rng = np.random.default_rng(17)
    n, dim, nq = 2500, 48, 192
    rows, cent = clustered(rng, n, dim)
    queries = (cent[rng.integers(0, len(cent), nq)] + 0.4 * rng.standard_normal((nq, dim))).astype(np.float32)
    planes = rng.standard_normal((64, dim)).astype(np.float32)
    gpu, ora = build_pair(gm, om, rows, m=8, m0=16, efc=60)
    # fingerprints: projected on the device with the reference's sequential order == the oracle's projection, bit for bit
    gpu.set_simhash_planes(planes)
    gpu.compute_simhash()
    bits = np.array([hxo.simhash_from_planes(planes, rows[i]) for i in range(n)], dtype=np.uint64)
    assert gpu.download_simhash(0, n).tolist() == bits.tolist()
    ora.put_simhash(np.arange(n, dtype=np.uint64), bits)
    qsim = np.array([hxo.simhash_from_planes(planes, q) for q in queries], dtype=np.uint64)

    cases = [
        ("SearchParams::new", hx.SearchParams.new(10), {}),
        ("new, ef 16 (sampling gates open)", hx.SearchParams.new(10).with_ef(16), {}),
        ("Always t=30 r=0.5", hx.SearchParams.new(10).with_ef(32).with_simhash_mode(hx.SimHashMode.Always),
         dict(threshold=30, sampling_ratio=0.5)),
        ("Always strict threshold", hx.SearchParams.new(5).with_ef(40).with_simhash_mode(hx.SimHashMode.Always),
         dict(threshold=40, sampling_ratio=1.0)),
        ("defer all (ratio 0) -> fallback picks", hx.SearchParams.new(10).with_ef(16).with_simhash_sampling_ratio(0.0), {}),
        ("Off + pre-sampling 0.3", hx.SearchParams.new(10).with_ef(16).with_simhash_mode(hx.SimHashMode.Off)
         .with_pre_simhash_sampling_ratio(0.3), {}),
        ("bypass windows (low yield)", hx.SearchParams.new(10).with_ef(24).with_simhash_bypass_tuning(2, 2, 0.9, 3), {}),
        ("adaptive flag off", hx.SearchParams.new(10).with_ef(20), dict(adaptive_enabled=False, threshold=28)),
        ("failure prob override", hx.SearchParams.new(10).with_ef(64).with_simhash_failure_prob(0.45), {}),
        ("throughput_profile_floor_92", hx.SearchParams.throughput_profile_floor_92(10), {}),
    ]
    seen_rng = seen_filter = seen_bypass = 0
    for what, params, idx in cases:
        gpu.set_simhash_config(threshold=idx.get("threshold", 43), sampling_ratio=idx.get("sampling_ratio", 0.8),
                               adaptive_enabled=idx.get("adaptive_enabled", True))
        cfg = oracle_cfg(params, threshold=idx.get("threshold", 43), sampling_ratio=idx.get("sampling_ratio", 0.8),
                         adaptive_enabled=idx.get("adaptive_enabled", True))
        ptot = compare(gpu, ora, queries, qsim, params, cfg, f"{gm.name}: {what}")
        seen_rng += ptot["rng_draws"]
        seen_filter += ptot["simhash_filtered"]
        seen_bypass += ptot["simhash_bypass_expansions"]
    assert seen_rng > 0                                   # the sampling paths really ran
    if om == hxo.COSINE:
        assert seen_filter > 0 and seen_bypass > 0        # and so did the threshold gate and the bypass windows
    else:
        assert seen_filter == 0                           # only cosine has an angular SimHash contract (policy.rs:67-88)

    # plain hx_search with SearchParams::new: fingerprints of the queries projected on the device from the planes
    gpu.set_simhash_config()
    params = hx.SearchParams.new(10)
    gi, gs, gc = gpu.search_batch(queries[:64], params)
    cfg = oracle_cfg(params)
    for q in range(64):
        oi, os_, _, _ = ora.search_policy(queries[q], 10, 100, cfg, int(qsim[q]))
        assert gi[q, :gc[q]].tolist() == oi.tolist() and gs[q, :gc[q]].tobytes() == os_.tobytes()
    # recall of the production default against the exact answer (the reference's gate for this mode)
    hit = 0
    for q in range(64):
        ex, _ = ora.search_exact(queries[q], 10)
        hit += len(set(gi[q, :gc[q]].tolist()) & set(ex.tolist()))
    assert hit / 640.0 >= 0.92


def test_policy_mode_requirements_and_strict_forwarding():
    rng = np.random.default_rng(3)
    rows, _ = clustered(rng, 400, 16)
    gpu, ora = build_pair(hx.Metric.Cosine, hxo.COSINE, rows, m=6, m0=12, efc=40)
    q = rows[:3] + 0.01
    with pytest.raises(hx.HelixDbError) as e:             # filtering needs the [0x12] rows
        gpu.search_batch(q, hx.SearchParams.new(5))
    assert e.value.variant == "InvalidVectorConfig"
    bits = rng.integers(0, 2**63, size=400, dtype=np.uint64)
    gpu.load_simhash(np.arange(400, dtype=np.uint64), bits)
    with pytest.raises(hx.HelixDbError) as e:             # no query fingerprint and no planes to project it
        gpu.search_batch(q, hx.SearchParams.new(5))
    assert e.value.variant == "InvalidVectorConfig"
    gi, gs, gc = gpu.search_ex(q, hx.SearchParams.new(5), query_simhash=np.array([1, 2, 3], dtype=np.uint64))
    assert gc.tolist() == [5, 5, 5]
    # the exhaustive specialisation through hx_search_ex == hx_search
    a = gpu.search_ex(q, hx.SearchParams.strict(5))
    b = gpu.search_batch(q, hx.SearchParams.strict(5))
    assert a[0].tolist() == b[0].tolist() and a[1].tobytes() == b[1].tobytes()
    with pytest.raises(hx.HelixDbError):
        gpu.load_simhash([10**9], [1])                    # a fingerprint for a node without a vector row
    assert hx.load_library().hx_order_code_from_simhash_bits(1 << 47) == 1 << 62
