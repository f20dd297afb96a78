"""Host-side mirror of the reference's parameter / candidate types (no GPU needed)."""
import numpy as np
import pytest

import helix_db_b200 as hx


def test_search_params_defaults_and_builder():
    p = hx.SearchParams.new(10)                      # mod.rs:482-500: ef = max(k,100), Adaptive
    assert p.k() == 10 and p.ef() == 100 and p.requires_query_simhash()
    assert hx.SearchParams.new(250).ef() == 250
    with pytest.raises(hx.VectorParameterError):
        hx.SearchParams.new(0)
    with pytest.raises(hx.VectorParameterError):
        hx.SearchParams.new(10).with_ef(9)           # SearchBeamWidth must cover k
    s = hx.SearchParams.new(4).with_ef(16).with_simhash_mode(hx.SimHashMode.Off).with_pre_simhash_sampling_ratio(1.0)
    assert not s.requires_query_simhash()            # the strict-exhaustive state (mod.rs:556-561)
    assert hx.SearchParams.new(4).with_simhash_mode(hx.SimHashMode.Off).with_pre_simhash_sampling_ratio(0.5) \
        .requires_query_simhash()
    c = s._c()
    assert (c.k, c.ef, c.simhash_mode, c.pre_sampling_ratio) == (4, 16, 0, 1.0)


def test_restricted_candidates_canonicalise_and_bound():
    c = hx.RestrictedVectorCandidates.from_ids([5, 3, 5, 9, 3])
    assert c.ids.tolist() == [3, 5, 9] and len(c) == 3 and c.contains(5) and not c.contains(4)
    assert hx.RestrictedVectorCandidates.from_ids([]).is_empty()
    with pytest.raises(hx.HelixDbError) as e:
        hx.RestrictedVectorCandidates.from_ids(np.arange(1_000_001, dtype=np.uint64))
    assert e.value.variant == "Query"
    assert len(hx.RestrictedVectorCandidates.from_ids(np.arange(1_000_000, dtype=np.uint64))) == 1_000_000


def test_search_result_equality_is_bitwise():
    a, b = hx.SearchResult(1, 0.5), hx.SearchResult(1, 0.5)
    assert a == b and a.entity_id() == 1 and float(a.score()) == 0.5
    assert hx.SearchResult(1, 0.0) != hx.SearchResult(1, -0.0)


def test_policy_builder_methods_validate_like_the_reference():
    p = hx.SearchParams.new(10)
    pol = p._policy()                                   # SearchParams::new defaults (mod.rs:482-500)
    assert (pol.bypass_min_frontier, pol.bypass_window_expansions, pol.read_budget_multiplier) == (24, 4, 3)
    assert pol.sampling_ratio_override < 0 and pol.failure_prob_override < 0 and p._c().pre_sampling_ratio < 0
    q = p.with_simhash_bypass_tuning(8, 2, 0.5, 5).with_simhash_sampling_ratio(0.25).with_simhash_failure_prob(0.2)
    pol = q._policy()
    assert (pol.bypass_min_frontier, pol.bypass_window_expansions, pol.read_budget_multiplier) == (8, 2, 5)
    assert abs(pol.sampling_ratio_override - 0.25) < 1e-7 and abs(pol.failure_prob_override - 0.2) < 1e-7
    for bad in ((0, 4, 0.1, 3), (24, 0, 0.1, 3), (24, 4, 1.5, 3), (24, 4, 0.1, 0)):   # mod.rs:563-592
        with pytest.raises(hx.VectorParameterError):
            hx.SearchParams.new(10).with_simhash_bypass_tuning(*bad)
    with pytest.raises(hx.VectorParameterError):
        hx.SearchParams.new(10).with_simhash_sampling_ratio(1.5)
    for bad in (0.0, 1.0, -0.1):                          # FailureProbability is the open unit interval
        with pytest.raises(hx.VectorParameterError):
            hx.SearchParams.new(10).with_simhash_failure_prob(bad)
    with pytest.raises(hx.VectorParameterError):
        hx.SearchParams.new(10).with_pre_simhash_sampling_ratio(-0.5)
    t = hx.SearchParams.throughput_profile_floor_92(100)  # mod.rs:615-621: ef = max(k, 48)
    assert t.ef() == 100 and t.requires_query_simhash()


def test_bench_config_is_shared_verbatim_by_both_arms(monkeypatch):
    """bench.py: the driver compares the `config` objects of our arm and of `--impl reference`; nothing that varies from run
    to run (setup timings) may live inside it, and the reference arm quotes OUR arm's config at the same N."""
    import sys as _sys
    import bench
    for n_gpus in (1, 8):
        monkeypatch.setattr(_sys, "argv", ["bench.py", "--gpus", str(n_gpus)])
        ours = bench.c2_config(bench.parse(), n_gpus)
        monkeypatch.setattr(_sys, "argv", ["bench.py", "--gpus", str(n_gpus), "--impl", "reference", "--steps", "3"])
        ref_args = bench.parse()
        theirs = bench.c2_config(ref_args, ref_args.gpus)
        assert ours == theirs and "setup" not in ours
        assert ours["workload"].startswith("C2: 1000000x768 f32 cosine HNSW top-10")
        assert ours["queries_per_step_per_gpu"] == 32768
