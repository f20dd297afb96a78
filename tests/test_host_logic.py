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
