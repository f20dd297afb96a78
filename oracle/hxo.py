"""ctypes binding of the CPU oracle (oracle/hx_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(helix-db_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libhx_oracle.so"
if os.environ.get("HXO_ORACLE_SO"):                # another build of the same checker (oracle/Makefile: `make asan`)
    _SO = Path(os.environ["HXO_ORACLE_SO"]).resolve()

EUCLIDEAN, COSINE, MANHATTAN = 0, 1, 2
METRICS = {"euclidean": EUCLIDEAN, "cosine": COSINE, "manhattan": MANHATTAN}

OK = 0
ERR_INDEX_NOT_FOUND = 1
ERR_INVALID_DIMENSION = 2
ERR_INVALID_VECTOR_COMPONENT = 3
ERR_ZERO_NORM_COSINE = 4
ERR_MAGNITUDE_EXCEEDED = 5
ERR_INVALID_VECTOR_CONFIG = 6
ERR_QUERY = 7
ERR_INVARIANT_VIOLATION = 8
ERR_INVALID_PARAMETER = 9


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (gcc only)."""
    src = _HERE / "hx_oracle.c"
    hdr = _HERE / "hx_oracle.h"
    if force or not _SO.exists() or _SO.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    return _SO


class FilteredBudgets(C.Structure):
    """FilteredGraphBudgets (restricted.rs:220-260)."""
    _fields_ = [(n, C.c_size_t) for n in ("ef_filtered", "routing_rows", "bridge_rows", "vector_payloads", "sampled_seeds")]


class FilteredStats(C.Structure):
    """The RestrictedSearchStats counters the filtered walk reports (restricted.rs)."""
    _fields_ = [(n, C.c_uint64) for n in ("vector_payload_requests", "distance_computations", "routing_rows", "bridge_rows",
                                          "bridge_frontier_pushes", "simhash_row_requests")] + \
               [("termination", C.c_uint32), ("reserved", C.c_uint32)]

    TERMINATION = {0: None, 1: "VectorBudget", 2: "BeamComplete", 3: "Exhausted", 4: "RoutingBudget", 5: "BridgeBudget"}

    def as_dict(self):
        d = {f: int(getattr(self, f)) for f, _ in self._fields_ if f != "reserved"}
        d["termination"] = self.TERMINATION[d["termination"]]
        return d


class Stats(C.Structure):
    _fields_ = [
        ("expansion_steps", C.c_uint64),
        ("neighbors_examined", C.c_uint64),
        ("distance_computations", C.c_uint64),
        ("vectors_loaded", C.c_uint64),
        ("upper_layer_steps", C.c_uint64),
    ]

    def as_dict(self):
        return {f: int(getattr(self, f)) for f, _ in self._fields_}



class PolicyCfg(C.Structure):
    """hxo_policy_cfg: SearchParams (mod.rs:411-500) + index config (config/indexes.rs:398-406)."""
    _fields_ = [("mode", C.c_int32), ("threshold", C.c_uint32), ("sampling_ratio", C.c_float),
                ("has_pre_override", C.c_int32), ("pre_override", C.c_float), ("adaptive_enabled", C.c_int32),
                ("failure_prob", C.c_float), ("bypass_min_frontier", C.c_uint32), ("bypass_window_expansions", C.c_uint32),
                ("bypass_min_filter_rate", C.c_float), ("read_budget_multiplier", C.c_uint32)]


class PolicyCtx(C.Structure):
    _fields_ = [("topk_ready", C.c_int32), ("ef", C.c_uint32), ("search_frontier_len", C.c_uint32),
                ("candidate_frontier_len", C.c_uint32), ("current", C.c_float), ("delta", C.c_float),
                ("bypass_state", C.c_int32), ("bypass_remaining", C.c_uint32), ("simhash_filter_reads", C.c_uint64),
                ("window_examined", C.c_uint64), ("window_filtered", C.c_uint64), ("window_expansions", C.c_uint64)]


class PolicyDecision(C.Structure):
    _fields_ = [("fetch_missing", C.c_int32), ("filter_cached", C.c_int32), ("has_threshold", C.c_int32),
                ("threshold", C.c_uint32), ("pre_kind", C.c_int32), ("pre_prob", C.c_float), ("samp_kind", C.c_int32),
                ("samp_prob", C.c_float), ("base_sampling_probability", C.c_float), ("bypassed", C.c_int32),
                ("next_state", C.c_int32), ("next_remaining", C.c_uint32), ("trigger", C.c_int32)]

    def sampling_probability(self):
        return 1.0 if self.samp_kind == SAMPLING_EXHAUSTIVE else float(self.samp_prob)

    def pre_probability(self):
        return 1.0 if self.pre_kind == SAMPLING_EXHAUSTIVE else float(self.pre_prob)


class Session(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("started", C.c_int32), ("key", C.c_uint32 * 8), ("block", C.c_uint64),
                ("buf", C.c_uint32 * 16), ("pos", C.c_uint32)]


class PolicyStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "simhash_filtered", "simhash_examined", "simhash_missing_hash", "simhash_passed_before_sampling",
        "simhash_passed_after_sampling", "simhash_bypass_expansions", "simhash_skipped_candidates",
        "pre_simhash_sample_kept", "pre_simhash_sample_dropped", "simhash_bypass_trigger_budget",
        "simhash_bypass_trigger_low_yield", "rng_draws")]

    def as_dict(self):
        return {f: int(getattr(self, f)) for f, _ in self._fields_}


SIMHASH_OFF, SIMHASH_ADAPTIVE, SIMHASH_ALWAYS = 0, 1, 2
BYPASS_READY, BYPASS_BYPASSING, BYPASS_COOLING = 0, 1, 2
TRIGGER_NONE, TRIGGER_READ_BUDGET, TRIGGER_LOW_YIELD, TRIGGER_BOTH = 0, 1, 2, 3
SAMPLING_EXHAUSTIVE, SAMPLING_FIXED, SAMPLING_ADAPTIVE = 0, 1, 2


def policy_defaults(**overrides) -> "PolicyCfg":
    cfg = PolicyCfg()
    lib().hxo_policy_defaults(C.byref(cfg))
    for key, val in overrides.items():
        setattr(cfg, key, val)
    return cfg


def policy_decide(metric, cfg, **ctx_fields) -> "PolicyDecision":
    ctx = PolicyCtx(**ctx_fields)
    out = PolicyDecision()
    lib().hxo_policy_decide(metric, C.byref(cfg), C.byref(ctx), C.byref(out))
    return out


def candidate_probability(decision, similarity_bits) -> float:
    return float(lib().hxo_candidate_probability(C.byref(decision), similarity_bits))


def simhash_from_planes(planes, v) -> int:
    pa, pp = _f32(planes)
    va, vp = _f32(v)
    return int(lib().hxo_simhash_from_planes(pp, vp, va.size))

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(str(_SO))
    fp = C.POINTER(C.c_float)
    u64p = C.POINTER(C.c_uint64)
    u32p = C.POINTER(C.c_uint32)
    u16p = C.POINTER(C.c_uint16)
    sz = C.c_size_t
    for name in ("hxo_euclid_scalar", "hxo_dot_scalar", "hxo_manhattan", "hxo_euclid_avx_fma", "hxo_dot_avx_fma",
                 "hxo_euclid_avx_fma_portable", "hxo_dot_avx_fma_portable", "hxo_euclidean_distance",
                 "hxo_dot_product"):
        f = getattr(L, name)
        f.restype = C.c_float
        f.argtypes = [fp, fp, sz]
    L.hxo_has_avx_fma.restype = C.c_int
    L.hxo_scaled_l2_norm.restype = C.c_double
    L.hxo_scaled_l2_norm.argtypes = [fp, sz]
    L.hxo_cosine_norm.restype = C.c_float
    L.hxo_cosine_norm.argtypes = [fp, sz]
    L.hxo_header.restype = C.c_float
    L.hxo_header.argtypes = [C.c_int, fp, sz]
    L.hxo_cosine_distance.restype = C.c_float
    L.hxo_cosine_distance.argtypes = [fp, C.c_float, fp, C.c_float, sz]
    L.hxo_distance.restype = C.c_float
    L.hxo_distance.argtypes = [C.c_int, fp, C.c_float, fp, C.c_float, sz]
    L.hxo_score_validate.restype = C.c_int
    L.hxo_score_validate.argtypes = [fp]
    L.hxo_component_limit.restype = C.c_int
    L.hxo_component_limit.argtypes = [C.c_int, sz, fp]
    L.hxo_validate_vector.restype = C.c_int
    L.hxo_validate_vector.argtypes = [C.c_int, sz, fp, sz, u32p]
    L.hxo_select_layer_from_uniform.restype = C.c_uint16
    L.hxo_select_layer_from_uniform.argtypes = [C.c_float, C.c_float]
    L.hxo_default_ml_for_m.restype = C.c_float
    L.hxo_default_ml_for_m.argtypes = [C.c_uint32]
    L.hxo_restricted_plan.restype = C.c_int
    L.hxo_restricted_plan.argtypes = [C.c_uint64, C.c_uint32]
    L.hxo_restricted_result_count.restype = C.c_int
    L.hxo_restricted_result_count.argtypes = [C.c_uint32, C.c_uint64, u32p]
    L.hxo_deterministic_sample_ids.restype = sz
    L.hxo_deterministic_sample_ids.argtypes = [u64p, sz, sz, u64p]
    L.hxo_fixture_xorshift_vector.restype = None
    L.hxo_fixture_xorshift_vector.argtypes = [C.c_uint64, C.c_uint32, fp]
    L.hxo_fixture_circle_vector.restype = None
    L.hxo_fixture_circle_vector.argtypes = [C.c_uint64, C.c_uint64, fp]
    L.hxo_fixture_skip_neighbors.restype = sz
    L.hxo_fixture_skip_neighbors.argtypes = [C.c_uint64, C.c_uint64, u64p, sz]
    L.hxo_index_new.restype = C.c_void_p
    L.hxo_index_new.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.hxo_index_free.restype = None
    L.hxo_index_free.argtypes = [C.c_void_p]
    L.hxo_index_len.restype = sz
    L.hxo_index_len.argtypes = [C.c_void_p]
    L.hxo_index_state.restype = C.c_int
    L.hxo_index_state.argtypes = [C.c_void_p, u64p, u16p]
    L.hxo_index_layer0_limit.restype = C.c_uint32
    L.hxo_index_layer0_limit.argtypes = [C.c_void_p]
    L.hxo_index_put_vector.restype = C.c_int
    L.hxo_index_put_vector.argtypes = [C.c_void_p, C.c_uint64, fp]
    L.hxo_index_put_vectors.restype = C.c_int
    L.hxo_index_put_vectors.argtypes = [C.c_void_p, u64p, fp, sz]
    L.hxo_index_put_neighbors.restype = C.c_int
    L.hxo_index_put_neighbors.argtypes = [C.c_void_p, C.c_uint16, C.c_uint64, u64p, sz]
    L.hxo_index_set_entry.restype = C.c_int
    L.hxo_index_set_entry.argtypes = [C.c_void_p, C.c_uint64, C.c_uint16]
    L.hxo_index_insert.restype = C.c_int
    L.hxo_index_insert.argtypes = [C.c_void_p, C.c_uint64, fp, C.c_uint16]
    L.hxo_index_upsert.restype = C.c_int
    L.hxo_index_upsert.argtypes = [C.c_void_p, C.c_uint64, fp, C.c_uint16]
    L.hxo_index_delete.restype = C.c_int
    L.hxo_index_delete.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]
    L.hxo_index_node_ids.restype = sz
    L.hxo_index_node_ids.argtypes = [C.c_void_p, u64p, sz]
    L.hxo_index_node_level.restype = C.c_int
    L.hxo_index_node_level.argtypes = [C.c_void_p, C.c_uint64]
    L.hxo_index_get_neighbors.restype = sz
    L.hxo_index_get_neighbors.argtypes = [C.c_void_p, C.c_uint16, C.c_uint64, u64p, sz]
    L.hxo_index_get_vector.restype = C.c_int
    L.hxo_index_get_vector.argtypes = [C.c_void_p, C.c_uint64, fp]
    L.hxo_search_layer_greedy.restype = C.c_int
    L.hxo_search_layer_greedy.argtypes = [C.c_void_p, fp, C.c_uint64, C.c_uint16, u64p]
    L.hxo_search.restype = C.c_int
    L.hxo_search.argtypes = [C.c_void_p, fp, C.c_uint32, C.c_uint32, C.c_uint32, u64p, fp, u32p, C.POINTER(Stats)]
    L.hxo_filtered_budgets.restype = None
    L.hxo_filtered_budgets.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, sz, C.POINTER(FilteredBudgets)]
    L.hxo_search_filtered_graph_budgets.restype = C.c_int
    L.hxo_search_filtered_graph_budgets.argtypes = [C.c_void_p, fp, C.c_uint32, C.POINTER(FilteredBudgets), u64p, sz, C.c_uint64,
                                                    u64p, fp, C.POINTER(C.c_uint32), C.POINTER(FilteredStats)]
    L.hxo_search_restricted.restype = C.c_int
    L.hxo_search_restricted.argtypes = [C.c_void_p, fp, C.c_uint32, C.c_uint32, u64p, sz, u64p, fp, u32p, u64p]
    L.hxo_search_exact.restype = C.c_int
    L.hxo_search_exact.argtypes = [C.c_void_p, fp, C.c_uint32, u64p, fp, u32p]
    L.hxo_search_batch.restype = C.c_double
    L.hxo_search_batch.argtypes = [C.c_void_p, fp, sz, C.c_uint32, C.c_uint32, C.c_int, u64p, fp, u32p,
                                   C.POINTER(Stats)]
    L.hxo_search_restricted_batch.restype = C.c_double
    L.hxo_search_restricted_batch.argtypes = [C.c_void_p, fp, sz, C.c_uint32, u64p, u64p, C.c_int, u64p, fp, u32p]
    L.hxo_search_exact_batch.restype = C.c_double
    L.hxo_search_exact_batch.argtypes = [C.c_void_p, fp, sz, C.c_uint32, C.c_int, u64p, fp, u32p]
    L.hxo_index_import_graph.restype = C.c_int
    L.hxo_index_import_graph.argtypes = [C.c_void_p, u16p, u32p, u32p, C.c_uint32, sz, u32p, u16p, u32p, u32p,
                                         C.c_uint32, C.c_uint64, C.c_uint16]
    L.hxo_policy_defaults.restype = None
    L.hxo_policy_defaults.argtypes = [C.POINTER(PolicyCfg)]
    L.hxo_policy_decide.restype = None
    L.hxo_policy_decide.argtypes = [C.c_int, C.POINTER(PolicyCfg), C.POINTER(PolicyCtx), C.POINTER(PolicyDecision)]
    L.hxo_candidate_probability.restype = C.c_float
    L.hxo_candidate_probability.argtypes = [C.POINTER(PolicyDecision), C.c_uint32]
    L.hxo_session_seeded.restype = None
    L.hxo_session_seeded.argtypes = [C.POINTER(Session), C.c_uint64]
    L.hxo_session_seed_for.restype = C.c_uint64
    L.hxo_session_seed_for.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
    L.hxo_session_next_u32.restype = C.c_uint32
    L.hxo_session_next_u32.argtypes = [C.POINTER(Session)]
    L.hxo_chacha_block.restype = None
    L.hxo_chacha_block.argtypes = [u32p, C.c_uint64, C.c_uint64, C.c_int, u32p]
    L.hxo_session_should_sample.restype = C.c_int
    L.hxo_session_should_sample.argtypes = [C.POINTER(Session), C.c_float]
    L.hxo_session_choose_index.restype = C.c_int64
    L.hxo_session_choose_index.argtypes = [C.POINTER(Session), C.c_uint64]
    L.hxo_simhash_from_planes.restype = C.c_uint64
    L.hxo_simhash_from_planes.argtypes = [fp, fp, C.c_uint32]
    L.hxo_simhash_collision_count.restype = C.c_uint32
    L.hxo_simhash_collision_count.argtypes = [C.c_uint64, C.c_uint64]
    L.hxo_order_code_from_simhash_bits.restype = C.c_uint64
    L.hxo_order_code_from_simhash_bits.argtypes = [C.c_uint64]
    L.hxo_index_put_simhash.restype = C.c_int
    L.hxo_index_put_simhash.argtypes = [C.c_void_p, u64p, u64p, sz]
    L.hxo_search_policy.restype = C.c_int
    L.hxo_search_policy.argtypes = [C.c_void_p, fp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(PolicyCfg), C.c_uint64,
                                    u64p, fp, u32p, C.POINTER(Stats), C.POINTER(PolicyStats)]
    L.hxo_search_policy_batch.restype = C.c_double
    L.hxo_search_policy_batch.argtypes = [C.c_void_p, fp, u64p, sz, C.c_uint32, C.c_uint32, C.POINTER(PolicyCfg), C.c_int,
                                          u64p, fp, u32p]
    _lib = L
    return L


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _u64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint64))


def _u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint32))


def _u16(a):
    a = np.ascontiguousarray(a, dtype=np.uint16)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint16))


class OracleError(Exception):
    def __init__(self, code, index=0):
        super().__init__(f"oracle status {code} (index {index})")
        self.code = code
        self.index = index


# ---- free functions ---------------------------------------------------------
def pair(name, u, v):
    ua, up = _f32(u)
    va, vp = _f32(v)
    assert ua.shape == va.shape
    return float(getattr(lib(), name)(up, vp, ua.size))


def distance(metric, p, q):
    pa, pp = _f32(p)
    qa, qp = _f32(q)
    L = lib()
    ph = L.hxo_header(metric, pp, pa.size)
    qh = L.hxo_header(metric, qp, qa.size)
    return float(L.hxo_distance(metric, pp, ph, qp, qh, pa.size))


def header(metric, v):
    va, vp = _f32(v)
    return float(lib().hxo_header(metric, vp, va.size))


def component_limit(metric, d):
    out = C.c_float(0)
    has = lib().hxo_component_limit(metric, d, C.byref(out))
    return float(out.value) if has else None


def validate_vector(metric, expected_d, v):
    va, vp = _f32(v)
    bad = C.c_uint32(0)
    rc = lib().hxo_validate_vector(metric, expected_d, vp, va.size, C.byref(bad))
    return rc, int(bad.value)


def xorshift_vectors(first_id, n, d=128):
    out = np.empty((n, d), dtype=np.float32)
    L = lib()
    for i in range(n):
        L.hxo_fixture_xorshift_vector(first_id + i, d, out[i].ctypes.data_as(C.POINTER(C.c_float)))
    return out


def circle_vector(entity_id, entity_count):
    out = np.empty(2, dtype=np.float32)
    lib().hxo_fixture_circle_vector(entity_id, entity_count, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def skip_neighbors(entity_id, entity_count):
    out = np.empty(160, dtype=np.uint64)
    n = lib().hxo_fixture_skip_neighbors(entity_id, entity_count, out.ctypes.data_as(C.POINTER(C.c_uint64)), 160)
    return out[:n].copy()


def deterministic_sample_ids(sorted_ids, limit):
    a, p = _u64(sorted_ids)
    out = np.empty(max(1, min(limit, a.size)), dtype=np.uint64)
    n = lib().hxo_deterministic_sample_ids(p, a.size, limit, out.ctypes.data_as(C.POINTER(C.c_uint64)))
    return out[:n].copy()


def restricted_result_count(k, n):
    out = C.c_uint32(0)
    rc = lib().hxo_restricted_result_count(k, n, C.byref(out))
    return rc, int(out.value)


# ---- index ------------------------------------------------------------------
class Index:
    """Flat in-memory HNSW index traversed exactly like the reference traverses its KV rows."""

    def __init__(self, metric, dim, m=16, m0=32, ef_construction=200):
        self.L = lib()
        self.metric, self.dim, self.m, self.m0, self.ef_construction = metric, dim, m, m0, ef_construction
        self.h = self.L.hxo_index_new(metric, dim, m, m0, ef_construction)
        if not self.h:
            raise OracleError(ERR_INVALID_VECTOR_CONFIG)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.hxo_index_free(self.h)
            self.h = None

    def __len__(self):
        return int(self.L.hxo_index_len(self.h))

    def _ck(self, rc, idx=0):
        if rc != OK:
            raise OracleError(rc, idx)

    def state(self):
        e, ml = C.c_uint64(0), C.c_uint16(0)
        if not self.L.hxo_index_state(self.h, C.byref(e), C.byref(ml)):
            return None
        return int(e.value), int(ml.value)

    @property
    def layer0_limit(self):
        return int(self.L.hxo_index_layer0_limit(self.h))

    def put_vector(self, node_id, v):
        va, vp = _f32(v)
        assert va.size == self.dim
        self._ck(self.L.hxo_index_put_vector(self.h, node_id, vp))

    def put_vectors(self, ids, rows):
        ia, ip = _u64(ids)
        ra, rp = _f32(rows)
        assert ra.size == ia.size * self.dim
        self._ck(self.L.hxo_index_put_vectors(self.h, ip, rp, ia.size))

    def put_neighbors(self, layer, node_id, nbrs):
        na, np_ = _u64(nbrs)
        self._ck(self.L.hxo_index_put_neighbors(self.h, layer, node_id, np_, na.size))

    def set_entry(self, entry, max_layer):
        self._ck(self.L.hxo_index_set_entry(self.h, entry, max_layer))

    def insert(self, node_id, v, layer):
        va, vp = _f32(v)
        assert va.size == self.dim
        self._ck(self.L.hxo_index_insert(self.h, node_id, vp, layer))

    def upsert(self, node_id, v, layer):
        """insert with VectorInsertContract::Upsert (mutation.rs:653-661): an existing item is deleted first."""
        va, vp = _f32(v)
        assert va.size == self.dim
        self._ck(self.L.hxo_index_upsert(self.h, node_id, vp, layer))

    def delete(self, node_id) -> bool:
        """VectorIndex::delete (mutation.rs:1606-1773); True when the item existed."""
        existed = C.c_int(0)
        self._ck(self.L.hxo_index_delete(self.h, node_id, C.byref(existed)))
        return bool(existed.value)

    def node_ids(self):
        n = len(self)
        out = np.empty(max(n, 1), dtype=np.uint64)
        w = self.L.hxo_index_node_ids(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64)), n)
        return out[:w].copy()

    def node_level(self, node_id):
        return int(self.L.hxo_index_node_level(self.h, node_id))

    def neighbors(self, layer, node_id, cap=4096):
        out = np.empty(cap, dtype=np.uint64)
        n = self.L.hxo_index_get_neighbors(self.h, layer, node_id, out.ctypes.data_as(C.POINTER(C.c_uint64)), cap)
        return out[:n].copy()

    def vector(self, node_id):
        out = np.empty(self.dim, dtype=np.float32)
        self._ck(self.L.hxo_index_get_vector(self.h, node_id, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def export_graph(self):
        """Returns {layer: (node_ids, offsets, neighbors)} CSR per layer plus (entry, max_layer)."""
        ids = self.node_ids()
        st = self.state()
        max_layer = st[1] if st else 0
        levels = np.array([self.node_level(int(i)) for i in ids], dtype=np.int32)
        top = int(max(max_layer, levels.max() if len(levels) else 0))
        out = {}
        for layer in range(0, top + 1):
            sel = ids[levels >= layer]
            offs = [0]
            nb = []
            for i in sel:
                r = self.neighbors(layer, int(i))
                nb.append(r)
                offs.append(offs[-1] + len(r))
            out[layer] = (sel.astype(np.uint64), np.array(offs, dtype=np.uint32),
                          np.concatenate(nb).astype(np.uint64) if nb else np.zeros(0, np.uint64))
        return out, st

    def search_layer_greedy(self, query, entry, layer):
        qa, qp = _f32(query)
        out = C.c_uint64(0)
        self._ck(self.L.hxo_search_layer_greedy(self.h, qp, entry, layer, C.byref(out)))
        return int(out.value)

    def search(self, query, k, ef=0, with_stats=False):
        qa, qp = _f32(query)
        ids = np.empty(max(k, 1), dtype=np.uint64)
        sc = np.empty(max(k, 1), dtype=np.float32)
        cnt = C.c_uint32(0)
        st = Stats()
        rc = self.L.hxo_search(self.h, qp, qa.size, k, ef, ids.ctypes.data_as(C.POINTER(C.c_uint64)),
                               sc.ctypes.data_as(C.POINTER(C.c_float)), C.byref(cnt), C.byref(st))
        self._ck(rc)
        n = int(cnt.value)
        if with_stats:
            return ids[:n].copy(), sc[:n].copy(), st.as_dict()
        return ids[:n].copy(), sc[:n].copy()

    def put_simhash(self, ids, bits):
        ia, ip = _u64(ids)
        ba, bp = _u64(bits)
        self._ck(self.L.hxo_index_put_simhash(self.h, ip, bp, ia.size))

    def search_policy(self, query, k, ef, cfg, query_simhash):
        """SearchSession::run with a non-exhaustive layer-0 policy; returns ids, scores, SearchStats, policy counters."""
        qa, qp = _f32(query)
        ids = np.empty(max(k, 1), dtype=np.uint64)
        sc = np.empty(max(k, 1), dtype=np.float32)
        cnt = C.c_uint32(0)
        st, ps = Stats(), PolicyStats()
        rc = self.L.hxo_search_policy(self.h, qp, qa.size, k, ef, C.byref(cfg), int(query_simhash),
                                      ids.ctypes.data_as(C.POINTER(C.c_uint64)), sc.ctypes.data_as(C.POINTER(C.c_float)),
                                      C.byref(cnt), C.byref(st), C.byref(ps))
        self._ck(rc)
        n = int(cnt.value)
        return ids[:n].copy(), sc[:n].copy(), st.as_dict(), ps.as_dict()

    def search_policy_batch(self, queries, k, ef, cfg, query_simhash, threads=1):
        qa, qp = _f32(queries)
        sa, sp = _u64(query_simhash)
        nq = qa.size // self.dim
        ids = np.zeros((nq, k), dtype=np.uint64)
        sc = np.zeros((nq, k), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        secs = self.L.hxo_search_policy_batch(self.h, qp, sp, nq, k, ef, C.byref(cfg), threads,
                                              ids.ctypes.data_as(C.POINTER(C.c_uint64)),
                                              sc.ctypes.data_as(C.POINTER(C.c_float)),
                                              cnt.ctypes.data_as(C.POINTER(C.c_uint32)))
        if secs < 0:
            raise OracleError(-1)
        return ids, sc, cnt, float(secs)

    def search_filtered_graph(self, query, k, cand_ids, query_simhash, ef=0, beam_percent=150, budgets=None):
        """restricted_filter_aware_search (restricted.rs:837-1148, directoryless).  Returns (ids, scores, stats dict)."""
        qa, qp = _f32(query)
        ca, cp = _u64(cand_ids)
        b = FilteredBudgets()
        if budgets is None:
            self.L.hxo_filtered_budgets(k, ef, beam_percent, ca.size, C.byref(b))
        else:
            b.ef_filtered, b.routing_rows, b.bridge_rows, b.vector_payloads, b.sampled_seeds = budgets
        ids = np.empty(max(k, 1), dtype=np.uint64)
        sc = np.empty(max(k, 1), dtype=np.float32)
        cnt = C.c_uint32(0)
        st = FilteredStats()
        self._ck(self.L.hxo_search_filtered_graph_budgets(self.h, qp, k, C.byref(b), cp, ca.size, int(query_simhash),
                                                          ids.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                          sc.ctypes.data_as(C.POINTER(C.c_float)), C.byref(cnt), C.byref(st)))
        n = int(cnt.value)
        return ids[:n].copy(), sc[:n].copy(), st.as_dict()

    def search_restricted(self, query, k, cand_ids):
        qa, qp = _f32(query)
        ca, cp = _u64(cand_ids)
        cap = max(1, min(k, max(ca.size, 1)))
        ids = np.empty(cap, dtype=np.uint64)
        sc = np.empty(cap, dtype=np.float32)
        cnt = C.c_uint32(0)
        nd = C.c_uint64(0)
        rc = self.L.hxo_search_restricted(self.h, qp, qa.size, k, cp, ca.size,
                                          ids.ctypes.data_as(C.POINTER(C.c_uint64)),
                                          sc.ctypes.data_as(C.POINTER(C.c_float)), C.byref(cnt), C.byref(nd))
        self._ck(rc)
        n = int(cnt.value)
        return ids[:n].copy(), sc[:n].copy()

    def search_exact(self, query, k):
        qa, qp = _f32(query)
        ids = np.empty(k, dtype=np.uint64)
        sc = np.empty(k, dtype=np.float32)
        cnt = C.c_uint32(0)
        self._ck(self.L.hxo_search_exact(self.h, qp, k, ids.ctypes.data_as(C.POINTER(C.c_uint64)),
                                         sc.ctypes.data_as(C.POINTER(C.c_float)), C.byref(cnt)))
        n = int(cnt.value)
        return ids[:n].copy(), sc[:n].copy()

    def search_batch(self, queries, k, ef=0, threads=1):
        qa, qp = _f32(queries)
        nq = qa.size // self.dim
        ids = np.zeros((nq, k), dtype=np.uint64)
        sc = np.zeros((nq, k), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        st = Stats()
        secs = self.L.hxo_search_batch(self.h, qp, nq, k, ef, threads, ids.ctypes.data_as(C.POINTER(C.c_uint64)),
                                       sc.ctypes.data_as(C.POINTER(C.c_float)),
                                       cnt.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(st))
        if secs < 0:
            raise OracleError(-1)
        return ids, sc, cnt, st.as_dict(), float(secs)

    def search_restricted_batch(self, queries, k, cand_ids, cand_offsets, threads=1):
        qa, qp = _f32(queries)
        nq = qa.size // self.dim
        ca, cp = _u64(cand_ids)
        oa, op = _u64(cand_offsets)
        ids = np.zeros((nq, k), dtype=np.uint64)
        sc = np.zeros((nq, k), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        secs = self.L.hxo_search_restricted_batch(self.h, qp, nq, k, cp, op, threads,
                                                  ids.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                  sc.ctypes.data_as(C.POINTER(C.c_float)),
                                                  cnt.ctypes.data_as(C.POINTER(C.c_uint32)))
        if secs < 0:
            raise OracleError(-1)
        return ids, sc, cnt, float(secs)

    def search_exact_batch(self, queries, k, threads=1):
        qa, qp = _f32(queries)
        nq = qa.size // self.dim
        ids = np.zeros((nq, k), dtype=np.uint64)
        sc = np.zeros((nq, k), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        secs = self.L.hxo_search_exact_batch(self.h, qp, nq, k, threads,
                                             ids.ctypes.data_as(C.POINTER(C.c_uint64)),
                                             sc.ctypes.data_as(C.POINTER(C.c_float)),
                                             cnt.ctypes.data_as(C.POINTER(C.c_uint32)))
        if secs < 0:
            raise OracleError(-1)
        return ids, sc, cnt, float(secs)

    def import_graph(self, levels, deg0, nbr0, layer0_stride, upper_node, upper_layer, upper_deg, upper_nbr,
                     upper_stride, entry_point, max_layer):
        la, lp = _u16(levels)
        da, dp = _u32(deg0)
        na, np_ = _u32(nbr0)
        una, unp = _u32(upper_node)
        ula, ulp = _u16(upper_layer)
        uda, udp = _u32(upper_deg)
        unb, unbp = _u32(upper_nbr)
        self._ck(self.L.hxo_index_import_graph(self.h, lp, dp, np_, layer0_stride, una.size, unp, ulp, udp, unbp,
                                               upper_stride, entry_point, max_layer))
