#!/usr/bin/env python
"""Generates tests/golden/hnsw_path_golden.json: frozen outputs of the oracle (oracle/hx_oracle.c — the CPU restatement of
the reference's algorithm, pinned by the reference's own known-answer tests in tests/test_oracle_kat.py) on inputs that are
fully determined by the reference's xorshift fixture generator (tests/index_lifecycle_scale.rs) and the layer sequence
stored in the file.

The reference itself is Rust and cannot be built here, so these are NOT outputs of the reference binary; they freeze the
pinned oracle so that (a) `-m "not gpu"` detects any drift of the oracle (compiler flags, refactors) and (b) `-m gpu` compares
the CUDA path with committed numbers, not only with whatever the oracle computes on the day.

    python tests/golden/make_golden.py        # rewrites the fixture (review the diff!)
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import hxo  # noqa: E402

N, DIM, M, M0, EFC = 1200, 48, 8, 16, 60
K, EF, NQ = 10, 40, 12
OUT = Path(__file__).with_name("hnsw_path_golden.json")


def inputs():
    rows = hxo.xorshift_vectors(1, N, DIM)                      # ids 1..N
    queries = hxo.xorshift_vectors(1_000_003, NQ, DIM)
    planes = hxo.xorshift_vectors(2_000_003, 64, DIM) - np.float32(0.5)
    ids = np.arange(1, N + 1, dtype=np.uint64)
    return ids, rows, queries, np.ascontiguousarray(planes, dtype=np.float32)


def scripted_levels():
    """select_layer(mod.rs:769-796) over a fixed uniform sequence (numpy's PCG64 stream 20240917); stored in the fixture so
    that the fixture does not depend on numpy's generator staying put."""
    rng = np.random.default_rng(20240917)
    ml = hxo.lib().hxo_default_ml_for_m(M)
    return [int(hxo.lib().hxo_select_layer_from_uniform(ml, float(u))) for u in rng.random(N, dtype=np.float32)]


def build(metric, ids, rows, levels):
    ix = hxo.Index(metric, DIM, m=M, m0=M0, ef_construction=EFC)
    for i in range(N):
        ix.insert(int(ids[i]), rows[i], int(levels[i]))
    return ix


def graph_digest(ix):
    graph, state = ix.export_graph()
    h = hashlib.sha256()
    for layer in sorted(graph):
        nodes, offs, nbrs = graph[layer]
        h.update(np.asarray(layer, np.uint32).tobytes())
        h.update(np.asarray(nodes, np.uint64).tobytes())
        h.update(np.asarray(offs, np.uint32).tobytes())
        h.update(np.asarray(nbrs, np.uint64).tobytes())
    h.update(np.asarray(state, np.uint64).tobytes())
    return h.hexdigest()


def hexbits(scores):
    return [format(int(b), "08x") for b in np.asarray(scores, dtype=np.float32).view(np.uint32)]


def compute(levels=None):
    ids, rows, queries, planes = inputs()
    levels = scripted_levels() if levels is None else levels
    out = {"_generator": "tests/golden/make_golden.py", "n": N, "dim": DIM, "m": M, "m0": M0, "ef_construction": EFC, "k": K,
           "ef": EF, "levels": levels, "metrics": {}}
    restricted = ids[::7].copy()
    filtered = ids[::5].copy()
    for name, metric in (("euclidean", hxo.EUCLIDEAN), ("cosine", hxo.COSINE), ("manhattan", hxo.MANHATTAN)):
        ix = build(metric, ids, rows, levels)
        ent = {"graph_sha256": graph_digest(ix), "strict": [], "restricted": []}
        for q in queries:
            oi, os_, st = ix.search(q, K, ef=EF, with_stats=True)
            ent["strict"].append({"ids": oi.tolist(), "score_bits": hexbits(os_),
                                  "stats": [st["expansion_steps"], st["neighbors_examined"], st["distance_computations"]]})
            ri, rs = ix.search_restricted(q, K, restricted)
            ent["restricted"].append({"ids": ri.tolist(), "score_bits": hexbits(rs)})
        if metric == hxo.COSINE:
            bits = np.array([hxo.simhash_from_planes(planes, rows[i]) for i in range(N)], dtype=np.uint64)
            ix.put_simhash(ids, bits)
            cfg = hxo.policy_defaults()
            ent["simhash_sha256"] = hashlib.sha256(bits.tobytes()).hexdigest()
            ent["default_mode"], ent["filtered_graph"] = [], []
            for q in queries:
                qs = hxo.simhash_from_planes(planes, q)
                pi, ps, _, pst = ix.search_policy(q, K, 100, cfg, qs)
                ent["default_mode"].append({"ids": pi.tolist(), "score_bits": hexbits(ps), "query_simhash": format(qs, "016x"),
                                            "policy_stats": {k_: int(v) for k_, v in pst.items()}})
                fi, fs, fst = ix.search_filtered_graph(q, K, filtered, qs, ef=100)
                ent["filtered_graph"].append({"ids": fi.tolist(), "score_bits": hexbits(fs),
                                              "stats": {k_: (v if isinstance(v, str) else int(v)) for k_, v in fst.items()}})
        out["metrics"][name] = ent
    return out


MUT_OUT = Path(__file__).with_name("mutation_golden.json")
N_DELETE, N_UPSERT = 150, 30


def compute_mutations(levels):
    """The write path after the build: VectorIndex::delete (mutation.rs:1606-2050: delete_from_layer, relink_neighbor,
    entry-candidate promotion) for a scripted list of ids that starts with the entry point, then Upsert (mutation.rs:653-661)
    of scripted ids with new vectors and layers.  Frozen per metric: the adjacency digest, the entry point and the strict
    top-k of the fixture's queries after each phase."""
    ids, rows, queries, _ = inputs()
    newrows = hxo.xorshift_vectors(3_000_003, N_UPSERT, DIM)
    out = {"_generator": "tests/golden/make_golden.py", "n_delete": N_DELETE, "n_upsert": N_UPSERT, "metrics": {}}
    for name, metric in (("euclidean", hxo.EUCLIDEAN), ("cosine", hxo.COSINE), ("manhattan", hxo.MANHATTAN)):
        ix = build(metric, ids, rows, levels)
        entry0 = ix.state()[0]
        victims = [entry0] + [int(i) for i in ids[3::8] if int(i) != entry0][:N_DELETE - 1]
        for v in victims:
            assert ix.delete(v)
        ent = {"deleted": victims[:4] + ["..."], "after_delete": {"graph_sha256": graph_digest(ix), "state": list(ix.state()),
                                                                "count": len(ix), "strict": []}}
        for q in queries:
            oi, os_ = ix.search(q, K, ef=EF)
            ent["after_delete"]["strict"].append({"ids": oi.tolist(), "score_bits": hexbits(os_)})
        targets = [int(i) for i in ids[5::40]][:N_UPSERT]              # some deleted before, some alive: both Upsert arms
        for j, t in enumerate(targets):
            ix.upsert(t, newrows[j], int(levels[(7 * j) % len(levels)]))
        ent["after_upsert"] = {"graph_sha256": graph_digest(ix), "state": list(ix.state()), "count": len(ix), "strict": []}
        for q in queries:
            oi, os_ = ix.search(q, K, ef=EF)
            ent["after_upsert"]["strict"].append({"ids": oi.tolist(), "score_bits": hexbits(os_)})
        out["metrics"][name] = ent
    return out


if __name__ == "__main__":
    data = compute()
    OUT.write_text(json.dumps(data, separators=(",", ":")) + "\n")
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes)")
    mut = compute_mutations(data["levels"])
    MUT_OUT.write_text(json.dumps(mut, separators=(",", ":")) + "\n")
    print(f"wrote {MUT_OUT} ({MUT_OUT.stat().st_size} bytes)")
