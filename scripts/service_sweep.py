#!/usr/bin/env python
"""Sweep the query service's launch shape on the C2 corpus (1M x 768 cosine, ef=100, k=10): concurrent one-query callers
through hx_service, bit-exactness against hx_search, q/s and latency percentiles.  Writes JSON lines to stdout."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import helix_db_b200 as hx  # noqa: E402
from helix_db_b200 import callers  # noqa: E402

SEED = 0x0DB9ED1A


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--queries", type=int, default=8192)
    ap.add_argument("--seconds", type=float, default=1.5)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    ix = hx.VectorIndex(hx.Metric.Cosine, hx.VectorIndexConfig("c2", "embedding", a.dim))
    t0 = time.perf_counter()
    ix.generate_vectors(0, a.n, SEED, 1024, 1.0, 32)
    ix.build(seed=SEED)
    print(json.dumps({"setup_s": round(time.perf_counter() - t0, 1)}), flush=True)
    q = ix.generate_queries(SEED, a.queries, first_query=10_000_000, n_centroids=1024, sigma=1.0, kind=32)
    k, ef = 10, 100
    ref_ids, ref_sc, ref_cnt = ix.search_batch(q, hx.SearchParams.strict(k, ef))
    shapes = [dict(ctas_per_sm=2), dict(ctas_per_sm=1), dict(ctas_per_sm=3), dict(ctas_per_sm=2, flags=1),
              dict(ctas_per_sm=2, flags=2), dict(ctas_per_sm=2, min_batch=32, batch_window_us=60),
              dict(ctas_per_sm=2, min_batch=8, batch_window_us=15), dict(ctas_per_sm=2, cta_warps=8),
              dict(ctas_per_sm=3, cta_warps=6)]
    if a.quick:
        shapes = shapes[:4]
    for shape in shapes:
        try:
            svc = ix.service(k, ef, capacity=2048, max_batch=128, **shape)
        except hx.HelixDbError as e:
            print(json.dumps({"shape": shape, "error": str(e)}), flush=True)
            continue
        info = svc.stats()
        for mode, ncall, nthr in (("tasks", 256, 8), ("tasks", 1024, 8), ("blocking", 256, 0), ("blocking", 64, 0),
                                  ("blocking", 16, 0), ("blocking", 1, 0)):
            rep, ids, sc, cnt = callers.run(svc, ix, q, k, ef, ncall, mode=mode, n_threads=nthr, seconds=a.seconds)
            same = bool(ids.tolist() == ref_ids.tolist() and sc.tobytes() == ref_sc.tobytes() and cnt.tolist() == ref_cnt.tolist())
            st = svc.stats()
            print(json.dumps({"shape": shape, "resolved": {x: info[x] for x in ("cta_warps", "rows_in_flight", "visited_cap",
                              "smem_bytes", "ctas_per_sm")}, **rep, "bit_exact_vs_hx_search": same,
                              "launches": st["launches"], "max_batch_seen": st["max_batch_seen"]}), flush=True)
        svc.close()
    # the round-1 path for comparison: blocking B = 1 hx_search per caller thread, no service
    for ncall in (1, 16, 64):
        rep, ids, sc, cnt = callers.run(None, ix, q[:2048], k, ef, ncall, mode="direct", seconds=a.seconds)
        print(json.dumps({"shape": "direct hx_search B=1 per thread", **rep}), flush=True)
    ix.close()


if __name__ == "__main__":
    main()
