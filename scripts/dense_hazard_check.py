#!/usr/bin/env python
"""Dense path, id-clustered neighbours: every hazard query has its 11 nearest rows at CONSECUTIVE ids inside one
64-row stripe of a 256-row tile.  A build that places tile rows in id order (HXD_INTERLEAVE = 0) keeps only HXD_T = 8 rows of
that (run, column-quarter) bucket and must lose true neighbours; the interleaved placement must not.  Prints one JSON line;
run once per library build (HELIX_B200_LIB selects the build).  Ground truth: the oracle's exact scan."""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import helix_db_b200 as hx  # noqa: E402
from oracle import hxo  # noqa: E402


def main():
    out = {"lib": os.environ.get("HELIX_B200_LIB", "default"), "version": hx.version()}
    # stride 1: consecutive ids (what real corpora do); stride 4: ids congruent mod 4 inside one tile — the adversarial layout
    # for the interleaved placement (every near-duplicate lands in the same column quarter again)
    for stride in (1, 4):
        out[f"stride_{stride}"] = fixture(stride)
    print(json.dumps(out))


def fixture(stride):
    rng = np.random.default_rng(77)
    n, dim, k, nh, nr = 40_000, 64, 10, 16, 240
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    hq = rng.standard_normal((nh, dim)).astype(np.float32)
    for h in range(nh):
        base = 2048 * h + 192                                  # tile row 192 + stride * i, i < 11: inside one 256-row tile
        for i in range(11):
            rows[base + stride * i] = hq[h] + np.float32(0.01 * (i + 1)) * rng.standard_normal(dim).astype(np.float32)
    queries = np.concatenate([hq, rng.standard_normal((nr, dim)).astype(np.float32)])
    ids = np.arange(n, dtype=np.uint64)
    out = {}
    for gm, om, name in ((hx.Metric.Cosine, hxo.COSINE, "cosine"), (hx.Metric.Euclidean, hxo.EUCLIDEAN, "euclidean")):
        gpu = hx.VectorIndex(gm, hx.VectorIndexConfig("hz", "embedding", dim), storage=1)
        gpu.load_vectors(ids, rows)
        gpu.load_graph(0, ids[:1], [0, 0], [])                 # the dense path needs a populated index, not a graph
        gpu.set_entry(int(ids[0]), 0)
        ora = hxo.Index(om, dim)
        ora.put_vectors(ids, rows)
        di, ds, dc = gpu.search_dense_batch(queries, hx.SearchParams.strict(k))
        hit_h = hit_r = exact_scores = 0
        for q in range(len(queries)):
            oi, os_ = ora.search_exact(queries[q], k)
            want = {int(i): os_[j].tobytes() for j, i in enumerate(oi)}
            got = di[q, :dc[q]].tolist()
            inter = len(set(got) & set(want))
            if q < nh:
                hit_h += inter
            else:
                hit_r += inter
            exact_scores += sum(1 for j, i in enumerate(got) if int(i) in want and ds[q, j].tobytes() == want[int(i)])
        out[name] = {"hazard_recall": hit_h / float(nh * k), "random_recall": hit_r / float(nr * k),
                     "returned_scores_bit_exact": exact_scores == hit_h + hit_r}
        gpu.close()
    return out


if __name__ == "__main__":
    main()
