#!/usr/bin/env python
"""Randomised insert / delete / upsert sequences on the oracle (test infrastructure), checking after every burst the
reference's structural invariants (V/index.rs:3617-3701: degree <= limit, no self link, strictly ascending rows), that no row
names a deleted node, that the entry point is a live node of the highest layer, and that searches only return live ids.
Meant to run on the sanitizer build of the checker: scripts/oracle_sanitize.sh."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import hxo  # noqa: E402

L = hxo.lib()
def check(ix, alive, lim_u, top_hint):
    st = ix.state()
    if not alive:
        assert st is None and len(ix) == 0
        return
    entry, top = st
    levels = {i: ix.node_level(i) for i in alive}
    assert entry in alive and top == max(levels.values()) == levels[entry]
    for layer in range(0, top_hint + 1):
        limit = ix.layer0_limit if layer == 0 else lim_u
        for i in alive:
            r = ix.neighbors(layer, i).tolist()
            assert len(r) <= limit and i not in r and all(a < b for a, b in zip(r, r[1:])), (layer, i, r)
            assert alive.issuperset(r), (layer, i, r)
for seed in range(6):
    rng = np.random.default_rng(seed)
    metric = [hxo.EUCLIDEAN, hxo.COSINE, hxo.MANHATTAN][seed % 3]
    dim, m = int(rng.integers(2, 20)), int(rng.integers(2, 9))
    ix = hxo.Index(metric, dim, m=m, m0=2 * m, ef_construction=max(2 * m, 16))
    ml = L.hxo_default_ml_for_m(max(m, 2))
    alive, next_id, top_hint = set(), 1, 0
    for step in range(700):
        op = rng.random()
        if op < 0.55 or len(alive) < 3:
            v = rng.standard_normal(dim).astype(np.float32) if seed % 2 else rng.integers(-2, 3, dim).astype(np.float32)
            if metric == hxo.COSINE and not np.any(v): v[0] = 1.0
            lv = int(L.hxo_select_layer_from_uniform(ml, float(rng.random(dtype=np.float32))))
            top_hint = max(top_hint, lv)
            ix.insert(next_id, v, lv); alive.add(next_id); next_id += 1
        elif op < 0.8:
            victim = int(rng.choice(sorted(alive)))
            assert ix.delete(victim) is True; alive.discard(victim)
            assert ix.delete(victim) is False
        else:
            target = int(rng.choice(sorted(alive)))
            v = rng.standard_normal(dim).astype(np.float32)
            lv = int(L.hxo_select_layer_from_uniform(ml, float(rng.random(dtype=np.float32))))
            top_hint = max(top_hint, lv)
            ix.upsert(target, v, lv)
        if step % 50 == 0:
            check(ix, alive, m, top_hint)
            if alive:
                q = rng.standard_normal(dim).astype(np.float32)
                got, _ = ix.search(q, 5, ef=32)
                assert alive.issuperset(got.tolist())
    for v in sorted(alive):
        assert ix.delete(v)
    check(ix, set(), m, top_hint)
    print('seed', seed, 'ok', next_id)
print('STRESS OK')
