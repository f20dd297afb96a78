#!/usr/bin/env python
"""Regenerates BASELINE.md §4 (round-2 measured table) from the committed bench lines under profiles/."""
import json
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
P = ROOT / "profiles"


def last(path):
    return json.loads((P / path).read_text().strip().split("\n")[-1])


d, n2, n8 = last("r02_bench_n1_final.json"), last("r02_bench_n2.json"), last("r02_bench_n8.json")
ref, ref8 = last("r02_bench_reference_arm.json"), last("r02_bench_n8_reference_arm.json")
fast, densefast, c5 = last("r02_bench_n1_c2_only_fast_box.json"), last("r02_bench_dense_c4shard_16warp_fast_box.json"), last("r02_bench_dense_c5_shape_2m_rows.json")
pf, cc, dc, e, dm = d["prefilter"], d["concurrent_callers"]["runs"], d["dense_c4"], d["euclid_d1536"], d["default_mode"]


def run(mode, c):
    return [r for r in cc if r["mode"] == mode and r["callers"] == c][0]


def k(x):
    return f"{x / 1e3:.0f} k"


def M(x):
    return f"{x / 1e6:.3f} M"


b256, t256, t1024, b1 = run("blocking", 256), run("tasks", 256), run("tasks", 1024), run("blocking", 1)
shapes = {s["shape"]: s for s in pf["reference_shapes"]}
rows = []
rows.append(
    f"| C2 1M×768 cosine HNSW top-10, ef=100, strict, 32 768 independent single-query traversals per step | 1 | **{M(d['value'])} q/s** "
    f"({d['ms_per_step']:.1f} ms/step, one launch per step; SM clock {d['clocks']['sm_mhz']:.0f} MHz under `sw_power_cap` — {M(fast['value'])} q/s on a "
    f"box that held {fast['clocks']['sm_mhz']:.0f} MHz, `r02_bench_n1_c2_only_fast_box.json`) | **{M(d['e2e']['value'])} q/s** (`hx_search`, pinned host "
    f"buffers, pipelined upload) | {d['recall_at_10']} | `k_hnsw_search_ring`: {d['roofline']['achieved']:.0f} GB/s algorithmic = "
    f"**{d['roofline']['frac']:.2f}** of measured HBM peak ({fast['roofline']['frac']:.2f} on the faster box) | {d['cpu_baseline']['value']:.0f} q/s "
    f"(16 threads), {d['cpu_baseline']['single_thread_qps']:.0f} q/s (1 thread); `--impl reference`: {ref['value']:.0f} q/s — **ids and score bits "
    f"identical to the device on all {d['parity']['queries_checked']} queries** (`parity`), ground truth identical to the oracle's exact scan |")
rows.append(
    f"| C2, the reference's calling pattern: ONE query per call from N concurrent callers (`hx_service`) | 1 | 256 blocking OS threads "
    f"**{k(b256['qps'])} q/s**, p50 {b256['p50_us']:.0f} µs, p99 {b256['p99_us']:.0f} µs; 256 logical callers on 8 threads (submit/poll) "
    f"**{k(t256['qps'])} q/s**, p99 {t256['p99_us']:.0f} µs; 1024 callers {k(t1024['qps'])} q/s; 1 caller {b1['qps']:.0f} q/s ({b1['p50_us']:.0f} µs) | same "
    f"(queries and results cross PCIe through host-mapped slots) | 0.9695, every answer bit-identical to `hx_search` | `k_hnsw_search_cta_ring`, "
    f"2 CTAs/SM | round 1: ≈ 9 k q/s (4 scratch sets, one blocking launch per query) |")
rows.append(
    f"| C2, production default `SearchParams::new` | 1 | **{k(dm['kernel_qps'])} q/s** (kernel; round 1 and mid-round 2: 820 k) | {k(dm['e2e_qps'])} q/s "
    f"(`hx_search_ex`); one query per call {dm['single_query_us']:.0f} µs | {dm['recall_at_10']} @ef=100, 0.982 @150, 0.989 @200 (the CPU port gets the same "
    f"value with identical ids: it is the mode's recall, not the device's) | `k_hnsw_search_policy`: {dm['alg_GBps']:.0f} GB/s algorithmic = "
    f"{dm['alg_GBps'] / 6572.2:.2f} of HBM peak; instruction-latency bound (DESIGN §8) | CPU port {dm['cpu_port_qps']:.0f} q/s (16 threads), ids identical |")
rows.append(
    f"| C3 1M×768 cosine, 100 queries × 10 000 label-filtered candidates | 1 | **{k(pf['value'])} q/s** ({pf['ms_per_step']:.2f} ms/step, ONE launch: scan + "
    f"warp-shuffle top-k) | {k(pf['e2e']['value'])} q/s with 8.3 MB of candidate ids crossing PCIe per step; **{k(pf['e2e_device_resident_sets']['value'])} "
    f"q/s** with the label sets resident (`hx_search_restricted_sets`) | 1.0 (exact) | `k_scan_topk`: {pf['roofline']['achieved']:.0f} GB/s = "
    f"**{pf['roofline']['frac']:.3f}** of HBM peak; ncu traffic 1.003× algorithmic | {pf['cpu_baseline']['value']:.0f} q/s (16 threads), bit-exact |")
for nm in ("prefilter-100", "prefilter-1000", "prefilter-10000", "prefilter-100000"):
    s = shapes[nm]
    a = s.get("acorn_walk")
    walk = "—" if not a else (
        f"filter-aware walk (`k_filtered_walk`, the reference's plan for this size): {k(a['qps'])} q/s vs exact scan {k(a['exact_scan_qps_same_queries'])} q/s "
        f"on the same 296 queries; walk recall {a['recall_at_10_vs_exact']:.2f}, {a['vectors_scored_per_query']:.0f} rows scored per query; bit-exact vs the "
        f"oracle's walk")
    rows.append(
        f"| reference shape `{nm}` (contiguous ids, `index_lifecycle_scale.rs:583-613`), 64 queries per step | 1 | {k(s['value'])} q/s (exact scan) | "
        f"{k(s['e2e'])} q/s | 1.0 | {walk} | {s['cpu_port_qps_1thread']:.0f} q/s (1 thread), bit-exact |")
rows.append(
    f"| C4 shard shape: 1024 queries × 1.25 M × 768 exhaustive top-10, tensor cores + exact f32 re-rank | 1 | **{k(dc['value'])} q/s** "
    f"({dc['ms_per_step']:.2f} ms/step) | {k(dc['e2e']['value'])} q/s ({dc['e2e']['ms_per_step']:.2f} ms/step, `hx_search_sharded(HX_SHARD_DENSE)`) | 1.0 vs "
    f"exact scan | `k_dense_scores` (tcgen05 cta_group::2 / TMEM / TMA): {dc['roofline']['kernel_ms_per_launch']:.2f} ms = {dc['roofline']['achieved']:.0f} "
    f"TFLOP/s = **{dc['roofline']['frac']:.2f}** of burst bf16 peak; {densefast['roofline']['kernel_ms_per_launch']:.2f} ms = "
    f"{densefast['roofline']['frac']:.2f} on a box that held its clocks (`r02_bench_dense_c4shard_16warp_fast_box.json`); round 1: 0.68 | — |")
rows.append(
    f"| C5 shard shape: 4096 queries × 2 M × 1536 (a real C5 shard has 12.5 M rows: time per row is what the kernel sees) | 1 | {k(c5['value'])} q/s "
    f"({c5['ms_per_step']:.1f} ms/step) | — | 1.0 vs exact scan | `k_dense_scores`: {c5['roofline']['kernel_ms_per_launch']:.1f} ms = "
    f"{c5['roofline']['achieved']:.0f} TFLOP/s = **{c5['roofline']['frac']:.2f}** of burst bf16 peak | — |")
rows.append(
    f"| 1M×1536 Euclidean HNSW top-10 (`index_lifecycle_scale.rs:1413` shape), 16 384 queries per step | 1 | {k(e['value'])} q/s ({e['ms_per_step']:.1f} "
    f"ms/step) | {k(e['e2e']['value'])} q/s (`hx_search`, pinned) | {e['recall_at_10']} | `k_hnsw_search_ring`: **{e['roofline']['frac']:.2f}** of HBM peak | "
    f"ids identical to the oracle on the CPU sample |")
rows.append(
    f"| C2 on 2 GPUs, replicas (queries partitioned, no collective) | 2 | **{M(n2['value'])} q/s** ({n2['ms_per_step']:.1f} ms/step) | "
    f"{M(n2['e2e']['value'])} q/s | 0.9695 | {n2['roofline']['frac']:.2f} | — |")
sh = n2["sharded"]
rows.append(
    f"| C2 on 2 GPUs, id-range shards behind the C ABI (`hx_shard_group`: local search → ONE `ncclAllGather` of 4 063 232 B per rank → merge kernel) | 2 | "
    f"{M(sh['value'])} q/s at ef = {sh['ef_per_shard']} per shard | — | {sh['recall_at_10']} (unsharded 0.9695) | — | — |")
d4 = n2["dense_c4"]
rows.append(
    f"| C4 shape on 2 GPUs (2 × 1.25 M rows, all-gather of per-shard top-10) | 2 | {k(d4['value'])} q/s ({d4['ms_per_step']:.2f} ms/step) | "
    f"{k(d4['e2e']['value'])} q/s | 1.0 | {d4['roofline']['frac']:.2f} | — |")
sh, dc8 = n8["sharded"], n8["dense_c4"]
rows.append(
    f"| C2 on 8 GPUs, replicas (queries partitioned, no collective), 32 768 queries per GPU per step | 8 | **{n8['value'] / 1e6:.2f} M q/s** "
    f"({n8['ms_per_step']:.1f} ms/step) | {n8['e2e']['value'] / 1e6:.2f} M q/s | 0.9695 | {n8['roofline']['frac']:.2f} | `--impl reference` on that box: "
    f"{ref8['value']:.0f} q/s ({ref8['cpu_baseline']['cores']} threads) |")
rows.append(
    f"| C2 on 8 GPUs, id-range shards (125 k rows each) behind the C ABI, per-shard ef tuned to iso-recall | 8 | **{sh['value'] / 1e6:.2f} M q/s** at "
    f"ef = {sh['ef_per_shard']} per shard ({sh['ms_per_step']:.1f} ms/step; {sh['value_at_unsharded_ef'] / 1e6:.2f} M q/s at the unsharded ef = 100; round 1: "
    f"0.87 M) | — | {sh['recall_at_10']} (unsharded 0.9695 on the same queries; ef = 24 gives 0.9621) | — | — |")
rows.append(
    f"| C4: 10 M × 768 sharded over 8 GPUs, 1024 queries, tensor cores + exact re-rank + ONE all-gather of per-shard top-10 | 8 | "
    f"{dc8['value'] / 1e3:.0f} k q/s (**{dc8['ms_per_step']:.2f} ms/step**; round 1: 3.49 ms) | {dc8['e2e']['value'] / 1e3:.0f} k q/s "
    f"({dc8['e2e']['ms_per_step']:.2f} ms/step, host buffers, every rank blocking) | 1.0 vs exact scan | `k_dense_scores` "
    f"{dc8['roofline']['kernel_ms_per_launch']:.2f} ms = {dc8['roofline']['frac']:.2f} of burst bf16 peak | — |")

hdr = ("| Config | GPUs | value (HBM-resident inputs) | e2e (host buffers through the C ABI) | recall@10 | roofline (dominant kernel) | CPU oracle, "
       "same box (threads) |\n|---|---|---|---|---|---|---|\n")
sec = f"""
## 4. Measured in round 2 (B200; `profiles/r02_*`; table generated by `scripts/baseline_table.py`)

Same corpus recipe, peaks and CPU port as §3.  One-GPU rows: `profiles/r02_bench_n1_final.json` (one default
`python bench.py` run: the C2 line with its `prefilter`, `concurrent_callers`, `dense_c4`, `euclid_d1536`, `default_mode` and
`parity` objects); multi-GPU rows: `profiles/r02_bench_n2.json` / `profiles/r02_bench_n8.json` (`--gpus N` under torchrun);
reference arm: `profiles/r02_bench_reference_arm.json`.  SM clocks differ between boxes of the pool (1.6–1.9 GHz under
`sw_power_cap`): across this round's runs the traversal kernel measured 1.00–1.05 M q/s and the dense kernel 1.45–1.53 ms.

""" + hdr + "\n".join(rows) + """

What changed against round 1, in one line each: strict-mode ids are compared with the oracle on every sampled query inside
the bench (it exits non-zero on a mismatch); one-query-per-call traffic went from ≈ 9 k to ≈ 500 k q/s (`hx_service`); C3 is one
launch at 0.99 of the HBM peak; the dense kernel went from 0.68 to 0.81–0.86 of the burst bf16 peak (CTA pairs, 16-warp
epilogue) and its step from 3.49 to 1.9 ms; the default-mode kernel lost its redundant top-k list (the beam's prefix is the
reference's tracker) and is now faster than the strict one, as a mode that reads fewer rows should be; the sharded path
lives behind the C ABI with NCCL inside the library and a per-shard beam tuned to iso-recall (8 GPUs: 0.87 → 2.76 M q/s);
the first-generation traversal kernels are gone (Manhattan and oversized rows run in the ring builds);
`compute-sanitizer` memcheck / racecheck / synccheck are clean (`profiles/r02_sanitizer_*`).  Build quality: at 200 k × 768
the batched device build and the sequential-equivalent build (the reference's insertion order) give recall@10 0.9818 vs
0.9830 at ef = 100 and 0.9973 both at ef = 200 (`profiles/r02_build_quality_batched_vs_sequential_200k.jsonl`).
"""
p = ROOT / "BASELINE.md"
s = p.read_text()
if "\n## 4. Measured in round 2" in s:
    s = s[:s.index("\n## 4. Measured in round 2")]
p.write_text(s.rstrip("\n") + "\n" + sec)
print("BASELINE.md §4 rewritten")
