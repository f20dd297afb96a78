#!/usr/bin/env python
"""bench.py — queries/sec @ recall@10, DBpedia-1M-shaped d=768 top-10 (BASELINE.json), on 1/2/4/8 B200.

A "step" is one pass of the hot path over one batch of Q independent single-query HNSW traversals
(config C2: 1M x 768 f32, cosine, m=16/m0=32/ef_construction=200, ef=100, k=10; no cross-query sharing).

  value   : whole-job queries/sec, queries already resident in HBM (hx_search_device), CUDA-event timed.
  e2e     : the same metric through the reference-facing C-ABI call hx_search with HOST (pinned) buffers:
            H2D of the queries and D2H of ids/scores/counts are inside the timed region.
  roofline: k_hnsw_search, algorithmic bytes E*(5+8*32) + Dc*(4+4d) per launch / CUDA-event kernel time,
            against the measured HBM peak in MEASURED_PEAKS.json.
  cpu_baseline / --impl reference: the CPU oracle (restatement of the reference's algorithm without its KV layer)
            traversing the IDENTICAL graph on all host cores, one query per thread.
  N > 1   : one process per GPU (torchrun).  Default "replica" mode partitions the QUERIES across full replicas
            (no data-path collective; weak scaling: Q queries per GPU per step).  The same line carries a
            "sharded" object: the 1M corpus split by id range, every rank searching every query on its shard,
            ONE NCCL all-gather of the per-shard top-k and the (score,id) merge kernel inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SEED = 0x0DB9ED1A
N_CENTROIDS = 1024
K = 10
EF = 100
# Synthetic corpus recipes (no network => no DBpedia download):
#  "embedding" (default): 1024-component Gaussian mixture in a 32-d latent space (sigma 1.0: overlapping clusters) pushed
#      through a fixed random 768x32 projection + 2 % isotropic noise, unit-normalised.  Intrinsic dimension ~32 like real
#      sentence embeddings; HNSW (m=16, ef=100) reaches recall@10 ~0.97 on it, as it does on DBpedia-OpenAI-1M.
#  "survey": SURVEY §8(d)'s literal example — 1024 isolated isotropic clusters in full 768-d (sigma 0.3).  Its landscape is
#      flat between clusters, so ANY greedy HNSW descent (reference algorithm included) lands in a wrong cluster for ~7 %
#      of the queries at 1M (profiles/r01_recall_diag_survey_mixture.json): recall@10 0.89 at ef=100, 0.91 at ef=200.
RECIPES = {"embedding": dict(kind=32, sigma=1.0), "survey": dict(kind=0, sigma=0.3)}
KIND = 32
SIGMA = 1.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--rows", dest="n", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--queries-per-step", type=int, default=32768)
    ap.add_argument("--metric", default="cosine", choices=["cosine", "euclidean"])
    ap.add_argument("--recall-queries", type=int, default=512)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sharded", action="store_true")
    ap.add_argument("--no-default-mode", action="store_true")
    ap.add_argument("--workload", default="hnsw", choices=["hnsw", "prefilter", "dense"])
    ap.add_argument("--recipe", default="embedding", choices=sorted(RECIPES))
    a = ap.parse_args()
    global KIND, SIGMA
    KIND, SIGMA = RECIPES[a.recipe]["kind"], RECIPES[a.recipe]["sigma"]
    return a


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms",
                                          "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx = max(mx, float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "samples": len(sm),
                "reasons": sorted(reasons)}


def available_cores():
    """Host threads this process can really use: affinity mask capped by the cgroup CPU quota (cpu.max)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        txt = Path("/sys/fs/cgroup/cpu.max").read_text().split()
        if txt[0] != "max":
            n = max(1, min(n, int(float(txt[0]) / float(txt[1]) + 0.5)))
    except Exception:
        pass
    return n


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel: str, shape: dict):
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu summary — only when the capture was taken on
    a launch of this very shape (a number for another launch size would not be "per launch like achieved"), else None."""
    p = ROOT / "profiles" / "ncu_summary.json"
    if p.exists():
        try:
            e = json.loads(p.read_text()).get(kernel, {})
            return e.get("dram_bytes_per_launch") if e.get("shape") == shape else None
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------------------------
def build_index(hx, args, device, first_id, n, storage=0):
    metric = hx.Metric.Cosine if args.metric == "cosine" else hx.Metric.Euclidean
    cfg = hx.VectorIndexConfig("dbpedia_1m_synthetic", "embedding", args.dim)   # m=16, m0=32, ef_c=200 defaults
    ix = hx.VectorIndex(metric, cfg, device=device, storage=storage)
    t0 = time.perf_counter()
    ix.generate_vectors(first_id, n, SEED, N_CENTROIDS, SIGMA, KIND)
    t1 = time.perf_counter()
    ix.build(seed=SEED)
    t2 = time.perf_counter()
    return ix, {"generate_s": round(t1 - t0, 2), "build_s": round(t2 - t1, 2)}


def oracle_from_device(hxo, ix, args):
    """Mirror the device index (vectors + graph) into the CPU oracle so both traverse identical adjacency."""
    metric = hxo.COSINE if args.metric == "cosine" else hxo.EUCLIDEAN
    ora = hxo.Index(metric, args.dim)
    g = ix.download_graph()
    n = g["n"]
    chunk = 65536
    for lo in range(0, n, chunk):
        ids, rows = ix.download_vectors(lo, min(chunk, n - lo))
        ora.put_vectors(ids, rows)
    ora.import_graph(g["levels"], g["deg0"], g["nbr0"], g["layer0_stride"], g["upper_node"], g["upper_layer"],
                     g["upper_deg"], g["upper_nbr"], g["upper_stride"], g["entry_point"], g["max_layer"])
    return ora


def exact_topk_device_full(hx, torch, ix, queries, n, first_id, k):
    """Ground truth by the exact scan kernel over ALL rows (same metric, same (score,id) tie rule).  The restricted
    entry point accepts at most 1e6 candidates (the reference's bound), so larger shards are scanned in 1M-row ranges
    whose top-k lists are merged by the (score,id) merge kernel.  Returns (ids u64 [B,k], scores f32 [B,k], counts)."""
    dev = torch.device("cuda", ix.device)
    B = len(queries)
    dq = torch.from_numpy(queries).to(dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    R = 1_000_000
    ranges = [(a, min(n, a + R)) for a in range(0, n, R)]
    a_ids = torch.zeros((len(ranges), B, k), dtype=torch.int64, device=dev)
    a_sc = torch.zeros((len(ranges), B, k), dtype=torch.float32, device=dev)
    a_cnt = torch.zeros((len(ranges), B), dtype=torch.int32, device=dev)
    step = 64
    for ri, (a, b) in enumerate(ranges):
        slots = torch.arange(a, b, dtype=torch.int32, device=dev)
        for lo in range(0, B, step):
            bb = min(step, B - lo)
            ix.search_restricted_device(dq[lo:lo + bb].data_ptr(), bb, hx.SearchParams.strict(k), slots.data_ptr(), 0,
                                        b - a, b - a, a_ids[ri, lo:lo + bb].data_ptr(), a_sc[ri, lo:lo + bb].data_ptr(),
                                        a_cnt[ri, lo:lo + bb].data_ptr(), stream)
        torch.cuda.synchronize(dev)
    ix.last_kernel_ms()
    if len(ranges) == 1:
        return a_ids[0].cpu().numpy().view(np.uint64), a_sc[0].cpu().numpy(), a_cnt[0].cpu().numpy()
    o_ids = torch.zeros((B, k), dtype=torch.int64, device=dev)
    o_sc = torch.zeros((B, k), dtype=torch.float32, device=dev)
    o_cnt = torch.zeros((B,), dtype=torch.int32, device=dev)
    hx.merge_topk_device(ix.device, a_ids.data_ptr(), a_sc.data_ptr(), a_cnt.data_ptr(), len(ranges), B, k,
                         o_ids.data_ptr(), o_sc.data_ptr(), o_cnt.data_ptr(), stream)
    torch.cuda.synchronize(dev)
    return o_ids.cpu().numpy().view(np.uint64), o_sc.cpu().numpy(), o_cnt.cpu().numpy()


def exact_topk_device(hx, torch, ix, queries, n, first_id, k):
    return exact_topk_device_full(hx, torch, ix, queries, n, first_id, k)[0]


def recall_at_k(found, truth):
    hit = 0
    for f, t in zip(found, truth):
        hit += len(set(f.tolist()) & set(t.tolist()))
    return hit / float(truth.size)


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import helix_db_b200 as hx

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream(dev).cuda_stream
    n, dim, Q, k = args.n, args.dim, args.queries_per_step, K
    hbm_peak, peak_src = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- setup (untimed): corpus on the device, graph built on the device ---------------------------------------
    ix, setup = build_index(hx, args, local_rank, 0, n)
    n_sets = args.steps + args.warmup
    # distinct queries every step and every rank (nothing can be answered from a previous step's cache lines)
    qsets = [ix.generate_queries(SEED, Q, first_query=(rank * n_sets + s) * Q, n_centroids=N_CENTROIDS, sigma=SIGMA, kind=KIND)
             for s in range(n_sets)]
    params = hx.SearchParams.strict(k, EF)
    d_q = [torch.from_numpy(q).to(dev) for q in qsets]
    o_ids = torch.zeros((Q, k), dtype=torch.int64, device=dev)
    o_sc = torch.zeros((Q, k), dtype=torch.float32, device=dev)
    o_cnt = torch.zeros((Q,), dtype=torch.int32, device=dev)

    def step_device(s):
        ix.search_device(d_q[s].data_ptr(), Q, params, o_ids.data_ptr(), o_sc.data_ptr(), o_cnt.data_ptr(), stream)

    # recall@10 on the first query set vs the exact scan
    rq = min(args.recall_queries, Q)
    truth = exact_topk_device(hx, torch, ix, qsets[0][:rq], n, 0, k)
    step_device(0)
    torch.cuda.synchronize(dev)
    recall = recall_at_k(o_ids[:rq].cpu().numpy().view(np.uint64), truth)

    # ---- value: W warm-up steps, then exactly K timed steps, barrier + synchronize on both sides ----------------------
    for s in range(args.warmup):
        step_device(s)
    barrier()
    ix.last_kernel_ms()                       # drop warm-up launches from the kernel-time accumulator
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        step_device(args.warmup + s)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    kernel_ms_total, kernel_launches = ix.last_kernel_ms()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * args.steps * Q / (ms_total / 1e3)

    # algorithmic bytes per launch: replay the timed query sets with the reference's counters switched on (untimed)
    st_sum = dict(expansion_steps=0, distance_computations=0, neighbors_examined=0, algorithmic_bytes=0)
    params_st = hx.SearchParams.strict(k, EF)
    params_st.collect_stats = True
    for s in range(args.steps):
        st = hx.SearchStats()
        ix.search_device(d_q[args.warmup + s].data_ptr(), Q, params_st, o_ids.data_ptr(), o_sc.data_ptr(),
                         o_cnt.data_ptr(), stream, st)
        for f in st_sum:
            st_sum[f] += int(getattr(st, f))
    ix.last_kernel_ms()
    bytes_per_launch = st_sum["algorithmic_bytes"] / args.steps
    kernel_ms = kernel_ms_total / max(kernel_launches, 1)
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0

    # ---- e2e: reference-facing C-ABI call with host (pinned) buffers -------------------------------------------------
    h_q = [torch.from_numpy(q).pin_memory() for q in qsets]
    h_ids = torch.zeros((Q, k), dtype=torch.int64).pin_memory()
    h_sc = torch.zeros((Q, k), dtype=torch.float32).pin_memory()
    h_cnt = torch.zeros((Q,), dtype=torch.int32).pin_memory()
    import ctypes as C
    L = hx.load_library()
    cp = params._c()

    def step_host(s):
        rc = L.hx_search(ix.h, C.cast(h_q[s].data_ptr(), C.POINTER(C.c_float)), Q, C.byref(cp),
                         C.cast(h_ids.data_ptr(), C.POINTER(C.c_uint64)), C.cast(h_sc.data_ptr(), C.POINTER(C.c_float)),
                         C.cast(h_cnt.data_ptr(), C.POINTER(C.c_uint32)), None)
        if rc != 0:
            raise RuntimeError(f"hx_search failed: {L.hx_last_error().decode()}")

    for s in range(args.warmup):
        step_host(s)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        step_host(args.warmup + s)
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    te = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps * Q / float(te.item())

    # ---- batch-1 single-stream latency (one query per call, sequential) ------------------------------------------------
    nb1 = 200
    for i in range(20):
        ix.search_device(d_q[0][i:i + 1].data_ptr(), 1, params, o_ids.data_ptr(), o_sc.data_ptr(), o_cnt.data_ptr(), stream)
    torch.cuda.synchronize(dev)
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record()
    for i in range(nb1):
        ix.search_device(d_q[0][i:i + 1].data_ptr(), 1, params, o_ids.data_ptr(), o_sc.data_ptr(), o_cnt.data_ptr(), stream)
    b1.record()
    torch.cuda.synchronize(dev)
    batch1_us = b0.elapsed_time(b1) / nb1 * 1e3
    ix.last_kernel_ms()

    # ---- production-default mode: SearchParams::new = SimHashMode::Adaptive (SimHash gate + sampling policy) ----------------
    # No network => no way to obtain the reference's StdRng(42) hyperplanes; a Gaussian table from numpy's seed 42 stands in
    # (the fingerprints are data to the kernel either way).  Host buffers through hx_search_ex, fingerprints of the queries
    # projected on the device inside the timed region.
    default_mode = None
    default_planes = None
    d_ids_first = None
    if not args.no_default_mode and args.metric == "cosine":
        planes = np.random.default_rng(42).standard_normal((64, dim)).astype(np.float32)
        ix.set_simhash_planes(planes)
        t0 = time.perf_counter()
        ix.compute_simhash()
        simhash_s = time.perf_counter() - t0
        pnew = hx.SearchParams.new(k)
        pnew.collect_stats = True
        st_d, ps_d = hx.SearchStats(), hx.PolicyStats()
        d_ids, _, d_cnt = ix.search_ex(qsets[0], pnew, stats=st_d, policy_stats=ps_d)
        d_ids_first = d_ids
        d_recall = recall_at_k(d_ids[:rq], truth)
        pnew.collect_stats = False
        cpn, poln = pnew._c(), pnew._policy()

        def step_default(s):
            rc = L.hx_search_ex(ix.h, C.cast(h_q[s].data_ptr(), C.POINTER(C.c_float)), Q, C.byref(cpn), C.byref(poln), None,
                                C.cast(h_ids.data_ptr(), C.POINTER(C.c_uint64)), C.cast(h_sc.data_ptr(), C.POINTER(C.c_float)),
                                C.cast(h_cnt.data_ptr(), C.POINTER(C.c_uint32)), None, None)
            if rc != 0:
                raise RuntimeError(f"hx_search_ex failed: {L.hx_last_error().decode()}")

        for s in range(args.warmup):
            step_default(s)
        barrier()
        kms_sum = 0.0
        t0 = time.perf_counter()
        for s in range(args.steps):
            step_default(args.warmup + s)
            kms_sum += ix.last_kernel_ms()[0]
        td = time.perf_counter() - t0
        # one query per call (the reference's usage: one query per tokio task)
        for i in range(10):
            ix.search_ex(qsets[0][i:i + 1], pnew)
        t0 = time.perf_counter()
        for i in range(100):
            ix.search_ex(qsets[0][i:i + 1], pnew)
        single_us = (time.perf_counter() - t0) / 100 * 1e6
        default_mode = {
            "params": "SearchParams::new(10): ef=100, SimHashMode::Adaptive, threshold 43, sampling 0.8, failure 0.1",
            "e2e_qps": round(args.steps * Q / td, 1), "kernel": "k_hnsw_search_policy",
            "kernel_ms_per_launch": round(kms_sum / args.steps, 4), "kernel_qps": round(args.steps * Q / (kms_sum / 1e3), 1),
            "recall_at_10": round(d_recall, 4),
            "distance_computations_per_query": round(st_d.distance_computations / Q, 1),
            "simhash_examined_per_query": round(ps_d.simhash_examined / Q, 1),
            "simhash_filtered_per_query": round(ps_d.simhash_filtered / Q, 1),
            "rng_draws_per_query": round(ps_d.rng_draws / Q, 2),
            "alg_GBps": round(st_d.algorithmic_bytes / (kms_sum / args.steps * 1e-3) / 1e9, 1) if kms_sum else None,
            "simhash_projection_s": round(simhash_s, 3),
            "single_query_us": round(single_us, 1),
            "note": "hyperplanes: numpy default_rng(42) Gaussian stand-in for the reference's StdRng(42) table",
        }
        default_planes = planes

    # ---- sharded path (north_star): id-range shards of the SAME corpus, one all-gather, merge ----------------------------
    sharded = None
    if world > 1 and not args.no_sharded:
        lo, hi = rank * n // world, (rank + 1) * n // world
        sx, s_setup = build_index(hx, args, local_rank, lo, hi - lo)
        # every rank searches the SAME queries (rank 0's sets) against its shard
        sq = [ix.generate_queries(SEED, Q, first_query=s * Q, n_centroids=N_CENTROIDS, sigma=SIGMA, kind=KIND) for s in range(n_sets)]
        d_sq = [torch.from_numpy(q).to(dev) for q in sq]
        from importlib import import_module
        sharding = import_module("helix_db_b200.sharding")
        searcher = sharding.ShardedSearcher(hx, sx, world, rank, Q, k, dev)

        def step_sharded(s):
            # local search -> ONE packed all-gather (12*k+4 bytes per query per shard) -> (score,id) merge kernel
            ids_, sc_, cnt_ = searcher.step(d_sq[s], params, stream)
            o_ids.copy_(ids_)

        truth_s = exact_topk_device(hx, torch, ix, sq[0][:rq], n, 0, k)
        step_sharded(0)
        torch.cuda.synchronize(dev)
        s_recall = recall_at_k(o_ids[:rq].cpu().numpy().view(np.uint64), truth_s)
        for s in range(args.warmup):
            step_sharded(s)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for s in range(args.steps):
            step_sharded(args.warmup + s)
        g1.record()
        barrier()
        ts = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        sharded = {"value": round(args.steps * Q / (float(ts.item()) / 1e3), 1), "unit": "queries/s",
                   "recall_at_10": round(s_recall, 4), "shard_vectors": hi - lo, "collective": "1 x all_gather_into_tensor "
                   f"of {(12 * k + 4) * Q} B per rank per step (NCCL) + hx_merge_topk_device",
                   "ms_per_step": round(float(ts.item()) / args.steps, 3), "shard_build_s": s_setup["build_s"]}
        sx.close()

    # ---- CPU baseline: the oracle on the box's host cores, same graph, bounded sample (rank 0, N = 1 only) -------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(args, ix, qsets[0], truth[:rq] if rq else None,
                           default_planes if default_mode is not None else None)
        if default_mode is not None and "default_mode_qps" in cpu:
            default_mode["cpu_port_qps"] = cpu.pop("default_mode_qps")
            default_mode["cpu_port_recall_at_10"] = cpu.pop("default_mode_recall")
            pi = cpu.pop("_default_ids")
            cpu.pop("default_mode_sample", None)
            default_mode["cpu_port_identical_to_device"] = bool(d_ids_first is not None and
                                                                pi.tolist() == d_ids_first[:len(pi)].tolist())

    impl = os.environ.get("HX_HNSW_IMPL", "ring")
    hnsw_kernel = {"ring": "k_hnsw_search_ring", "tma": "k_hnsw_search_tma", "ldg": "k_hnsw_search_warp"}.get(impl, impl)
    if rank == 0:
        line = {
            "metric": "queries/sec @ recall@10, DBpedia-1M d=768 top-10, 1/2/4/8 B200 vs CPU ref",
            "value": round(value, 1), "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "recall_at_10": round(recall, 4),
            "config": {
                "workload": f"C2: {n}x{dim} f32 {args.metric} HNSW top-10 (m=16, m0=32, ef_construction=200, ef={EF}), "
                            f"independent single-query traversals (batch=1 semantics, no cross-query sharing), "
                            f"{Q} queries per GPU per step",
                "queries_per_step_per_gpu": Q, "parallelism": "single GPU" if world == 1 else
                f"{world} full replicas, queries partitioned, no data-path collective",
                "l2": "corpus 3.07 GB >> 126 MB L2; distinct queries every step and rank",
                "graph": "built on the device (hx_index_build), identical adjacency mirrored into the CPU oracle",
                "data_recipe": (f"{args.recipe}: unit-normalised {N_CENTROIDS}-component Gaussian mixture, sigma={SIGMA}, "
                                + (f"rank-{KIND} latent space -> fixed random projection to {dim}-d + 2% noise"
                                   if KIND else f"isolated isotropic clusters in {dim}-d") + ", seed=0x0DB9ED1A"),
                "setup": setup,
            },
            "e2e": {"value": round(e2e_value, 1), "unit": "queries/s", "h2d_bytes_per_step": Q * dim * 4,
                    "d2h_bytes_per_step": Q * (k * 12 + 4) + Q * 4 + 4,
                    "api": "hx_search (C ABI, pinned host buffers, blocking)"},
            "gpu_launches": args.steps * (1 if impl == "ring" else 2),
            "launches_per_step": ({hnsw_kernel: 1} if impl == "ring" else {"k_validate_and_header": 1, hnsw_kernel: 1}),
            "roofline": {"bound": "hbm", "kernel": hnsw_kernel, "achieved": round(achieved, 1), "peak": hbm_peak,
                         "unit": "GB/s", "frac": round(achieved / hbm_peak, 4), "traffic": ncu_traffic("k_hnsw_search", {"queries": Q, "rows": n, "dim": dim, "ef": EF}),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": int(bytes_per_launch),
                         "kernel_ms_per_launch": round(kernel_ms, 4),
                         "expansions_per_query": round(st_sum["expansion_steps"] / (args.steps * Q), 1),
                         "distance_computations_per_query": round(st_sum["distance_computations"] / (args.steps * Q), 1)},
            "single_stream_batch1": {"latency_us": round(batch1_us, 1), "qps": round(1e6 / batch1_us, 1),
                                     "note": "one query per hx_search_device call, calls issued back to back"},
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if sharded is not None:
            line["sharded"] = sharded
        if default_mode is not None:
            line["default_mode"] = default_mode
        print(json.dumps(line), flush=True)
    ix.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(args, ix, queries, truth, planes=None):
    """The oracle (C restatement of the reference, no KV layer => an upper bound on the reference's CPU throughput)."""
    from oracle import hxo

    cores = available_cores()
    t0 = time.perf_counter()
    ora = oracle_from_device(hxo, ix, args)
    mirror_s = time.perf_counter() - t0
    probe = min(256, len(queries))
    _, _, _, _, secs = ora.search_batch(queries[:probe], K, EF, threads=cores)
    qps_probe = probe / max(secs, 1e-9)
    sample = int(max(probe, min(len(queries), qps_probe * args.cpu_seconds)))
    ids, sc, cnt, st, secs = ora.search_batch(queries[:sample], K, EF, threads=cores)
    rec = recall_at_k(ids[:len(truth)], truth) if truth is not None and len(truth) <= sample else None
    _, _, _, _, secs1 = ora.search_batch(queries[:min(sample, 256)], K, EF, threads=1)
    extra = {}
    if planes is not None:   # the production-default mode on the CPU: same fingerprints, same policy
        n = args.n
        ora.put_simhash(np.arange(n, dtype=np.uint64), ix.download_simhash(0, n))
        dq = min(sample, 4096)
        qsim = np.array([hxo.simhash_from_planes(planes, q) for q in queries[:dq]], dtype=np.uint64)
        cfg = hxo.policy_defaults()
        pi, _, pc, psecs = ora.search_policy_batch(queries[:dq], K, EF, cfg, qsim, threads=cores)
        extra["default_mode_qps"] = round(dq / psecs, 1)
        extra["default_mode_recall"] = round(recall_at_k(pi[:len(truth)], truth), 4) if truth is not None and len(truth) <= dq else None
        extra["default_mode_sample"] = dq
        extra["_default_ids"] = pi
    out = {"value": round(sample / secs, 1), "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"{sample} queries of the first step's set, one query per thread, {cores} threads, "
                      f"identical graph and vectors",
            "single_thread_qps": round(min(sample, 256) / secs1, 1), "recall_at_10": None if rec is None else round(rec, 4),
            "distance_computations_per_query": round(st["distance_computations"] / sample, 1),
            "mirror_s": round(mirror_s, 1)}
    out.update(extra)
    return out


# ------------------------------------------------------------------------------------------------------------------
def run_prefilter(args):
    """Config C3: 1M x 768 prefiltered top-10 (graph-label filter), exact brute-force scan path, single B200.

    A step = 100 queries, query b restricted to the label set {id : id mod 100 == b} (10 000 candidates = 1 % of the
    corpus each), so one step reads every row of the 3.07 GB corpus exactly once (>> 126 MB L2)."""
    import ctypes as C

    import torch

    import helix_db_b200 as hx

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream(dev).cuda_stream
    n, dim, k, sel = args.n, args.dim, K, 100
    B = sel
    hbm_peak, peak_src = measured_peaks()
    metric = hx.Metric.Cosine if args.metric == "cosine" else hx.Metric.Euclidean
    ix = hx.VectorIndex(metric, hx.VectorIndexConfig("dbpedia_1m_synthetic", "embedding", dim), device=local_rank)
    ix.generate_vectors(0, n, SEED, N_CENTROIDS, SIGMA, KIND)
    ix.load_graph(0, np.array([0], np.uint64), np.array([0, 0], np.uint32), np.zeros(0, np.uint64))
    ix.set_entry(0, 0)
    n_sets = args.steps + args.warmup
    qsets = [ix.generate_queries(SEED, B, first_query=(rank * n_sets + s) * B, n_centroids=N_CENTROIDS, sigma=SIGMA,
                                 kind=KIND) for s in range(n_sets)]
    cand_lists = [np.arange(b, n, sel, dtype=np.uint64) for b in range(B)]
    cand_ids = np.concatenate(cand_lists)
    offs = np.zeros(B + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(c) for c in cand_lists])
    per_q = int(max(len(c) for c in cand_lists))
    total = int(offs[-1])
    d_slots = torch.from_numpy(cand_ids.astype(np.uint32).view(np.int32)).to(dev)      # ids == slots here (first_id 0)
    d_offs = torch.from_numpy(offs.view(np.int64)).to(dev)
    d_q = [torch.from_numpy(q).to(dev) for q in qsets]
    o_ids = torch.zeros((B, k), dtype=torch.int64, device=dev)
    o_sc = torch.zeros((B, k), dtype=torch.float32, device=dev)
    o_cnt = torch.zeros((B,), dtype=torch.int32, device=dev)
    params = hx.SearchParams.strict(k)

    def step_device(s):
        ix.search_restricted_device(d_q[s].data_ptr(), B, params, d_slots.data_ptr(), d_offs.data_ptr(), total, per_q,
                                    o_ids.data_ptr(), o_sc.data_ptr(), o_cnt.data_ptr(), stream)

    for s in range(args.warmup):
        step_device(s)
    torch.cuda.synchronize(dev)
    ix.last_kernel_ms()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        step_device(args.warmup + s)
    e1.record()
    torch.cuda.synchronize(dev)
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    kms, kl = ix.last_kernel_ms()
    value = args.steps * B / (ms_total / 1e3)
    bytes_per_launch = total * (4 * dim + (4 if args.metric == "cosine" else 0))
    kernel_ms = kms / max(kl, 1)
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    dev_ids = o_ids.cpu().numpy().view(np.uint64).copy()
    dev_sc = o_sc.cpu().numpy().copy()

    # e2e: candidate ids, offsets and queries in pinned host memory -> hx_search_restricted_multi -> host results
    L = hx.load_library()
    cp = params._c()
    h_q = [torch.from_numpy(q).pin_memory() for q in qsets]
    h_c = torch.from_numpy(cand_ids.view(np.int64)).pin_memory()
    h_o = torch.from_numpy(offs.view(np.int64)).pin_memory()
    h_ids = torch.zeros((B, k), dtype=torch.int64).pin_memory()
    h_sc = torch.zeros((B, k), dtype=torch.float32).pin_memory()
    h_cnt = torch.zeros((B,), dtype=torch.int32).pin_memory()

    def step_host(s):
        rc = L.hx_search_restricted_multi(ix.h, C.cast(h_q[s].data_ptr(), C.POINTER(C.c_float)), B, C.byref(cp),
                                          C.cast(h_c.data_ptr(), C.POINTER(C.c_uint64)),
                                          C.cast(h_o.data_ptr(), C.POINTER(C.c_uint64)),
                                          C.cast(h_ids.data_ptr(), C.POINTER(C.c_uint64)),
                                          C.cast(h_sc.data_ptr(), C.POINTER(C.c_float)),
                                          C.cast(h_cnt.data_ptr(), C.POINTER(C.c_uint32)), None)
        if rc != 0:
            raise RuntimeError(L.hx_last_error().decode())

    for s in range(args.warmup):
        step_host(s)
    t0 = time.perf_counter()
    for s in range(args.steps):
        step_host(args.warmup + s)
    t1 = time.perf_counter()
    e2e_value = args.steps * B / (t1 - t0)
    same = bool(h_ids.numpy().view(np.uint64).tolist() == dev_ids.tolist() and h_sc.numpy().tobytes() == dev_sc.tobytes())

    # e2e with the label sets resident on the device (hx_candidates: uploaded + mapped once, like a label bitmap cached per
    # snapshot): per step only the queries and 24 bytes per query cross PCIe
    t0 = time.perf_counter()
    dsets = [ix.cache_candidates(hx.RestrictedVectorCandidates(c)) for c in cand_lists]
    cache_s = time.perf_counter() - t0
    set_arr = (C.c_void_p * B)(*[d.h for d in dsets])

    def step_host_sets(s):
        rc = L.hx_search_restricted_sets(ix.h, C.cast(h_q[s].data_ptr(), C.POINTER(C.c_float)), B, C.byref(cp), set_arr, B,
                                         C.cast(h_ids.data_ptr(), C.POINTER(C.c_uint64)),
                                         C.cast(h_sc.data_ptr(), C.POINTER(C.c_float)),
                                         C.cast(h_cnt.data_ptr(), C.POINTER(C.c_uint32)), None)
        if rc != 0:
            raise RuntimeError(L.hx_last_error().decode())

    for s in range(args.warmup):
        step_host_sets(s)
    t0 = time.perf_counter()
    for s in range(args.steps):
        step_host_sets(args.warmup + s)
    t1 = time.perf_counter()
    e2e_sets = args.steps * B / (t1 - t0)
    same_sets = bool(h_ids.numpy().view(np.uint64).tolist() == dev_ids.tolist() and h_sc.numpy().tobytes() == dev_sc.tobytes())

    cpu = None
    if not args.no_cpu and rank == 0:
        from oracle import hxo
        om = hxo.COSINE if args.metric == "cosine" else hxo.EUCLIDEAN
        ora = hxo.Index(om, dim)
        for lo in range(0, n, 65536):
            ids_, rows_ = ix.download_vectors(lo, min(65536, n - lo))
            ora.put_vectors(ids_, rows_)
        ora.set_entry(0, 0)
        cores = available_cores()
        qs = qsets[args.warmup + args.steps - 1]
        ci, cs, cc, secs = ora.search_restricted_batch(qs, k, cand_ids, offs, threads=cores)
        parity = bool(ci.tolist() == dev_ids.tolist() and cs.tobytes() == dev_sc.tobytes())
        cpu = {"value": round(B / secs, 1), "unit": "queries/s", "cores": cores, "kind": "port",
               "sample": f"the last step's {B} queries x {per_q} candidates, one query per thread, {cores} threads",
               "bit_exact_vs_device": parity}
    if rank == 0:
        line = {
            "metric": "queries/sec, DBpedia-1M d=768 prefiltered top-10 (graph-label filter), exact scan (config C3)",
            "value": round(value, 1), "unit": "queries/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "recall_at_10": 1.0,
            "config": {"workload": f"C3: {n}x{dim} f32 {args.metric}, {B} queries per step, query b restricted to "
                                   f"{{id : id mod {sel} == b}} ({per_q} candidates each): every row read once per step",
                       "l2": "3.07 GB streamed per step >> 126 MB L2", "recipe": args.recipe},
            "e2e": {"value": round(e2e_value, 1), "unit": "queries/s", "h2d_bytes_per_step": B * dim * 4 + total * 8 + (B + 1) * 8,
                    "d2h_bytes_per_step": B * (k * 12 + 4) + B * 4 + 4, "api": "hx_search_restricted_multi (C ABI, pinned host)",
                    "identical_to_device_path": same},
            "e2e_device_resident_sets": {"value": round(e2e_sets, 1), "unit": "queries/s",
                                         "h2d_bytes_per_step": B * dim * 4 + B * 24, "d2h_bytes_per_step": B * (k * 12 + 4) + B * 4 + 4,
                                         "api": "hx_search_restricted_sets (label sets uploaded once with hx_candidates_create)",
                                         "sets_upload_s": round(cache_s, 3), "identical_to_device_path": same_sets},
            "gpu_launches": args.steps * 3,
            "launches_per_step": {"k_validate_and_header": 1, "k_scan": 1, "k_select": 1},
            "roofline": {"bound": "hbm", "kernel": "k_scan", "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s",
                         "frac": round(achieved / hbm_peak, 4), "traffic": ncu_traffic("k_scan", {"queries": B, "candidates": per_q, "dim": dim}), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(bytes_per_launch), "kernel_ms_per_launch": round(kernel_ms, 4)},
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    ix.close()


# ------------------------------------------------------------------------------------------------------------------
def run_dense(args):
    """Configs C4/C5 shape: exhaustive top-10 of a large query batch through the tensor cores (hx_search_dense: tcgen05
    bf16 contraction -> per-run nominees -> exact fp32 re-rank).  With --gpus N the corpus (--n rows in total) is split by
    id range, every rank scores every query against its shard, ONE NCCL all-gather moves the per-shard top-k and the
    merge kernel selects the global top-k by (score, id)."""
    import torch

    import helix_db_b200 as hx

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream(dev).cuda_stream
    n_total, dim, k = args.n, args.dim, K
    B = args.queries_per_step if args.queries_per_step != 8192 else 1024
    lo, hi = rank * n_total // world, (rank + 1) * n_total // world
    n = hi - lo
    metric = hx.Metric.Cosine if args.metric == "cosine" else hx.Metric.Euclidean
    ix = hx.VectorIndex(metric, hx.VectorIndexConfig("dense", "embedding", dim), device=local_rank, storage=1)
    ix.generate_vectors(lo, n, SEED, N_CENTROIDS, SIGMA, KIND)
    ix.load_graph(0, np.array([lo], np.uint64), np.array([0, 0], np.uint32), np.zeros(0, np.uint64))
    ix.set_entry(lo, 0)
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    tf_peak = float(peaks.get("bf16_tflops", 1590.0))
    # every rank answers the SAME queries
    qsets = [ix.generate_queries(SEED, B, first_query=s * B, n_centroids=N_CENTROIDS, sigma=SIGMA, kind=KIND)
             for s in range(args.steps + args.warmup)]
    params = hx.SearchParams.strict(k)
    sharding = None
    if world > 1:
        from importlib import import_module
        sharding = import_module("helix_db_b200.sharding")
        pack = torch.zeros((B, 3 * k + 1), dtype=torch.int32, device=dev)
        apack = torch.zeros((world, B, 3 * k + 1), dtype=torch.int32, device=dev)
        o_ids = torch.zeros((B, k), dtype=torch.int64, device=dev)
        o_sc = torch.zeros((B, k), dtype=torch.float32, device=dev)
        o_cnt = torch.zeros((B,), dtype=torch.int32, device=dev)

    def merge_across_ranks(ids, sc, cnt, nq, pk, apk, oi, osc, ocn):
        sharding.pack_topk(torch.from_numpy(ids.view(np.int64).copy()).to(dev), torch.from_numpy(sc.copy()).to(dev),
                           torch.from_numpy(cnt.astype(np.int32)).to(dev), pk)
        sharding.all_gather_topk(pk, world, apk)
        a_ids, a_sc, a_cnt = sharding.unpack_topk(apk, k)
        hx.merge_topk_device(local_rank, a_ids.data_ptr(), a_sc.data_ptr(), a_cnt.data_ptr(), world, nq, k,
                             oi.data_ptr(), osc.data_ptr(), ocn.data_ptr(), stream)
        torch.cuda.synchronize(dev)
        return oi.cpu().numpy().view(np.uint64)

    def step(s):
        ids, sc, cnt = ix.search_dense_batch(qsets[s], params)
        kms = ix.last_kernel_ms()[0]
        if world == 1:
            return ids, kms
        return merge_across_ranks(ids, sc, cnt, B, pack, apack, o_ids, o_sc, o_cnt), kms

    for s in range(args.warmup):
        step(s)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    kms, t0 = 0.0, time.perf_counter()
    last = None
    for s in range(args.steps):
        last, km = step(args.warmup + s)
        kms += km
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    tw = torch.tensor([wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    wall = float(tw.item())
    clocks = sampler.stop() if rank == 0 else None
    ldb = (dim + 63) // 64 * 64
    flop = 2.0 * B * n * ldb
    kernel_ms = kms / args.steps
    # recall of the (merged) answer vs the exact scan of every shard, merged the same way
    rq = min(64, B)
    gi, gs, gc = exact_topk_device_full(hx, torch, ix, qsets[args.warmup + args.steps - 1][:rq], n, lo, k)
    if world > 1:
        p2 = torch.zeros((rq, 3 * k + 1), dtype=torch.int32, device=dev)
        ap2 = torch.zeros((world, rq, 3 * k + 1), dtype=torch.int32, device=dev)
        t_o = torch.zeros((rq, k), dtype=torch.int64, device=dev)
        t_s = torch.zeros((rq, k), dtype=torch.float32, device=dev)
        t_c = torch.zeros((rq,), dtype=torch.int32, device=dev)
        truth = merge_across_ranks(gi, gs, gc, rq, p2, ap2, t_o, t_s, t_c)
    else:
        truth = gi
    rec = recall_at_k(last[:rq], truth)
    if rank == 0:
        line = {"metric": "queries/sec, exhaustive top-10 through the tensor cores (configs C4/C5 shape)",
                "value": round(args.steps * B / wall, 1), "unit": "queries/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(wall / args.steps * 1e3, 3), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16 (fp32 accumulate) + f32 exact re-rank",
                "data": "synthetic", "recall_at_10": round(rec, 4),
                "config": {"workload": f"dense: {B} queries x {n_total} rows x d={dim}, k={k}, host buffers (hx_search_dense)"
                                       + (f", {world} id-range shards of {n} rows, 1 all-gather + merge" if world > 1 else ""),
                           "recipe": args.recipe},
                "roofline": {"bound": "tensor", "kernel": "k_dense_scores", "achieved": round(flop / (kernel_ms * 1e-3) / 1e12, 1),
                             "peak": tf_peak, "unit": "TFLOP/s", "frac": round(flop / (kernel_ms * 1e-3) / 1e12 / tf_peak, 4),
                             "traffic": ncu_traffic("k_dense_scores", {"queries": B, "rows": n, "dim": dim}), "flop_per_launch_per_gpu": flop,
                             "kernel_ms_per_launch": round(kernel_ms, 4),
                             "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst); per GPU"},
                "gpu_launches": args.steps * 7, "clocks": clocks}
        print(json.dumps(line), flush=True)
    ix.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Rust reference cannot be built
    here: no rustc/cargo, un-vendored SlateDB fork) on all host cores, same config, metric and unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    import helix_db_b200 as hx
    from oracle import hxo

    if not torch.cuda.is_available():
        print(json.dumps({"impl": "reference", "unavailable": "the graph of config C2 is built on the device "
                          "(hours on the CPU); no GPU visible in this process"}))
        return
    n, dim, Q = args.n, args.dim, args.queries_per_step
    ix, setup = build_index(hx, args, 0, 0, n)
    cores = available_cores()
    queries = ix.generate_queries(SEED, Q, first_query=0, n_centroids=N_CENTROIDS, sigma=SIGMA, kind=KIND)
    rq = min(args.recall_queries, Q)
    truth = exact_topk_device(hx, torch, ix, queries[:rq], n, 0, K)
    ora = oracle_from_device(hxo, ix, args)
    ix.close()
    # a step = a bounded sample of the workload sized so that the whole run ends within a few minutes
    _, _, _, _, secs = ora.search_batch(queries[:256], K, EF, threads=cores)
    per_step = int(max(256, min(Q, (256 / max(secs, 1e-9)) * (args.cpu_seconds / max(args.steps, 1)))))
    for _ in range(args.warmup):
        ora.search_batch(queries[:per_step], K, EF, threads=cores)
    total = 0.0
    last = None
    for _ in range(args.steps):
        last = ora.search_batch(queries[:per_step], K, EF, threads=cores)
        total += last[4]
    value = args.steps * per_step / total
    rec = recall_at_k(last[0][:min(rq, per_step)], truth[:min(rq, per_step)])
    line = {
        "impl": "reference",
        "metric": "queries/sec @ recall@10, DBpedia-1M d=768 top-10, 1/2/4/8 B200 vs CPU ref",
        "value": round(value, 1), "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(total / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "recall_at_10": round(rec, 4),
        "config": {"workload": f"C2: {n}x{dim} f32 {args.metric} HNSW top-10 (m=16, m0=32, ef_construction=200, ef={EF}), "
                               f"one query per host thread, {per_step} queries per step (bounded sample of the {Q}-query step)",
                   "graph": "built on the device (setup only), identical adjacency for both arms", "setup": setup},
        "cpu_baseline": {"value": round(value, 1), "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} queries per step x {args.steps} steps, {cores} threads"},
        "e2e": {"value": round(value, 1), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "prefilter":
        run_prefilter(a)
    elif a.workload == "dense":
        run_dense(a)
    else:
        run_ours(a)
