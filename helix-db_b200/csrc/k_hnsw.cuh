// k_hnsw.cuh — what every build of the HNSW traversal shares: the launch arguments, the sorted beam and the statement of
// why a sorted beam + tie stack is the reference's two heaps.  The kernels themselves are in k_hnsw_ring.cuh (strict
// exhaustive mode: warp-per-query and CTA-per-query builds) and k_hnsw_policy.cuh (production default), the insertion beam of
// the device build in k_build.cu.
//
// Restated, for one device-resident shard:
//   VectorIndex::search_layer_greedy            search/vector/search.rs:169-224
//   VectorIndex::search_layer0_with_simhash     search/vector/search.rs:267-1067, STRICT_EXHAUSTIVE=true
//   SearchSession::run (layer loop, take k)     search/vector/search.rs:1150-1156,1229
//
// Equivalence with the reference's two heaps (`candidates` min-heap, `w` max-heap bounded by ef):
//  * every admitted candidate is pushed to both heaps; `w` evicts its max when it exceeds ef.  An entry
//    evicted from `w` stays in `candidates`, but all unexpanded members of `w` order before it, and when it
//    is finally popped it triggers `current_dist > w.max` (break) — unless its score EQUALS w.max, in which
//    case the reference expands it.  So: beam = `w` kept sorted by (score,id) with an `expanded` bit;
//    candidates = unexpanded beam entries  +  a small stack of evicted-unexpanded entries whose score equals
//    the current w.max score (dropped as soon as w.max decreases)  +  a `dropped` flag that only matters for
//    the expansion_steps counter (the reference counts the popped entry that triggers the break).
//  * admission is sequential in neighbour-id order with the reference's exact test
//    `dist < w.max.score || len < ef` (search.rs:935) — score only, strict.
#pragma once
#include "hx_common.cuh"

#define HX_TIE_CAP 32

struct HxHnswArgs {
  const float* queries;     // [B][dim] device
  const float* q_hdr;       // [B] query header (cosine norm)
  const uint32_t* q_status; // [B] != 0 => query failed validation, skip
  uint32_t B, k, ef;
  uint64_t* out_ids;        // [B][k]
  float* out_scores;        // [B][k]
  uint32_t* out_counts;     // [B]
  uint32_t* q_stats;        // optional [B][4]: expansion_steps, neighbors_examined, distance_computations, upper_steps
  uint32_t* err_flags;
  uint32_t fr_cap;          // frontier capacity (>= max row length, multiple of 32)
  // ring build, host-buffer calls: validation fused into the search (each warp validates its own query) and the search
  // gated on `avail` (queries [0, *avail) have landed), so the kernel starts while the queries are still crossing PCIe
  uint32_t fused_validate;
  uint32_t* q_status_w;     // [B] written when fused_validate
  float* q_hdr_w;           // [B]
  float limit;
  int32_t has_limit;
  const uint32_t* avail;    // nullptr: everything is resident
  // ring builds: per-query outcome.  q_err[qi] = error flags raised by query qi alone (a tie-stack or visited-set overflow
  // fails that query, not the batch); done[qi] (host-mapped, service path) = 0x80000000 | flags once the results are visible
  uint32_t* q_err;          // optional [B]
  uint32_t* done;           // optional [B]
};

// ---- sorted beam in shared memory, maintained by warp 0 -----------------------------------------------
// key = score_bits << 32 | slot << 1 | expanded
struct HxBeam {
  uint64_t* a;
  uint32_t len;
};

// Insert `key` (not present) keeping ascending order; if full (len == ef) the max entry is evicted and
// returned through *evicted (else HX_KEY_MAX). Warp-cooperative: all 32 lanes call with identical arguments.
__device__ __forceinline__ void hx_beam_insert(HxBeam& b, uint32_t ef, uint64_t key, uint64_t* evicted, uint32_t lane) {
  uint32_t cnt = 0;
  for (uint32_t i = lane; i < b.len; i += 32) cnt += (b.a[i] < key) ? 1u : 0u;
  const uint32_t pos = hx_warp_sum(cnt);
  uint32_t end;  // entries [pos, end) move up by one
  if (b.len == ef) {
    *evicted = b.a[b.len - 1];
    end = b.len - 1;
  } else {
    *evicted = HX_KEY_MAX;
    end = b.len;
    b.len += 1;
  }
  __syncwarp();
  for (int hi = (int)end; hi > (int)pos; hi -= 32) {
    int i = hi - 1 - (int)lane;
    bool ok = i >= (int)pos;
    uint64_t v = 0;
    if (ok) v = b.a[i];
    __syncwarp();
    if (ok) b.a[i + 1] = v;
    __syncwarp();
  }
  if (lane == 0) b.a[pos] = key;
  __syncwarp();
}
