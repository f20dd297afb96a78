// k_hnsw.cuh — HNSW traversal kernel: upper-layer greedy descent + layer-0 beam search.
//
// Restates, for one device-resident shard:
//   VectorIndex::search_layer_greedy            search/vector/search.rs:169-224
//   VectorIndex::search_layer0_with_simhash     search/vector/search.rs:267-1067, STRICT_EXHAUSTIVE=true
//   SearchSession::run (layer loop, take k)     search/vector/search.rs:1150-1156,1229
//
// One CTA (256 threads = 32 octets) per query; CTAs are persistent over the query list.
// Per layer-0 expansion:   warp 0 picks the nearest unexpanded beam entry, reads its neighbour row
// (one coalesced 128-byte load for m0 = 32), filters it against the visited stamps and marks the
// survivors  ->  all 32 octets score one neighbour each with the bit-exact octet kernel (24 x LDG.128
// per thread for d = 768, all independent: ~98 KB in flight per CTA)  ->  warp 0 admits the scores in
// neighbour-id order into the sorted beam.
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

#define HX_HNSW_THREADS 256
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
  uint8_t* stamps;          // [gridDim.x][stamp_stride] visited stamps
  uint32_t* epochs;         // [gridDim.x]
  size_t stamp_stride;
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

// NB = 128-bit loads in flight per thread while scoring (8: throughput build, 4 CTAs/SM; 24: latency build for small batches)
template <int METRIC, int NB>
__global__ void __launch_bounds__(HX_HNSW_THREADS) k_hnsw_search(HxDev ix, HxHnswArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* sq = reinterpret_cast<float*>(smem);                                 // [ld]
  uint64_t* beam_mem = reinterpret_cast<uint64_t*>(smem + (size_t)ix.ld * 4); // [ef]
  uint64_t* tie = beam_mem + a.ef;                                            // [HX_TIE_CAP]
  uint32_t* frontier = reinterpret_cast<uint32_t*>(tie + HX_TIE_CAP);         // [fr_cap]
  float* fdist = reinterpret_cast<float*>(frontier + a.fr_cap);              // [fr_cap]
  __shared__ uint32_t s_nf, s_cur, s_done, s_epoch, s_changed;
  __shared__ float s_cur_dist;

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, t = tid & 7u, oct = tid >> 3;
  uint8_t* stamp = a.stamps + (size_t)blockIdx.x * a.stamp_stride;

  for (uint32_t qi = blockIdx.x; qi < a.B; qi += gridDim.x) {
    if (a.q_status[qi] != 0u || !ix.populated) {   // uniform per CTA
      if (tid == 0) a.out_counts[qi] = 0;
      continue;
    }
    const float q_hdr = a.q_hdr[qi];
    for (uint32_t i = tid; i < ix.ld; i += HX_HNSW_THREADS)
      sq[i] = i < ix.dim ? a.queries[(size_t)qi * ix.dim + i] : 0.0f;
    if (tid == 0) s_epoch = a.epochs[blockIdx.x] + 1u;
    __syncthreads();
    uint32_t epoch = s_epoch;
    if (epoch >= 256u) {   // stamp wrap: clear this CTA's stamp array (once per 255 queries)
      uint4* s4 = reinterpret_cast<uint4*>(stamp);
      const size_t n16 = a.stamp_stride >> 4;
      for (size_t i = tid; i < n16; i += HX_HNSW_THREADS) s4[i] = make_uint4(0, 0, 0, 0);
      epoch = 1u;
    }
    __syncthreads();
    if (tid == 0) a.epochs[blockIdx.x] = epoch;
    const uint8_t ep8 = (uint8_t)epoch;

    // ---- entry point -------------------------------------------------------------------------------
    uint32_t cur = ix.entry_slot;
    {
      float s = 0.f;
      if (METRIC == HXM_MANHATTAN) {
        if (tid == 0) s = hx_manhattan_seq(ix.vec + (size_t)cur * ix.ld, sq, ix.dim);
      } else if (oct == 0) {
        s = hx_octet_score<METRIC, NB>(ix, sq, q_hdr, cur, t);
      }
      if (tid == 0) {
        if (!hx_score_ok(s)) atomicOr(a.err_flags, HXF_INVALID_SCORE);
        s_cur_dist = s;
      }
    }
    __syncthreads();
    float cur_dist = s_cur_dist;
    uint32_t upper_steps = 0;

    // ---- upper layers: greedy descent (search.rs:169-224) --------------------------------------------
    // The reference's per-layer visited set only avoids re-scoring: a node scored earlier had
    // dist >= current_dist at that time >= current_dist now, so it can never pass the strict `<` again.
    for (int layer = ix.max_layer; layer >= 1; --layer) {
      for (;;) {
        uint32_t deg = 0;
        const uint32_t* row = nullptr;
        {
          const uint32_t off = ix.upper_off[cur];
          if (off != HX_ABSENT && (int)ix.level[cur] >= layer) {
            deg = ix.upper_deg[off + (uint32_t)layer - 1u];
            row = ix.upper_nbr + (size_t)(off + (uint32_t)layer - 1u) * ix.stride_u;
          }
        }
        if (METRIC == HXM_MANHATTAN) {
          for (uint32_t f = tid; f < deg; f += HX_HNSW_THREADS)
            fdist[f] = hx_manhattan_seq(ix.vec + (size_t)row[f] * ix.ld, sq, ix.dim);
        } else {
          for (uint32_t f = oct; f < deg; f += 32) {
            float s = hx_octet_score<METRIC, NB>(ix, sq, q_hdr, row[f], t);
            if (t == 0) fdist[f] = s;
          }
        }
        __syncthreads();
        if (warp == 0) {
          // first neighbour (list order = id order) holding the minimum score, if strictly below current
          float best = cur_dist;
          uint32_t best_i = HX_ABSENT;
          bool bad = false;
          for (uint32_t base = 0; base < deg; base += 32) {
            uint32_t f = base + lane;
            float s = f < deg ? fdist[f] : __int_as_float(0x7f800000);
            if (f < deg && !hx_score_ok(s)) bad = true;
            // sequential semantics: replace when strictly smaller; ties keep the earlier one
            float m = s;
            uint32_t mi = f;
            for (int o = 16; o > 0; o >>= 1) {
              float om = __shfl_xor_sync(0xffffffffu, m, o);
              uint32_t oi = __shfl_xor_sync(0xffffffffu, mi, o);
              if (om < m || (om == m && oi < mi)) { m = om; mi = oi; }
            }
            if (m < best) { best = m; best_i = mi; }
          }
          if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(a.err_flags, HXF_INVALID_SCORE);
          if (lane == 0) {
            if (best_i != HX_ABSENT) { s_cur = row[best_i]; s_cur_dist = best; s_changed = 1u; }
            else s_changed = 0u;
          }
        }
        __syncthreads();
        const uint32_t changed = s_changed;
        if (changed) { cur = s_cur; cur_dist = s_cur_dist; upper_steps++; }
        __syncthreads();
        if (!changed) break;
      }
    }

    // ---- layer 0: beam search ---------------------------------------------------------------------------
    HxBeam beam{beam_mem, 0u};
    uint32_t tie_len = 0, dropped = 0;           // warp-0 state (replicated in its 32 lanes)
    uint32_t st_steps = 0, st_examined = 0, st_dc = 1;
    if (warp == 0) {
      if (lane == 0) {
        beam_mem[0] = hx_make_key(cur_dist, cur << 1);
        stamp[cur] = ep8;
      }
      beam.len = 1;
      __syncwarp();
    }
    for (;;) {
      if (warp == 0) {
        // -- pop the nearest candidate (search.rs:538-551)
        uint32_t first = HX_ABSENT;
        for (uint32_t i = lane; i < beam.len; i += 32)
          if (!(beam_mem[i] & 1ull)) { first = i; break; }
        first = hx_warp_min(first);
        uint32_t cur_slot = HX_ABSENT;
        if (first != HX_ABSENT) {
          uint64_t key = beam_mem[first];
          __syncwarp();
          if (lane == 0) beam_mem[first] = key | 1ull;
          cur_slot = (uint32_t)(key & 0xffffffffu) >> 1;
          st_steps++;
        } else if (tie_len > 0) {
          // evicted entry whose score equals w.max: `current_dist > w.max` is false, the reference expands it
          uint64_t key = tie[tie_len - 1];
          tie_len--;
          cur_slot = (uint32_t)(key & 0xffffffffu) >> 1;
          st_steps++;
        } else if (dropped) {
          st_steps++;   // the reference pops one evicted candidate and breaks on it
        }
        uint32_t nf = 0;
        if (cur_slot != HX_ABSENT) {
          // -- neighbour row, visited filter, mark (search.rs:555-593, 830-831)
          const uint32_t deg = ix.deg0[cur_slot];
          st_examined += ix.raw0[cur_slot];
          const uint32_t* row = ix.nbr0 + (size_t)cur_slot * ix.stride0;
          for (uint32_t base = 0; base < deg; base += 32) {
            const uint32_t i = base + lane;
            uint32_t nb = 0;
            bool fresh = false;
            if (i < deg) {
              nb = row[i];
              fresh = stamp[nb] != ep8;
            }
            const uint32_t mask = __ballot_sync(0xffffffffu, fresh);
            if (fresh) {
              frontier[nf + __popc(mask & ((1u << lane) - 1u))] = nb;
              stamp[nb] = ep8;
            }
            nf += __popc(mask);
          }
          st_dc += nf;
        }
        if (lane == 0) {
          s_nf = nf;
          s_done = (cur_slot == HX_ABSENT) ? 1u : 0u;
        }
      }
      __syncthreads();
      if (s_done) break;
      const uint32_t nf = s_nf;
      // -- score the frontier (search.rs:909-933): one octet per neighbour
      if (METRIC == HXM_MANHATTAN) {
        for (uint32_t f = tid; f < nf; f += HX_HNSW_THREADS)
          fdist[f] = hx_manhattan_seq(ix.vec + (size_t)frontier[f] * ix.ld, sq, ix.dim);
      } else {
        for (uint32_t f = oct; f < nf; f += 32) {
          float s = hx_octet_score<METRIC, NB>(ix, sq, q_hdr, frontier[f], t);
          if (t == 0) fdist[f] = s;
        }
      }
      __syncthreads();
      // -- admit in neighbour-id order (search.rs:934-953)
      if (warp == 0) {
        for (uint32_t base = 0; base < nf; base += 32) {
          const uint32_t f = base + lane;
          float s = f < nf ? fdist[f] : 0.f;
          uint32_t sbits = 0;
          bool pass = false;
          if (f < nf) {
            if (!hx_score_ok(s)) atomicOr(a.err_flags, HXF_INVALID_SCORE);
            sbits = __float_as_uint(s);
            const uint32_t wmax = (uint32_t)(beam_mem[beam.len - 1] >> 32);
            pass = (sbits < wmax) || (beam.len < a.ef);   // w.max only decreases once full: a fail now is final
          }
          uint32_t mask = __ballot_sync(0xffffffffu, pass);
          while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const uint32_t xb = __shfl_sync(0xffffffffu, sbits, src);
            const uint32_t xslot = __shfl_sync(0xffffffffu, f < nf ? frontier[f] : 0u, src);
            const uint32_t wmax = (uint32_t)(beam_mem[beam.len - 1] >> 32);
            if (!((xb < wmax) || (beam.len < a.ef))) continue;
            const uint32_t old_wmax = wmax;
            const bool was_full = beam.len == a.ef;
            uint64_t ev;
            hx_beam_insert(beam, a.ef, ((uint64_t)xb << 32) | ((uint64_t)xslot << 1), &ev, lane);
            if (lane == 0) {   // warm the row we will need if this candidate is expanded
              hx_prefetch_l2(ix.nbr0 + (size_t)xslot * ix.stride0);
              hx_prefetch_l2(ix.deg0 + xslot);
            }
            if (was_full) {
              const uint32_t new_wmax = (uint32_t)(beam_mem[beam.len - 1] >> 32);
              if (new_wmax < old_wmax && tie_len) { dropped = 1; tie_len = 0; }
              if (!(ev & 1ull)) {   // evicted while still unexpanded
                if ((uint32_t)(ev >> 32) == new_wmax) {
                  if (tie_len < HX_TIE_CAP) {
                    if (lane == 0) tie[tie_len] = ev;
                    tie_len++;
                  } else {
                    if (lane == 0) atomicOr(a.err_flags, HXF_TIE_OVERFLOW);
                    dropped = 1;
                  }
                } else {
                  dropped = 1;
                }
              }
              __syncwarp();
            }
          }
        }
      }
      // no barrier needed here: only warp 0 touches the beam, and the next barrier orders frontier reuse
    }

    // ---- results: beam is sorted by (score,id); take k (search.rs:994-1004,1229) -------------------------
    if (warp == 0 && lane == 0) s_nf = beam.len;
    __syncthreads();
    const uint32_t len = s_nf;
    const uint32_t cnt = len < a.k ? len : a.k;
    for (uint32_t i = tid; i < cnt; i += HX_HNSW_THREADS) {
      const uint64_t key = beam_mem[i];
      a.out_ids[(size_t)qi * a.k + i] = ix.ids[(uint32_t)(key & 0xffffffffu) >> 1];
      a.out_scores[(size_t)qi * a.k + i] = hx_key_score(key);
    }
    if (tid == 0) {
      a.out_counts[qi] = cnt;
      if (a.q_stats) {
        a.q_stats[(size_t)qi * 4 + 0] = st_steps;
        a.q_stats[(size_t)qi * 4 + 1] = st_examined;
        a.q_stats[(size_t)qi * 4 + 2] = st_dc;
        a.q_stats[(size_t)qi * 4 + 3] = upper_steps;
      }
    }
    __syncthreads();
  }
}

// ---- warp-per-query variant (throughput) ---------------------------------------------------------------------------
// A layer-0 expansion of a converged beam discovers only a handful of unvisited neighbours, so a 256-thread CTA per
// query leaves most octets idle and — at 4 CTAs per SM — keeps only 4 dependent pointer chases in flight per SM.  Here
// every WARP owns one query (its 4 octets score 4 neighbours per round), 8 warps per CTA, up to 32 queries in flight per
// SM: enough independent row fetches to cover HBM latency.  Same algorithm, same order of every float operation and of
// every admission as k_hnsw_search; __syncthreads became __syncwarp.  `wstride` = shared-memory bytes per warp.
template <int METRIC, int NB, int MINB>
__global__ void __launch_bounds__(HX_HNSW_THREADS, MINB) k_hnsw_search_warp(HxDev ix, HxHnswArgs a, uint32_t wstride) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, t = lane & 7u, oct = lane >> 3;
  const uint32_t warps_per_cta = blockDim.x >> 5;
  const uint32_t gw = blockIdx.x * warps_per_cta + warp;          // global warp id == stamp slot
  const uint32_t total_warps = gridDim.x * warps_per_cta;
  unsigned char* wmem = smem + (size_t)warp * wstride;
  float* sq = reinterpret_cast<float*>(wmem);                                  // [ld]
  uint64_t* beam_mem = reinterpret_cast<uint64_t*>(wmem + (size_t)ix.ld * 4);  // [ef]
  uint64_t* tie = beam_mem + a.ef;                                             // [HX_TIE_CAP]
  uint32_t* frontier = reinterpret_cast<uint32_t*>(tie + HX_TIE_CAP);          // [fr_cap]
  float* fdist = reinterpret_cast<float*>(frontier + a.fr_cap);               // [fr_cap]
  uint8_t* stamp = a.stamps + (size_t)gw * a.stamp_stride;
  const unsigned FULL = 0xffffffffu;

  for (uint32_t qi = gw; qi < a.B; qi += total_warps) {
    if (a.q_status[qi] != 0u || !ix.populated) {
      if (lane == 0) a.out_counts[qi] = 0;
      continue;
    }
    const float q_hdr = a.q_hdr[qi];
    for (uint32_t i = lane; i < ix.ld; i += 32) sq[i] = i < ix.dim ? a.queries[(size_t)qi * ix.dim + i] : 0.0f;
    uint32_t epoch = 0;
    if (lane == 0) epoch = a.epochs[gw] + 1u;
    epoch = __shfl_sync(FULL, epoch, 0);
    if (epoch >= 256u) {
      uint4* s4 = reinterpret_cast<uint4*>(stamp);
      const size_t n16 = a.stamp_stride >> 4;
      for (size_t i = lane; i < n16; i += 32) s4[i] = make_uint4(0, 0, 0, 0);
      epoch = 1u;
    }
    __syncwarp();
    if (lane == 0) a.epochs[gw] = epoch;
    const uint8_t ep8 = (uint8_t)epoch;

    // ---- entry point
    uint32_t cur = ix.entry_slot;
    float cur_dist = 0.f;
    {
      float s = 0.f;
      if (METRIC == HXM_MANHATTAN) {
        if (lane == 0) s = hx_manhattan_seq(ix.vec + (size_t)cur * ix.ld, sq, ix.dim);
      } else if (oct == 0) {
        s = hx_octet_score<METRIC, NB>(ix, sq, q_hdr, cur, t);
      }
      s = __shfl_sync(FULL, s, 0);
      if (!hx_score_ok(s) && lane == 0) atomicOr(a.err_flags, HXF_INVALID_SCORE);
      cur_dist = s;
    }
    uint32_t upper_steps = 0;

    // ---- upper layers: greedy descent (search.rs:169-224)
    for (int layer = ix.max_layer; layer >= 1; --layer) {
      for (;;) {
        uint32_t deg = 0;
        const uint32_t* row = nullptr;
        {
          const uint32_t off = ix.upper_off[cur];
          if (off != HX_ABSENT && (int)ix.level[cur] >= layer) {
            deg = ix.upper_deg[off + (uint32_t)layer - 1u];
            row = ix.upper_nbr + (size_t)(off + (uint32_t)layer - 1u) * ix.stride_u;
          }
        }
        if (METRIC == HXM_MANHATTAN) {
          for (uint32_t f = lane; f < deg; f += 32) fdist[f] = hx_manhattan_seq(ix.vec + (size_t)row[f] * ix.ld, sq, ix.dim);
        } else {
          for (uint32_t f = oct; f < deg; f += 4) {
            float s = hx_octet_score<METRIC, NB>(ix, sq, q_hdr, row[f], t);
            if (t == 0) fdist[f] = s;
          }
        }
        __syncwarp();
        float best = cur_dist;
        uint32_t best_i = HX_ABSENT;
        bool bad = false;
        for (uint32_t base = 0; base < deg; base += 32) {
          uint32_t f = base + lane;
          float s = f < deg ? fdist[f] : __int_as_float(0x7f800000);
          if (f < deg && !hx_score_ok(s)) bad = true;
          float m = s;
          uint32_t mi = f;
          for (int o = 16; o > 0; o >>= 1) {
            float om = __shfl_xor_sync(FULL, m, o);
            uint32_t oi = __shfl_xor_sync(FULL, mi, o);
            if (om < m || (om == m && oi < mi)) { m = om; mi = oi; }
          }
          if (m < best) { best = m; best_i = mi; }
        }
        if (__any_sync(FULL, bad) && lane == 0) atomicOr(a.err_flags, HXF_INVALID_SCORE);
        __syncwarp();
        if (best_i == HX_ABSENT) break;
        cur = row[best_i];
        cur_dist = best;
        upper_steps++;
      }
    }

    // ---- layer 0: beam search
    HxBeam beam{beam_mem, 1u};
    uint32_t tie_len = 0, dropped = 0;
    uint32_t st_steps = 0, st_examined = 0, st_dc = 1;
    if (lane == 0) {
      beam_mem[0] = hx_make_key(cur_dist, cur << 1);
      stamp[cur] = ep8;
    }
    __syncwarp();
    for (;;) {
      uint32_t first = HX_ABSENT;
      for (uint32_t i = lane; i < beam.len; i += 32)
        if (!(beam_mem[i] & 1ull)) { first = i; break; }
      first = hx_warp_min(first);
      uint32_t cur_slot = HX_ABSENT;
      if (first != HX_ABSENT) {
        uint64_t key = beam_mem[first];
        __syncwarp();
        if (lane == 0) beam_mem[first] = key | 1ull;
        cur_slot = (uint32_t)(key & 0xffffffffu) >> 1;
        st_steps++;
      } else if (tie_len > 0) {
        uint64_t key = tie[tie_len - 1];
        tie_len--;
        cur_slot = (uint32_t)(key & 0xffffffffu) >> 1;
        st_steps++;
      } else if (dropped) {
        st_steps++;
      }
      if (cur_slot == HX_ABSENT) break;
      uint32_t nf = 0;
      {
        const uint32_t deg = ix.deg0[cur_slot];
        st_examined += ix.raw0[cur_slot];
        const uint32_t* row = ix.nbr0 + (size_t)cur_slot * ix.stride0;
        for (uint32_t base = 0; base < deg; base += 32) {
          const uint32_t i = base + lane;
          uint32_t nb = 0;
          bool fresh = false;
          if (i < deg) {
            nb = row[i];
            fresh = stamp[nb] != ep8;
          }
          const uint32_t mask = __ballot_sync(FULL, fresh);
          if (fresh) {
            frontier[nf + __popc(mask & ((1u << lane) - 1u))] = nb;
            stamp[nb] = ep8;
          }
          nf += __popc(mask);
        }
        st_dc += nf;
      }
      __syncwarp();
      if (METRIC == HXM_MANHATTAN) {
        for (uint32_t f = lane; f < nf; f += 32) fdist[f] = hx_manhattan_seq(ix.vec + (size_t)frontier[f] * ix.ld, sq, ix.dim);
      } else {
        for (uint32_t f = oct; f < nf; f += 4) {
          float s = hx_octet_score<METRIC, NB>(ix, sq, q_hdr, frontier[f], t);
          if (t == 0) fdist[f] = s;
        }
      }
      __syncwarp();
      for (uint32_t base = 0; base < nf; base += 32) {
        const uint32_t f = base + lane;
        float s = f < nf ? fdist[f] : 0.f;
        uint32_t sbits = 0;
        bool pass = false;
        if (f < nf) {
          if (!hx_score_ok(s)) atomicOr(a.err_flags, HXF_INVALID_SCORE);
          sbits = __float_as_uint(s);
          const uint32_t wmax = (uint32_t)(beam_mem[beam.len - 1] >> 32);
          pass = (sbits < wmax) || (beam.len < a.ef);
        }
        uint32_t mask = __ballot_sync(FULL, pass);
        while (mask) {
          const int src = __ffs(mask) - 1;
          mask &= mask - 1;
          const uint32_t xb = __shfl_sync(FULL, sbits, src);
          const uint32_t xslot = __shfl_sync(FULL, f < nf ? frontier[f] : 0u, src);
          const uint32_t wmax = (uint32_t)(beam_mem[beam.len - 1] >> 32);
          if (!((xb < wmax) || (beam.len < a.ef))) continue;
          const uint32_t old_wmax = wmax;
          const bool was_full = beam.len == a.ef;
          uint64_t ev;
          hx_beam_insert(beam, a.ef, ((uint64_t)xb << 32) | ((uint64_t)xslot << 1), &ev, lane);
          if (lane == 0) {
            hx_prefetch_l2(ix.nbr0 + (size_t)xslot * ix.stride0);
            hx_prefetch_l2(ix.deg0 + xslot);
          }
          if (was_full) {
            const uint32_t new_wmax = (uint32_t)(beam_mem[beam.len - 1] >> 32);
            if (new_wmax < old_wmax && tie_len) { dropped = 1; tie_len = 0; }
            if (!(ev & 1ull)) {
              if ((uint32_t)(ev >> 32) == new_wmax) {
                if (tie_len < HX_TIE_CAP) {
                  if (lane == 0) tie[tie_len] = ev;
                  tie_len++;
                } else {
                  if (lane == 0) atomicOr(a.err_flags, HXF_TIE_OVERFLOW);
                  dropped = 1;
                }
              } else {
                dropped = 1;
              }
            }
            __syncwarp();
          }
        }
      }
      __syncwarp();
    }

    // ---- results
    const uint32_t len = beam.len;
    const uint32_t cnt = len < a.k ? len : a.k;
    for (uint32_t i = lane; i < cnt; i += 32) {
      const uint64_t key = beam_mem[i];
      a.out_ids[(size_t)qi * a.k + i] = ix.ids[(uint32_t)(key & 0xffffffffu) >> 1];
      a.out_scores[(size_t)qi * a.k + i] = hx_key_score(key);
    }
    if (lane == 0) {
      a.out_counts[qi] = cnt;
      if (a.q_stats) {
        a.q_stats[(size_t)qi * 4 + 0] = st_steps;
        a.q_stats[(size_t)qi * 4 + 1] = st_examined;
        a.q_stats[(size_t)qi * 4 + 2] = st_dc;
        a.q_stats[(size_t)qi * 4 + 3] = upper_steps;
      }
    }
    __syncwarp();
  }
}

// ---- warp-per-query with TMA-staged rows (default throughput build) --------------------------------------------------
// ncu on k_hnsw_search_warp (profiles/r01_ncu_k_hnsw_search_warp_octet_details.txt) shows the limit: 58 % of the warp
// cycles are long-scoreboard stalls with 0.2 eligible warps per scheduler — register-fed LDG.128s keep only ~5 loads per
// thread in flight, and a row needs 24 of them, so every scored row costs several dependent DRAM round trips.
// Here the rows of a round are fetched by the TMA engine instead: lane i issues ONE cp.async.bulk (global -> shared,
// 4*ld bytes) for row i, all R rows of the round are in flight at once with zero registers, an mbarrier counts the
// bytes, and the octets then reduce the rows out of shared memory (LDS.128) in exactly the same order as before.
//   shared memory per warp: query | R rows | beam | tie stack | frontier | scores | row headers | mbarrier
#define HX_TMA_MAX_THREADS 512
template <int METRIC>
__global__ void __launch_bounds__(HX_TMA_MAX_THREADS, 1) k_hnsw_search_tma(HxDev ix, HxHnswArgs a, uint32_t wstride, uint32_t R) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, t = lane & 7u, oct = lane >> 3;
  const uint32_t warps_per_cta = blockDim.x >> 5;
  const uint32_t gw = blockIdx.x * warps_per_cta + warp;
  const uint32_t total_warps = gridDim.x * warps_per_cta;
  unsigned char* wmem = smem + (size_t)warp * wstride;
  float* sq = reinterpret_cast<float*>(wmem);                                         // [ld]
  float* rowbuf = sq + ix.ld;                                                         // [R][ld]
  uint64_t* beam_mem = reinterpret_cast<uint64_t*>(rowbuf + (size_t)R * ix.ld);      // [ef]
  uint64_t* tie = beam_mem + a.ef;                                                    // [HX_TIE_CAP]
  uint64_t* bar = tie + HX_TIE_CAP;                                                   // [1]
  uint32_t* frontier = reinterpret_cast<uint32_t*>(bar + 1);                          // [fr_cap]
  float* fdist = reinterpret_cast<float*>(frontier + a.fr_cap);                      // [fr_cap]
  float* fhdr = fdist + a.fr_cap;                                                     // [fr_cap]
  uint8_t* stamp = a.stamps + (size_t)gw * a.stamp_stride;
  const unsigned FULL = 0xffffffffu;
  const uint32_t rowbytes = ix.ld * 4u;
  uint32_t phase = 0;
  if (lane == 0) {
    hx_mbar_init(bar, 1);
    hx_fence_mbar_init();
  }
  __syncwarp();

  // score `cnt` rows whose slots sit in list[0..cnt) (shared or global memory) into fdist[0..cnt)
  auto score_rows = [&](const uint32_t* list, uint32_t cnt, float q_hdr) {
    for (uint32_t base = 0; base < cnt; base += R) {
      const uint32_t rows = min(R, cnt - base);
      if (lane == 0) hx_mbar_expect_tx(bar, rows * rowbytes);
      __syncwarp();
      if (lane < rows) {
        const uint32_t slot = list[base + lane];
        hx_bulk_g2s(rowbuf + (size_t)lane * ix.ld, ix.vec + (size_t)slot * ix.ld, rowbytes, bar);
        if (METRIC == HXM_COSINE) fhdr[lane] = __ldg(ix.hdr + slot);
      }
      hx_mbar_wait(bar, phase);
      phase ^= 1u;
      __syncwarp();
      if (METRIC == HXM_MANHATTAN) {
        if (lane < rows) fdist[base + lane] = hx_octet_score_smem<METRIC>(rowbuf + (size_t)lane * ix.ld, sq, q_hdr, 0.f, ix.dim, 0);
      } else {
        for (uint32_t r = oct; r < rows; r += 4) {
          float s = hx_octet_score_smem<METRIC>(rowbuf + (size_t)r * ix.ld, sq, q_hdr, METRIC == HXM_COSINE ? fhdr[r] : 0.f, ix.dim, t);
          if (t == 0) fdist[base + r] = s;
        }
      }
      __syncwarp();   // every lane is done with rowbuf before the next round overwrites it
    }
  };

  for (uint32_t qi = gw; qi < a.B; qi += total_warps) {
    if (a.q_status[qi] != 0u || !ix.populated) {
      if (lane == 0) a.out_counts[qi] = 0;
      continue;
    }
    const float q_hdr = a.q_hdr[qi];
    for (uint32_t i = lane; i < ix.ld; i += 32) sq[i] = i < ix.dim ? a.queries[(size_t)qi * ix.dim + i] : 0.0f;
    uint32_t epoch = 0;
    if (lane == 0) epoch = a.epochs[gw] + 1u;
    epoch = __shfl_sync(FULL, epoch, 0);
    if (epoch >= 256u) {
      uint4* s4 = reinterpret_cast<uint4*>(stamp);
      const size_t n16 = a.stamp_stride >> 4;
      for (size_t i = lane; i < n16; i += 32) s4[i] = make_uint4(0, 0, 0, 0);
      epoch = 1u;
    }
    __syncwarp();
    if (lane == 0) a.epochs[gw] = epoch;
    const uint8_t ep8 = (uint8_t)epoch;

    // ---- entry point
    uint32_t cur = ix.entry_slot;
    if (lane == 0) frontier[0] = cur;
    __syncwarp();
    score_rows(frontier, 1, q_hdr);
    float cur_dist = fdist[0];
    if (!hx_score_ok(cur_dist) && lane == 0) atomicOr(a.err_flags, HXF_INVALID_SCORE);
    uint32_t upper_steps = 0;
    __syncwarp();

    // ---- upper layers: greedy descent (search.rs:169-224)
    for (int layer = ix.max_layer; layer >= 1; --layer) {
      for (;;) {
        uint32_t deg = 0;
        const uint32_t* row = nullptr;
        {
          const uint32_t off = ix.upper_off[cur];
          if (off != HX_ABSENT && (int)ix.level[cur] >= layer) {
            deg = ix.upper_deg[off + (uint32_t)layer - 1u];
            row = ix.upper_nbr + (size_t)(off + (uint32_t)layer - 1u) * ix.stride_u;
          }
        }
        score_rows(row, deg, q_hdr);
        float best = cur_dist;
        uint32_t best_i = HX_ABSENT;
        bool bad = false;
        for (uint32_t base = 0; base < deg; base += 32) {
          uint32_t f = base + lane;
          float s = f < deg ? fdist[f] : __int_as_float(0x7f800000);
          if (f < deg && !hx_score_ok(s)) bad = true;
          float m = s;
          uint32_t mi = f;
          for (int o = 16; o > 0; o >>= 1) {
            float om = __shfl_xor_sync(FULL, m, o);
            uint32_t oi = __shfl_xor_sync(FULL, mi, o);
            if (om < m || (om == m && oi < mi)) { m = om; mi = oi; }
          }
          if (m < best) { best = m; best_i = mi; }
        }
        if (__any_sync(FULL, bad) && lane == 0) atomicOr(a.err_flags, HXF_INVALID_SCORE);
        __syncwarp();
        if (best_i == HX_ABSENT) break;
        cur = row[best_i];
        cur_dist = best;
        upper_steps++;
      }
    }

    // ---- layer 0: beam search
    HxBeam beam{beam_mem, 1u};
    uint32_t tie_len = 0, dropped = 0;
    uint32_t st_steps = 0, st_examined = 0, st_dc = 1;
    if (lane == 0) {
      beam_mem[0] = hx_make_key(cur_dist, cur << 1);
      stamp[cur] = ep8;
    }
    __syncwarp();
    for (;;) {
      uint32_t first = HX_ABSENT;
      for (uint32_t i = lane; i < beam.len; i += 32)
        if (!(beam_mem[i] & 1ull)) { first = i; break; }
      first = hx_warp_min(first);
      uint32_t cur_slot = HX_ABSENT;
      if (first != HX_ABSENT) {
        uint64_t key = beam_mem[first];
        __syncwarp();
        if (lane == 0) beam_mem[first] = key | 1ull;
        cur_slot = (uint32_t)(key & 0xffffffffu) >> 1;
        st_steps++;
      } else if (tie_len > 0) {
        uint64_t key = tie[tie_len - 1];
        tie_len--;
        cur_slot = (uint32_t)(key & 0xffffffffu) >> 1;
        st_steps++;
      } else if (dropped) {
        st_steps++;
      }
      if (cur_slot == HX_ABSENT) break;
      uint32_t nf = 0;
      {
        const uint32_t deg = ix.deg0[cur_slot];
        st_examined += ix.raw0[cur_slot];
        const uint32_t* row = ix.nbr0 + (size_t)cur_slot * ix.stride0;
        for (uint32_t base = 0; base < deg; base += 32) {
          const uint32_t i = base + lane;
          uint32_t nb = 0;
          bool fresh = false;
          if (i < deg) {
            nb = row[i];
            fresh = stamp[nb] != ep8;
          }
          const uint32_t mask = __ballot_sync(FULL, fresh);
          if (fresh) {
            frontier[nf + __popc(mask & ((1u << lane) - 1u))] = nb;
            stamp[nb] = ep8;
          }
          nf += __popc(mask);
        }
        st_dc += nf;
      }
      __syncwarp();
      score_rows(frontier, nf, q_hdr);
      for (uint32_t base = 0; base < nf; base += 32) {
        const uint32_t f = base + lane;
        float s = f < nf ? fdist[f] : 0.f;
        uint32_t sbits = 0;
        bool pass = false;
        if (f < nf) {
          if (!hx_score_ok(s)) atomicOr(a.err_flags, HXF_INVALID_SCORE);
          sbits = __float_as_uint(s);
          const uint32_t wmax = (uint32_t)(beam_mem[beam.len - 1] >> 32);
          pass = (sbits < wmax) || (beam.len < a.ef);
        }
        uint32_t mask = __ballot_sync(FULL, pass);
        while (mask) {
          const int src = __ffs(mask) - 1;
          mask &= mask - 1;
          const uint32_t xb = __shfl_sync(FULL, sbits, src);
          const uint32_t xslot = __shfl_sync(FULL, f < nf ? frontier[f] : 0u, src);
          const uint32_t wmax = (uint32_t)(beam_mem[beam.len - 1] >> 32);
          if (!((xb < wmax) || (beam.len < a.ef))) continue;
          const uint32_t old_wmax = wmax;
          const bool was_full = beam.len == a.ef;
          uint64_t ev;
          hx_beam_insert(beam, a.ef, ((uint64_t)xb << 32) | ((uint64_t)xslot << 1), &ev, lane);
          if (lane == 0) {
            hx_prefetch_l2(ix.nbr0 + (size_t)xslot * ix.stride0);
            hx_prefetch_l2(ix.deg0 + xslot);
          }
          if (was_full) {
            const uint32_t new_wmax = (uint32_t)(beam_mem[beam.len - 1] >> 32);
            if (new_wmax < old_wmax && tie_len) { dropped = 1; tie_len = 0; }
            if (!(ev & 1ull)) {
              if ((uint32_t)(ev >> 32) == new_wmax) {
                if (tie_len < HX_TIE_CAP) {
                  if (lane == 0) tie[tie_len] = ev;
                  tie_len++;
                } else {
                  if (lane == 0) atomicOr(a.err_flags, HXF_TIE_OVERFLOW);
                  dropped = 1;
                }
              } else {
                dropped = 1;
              }
            }
            __syncwarp();
          }
        }
      }
      __syncwarp();
    }

    // ---- results
    const uint32_t len = beam.len;
    const uint32_t cnt = len < a.k ? len : a.k;
    for (uint32_t i = lane; i < cnt; i += 32) {
      const uint64_t key = beam_mem[i];
      a.out_ids[(size_t)qi * a.k + i] = ix.ids[(uint32_t)(key & 0xffffffffu) >> 1];
      a.out_scores[(size_t)qi * a.k + i] = hx_key_score(key);
    }
    if (lane == 0) {
      a.out_counts[qi] = cnt;
      if (a.q_stats) {
        a.q_stats[(size_t)qi * 4 + 0] = st_steps;
        a.q_stats[(size_t)qi * 4 + 1] = st_examined;
        a.q_stats[(size_t)qi * 4 + 2] = st_dc;
        a.q_stats[(size_t)qi * 4 + 3] = upper_steps;
      }
    }
    __syncwarp();
  }
}

// ---- CTA-per-query with TMA-staged rows (latency build, B < #SMs) ---------------------------------------------------------
// One CTA owns the query; warp 0 walks the beam and, as soon as the unvisited neighbours of an expansion are known, issues
// one cp.async.bulk per neighbour row (up to RC rows, 4*ld bytes each) — the whole frontier is in flight in one DRAM round
// trip instead of the ~4 dependent ones the register-fed loop needs; all 32 octets then reduce one row each out of shared
// memory.  Bit-identical to the other builds.
template <int METRIC>
__global__ void __launch_bounds__(HX_HNSW_THREADS, 1) k_hnsw_search_cta_tma(HxDev ix, HxHnswArgs a, uint32_t RC) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* sq = reinterpret_cast<float*>(smem);                                         // [ld]
  float* rowbuf = sq + ix.ld;                                                         // [RC][ld]
  uint64_t* beam_mem = reinterpret_cast<uint64_t*>(rowbuf + (size_t)RC * ix.ld);     // [ef]
  uint64_t* tie = beam_mem + a.ef;                                                    // [HX_TIE_CAP]
  uint64_t* bar = tie + HX_TIE_CAP;                                                   // [1]
  uint32_t* frontier = reinterpret_cast<uint32_t*>(bar + 1);                          // [fr_cap]
  float* fdist = reinterpret_cast<float*>(frontier + a.fr_cap);                      // [fr_cap]
  float* fhdr = fdist + a.fr_cap;                                                     // [fr_cap]
  __shared__ uint32_t s_nf, s_cur, s_done, s_epoch, s_changed;
  __shared__ float s_cur_dist;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, t = tid & 7u, oct = tid >> 3;
  uint8_t* stamp = a.stamps + (size_t)blockIdx.x * a.stamp_stride;
  const uint32_t rowbytes = ix.ld * 4u;
  uint32_t phase = 0;
  if (tid == 0) {
    hx_mbar_init(bar, 1);
    hx_fence_mbar_init();
  }
  __syncthreads();

  // warp 0 only: start the copies of rows list[base .. base+rows)
  auto issue_rows = [&](const uint32_t* list, uint32_t base, uint32_t rows) {
    if (lane == 0) hx_mbar_expect_tx(bar, rows * rowbytes);
    __syncwarp();
    if (lane < rows) {
      const uint32_t slot = list[base + lane];
      hx_bulk_g2s(rowbuf + (size_t)lane * ix.ld, ix.vec + (size_t)slot * ix.ld, rowbytes, bar);
      if (METRIC == HXM_COSINE) fhdr[lane] = __ldg(ix.hdr + slot);
    }
  };
  // all threads: wait for the pass, reduce it; `first_issued` = warp 0 already issued the first pass
  auto score_rows = [&](const uint32_t* list, uint32_t cnt, float q_hdr, bool first_issued) {
    for (uint32_t base = 0; base < cnt; base += RC) {
      const uint32_t rows = min(RC, cnt - base);
      if (warp == 0 && !(first_issued && base == 0)) issue_rows(list, base, rows);
      hx_mbar_wait(bar, phase);
      phase ^= 1u;
      __syncthreads();   // fhdr visible; everyone past the wait
      if (METRIC == HXM_MANHATTAN) {
        if (tid < rows) fdist[base + tid] = hx_octet_score_smem<METRIC>(rowbuf + (size_t)tid * ix.ld, sq, q_hdr, 0.f, ix.dim, 0);
      } else {
        for (uint32_t r = oct; r < rows; r += HX_HNSW_THREADS / 8) {
          float s = hx_octet_score_smem<METRIC>(rowbuf + (size_t)r * ix.ld, sq, q_hdr, METRIC == HXM_COSINE ? fhdr[r] : 0.f, ix.dim, t);
          if (t == 0) fdist[base + r] = s;
        }
      }
      __syncthreads();
    }
  };

  for (uint32_t qi = blockIdx.x; qi < a.B; qi += gridDim.x) {
    if (a.q_status[qi] != 0u || !ix.populated) {
      if (tid == 0) a.out_counts[qi] = 0;
      continue;
    }
    const float q_hdr = a.q_hdr[qi];
    for (uint32_t i = tid; i < ix.ld; i += HX_HNSW_THREADS) sq[i] = i < ix.dim ? a.queries[(size_t)qi * ix.dim + i] : 0.0f;
    if (tid == 0) s_epoch = a.epochs[blockIdx.x] + 1u;
    __syncthreads();
    uint32_t epoch = s_epoch;
    if (epoch >= 256u) {
      uint4* s4 = reinterpret_cast<uint4*>(stamp);
      const size_t n16 = a.stamp_stride >> 4;
      for (size_t i = tid; i < n16; i += HX_HNSW_THREADS) s4[i] = make_uint4(0, 0, 0, 0);
      epoch = 1u;
    }
    __syncthreads();
    if (tid == 0) a.epochs[blockIdx.x] = epoch;
    const uint8_t ep8 = (uint8_t)epoch;

    // ---- entry point
    uint32_t cur = ix.entry_slot;
    if (tid == 0) frontier[0] = cur;
    __syncthreads();
    score_rows(frontier, 1, q_hdr, false);
    float cur_dist = fdist[0];
    if (tid == 0 && !hx_score_ok(cur_dist)) atomicOr(a.err_flags, HXF_INVALID_SCORE);
    uint32_t upper_steps = 0;
    __syncthreads();

    // ---- upper layers
    for (int layer = ix.max_layer; layer >= 1; --layer) {
      for (;;) {
        uint32_t deg = 0;
        const uint32_t* row = nullptr;
        {
          const uint32_t off = ix.upper_off[cur];
          if (off != HX_ABSENT && (int)ix.level[cur] >= layer) {
            deg = ix.upper_deg[off + (uint32_t)layer - 1u];
            row = ix.upper_nbr + (size_t)(off + (uint32_t)layer - 1u) * ix.stride_u;
          }
        }
        score_rows(row, deg, q_hdr, false);
        if (warp == 0) {
          float best = cur_dist;
          uint32_t best_i = HX_ABSENT;
          bool bad = false;
          for (uint32_t base = 0; base < deg; base += 32) {
            uint32_t f = base + lane;
            float s = f < deg ? fdist[f] : __int_as_float(0x7f800000);
            if (f < deg && !hx_score_ok(s)) bad = true;
            float m = s;
            uint32_t mi = f;
            for (int o = 16; o > 0; o >>= 1) {
              float om = __shfl_xor_sync(0xffffffffu, m, o);
              uint32_t oi = __shfl_xor_sync(0xffffffffu, mi, o);
              if (om < m || (om == m && oi < mi)) { m = om; mi = oi; }
            }
            if (m < best) { best = m; best_i = mi; }
          }
          if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(a.err_flags, HXF_INVALID_SCORE);
          if (lane == 0) {
            if (best_i != HX_ABSENT) { s_cur = row[best_i]; s_cur_dist = best; s_changed = 1u; }
            else s_changed = 0u;
          }
        }
        __syncthreads();
        const uint32_t changed = s_changed;
        if (changed) { cur = s_cur; cur_dist = s_cur_dist; upper_steps++; }
        __syncthreads();
        if (!changed) break;
      }
    }

    // ---- layer 0
    HxBeam beam{beam_mem, 0u};
    uint32_t tie_len = 0, dropped = 0;
    uint32_t st_steps = 0, st_examined = 0, st_dc = 1;
    if (warp == 0) {
      if (lane == 0) {
        beam_mem[0] = hx_make_key(cur_dist, cur << 1);
        stamp[cur] = ep8;
      }
      beam.len = 1;
      __syncwarp();
    }
    for (;;) {
      if (warp == 0) {
        uint32_t first = HX_ABSENT;
        for (uint32_t i = lane; i < beam.len; i += 32)
          if (!(beam_mem[i] & 1ull)) { first = i; break; }
        first = hx_warp_min(first);
        uint32_t cur_slot = HX_ABSENT;
        if (first != HX_ABSENT) {
          uint64_t key = beam_mem[first];
          __syncwarp();
          if (lane == 0) beam_mem[first] = key | 1ull;
          cur_slot = (uint32_t)(key & 0xffffffffu) >> 1;
          st_steps++;
        } else if (tie_len > 0) {
          uint64_t key = tie[tie_len - 1];
          tie_len--;
          cur_slot = (uint32_t)(key & 0xffffffffu) >> 1;
          st_steps++;
        } else if (dropped) {
          st_steps++;
        }
        uint32_t nf = 0;
        if (cur_slot != HX_ABSENT) {
          const uint32_t deg = ix.deg0[cur_slot];
          st_examined += ix.raw0[cur_slot];
          const uint32_t* row = ix.nbr0 + (size_t)cur_slot * ix.stride0;
          for (uint32_t base = 0; base < deg; base += 32) {
            const uint32_t i = base + lane;
            uint32_t nb = 0;
            bool fresh = false;
            if (i < deg) {
              nb = row[i];
              fresh = stamp[nb] != ep8;
            }
            const uint32_t mask = __ballot_sync(0xffffffffu, fresh);
            if (fresh) {
              frontier[nf + __popc(mask & ((1u << lane) - 1u))] = nb;
              stamp[nb] = ep8;
            }
            nf += __popc(mask);
          }
          st_dc += nf;
          __syncwarp();
          if (nf) issue_rows(frontier, 0, min(RC, nf));   // the frontier's rows leave for shared memory right away
        }
        if (lane == 0) {
          s_nf = nf;
          s_done = (cur_slot == HX_ABSENT) ? 1u : 0u;
        }
      }
      __syncthreads();
      if (s_done) break;
      const uint32_t nf = s_nf;
      score_rows(frontier, nf, q_hdr, true);
      if (warp == 0) {
        for (uint32_t base = 0; base < nf; base += 32) {
          const uint32_t f = base + lane;
          float s = f < nf ? fdist[f] : 0.f;
          uint32_t sbits = 0;
          bool pass = false;
          if (f < nf) {
            if (!hx_score_ok(s)) atomicOr(a.err_flags, HXF_INVALID_SCORE);
            sbits = __float_as_uint(s);
            const uint32_t wmax = (uint32_t)(beam_mem[beam.len - 1] >> 32);
            pass = (sbits < wmax) || (beam.len < a.ef);
          }
          uint32_t mask = __ballot_sync(0xffffffffu, pass);
          while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const uint32_t xb = __shfl_sync(0xffffffffu, sbits, src);
            const uint32_t xslot = __shfl_sync(0xffffffffu, f < nf ? frontier[f] : 0u, src);
            const uint32_t wmax = (uint32_t)(beam_mem[beam.len - 1] >> 32);
            if (!((xb < wmax) || (beam.len < a.ef))) continue;
            const uint32_t old_wmax = wmax;
            const bool was_full = beam.len == a.ef;
            uint64_t ev;
            hx_beam_insert(beam, a.ef, ((uint64_t)xb << 32) | ((uint64_t)xslot << 1), &ev, lane);
            if (lane == 0) {
              hx_prefetch_l2(ix.nbr0 + (size_t)xslot * ix.stride0);
              hx_prefetch_l2(ix.deg0 + xslot);
            }
            if (was_full) {
              const uint32_t new_wmax = (uint32_t)(beam_mem[beam.len - 1] >> 32);
              if (new_wmax < old_wmax && tie_len) { dropped = 1; tie_len = 0; }
              if (!(ev & 1ull)) {
                if ((uint32_t)(ev >> 32) == new_wmax) {
                  if (tie_len < HX_TIE_CAP) {
                    if (lane == 0) tie[tie_len] = ev;
                    tie_len++;
                  } else {
                    if (lane == 0) atomicOr(a.err_flags, HXF_TIE_OVERFLOW);
                    dropped = 1;
                  }
                } else {
                  dropped = 1;
                }
              }
              __syncwarp();
            }
          }
        }
      }
    }

    if (warp == 0 && lane == 0) s_nf = beam.len;
    __syncthreads();
    const uint32_t len = s_nf;
    const uint32_t cnt = len < a.k ? len : a.k;
    for (uint32_t i = tid; i < cnt; i += HX_HNSW_THREADS) {
      const uint64_t key = beam_mem[i];
      a.out_ids[(size_t)qi * a.k + i] = ix.ids[(uint32_t)(key & 0xffffffffu) >> 1];
      a.out_scores[(size_t)qi * a.k + i] = hx_key_score(key);
    }
    if (tid == 0) {
      a.out_counts[qi] = cnt;
      if (a.q_stats) {
        a.q_stats[(size_t)qi * 4 + 0] = st_steps;
        a.q_stats[(size_t)qi * 4 + 1] = st_examined;
        a.q_stats[(size_t)qi * 4 + 2] = st_dc;
        a.q_stats[(size_t)qi * 4 + 3] = upper_steps;
      }
    }
    __syncthreads();
  }
}
