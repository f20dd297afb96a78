// k_hnsw_ring.cuh — second-generation HNSW traversal kernels (default for Euclidean / cosine).
//
// Same algorithm, same float operations in the same order and the same admission order as k_hnsw.cuh
// (search.rs:169-224 greedy descent, search.rs:267-1067 strict-exhaustive layer 0) — results, scores and
// SearchStats counters are bit-identical.  What changed is how the bytes move:
//
//  * visited set: an open-addressing hash set of slot numbers per resident query (8192 entries for ef = 100),
//    tested-and-set with ONE atomicCAS per neighbour.  All tables together are a few tens of MB and live in the
//    126 MB L2, where the byte-per-node stamps of the first generation cost n bytes per resident query (2.4 GB at
//    1 M nodes, 24 GB at 10 M) and one DRAM sector read + one write-back per neighbour examined.  A table that
//    fills past 13/16 is re-hashed into a 16x larger one from a small pool, so the set is always exact.
//  * rows: a ring of R shared-memory slots per warp, one mbarrier per slot.  Row j+R is requested the moment row j
//    has been reduced, so R rows stay in flight for the whole frontier instead of R-row rounds that each pay a full
//    DRAM round trip.  Bulk copies carry an L2 evict-first hint: vector rows are streamed once, while neighbour rows
//    and the visited tables are what should stay cached.
//  * one row is reduced by the whole warp: lane L owns FMA chain L (elements i = L mod 32, increasing i) — exactly
//    the 32 chains of the reference's 4 x 8-lane AVX accumulators — the query sits in registers, and the tree is
//    xor 8, 16 ((s1+s2)+(s3+s4)), then 4, 2, 1 (hsum256).  Same tree as the octet kernel, so same bits.
//  * queries are handed out through an atomic counter (no static striding: the last round of a batch no longer
//    leaves most warps idle).
#pragma once
#include "k_hnsw.cuh"

#define HX_VT_EMPTY 0xFFFFFFFFu
#define HXF_VT_OVERFLOW 8u

struct HxRingArgs {
  uint32_t* vtab;        // [slots][vt_cap] visited hash tables (slot = global warp id / CTA id)
  uint32_t vt_cap;       // power of two
  uint32_t* pool;        // [pool_n][pool_cap] overflow tables
  uint32_t* pool_busy;   // [pool_n]
  uint32_t pool_n, pool_cap;
  uint32_t* counter;     // next query index (zeroed before the launch)
  uint32_t l2_hint;      // 1: evict-first cache hint on row copies
};

__device__ __forceinline__ uint64_t hx_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void hx_bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          hx_smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(hx_smem_u32(bar)), "l"(policy)
      : "memory");
}

// ---- visited hash set --------------------------------------------------------------------------------------------------
struct HxVisited {
  uint32_t* tab;
  uint32_t mask, shift, limit;
};
__device__ __forceinline__ HxVisited hx_vt_make(uint32_t* tab, uint32_t cap) {
  HxVisited v;
  v.tab = tab;
  v.mask = cap - 1u;
  v.shift = 32u - (uint32_t)__ffs((int)cap) + 1u;   // 32 - log2(cap)
  v.limit = cap - (cap >> 2) + (cap >> 4);          // 13/16
  return v;
}
// true when `key` was not in the set (it is now)
__device__ __forceinline__ bool hx_vt_test_and_set(const HxVisited& v, uint32_t key) {
  uint32_t h = (key * 2654435761u) >> v.shift;
  for (;;) {
    const uint32_t old = atomicCAS(v.tab + h, HX_VT_EMPTY, key);
    if (old == HX_VT_EMPTY) return true;
    if (old == key) return false;
    h = (h + 1u) & v.mask;
  }
}
__device__ __forceinline__ void hx_vt_clear_warp(uint32_t* tab, uint32_t cap, uint32_t lane) {
  uint4* t4 = reinterpret_cast<uint4*>(tab);
  const uint4 e = make_uint4(HX_VT_EMPTY, HX_VT_EMPTY, HX_VT_EMPTY, HX_VT_EMPTY);
  for (uint32_t i = lane; i < (cap >> 2); i += 32) t4[i] = e;
  __syncwarp();
}
// Move the set into a pool table (warp-cooperative).  Returns the pool index, or -1 when the pool is exhausted.
__device__ __forceinline__ int hx_vt_grow_warp(HxVisited& v, const HxRingArgs& r, uint32_t lane) {
  int got = -1;
  if (lane == 0) {
    // A holder never waits for anything, so waiting for a free table cannot deadlock; only an empty pool fails.
    while (got < 0 && r.pool_n) {
      for (uint32_t i = 0; i < r.pool_n; ++i)
        if (atomicCAS(r.pool_busy + i, 0u, 1u) == 0u) { got = (int)i; break; }
      if (got < 0) __nanosleep(2000);
    }
  }
  got = __shfl_sync(0xffffffffu, got, 0);
  if (got < 0) return -1;
  uint32_t* nt = r.pool + (size_t)got * r.pool_cap;
  hx_vt_clear_warp(nt, r.pool_cap, lane);
  HxVisited nv = hx_vt_make(nt, r.pool_cap);
  const uint32_t old_cap = v.mask + 1u;
  for (uint32_t i = lane; i < old_cap; i += 32) {
    const uint32_t key = v.tab[i];
    if (key != HX_VT_EMPTY) hx_vt_test_and_set(nv, key);
  }
  __syncwarp();
  v = nv;
  return got;
}

// ---- sorted beam, faster insertion -----------------------------------------------------------------------------------
// Same contract as hx_beam_insert; position by one redux, the shift staged through registers 128 entries per pass.
__device__ __forceinline__ void hx_beam_insert2(HxBeam& b, uint32_t ef, uint64_t key, uint64_t* evicted, uint32_t lane) {
  const unsigned FULL = 0xffffffffu;
  uint32_t cnt = 0;
  for (uint32_t i = lane; i < b.len; i += 32) cnt += (b.a[i] < key) ? 1u : 0u;
  const uint32_t pos = __reduce_add_sync(FULL, cnt);
  uint32_t end;   // entries [pos, end) move up by one
  if (b.len == ef) {
    *evicted = b.a[b.len - 1];
    end = b.len - 1;
  } else {
    *evicted = HX_KEY_MAX;
    end = b.len;
    b.len += 1;
  }
  __syncwarp();
  for (int hi = (int)end; hi > (int)pos; hi -= 128) {
    uint64_t v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = hi - 1 - (int)lane - 32 * u;
      if (i >= (int)pos) v[u] = b.a[i];
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = hi - 1 - (int)lane - 32 * u;
      if (i >= (int)pos) b.a[i + 1] = v[u];
    }
    __syncwarp();
  }
  if (lane == 0) b.a[pos] = key;
  __syncwarp();
}

// index of the first beam entry whose `expanded` bit is clear, HX_ABSENT when there is none
__device__ __forceinline__ uint32_t hx_beam_first_unexpanded(const uint64_t* beam, uint32_t len, uint32_t lane) {
  uint32_t first = HX_ABSENT;
  for (uint32_t i = lane; i < len; i += 32)
    if (!(beam[i] & 1ull)) { first = i; break; }
  return __reduce_min_sync(0xffffffffu, first);
}

// ---- one row reduced by one warp -----------------------------------------------------------------------------------------
// Squared L2 / dot product of the query with the shared-memory row, lane L = chain L (simple_avx.rs:128-238).
// qr[c] = query[32c + lane] for c < QCH (QCH >= chunks); QCH == 0: the query is read from shared memory `sq`.
// `qg` = the query in global memory (tail elements and the f64 fallback only).
template <bool IS_DOT, int QCH>
__device__ __forceinline__ float hx_warp_row(const float* __restrict__ row_s, const float* qr, const float* __restrict__ sq,
                                             const float* __restrict__ qg, uint32_t dim, uint32_t lane) {
  const unsigned FULL = 0xffffffffu;
  const uint32_t chunks = dim >> 5;
  float acc = 0.f;
  if (QCH > 0) {
#pragma unroll
    for (int c = 0; c < (QCH > 0 ? QCH : 1); ++c) {
      if ((uint32_t)c < chunks) {
        const float x = row_s[c * 32 + lane];
        const float q = qr[c];
        if (IS_DOT) {
          acc = __fmaf_rn(q, x, acc);
        } else {
          const float d = __fsub_rn(q, x);
          acc = __fmaf_rn(d, d, acc);
        }
      }
    }
  } else {
#pragma unroll 8
    for (uint32_t c = 0; c < chunks; ++c) {
      const float x = row_s[c * 32 + lane];
      const float q = sq[c * 32 + lane];
      if (IS_DOT) {
        acc = __fmaf_rn(q, x, acc);
      } else {
        const float d = __fsub_rn(q, x);
        acc = __fmaf_rn(d, d, acc);
      }
    }
  }
  // L = 8a + j: (s1+s2), (s3+s4) <-> a^1; their sum <-> a^2; hsum256: j^4, j^2, j^1
  acc = __fadd_rn(acc, __shfl_xor_sync(FULL, acc, 8));
  acc = __fadd_rn(acc, __shfl_xor_sync(FULL, acc, 16));
  acc = __fadd_rn(acc, __shfl_xor_sync(FULL, acc, 4));
  acc = __fadd_rn(acc, __shfl_xor_sync(FULL, acc, 2));
  acc = __fadd_rn(acc, __shfl_xor_sync(FULL, acc, 1));
  float result = acc;
  for (uint32_t i = chunks << 5; i < dim; ++i) {   // scalar tail: separately rounded mul + add
    const float a = __ldg(qg + i), b = row_s[i];
    if (IS_DOT) {
      result = __fadd_rn(result, __fmul_rn(a, b));
    } else {
      const float d = __fsub_rn(a, b);
      result = __fadd_rn(result, __fmul_rn(d, d));
    }
  }
  return result;
}

template <int METRIC, int QCH>
__device__ __forceinline__ float hx_warp_score(const float* __restrict__ row_s, const float* qr, const float* __restrict__ sq,
                                               const float* __restrict__ qg, float q_hdr, float row_hdr, uint32_t dim,
                                               uint32_t lane) {
  if (METRIC == HXM_EUCLIDEAN) return hx_warp_row<false, QCH>(row_s, qr, sq, qg, dim, lane);
  const float pq = hx_warp_row<true, QCH>(row_s, qr, sq, qg, dim, lane);
  return hx_cosine_finish(pq, q_hdr, row_hdr, qg, row_s, dim);
}

// ---- warp-per-query, ring-staged rows (throughput build) -------------------------------------------------------------------
// shared memory per warp: [query (QCH == 0 only)] | R row slots | beam | tie stack | R mbarriers | frontier | scores | headers
#define HX_RING_MAX_THREADS 512
template <int METRIC, int QCH>
__global__ void __launch_bounds__(HX_RING_MAX_THREADS, 1) k_hnsw_search_ring(HxDev ix, HxHnswArgs a, HxRingArgs rg,
                                                                            uint32_t wstride, uint32_t R) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t warps_per_cta = blockDim.x >> 5;
  const uint32_t gw = blockIdx.x * warps_per_cta + warp;
  unsigned char* wmem = smem + (size_t)warp * wstride;
  float* sq = reinterpret_cast<float*>(wmem);                                            // [ld] when QCH == 0
  float* ring = sq + (QCH == 0 ? ix.ld : 0u);                                            // [R][ld]
  uint64_t* beam_mem = reinterpret_cast<uint64_t*>(ring + (size_t)R * ix.ld);           // [ef]
  uint64_t* tie = beam_mem + a.ef;                                                       // [HX_TIE_CAP]
  uint64_t* bars = tie + HX_TIE_CAP;                                                     // [R]
  uint32_t* frontier = reinterpret_cast<uint32_t*>(bars + R);                            // [fr_cap]
  float* fdist = reinterpret_cast<float*>(frontier + a.fr_cap);                         // [fr_cap]
  float* fhdr = fdist + a.fr_cap;                                                        // [fr_cap]
  const unsigned FULL = 0xffffffffu;
  const uint32_t rowbytes = ix.ld * 4u;
  const uint64_t policy = hx_policy_evict_first();
  uint32_t ph = 0;   // phase parity of every slot's mbarrier (bit s), warp-uniform
  if (lane < R) hx_mbar_init(bars + lane, 1);
  hx_fence_mbar_init();
  __syncwarp();

  float qr[QCH > 0 ? QCH : 1];
  const float* qg = nullptr;
  float q_hdr = 0.f;

  auto issue = [&](uint32_t s, uint32_t slot) {   // one lane
    hx_mbar_expect_tx(bars + s, rowbytes);
    if (rg.l2_hint) hx_bulk_g2s_hint(ring + (size_t)s * ix.ld, ix.vec + (size_t)slot * ix.ld, rowbytes, bars + s, policy);
    else hx_bulk_g2s(ring + (size_t)s * ix.ld, ix.vec + (size_t)slot * ix.ld, rowbytes, bars + s);
  };
  // score the rows whose slots are list[0..cnt) into fdist[0..cnt)
  auto score_list = [&](const uint32_t* list, uint32_t cnt) {
    if (lane < min(R, cnt)) issue(lane, list[lane]);
    if (METRIC == HXM_COSINE)
      for (uint32_t f = lane; f < cnt; f += 32) fhdr[f] = __ldg(ix.hdr + list[f]);
    __syncwarp();
    uint32_t s = 0;
    for (uint32_t j = 0; j < cnt; ++j) {
      hx_mbar_wait(bars + s, (ph >> s) & 1u);
      ph ^= 1u << s;
      const float sc = hx_warp_score<METRIC, QCH>(ring + (size_t)s * ix.ld, qr, sq, qg, q_hdr,
                                                  METRIC == HXM_COSINE ? fhdr[j] : 0.f, ix.dim, lane);
      if (lane == 0) fdist[j] = sc;
      __syncwarp();   // every lane is done with slot s
      if (j + R < cnt && lane == 0) issue(s, list[j + R]);
      s = (s + 1 == R) ? 0u : s + 1;
    }
    __syncwarp();
  };

  for (;;) {
    uint32_t qi = 0;
    if (lane == 0) qi = atomicAdd(rg.counter, 1u);
    qi = __shfl_sync(FULL, qi, 0);
    if (qi >= a.B) break;
    if (a.q_status[qi] != 0u || !ix.populated) {
      if (lane == 0) a.out_counts[qi] = 0;
      continue;
    }
    q_hdr = a.q_hdr[qi];
    qg = a.queries + (size_t)qi * ix.dim;
    if (QCH > 0) {
#pragma unroll
      for (int c = 0; c < (QCH > 0 ? QCH : 1); ++c) qr[c] = (uint32_t)(c * 32) + lane < ix.dim ? qg[c * 32 + lane] : 0.f;
    } else {
      for (uint32_t i = lane; i < ix.ld; i += 32) sq[i] = i < ix.dim ? qg[i] : 0.0f;
    }
    HxVisited vt = hx_vt_make(rg.vtab + (size_t)gw * rg.vt_cap, rg.vt_cap);
    int pool_idx = -1;
    bool failed = false;
    hx_vt_clear_warp(vt.tab, rg.vt_cap, lane);

    // ---- entry point
    uint32_t cur = ix.entry_slot;
    if (lane == 0) frontier[0] = cur;
    __syncwarp();
    score_list(frontier, 1);
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
        for (uint32_t f = lane; f < deg; f += 32) frontier[f] = row[f];
        __syncwarp();
        score_list(frontier, deg);
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
        cur = frontier[best_i];
        cur_dist = best;
        upper_steps++;
        __syncwarp();
      }
    }

    // ---- layer 0: beam search
    HxBeam beam{beam_mem, 1u};
    uint32_t tie_len = 0, dropped = 0;
    uint32_t st_steps = 0, st_examined = 0, st_dc = 1;
    if (lane == 0) {
      beam_mem[0] = hx_make_key(cur_dist, cur << 1);
      hx_vt_test_and_set(vt, cur);
    }
    __syncwarp();
    for (;;) {
      const uint32_t first = hx_beam_first_unexpanded(beam_mem, beam.len, lane);
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
        const uint32_t* row = ix.nbr0 + (size_t)cur_slot * ix.stride0;
        uint32_t nb = row[lane];                    // stride0 >= 32: in bounds; issued together with the degree
        const uint32_t deg = ix.deg0[cur_slot];
        st_examined += ix.raw0[cur_slot];
        if (st_dc + deg > vt.limit) {               // keep the visited set below 13/16 full
          if (pool_idx >= 0 || (pool_idx = hx_vt_grow_warp(vt, rg, lane)) < 0) {
            if (lane == 0) atomicOr(a.err_flags, HXF_VT_OVERFLOW);
            failed = true;
            break;
          }
        }
        for (uint32_t base = 0; base < deg; base += 32) {
          const uint32_t i = base + lane;
          if (base) nb = i < deg ? row[i] : 0u;
          bool fresh = false;
          if (i < deg) fresh = hx_vt_test_and_set(vt, nb);
          const uint32_t mask = __ballot_sync(FULL, fresh);
          if (fresh) frontier[nf + __popc(mask & ((1u << lane) - 1u))] = nb;
          nf += __popc(mask);
        }
        st_dc += nf;
      }
      __syncwarp();
      score_list(frontier, nf);
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
          hx_beam_insert2(beam, a.ef, ((uint64_t)xb << 32) | ((uint64_t)xslot << 1), &ev, lane);
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
    const uint32_t len = failed ? 0u : beam.len;
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
    if (pool_idx >= 0 && lane == 0) {
      __threadfence();
      atomicExch(rg.pool_busy + pool_idx, 0u);
    }
    __syncwarp();
  }
}
