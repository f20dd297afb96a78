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
#include "k_util.cuh"

#define HX_VT_EMPTY 0xFFFFFFFFu
#define HXF_VT_OVERFLOW 8u
#define HXF_COPY_TIMEOUT 16u

struct HxRingArgs {
  uint32_t* vtab;        // [slots][vt_cap] visited hash tables (slot = global warp id / CTA id)
  uint32_t vt_cap;       // power of two
  uint32_t* pool;        // [pool_n][pool_cap] overflow tables
  uint32_t* pool_busy;   // [pool_n]
  uint32_t pool_n, pool_cap;
  uint32_t* counter;     // next query index (zeroed before the launch)
  uint32_t l2_hint;      // 1: evict-first cache hint on row copies
  uint32_t batch_admit;  // 1: one-pass admission of a frontier (latency build, register beam)
  uint32_t prefetch_below;  // warp ring build: L2-prefetch the neighbour row of an admitted entry only below this beam position
  uint32_t l2_spec;      // 1: latency build prefetches the predicted next expansion's rows into L2
  unsigned long long* prof;   // optional [8] cycle sums of the latency build's phases (HX_PHASE_PROF=1, diagnostics only)
  // overflow regions of the tie stack (evicted-unexpanded entries whose score equals w.max): the reference's candidates heap
  // is unbounded, so a corpus with many exactly equal scores (duplicate rows, quantised data) must not fail the search
  uint64_t* tie_pool;         // [tie_pool_n][tie_pool_cap]
  uint32_t* tie_busy;         // [tie_pool_n]
  uint32_t tie_pool_n, tie_pool_cap;
};

// ---- tie stack: HX_TIE_CAP entries in shared memory, then a pool region in global memory ---------------------------------
struct HxTie {
  uint64_t* s;      // shared memory [HX_TIE_CAP]
  uint64_t* ovf;    // claimed overflow region or nullptr
  int ovf_idx;
  uint32_t len;
};
__device__ __forceinline__ HxTie hx_tie_make(uint64_t* s) {
  HxTie t;
  t.s = s; t.ovf = nullptr; t.ovf_idx = -1; t.len = 0;
  return t;
}
// warp-uniform; false = no overflow region free / region full (the caller fails THIS query only)
__device__ __forceinline__ bool hx_tie_push(HxTie& t, uint64_t e, const HxRingArgs& rg, uint32_t lane) {
  if (t.len < HX_TIE_CAP) {
    if (lane == 0) t.s[t.len] = e;
    t.len++;
    return true;
  }
  if (t.ovf_idx < 0) {
    int got = -1;
    if (lane == 0) {
      // BOUNDED wait (about a millisecond): holders release their region at the end of their query, but an unbounded wait
      // could deadlock against the visited-table pool (one query holding a region and waiting for a table, another the
      // reverse).  When the wait runs out only THIS query fails.
      for (uint32_t tries = 0; tries < 512u && got < 0; ++tries) {
        const uint32_t start = (uint32_t)(clock64() >> 4) % (rg.tie_pool_n ? rg.tie_pool_n : 1u);
        for (uint32_t j = 0; j < rg.tie_pool_n; ++j) {
          const uint32_t i = (start + j) % rg.tie_pool_n;
          if (atomicCAS(rg.tie_busy + i, 0u, 1u) == 0u) { got = (int)i; break; }
        }
        if (got < 0) __nanosleep(2000);
      }
    }
    got = __shfl_sync(0xffffffffu, got, 0);
    if (got < 0) return false;
    t.ovf_idx = got;
    t.ovf = rg.tie_pool + (size_t)got * rg.tie_pool_cap;
  }
  const uint32_t o = t.len - HX_TIE_CAP;
  if (o >= rg.tie_pool_cap) return false;
  if (lane == 0) t.ovf[o] = e;
  t.len++;
  __syncwarp();
  return true;
}
__device__ __forceinline__ uint64_t hx_tie_pop(HxTie& t) {
  t.len--;
  return t.len >= HX_TIE_CAP ? ((volatile uint64_t*)t.ovf)[t.len - HX_TIE_CAP] : t.s[t.len];
}
__device__ __forceinline__ void hx_tie_release(HxTie& t, const HxRingArgs& rg, uint32_t lane) {
  if (t.ovf_idx >= 0 && lane == 0) {
    __threadfence();
    atomicExch(rg.tie_busy + t.ovf_idx, 0u);
  }
  t.ovf_idx = -1;
  t.ovf = nullptr;
  t.len = 0;
}

// per-query completion: error flags of THIS query (q_err) and, for the service path, the host-mapped done word.
// Called by one whole warp after it has written the query's results.
__device__ __forceinline__ void hx_query_done(const HxHnswArgs& a, uint32_t qi, uint32_t qflags, uint32_t lane) {
  qflags = __reduce_or_sync(0xffffffffu, qflags);
  if (a.done) __threadfence_system();   // results (possibly in host-mapped memory) before the done word
  __syncwarp();
  if (lane == 0) {
    if (qflags) atomicOr(a.err_flags, qflags);
    if (a.q_err) a.q_err[qi] = qflags;
    if (a.done) *((volatile uint32_t*)a.done + qi) = 0x80000000u | qflags;
  }
}

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

// acquire load at system scope: the writer is the copy engine (a host-to-device copy of the availability counter)
__device__ __forceinline__ uint32_t hx_ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// bulk prefetch of `bytes` (multiple of 16) into L2, no shared-memory destination
__device__ __forceinline__ void hx_bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
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
__device__ __forceinline__ void hx_beam_insert2(HxBeam& b, uint32_t ef, uint64_t key, uint64_t* evicted, uint32_t lane,
                                                uint32_t* pos_out = nullptr) {
  const unsigned FULL = 0xffffffffu;
  uint32_t cnt = 0;
  for (uint32_t i = lane; i < b.len; i += 32) cnt += (b.a[i] < key) ? 1u : 0u;
  const uint32_t pos = __reduce_add_sync(FULL, cnt);
  if (pos_out) *pos_out = pos;
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
#define HX_RING_QCH48_THREADS 320   // 1024 < d <= 1536: 48 query registers per lane; 6 KB rows leave room for at most 9 warps anyway
template <int METRIC, int QCH>
__global__ void __launch_bounds__(QCH >= 48 ? HX_RING_QCH48_THREADS : HX_RING_MAX_THREADS, 1) k_hnsw_search_ring(HxDev ix, HxHnswArgs a, HxRingArgs rg,
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
  constexpr int MS = METRIC == HXM_MANHATTAN ? HXM_EUCLIDEAN : METRIC;   // the warp reduction is never run for Manhattan

  auto issue = [&](uint32_t s, uint32_t slot) {   // one lane
    hx_mbar_expect_tx(bars + s, rowbytes);
    if (rg.l2_hint) hx_bulk_g2s_hint(ring + (size_t)s * ix.ld, ix.vec + (size_t)slot * ix.ld, rowbytes, bars + s, policy);
    else hx_bulk_g2s(ring + (size_t)s * ix.ld, ix.vec + (size_t)slot * ix.ld, rowbytes, bars + s);
  };
  // score the rows whose slots are list[0..cnt) into fdist[0..cnt)
  auto score_list = [&](const uint32_t* list, uint32_t cnt) {
    if (METRIC == HXM_MANHATTAN) {
      // one strictly sequential chain per row (simple.rs:186-202): lane f walks row f, straight from global memory
      for (uint32_t f = lane; f < cnt; f += 32) fdist[f] = hx_manhattan_seq(ix.vec + (size_t)list[f] * ix.ld, sq, ix.dim);
      __syncwarp();
      return;
    }
    if (R == 0) {   // a row does not fit next to the query state (dim in the tens of thousands): reduce it from global memory
      for (uint32_t j = 0; j < cnt; ++j) {
        const uint32_t slot = list[j];
        const float sc = hx_warp_score<MS, QCH>(ix.vec + (size_t)slot * ix.ld, qr, sq, qg, q_hdr,
                                                METRIC == HXM_COSINE ? __ldg(ix.hdr + slot) : 0.f, ix.dim, lane);
        if (lane == 0) fdist[j] = sc;
      }
      __syncwarp();
      return;
    }
    if (lane < min(R, cnt)) issue(lane, list[lane]);
    if (METRIC == HXM_COSINE)
      for (uint32_t f = lane; f < cnt; f += 32) fhdr[f] = __ldg(ix.hdr + list[f]);
    __syncwarp();
    uint32_t s = 0;
    for (uint32_t j = 0; j < cnt; ++j) {
      hx_mbar_wait(bars + s, (ph >> s) & 1u);
      ph ^= 1u << s;
      const float sc = hx_warp_score<MS, QCH>(ring + (size_t)s * ix.ld, qr, sq, qg, q_hdr,
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
    qg = a.queries + (size_t)qi * ix.dim;
    uint32_t qflags = 0;   // error flags raised by THIS query
    if (a.avail) {   // the query may still be on its way from the host
      uint32_t gone = 0;
      if (lane == 0) {
        // a few seconds without progress means the copy stream died: give the launch up instead of hanging the device
        uint32_t spins = 0;
        while (hx_ld_acquire_sys(a.avail) <= qi) {
          __nanosleep(200);
          if ((++spins & 0xfffu) == 0u && ((*(volatile uint32_t*)a.err_flags & HXF_COPY_TIMEOUT) || spins > (1u << 24))) {
            atomicOr(a.err_flags, HXF_COPY_TIMEOUT);
            gone = 1;
            break;
          }
        }
      }
      gone = __shfl_sync(FULL, gone, 0);
      if (gone) {
        if (lane == 0) { a.out_counts[qi] = 0; if (a.q_status_w) a.q_status_w[qi] = 0; }
        hx_query_done(a, qi, HXF_COPY_TIMEOUT, lane);
        continue;
      }
    }
    if (a.fused_validate) {   // ValidatedMetricVector::try_new + D::new_header by the warp that owns the query
      float h;
      const uint32_t st = hx_validate_warp(qg, ix.dim, ix.metric, a.limit, a.has_limit, lane, &h);
      if (lane == 0) { a.q_status_w[qi] = st; a.q_hdr_w[qi] = h; }
      if (st != 0u || !ix.populated) {
        if (lane == 0) a.out_counts[qi] = 0;
        hx_query_done(a, qi, 0u, lane);
        continue;
      }
      q_hdr = h;
    } else {
      if (a.q_status[qi] != 0u || !ix.populated) {
        if (lane == 0) a.out_counts[qi] = 0;
        hx_query_done(a, qi, 0u, lane);
        continue;
      }
      q_hdr = a.q_hdr[qi];
    }
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
    if (!hx_score_ok(cur_dist)) qflags |= HXF_INVALID_SCORE;
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
        if (bad) qflags |= HXF_INVALID_SCORE;
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
    HxTie tq = hx_tie_make(tie);
    uint32_t dropped = 0;
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
      } else if (tq.len > 0) {
        const uint64_t key = hx_tie_pop(tq);
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
            qflags |= HXF_VT_OVERFLOW;
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
          if (!hx_score_ok(s)) qflags |= HXF_INVALID_SCORE;
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
          uint32_t ipos;
          hx_beam_insert2(beam, a.ef, ((uint64_t)xb << 32) | ((uint64_t)xslot << 1), &ev, lane, &ipos);
          // warm the neighbour row only for entries that enter the near half of the beam: those are the ones that get
          // expanded (~100 of ~500 admissions per query); the rest would just be DRAM traffic
          if (lane == 0 && ipos < rg.prefetch_below) {
            hx_prefetch_l2(ix.nbr0 + (size_t)xslot * ix.stride0);
            hx_prefetch_l2(ix.deg0 + xslot);
          }
          if (was_full) {
            const uint32_t new_wmax = (uint32_t)(beam_mem[beam.len - 1] >> 32);
            if (new_wmax < old_wmax && tq.len) { dropped = 1; tq.len = 0; }
            if (!(ev & 1ull)) {
              if ((uint32_t)(ev >> 32) == new_wmax) {
                if (!hx_tie_push(tq, ev, rg, lane)) { qflags |= HXF_TIE_OVERFLOW; failed = true; }
              } else {
                dropped = 1;
              }
            }
            __syncwarp();
          }
        }
      }
      __syncwarp();
      if (failed) break;
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
    hx_tie_release(tq, rg, lane);
    hx_query_done(a, qi, qflags, lane);
    __syncwarp();
  }
}

// ---- sorted beam held in the registers of one warp (latency build) ----------------------------------------------------------
// Position p = 32 r + lane lives in v[r] of that lane; positions >= len hold HX_KEY_MAX, whose bit 0 ("expanded") is set,
// so neither the rank count nor the search for the first unexpanded entry needs a length test.  An insertion is one
// redux for the rank and one shuffle-up per register for the shift — no shared-memory round trips on the admission chain.
template <int NB>
struct HxRegBeam {
  uint64_t v[NB];
  uint32_t len;
};
template <int NB>
__device__ __forceinline__ uint64_t hx_rbeam_get(const HxRegBeam<NB>& b, uint32_t p) {
  uint64_t t = 0;   // mask-select (not `if`): keeps v[] in registers — a compare chain is turned into an indexed local load
#pragma unroll
  for (int r = 0; r < NB; ++r) t |= b.v[r] & (0ull - (uint64_t)((uint32_t)r == (p >> 5)));
  return __shfl_sync(0xffffffffu, t, p & 31u);
}
// Insert `key` (absent, and below the current maximum when the beam is full).  Returns the evicted entry or HX_KEY_MAX.
template <int NB>
__device__ __forceinline__ uint64_t hx_rbeam_insert(HxRegBeam<NB>& b, uint32_t ef, uint64_t key, uint32_t lane) {
  const unsigned FULL = 0xffffffffu;
  uint32_t cnt = 0;
#pragma unroll
  for (int r = 0; r < NB; ++r) cnt += (b.v[r] < key) ? 1u : 0u;
  const uint32_t pos = __reduce_add_sync(FULL, cnt);
  const bool full = b.len == ef;
  uint64_t ev = HX_KEY_MAX;
  if (full) ev = hx_rbeam_get(b, ef - 1u);
#pragma unroll
  for (int r = NB - 1; r >= 0; --r) {
    const uint64_t up = __shfl_up_sync(FULL, b.v[r], 1);
    const uint64_t carry = r > 0 ? __shfl_sync(FULL, b.v[r > 0 ? r - 1 : 0], 31) : 0ull;
    const uint64_t prev = lane == 0 ? carry : up;   // the entry at position p - 1
    const uint32_t p = (uint32_t)r * 32u + lane;
    if (p > pos) b.v[r] = prev;
    else if (p == pos) b.v[r] = key;
    if (p >= ef) b.v[r] = HX_KEY_MAX;               // what slid past the end of a full beam is gone
  }
  if (!full) b.len += 1;
  return ev;
}
template <int NB>
__device__ __forceinline__ uint32_t hx_rbeam_first_unexpanded(const HxRegBeam<NB>& b) {
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    const uint32_t m = __ballot_sync(0xffffffffu, !(b.v[r] & 1ull));
    if (m) return (uint32_t)r * 32u + (uint32_t)__ffs((int)m) - 1u;
  }
  return HX_ABSENT;
}

// Admit up to 32 scored neighbours (lane f holds candidate f, in neighbour-id order) in ONE pass, with exactly the outcome
// of the reference's sequential loop (search.rs:934-953: `dist < w.max || len < ef`, push, evict the maximum):
//  * sequentially, candidate c is admitted iff fewer than ef members of (beam  U  earlier admitted candidates) have a
//    score <= s_c  (w.max is the ef-th smallest member, and the test is on the score alone, strict).  A rejected earlier
//    candidate c' with s_c' <= s_c forces c's rejection too, so "earlier admitted" may be replaced by "earlier":
//        admitted(c)  <=>  #{x in beam : s_x <= s_c} + #{c' before c : s_c' <= s_c} < ef      (induction on id order)
//  * the beam afterwards is the ef smallest keys of beam U admitted; evictions happen in descending key order, so the
//    tie stack (evicted-unexpanded entries whose score equals the final w.max, pushed in eviction order) and `dropped`
//    (any other evicted-unexpanded entry; or older ties once w.max has decreased) follow from the merged ranks.
// `stage` = ef u64 of shared memory.  Returns the mask of admitted lanes.
template <int NB>
__device__ __forceinline__ uint32_t hx_rbeam_admit_batch(HxRegBeam<NB>& b, uint32_t& wmax, HxTie& tq, const HxRingArgs& rg,
                                                         uint32_t& dropped, uint64_t* stage, uint32_t ef, uint32_t sbits,
                                                         uint32_t slot, bool valid, uint32_t lane, uint32_t& qflags) {
  const unsigned FULL = 0xffffffffu;
  const bool pass = valid && ((sbits < wmax) || (b.len < ef));   // necessary: w.max never increases
  const uint32_t pm = __ballot_sync(FULL, pass);
  if (!pm) return 0u;
  // ranks against the beam by binary search over its sorted shared-memory mirror (`stage` holds the beam as of the
  // previous admission; bit 0 of a key never decides an order because slots are distinct): cW = #keys below the score
  // bound, rW = #keys below the candidate's own key.  Both searches walk the same (warp-uniform) halving schedule.
  const uint64_t key = ((uint64_t)sbits << 32) | ((uint64_t)slot << 1);
  const uint64_t bound = ((uint64_t)sbits + 1ull) << 32;   // every key with score <= s is below it
  uint32_t cW = 0, rW = 0;
  {
    uint32_t n = b.len, bb = 0, bk = 0;
    if (n) {
      while (n > 1) {
        const uint32_t half = n >> 1;
        const uint64_t xb = stage[bb + half - 1], xk = stage[bk + half - 1];
        bb = xb < bound ? bb + half : bb;
        bk = xk < key ? bk + half : bk;
        n -= half;
      }
      cW = bb + (stage[bb] < bound ? 1u : 0u);
      rW = bk + (stage[bk] < key ? 1u : 0u);
    }
  }
  uint32_t cC = 0;   // earlier passing candidates with a score <= mine (rejected ones can only have larger scores)
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const uint32_t sj = __shfl_sync(FULL, sbits, j);
    cC += (((pm >> j) & 1u) && j < (int)lane && sj <= sbits) ? 1u : 0u;
  }
  const bool adm = pass && (cW + cC < ef);
  const uint32_t am = __ballot_sync(FULL, adm);
  if (!am) return 0u;
  const uint32_t n_adm = (uint32_t)__popc(am);
  uint32_t sh[NB];
#pragma unroll
  for (int r = 0; r < NB; ++r) sh[r] = 0;
  uint32_t rA = 0;
  for (uint32_t m = am; m; m &= m - 1) {
    const int j = __ffs((int)m) - 1;
    const uint64_t ka = __shfl_sync(FULL, key, j);
#pragma unroll
    for (int r = 0; r < NB; ++r) sh[r] += (ka < b.v[r]) ? 1u : 0u;   // sentinels (HX_KEY_MAX) count too: never used
    if (adm && ka < key) rA++;
  }
  const uint32_t old_len = b.len, total = old_len + n_adm, new_len = total < ef ? total : ef;
  const bool was_full = old_len == ef;
  const uint32_t old_wmax = wmax;
  uint32_t q[NB];
  __syncwarp();   // every lane has finished its binary searches over `stage` before any lane rewrites it (racecheck)
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    const uint32_t p = (uint32_t)r * 32u + lane;
    q[r] = p < old_len ? p + sh[r] : HX_ABSENT;
    if (q[r] < ef) stage[q[r]] = b.v[r];
  }
  const uint32_t qc = adm ? rW + rA : HX_ABSENT;
  if (qc < ef) stage[qc] = key;
  __syncwarp();
  if (new_len == ef) wmax = (uint32_t)(stage[ef - 1u] >> 32);
  if (total > ef) {   // evictions: only a full beam evicts, and it is full now
    if (was_full && wmax < old_wmax && tq.len) { dropped = 1; tq.len = 0; }
    bool ev_tie = false, ev_other = false;
#pragma unroll
    for (int r = 0; r < NB; ++r)
      if (q[r] != HX_ABSENT && q[r] >= ef && !(b.v[r] & 1ull)) {
        if ((uint32_t)(b.v[r] >> 32) == wmax) ev_tie = true; else ev_other = true;
      }
    if (qc != HX_ABSENT && qc >= ef) {
      if (sbits == wmax) ev_tie = true; else ev_other = true;
    }
    if (__any_sync(FULL, ev_other)) dropped = 1;
    if (__any_sync(FULL, ev_tie)) {
      // rare: exact score ties at the beam boundary; push in eviction order (descending merged rank)
      for (uint32_t qq = total; qq-- > ef;) {
        uint64_t mine = HX_KEY_MAX;
#pragma unroll
        for (int r = 0; r < NB; ++r)
          if (q[r] == qq) mine = b.v[r];
        if (qc == qq) mine = key;
        const uint32_t owner = __ballot_sync(FULL, mine != HX_KEY_MAX);
        const uint64_t e = __shfl_sync(FULL, mine, owner ? __ffs((int)owner) - 1 : 0);
        if (!owner || (e & 1ull) || (uint32_t)(e >> 32) != wmax) continue;
        if (!hx_tie_push(tq, e, rg, lane)) { qflags |= HXF_TIE_OVERFLOW; dropped = 1; }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    const uint32_t p = (uint32_t)r * 32u + lane;
    b.v[r] = p < new_len ? stage[p] : HX_KEY_MAX;
  }
  b.len = new_len;
  __syncwarp();
  return am;
}

// visited set in shared memory (the table of the latency build until it outgrows it): shared-space CAS, no generic path
__device__ __forceinline__ bool hx_vt_test_and_set_smem(uint32_t tab_s32, uint32_t mask, uint32_t shift, uint32_t key) {
  uint32_t h = (key * 2654435761u) >> shift;
  for (;;) {
    uint32_t old;
    asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(tab_s32 + h * 4u), "r"(HX_VT_EMPTY), "r"(key) : "memory");
    if (old == HX_VT_EMPTY) return true;
    if (old == key) return false;
    h = (h + 1u) & mask;
  }
}

// ---- CTA-per-query, ring-staged rows (latency build, B < #SMs) -----------------------------------------------------------
// One CTA owns the query.  Warp 0 walks the beam — held in its REGISTERS (NB x 32 entries; NB = 0: shared memory, any
// ef): pops the nearest unexpanded entry, reads its neighbour row (L2: prefetched when the entry was admitted; the row of
// the entry most likely to be popped next is already loaded while the current frontier is being scored), and
// tests-and-sets the visited hash set — in SHARED memory, so the visited filter costs no global round trip.  Then every
// warp w issues the bulk copies of the rows it will reduce (rows w, w+W, ...; one mbarrier per row, the query in
// registers), reduces them as they land, and warp 0 admits the scores in neighbour-id order.  Bit-identical to every
// other build.
#define HX_CTA_RING_MAX_THREADS 384
template <int METRIC, int QCH, int NB>
__global__ void __launch_bounds__(QCH >= 48 ? 256 : HX_CTA_RING_MAX_THREADS, 1)
    k_hnsw_search_cta_ring(HxDev ix, HxHnswArgs a, HxRingArgs rg, uint32_t RC, uint32_t vt_smem_cap) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* sq = reinterpret_cast<float*>(smem);                                            // [ld] when QCH == 0
  float* ring = sq + (QCH == 0 ? ix.ld : 0u);                                            // [RC][ld]
  uint64_t* beam_mem = reinterpret_cast<uint64_t*>(ring + (size_t)RC * ix.ld);          // [ef] beam (NB == 0) / merge stage
  uint64_t* tie = beam_mem + a.ef;                                                       // [HX_TIE_CAP]
  uint64_t* bars = tie + HX_TIE_CAP;                                                     // [RC]
  uint32_t* frontier = reinterpret_cast<uint32_t*>(bars + RC);                           // [fr_cap]
  float* fdist = reinterpret_cast<float*>(frontier + a.fr_cap);                         // [fr_cap]
  uint32_t* vts = reinterpret_cast<uint32_t*>(fdist + a.fr_cap);                        // [vt_smem_cap]
  __shared__ uint32_t s_nf, s_cur, s_done, s_changed;
  __shared__ float s_cur_dist;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, W = blockDim.x >> 5;
  const unsigned FULL = 0xffffffffu;
  const uint32_t rowbytes = ix.ld * 4u;
  const uint64_t policy = hx_policy_evict_first();
  uint32_t ph = 0;   // phase parity of every row slot's mbarrier, identical in all threads
  if (tid < RC) hx_mbar_init(bars + tid, 1);
  hx_fence_mbar_init();
  __syncthreads();

  float qr[QCH > 0 ? QCH : 1];
  const float* qg = nullptr;
  float q_hdr = 0.f;

  // all threads: reduce list[0..cnt) (shared memory, visible to every warp) into fdist.  Warp w owns slots and rows
  // w, w+W, ...: it issues their copies itself (lane j -> its j-th row), so a slot is only ever touched by one warp.
  auto score_list = [&](const uint32_t* list, uint32_t cnt, auto&& after_issue) {
    if (METRIC == HXM_MANHATTAN) {
      // one strictly sequential chain per row (simple.rs:186-202): thread f walks row f, straight from global memory
      for (uint32_t f = tid; f < cnt; f += blockDim.x) fdist[f] = hx_manhattan_seq(ix.vec + (size_t)list[f] * ix.ld, sq, ix.dim);
      if (cnt) after_issue();
      __syncthreads();
      return;
    }
    for (uint32_t base = 0; base < cnt; base += RC) {
      const uint32_t rows = min(RC, cnt - base);
      const uint32_t mine = lane * W + warp;   // the row this lane issues
      float rh = 0.f;
      if (mine < rows) {
        const uint32_t slot = list[base + mine];
        hx_mbar_expect_tx(bars + mine, rowbytes);
        if (rg.l2_hint) hx_bulk_g2s_hint(ring + (size_t)mine * ix.ld, ix.vec + (size_t)slot * ix.ld, rowbytes, bars + mine, policy);
        else hx_bulk_g2s(ring + (size_t)mine * ix.ld, ix.vec + (size_t)slot * ix.ld, rowbytes, bars + mine);
        if (METRIC == HXM_COSINE) rh = __ldg(ix.hdr + slot);
      }
      if (base == 0) after_issue();   // work that may overlap the copies' flight (warp 0: speculative L2 prefetch)
      uint32_t j = 0;
      for (uint32_t r = warp; r < rows; r += W, ++j) {
        const float row_hdr = __shfl_sync(FULL, rh, j);
        hx_mbar_wait(bars + r, (ph >> r) & 1u);
        const float sc = hx_warp_score<(METRIC == HXM_MANHATTAN ? HXM_EUCLIDEAN : METRIC), QCH>(
            ring + (size_t)r * ix.ld, qr, sq, qg, q_hdr, row_hdr, ix.dim, lane);
        if (lane == 0) fdist[base + r] = sc;
      }
      ph ^= rows >= 32u ? FULL : ((1u << rows) - 1u);
      __syncthreads();   // scores visible to warp 0 (and `list` free to be rewritten)
    }
  };

  for (uint32_t qi = blockIdx.x; qi < a.B; qi += gridDim.x) {
    uint32_t qflags = 0;   // error flags raised by THIS query (warp 0's copy is the one reported)
    if (a.q_status[qi] != 0u || !ix.populated) {   // uniform per CTA
      if (tid == 0) a.out_counts[qi] = 0;
      if (warp == 0) hx_query_done(a, qi, 0u, lane);
      continue;
    }
    q_hdr = a.q_hdr[qi];
    qg = a.queries + (size_t)qi * ix.dim;
    if (QCH > 0) {
#pragma unroll
      for (int c = 0; c < (QCH > 0 ? QCH : 1); ++c) qr[c] = (uint32_t)(c * 32) + lane < ix.dim ? qg[c * 32 + lane] : 0.f;
    } else {
      for (uint32_t i = tid; i < ix.ld; i += blockDim.x) sq[i] = i < ix.dim ? qg[i] : 0.0f;
    }
    for (uint32_t i = tid; i < vt_smem_cap; i += blockDim.x) vts[i] = HX_VT_EMPTY;
    HxVisited vt = hx_vt_make(vts, vt_smem_cap);   // warp 0's copy is the live one
    const uint32_t vts_s32 = hx_smem_u32(vts);
    int pool_idx = -1;

    // ---- entry point
    uint32_t cur = ix.entry_slot;
    if (tid == 0) frontier[0] = cur;
    __syncthreads();
    score_list(frontier, 1, [] {});
    float cur_dist = fdist[0];
    if (tid == 0 && !hx_score_ok(cur_dist)) qflags |= HXF_INVALID_SCORE;
    uint32_t upper_steps = 0;
    __syncthreads();

    // ---- upper layers: greedy descent (search.rs:169-224)
    for (int layer = ix.max_layer; layer >= 1; --layer) {
      for (;;) {
        uint32_t deg = 0;
        {
          const uint32_t off = ix.upper_off[cur];
          if (off != HX_ABSENT && (int)ix.level[cur] >= layer) {
            deg = ix.upper_deg[off + (uint32_t)layer - 1u];
            const uint32_t* row = ix.upper_nbr + (size_t)(off + (uint32_t)layer - 1u) * ix.stride_u;
            for (uint32_t f = tid; f < deg; f += blockDim.x) frontier[f] = row[f];
          }
        }
        __syncthreads();
        score_list(frontier, deg, [] {});
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
              float om = __shfl_xor_sync(FULL, m, o);
              uint32_t oi = __shfl_xor_sync(FULL, mi, o);
              if (om < m || (om == m && oi < mi)) { m = om; mi = oi; }
            }
            if (m < best) { best = m; best_i = mi; }
          }
          if (bad) qflags |= HXF_INVALID_SCORE;
          if (lane == 0) {
            if (best_i != HX_ABSENT) { s_cur = frontier[best_i]; s_cur_dist = best; s_changed = 1u; }
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

    // ---- layer 0 (beam state lives in warp 0)
    HxBeam beam{beam_mem, 0u};                 // NB == 0
    HxRegBeam<(NB > 0 ? NB : 1)> rb;           // NB > 0
#pragma unroll
    for (int r = 0; r < (NB > 0 ? NB : 1); ++r) rb.v[r] = HX_KEY_MAX;
    rb.len = 0;
    uint32_t wmax = 0xffffffffu;               // score bits of the last entry once the beam is full
    HxTie tq = hx_tie_make(tie);
    uint32_t dropped = 0;
    uint32_t st_steps = 0, st_examined = 0, st_dc = 1;
    bool failed = false;
    // speculative neighbour row: the row of the entry predicted to be popped next, loaded during the scoring phase
    uint32_t sp_slot = HX_ABSENT, sp_nb = 0, sp_deg = 0, sp_raw = 0;
    if (warp == 0) {
      const uint64_t key0 = hx_make_key(cur_dist, cur << 1);
      if (NB > 0) {
        hx_rbeam_insert(rb, a.ef, key0, lane);
        if (rb.len == a.ef) wmax = (uint32_t)(hx_rbeam_get(rb, a.ef - 1u) >> 32);
        if (lane == 0) beam_mem[0] = key0;   // sorted mirror of the beam for the one-pass admission's binary searches
      } else {
        if (lane == 0) beam_mem[0] = key0;
        beam.len = 1;
      }
      if (lane == 0) hx_vt_test_and_set_smem(vts_s32, vt.mask, vt.shift, cur);
      __syncwarp();
    }
    long long pt0 = 0, pt1 = 0, pt2 = 0, pt3 = 0, pt4 = 0, pt5 = 0;
    unsigned long long pa[6] = {0, 0, 0, 0, 0, 0};
    const bool prof = rg.prof != nullptr && tid == 0;
    for (;;) {
      if (warp == 0) {
        if (prof) pt0 = clock64();
        // -- pop the nearest candidate (search.rs:538-551)
        uint32_t cur_slot = HX_ABSENT;
        uint32_t first;
        if (NB > 0) first = hx_rbeam_first_unexpanded(rb);
        else first = hx_beam_first_unexpanded(beam_mem, beam.len, lane);
        if (first != HX_ABSENT) {
          uint64_t key;
          if (NB > 0) {
            key = hx_rbeam_get(rb, first);
#pragma unroll
            for (int r = 0; r < (NB > 0 ? NB : 1); ++r)
              if ((uint32_t)r * 32u + lane == first) rb.v[r] |= 1ull;
          } else {
            key = beam_mem[first];
            __syncwarp();
            if (lane == 0) beam_mem[first] = key | 1ull;
          }
          cur_slot = (uint32_t)(key & 0xffffffffu) >> 1;
          st_steps++;
        } else if (tq.len > 0) {
          const uint64_t key = hx_tie_pop(tq);
          cur_slot = (uint32_t)(key & 0xffffffffu) >> 1;
          st_steps++;
        } else if (dropped) {
          st_steps++;
        }
        if (failed) cur_slot = HX_ABSENT;   // a tie-stack overflow during the previous admission ends this query
        uint32_t nf = 0;
        if (cur_slot != HX_ABSENT) {
          const uint32_t* row = ix.nbr0 + (size_t)cur_slot * ix.stride0;
          if (prof) pt1 = clock64();
          uint32_t nb, deg, raw;
          if (cur_slot == sp_slot) {   // predicted: the row is already in registers
            if (prof) pa[5]++;
            nb = sp_nb; deg = sp_deg; raw = sp_raw;
          } else {
            nb = row[lane];            // stride0 >= 32: in bounds; issued together with the degree
            deg = ix.deg0[cur_slot];
            raw = ix.raw0[cur_slot];
          }
          st_examined += raw;
          if (st_dc + deg > vt.limit) {
            if (pool_idx >= 0 || (pool_idx = hx_vt_grow_warp(vt, rg, lane)) < 0) {
              qflags |= HXF_VT_OVERFLOW;
              failed = true;
              cur_slot = HX_ABSENT;
            }
          }
          if (prof) { pt2 = clock64(); pa[0] += pt1 - pt0; pa[1] += pt2 - pt1; }
          if (!failed) {
            for (uint32_t base = 0; base < deg; base += 32) {
              const uint32_t i = base + lane;
              if (base) nb = i < deg ? row[i] : 0u;
              bool fresh = false;
              if (i < deg) fresh = pool_idx < 0 ? hx_vt_test_and_set_smem(vts_s32, vt.mask, vt.shift, nb) : hx_vt_test_and_set(vt, nb);
              const uint32_t mask = __ballot_sync(FULL, fresh);
              if (fresh) frontier[nf + __popc(mask & ((1u << lane) - 1u))] = nb;
              nf += __popc(mask);
            }
            st_dc += nf;
            if (prof) { pt3 = clock64(); pa[2] += pt3 - pt2; }
          }
        }
        if (lane == 0) {
          s_nf = nf;
          s_done = (cur_slot == HX_ABSENT) ? 1u : 0u;
        }
      }
      __syncthreads();
      if (s_done) break;
      const uint32_t nf = s_nf;
      if (warp == 0) {
        // predict the next pop: the first unexpanded entry as the beam stands now (right unless a new score beats it)
        uint32_t pf;
        if (NB > 0) pf = hx_rbeam_first_unexpanded(rb);
        else pf = hx_beam_first_unexpanded(beam_mem, beam.len, lane);
        sp_slot = HX_ABSENT;
        if (pf != HX_ABSENT) {
          const uint64_t pk = NB > 0 ? hx_rbeam_get(rb, pf) : beam_mem[pf];
          sp_slot = (uint32_t)(pk & 0xffffffffu) >> 1;
          sp_nb = ix.nbr0[(size_t)sp_slot * ix.stride0 + lane];
          sp_deg = ix.deg0[sp_slot];
          sp_raw = ix.raw0[sp_slot];
        }
      }
      score_list(frontier, nf, [&] {
        // Rows of the predicted next expansion: start pulling its unvisited neighbours' vectors into L2 now, while this
        // frontier is in flight / being reduced.  Read-only probe of the visited set; a wrong guess only costs bandwidth.
        if (warp == 0 && rg.l2_spec && sp_slot != HX_ABSENT && lane < sp_deg) {
          bool vis;
          if (pool_idx < 0) {
            uint32_t h = (sp_nb * 2654435761u) >> vt.shift;
            for (;;) {
              const uint32_t cur_e = ((volatile uint32_t*)vts)[h];
              if (cur_e == HX_VT_EMPTY) { vis = false; break; }
              if (cur_e == sp_nb) { vis = true; break; }
              h = (h + 1u) & vt.mask;
            }
          } else {
            uint32_t h = (sp_nb * 2654435761u) >> vt.shift;
            for (;;) {
              const uint32_t cur_e = ((volatile uint32_t*)vt.tab)[h];
              if (cur_e == HX_VT_EMPTY) { vis = false; break; }
              if (cur_e == sp_nb) { vis = true; break; }
              h = (h + 1u) & vt.mask;
            }
          }
          if (!vis) hx_bulk_prefetch_l2(ix.vec + (size_t)sp_nb * ix.ld, rowbytes);
        }
      });
      if (prof) { pt4 = clock64(); pa[3] += pt4 - pt3; }
      // -- admit in neighbour-id order (search.rs:934-953)
      if (warp == 0) {
        for (uint32_t base = 0; base < nf; base += 32) {
          const uint32_t f = base + lane;
          float s = f < nf ? fdist[f] : 0.f;
          uint32_t sbits = 0;
          bool pass = false;
          const uint32_t blen = NB > 0 ? rb.len : beam.len;
          if (NB == 0 && blen) wmax = (uint32_t)(beam_mem[blen - 1] >> 32);
          if (f < nf) {
            if (!hx_score_ok(s)) qflags |= HXF_INVALID_SCORE;
            sbits = __float_as_uint(s);
            pass = (sbits < wmax) || (blen < a.ef);   // w.max only decreases once full: a fail now is final
          }
          uint32_t mask = __ballot_sync(FULL, pass);
          const uint32_t myslot = f < nf ? frontier[f] : 0u;
          if (NB > 0 && rg.batch_admit) {
            const uint32_t am = hx_rbeam_admit_batch(rb, wmax, tq, rg, dropped, beam_mem, a.ef, sbits, myslot, f < nf,
                                                     lane, qflags);
            if (qflags & HXF_TIE_OVERFLOW) failed = true;
            if ((am >> lane) & 1u) {   // warm the rows we will need if these candidates are expanded
              hx_prefetch_l2(ix.nbr0 + (size_t)myslot * ix.stride0);
              hx_prefetch_l2(ix.deg0 + myslot);
            }
            continue;
          }
          while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const uint32_t xb = __shfl_sync(FULL, sbits, src);
            const uint32_t xslot = __shfl_sync(FULL, myslot, src);
            const uint32_t len_now = NB > 0 ? rb.len : beam.len;
            if (NB == 0) wmax = (uint32_t)(beam_mem[len_now - 1] >> 32);
            if (!((xb < wmax) || (len_now < a.ef))) continue;
            const uint32_t old_wmax = wmax;
            const bool was_full = len_now == a.ef;
            uint64_t ev;
            const uint64_t nkey = ((uint64_t)xb << 32) | ((uint64_t)xslot << 1);
            if (NB > 0) {
              ev = hx_rbeam_insert(rb, a.ef, nkey, lane);
              if (rb.len == a.ef) wmax = (uint32_t)(hx_rbeam_get(rb, a.ef - 1u) >> 32);
            } else {
              hx_beam_insert2(beam, a.ef, nkey, &ev, lane);
            }
            if (lane == 0) {   // warm the row we will need if this candidate is expanded
              hx_prefetch_l2(ix.nbr0 + (size_t)xslot * ix.stride0);
              hx_prefetch_l2(ix.deg0 + xslot);
            }
            if (was_full) {
              const uint32_t new_wmax = NB > 0 ? wmax : (uint32_t)(beam_mem[beam.len - 1] >> 32);
              if (new_wmax < old_wmax && tq.len) { dropped = 1; tq.len = 0; }
              if (!(ev & 1ull)) {   // evicted while still unexpanded
                if ((uint32_t)(ev >> 32) == new_wmax) {
                  if (!hx_tie_push(tq, ev, rg, lane)) { qflags |= HXF_TIE_OVERFLOW; failed = true; }
                } else {
                  dropped = 1;
                }
              }
              __syncwarp();
            }
          }
        }
      }
      if (prof) { pt5 = clock64(); pa[4] += pt5 - pt4; }
      if (warp == 0) __syncwarp();   // admission's reads of frontier[] before the next expansion rewrites it (racecheck)
      // no CTA barrier here: only warp 0 touches the beam; the next barrier orders the reuse of frontier / fdist
    }
    if (prof) {
      for (int i = 0; i < 6; ++i) atomicAdd(rg.prof + i, pa[i]);
      atomicAdd(rg.prof + 6, (unsigned long long)st_steps);
    }

    // ---- results: the beam is sorted by (score,id); take k (search.rs:994-1004,1229)
    if (warp == 0) {
      const uint32_t len = failed ? 0u : (NB > 0 ? rb.len : beam.len);
      const uint32_t cnt = len < a.k ? len : a.k;
      if (NB > 0) {
#pragma unroll
        for (int r = 0; r < (NB > 0 ? NB : 1); ++r) {
          const uint32_t p = (uint32_t)r * 32u + lane;
          if (p < cnt) {
            a.out_ids[(size_t)qi * a.k + p] = ix.ids[(uint32_t)(rb.v[r] & 0xffffffffu) >> 1];
            a.out_scores[(size_t)qi * a.k + p] = hx_key_score(rb.v[r]);
          }
        }
      } else {
        for (uint32_t i = lane; i < cnt; i += 32) {
          const uint64_t key = beam_mem[i];
          a.out_ids[(size_t)qi * a.k + i] = ix.ids[(uint32_t)(key & 0xffffffffu) >> 1];
          a.out_scores[(size_t)qi * a.k + i] = hx_key_score(key);
        }
      }
      if (lane == 0) {
        a.out_counts[qi] = cnt;
        if (a.q_stats) {
          a.q_stats[(size_t)qi * 4 + 0] = st_steps;
          a.q_stats[(size_t)qi * 4 + 1] = st_examined;
          a.q_stats[(size_t)qi * 4 + 2] = st_dc;
          a.q_stats[(size_t)qi * 4 + 3] = upper_steps;
        }
        if (pool_idx >= 0) {
          __threadfence();
          atomicExch(rg.pool_busy + pool_idx, 0u);
        }
      }
      hx_tie_release(tq, rg, lane);
      hx_query_done(a, qi, qflags, lane);
    }
    __syncthreads();
  }
}
