// hx_common.cuh — device-side building blocks shared by every kernel of libhelix_b200.
//
// Distance arithmetic reproduces the reference's AVX+FMA kernels bit for bit
// (crates/db/src/search/vector/spaces/simple_avx.rs:128-238): the reference keeps four
// 8-lane accumulators, i.e. 32 independent FMA chains where chain L sees the elements
// i with i mod 32 == L in increasing i; it then adds (s1+s2)+(s3+s4) lane-wise, reduces
// the 8 lanes with hsum256 (hi128+lo128, movehl, shuffle 0x55) and finishes the
// d mod 32 tail with separately rounded scalar mul+add.  Here one *octet* (8 consecutive
// threads of a warp) owns one (query,row) pair; thread t holds chains 4t..4t+3, loads
// one float4 (128-bit, coalesced: the octet reads one full 128-byte line per step) and
// the tree is three xor-shuffles (2,4,1) plus three in-thread adds.  IEEE add/mul are
// commutative, so only the tree shape matters, and it is the same tree.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

#define HX_ABSENT 0xFFFFFFFFu
#define HX_KEY_MAX 0xFFFFFFFFFFFFFFFFull

enum { HXM_EUCLIDEAN = 0, HXM_COSINE = 1, HXM_MANHATTAN = 2 };

// device error flags (OR-ed into a global word by kernels; host maps to hx_status)
enum {
  HXF_INVALID_SCORE = 1u,    // NaN / Inf / negative score => InvariantViolation (model.rs:21-28)
  HXF_TIE_OVERFLOW = 2u,     // more exact score ties at the beam boundary than the tie stack holds
  HXF_BEAM_CAPACITY = 4u
};

// Device image of one index shard (passed by value to kernels).
struct HxDev {
  const float* vec;            // [n][ld] row-major, ld = round_up(dim,32), zero padded, 128-byte aligned rows
  const float* hdr;            // [n] row header: cosine norm (cosine.rs:89-93) or 0.0 bias
  const uint64_t* ids;         // [n] slot -> external id, ascending
  uint32_t* nbr0;              // [n][stride0] layer-0 neighbour slots, ascending (== ascending id)
  uint16_t* deg0;              // [n] neighbours present in nbr0
  uint16_t* raw0;              // [n] full row length incl. neighbours without a vector row (neighbors_examined)
  const uint32_t* upper_off;   // [n] first upper row of the node (layer 1) or HX_ABSENT
  uint32_t* upper_nbr;         // [n_upper_rows][stride_u]
  uint16_t* upper_deg;         // [n_upper_rows]
  const uint8_t* level;        // [n] highest layer holding a row for the node
  uint32_t n, dim, ld, stride0, stride_u;
  int32_t metric;
  uint32_t entry_slot;
  int32_t max_layer;
  int32_t populated;
};

__device__ __forceinline__ float4 hx_ldg4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void hx_prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier wrappers ---------------------------------------------------
__device__ __forceinline__ uint32_t hx_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hx_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hx_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void hx_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void hx_mbar_arrive(uint64_t* bar) {   // release at CTA scope: earlier shared-memory writes are visible to the waiters
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(hx_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void hx_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hx_smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned); completion is counted on `bar`
__device__ __forceinline__ void hx_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   hx_smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(hx_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool hx_mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok)
               : "r"(hx_smem_u32(bar)), "r"(parity)
               : "memory");
  return ok != 0;
}
__device__ __forceinline__ void hx_mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!hx_mbar_try_wait(bar, parity)) {}
}

// ---- octet (8-thread) reduction in hsum256 order ------------------------------------------
// acc.{x,y,z,w} of thread t are chains 4t..4t+3.  With L = 8a + j: a = t>>1, j = 4(t&1)+component.
__device__ __forceinline__ float hx_octet_reduce(float4 acc) {
  // shuffle mask = the caller's octet only, so octets of one warp may diverge independently
  const unsigned full = 0xFFu << (threadIdx.x & 24u);
  float4 o;
  // (s1+s2) and (s3+s4): partner accumulator a^1  <->  t^2
  o.x = __shfl_xor_sync(full, acc.x, 2); o.y = __shfl_xor_sync(full, acc.y, 2);
  o.z = __shfl_xor_sync(full, acc.z, 2); o.w = __shfl_xor_sync(full, acc.w, 2);
  acc.x = __fadd_rn(acc.x, o.x); acc.y = __fadd_rn(acc.y, o.y);
  acc.z = __fadd_rn(acc.z, o.z); acc.w = __fadd_rn(acc.w, o.w);
  // (s1+s2)+(s3+s4): t^4
  o.x = __shfl_xor_sync(full, acc.x, 4); o.y = __shfl_xor_sync(full, acc.y, 4);
  o.z = __shfl_xor_sync(full, acc.z, 4); o.w = __shfl_xor_sync(full, acc.w, 4);
  acc.x = __fadd_rn(acc.x, o.x); acc.y = __fadd_rn(acc.y, o.y);
  acc.z = __fadd_rn(acc.z, o.z); acc.w = __fadd_rn(acc.w, o.w);
  // hsum256: x128[j] = lane[4+j] + lane[j]  <->  t^1
  o.x = __shfl_xor_sync(full, acc.x, 1); o.y = __shfl_xor_sync(full, acc.y, 1);
  o.z = __shfl_xor_sync(full, acc.z, 1); o.w = __shfl_xor_sync(full, acc.w, 1);
  acc.x = __fadd_rn(acc.x, o.x); acc.y = __fadd_rn(acc.y, o.y);
  acc.z = __fadd_rn(acc.z, o.z); acc.w = __fadd_rn(acc.w, o.w);
  // x64 = x128 + movehl(x128): [0]+[2], [1]+[3];  x32 = x64[0] + x64[1]
  float x64_0 = __fadd_rn(acc.x, acc.z);
  float x64_1 = __fadd_rn(acc.y, acc.w);
  return __fadd_rn(x64_0, x64_1);
}

#define HX_L2_STEP(ACC, Q, R)                                   \
  {                                                             \
    float dx = __fsub_rn((Q).x, (R).x), dy = __fsub_rn((Q).y, (R).y); \
    float dz = __fsub_rn((Q).z, (R).z), dw = __fsub_rn((Q).w, (R).w); \
    (ACC).x = __fmaf_rn(dx, dx, (ACC).x); (ACC).y = __fmaf_rn(dy, dy, (ACC).y); \
    (ACC).z = __fmaf_rn(dz, dz, (ACC).z); (ACC).w = __fmaf_rn(dw, dw, (ACC).w); \
  }
#define HX_DOT_STEP(ACC, Q, R)                                  \
  {                                                             \
    (ACC).x = __fmaf_rn((Q).x, (R).x, (ACC).x); (ACC).y = __fmaf_rn((Q).y, (R).y, (ACC).y); \
    (ACC).z = __fmaf_rn((Q).z, (R).z, (ACC).z); (ACC).w = __fmaf_rn((Q).w, (R).w, (ACC).w); \
  }

// Squared L2 (IS_DOT=false) or dot product (IS_DOT=true) of the query `q` (any address space,
// 16-byte aligned, at least round_up(dim,32) floats) with the global row `row`, computed by the
// 8 threads of an octet (t = 0..7).  Every thread returns the same bits.
// dim >= 32: euclid_similarity_avx_fma / dot_similarity_avx_fma order (simple_avx.rs:128-238);
// dim <  32: the scalar loops (simple.rs:204-234), which are exactly the "tail" below with m = 0.
// One batch of NB consecutive 32-element chunks: all NB 128-bit loads are issued before the first FMA, so NB*16 bytes
// per thread are in flight; the FMAs then run in chunk order (the accumulation order is unchanged by NB).
template <bool IS_DOT, int NB, bool ROW_GLOBAL = true>
__device__ __forceinline__ void hx_octet_batch(float4& acc, const float4* __restrict__ r4, const float4* __restrict__ q4,
                                               uint32_t c) {
  float4 r[NB];
#pragma unroll
  for (int i = 0; i < NB; ++i) r[i] = ROW_GLOBAL ? hx_ldg4(r4 + (c + i) * 8) : r4[(c + i) * 8];   // smem rows: plain LDS.128
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const float4 q = q4[(c + i) * 8];
    if (IS_DOT) { HX_DOT_STEP(acc, q, r[i]) } else { HX_L2_STEP(acc, q, r[i]) }
  }
}

template <bool IS_DOT, int NB = 8, bool ROW_GLOBAL = true>
__device__ __forceinline__ float hx_octet_kernel(const float* __restrict__ row, const float* __restrict__ q,
                                                 uint32_t dim, uint32_t t) {
  const uint32_t chunks = dim >> 5;
  const float4* r4 = reinterpret_cast<const float4*>(row) + t;
  const float4* q4 = reinterpret_cast<const float4*>(q) + t;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t c = 0;
  for (; c + NB <= chunks; c += NB) hx_octet_batch<IS_DOT, NB, ROW_GLOBAL>(acc, r4, q4, c);
  if (NB > 16 && c + 16 <= chunks) { hx_octet_batch<IS_DOT, 16, ROW_GLOBAL>(acc, r4, q4, c); c += 16; }
  if (NB > 8 && c + 8 <= chunks) { hx_octet_batch<IS_DOT, 8, ROW_GLOBAL>(acc, r4, q4, c); c += 8; }
  if (NB > 4 && c + 4 <= chunks) { hx_octet_batch<IS_DOT, 4, ROW_GLOBAL>(acc, r4, q4, c); c += 4; }
  if (NB > 2 && c + 2 <= chunks) { hx_octet_batch<IS_DOT, 2, ROW_GLOBAL>(acc, r4, q4, c); c += 2; }
  if (NB > 1 && c < chunks) { hx_octet_batch<IS_DOT, 1, ROW_GLOBAL>(acc, r4, q4, c); c += 1; }
  float result = hx_octet_reduce(acc);   // all-zero accumulators reduce to +0.0 == the scalar loop's 0.0 start
  const uint32_t m = chunks << 5;
  for (uint32_t i = m; i < dim; ++i) {   // `result += d * d` / `result += a * b`: two roundings, never fused
    float a = q[i], b = ROW_GLOBAL ? __ldg(row + i) : row[i];
    if (IS_DOT) {
      result = __fadd_rn(result, __fmul_rn(a, b));
    } else {
      float d = __fsub_rn(a, b);
      result = __fadd_rn(result, __fmul_rn(d, d));
    }
  }
  return result;
}

// Manhattan: strictly sequential scalar sum (simple.rs:186-202). One thread per row.
__device__ __forceinline__ float hx_manhattan_seq(const float* __restrict__ row, const float* __restrict__ q,
                                                  uint32_t dim) {
  float distance = 0.0f;
  const uint32_t v = dim >> 2;
  const float4* r4 = reinterpret_cast<const float4*>(row);
  const float4* q4 = reinterpret_cast<const float4*>(q);
  for (uint32_t i = 0; i < v; ++i) {
    float4 r = __ldg(r4 + i);
    float4 a = q4[i];
    distance = __fadd_rn(distance, fabsf(__fsub_rn(a.x, r.x)));
    distance = __fadd_rn(distance, fabsf(__fsub_rn(a.y, r.y)));
    distance = __fadd_rn(distance, fabsf(__fsub_rn(a.z, r.z)));
    distance = __fadd_rn(distance, fabsf(__fsub_rn(a.w, r.w)));
  }
  for (uint32_t i = v << 2; i < dim; ++i)
    distance = __fadd_rn(distance, fabsf(__fsub_rn(q[i], __ldg(row + i))));
  return distance;
}

// scaled_l2_norm (cosine.rs:12-36), sequential f64 — used for headers and the stable fallback.
__device__ __forceinline__ double hx_scaled_l2_norm(const float* __restrict__ v, uint32_t dim) {
  double scale = 0.0, scaled_sum = 1.0;
  for (uint32_t i = 0; i < dim; ++i) {
    double magnitude = (double)fabsf(v[i]);
    if (magnitude == 0.0) continue;
    if (scale < magnitude) {
      double ratio = __ddiv_rn(scale, magnitude);
      scaled_sum = __dadd_rn(1.0, __dmul_rn(__dmul_rn(scaled_sum, ratio), ratio));
      scale = magnitude;
    } else {
      double ratio = __ddiv_rn(magnitude, scale);
      scaled_sum = __dadd_rn(scaled_sum, __dmul_rn(ratio, ratio));
    }
  }
  if (scale == 0.0) return 0.0;
  return __dmul_rn(scale, __dsqrt_rn(scaled_sum));
}

// Cosine::norm_no_header (cosine.rs:120-122)
__device__ __forceinline__ float hx_cosine_norm(const float* __restrict__ v, uint32_t dim) {
  double n = hx_scaled_l2_norm(v, dim);
  const double mx = (double)FLT_MAX;
  if (n > mx) n = mx;
  return __double2float_rn(n);
}

// Warp-cooperative scaled_l2_norm with the SAME operations in the SAME order (cosine.rs:12-36).  The recurrence is
// sequential only through `scale` (the running maximum) and the order of the additions.  For a block of 32 elements in
// which no magnitude exceeds the current scale — every block but the handful where the running maximum still grows —
// the 32 divisions and squarings are independent (same divisor), so the lanes compute them in parallel and only the 32
// additions are replayed in element order.  A block containing a new maximum is replayed element by element.
// All lanes return the same bits.
__device__ __forceinline__ float hx_cosine_norm_warp(const float* __restrict__ v, uint32_t dim, uint32_t lane) {
  const unsigned FULL = 0xffffffffu;
  double scale = 0.0, scaled_sum = 1.0;
  for (uint32_t base = 0; base < dim; base += 32) {
    const uint32_t i = base + lane;
    const double mag = i < dim ? (double)fabsf(v[i]) : 0.0;
    const uint32_t cnt = min(32u, dim - base);
    if (__any_sync(FULL, scale < mag)) {
      for (uint32_t j = 0; j < cnt; ++j) {
        const double m = __shfl_sync(FULL, mag, j);
        if (m == 0.0) continue;
        if (scale < m) {
          const double ratio = __ddiv_rn(scale, m);
          scaled_sum = __dadd_rn(1.0, __dmul_rn(__dmul_rn(scaled_sum, ratio), ratio));
          scale = m;
        } else {
          const double ratio = __ddiv_rn(m, scale);
          scaled_sum = __dadd_rn(scaled_sum, __dmul_rn(ratio, ratio));
        }
      }
    } else {
      double sq = 0.0;
      const bool nz = mag != 0.0;
      if (nz) {
        const double ratio = __ddiv_rn(mag, scale);
        sq = __dmul_rn(ratio, ratio);
      }
      const uint32_t nzmask = __ballot_sync(FULL, nz);
      for (uint32_t j = 0; j < cnt; ++j) {
        const double x = __shfl_sync(FULL, sq, j);
        if ((nzmask >> j) & 1u) scaled_sum = __dadd_rn(scaled_sum, x);
      }
    }
  }
  double n = scale == 0.0 ? 0.0 : __dmul_rn(scale, __dsqrt_rn(scaled_sum));
  const double mx = (double)FLT_MAX;
  if (n > mx) n = mx;
  return __double2float_rn(n);
}

// stable_half_cosine (cosine.rs:39-59): f64 fallback, sequential.
static __device__ __noinline__ float hx_stable_half_cosine(const float* __restrict__ p, const float* __restrict__ q,
                                                    uint32_t dim) {
  double pn = hx_scaled_l2_norm(p, dim), qn = hx_scaled_l2_norm(q, dim);
  if (pn == 0.0 || qn == 0.0) return __int_as_float(0x7fc00000);
  double dot = 0.0;
  for (uint32_t i = 0; i < dim; ++i) dot = __dadd_rn(dot, __dmul_rn((double)p[i], (double)q[i]));
  double c = __ddiv_rn(dot, __dmul_rn(pn, qn));
  if (c < -1.0) c = -1.0;
  if (c > 1.0) c = 1.0;
  return __double2float_rn(__dmul_rn(__dadd_rn(1.0, -c), 0.5));
}

// Cosine::distance (cosine.rs:96-118) given pq = dot(query,row); qn = query norm, rn = row norm.
__device__ __forceinline__ float hx_cosine_finish(float pq, float qn, float rn, const float* __restrict__ q,
                                                  const float* __restrict__ row, uint32_t dim) {
  float pnqn = __fmul_rn(qn, rn);
  float apn = fabsf(pnqn);
  bool normal = (apn >= FLT_MIN) && (apn <= FLT_MAX);     // f32::is_normal
  bool finite_pq = fabsf(pq) <= FLT_MAX;                    // f32::is_finite
  if (qn > 0.0f && rn > 0.0f && qn != FLT_MAX && rn != FLT_MAX && normal && finite_pq) {
    float c = __fdiv_rn(pq, pnqn);
    if (c < -1.0f) c = -1.0f;
    if (c > 1.0f) c = 1.0f;
    return __fmul_rn(__fsub_rn(1.0f, c), 0.5f);             // (1 - cos) / 2: x/2 == x*0.5 exactly
  }
  return hx_stable_half_cosine(q, row, dim);
}

// Full metric score of one row for one octet. For Manhattan only thread t==0's value is meaningful
// work-wise but all threads compute the same sequential sum (callers normally use the per-lane path).
template <int METRIC, int NB = 8>
__device__ __forceinline__ float hx_octet_score(const HxDev& ix, const float* __restrict__ q, float q_hdr,
                                                uint32_t slot, uint32_t t) {
  const float* row = ix.vec + (size_t)slot * ix.ld;
  if (METRIC == HXM_EUCLIDEAN) return hx_octet_kernel<false, NB>(row, q, ix.dim, t);
  if (METRIC == HXM_COSINE) {
    float pq = hx_octet_kernel<true, NB>(row, q, ix.dim, t);
    return hx_cosine_finish(pq, q_hdr, __ldg(ix.hdr + slot), q, row, ix.dim);
  }
  return hx_manhattan_seq(row, q, ix.dim);
}

// Same score, the row already staged in shared memory (TMA path); `row_hdr` = the row's header.
template <int METRIC, int NB = 8>
__device__ __forceinline__ float hx_octet_score_smem(const float* __restrict__ row_smem, const float* __restrict__ q,
                                                     float q_hdr, float row_hdr, uint32_t dim, uint32_t t) {
  if (METRIC == HXM_EUCLIDEAN) return hx_octet_kernel<false, NB, false>(row_smem, q, dim, t);
  if (METRIC == HXM_COSINE) {
    float pq = hx_octet_kernel<true, NB, false>(row_smem, q, dim, t);
    return hx_cosine_finish(pq, q_hdr, row_hdr, q, row_smem, dim);
  }
  float distance = 0.0f;   // Manhattan: one sequential chain (simple.rs:186-202)
  for (uint32_t i = 0; i < dim; ++i) distance = __fadd_rn(distance, fabsf(__fsub_rn(q[i], row_smem[i])));
  return distance;
}

// DistanceScore::try_new (parameters.rs:241-258): finite, >= 0, -0 -> +0.  Returns false if invalid.
__device__ __forceinline__ bool hx_score_ok(float& s) {
  if (!(s >= 0.0f) || !(s <= FLT_MAX)) return false;
  if (s == 0.0f) s = 0.0f;
  return true;
}

// (score, id) total order as ONE unsigned compare: scores are finite and non-negative, so their bit
// patterns order like the values; slots are ascending in id (model.rs:41-61).
__device__ __forceinline__ uint64_t hx_make_key(float score, uint32_t low) {
  return ((uint64_t)__float_as_uint(score) << 32) | (uint64_t)low;
}
__device__ __forceinline__ float hx_key_score(uint64_t key) { return __uint_as_float((uint32_t)(key >> 32)); }

__device__ __forceinline__ uint32_t hx_warp_sum(uint32_t v) {
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
__device__ __forceinline__ uint32_t hx_warp_min(uint32_t v) {
  v = min(v, __shfl_xor_sync(0xffffffffu, v, 16));
  v = min(v, __shfl_xor_sync(0xffffffffu, v, 8));
  v = min(v, __shfl_xor_sync(0xffffffffu, v, 4));
  v = min(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = min(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return v;
}
