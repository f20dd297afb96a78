// k_util.cuh — small supporting kernels: query/row validation + headers, synthetic data, top-k merge.
#pragma once
#include "hx_common.cuh"

// status word per query/row: code << 24 | component index (index clipped to 24 bits)
#define HX_ST_OK 0u
#define HX_ST_COMPONENT 3u   // non-finite component            (domain.rs:127-131)
#define HX_ST_ZERO_NORM 4u   // cosine zero vector              (domain.rs:132-134)
#define HX_ST_MAGNITUDE 5u   // |component| > metric limit      (domain.rs:135-149)

// ValidatedMetricVector::try_new (domain.rs:113-154) for each of `count` vectors of `dim` floats with
// row stride `ld`, plus D::new_header (cosine norm, cosine.rs:89-93). One warp per vector.
// Check order: finiteness (first bad index) -> cosine zero norm -> magnitude (first bad index).
// One warp validates one vector: returns the status word, *hdr_out = the cosine norm (0 for the other metrics / on error).
__device__ __forceinline__ uint32_t hx_validate_warp(const float* __restrict__ x, uint32_t dim, int metric, float limit,
                                                     int has_limit, uint32_t lane, float* hdr_out) {
  uint32_t bad_fin = HX_ABSENT, bad_mag = HX_ABSENT, nonzero = 0;
  for (uint32_t i = lane; i < dim; i += 32) {
    const float c = x[i];
    const float ac = fabsf(c);
    if (!(ac <= FLT_MAX) && bad_fin == HX_ABSENT) bad_fin = i;
    if (has_limit && ac > limit && bad_mag == HX_ABSENT) bad_mag = i;
    if (!(c == 0.0f)) nonzero = 1;
  }
  bad_fin = hx_warp_min(bad_fin);
  bad_mag = hx_warp_min(bad_mag);
  nonzero = __any_sync(0xffffffffu, nonzero) ? 1u : 0u;
  uint32_t st = HX_ST_OK;
  if (bad_fin != HX_ABSENT) st = (HX_ST_COMPONENT << 24) | (bad_fin & 0xffffffu);
  else if (metric == HXM_COSINE && !nonzero) st = (HX_ST_ZERO_NORM << 24);
  else if (bad_mag != HX_ABSENT) st = (HX_ST_MAGNITUDE << 24) | (bad_mag & 0xffffffu);
  float h = 0.0f;
  if (metric == HXM_COSINE && st == HX_ST_OK) h = hx_cosine_norm_warp(x, dim, lane);   // st is warp-uniform
  *hdr_out = h;
  return st;
}

static __global__ void k_validate_and_header(const float* __restrict__ v, size_t count, uint32_t dim, size_t ld, int metric,
                                      float limit, int has_limit, float* __restrict__ hdr_out,
                                      uint32_t* __restrict__ status_out) {
  const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (w >= count) return;
  float h;
  const uint32_t st = hx_validate_warp(v + w * ld, dim, metric, limit, has_limit, lane, &h);
  if (lane == 0) {
    status_out[w] = st;
    hdr_out[w] = h;
  }
}

// ---- synthetic data (bench only): unit-normalised Gaussian mixture, SURVEY §8(d) ---------------------------
__device__ __forceinline__ uint64_t hx_mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}
__device__ __forceinline__ float hx_gauss(uint64_t key) {
  const uint64_t h = hx_mix64(key);
  const float u1 = ((float)((h >> 40) & 0xffffffu) + 1.0f) * (1.0f / 16777217.0f);   // (0,1)
  const float u2 = (float)((h >> 8) & 0xffffffu) * (1.0f / 16777216.0f);            // [0,1)
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}
// row i = normalise(centroid[c(i)] + sigma * N(0,I)); one warp per row; `stream_tag` separates corpus / queries
static __global__ void k_generate_mixture(float* __restrict__ out, size_t count, uint32_t dim, size_t ld, uint64_t seed,
                                   uint32_t n_centroids, float sigma, uint64_t first_index, uint64_t stream_tag) {
  const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (w >= count) return;
  const uint64_t gi = first_index + w;
  const uint64_t c = hx_mix64(seed ^ hx_mix64(gi * 0x9e3779b97f4a7c15ull + stream_tag)) % n_centroids;
  float* row = out + w * ld;
  float ss = 0.f;
  for (uint32_t j = lane; j < dim; j += 32) {
    const float cen = hx_gauss(seed * 0x100000001b3ull + (c << 20) + j + 0x5555000000000000ull);
    const float noise = hx_gauss(hx_mix64(seed + stream_tag) ^ (gi * 0xd1342543de82ef95ull + j));
    const float x = cen + sigma * noise;
    row[j] = x;
    ss += x * x;
  }
  ss += __shfl_xor_sync(0xffffffffu, ss, 16); ss += __shfl_xor_sync(0xffffffffu, ss, 8);
  ss += __shfl_xor_sync(0xffffffffu, ss, 4);  ss += __shfl_xor_sync(0xffffffffu, ss, 2);
  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
  const float inv = ss > 0.f ? rsqrtf(ss) : 0.f;
  for (uint32_t j = lane; j < dim; j += 32) row[j] *= inv;
  for (uint32_t j = dim + lane; j < ld; j += 32) row[j] = 0.f;
}

// Embedding-like corpus: a Gaussian mixture in an r-dimensional latent space (overlapping clusters), pushed through a
// fixed random projection A[dim][r] plus a little isotropic noise, then unit-normalised.  Unlike isolated isotropic
// clusters in 768-d (intrinsic dimension 768, where every graph index degrades), this has intrinsic dimension ~r,
// like real sentence/document embeddings.  One warp per row; requires r <= 64.
static __global__ void k_generate_proj(float* __restrict__ A, uint32_t dim, uint32_t r, uint64_t seed) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < dim * r) A[i] = hx_gauss(seed * 0x9e3779b97f4a7c15ull + 0x7777000000000000ull + i) * rsqrtf((float)r);
}
static __global__ void k_generate_latent(float* __restrict__ out, size_t count, uint32_t dim, size_t ld, uint64_t seed,
                                  uint32_t n_centroids, float sigma, uint64_t first_index, uint64_t stream_tag,
                                  const float* __restrict__ A, uint32_t r, float eps) {
  const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (w >= count) return;
  const uint64_t gi = first_index + w;
  const uint64_t c = hx_mix64(seed ^ hx_mix64(gi * 0x9e3779b97f4a7c15ull + stream_tag)) % n_centroids;
  // latent coordinates: lane holds z[lane] and z[lane+32]
  float z0 = 0.f, z1 = 0.f;
  if (lane < r)
    z0 = hx_gauss(seed * 0x100000001b3ull + (c << 20) + lane + 0x5555000000000000ull) +
         sigma * hx_gauss(hx_mix64(seed + stream_tag) ^ (gi * 0xd1342543de82ef95ull + lane));
  if (lane + 32 < r)
    z1 = hx_gauss(seed * 0x100000001b3ull + (c << 20) + lane + 32 + 0x5555000000000000ull) +
         sigma * hx_gauss(hx_mix64(seed + stream_tag) ^ (gi * 0xd1342543de82ef95ull + lane + 32));
  float* row = out + w * ld;
  float ss = 0.f;
  for (uint32_t j = lane; j < ((dim + 31) / 32) * 32; j += 32) {
    float x = 0.f;
    const float* a = A + (size_t)(j < dim ? j : 0) * r;
    for (uint32_t k = 0; k < r; ++k) {
      const float zk = __shfl_sync(0xffffffffu, k < 32 ? z0 : z1, k & 31);
      x += a[k] * zk;
    }
    if (j < dim) {
      x += eps * hx_gauss(hx_mix64(seed + stream_tag + 0x33) ^ (gi * 0x2545f4914f6cdd1dull + j));
      row[j] = x;
      ss += x * x;
    }
  }
  ss += __shfl_xor_sync(0xffffffffu, ss, 16); ss += __shfl_xor_sync(0xffffffffu, ss, 8);
  ss += __shfl_xor_sync(0xffffffffu, ss, 4);  ss += __shfl_xor_sync(0xffffffffu, ss, 2);
  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
  const float inv = ss > 0.f ? rsqrtf(ss) : 0.f;
  for (uint32_t j = lane; j < dim; j += 32) row[j] *= inv;
  for (uint32_t j = dim + lane; j < ld; j += 32) row[j] = 0.f;
}

static __global__ void k_iota_ids(uint64_t* ids, size_t n, uint64_t first) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ids[i] = first + i;
}

// ---- merge of per-shard top-k lists: one warp per query, lane s walks shard s's sorted list -------
// Per-shard lists live at base + s * stride (BYTE strides: the gathered blocks of the sharded path are one block per
// shard, not one array per field); every list holds k_in entries per query, the merge emits k_out.
static __global__ void k_merge_topk(const unsigned char* __restrict__ all_ids, const unsigned char* __restrict__ all_scores,
                             const unsigned char* __restrict__ all_counts, size_t stride_ids, size_t stride_scores,
                             size_t stride_counts, uint32_t n_shards, size_t B, uint32_t k_in, uint32_t k_out,
                             uint64_t* __restrict__ out_ids, float* __restrict__ out_scores,
                             uint32_t* __restrict__ out_counts) {
  const size_t q = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (q >= B) return;
  // shards beyond 32 are folded: lane handles shards lane, lane+32, ... by always exposing its best head
  uint32_t out_n = 0;
  // per-lane cursor for up to 4 shards per lane (n_shards <= 128)
  uint32_t cur[4] = {0, 0, 0, 0};
  uint32_t cnt[4] = {0, 0, 0, 0};
  for (int j = 0; j < 4; ++j) {
    const uint32_t s = lane + 32u * (uint32_t)j;
    if (s < n_shards) {
      const uint32_t c = reinterpret_cast<const uint32_t*>(all_counts + (size_t)s * stride_counts)[q];
      cnt[j] = c < k_in ? c : k_in;
    }
  }
  while (out_n < k_out) {
    // best head among this lane's shards by (score, id)
    uint32_t bs = 0x7f800000u + 1u;   // larger than any finite score's bits
    uint64_t bid = ~0ull;
    int bj = -1;
    for (int j = 0; j < 4; ++j) {
      const uint32_t s = lane + 32u * (uint32_t)j;
      if (s < n_shards && cur[j] < cnt[j]) {
        const size_t o = q * k_in + cur[j];
        const uint32_t sb = __float_as_uint(reinterpret_cast<const float*>(all_scores + (size_t)s * stride_scores)[o]);
        const uint64_t id = reinterpret_cast<const uint64_t*>(all_ids + (size_t)s * stride_ids)[o];
        if (sb < bs || (sb == bs && id < bid)) { bs = sb; bid = id; bj = j; }
      }
    }
    uint32_t ms = bs;
    uint64_t mid = bid;
    uint32_t ml = lane;
    for (int o = 16; o > 0; o >>= 1) {
      const uint32_t os = __shfl_xor_sync(0xffffffffu, ms, o);
      const uint64_t oid = __shfl_xor_sync(0xffffffffu, mid, o);
      const uint32_t ol = __shfl_xor_sync(0xffffffffu, ml, o);
      if (os < ms || (os == ms && (oid < mid || (oid == mid && ol < ml)))) { ms = os; mid = oid; ml = ol; }
    }
    if (ms > 0x7f800000u) break;   // every list exhausted
    if (lane == ml && bj >= 0) cur[bj]++;
    if (lane == 0) {
      out_ids[q * k_out + out_n] = mid;
      out_scores[q * k_out + out_n] = __uint_as_float(ms);
    }
    out_n++;
  }
  if (lane == 0) out_counts[q] = out_n;
}
