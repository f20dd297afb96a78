// k_build.cu — HNSW construction on the device.
//
// Restates the reference's insertion path as *batched concurrent insertion*:
//   insert_hnsw                 search/vector/mutation.rs:787-895
//   search_layer_beam           search/vector/mutation.rs:904-1005   (beam = max(ef_construction, degree limit))
//   select_neighbors_heuristic  search/vector/mutation.rs:1072-1098  (items only for the first 2*M candidates)
//   select_diverse              search/vector/mod.rs:809-856          (reject c if some selected s has d(c,s) < d(c,q); back-fill)
//   add_bidirectional_link      search/vector/mutation.rs:1498-1591  (append, re-rank all neighbours, select_diverse, drop the
//                                                                     reciprocal edge of every rejected neighbour)
//
// The reference inserts one node at a time.  Here nodes are inserted in id order in rounds whose size grows with the
// graph (never more than half of what is already linked, capped).  Within a round every new node searches the graph as
// it stood at the start of the round (kernel A), the proposed links are grouped per target row (B0), every target whose
// row would overflow is pruned with the reference's diversity rule (B1), and an edge survives only if BOTH endpoints
// keep it (B2/B3) — which is exactly the invariant the reference's reciprocal cleanup maintains: every edge is
// bidirectional, rows are strictly ascending, no self links, degree <= limit (index.rs:3617-3701).
// The graph differs from a sequential build (nodes of one round do not see each other) so id parity with the oracle's
// build is not claimed; structural invariants and recall are tested instead.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "hx_index.hpp"
#include "k_hnsw.cuh"

#define HXB_THREADS 256
#define HXB_CAP_IN 64          // incoming proposals kept per target row and round
#define HXB_MAX_CAND 192       // degree limit (<=128) + HXB_CAP_IN

struct HxBuildArgs {
  // graph being built (same arrays as HxDev, mutable)
  const uint32_t* upper_row_node;    // [rows]
  const uint8_t* upper_row_layer;    // [rows] 1-based
  uint32_t lim0, limu, efc;
  uint32_t batch_start, batch_size;
  uint32_t cur_entry;                // entry slot of the frozen graph
  int32_t cur_max_layer;             // -1 when the graph is empty
  // proposals
  uint32_t* P;                       // [batch][pstride]: layer 0 at 0, layer l>=1 at lim0 + (l-1)*limu
  uint16_t* Pn;                      // [batch][maxL+1]
  uint32_t pstride, maxL;
  // visited stamps (kernel A)
  uint8_t* stamps;
  uint32_t* epochs;
  size_t stamp_stride;
  // grouping / pruning state, unified row space: row = slot (layer 0) or n + upper row
  uint32_t* cnt;                     // [n+rows] proposals received this round
  uint32_t* inc;                     // [n+rows][HXB_CAP_IN]
  uint32_t* tidx;                    // [n+rows] index in targets[] or HX_ABSENT
  uint32_t* targets;                 // [max_targets] unified rows
  uint32_t* R;                       // [max_targets][rstride] retained neighbours chosen by the target
  uint16_t* Rn;                      // [max_targets]
  uint32_t rstride, max_targets;
  uint32_t* victims;                 // [n+rows]
  uint32_t* vflag;                   // [n+rows]
  uint32_t* counters;                // [0]=n_targets [1]=n_victims [2]=dropped proposals [3]=err flags
};

__device__ __forceinline__ uint32_t hxb_row_id(const HxDev& ix, int layer, uint32_t node) {
  return layer == 0 ? node : ix.n + ix.upper_off[node] + (uint32_t)layer - 1u;
}
__device__ __forceinline__ uint32_t* hxb_row_nbr(const HxDev& ix, uint32_t row) {
  return row < ix.n ? ix.nbr0 + (size_t)row * ix.stride0 : ix.upper_nbr + (size_t)(row - ix.n) * ix.stride_u;
}
__device__ __forceinline__ uint16_t* hxb_row_deg(const HxDev& ix, uint32_t row) {
  return row < ix.n ? ix.deg0 + row : ix.upper_deg + (row - ix.n);
}

// distance between two stored rows, octet-cooperative (all 8 threads return the same bits)
__device__ __forceinline__ float hxb_pair(const HxDev& ix, uint32_t a, uint32_t b, uint32_t t) {
  const float* ra = ix.vec + (size_t)a * ix.ld;
  switch (ix.metric) {
    case HXM_EUCLIDEAN: return hx_octet_kernel<false>(ix.vec + (size_t)b * ix.ld, ra, ix.dim, t);
    case HXM_COSINE: {
      const float* rb = ix.vec + (size_t)b * ix.ld;
      float pq = hx_octet_kernel<true>(rb, ra, ix.dim, t);
      return hx_cosine_finish(pq, ix.hdr[a], ix.hdr[b], ra, rb, ix.dim);
    }
    default: return hx_manhattan_seq(ix.vec + (size_t)b * ix.ld, ra, ix.dim);
  }
}
__device__ __forceinline__ float hxb_qscore(const HxDev& ix, const float* sq, float q_hdr, uint32_t slot, uint32_t t) {
  switch (ix.metric) {
    case HXM_EUCLIDEAN: return hx_octet_score<HXM_EUCLIDEAN>(ix, sq, q_hdr, slot, t);
    case HXM_COSINE: return hx_octet_score<HXM_COSINE>(ix, sq, q_hdr, slot, t);
    default: return hx_octet_score<HXM_MANHATTAN>(ix, sq, q_hdr, slot, t);
  }
}

// select_diverse (mod.rs:809-856) executed by a whole CTA.
//   cand_key[i] (ascending by (score,id)): score bits << 32 | slot;  only the first `ncand` are resolvable.
//   Writes up to m slots to sel[] (selection order) and returns the count through *s_nsel (shared).
__device__ void hxb_select_diverse(const HxDev& ix, const uint64_t* cand_key, uint32_t ncand, uint32_t m, uint32_t* sel,
                                   uint8_t* taken, uint32_t* s_nsel, uint32_t* s_reject) {
  const uint32_t tid = threadIdx.x, t = tid & 7u, oct = tid >> 3;
  if (tid == 0) { *s_nsel = 0; *s_reject = 0; }
  for (uint32_t i = tid; i < ncand; i += HXB_THREADS) taken[i] = 0;
  __syncthreads();
  for (uint32_t i = 0; i < ncand; ++i) {
    const uint32_t nsel = *s_nsel;
    if (nsel >= m) break;
    const uint32_t c = (uint32_t)(cand_key[i] & 0xffffffffu);
    const float thr = __uint_as_float((uint32_t)(cand_key[i] >> 32));
    for (uint32_t j = oct; j < nsel; j += HXB_THREADS / 8) {
      float pd = hxb_pair(ix, c, sel[j], t);
      if (t == 0 && pd < thr) *s_reject = 1;   // benign race: every writer stores 1
    }
    __syncthreads();
    if (tid == 0) {
      if (!*s_reject) { sel[nsel] = c; taken[i] = 1; *s_nsel = nsel + 1; }
      *s_reject = 0;
    }
    __syncthreads();
  }
  __syncthreads();   // every thread has read *s_nsel at the loop's exit test before thread 0 rewrites it (racecheck)
  if (tid == 0) {   // back-fill with the closest remaining (mod.rs:842-853)
    uint32_t nsel = *s_nsel;
    for (uint32_t i = 0; i < ncand && nsel < m; ++i)
      if (!taken[i]) { sel[nsel++] = (uint32_t)(cand_key[i] & 0xffffffffu); taken[i] = 1; }
    *s_nsel = nsel;
  }
  __syncthreads();
}

// ---- kernel A: search + neighbour selection for every node of the round ----------------------------------------------
// One CTA per new node.  Greedy descent above the node's level (mutation.rs:1008-1064), then per layer a beam search
// with the reference's admission rule `len < ef || d < w.max` (mutation.rs:953-979) and select_diverse on the result.
__global__ void __launch_bounds__(HXB_THREADS) k_build_search(HxDev ix, HxBuildArgs a, uint32_t ef_cap, uint32_t fr_cap) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* sq = reinterpret_cast<float*>(smem);
  uint64_t* beam_mem = reinterpret_cast<uint64_t*>(smem + (size_t)ix.ld * 4);
  uint32_t* frontier = reinterpret_cast<uint32_t*>(beam_mem + ef_cap);
  float* fdist = reinterpret_cast<float*>(frontier + fr_cap);
  uint32_t* sel = reinterpret_cast<uint32_t*>(fdist + fr_cap);        // [128]
  uint8_t* taken = reinterpret_cast<uint8_t*>(sel + 128);             // [256]
  __shared__ uint64_t s_tie[HX_TIE_CAP];   // evicted-unexpanded entries whose score equals w.max (see k_hnsw.cuh)
  __shared__ uint32_t s_nf, s_cur, s_done, s_epoch, s_changed, s_nsel, s_reject, s_len;
  __shared__ float s_cur_dist;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, t = tid & 7u, oct = tid >> 3;
  uint8_t* stamp = a.stamps + (size_t)blockIdx.x * a.stamp_stride;

  for (uint32_t bi = blockIdx.x; bi < a.batch_size; bi += gridDim.x) {
    const uint32_t u = a.batch_start + bi;
    const int node_layer = ix.level[u];
    uint16_t* pn = a.Pn + (size_t)bi * (a.maxL + 1);
    for (uint32_t l = tid; l <= a.maxL; l += HXB_THREADS) pn[l] = 0;
    if (a.cur_max_layer < 0) { __syncthreads(); continue; }   // empty graph: the node only gets empty rows
    const float* urow = ix.vec + (size_t)u * ix.ld;
    const float q_hdr = ix.hdr[u];
    for (uint32_t i = tid; i < ix.ld; i += HXB_THREADS) sq[i] = urow[i];
    __syncthreads();
    uint32_t cur = a.cur_entry;
    {
      float s = 0.f;
      if (oct == 0) s = hxb_qscore(ix, sq, q_hdr, cur, t);
      if (tid == 0) s_cur_dist = s;
    }
    __syncthreads();
    float cur_dist = s_cur_dist;

    for (int layer = a.cur_max_layer; layer >= 0; --layer) {
      if (layer > node_layer) {
        // ---- greedy step(s) on this layer
        for (;;) {
          uint32_t deg = 0;
          const uint32_t* row = nullptr;
          if ((int)ix.level[cur] >= layer) {
            const uint32_t r = hxb_row_id(ix, layer, cur);
            deg = *hxb_row_deg(ix, r);
            row = hxb_row_nbr(ix, r);
          }
          for (uint32_t f = oct; f < deg; f += HXB_THREADS / 8) {
            float s = hxb_qscore(ix, sq, q_hdr, row[f], t);
            if (t == 0) fdist[f] = s;
          }
          __syncthreads();
          if (warp == 0) {
            float best = cur_dist;
            uint32_t best_i = HX_ABSENT;
            for (uint32_t base = 0; base < deg; base += 32) {
              uint32_t f = base + lane;
              float m = f < deg ? fdist[f] : __int_as_float(0x7f800000);
              uint32_t mi = f;
              for (int o = 16; o > 0; o >>= 1) {
                float om = __shfl_xor_sync(0xffffffffu, m, o);
                uint32_t oi = __shfl_xor_sync(0xffffffffu, mi, o);
                if (om < m || (om == m && oi < mi)) { m = om; mi = oi; }
              }
              if (m < best) { best = m; best_i = mi; }
            }
            if (lane == 0) {
              if (best_i != HX_ABSENT) { s_cur = row[best_i]; s_cur_dist = best; s_changed = 1u; }
              else s_changed = 0u;
            }
          }
          __syncthreads();
          const uint32_t changed = s_changed;
          if (changed) { cur = s_cur; cur_dist = s_cur_dist; }
          __syncthreads();
          if (!changed) break;
        }
        continue;
      }
      // ---- beam search on `layer` (search_layer_beam)
      const uint32_t limit = layer == 0 ? a.lim0 : a.limu;
      uint32_t ef = layer == 0 ? max(a.efc, a.lim0) : max(a.efc, 2u * a.limu);   // mutation.rs:821-827
      if (ef > ef_cap) ef = ef_cap;
      // fresh visited set for this layer
      if (tid == 0) s_epoch = a.epochs[blockIdx.x] + 1u;
      __syncthreads();
      uint32_t epoch = s_epoch;
      if (epoch >= 256u) {
        uint4* s4 = reinterpret_cast<uint4*>(stamp);
        const size_t n16 = a.stamp_stride >> 4;
        for (size_t i = tid; i < n16; i += HXB_THREADS) s4[i] = make_uint4(0, 0, 0, 0);
        epoch = 1u;
      }
      __syncthreads();
      if (tid == 0) a.epochs[blockIdx.x] = epoch;
      const uint8_t ep8 = (uint8_t)epoch;
      HxBeam beam{beam_mem, 0u};
      uint32_t tie_len = 0;
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
          } else if (tie_len > 0) {   // `current.score > w.max` is strict: a candidate tied with w.max is still expanded
            tie_len--;
            cur_slot = (uint32_t)(s_tie[tie_len] & 0xffffffffu) >> 1;
          }
          uint32_t nf = 0;
          if (cur_slot != HX_ABSENT && (int)ix.level[cur_slot] >= layer) {
            const uint32_t r = hxb_row_id(ix, layer, cur_slot);
            const uint32_t deg = *hxb_row_deg(ix, r);
            const uint32_t* row = hxb_row_nbr(ix, r);
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
          }
          if (lane == 0) {
            s_nf = nf;
            s_done = (cur_slot == HX_ABSENT) ? 1u : 0u;
          }
        }
        __syncthreads();
        if (s_done) break;
        const uint32_t nf = s_nf;
        for (uint32_t f = oct; f < nf; f += HXB_THREADS / 8) {
          float s = hxb_qscore(ix, sq, q_hdr, frontier[f], t);
          if (t == 0) fdist[f] = s;
        }
        __syncthreads();
        if (warp == 0) {
          for (uint32_t base = 0; base < nf; base += 32) {
            const uint32_t f = base + lane;
            float s = f < nf ? fdist[f] : 0.f;
            uint32_t sbits = 0;
            bool pass = false;
            if (f < nf) {
              if (!hx_score_ok(s)) atomicOr(a.counters + 3, HXF_INVALID_SCORE);
              sbits = __float_as_uint(s);
              const uint32_t wmax = (uint32_t)(beam_mem[beam.len - 1] >> 32);
              pass = (sbits < wmax) || (beam.len < ef);
            }
            uint32_t mask = __ballot_sync(0xffffffffu, pass);
            while (mask) {
              const int src = __ffs(mask) - 1;
              mask &= mask - 1;
              const uint32_t xb = __shfl_sync(0xffffffffu, sbits, src);
              const uint32_t xslot = __shfl_sync(0xffffffffu, f < nf ? frontier[f] : 0u, src);
              const uint32_t wmax = (uint32_t)(beam_mem[beam.len - 1] >> 32);
              if (!((xb < wmax) || (beam.len < ef))) continue;
              const bool was_full = beam.len == ef;
              uint64_t ev;
              hx_beam_insert(beam, ef, ((uint64_t)xb << 32) | ((uint64_t)xslot << 1), &ev, lane);
              if (was_full) {
                const uint32_t new_wmax = (uint32_t)(beam_mem[beam.len - 1] >> 32);
                if (new_wmax < wmax) tie_len = 0;   // older ties are now above w.max: their pop would end the loop
                if (!(ev & 1ull) && (uint32_t)(ev >> 32) == new_wmax) {
                  if (tie_len < HX_TIE_CAP) {
                    if (lane == 0) s_tie[tie_len] = ev;
                    tie_len++;
                  } else if (lane == 0) {
                    atomicOr(a.counters + 3, HXF_TIE_OVERFLOW);
                  }
                }
                __syncwarp();
              }
            }
          }
        }
      }
      if (warp == 0 && lane == 0) s_len = beam.len;
      __syncthreads();
      const uint32_t len = s_len;
      // strip the expanded bit: key = score << 32 | slot
      for (uint32_t i = tid; i < len; i += HXB_THREADS) {
        const uint64_t k = beam_mem[i];
        beam_mem[i] = (k & 0xffffffff00000000ull) | ((k & 0xffffffffull) >> 1);
      }
      __syncthreads();
      const uint32_t ncand = min(len, 2u * limit);   // select_neighbors_heuristic resolves the first 2*M candidates
      hxb_select_diverse(ix, beam_mem, ncand, limit, sel, taken, &s_nsel, &s_reject);
      const uint32_t nsel = s_nsel;
      uint32_t* prow = a.P + (size_t)bi * a.pstride + (layer == 0 ? 0u : a.lim0 + (uint32_t)(layer - 1) * a.limu);
      for (uint32_t i = tid; i < nsel; i += HXB_THREADS) prow[i] = sel[i];
      if (tid == 0) pn[layer] = (uint16_t)nsel;
      // next layer starts from the closest candidate (mutation.rs:876-878)
      if (len) {
        cur = (uint32_t)(beam_mem[0] & 0xffffffffu);
        cur_dist = __uint_as_float((uint32_t)(beam_mem[0] >> 32));
      }
      __syncthreads();
    }
  }
}

// write `vals[0..n)` (unique) into the row in ascending order (stage_neighbors_vec_for_mutation, mutation.rs:1291-1307)
__device__ __forceinline__ void hxb_store_sorted(uint32_t* dst, uint16_t* deg, const uint32_t* vals, uint32_t n, uint32_t lane) {
  for (uint32_t i = lane; i < n; i += 32) {
    const uint32_t x = vals[i];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < n; ++j) rank += (vals[j] < x) ? 1u : 0u;
    dst[rank] = x;
  }
  if (lane == 0) *deg = (uint16_t)n;
}

// ---- sequential mode: add_bidirectional_link for ONE new node, targets in selection order --------------------------------
// mutation.rs:1498-1591 one link at a time: append `u` to v's row; if the row overflows, rank ALL its neighbours by distance
// to v, select_diverse to the limit, and remove the reciprocal edge from every rejected neighbour's own row — before the
// next link of the same insert is processed (a later target sees the rows the earlier ones rewrote).  One CTA; with
// k_build_search on a round of one node this reproduces insert_hnsw's graph exactly (tests/test_gpu_build_seq.py).
__device__ __forceinline__ void hxb_remove_edge(const HxDev& ix, uint32_t r, uint32_t x) {   // one thread
  uint32_t* row = hxb_row_nbr(ix, r);
  uint16_t* degp = hxb_row_deg(ix, r);
  const uint32_t deg = *degp;
  uint32_t w = 0;
  for (uint32_t j = 0; j < deg; ++j) {
    const uint32_t y = row[j];
    if (y != x) row[w++] = y;
  }
  *degp = (uint16_t)w;
}

__global__ void __launch_bounds__(HXB_THREADS) k_build_link_seq(HxDev ix, HxBuildArgs a) {
  __shared__ uint64_t keys[HXB_MAX_CAND];
  __shared__ uint64_t sorted[HXB_MAX_CAND];
  __shared__ uint32_t cand[HXB_MAX_CAND];
  __shared__ uint32_t sel[128];
  __shared__ uint8_t taken[HXB_MAX_CAND];
  __shared__ uint32_t s_nsel, s_reject;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, t = tid & 7u, oct = tid >> 3;
  const uint32_t u = a.batch_start;
  const int node_layer = ix.level[u];
  const int top = node_layer < a.cur_max_layer ? node_layer : a.cur_max_layer;
  for (int layer = top; layer >= 0; --layer) {
    const uint32_t limit = layer == 0 ? a.lim0 : a.limu;
    const uint32_t np = a.Pn[layer];
    const uint32_t* prow = a.P + (layer == 0 ? 0u : a.lim0 + (uint32_t)(layer - 1) * a.limu);
    const uint32_t ru = hxb_row_id(ix, layer, u);
    if (warp == 0) hxb_store_sorted(hxb_row_nbr(ix, ru), hxb_row_deg(ix, ru), prow, np, lane);   // stage_new_neighbors
    __syncthreads();
    for (uint32_t pi = 0; pi < np; ++pi) {
      const uint32_t v = prow[pi];
      const uint32_t rv = hxb_row_id(ix, layer, v);
      uint32_t* row = hxb_row_nbr(ix, rv);
      uint16_t* degp = hxb_row_deg(ix, rv);
      const uint32_t deg = *degp;
      bool has = false;
      for (uint32_t j = 0; j < deg; ++j) has |= (row[j] == u);
      const uint32_t nc = deg + (has ? 0u : 1u);
      for (uint32_t i = tid; i < nc; i += HXB_THREADS) cand[i] = i < deg ? row[i] : u;
      __syncthreads();
      if (nc <= limit) {
        if (warp == 0) hxb_store_sorted(row, degp, cand, nc, lane);
        __syncthreads();
        continue;
      }
      for (uint32_t i = oct; i < nc; i += HXB_THREADS / 8) {
        float d = hxb_pair(ix, v, cand[i], t);
        if (t == 0) {
          if (!hx_score_ok(d)) atomicOr(a.counters + 3, HXF_INVALID_SCORE);
          keys[i] = hx_make_key(d, cand[i]);
        }
      }
      __syncthreads();
      for (uint32_t i = tid; i < nc; i += HXB_THREADS) {   // distances.sort(): rank sort by (score, id)
        const uint64_t k = keys[i];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < nc; ++j) rank += (keys[j] < k) ? 1u : 0u;
        sorted[rank] = k;
      }
      __syncthreads();
      hxb_select_diverse(ix, sorted, nc, limit, sel, taken, &s_nsel, &s_reject);
      const uint32_t nsel = s_nsel;
      if (warp == 0) hxb_store_sorted(row, degp, sel, nsel, lane);
      __syncthreads();
      if (tid == 0) {   // remove the reciprocal edge of every rejected neighbour (u included, if v rejected it)
        for (uint32_t i = 0; i < nc; ++i) {
          const uint32_t c = cand[i];
          bool kept = false;
          for (uint32_t j = 0; j < nsel; ++j) kept |= (sel[j] == c);
          if (!kept) hxb_remove_edge(ix, hxb_row_id(ix, layer, c), v);
        }
      }
      __syncthreads();
    }
  }
}

// ---- kernel B0: group proposals by target row ------------------------------------------------------------------------------
__global__ void k_build_scatter(HxDev ix, HxBuildArgs a) {
  const uint32_t bi = blockIdx.x;
  const uint32_t u = a.batch_start + bi;
  const int node_layer = ix.level[u];
  const uint16_t* pn = a.Pn + (size_t)bi * (a.maxL + 1);
  for (int layer = 0; layer <= node_layer && layer <= (int)a.maxL; ++layer) {
    const uint32_t cntp = pn[layer];
    const uint32_t* prow = a.P + (size_t)bi * a.pstride + (layer == 0 ? 0u : a.lim0 + (uint32_t)(layer - 1) * a.limu);
    for (uint32_t j = threadIdx.x; j < cntp; j += blockDim.x) {
      const uint32_t v = prow[j];
      const uint32_t r = hxb_row_id(ix, layer, v);
      const uint32_t pos = atomicAdd(a.cnt + r, 1u);
      if (pos < HXB_CAP_IN) a.inc[(size_t)r * HXB_CAP_IN + pos] = u;
      else atomicAdd(a.counters + 2, 1u);
      if (pos == 0) {
        const uint32_t ti = atomicAdd(a.counters + 0, 1u);
        if (ti < a.max_targets) {
          a.targets[ti] = r;
          a.tidx[r] = ti;
        }
      }
    }
  }
}

// ---- kernel B1: every target decides which neighbours it keeps (add_bidirectional_link's prune) ------------------------------
__global__ void __launch_bounds__(HXB_THREADS) k_build_prune(HxDev ix, HxBuildArgs a) {
  __shared__ uint64_t keys[HXB_MAX_CAND];
  __shared__ uint64_t sorted[HXB_MAX_CAND];
  __shared__ uint32_t sel[128];
  __shared__ uint8_t taken[HXB_MAX_CAND];
  __shared__ uint32_t s_nsel, s_reject;
  const uint32_t tid = threadIdx.x, t = tid & 7u, oct = tid >> 3;
  const uint32_t n_targets = min(a.counters[0], a.max_targets);
  for (uint32_t ti = blockIdx.x; ti < n_targets; ti += gridDim.x) {
    const uint32_t r = a.targets[ti];
    const bool l0 = r < ix.n;
    const uint32_t v = l0 ? r : a.upper_row_node[r - ix.n];
    const uint32_t limit = l0 ? a.lim0 : a.limu;
    const uint32_t deg = *hxb_row_deg(ix, r);
    const uint32_t* row = hxb_row_nbr(ix, r);
    const uint32_t nin = min(a.cnt[r], (uint32_t)HXB_CAP_IN);
    const uint32_t nc = deg + nin;
    uint32_t* Rrow = a.R + (size_t)ti * a.rstride;
    if (nc <= limit) {   // no overflow: keep everything (mutation.rs:1515 `if to_neighbors.len() > maximum_neighbors`)
      for (uint32_t i = tid; i < nc; i += HXB_THREADS) Rrow[i] = i < deg ? row[i] : a.inc[(size_t)r * HXB_CAP_IN + (i - deg)];
      if (tid == 0) a.Rn[ti] = (uint16_t)nc;
      __syncthreads();
      continue;
    }
    // distances from the target to every current + incoming neighbour (mutation.rs:1521-1541)
    for (uint32_t i = oct; i < nc; i += HXB_THREADS / 8) {
      const uint32_t c = i < deg ? row[i] : a.inc[(size_t)r * HXB_CAP_IN + (i - deg)];
      float d = hxb_pair(ix, v, c, t);
      if (t == 0) {
        if (!hx_score_ok(d)) atomicOr(a.counters + 3, HXF_INVALID_SCORE);
        keys[i] = hx_make_key(d, c);
      }
    }
    __syncthreads();
    // distances.sort(): rank sort by (score, id)
    for (uint32_t i = tid; i < nc; i += HXB_THREADS) {
      const uint64_t k = keys[i];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < nc; ++j) rank += (keys[j] < k) ? 1u : 0u;
      sorted[rank] = k;
    }
    __syncthreads();
    hxb_select_diverse(ix, sorted, nc, limit, sel, taken, &s_nsel, &s_reject);
    const uint32_t nsel = s_nsel;
    for (uint32_t i = tid; i < nsel; i += HXB_THREADS) Rrow[i] = sel[i];
    if (tid == 0) a.Rn[ti] = (uint16_t)nsel;
    __syncthreads();
  }
}

// is x in the retained set chosen by target index ti
__device__ __forceinline__ bool hxb_in_R(const HxBuildArgs& a, uint32_t ti, uint32_t x) {
  const uint32_t* Rrow = a.R + (size_t)ti * a.rstride;
  const uint32_t n = a.Rn[ti];
  for (uint32_t i = 0; i < n; ++i)
    if (Rrow[i] == x) return true;
  return false;
}

// ---- kernel B2: mutual consent — an edge survives only if both endpoints keep it ------------------------------------------------
// One warp per item; items = targets, then the (node, layer) rows of the round's new nodes.
__global__ void __launch_bounds__(256) k_build_finalize(HxDev ix, HxBuildArgs a, uint32_t n_new_rows, const uint32_t* new_rows) {
  __shared__ uint32_t keep_all[8][HXB_MAX_CAND];
  const uint32_t warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  uint32_t* keep = keep_all[warp_in_block];
  const uint32_t n_targets = min(a.counters[0], a.max_targets);
  const uint32_t item = blockIdx.x * 8 + warp_in_block;
  const uint32_t batch_end = a.batch_start + a.batch_size;
  if (item < n_targets) {
    const uint32_t ti = item;
    const uint32_t r = a.targets[ti];
    const bool l0 = r < ix.n;
    const uint32_t v = l0 ? r : a.upper_row_node[r - ix.n];
    const int layer = l0 ? 0 : (int)a.upper_row_layer[r - ix.n];
    const uint32_t* Rrow = a.R + (size_t)ti * a.rstride;
    const uint32_t nR = a.Rn[ti];
    // consent of each retained neighbour
    uint32_t nk = 0;
    for (uint32_t base = 0; base < nR; base += 32) {
      const uint32_t i = base + lane;
      bool ok = false;
      uint32_t w = 0;
      if (i < nR) {
        w = Rrow[i];
        if (w >= a.batch_start && w < batch_end) ok = true;            // new node proposed this edge itself
        else {
          const uint32_t tw = a.tidx[hxb_row_id(ix, layer, w)];
          ok = (tw == HX_ABSENT) ? true : hxb_in_R(a, tw, v);          // untouched rows still hold the old symmetric edge
        }
      }
      const uint32_t mask = __ballot_sync(0xffffffffu, ok);
      if (ok) keep[nk + __popc(mask & ((1u << lane) - 1u))] = w;
      nk += __popc(mask);
    }
    __syncwarp();
    // old neighbours this row no longer keeps: remove the reciprocal edge from rows nobody rewrites this round
    uint32_t* row = hxb_row_nbr(ix, r);
    uint16_t* degp = hxb_row_deg(ix, r);
    const uint32_t deg = *degp;
    for (uint32_t base = 0; base < deg; base += 32) {
      const uint32_t i = base + lane;
      if (i < deg) {
        const uint32_t w = row[i];
        bool kept = false;
        for (uint32_t j = 0; j < nk; ++j) kept |= (keep[j] == w);
        if (!kept) {
          const uint32_t rw = hxb_row_id(ix, layer, w);
          if (a.tidx[rw] == HX_ABSENT) {
            uint32_t* wr = hxb_row_nbr(ix, rw);
            const uint32_t wd = *hxb_row_deg(ix, rw);
            for (uint32_t j = 0; j < wd; ++j)
              if (wr[j] == v) { wr[j] = HX_ABSENT; break; }
            if (atomicExch(a.vflag + rw, 1u) == 0u) a.victims[atomicAdd(a.counters + 1, 1u)] = rw;
          }
        }
      }
    }
    __syncwarp();
    hxb_store_sorted(row, degp, keep, nk, lane);
  } else if (item - n_targets < n_new_rows) {
    const uint32_t r = new_rows[item - n_targets];
    const bool l0 = r < ix.n;
    const uint32_t u = l0 ? r : a.upper_row_node[r - ix.n];
    const int layer = l0 ? 0 : (int)a.upper_row_layer[r - ix.n];
    const uint32_t bi = u - a.batch_start;
    uint32_t np = 0;
    const uint32_t* prow = a.P + (size_t)bi * a.pstride + (layer == 0 ? 0u : a.lim0 + (uint32_t)(layer - 1) * a.limu);
    if (layer <= (int)a.maxL) np = a.Pn[(size_t)bi * (a.maxL + 1) + layer];
    uint32_t nk = 0;
    for (uint32_t base = 0; base < np; base += 32) {
      const uint32_t i = base + lane;
      bool ok = false;
      uint32_t v = 0;
      if (i < np) {
        v = prow[i];
        const uint32_t tv = a.tidx[hxb_row_id(ix, layer, v)];
        ok = (tv != HX_ABSENT) && hxb_in_R(a, tv, u);
      }
      const uint32_t mask = __ballot_sync(0xffffffffu, ok);
      if (ok) keep[nk + __popc(mask & ((1u << lane) - 1u))] = v;
      nk += __popc(mask);
    }
    __syncwarp();
    hxb_store_sorted(hxb_row_nbr(ix, r), hxb_row_deg(ix, r), keep, nk, lane);
  }
}

// ---- kernel B3: reset per-round state, compact rows that lost edges ---------------------------------------------------------------
__global__ void k_build_cleanup(HxDev ix, HxBuildArgs a) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n_targets = min(a.counters[0], a.max_targets);
  const uint32_t n_victims = a.counters[1];
  if (i < n_targets) {
    const uint32_t r = a.targets[i];
    a.cnt[r] = 0;
    a.tidx[r] = HX_ABSENT;
  }
  if (i < n_victims) {
    const uint32_t r = a.victims[i];
    a.vflag[r] = 0;
    uint32_t* row = hxb_row_nbr(ix, r);
    uint16_t* degp = hxb_row_deg(ix, r);
    const uint32_t deg = *degp;
    uint32_t w = 0;
    for (uint32_t j = 0; j < deg; ++j) {
      const uint32_t x = row[j];
      if (x != HX_ABSENT) row[w++] = x;
    }
    *degp = (uint16_t)w;
  }
}

__global__ void k_build_reset_counters(uint32_t* counters) {
  if (threadIdx.x < 2) counters[threadIdx.x] = 0;
}

__global__ void k_copy_u16(uint16_t* dst, const uint16_t* src, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------------------------------
static inline uint64_t splitmix64(uint64_t& s) {
  uint64_t z = (s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// select_layer_from_uniform (mod.rs:774-796)
static uint16_t select_layer_from_uniform(float ml, float uniform) {
  if (!(std::isfinite(ml) && ml > 0.0f)) ml = 1.0f / std::log(16.0f);
  if (std::isfinite(uniform)) {
    const float lo = FLT_MIN, hi = 1.0f - FLT_EPSILON;
    uniform = std::min(std::max(uniform, lo), hi);
  } else {
    uniform = 0.5f;
  }
  const float sampled = std::floor(-std::log(uniform) * ml);
  if (!std::isfinite(sampled) || sampled <= 0.0f) return 0;
  return (uint16_t)std::min(sampled, 63.0f);
}

template <typename T>
static hx_status dalloc(T** p, size_t count) {
  cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T));
  if (e != cudaSuccess) {
    hx_set_error("cudaMalloc(%zu bytes) failed during build: %s", count * sizeof(T), cudaGetErrorString(e));
    return HX_ERR_OUT_OF_MEMORY;
  }
  return HX_OK;
}

hx_status hx_build_impl(hx_index* ix, const uint16_t* levels_in, uint64_t seed, int sequential) {
  const size_t n = ix->n;
  if (n == 0) return HX_OK;
  const uint32_t m = ix->cfg.m, lim0 = ix->lim0;
  if (lim0 > 128 || m > 128) {
    hx_set_error("device build supports degree limits up to 128 (m=%u, layer-0 limit=%u)", m, lim0);
    return HX_ERR_INVALID_VECTOR_CONFIG;
  }
  ix->free_graph();
  ix->graph_dirty.store(false, std::memory_order_release);   // the build writes the device graph itself
  // ---- levels -----------------------------------------------------------------------------------------------------
  std::vector<uint8_t> level(n);
  const float ml = 1.0f / std::log((float)std::max(m, 2u));   // default_ml_for_m (mod.rs:705-709)
  uint64_t st = seed ^ 0x5851f42d4c957f2dull;
  int top = 0;
  for (size_t i = 0; i < n; ++i) {
    uint16_t l;
    if (levels_in) l = std::min<uint16_t>(levels_in[i], 63);
    else {
      const float u = (float)(splitmix64(st) >> 40) * (1.0f / 16777216.0f);
      l = select_layer_from_uniform(ml, u);
    }
    level[i] = (uint8_t)l;
    top = std::max<int>(top, l);
  }
  std::vector<uint32_t> upper_off(n, HX_ABSENT);
  size_t rows = 0;
  for (size_t s = 0; s < n; ++s)
    if (level[s]) { upper_off[s] = (uint32_t)rows; rows += level[s]; }
  std::vector<uint32_t> row_node(std::max<size_t>(rows, 1));
  std::vector<uint8_t> row_layer(std::max<size_t>(rows, 1));
  for (size_t s = 0; s < n; ++s)
    for (uint32_t l = 1; l <= level[s]; ++l) {
      row_node[upper_off[s] + l - 1] = (uint32_t)s;
      row_layer[upper_off[s] + l - 1] = (uint8_t)l;
    }
  // ---- graph arrays ------------------------------------------------------------------------------------------------
  ix->stride0 = (lim0 + 31) / 32 * 32;
  ix->stride_u = (m + 15) / 16 * 16;
  ix->n_upper_rows = rows;
  hx_status rc;
  const size_t rcap = std::max(ix->cap_rows, n);   // per-row arrays follow the vector arrays' capacity (hx_mirror.inl)
  if ((rc = dalloc(&ix->d_nbr0, rcap * (size_t)ix->stride0))) return rc;
  if ((rc = dalloc(&ix->d_deg0, rcap))) return rc;
  if ((rc = dalloc(&ix->d_raw0, rcap))) return rc;
  if ((rc = dalloc(&ix->d_upper_off, rcap))) return rc;
  if ((rc = dalloc(&ix->d_upper_nbr, rows * (size_t)ix->stride_u))) return rc;
  if ((rc = dalloc(&ix->d_upper_deg, rows))) return rc;
  if ((rc = dalloc(&ix->d_level, rcap))) return rc;
  if (rcap > n) {
    HX_CUDA(cudaMemset(ix->d_deg0, 0, rcap * sizeof(uint16_t)));
    HX_CUDA(cudaMemset(ix->d_raw0, 0, rcap * sizeof(uint16_t)));
    HX_CUDA(cudaMemset(ix->d_upper_off, 0xFF, rcap * sizeof(uint32_t)));
    HX_CUDA(cudaMemset(ix->d_level, 0, rcap));
  }
  HX_CUDA(cudaMemset(ix->d_nbr0, 0, n * (size_t)ix->stride0 * sizeof(uint32_t)));
  HX_CUDA(cudaMemset(ix->d_deg0, 0, n * sizeof(uint16_t)));
  HX_CUDA(cudaMemset(ix->d_upper_nbr, 0, std::max<size_t>(rows, 1) * ix->stride_u * sizeof(uint32_t)));
  HX_CUDA(cudaMemset(ix->d_upper_deg, 0, std::max<size_t>(rows, 1) * sizeof(uint16_t)));
  HX_CUDA(cudaMemcpy(ix->d_upper_off, upper_off.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
  HX_CUDA(cudaMemcpy(ix->d_level, level.data(), n, cudaMemcpyHostToDevice));
  // ---- build-only state ----------------------------------------------------------------------------------------------
  const size_t urows = n + rows;
  uint32_t max_batch = (uint32_t)std::min<size_t>(16384, std::max<size_t>(1, n / 64));
  if (sequential) max_batch = 1;   // HX_BUILD_SEQUENTIAL: one insert at a time, links applied in selection order
  if (ix->tune.build_max_batch > 0 && !sequential)   // experiment knob: rounds never exceed this many nodes
    max_batch = (uint32_t)std::min<int32_t>(ix->tune.build_max_batch, 65536);
  const uint32_t maxL_all = (uint32_t)top;
  const uint32_t pstride = lim0 + maxL_all * m;
  const uint32_t rstride = std::max(lim0, m);
  const size_t max_targets = std::min<size_t>(urows, (size_t)max_batch * pstride);
  uint32_t *d_row_node = nullptr, *d_cnt = nullptr, *d_inc = nullptr, *d_tidx = nullptr, *d_targets = nullptr,
           *d_R = nullptr, *d_victims = nullptr, *d_vflag = nullptr, *d_counters = nullptr, *d_P = nullptr,
           *d_epochs = nullptr, *d_new_rows = nullptr;
  uint8_t *d_row_layer = nullptr, *d_stamps = nullptr;
  uint16_t *d_Rn = nullptr, *d_Pn = nullptr;
  const int per_sm = 4;
  const uint32_t grid_cap = (uint32_t)ix->sm_count * per_sm;
  const size_t stamp_stride = ((n + 15) / 16) * 16;
  auto cleanup = [&]() {
    cudaFree(d_row_node); cudaFree(d_cnt); cudaFree(d_inc); cudaFree(d_tidx); cudaFree(d_targets); cudaFree(d_R);
    cudaFree(d_victims); cudaFree(d_vflag); cudaFree(d_counters); cudaFree(d_P); cudaFree(d_epochs);
    cudaFree(d_new_rows); cudaFree(d_row_layer); cudaFree(d_stamps); cudaFree(d_Rn); cudaFree(d_Pn);
  };
#define HXB_TRY(expr)            \
  do {                           \
    hx_status _rc = (expr);      \
    if (_rc) { cleanup(); return _rc; } \
  } while (0)
#define HXB_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      hx_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      cleanup();                                                                             \
      return HX_ERR_CUDA;                                                                    \
    }                                                                                        \
  } while (0)
  HXB_TRY(dalloc(&d_row_node, rows));
  HXB_TRY(dalloc(&d_row_layer, rows));
  HXB_TRY(dalloc(&d_cnt, urows));
  HXB_TRY(dalloc(&d_inc, urows * HXB_CAP_IN));
  HXB_TRY(dalloc(&d_tidx, urows));
  HXB_TRY(dalloc(&d_targets, max_targets));
  HXB_TRY(dalloc(&d_R, max_targets * rstride));
  HXB_TRY(dalloc(&d_Rn, max_targets));
  HXB_TRY(dalloc(&d_victims, urows));
  HXB_TRY(dalloc(&d_vflag, urows));
  HXB_TRY(dalloc(&d_counters, 8));
  HXB_TRY(dalloc(&d_P, (size_t)max_batch * pstride));
  HXB_TRY(dalloc(&d_Pn, (size_t)max_batch * (maxL_all + 1)));
  HXB_TRY(dalloc(&d_new_rows, (size_t)max_batch * (maxL_all + 1)));
  HXB_TRY(dalloc(&d_stamps, (size_t)grid_cap * stamp_stride));
  HXB_TRY(dalloc(&d_epochs, grid_cap));
  HXB_CUDA(cudaMemcpy(d_row_node, row_node.data(), std::max<size_t>(rows, 1) * sizeof(uint32_t), cudaMemcpyHostToDevice));
  HXB_CUDA(cudaMemcpy(d_row_layer, row_layer.data(), std::max<size_t>(rows, 1), cudaMemcpyHostToDevice));
  HXB_CUDA(cudaMemset(d_cnt, 0, urows * sizeof(uint32_t)));
  HXB_CUDA(cudaMemset(d_tidx, 0xff, urows * sizeof(uint32_t)));
  HXB_CUDA(cudaMemset(d_vflag, 0, urows * sizeof(uint32_t)));
  HXB_CUDA(cudaMemset(d_counters, 0, 8 * sizeof(uint32_t)));
  HXB_CUDA(cudaMemset(d_stamps, 0, (size_t)grid_cap * stamp_stride));
  HXB_CUDA(cudaMemset(d_epochs, 0, grid_cap * sizeof(uint32_t)));

  HxBuildArgs a{};
  a.upper_row_node = d_row_node;
  a.upper_row_layer = d_row_layer;
  a.lim0 = lim0;
  a.limu = m;
  a.efc = ix->cfg.ef_construction;
  a.P = d_P;
  a.Pn = d_Pn;
  a.pstride = pstride;
  a.maxL = maxL_all;
  a.stamps = d_stamps;
  a.epochs = d_epochs;
  a.stamp_stride = stamp_stride;
  a.cnt = d_cnt;
  a.inc = d_inc;
  a.tidx = d_tidx;
  a.targets = d_targets;
  a.R = d_R;
  a.Rn = d_Rn;
  a.rstride = rstride;
  a.max_targets = (uint32_t)max_targets;
  a.victims = d_victims;
  a.vflag = d_vflag;
  a.counters = d_counters;

  const uint32_t ef_need = std::max(a.efc, std::max(lim0, 2 * m));
  const uint32_t ef_cap = std::min<uint32_t>(std::max<uint32_t>(ef_need, 16), 4096);
  const uint32_t fr_cap = (std::max(ix->stride0, ix->stride_u) + 31) / 32 * 32;
  const uint32_t smemA = ix->ld * 4u + ef_cap * 8u + fr_cap * 8u + 128 * 4u + 256;
  if (smemA > 48 * 1024) HXB_CUDA(cudaFuncSetAttribute(k_build_search, cudaFuncAttributeMaxDynamicSharedMemorySize, smemA));

  // the staged HxDev the kernels see (graph arrays are live)
  ix->populated = false;
  ix->max_layer = 0;
  HxDev dev = ix->dev();
  size_t inserted = 0;
  int cur_max_layer = -1;
  uint32_t cur_entry = 0;
  std::vector<uint32_t> new_rows;
  while (inserted < n) {
    size_t batch = inserted == 0 ? 1 : std::min<size_t>(std::max<size_t>(1, inserted / 2), max_batch);
    batch = std::min(batch, n - inserted);
    a.batch_start = (uint32_t)inserted;
    a.batch_size = (uint32_t)batch;
    a.cur_entry = cur_entry;
    a.cur_max_layer = cur_max_layer;
    // rows owned by this round's nodes (for the consent pass)
    new_rows.clear();
    int batch_top = 0;
    uint32_t batch_top_node = (uint32_t)inserted;
    for (size_t i = 0; i < batch; ++i) {
      const uint32_t u = (uint32_t)(inserted + i);
      new_rows.push_back(u);
      for (uint32_t l = 1; l <= level[u]; ++l) new_rows.push_back((uint32_t)(n + upper_off[u] + l - 1));
      if ((int)level[u] > batch_top) { batch_top = level[u]; batch_top_node = u; }
    }
    HXB_CUDA(cudaMemcpyAsync(d_new_rows, new_rows.data(), new_rows.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, 0));
    k_build_reset_counters<<<1, 32>>>(d_counters);
    const uint32_t gridA = (uint32_t)std::min<size_t>(batch, grid_cap);
    k_build_search<<<gridA, HXB_THREADS, smemA>>>(dev, a, ef_cap, fr_cap);
    HXB_CUDA(cudaGetLastError());
    if (sequential) {
      if (cur_max_layer >= 0) k_build_link_seq<<<1, HXB_THREADS>>>(dev, a);
      // (an empty graph: the node only gets its empty rows, which the zero-initialised arrays already are)
    } else if (cur_max_layer >= 0) {
      k_build_scatter<<<(unsigned)batch, 64>>>(dev, a);
      k_build_prune<<<(unsigned)std::min<size_t>(max_targets, (size_t)grid_cap * 2), HXB_THREADS>>>(dev, a);
    }
    // number of targets is only known on the device: size the consent grid by its upper bound for this round
    if (!sequential) {
      const size_t tb = std::min<size_t>(max_targets, batch * (size_t)pstride);
      const size_t items = (cur_max_layer >= 0 ? tb : 0) + new_rows.size();
      k_build_finalize<<<(unsigned)((items + 7) / 8), 256>>>(dev, a, (uint32_t)new_rows.size(), d_new_rows);
      k_build_cleanup<<<(unsigned)((urows + 255) / 256), 256>>>(dev, a);
    }
    HXB_CUDA(cudaGetLastError());
    HXB_CUDA(cudaStreamSynchronize(0));
    // entry point / max layer (mutation.rs:706-739,769-772)
    if (cur_max_layer < 0) {
      cur_max_layer = level[inserted];
      cur_entry = (uint32_t)inserted;
    } else if (batch_top > cur_max_layer) {
      cur_max_layer = batch_top;
      cur_entry = batch_top_node;
    }
    inserted += batch;
  }
  uint32_t counters[8];
  HXB_CUDA(cudaMemcpy(counters, d_counters, sizeof(counters), cudaMemcpyDeviceToHost));
  k_copy_u16<<<(unsigned)((n + 255) / 256), 256>>>(ix->d_raw0, ix->d_deg0, n);
  HXB_CUDA(cudaDeviceSynchronize());
  cleanup();
  if (counters[3] & HXF_INVALID_SCORE) {
    hx_set_error("vector distance kernel emitted an invalid score during build");
    return HX_ERR_INVARIANT_VIOLATION;
  }
  if (counters[3] & HXF_TIE_OVERFLOW) {
    hx_set_error("more than %d exact score ties at the insertion beam's boundary", HX_TIE_CAP);
    return HX_ERR_INVARIANT_VIOLATION;
  }
  ix->populated = true;
  ix->entry_slot = cur_entry;
  ix->entry_id = ix->ids_sorted[cur_entry];
  ix->max_layer = cur_max_layer;
  ix->graph_dirty = false;
  ix->staged.clear();
  return HX_OK;
}
