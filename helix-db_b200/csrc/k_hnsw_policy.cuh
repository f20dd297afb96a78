// k_hnsw_policy.cuh — layer-0 search with the reference's SimHash filtering / sampling / adaptive-bypass policy
// (SimHashMode::{Adaptive, Always}, or Off with a pre-sampling override < 1): the production-default path.
//
// Restates search.rs:267-1067 with STRICT_EXHAUSTIVE = false, policy.rs:52-597 (Layer0Policy::decide and its helpers),
// randomness.rs:96-165 (SearchSession) and unaligned_vector/simhash.rs:20-55.  The parity tests compare this kernel bit
// for bit — ids, scores, SearchStats and the SimHash counters — with the CPU restatement of the same lines kept with the
// test infrastructure, whose header states what the reference itself pins: the policy functions are pinned by the
// reference's policy tests; the session RNG stream (rand 0.10 StdRng = ChaCha12) is restated from the published
// algorithm and is consulted only for frontiers larger than max(ef/4, 8).
//
// One warp per query, rows through the same mbarrier ring and the same L2-resident visited hash set as k_hnsw_ring.cuh.
// Per expansion: unvisited neighbours by a read-only probe -> decision (warp-uniform scalars) -> SimHash threshold gate
// (one 8-byte fingerprint per neighbour, popcount) -> sampling -> only the survivors' vector rows are fetched — the
// point of the filter: fewer 3 KB row reads per expansion.
#pragma once
#include "k_hnsw_ring.cuh"

enum { HXP_OFF = 0, HXP_ADAPTIVE = 1, HXP_ALWAYS = 2 };                 // hx_simhash_mode
enum { HXP_READY = 0, HXP_BYPASSING = 1, HXP_COOLING = 2 };
enum { HXP_EXHAUSTIVE = 0, HXP_FIXED = 1, HXP_ADAPTIVE_S = 2 };

struct HxPolicyCfg {   // SearchParams + VectorIndexConfig fields the policy reads
  int32_t mode;
  uint32_t threshold;
  float sampling_ratio;
  int32_t has_pre_override;
  float pre_override;
  int32_t adaptive_enabled;
  float failure_prob;
  uint32_t bypass_min_frontier, bypass_window_expansions;
  float bypass_min_filter_rate;
  uint32_t read_budget_multiplier;
  float threshold_margin;   // sqrt(64 * ln(1/failure) / 2), computed once on the host (policy.rs:592)
};

struct HxPolicyArgs {
  HxPolicyCfg cfg;
  const uint64_t* node_simhash;   // [n] slot order
  const uint8_t* node_has_simhash;   // [n] or nullptr (= every node has one)
  const uint64_t* query_simhash;  // [B]
  unsigned long long* pstats;     // optional [12] sums (hx_policy_stats order)
};

struct HxDecision {
  int fetch_missing, filter_cached, has_threshold;
  uint32_t threshold;
  int pre_kind;
  float pre_prob;
  int samp_kind;
  float samp_prob;
  int bypassed, next_state;
  uint32_t next_remaining;
  int trigger;
};

__device__ __forceinline__ float hxp_clamp(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

// policy.rs:558-574
__device__ __forceinline__ float hxp_adaptive_sampling_ratio(float base, uint32_t ef, uint32_t search_len, float current,
                                                             float delta) {
  if (base >= 1.0f) return base;
  if (search_len < max(ef / 3u, 8u)) return 1.0f;
  if (delta <= 1e-6f) return base;
  const float rq = hxp_clamp(__fsub_rn(1.0f, hxp_clamp(__fdiv_rn(current, delta), 0.0f, 1.0f)), 0.0f, 1.0f);
  const float v = hxp_clamp(__fadd_rn(base, __fmul_rn(__fsub_rn(1.0f, base), rq)), base, 1.0f);
  const float cap = 0.90f > base ? 0.90f : base;
  return v < cap ? v : cap;
}
// policy.rs:576-597.  The margin sqrt(64 ln(1/eps) / 2) is a per-query constant (host libm, passed in); acos goes through
// double precision and is rounded once to f32 (the integer threshold only changes at a floor boundary).  FP64 is slow on
// this part, and delta only moves when the running top-k improves: callers memoise on delta's bits.
__device__ __forceinline__ uint32_t hxp_adaptive_threshold(bool topk_ready, float delta, uint32_t configured, float margin) {
  if (configured == 0u) return 0u;
  if (!topk_ready) return 1u;
  const float normalized = hxp_clamp(delta, 0.0f, 1.0f);
  const float cs = hxp_clamp(__fsub_rn(1.0f, __fmul_rn(2.0f, normalized)), -1.0f, 1.0f);
  const float ac = (float)acos((double)cs);
  const float collision = __fsub_rn(1.0f, __fdiv_rn(ac, 3.14159274101257324f));
  const float t = hxp_clamp(floorf(__fsub_rn(__fmul_rn(64.0f, collision), margin)), 1.0f, 64.0f);
  const uint32_t v = (uint32_t)t;
  return v < configured ? v : configured;
}

// Layer0Policy::decide (policy.rs:119-175) with AdaptiveBypassPolicy::decide (:222-291)
__device__ __forceinline__ HxDecision hxp_decide(int metric, const HxPolicyCfg& c, bool topk_ready, uint32_t ef,
                                                 uint32_t search_len, uint32_t frontier_len, float current, float delta,
                                                 int bstate, uint32_t bremaining, uint64_t filter_reads, uint64_t w_examined,
                                                 uint64_t w_filtered, uint64_t w_expansions, uint32_t& memo_key,
                                                 uint32_t& memo_thr) {
  HxDecision d;
  d.fetch_missing = d.filter_cached = d.has_threshold = 0;
  d.threshold = 0;
  d.bypassed = 0;
  d.next_state = HXP_READY;
  d.next_remaining = 0;
  d.trigger = 0;
  bool bypassed = false;
  if (c.mode == HXP_ADAPTIVE) {
    const uint32_t window = c.bypass_window_expansions;
    uint64_t rb = (uint64_t)ef * c.read_budget_multiplier;
    if (rb < c.bypass_min_frontier) rb = c.bypass_min_frontier;
    bool decided = false;
    if (bstate == HXP_BYPASSING) {
      bypassed = true;
      if (bremaining - 1u > 0u) { d.next_state = HXP_BYPASSING; d.next_remaining = bremaining - 1u; }
      else { d.next_state = HXP_COOLING; d.next_remaining = window; }
      decided = true;
    } else if (bstate == HXP_COOLING && bremaining > 1u) {
      d.next_state = HXP_COOLING;
      d.next_remaining = bremaining - 1u;
      decided = true;
    }
    if (!decided) {
      const bool budget = filter_reads >= rb;
      const float rate = w_examined == 0 ? 1.0f : __fdiv_rn((float)w_filtered, (float)w_examined);
      const bool low = w_expansions >= window && rate < c.bypass_min_filter_rate;
      const int trg = (budget ? 1 : 0) | (low ? 2 : 0);
      if (!(frontier_len < c.bypass_min_frontier || trg == 0)) {
        bypassed = true;
        d.trigger = trg;
        if (window - 1u > 0u) { d.next_state = HXP_BYPASSING; d.next_remaining = window - 1u; }
        else { d.next_state = HXP_COOLING; d.next_remaining = window; }
      }
    }
  }
  int filtering = 0;   // 0 disabled, 1 fixed, 2 adaptive (policy.rs:67-88)
  if (c.mode != HXP_OFF && metric == HXM_COSINE) filtering = c.mode == HXP_ALWAYS ? 1 : (c.adaptive_enabled ? 2 : 1);
  int base_kind;
  float base_prob;
  if (c.mode == HXP_OFF) { base_kind = HXP_EXHAUSTIVE; base_prob = 1.0f; }
  else if (c.mode == HXP_ADAPTIVE && c.adaptive_enabled) {
    base_kind = HXP_ADAPTIVE_S;
    base_prob = hxp_adaptive_sampling_ratio(c.sampling_ratio, ef, search_len, current, delta);
  } else { base_kind = HXP_FIXED; base_prob = c.sampling_ratio; }
  const uint32_t gate = max(ef / 4u, 8u);
  if (base_prob <= 0.0f || base_prob >= 1.0f || frontier_len > gate) { d.samp_kind = base_kind; d.samp_prob = base_prob; }
  else { d.samp_kind = HXP_EXHAUSTIVE; d.samp_prob = 1.0f; }
  const float pre_base = c.has_pre_override ? c.pre_override : base_prob;
  if (pre_base >= 1.0f || frontier_len <= gate) { d.pre_kind = HXP_EXHAUSTIVE; d.pre_prob = 1.0f; }
  else {
    float ratio = hxp_clamp(__fmul_rn(pre_base, 0.65f), 0.25f, 0.9f);
    const uint64_t twice = (uint64_t)ef * 2u;
    if (pre_base <= 0.0f) ratio = 0.0f;
    else if ((uint64_t)frontier_len > (twice > 32u ? twice : 32u)) { const float r = __fmul_rn(ratio, 0.8f); ratio = r > 0.20f ? r : 0.20f; }
    d.pre_kind = HXP_FIXED;
    d.pre_prob = ratio;
  }
  if (bypassed) { d.bypassed = 1; return d; }
  if (filtering == 0) return d;
  d.fetch_missing = d.filter_cached = d.has_threshold = 1;
  if (filtering == 1) {
    d.threshold = c.threshold;
  } else {
    const uint32_t mk = topk_ready ? __float_as_uint(delta) : 0xFFFFFFFEu;   // valid scores are below 0x7f800000
    if (mk != memo_key) {
      memo_thr = hxp_adaptive_threshold(topk_ready, delta, c.threshold, c.threshold_margin);
      memo_key = mk;
    }
    d.threshold = memo_thr;
  }
  return d;
}
// SamplingDecision::candidate_probability (policy.rs:415-430)
__device__ __forceinline__ float hxp_candidate_probability(const HxDecision& d, uint32_t sim) {
  const float base = d.samp_kind == HXP_EXHAUSTIVE ? 1.0f : d.samp_prob;
  if (d.samp_kind != HXP_ADAPTIVE_S) return base;
  if (base <= 0.0f || base >= 1.0f) return base;
  const float sr = __fdiv_rn((float)(sim < 64u ? sim : 64u), 64.0f);
  const float tr = d.has_threshold ? __fdiv_rn((float)d.threshold, 64.0f) : 0.0f;
  float diff = __fsub_rn(sr, tr);
  if (!(diff > 0.0f)) diff = 0.0f;
  return hxp_clamp(__fadd_rn(base, __fmul_rn(__fsub_rn(1.0f, base), diff)), base, 1.0f);
}

// ---- SearchSession over ChaCha12 (one lane; state in shared memory: key[8] | buf[16] | pos | started | block lo/hi) ----
struct HxSession {
  uint32_t* m;     // 28 words of shared memory
  uint64_t seed;
};
__device__ __forceinline__ uint32_t hxp_rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
#define HXP_QR(a, b, c, d) \
  a += b; d ^= a; d = hxp_rotl(d, 16); c += d; b ^= c; b = hxp_rotl(b, 12); \
  a += b; d ^= a; d = hxp_rotl(d, 8);  c += d; b ^= c; b = hxp_rotl(b, 7);
__device__ __noinline__ void hxp_session_refill(uint32_t* m) {
  uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7],
                     m[26], m[27], 0u, 0u};
  uint32_t x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = in[i];
  for (int i = 0; i < 6; ++i) {
    HXP_QR(x[0], x[4], x[8], x[12]) HXP_QR(x[1], x[5], x[9], x[13]) HXP_QR(x[2], x[6], x[10], x[14]) HXP_QR(x[3], x[7], x[11], x[15])
    HXP_QR(x[0], x[5], x[10], x[15]) HXP_QR(x[1], x[6], x[11], x[12]) HXP_QR(x[2], x[7], x[8], x[13]) HXP_QR(x[3], x[4], x[9], x[14])
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) m[8 + i] = x[i] + in[i];
  const uint64_t blk = (((uint64_t)m[27] << 32) | m[26]) + 1ull;
  m[26] = (uint32_t)blk;
  m[27] = (uint32_t)(blk >> 32);
  m[24] = 0;
}
__device__ __forceinline__ uint32_t hxp_session_next(HxSession& s) {   // one lane only
  uint32_t* m = s.m;
  if (!m[25]) {   // seed_from_u64: PCG32 fills the key
    uint64_t st = s.seed;
    for (int i = 0; i < 8; ++i) {
      st = st * 6364136223846793005ull + 11634580027462260723ull;
      const uint32_t xs = (uint32_t)(((st >> 18) ^ st) >> 27);
      const uint32_t rot = (uint32_t)(st >> 59);
      m[i] = (xs >> rot) | (xs << ((32u - rot) & 31u));
    }
    m[26] = m[27] = 0;
    m[24] = 16;
    m[25] = 1;
  }
  if (m[24] >= 16u) hxp_session_refill(m);
  return m[8 + m[24]++];
}
__device__ __forceinline__ bool hxp_should_sample(HxSession& s, float ratio) {
  if (ratio >= 1.0f) return true;
  if (ratio <= 0.0f) return false;
  const float u = __fmul_rn((float)(hxp_session_next(s) >> 8), 1.0f / 16777216.0f);
  return u < ratio;
}
__device__ __forceinline__ uint32_t hxp_choose_index(HxSession& s, uint32_t count) {   // count > 0
  const uint64_t mm = (uint64_t)hxp_session_next(s) * count;
  uint32_t result = (uint32_t)(mm >> 32);
  const uint32_t lo = (uint32_t)mm;
  if (lo > (uint32_t)(0u - count)) {
    const uint32_t nh = (uint32_t)(((uint64_t)hxp_session_next(s) * count) >> 32);
    if ((uint64_t)lo + nh > 0xffffffffull) result += 1u;
  }
  return result;
}

// read-only membership probe of the visited set
__device__ __forceinline__ bool hx_vt_contains(const HxVisited& v, uint32_t key) {
  uint32_t h = (key * 2654435761u) >> v.shift;
  for (;;) {
    const uint32_t cur = *((volatile uint32_t*)(v.tab + h));
    if (cur == HX_VT_EMPTY) return false;
    if (cur == key) return true;
    h = (h + 1u) & v.mask;
  }
}

// shared memory per warp: query | R row slots | beam[ef] | topk[k] | tie | R mbarriers | frontier | fdist | fhdr(aliases fsim) |
//                         fstate (bytes) | session (28 words)
#define HX_POLICY_MAX_THREADS 512   // warp-per-query build; the CTA-per-query build launches 8 warps (256 threads: up to 255 registers, no spills)
// CTA = false: one warp per query (throughput).  CTA = true (B < #SMs): one CTA per query — warp 0 runs the very same
// per-query code, and whenever it has rows to score it wakes the other warps, which issue and reduce their share of the
// rows (row r -> warp r mod W), exactly like the latency build of the exhaustive kernel.
template <int METRIC, int QCH, bool CTA>
__global__ void __launch_bounds__(CTA ? 256 : HX_POLICY_MAX_THREADS, 1) k_hnsw_search_policy(HxDev ix, HxHnswArgs a, HxRingArgs rg,
                                                                                HxPolicyArgs pa, uint32_t wstride, uint32_t R) {
  extern __shared__ __align__(128) unsigned char smem[];
  // CTA mode: warp 0 (the query) and the helper warps meet from different places in the code, so the rendezvous is a pair
  // of mbarriers, not __syncthreads(): s_bar_cmd (1 arrival: warp 0 posts a command) and s_bar_done (one arrival per helper
  // warp: its share of the rows is scored and visible)
  __shared__ uint32_t s_cmd, s_cnt;      // warp 0 -> helpers (1 = score s_cnt rows of `frontier`, 2 = done)
  __shared__ __align__(8) uint64_t s_bar_cmd, s_bar_done;
  __shared__ float s_qhdr;
  __shared__ const float* s_qg;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t warps_per_cta = blockDim.x >> 5;
  const uint32_t W = warps_per_cta;
  const uint32_t gw = CTA ? blockIdx.x : blockIdx.x * warps_per_cta + warp;
  unsigned char* wmem = smem + (CTA ? (size_t)0 : (size_t)warp * wstride);
  constexpr bool Q_SMEM = QCH == 0 || METRIC == HXM_MANHATTAN;   // Manhattan walks the query sequentially: keep it in smem
  float* sq = reinterpret_cast<float*>(wmem);                                            // [ld] when Q_SMEM
  float* ring = sq + (Q_SMEM ? ix.ld : 0u);                                              // [R][ld]
  uint64_t* beam_mem = reinterpret_cast<uint64_t*>(ring + (size_t)R * ix.ld);           // [ef]
  uint64_t* topk_mem = beam_mem + a.ef;                                                  // [k]
  uint64_t* tie = topk_mem + a.k;                                                        // [HX_TIE_CAP]
  uint64_t* bars = tie + HX_TIE_CAP;                                                     // [R]
  uint32_t* frontier = reinterpret_cast<uint32_t*>(bars + R);                            // [fr_cap]
  float* fdist = reinterpret_cast<float*>(frontier + a.fr_cap);                         // [fr_cap]
  float* fhdr = fdist + a.fr_cap;                                                        // [fr_cap]
  uint32_t* fsim = reinterpret_cast<uint32_t*>(fhdr);                                    // aliases fhdr (used before scoring)
  uint32_t* sess_mem = reinterpret_cast<uint32_t*>(fhdr + a.fr_cap);                     // [28]
  uint8_t* fstate = reinterpret_cast<uint8_t*>(sess_mem + 28);                           // [fr_cap]: 0 dropped, 1 sampled, 2 deferred
  const unsigned FULL = 0xffffffffu;
  const uint32_t rowbytes = ix.ld * 4u;
  const uint64_t policy = hx_policy_evict_first();
  uint32_t ph = 0;
  uint32_t pc = 0, pd = 0;   // phase parity of the command / done barriers
  if (CTA) {
    if (threadIdx.x < R) hx_mbar_init(bars + threadIdx.x, 1);
    if (threadIdx.x == 0) {
      hx_mbar_init(&s_bar_cmd, 1);
      hx_mbar_init(&s_bar_done, W > 1 ? W - 1 : 1);
    }
    hx_fence_mbar_init();
    __syncthreads();
  } else {
    if (lane < R) hx_mbar_init(bars + lane, 1);
    hx_fence_mbar_init();
    __syncwarp();
  }
  float qr[(!Q_SMEM && QCH > 0) ? QCH : 1];
  const float* qg = nullptr;
  float q_hdr = 0.f;
  uint32_t ps[12];   // per-warp sums (a warp handles a handful of queries per launch: 32 bits are plenty)
#pragma unroll
  for (int i = 0; i < 12; ++i) ps[i] = 0;   // lane 0's copy is the one that is reported

  auto issue = [&](uint32_t s, uint32_t slot) {
    hx_mbar_expect_tx(bars + s, rowbytes);
    if (rg.l2_hint) hx_bulk_g2s_hint(ring + (size_t)s * ix.ld, ix.vec + (size_t)slot * ix.ld, rowbytes, bars + s, policy);
    else hx_bulk_g2s(ring + (size_t)s * ix.ld, ix.vec + (size_t)slot * ix.ld, rowbytes, bars + s);
  };
  // CTA mode: every warp's share of one pass of `rows` rows starting at list[base] (slot = row index within the pass)
  auto score_share = [&](const uint32_t* list, uint32_t base, uint32_t rows) {
    const uint32_t mine = lane * W + warp;
    if (METRIC == HXM_MANHATTAN) {   // one strictly sequential chain per row (simple.rs:186-202): one thread per row, from global memory
      if (mine < rows) fdist[base + mine] = hx_manhattan_seq(ix.vec + (size_t)list[base + mine] * ix.ld, sq, ix.dim);
      return;
    }
    float rh = 0.f;
    if (mine < rows) {
      const uint32_t slot = list[base + mine];
      issue(mine, slot);
      if (METRIC == HXM_COSINE) rh = __ldg(ix.hdr + slot);
    }
    uint32_t j = 0;
    for (uint32_t r = warp; r < rows; r += W, ++j) {
      const float row_hdr = __shfl_sync(FULL, rh, j);
      hx_mbar_wait(bars + r, (ph >> r) & 1u);
      const float sc = hx_warp_score<(METRIC == HXM_MANHATTAN ? HXM_EUCLIDEAN : METRIC), (Q_SMEM ? 0 : QCH)>(
          ring + (size_t)r * ix.ld, qr, sq, qg, q_hdr, row_hdr, ix.dim, lane);
      if (lane == 0) fdist[base + r] = sc;
    }
    ph ^= rows >= 32u ? FULL : ((1u << rows) - 1u);
  };
  if (CTA && warp != 0) {   // helpers: score on command until warp 0 says done
    for (;;) {
      hx_mbar_wait(&s_bar_cmd, pc);
      pc ^= 1u;
      if (s_cmd == 2u) return;
      const uint32_t cnt = s_cnt;
      q_hdr = s_qhdr;
      qg = s_qg;
      // a row slot is only ever touched by the warp that owns it (slot r -> warp r mod W): passes need no rendezvous
      for (uint32_t base = 0; base < cnt; base += R) score_share(frontier, base, min(R, cnt - base));
      __syncwarp();
      if (lane == 0) hx_mbar_arrive(&s_bar_done);
    }
  }
  auto score_list = [&](const uint32_t* list, uint32_t cnt) {
    if (CTA) {   // list == frontier (always, in this kernel)
      if (cnt == 0) return;
      if (lane == 0) {
        s_cmd = 1u; s_cnt = cnt; s_qhdr = q_hdr; s_qg = qg;
        hx_mbar_arrive(&s_bar_cmd);
      }
      __syncwarp();
      for (uint32_t base = 0; base < cnt; base += R) score_share(list, base, min(R, cnt - base));
      if (W > 1) {   // every helper's scores are in fdist (and `list` may be rewritten)
        hx_mbar_wait(&s_bar_done, pd);
        pd ^= 1u;
      }
      __syncwarp();
      return;
    }
    if (METRIC == HXM_MANHATTAN) {   // lane f walks row f's sequential chain, straight from global memory
      for (uint32_t f = lane; f < cnt; f += 32) fdist[f] = hx_manhattan_seq(ix.vec + (size_t)list[f] * ix.ld, sq, ix.dim);
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
      const float sc = hx_warp_score<(METRIC == HXM_MANHATTAN ? HXM_EUCLIDEAN : METRIC), (Q_SMEM ? 0 : QCH)>(
          ring + (size_t)s * ix.ld, qr, sq, qg, q_hdr, METRIC == HXM_COSINE ? fhdr[j] : 0.f, ix.dim, lane);
      if (lane == 0) fdist[j] = sc;
      __syncwarp();
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
    uint32_t qflags = 0;   // error flags raised by THIS query
    if (a.q_status[qi] != 0u || !ix.populated) {
      if (lane == 0) a.out_counts[qi] = 0;
      hx_query_done(a, qi, 0u, lane);
      continue;
    }
    q_hdr = a.q_hdr[qi];
    qg = a.queries + (size_t)qi * ix.dim;
    if (Q_SMEM) {
      for (uint32_t i = lane; i < ix.ld; i += 32) sq[i] = i < ix.dim ? qg[i] : 0.0f;
    } else {
#pragma unroll
      for (int c = 0; c < ((!Q_SMEM && QCH > 0) ? QCH : 1); ++c) qr[c] = (uint32_t)(c * 32) + lane < ix.dim ? qg[c * 32 + lane] : 0.f;
    }
    const uint64_t qsim = pa.query_simhash[qi];
    // CTA build: the visited set of the one query lives in shared memory (after the warp-0 region); it moves to a pool
    // table in global memory only if it outgrows it
    uint32_t* vt_home = CTA ? reinterpret_cast<uint32_t*>(smem + wstride) : rg.vtab + (size_t)gw * rg.vt_cap;
    HxVisited vt = hx_vt_make(vt_home, rg.vt_cap);
    int pool_idx = -1;
    bool failed = false;
    hx_vt_clear_warp(vt.tab, rg.vt_cap, lane);

    // ---- entry point + upper layers: identical to the exhaustive builds (search.rs:1150-1156)
    uint32_t cur = ix.entry_slot;
    if (lane == 0) frontier[0] = cur;
    __syncwarp();
    score_list(frontier, 1);
    float cur_dist = fdist[0];
    if (!hx_score_ok(cur_dist)) qflags |= HXF_INVALID_SCORE;
    uint32_t upper_steps = 0;
    __syncwarp();
    for (int layer = ix.max_layer; layer >= 1; --layer) {
      for (;;) {
        uint32_t deg = 0;
        {
          const uint32_t off = ix.upper_off[cur];
          if (off != HX_ABSENT && (int)ix.level[cur] >= layer) {
            deg = ix.upper_deg[off + (uint32_t)layer - 1u];
            const uint32_t* row = ix.upper_nbr + (size_t)(off + (uint32_t)layer - 1u) * ix.stride_u;
            for (uint32_t f = lane; f < deg; f += 32) frontier[f] = row[f];
          }
        }
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

    // ---- layer 0 with the policy (search.rs:500-992)
    HxBeam beam{beam_mem, 1u};
    // The reference's top-k tracker (a max-heap of the k best admitted candidates, search.rs:933-938) needs no storage of its
    // own here: every admitted candidate enters the beam, the beam keeps the ef >= k best of them sorted, so the tracker IS
    // the beam's first min(k, len) entries and its maximum is beam[min(k, len) - 1].
    const uint32_t topk_target = a.k > 1u ? a.k : 1u;   // == a.k (k >= 1)
    HxTie tq = hx_tie_make(tie);
    uint32_t dropped = 0;
    uint32_t st_steps = 0, st_examined = 0, st_dc = 1;
    uint32_t fill = 0;                                   // simhash_fill_slots
    uint64_t w_examined = 0, w_filtered = 0, w_expansions = 0;
    int bstate = HXP_READY;
    uint32_t bremaining = 0;
    uint32_t vcount = 1;                                 // entries in the visited set
    uint32_t memo_key = 0xffffffffu, memo_thr = 0;       // adaptive threshold memo (keyed by delta's bits)
    HxSession sess{sess_mem, qsim ^ ((ix.ids[cur] << 17) | (ix.ids[cur] >> 47)) ^ (((uint64_t)a.ef << 7) | ((uint64_t)a.ef >> 57))};
    if (lane == 0) {
      const uint64_t k0 = hx_make_key(cur_dist, cur << 1);
      beam_mem[0] = k0;
      sess_mem[25] = 0;
      hx_vt_test_and_set(vt, cur);
    }
    __syncwarp();
    for (;;) {
      const uint32_t first = hx_beam_first_unexpanded(beam_mem, beam.len, lane);
      uint32_t cur_slot = HX_ABSENT;
      uint32_t cur_bits = 0;
      if (first != HX_ABSENT) {
        uint64_t key = beam_mem[first];
        __syncwarp();
        if (lane == 0) beam_mem[first] = key | 1ull;
        cur_slot = (uint32_t)(key & 0xffffffffu) >> 1;
        cur_bits = (uint32_t)(key >> 32);
        st_steps++;
      } else if (tq.len > 0) {
        const uint64_t key = hx_tie_pop(tq);
        cur_slot = (uint32_t)(key & 0xffffffffu) >> 1;
        cur_bits = (uint32_t)(key >> 32);
        st_steps++;
      } else if (dropped) {
        st_steps++;
      }
      if (cur_slot == HX_ABSENT) break;
      // -- unvisited neighbours (search.rs:583-593): a read-only probe, nothing is marked yet
      uint32_t nf = 0;
      {
        const uint32_t* row = ix.nbr0 + (size_t)cur_slot * ix.stride0;
        uint32_t nb = row[lane];
        const uint32_t deg = ix.deg0[cur_slot];
        st_examined += ix.raw0[cur_slot];
        if (vcount + deg > vt.limit) {
          if (pool_idx >= 0 || (pool_idx = hx_vt_grow_warp(vt, rg, lane)) < 0) {
            qflags |= HXF_VT_OVERFLOW;
            failed = true;
            break;
          }
        }
        const bool want_sim = METRIC == HXM_COSINE && pa.cfg.mode != HXP_OFF;   // the gate can only be active then
        for (uint32_t base = 0; base < deg; base += 32) {
          const uint32_t i = base + lane;
          if (base) nb = i < deg ? row[i] : 0u;
          // the fingerprint is requested together with the visited probe (both depend on the row only): one round trip
          uint64_t h = 0;
          bool hh = false;
          const bool early = rg.l2_spec != 0;
          if (early && want_sim && i < deg) {
            hh = pa.node_has_simhash == nullptr || pa.node_has_simhash[nb] != 0;
            h = pa.node_simhash[nb];
          }
          const bool fresh = i < deg && !hx_vt_contains(vt, nb);
          if (!early && want_sim && fresh) {
            hh = pa.node_has_simhash == nullptr || pa.node_has_simhash[nb] != 0;
            h = pa.node_simhash[nb];
          }
          const uint32_t mask = __ballot_sync(FULL, fresh);
          if (fresh) {
            const uint32_t pos = nf + __popc(mask & ((1u << lane) - 1u));
            frontier[pos] = nb;
            fsim[pos] = hh ? 64u - (uint32_t)__popcll(h ^ qsim) : 0x80000000u;   // bit 31: no fingerprint
          }
          nf += __popc(mask);
        }
      }
      __syncwarp();
      if (nf == 0) continue;
      // -- decision (search.rs:603-649)
      const bool topk_ready = beam.len >= topk_target;
      const float delta = hx_key_score(beam_mem[min(beam.len, topk_target) - 1u]);   // beam.len >= 1: the entry point
      const HxDecision dec = hxp_decide(METRIC, pa.cfg, topk_ready, a.ef, beam.len, nf, __uint_as_float(cur_bits), delta, bstate,
                                        bremaining, 0ull, w_examined, w_filtered, w_expansions, memo_key, memo_thr);
      bstate = dec.next_state;
      bremaining = dec.next_remaining;
      if (dec.trigger & 1) ps[9]++;
      if (dec.trigger & 2) ps[10]++;
      const float active_ratio = dec.samp_kind == HXP_EXHAUSTIVE ? 1.0f : dec.samp_prob;
      const uint32_t active_threshold = dec.has_threshold ? dec.threshold : 0u;
      // -- stage-0 pre-sampling (search.rs:651-678): sequential draws, in place
      uint32_t nsf = nf;
      const bool pre_enabled = dec.pre_kind != HXP_EXHAUSTIVE;
      if (pre_enabled) {
        uint32_t kept = 0, rejected = 0;
        if (lane == 0) {
          for (uint32_t i = 0; i < nf; ++i) {
            if (hxp_should_sample(sess, dec.pre_prob)) { fsim[kept] = fsim[i]; frontier[kept++] = frontier[i]; }
            else rejected++;
          }
          if (kept == 0) {   // never leave a non-empty frontier unexplored: one uniform pick (:666-672)
            const uint32_t idx = hxp_choose_index(sess, nf);
            frontier[0] = frontier[idx];
            fsim[0] = fsim[idx];
            kept = 1;
          }
        }
        kept = __shfl_sync(FULL, kept, 0);
        rejected = __shfl_sync(FULL, rejected, 0);
        ps[8] += rejected;
        ps[7] += kept;
        nsf = kept;
        __syncwarp();
      }
      if (dec.bypassed) { ps[5]++; ps[6] += nsf; }
      // -- threshold gate + sampling (search.rs:709-786)
      const bool should_sample = !pre_enabled && dec.samp_kind != HXP_EXHAUSTIVE && active_ratio > 0.0f;
      uint32_t examined_round = 0, filtered_round = 0, n_filtered_new = 0;
      for (uint32_t base = 0; base < nsf; base += 32) {
        const uint32_t f = base + lane;
        uint32_t nb = 0, sim = 32u;
        bool has_hash = false, filtered = false;
        if (f < nsf) {
          nb = frontier[f];
          const uint32_t raw = fsim[f];
          has_hash = dec.filter_cached && !(raw & 0x80000000u);
          if (has_hash) {
            sim = raw;
            filtered = sim < active_threshold;
          }
          fsim[f] = sim;
          fstate[f] = filtered ? 0 : (should_sample ? 4 : (active_ratio <= 0.0f ? 2 : 1));   // 4 = to be drawn
          if (filtered) hx_vt_test_and_set(vt, nb);   // visited.insert (always new: the frontier is unvisited and unique)
        }
        const uint32_t hm = __ballot_sync(FULL, has_hash), fm = __ballot_sync(FULL, filtered);
        examined_round += __popc(hm);
        filtered_round += __popc(fm);
        if (dec.fetch_missing) ps[2] += __popc(__ballot_sync(FULL, f < nsf && !has_hash));
      }
      n_filtered_new = filtered_round;
      vcount += n_filtered_new;
      ps[1] += examined_round;
      ps[0] += filtered_round;
      ps[3] += nsf - filtered_round;
      {   // virtual beam-fill slots (search.rs:742-748): each filtered node takes one while the effective beam is not full
        const uint32_t room = beam.len + fill < a.ef ? a.ef - (beam.len + fill) : 0u;
        fill += n_filtered_new < room ? n_filtered_new : room;
      }
      __syncwarp();
      if (should_sample) {   // sequential Bernoulli draws in neighbour order
        if (lane == 0) {
          for (uint32_t i = 0; i < nsf; ++i)
            if (fstate[i] == 4) fstate[i] = hxp_should_sample(sess, hxp_candidate_probability(dec, fsim[i])) ? 1 : 2;
        }
        __syncwarp();
      }
      if (dec.filter_cached && examined_round > 0) {   // search.rs:788-800
        w_examined += examined_round;
        w_filtered += filtered_round;
        w_expansions += 1;
        if (w_expansions > pa.cfg.bypass_window_expansions) {
          w_examined /= 2;
          w_filtered /= 2;
          w_expansions = pa.cfg.bypass_window_expansions / 2;
        }
      }
      // -- compact the sampled ones (order kept); fallback when nothing was sampled (search.rs:802-823)
      uint32_t ns = 0, nd = 0, best_sim = 0;
      for (uint32_t base = 0; base < nsf; base += 32) {
        const uint32_t f = base + lane;
        const uint32_t stt = f < nsf ? fstate[f] : 0u;
        const uint32_t sm = __ballot_sync(FULL, stt == 1u), dm = __ballot_sync(FULL, stt == 2u);
        const uint32_t nb = f < nsf ? frontier[f] : 0u;
        const uint32_t mysim = stt == 2u ? fsim[f] : 0u;
        best_sim = max(best_sim, __reduce_max_sync(FULL, mysim));
        __syncwarp();
        if (stt == 1u) frontier[ns + __popc(sm & ((1u << lane) - 1u))] = nb;   // ns + rank <= f: never overwrites unread data of later chunks
        ns += __popc(sm);
        nd += __popc(dm);
        __syncwarp();
      }
      if (active_ratio > 0.0f && ns == 0 && nd > 0) {
        // the compaction above wrote nothing (ns == 0), so frontier / fstate / fsim still describe the gate's output
        uint32_t pick = HX_ABSENT;
        if (lane == 0) {
          uint32_t nbest = 0;
          for (uint32_t i = 0; i < nsf; ++i)
            if (fstate[i] == 2 && fsim[i] == best_sim) nbest++;
          const uint32_t idx = hxp_choose_index(sess, nbest);
          uint32_t seen = 0;
          for (uint32_t i = 0; i < nsf; ++i)
            if (fstate[i] == 2 && fsim[i] == best_sim && seen++ == idx) { pick = frontier[i]; break; }
          frontier[0] = pick;
        }
        ns = 1;
        __syncwarp();
      }
      ps[4] += ns;
      // -- mark the sampled ones visited (search.rs:830-836)
      for (uint32_t f = lane; f < ns; f += 32) hx_vt_test_and_set(vt, frontier[f]);
      vcount += ns;
      st_dc += ns;
      __syncwarp();
      if (ns == 0) continue;
      // -- score and admit (search.rs:909-953)
      score_list(frontier, ns);
      // One pass decides which scores can be admitted at all: once the effective beam (entries + fill slots) is full it stays
      // full and w.max only decreases, so a score that fails `dist < w.max || len + fill < ef` now fails it at its turn too;
      // the survivors are admitted one by one in neighbour order with the test repeated on the live state (search.rs:909-953)
      for (uint32_t base = 0; base < ns; base += 32) {
       const uint32_t fl = base + lane;
       float sl = fl < ns ? fdist[fl] : 0.f;
       uint32_t sbl = 0;
       bool pass = false;
       if (fl < ns) {
         if (!hx_score_ok(sl)) qflags |= HXF_INVALID_SCORE;
         sbl = __float_as_uint(sl);
         pass = sbl < (uint32_t)(beam_mem[beam.len - 1] >> 32) || beam.len + fill < a.ef;
       }
       const uint32_t slot_l = fl < ns ? frontier[fl] : 0u;
       uint32_t pmask = __ballot_sync(FULL, pass);
       while (pmask) {
        const int src = __ffs(pmask) - 1;
        pmask &= pmask - 1;
        const uint32_t sbits = __shfl_sync(FULL, sbl, src);
        const uint32_t xslot = __shfl_sync(FULL, slot_l, src);
        const uint32_t wmax = (uint32_t)(beam_mem[beam.len - 1] >> 32);
        if (!(sbits < wmax || beam.len + fill < a.ef)) continue;
        const uint64_t nkey = ((uint64_t)sbits << 32) | ((uint64_t)xslot << 1);
        const bool was_full = beam.len == a.ef;
        const uint32_t old_wmax = wmax;
        uint64_t ev;
        hx_beam_insert2(beam, a.ef, nkey, &ev, lane);
        if (!was_full && beam.len + fill > a.ef) fill--;   // a real candidate replaced a virtual fill slot (:940-944)
        if (lane == 0) {
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
      if (sess_mem[25]) {
        const uint64_t blk = ((uint64_t)sess_mem[27] << 32) | sess_mem[26];
        ps[11] += (uint32_t)(blk * 16ull - (16ull - sess_mem[24]));
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
  if (CTA) {   // warp 0: release the helpers
    if (lane == 0) {
      s_cmd = 2u;
      hx_mbar_arrive(&s_bar_cmd);
    }
  }
  if (pa.pstats && lane == 0)
    for (int i = 0; i < 12; ++i)
      if (ps[i]) atomicAdd(pa.pstats + i, (unsigned long long)ps[i]);
}
