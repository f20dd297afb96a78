// k_filtered.cuh — filter-aware (ACORN-style) restricted search on the device.
//
// Restates VectorIndex::restricted_filter_aware_search (search/vector/restricted.rs:837-1148) for generations without the
// SimHash routing directory: the walk the reference takes when |C| > 256 (restricted.rs:426-453).  Seeds = the evenly
// spaced sample of the candidate set (:321-342) + the entry point when it is a candidate; layer-0 rows of the best scored
// candidates are "routed" 16 at a time (:1007-1041); neighbours that are NOT candidates become BRIDGES, ranked by the
// Hamming distance between their SimHash and the query's (:706-751) and expanded up to 256 at a time without ever
// touching their vectors (:1051-1093); at most 800 candidate vectors are scored in total (:55, budgets :232-260).
//
// One CTA per query.  Every container of the reference keeps its order semantics:
//   frontier  BinaryHeap<Reverse<Candidate>>  -> ascending key array F in shared memory (key = score bits << 32 | slot;
//                                               slots ascend with ids, so key order == (score, id) order)
//   top       BinaryHeap<Candidate> (<= ef_filtered)  -> T = ALL scored keys ascending; top = T[0 .. ef_filtered)
//   bridges   BinaryHeap<Reverse<(hamming, id)>>  -> unsorted key array in global memory; a pop of m entries is an exact
//                                               selection of the m smallest (histogram over the 65 Hamming values, then a
//                                               radix select over the slot inside the cut-off bin), sorted before use
//   attempted / scored / expanded / queued / eligible_seen (HashSet<NodeId>) -> flag bits in a per-CTA stamp array
//                                               indexed by slot, (epoch << 8 | flags): no clearing between queries
// The eligible list is appended in discovery order (batch order, row order) by one warp with ballot compaction, because
// `eligible.truncate(..)` (:1100) makes that order part of the answer.  Scores use the same bit-exact octet kernels as every
// other path.  The result equals the oracle's restatement id for id and bit for bit (tests/test_gpu_filtered.py).
#pragma once
#include "hx_common.cuh"

#define HXG_THREADS 256
#define HXG_MAX_SCORED 800u        // FILTERED_VECTOR_PAYLOAD_LIMIT (restricted.rs:55)
#define HXG_FRONTIER_BATCH 16u     // FRONTIER_BATCH_SIZE
#define HXG_BRIDGE_BATCH 256u      // BRIDGE_BATCH_SIZE
#define HXG_F_ATT 1u
#define HXG_F_SCORED 2u
#define HXG_F_EXPANDED 4u
#define HXG_F_QUEUED 8u
#define HXG_F_ELIG 16u
#define HXG_ERR_MISSING_SIMHASH 32u   // a graph neighbour / entry point without its mandatory SimHash companion (:737-742)
#define HXG_ERR_CAPACITY 64u          // internal list capacity exceeded (cannot happen for validated budgets)

struct HxFilteredArgs {
  const float* queries;       // [B][dim]
  const float* q_hdr;         // [B]
  const uint32_t* q_status;   // [B]
  const uint64_t* q_simhash;  // [B]
  uint32_t B, k;
  // candidate set (shared by the B queries): membership bitmap over slots + the seeds resolved on the host
  const uint32_t* allowed_bits;   // [ceil(n/32)]
  const uint32_t* seed_slots;     // [n_seed_att] sampled seeds (HX_ABSENT: id without a vector row); all are "attempted"
  uint32_t n_seed_att;
  const uint32_t* init_slots;     // [n_init] what is scored first (sample + entry point, resolved, truncated to the budget)
  uint32_t n_init;
  uint32_t entry_allowed;
  // budgets (FilteredGraphBudgets, restricted.rs:220-260)
  uint32_t ef_filtered, routing_rows, bridge_rows, vector_payloads;
  // node fingerprints
  const uint64_t* simhash;
  const uint8_t* has_simhash;     // nullptr: every node has one
  // per-CTA scratch
  uint32_t* stamps;               // [grid][n]
  uint32_t* epochs;               // [grid]
  uint64_t* bridge;               // [grid][2][bridge_cap]  (double buffered for the compaction after a pop)
  uint32_t bridge_cap;
  uint32_t* elig;                 // [grid][elig_cap]
  uint32_t elig_cap;
  uint32_t* counter;              // query ticket
  // outputs
  uint64_t* out_ids;
  float* out_scores;
  uint32_t* out_counts;
  uint32_t* q_err;                // [B] per-query error flags
  unsigned long long* stats;      // [8]: payload requests, distance computations, routing rows, bridge rows, bridge pushes,
                                  //      iterations, terminations (packed counts are not needed: summed per call)
};

// bitonic sort of `n` (<= cap, cap a power of two) u64 keys in shared memory, ascending; padding with HX_KEY_MAX
__device__ __forceinline__ void hxg_sort_smem(uint64_t* s, uint32_t cap, uint32_t tid) {
  for (uint32_t size = 2; size <= cap; size <<= 1)
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (uint32_t p = tid; p < cap / 2; p += HXG_THREADS) {
        const uint32_t i = 2 * p - (p & (stride - 1));
        const uint32_t j = i + stride;
        const bool up = ((i & size) == 0);
        const uint64_t x = s[i], y = s[j];
        if ((x > y) == up) { s[i] = y; s[j] = x; }
      }
    }
  __syncthreads();
}

// merge the sorted run `add[0..na)` into the sorted array `dst[0..nd)` (both ascending, keys distinct), through `tmp`
__device__ __forceinline__ void hxg_merge_sorted(uint64_t* dst, uint32_t nd, const uint64_t* add, uint32_t na, uint64_t* tmp,
                                                 uint32_t tid) {
  for (uint32_t i = tid; i < nd; i += HXG_THREADS) {          // old element i moves up by #(add < dst[i])
    const uint64_t x = dst[i];
    uint32_t lo = 0, hi = na;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (add[mid] < x) lo = mid + 1; else hi = mid; }
    tmp[i + lo] = x;
  }
  for (uint32_t j = tid; j < na; j += HXG_THREADS) {          // new element j lands at j + #(dst < add[j])
    const uint64_t x = add[j];
    uint32_t lo = 0, hi = nd;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (dst[mid] < x) lo = mid + 1; else hi = mid; }
    tmp[j + lo] = x;
  }
  __syncthreads();
  for (uint32_t i = tid; i < nd + na; i += HXG_THREADS) dst[i] = tmp[i];
  __syncthreads();
}

template <int METRIC>
__global__ void __launch_bounds__(HXG_THREADS) k_filtered_walk(HxDev ix, HxFilteredArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* sq = reinterpret_cast<float*>(smem);                                   // [ld]
  uint64_t* F = reinterpret_cast<uint64_t*>(smem + (size_t)ix.ld * 4);          // [HXG_MAX_SCORED] frontier (ascending)
  uint64_t* T = F + HXG_MAX_SCORED;                                             // [HXG_MAX_SCORED] every scored key (ascending)
  uint64_t* tmp = T + HXG_MAX_SCORED;                                           // [HXG_MAX_SCORED] merge scratch
  uint64_t* newk = tmp + HXG_MAX_SCORED;                                        // [1024] new keys of one scoring batch (sorted)
  uint64_t* bbatch = newk + 1024;                                               // [256] bridge batch keys
  uint32_t* rbatch = reinterpret_cast<uint32_t*>(bbatch + HXG_BRIDGE_BATCH);    // [16] routing batch slots
  __shared__ uint32_t s_hist[72];
  __shared__ uint32_t s_n, s_m, s_cut, s_need, s_cnt, s_cnt2, s_flag, s_prefix, s_take;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, t = tid & 7u, oct = tid >> 3;
  const unsigned FULL = 0xffffffffu;
  volatile uint32_t* stamp = a.stamps + (size_t)blockIdx.x * ix.n;   // volatile: flags written by one lane are read by others
  uint64_t* BRa = a.bridge + (size_t)blockIdx.x * 2 * a.bridge_cap;
  uint32_t* EL = a.elig + (size_t)blockIdx.x * a.elig_cap;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_n = atomicAdd(a.counter, 1u);
    __syncthreads();
    const uint32_t qi = s_n;
    if (qi >= a.B) break;
    if (a.q_status[qi] != 0u || !ix.populated) {
      if (tid == 0) { a.out_counts[qi] = 0; a.q_err[qi] = 0; }
      continue;
    }
    // epoch of this query in the stamp array (flags of older epochs read as 0)
    if (tid == 0) {
      uint32_t e = a.epochs[blockIdx.x] + 1u;
      if (e >= (1u << 24)) e = 0xffffffffu;   // wrap: clear below
      s_flag = e;
    }
    __syncthreads();
    uint32_t epoch = s_flag;
    if (epoch == 0xffffffffu) {
      for (uint32_t i = tid; i < ix.n; i += HXG_THREADS) stamp[i] = 0u;
      epoch = 1u;
    }
    __syncthreads();
    if (tid == 0) a.epochs[blockIdx.x] = epoch;
    const uint32_t ebits = epoch << 8;
    auto flags_of = [&](uint32_t slot) -> uint32_t {
      const uint32_t v = stamp[slot];
      return (v >> 8) == epoch ? (v & 0xffu) : 0u;
    };
    auto set_flag = [&](uint32_t slot, uint32_t f) { stamp[slot] = ebits | (flags_of(slot) | f); };
    const float q_hdr = a.q_hdr[qi];
    const uint64_t qsim = a.q_simhash[qi];
    for (uint32_t i = tid; i < ix.ld; i += HXG_THREADS) sq[i] = i < ix.dim ? a.queries[(size_t)qi * ix.dim + i] : 0.0f;
    __syncthreads();
    uint32_t nF = 0, nT = 0, nBR = 0, cur = 0;   // cur: which half of the bridge double buffer is live
    uint32_t st_payload = 0, st_routing = 0, st_bridge = 0, st_pushes = 0, st_iters = 0;
    uint32_t qerr = 0;
    uint64_t* BR = BRa;

    // score the slots list[0..cnt) (cnt <= 1024; not yet scored), push into frontier and top (:661-704)
    auto score_batch = [&](const uint32_t* list, uint32_t cnt) {
      for (uint32_t i = tid; i < 1024; i += HXG_THREADS) newk[i] = HX_KEY_MAX;
      __syncthreads();
      for (uint32_t f = oct; f < cnt; f += HXG_THREADS / 8) {
        const uint32_t slot = list[f];
        float s;
        if (METRIC == HXM_MANHATTAN) s = hx_manhattan_seq(ix.vec + (size_t)slot * ix.ld, sq, ix.dim);
        else s = hx_octet_score<METRIC>(ix, sq, q_hdr, slot, t);
        if (t == 0) {
          if (!hx_score_ok(s)) atomicOr(&s_flag, HXF_INVALID_SCORE);
          newk[f] = hx_make_key(s, slot);
          stamp[slot] = ebits | (flags_of(slot) | HXG_F_SCORED);
        }
      }
      __syncthreads();
      uint32_t cap = 32;
      while (cap < cnt) cap <<= 1;
      hxg_sort_smem(newk, cap, tid);
      hxg_merge_sorted(F, nF, newk, cnt, tmp, tid);
      hxg_merge_sorted(T, nT, newk, cnt, tmp, tid);
      nF += cnt;
      nT += cnt;
      st_payload += cnt;
    };

    // classify the layer-0 rows of `nodes[0..cnt)` in order (:1028-1041, :1072-1083): candidates not yet attempted join
    // the eligible list once (discovery order), everything else becomes a bridge once.  One warp, ballot compaction.
    // Rows are independent to FETCH but ordered to COMMIT (eligible order is part of the answer): the eight warps take the
    // rows round-robin, each prefetches its row's first 32 neighbours — the row itself, their candidate bits, their flags
    // (to warm L2) and fingerprints: the ~1.5 us of dependent latency — and then commits in row order behind a turn
    // counter in shared memory; the commit re-reads the flags (an earlier row of the same batch may have claimed the slot).
    uint32_t nEL = 0;
    auto classify = [&](const uint32_t* nodes, bool keys64, const uint64_t* nodes64, uint32_t cnt) {
      __shared__ volatile uint32_t s_turn, s_el, s_br, s_push;
      if (tid == 0) { s_turn = 0; s_el = nEL; s_br = nBR; s_push = 0; }
      __syncthreads();
      uint32_t err = 0;
      for (uint32_t r = warp; r < cnt; r += HXG_THREADS / 32) {
        const uint32_t node = keys64 ? (uint32_t)(nodes64[r] & 0xffffffffu) : nodes[r];
        const uint32_t deg = ix.deg0[node];
        const uint32_t* row = ix.nbr0 + (size_t)node * ix.stride0;
        // ---- prefetch phase (runs concurrently in all warps)
        uint32_t nb0 = 0;
        bool allowed0 = false;
        uint64_t sim0 = 0;
        bool has0 = true;
        if (lane < deg) {
          nb0 = row[lane];
          allowed0 = (a.allowed_bits[nb0 >> 5] >> (nb0 & 31u)) & 1u;
          (void)stamp[nb0];
          if (!allowed0) {
            sim0 = a.simhash[nb0];
            has0 = a.has_simhash ? a.has_simhash[nb0] != 0 : true;
          }
        }
        // ---- commit phase, in row order
        if (lane == 0)
          while (s_turn != r) __nanosleep(40);
        __syncwarp();
        uint32_t el = s_el, br = s_br, pushes = 0;
        for (uint32_t base = 0; base < deg; base += 32) {
          const uint32_t j = base + lane;
          uint32_t nb = nb0, fl = 0;
          bool allowed = allowed0, has = has0, is_el = false, is_br = false;
          uint64_t sim = sim0;
          if (j < deg) {
            if (base) {
              nb = row[j];
              allowed = (a.allowed_bits[nb >> 5] >> (nb & 31u)) & 1u;
              if (!allowed) {
                sim = a.simhash[nb];
                has = a.has_simhash ? a.has_simhash[nb] != 0 : true;
              }
            }
            fl = flags_of(nb);
            if (allowed) is_el = !(fl & (HXG_F_ATT | HXG_F_ELIG));
            else is_br = !(fl & HXG_F_QUEUED);
          }
          // a row is ascending and unique, so lanes of one chunk never name the same slot: flags can be set in parallel
          if (is_el) stamp[nb] = ebits | (fl | HXG_F_ELIG);
          if (is_br) stamp[nb] = ebits | (fl | HXG_F_QUEUED);
          const uint32_t me = __ballot_sync(FULL, is_el), mb = __ballot_sync(FULL, is_br);
          if (is_el) {
            const uint32_t p = el + __popc(me & ((1u << lane) - 1u));
            if (p < a.elig_cap) EL[p] = nb; else err |= HXG_ERR_CAPACITY;
          }
          if (is_br) {
            const uint32_t p = br + __popc(mb & ((1u << lane) - 1u));
            if (!has) err |= HXG_ERR_MISSING_SIMHASH;
            if (p < a.bridge_cap) BR[p] = ((uint64_t)__popcll(sim ^ qsim) << 32) | nb; else err |= HXG_ERR_CAPACITY;
          }
          el += __popc(me);
          br += __popc(mb);
          pushes += __popc(mb);
          __syncwarp();
        }
        __threadfence_block();
        if (lane == 0) {
          s_el = el;
          s_br = br;
          s_push = s_push + pushes;
          __threadfence_block();
          s_turn = r + 1u;
        }
        __syncwarp();
      }
      err = __reduce_or_sync(FULL, err);
      if (lane == 0 && err) atomicOr(&s_flag, err);
      __syncthreads();
      nEL = min((uint32_t)s_el, a.elig_cap);
      nBR = min((uint32_t)s_br, a.bridge_cap);
      st_pushes += s_push;
      __syncthreads();
    };

    // ---- seeds (:926-963)
    if (tid == 0) s_flag = 0;
    __syncthreads();
    for (uint32_t i = tid; i < a.n_seed_att; i += HXG_THREADS) {
      const uint32_t sl = a.seed_slots[i];
      if (sl != HX_ABSENT) stamp[sl] = ebits | HXG_F_ATT;   // seeds are distinct: no read-modify-write race
    }
    __syncthreads();
    if (a.entry_allowed && tid == 0) set_flag(ix.entry_slot, HXG_F_ATT);
    __syncthreads();
    if (a.n_init) score_batch(a.init_slots, a.n_init);
    if (!a.entry_allowed) {   // the entry point is not a candidate: it is the first bridge (:984-993)
      if (tid == 0) {
        const uint32_t e = ix.entry_slot;
        const bool has = a.has_simhash ? a.has_simhash[e] != 0 : true;
        if (!has) atomicOr(&s_flag, HXG_ERR_MISSING_SIMHASH);
        set_flag(e, HXG_F_QUEUED);
        BR[0] = ((uint64_t)__popcll(a.simhash[e] ^ qsim) << 32) | e;
      }
      nBR = 1;
      st_pushes = 1;
      __syncthreads();
    }

    // ---- the walk (:995-1133)
    for (;;) {
      __syncthreads();
      if (s_flag) break;                                                   // an error ends this query
      if (st_payload >= a.vector_payloads) break;                          // VectorBudget
      if (nBR == 0 && nT >= a.ef_filtered && nF > 0 && F[0] > T[a.ef_filtered - 1u]) break;   // BeamComplete
      st_iters++;
      // routing batch: the best scored candidates not yet expanded (every scored node enters the frontier exactly once,
      // so `expanded.insert` never fails; the flag is kept for the stats and for symmetry with the reference)
      uint32_t nr = min(nF, HXG_FRONTIER_BATCH);
      for (uint32_t i = tid; i < nr; i += HXG_THREADS) {
        const uint32_t slot = (uint32_t)(F[i] & 0xffffffffu);
        rbatch[i] = slot;
        stamp[slot] = ebits | (flags_of(slot) | HXG_F_EXPANDED);
      }
      __syncthreads();
      if (nr) {   // pop: shift the frontier
        for (uint32_t i = tid; i + nr < nF; i += HXG_THREADS) tmp[i] = F[i + nr];
        __syncthreads();
        for (uint32_t i = tid; i + nr < nF; i += HXG_THREADS) F[i] = tmp[i];
        nF -= nr;
        __syncthreads();
      }
      if (nr == 0 && nBR == 0) break;                                      // Exhausted
      uint32_t routing_remaining = a.routing_rows > st_routing ? a.routing_rows - st_routing : 0u;
      if (routing_remaining == 0) break;                                   // RoutingBudget
      if (nr > routing_remaining) nr = routing_remaining;
      nEL = 0;
      if (nr) {
        st_routing += nr;
        routing_remaining -= nr;
        classify(rbatch, false, nullptr, nr);
      }
      // bridge batch (:1051-1062): the m smallest (hamming, id) of the bridge frontier, in order
      const uint32_t bridge_remaining = a.bridge_rows > st_bridge ? a.bridge_rows - st_bridge : 0u;
      uint32_t m = min(min(bridge_remaining, routing_remaining), min(HXG_BRIDGE_BATCH, nBR));
      if (m) {
        // 1. histogram over the Hamming distance (0..64)
        for (uint32_t i = tid; i < 72; i += HXG_THREADS) s_hist[i] = 0;
        if (tid == 0) { s_cnt = 0; s_cnt2 = 0; }
        __syncthreads();
        for (uint32_t i = tid; i < nBR; i += HXG_THREADS) atomicAdd(&s_hist[(uint32_t)(BR[i] >> 32)], 1u);
        __syncthreads();
        if (tid == 0) {
          uint32_t acc = 0, h = 0;
          for (; h < 65; ++h) { if (acc + s_hist[h] >= m) break; acc += s_hist[h]; }
          s_cut = h;            // cut-off Hamming value
          s_need = m - acc;     // how many of the cut-off bin are taken (the smallest slots)
          s_m = s_hist[h];      // size of the cut-off bin
        }
        __syncthreads();
        const uint32_t cut = s_cut, need = s_need, binsz = s_m;
        // 2. inside the cut-off bin: the `need`-th smallest slot by a 4-pass radix select (only when the bin is not taken whole)
        uint32_t thr_slot = 0xffffffffu;   // take slots <= thr_slot of the cut-off bin
        if (need < binsz) {
          uint32_t prefix = 0, kth = need;   // kth (1-based) smallest
          for (int pass = 3; pass >= 0; --pass) {
            const uint32_t shift = (uint32_t)pass * 8u;
            __shared__ uint32_t s_h256[256];
            for (uint32_t i = tid; i < 256; i += HXG_THREADS) s_h256[i] = 0;
            __syncthreads();
            const uint32_t himask = pass == 3 ? 0u : (0xffffffffu << (shift + 8u));
            for (uint32_t i = tid; i < nBR; i += HXG_THREADS) {
              const uint64_t key = BR[i];
              const uint32_t slot = (uint32_t)(key & 0xffffffffu);
              if ((uint32_t)(key >> 32) == cut && (slot & himask) == (prefix & himask)) atomicAdd(&s_h256[(slot >> shift) & 0xffu], 1u);
            }
            __syncthreads();
            if (tid == 0) {
              uint32_t acc = 0, b = 0;
              for (; b < 256; ++b) { if (acc + s_h256[b] >= kth) break; acc += s_h256[b]; }
              s_prefix = prefix | (b << shift);
              s_take = kth - acc;
            }
            __syncthreads();
            prefix = s_prefix;
            kth = s_take;
            __syncthreads();
          }
          thr_slot = prefix;   // slots are distinct, so exactly `need` bin members are <= prefix
        }
        // 3. move the selected keys to the batch, the rest to the other half of the double buffer
        uint64_t* BRn = BRa + (size_t)(cur ^ 1u) * a.bridge_cap;
        for (uint32_t i = tid; i < HXG_BRIDGE_BATCH; i += HXG_THREADS) bbatch[i] = HX_KEY_MAX;
        __syncthreads();
        for (uint32_t i = tid; i < nBR; i += HXG_THREADS) {
          const uint64_t key = BR[i];
          const uint32_t ham = (uint32_t)(key >> 32), slot = (uint32_t)(key & 0xffffffffu);
          const bool sel = ham < cut || (ham == cut && slot <= thr_slot);
          if (sel) bbatch[atomicAdd(&s_cnt, 1u)] = key;
          else BRn[atomicAdd(&s_cnt2, 1u)] = key;
        }
        __syncthreads();
        nBR = s_cnt2;
        cur ^= 1u;
        BR = BRn;
        hxg_sort_smem(bbatch, HXG_BRIDGE_BATCH, tid);     // pop order: ascending (hamming, id)
        st_routing += m;
        st_bridge += m;
        classify(nullptr, true, bbatch, m);
      } else if (nr == 0 && nBR != 0) {
        break;                                                             // BridgeBudget
      }
      const uint32_t vector_remaining = a.vector_payloads > st_payload ? a.vector_payloads - st_payload : 0u;
      if (vector_remaining == 0) break;                                    // VectorBudget
      const uint32_t take = min(min(vector_remaining, a.ef_filtered), nEL);
      // eligible_seen lives for one iteration: drop the flag of every discovered node, mark the taken ones attempted
      for (uint32_t i = tid; i < nEL; i += HXG_THREADS) {
        const uint32_t slot = EL[i];
        const uint32_t fl = flags_of(slot) & ~HXG_F_ELIG;
        stamp[slot] = ebits | (i < take ? (fl | HXG_F_ATT) : fl);
      }
      __syncthreads();
      if (take == 0) continue;
      score_batch(EL, take);
    }
    __syncthreads();
    qerr = s_flag;
    // ---- results: the k smallest of everything scored (:1135-1147)
    const uint32_t kk = a.k < nT ? a.k : nT;
    const uint32_t cnt = qerr ? 0u : kk;
    for (uint32_t i = tid; i < cnt; i += HXG_THREADS) {
      a.out_ids[(size_t)qi * a.k + i] = ix.ids[(uint32_t)(T[i] & 0xffffffffu)];
      a.out_scores[(size_t)qi * a.k + i] = hx_key_score(T[i]);
    }
    if (tid == 0) {
      a.out_counts[qi] = cnt;
      a.q_err[qi] = qerr;
      if (a.stats) {
        atomicAdd(a.stats + 0, (unsigned long long)st_payload);
        atomicAdd(a.stats + 1, (unsigned long long)st_payload);
        atomicAdd(a.stats + 2, (unsigned long long)st_routing);
        atomicAdd(a.stats + 3, (unsigned long long)st_bridge);
        atomicAdd(a.stats + 4, (unsigned long long)st_pushes);
        atomicAdd(a.stats + 5, (unsigned long long)st_iters);
      }
    }
  }
}
