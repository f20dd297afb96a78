// hx_filtered.inl — host side of the filter-aware (ACORN-style) restricted search (included by hx_api.cu).
// restricted.rs:426-453 plans this walk for |C| > 256 (or > 4 MiB of candidate vectors); the device library offers it next to
// the exact scan (hx_search_restricted): the scan is exact and HBM-bound in |C|, the walk scores at most 800 vectors.

static __global__ void k_set_bits(const uint32_t* __restrict__ slots, uint64_t n, uint32_t* __restrict__ bits) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = slots[i];
  if (s != HX_ABSENT) atomicOr(bits + (s >> 5), 1u << (s & 31u));
}

extern "C" void hx_filtered_budgets(uint32_t k, uint32_t ef, uint32_t beam_percent, uint64_t n_cand, hx_filtered_budgets_t* b) {
  // FilteredGraphBudgets::with_beam_percent (restricted.rs:232-260)
  if (!b) return;
  const uint64_t kk = std::min<uint64_t>(k, n_cand);
  if (beam_percent == 0) beam_percent = 150;
  if (ef == 0) ef = std::max(k, 100u);
  uint64_t ef_f = (uint64_t)ef * beam_percent / 100;
  ef_f = std::max<uint64_t>(ef_f, kk * 4);
  ef_f = std::min<uint64_t>(ef_f, n_cand);
  b->ef_filtered = (uint32_t)ef_f;
  b->routing_rows = (uint32_t)std::min<uint64_t>(ef_f * 16, 0xffffffffull);
  b->bridge_rows = (uint32_t)std::min<uint64_t>(ef_f * 8, 0xffffffffull);
  b->vector_payloads = (uint32_t)std::min<uint64_t>(800, n_cand);
  b->sampled_seeds = (uint32_t)std::min<uint64_t>(64, n_cand);
}

extern "C" hx_status hx_search_filtered_graph(hx_index* ix, const float* queries, size_t B, const hx_search_params* p,
                                              const hx_filtered_budgets_t* budgets_or_null, const uint64_t* cand_ids,
                                              size_t n_cand, const uint64_t* query_simhash, uint64_t* out_ids,
                                              float* out_scores, uint32_t* out_counts, hx_filtered_stats* stats) {
  if (!ix) {
    hx_set_error("null index handle");
    return HX_ERR_INDEX_NOT_FOUND;
  }
  uint32_t k, ef;
  hx_status rc = check_params(ix, p, &k, &ef);
  if (rc) return rc;
  if (stats) memset(stats, 0, sizeof(*stats));
  if (B == 0) return HX_OK;
  if (!queries || !out_ids || !out_scores || !out_counts || (n_cand && !cand_ids)) return HX_ERR_INVALID_PARAMETER;
  if (n_cand > 1000000ull) {
    hx_set_error("restricted vector search accepts at most 1000000 unique candidates");
    return HX_ERR_QUERY;
  }
  if (n_cand == 0) {   // RestrictedVectorCandidates::Empty => Ok(vec![]) before any I/O (restricted.rs:539-541)
    for (size_t b = 0; b < B; ++b) out_counts[b] = 0;
    return HX_OK;
  }
  if (std::min<uint64_t>(k, n_cand) > 800) {
    hx_set_error("restricted vector search result count must be at most 800, got %llu",
                 (unsigned long long)std::min<uint64_t>(k, n_cand));
    return HX_ERR_QUERY;
  }
  for (size_t i = 1; i < n_cand; ++i)
    if (cand_ids[i] <= cand_ids[i - 1]) {
      hx_set_error("candidate ids must be ascending and unique (RoaringTreemap iteration order)");
      return HX_ERR_INVALID_PARAMETER;
    }
  HX_CUDA(cudaSetDevice(ix->device));
  if (ix->n == 0 || !ix->populated) return answer_empty_index(ix, queries, B, out_counts, nullptr);
  if ((rc = hx_finalize_graph(ix))) return rc;
  if (!ix->d_nbr0) return answer_empty_index(ix, queries, B, out_counts, nullptr);
  if (!ix->d_simhash) {
    hx_set_error("SimHash rows not loaded: the filtered walk ranks its bridges by SimHash (hx_index_load_simhash / "
                 "hx_index_compute_simhash)");
    return HX_ERR_INVALID_VECTOR_CONFIG;
  }
  if (!query_simhash && !ix->d_planes_t) {
    hx_set_error("no query fingerprints given and no hyperplanes set (hx_index_set_simhash_planes)");
    return HX_ERR_INVALID_VECTOR_CONFIG;
  }
  hx_filtered_budgets_t bud;
  if (budgets_or_null) bud = *budgets_or_null;
  else hx_filtered_budgets(k, ef, 0, n_cand, &bud);
  if (bud.vector_payloads > HXG_MAX_SCORED || bud.ef_filtered == 0) {
    hx_set_error("filtered-graph budgets: vector_payloads must be <= %u and ef_filtered > 0", HXG_MAX_SCORED);
    return HX_ERR_INVALID_PARAMETER;
  }
  // seeds (restricted.rs:926-963): the evenly spaced sample, then the entry point when it is a candidate
  std::vector<uint32_t> seed_slots, init_slots;
  {
    const size_t ns = std::min<size_t>(bud.sampled_seeds, n_cand);
    std::vector<uint64_t> sample(ns);
    if (ns == n_cand) {
      for (size_t i = 0; i < ns; ++i) sample[i] = cand_ids[i];
    } else if (ns == 1) {
      sample[0] = cand_ids[0];
    } else if (ns > 1) {
      const unsigned __int128 last = (unsigned __int128)(n_cand - 1);
      for (size_t sidx = 0; sidx < ns; ++sidx) sample[sidx] = cand_ids[(size_t)((unsigned __int128)sidx * last / (ns - 1))];
    }
    bool entry_sampled = false;
    for (uint64_t id : sample) {
      uint32_t slot = HX_ABSENT;
      if (hx_slot_of(ix, id, &slot)) init_slots.push_back(slot);
      seed_slots.push_back(slot);
      if (id == ix->entry_id) entry_sampled = true;
    }
    const bool entry_allowed = std::binary_search(cand_ids, cand_ids + n_cand, ix->entry_id);
    if (entry_allowed && !entry_sampled) init_slots.push_back(ix->entry_slot);
    if (init_slots.size() > bud.vector_payloads) init_slots.resize(bud.vector_payloads);
  }
  const bool entry_allowed = std::binary_search(cand_ids, cand_ids + n_cand, ix->entry_id);

  HxScratch* s = nullptr;
  if ((rc = hx_acquire_scratch(ix, &s))) return rc;
  ScratchGuard guard{ix, s};
  uint32_t launches = 0;
  if ((rc = stage_queries(ix, s, queries, B, &launches))) return rc;
  const uint32_t grid = (uint32_t)std::min<size_t>(B, (size_t)ix->sm_count);
  const uint32_t words = (uint32_t)((ix->n + 31) / 32);
  const uint64_t br_need = std::min<uint64_t>((uint64_t)ix->n, (uint64_t)bud.routing_rows * ix->stride0) + 64;
  const uint32_t bridge_cap = (uint32_t)std::min<uint64_t>(br_need, (uint64_t)ix->n + 64);
  const uint32_t elig_cap = (HXG_FRONTIER_BATCH + HXG_BRIDGE_BATCH) * ix->stride0;
  if ((rc = s->d_cand_ids.reserve(n_cand)) || (rc = s->d_cand_slots.reserve(n_cand)) || (rc = s->d_fg_bits.reserve(words)) ||
      (rc = s->d_fg_seed.reserve(seed_slots.size() + init_slots.size() + 2)) || (rc = s->d_qsim.reserve(B)) ||
      (rc = s->d_fg_bridge.reserve((size_t)grid * 2 * bridge_cap)) || (rc = s->d_fg_elig.reserve((size_t)grid * elig_cap)) ||
      (rc = s->d_out_ids.reserve(B * (size_t)k)) || (rc = s->d_out_scores.reserve(B * (size_t)k)) ||
      (rc = s->d_out_counts.reserve(B)) || (rc = s->d_qerr.reserve(B)) || (rc = s->d_pstats.reserve(12)) ||
      (rc = s->d_err.reserve(4)))
    return rc;
  // stamps: [grid][n], zero only when (re)allocated — epochs make older flags invisible
  if (s->fg_stamp_rows != ix->n || s->fg_stamp_grid < grid || !s->d_fg_stamps.p) {
    if ((rc = s->d_fg_stamps.reserve((size_t)grid * ix->n)) || (rc = s->d_fg_epochs.reserve(grid))) return rc;
    HX_CUDA(cudaMemsetAsync(s->d_fg_stamps.p, 0, (size_t)grid * ix->n * sizeof(uint32_t), s->stream));
    HX_CUDA(cudaMemsetAsync(s->d_fg_epochs.p, 0, grid * sizeof(uint32_t), s->stream));
    s->fg_stamp_rows = ix->n;
    s->fg_stamp_grid = grid;
  }
  HX_CUDA(cudaMemcpyAsync(s->d_cand_ids.p, cand_ids, n_cand * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
  k_map_candidates<<<(unsigned)((n_cand + 255) / 256), 256, 0, s->stream>>>(ix->d_ids, (uint32_t)ix->n, s->d_cand_ids.p, n_cand,
                                                                            s->d_cand_slots.p, ix->contiguous ? 1 : 0,
                                                                            ix->first_id, ix->d_deleted);
  HX_CUDA(cudaMemsetAsync(s->d_fg_bits.p, 0, words * sizeof(uint32_t), s->stream));
  k_set_bits<<<(unsigned)((n_cand + 255) / 256), 256, 0, s->stream>>>(s->d_cand_slots.p, n_cand, s->d_fg_bits.p);
  launches += 2;
  std::vector<uint32_t> seeds_host(seed_slots);
  seeds_host.insert(seeds_host.end(), init_slots.begin(), init_slots.end());
  if (!seeds_host.empty())
    HX_CUDA(cudaMemcpyAsync(s->d_fg_seed.p, seeds_host.data(), seeds_host.size() * sizeof(uint32_t), cudaMemcpyHostToDevice,
                            s->stream));
  if (query_simhash) {
    HX_CUDA(cudaMemcpyAsync(s->d_qsim.p, query_simhash, B * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
  } else {
    k_simhash_project<<<(unsigned)((B + 7) / 8), 512, 0, s->stream>>>(s->d_queries.p, B, ix->cfg.dimension, ix->cfg.dimension,
                                                                     ix->d_planes_t, s->d_qsim.p, nullptr);
    launches++;
  }
  HX_CUDA(cudaMemsetAsync(s->d_err.p, 0, 2 * sizeof(uint32_t), s->stream));
  HX_CUDA(cudaMemsetAsync(s->d_pstats.p, 0, 12 * sizeof(unsigned long long), s->stream));
  HxFilteredArgs a{};
  a.queries = s->d_queries.p;
  a.q_hdr = s->d_qhdr.p;
  a.q_status = s->d_qstatus.p;
  a.q_simhash = s->d_qsim.p;
  a.B = (uint32_t)B;
  a.k = k;
  a.allowed_bits = s->d_fg_bits.p;
  a.seed_slots = s->d_fg_seed.p;
  a.n_seed_att = (uint32_t)seed_slots.size();
  a.init_slots = s->d_fg_seed.p + seed_slots.size();
  a.n_init = (uint32_t)init_slots.size();
  a.entry_allowed = entry_allowed ? 1u : 0u;
  a.ef_filtered = bud.ef_filtered;
  a.routing_rows = bud.routing_rows;
  a.bridge_rows = bud.bridge_rows;
  a.vector_payloads = bud.vector_payloads;
  a.simhash = ix->d_simhash;
  a.has_simhash = ix->simhash_count == ix->n ? nullptr : ix->d_has_simhash;
  a.stamps = s->d_fg_stamps.p;
  a.epochs = s->d_fg_epochs.p;
  a.bridge = s->d_fg_bridge.p;
  a.bridge_cap = bridge_cap;
  a.elig = s->d_fg_elig.p;
  a.elig_cap = elig_cap;
  a.counter = s->d_err.p + 1;
  a.out_ids = s->d_out_ids.p;
  a.out_scores = s->d_out_scores.p;
  a.out_counts = s->d_out_counts.p;
  a.q_err = s->d_qerr.p;
  a.stats = s->d_pstats.p;
  const HxDev dev = ix->dev();
  const size_t smem = (size_t)ix->ld * 4 + 3 * (size_t)HXG_MAX_SCORED * 8 + 1024 * 8 + HXG_BRIDGE_BATCH * 8 + 64;
  HX_CUDA(cudaEventRecord(s->ev0, s->stream));
#define HXG_LAUNCH(M)                                                                                               \
  do {                                                                                                              \
    if (smem > 48 * 1024)                                                                                           \
      HX_CUDA(cudaFuncSetAttribute(k_filtered_walk<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    k_filtered_walk<M><<<grid, HXG_THREADS, smem, s->stream>>>(dev, a);                                             \
  } while (0)
  switch (ix->cfg.metric) {
    case HX_METRIC_EUCLIDEAN: HXG_LAUNCH(HXM_EUCLIDEAN); break;
    case HX_METRIC_COSINE: HXG_LAUNCH(HXM_COSINE); break;
    default: HXG_LAUNCH(HXM_MANHATTAN); break;
  }
#undef HXG_LAUNCH
  HX_CUDA(cudaGetLastError());
  HX_CUDA(cudaEventRecord(s->ev1, s->stream));
  launches++;
  if ((rc = s->h_status.reserve(B)) || (rc = s->h_qerr.reserve(B))) return rc;
  unsigned long long hst[12] = {0};
  HX_CUDA(cudaMemcpyAsync(out_ids, s->d_out_ids.p, B * (size_t)k * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
  HX_CUDA(cudaMemcpyAsync(out_scores, s->d_out_scores.p, B * (size_t)k * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
  HX_CUDA(cudaMemcpyAsync(out_counts, s->d_out_counts.p, B * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
  HX_CUDA(cudaMemcpyAsync(s->h_status.p, s->d_qstatus.p, B * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
  HX_CUDA(cudaMemcpyAsync(s->h_qerr.p, s->d_qerr.p, B * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
  HX_CUDA(cudaMemcpyAsync(hst, s->d_pstats.p, sizeof(hst), cudaMemcpyDeviceToHost, s->stream));
  HX_CUDA(cudaStreamSynchronize(s->stream));
  for (size_t b = 0; b < B; ++b) {
    if (s->h_status.p[b] != HX_ST_OK) return status_from_word(s->h_status.p[b], b, "query");
    const uint32_t e = s->h_qerr.p[b];
    if (e & HXG_ERR_MISSING_SIMHASH) {
      hx_set_error("query %zu: a graph neighbour or the entry point has no SimHash row (missing simhash)", b);
      return HX_ERR_INVARIANT_VIOLATION;
    }
    if (e & HXG_ERR_CAPACITY) {
      hx_set_error("query %zu: internal list capacity exceeded in the filtered walk", b);
      return HX_ERR_INVARIANT_VIOLATION;
    }
    if (e & HXF_INVALID_SCORE) {
      hx_set_error("vector distance kernel emitted an invalid score");
      return HX_ERR_INVARIANT_VIOLATION;
    }
  }
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, s->ev0, s->ev1) == cudaSuccess) {
    ix->last_kernel_ms = ms;
    ix->last_kernel_launches = 1;
  }
  if (stats) {
    stats->vector_payload_requests = hst[0];
    stats->distance_computations = hst[1];
    stats->routing_rows = hst[2];
    stats->bridge_rows = hst[3];
    stats->bridge_frontier_pushes = hst[4];
    stats->iterations = hst[5];
    stats->kernel_launches = launches;
  }
  return HX_OK;
}
