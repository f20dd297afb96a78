// hx_mirror.inl — incremental maintenance of the device mirror (included by hx_api.cu).
//
// The reference mutates one node at a time (insert_hnsw, mutation.rs:642-780; delete_from_layer, :1819) and tells its
// resident cache which rows became stale (VectorMemoryDirtyRows: dirty_nodes + dirty_upper_neighbors,
// memory_store.rs:105-130); readers use the cache only when its hydration sequence equals the request's snapshot sequence
// (read_index.rs:53-65).  The device mirror follows the same protocol: the Rust side applies a committed write as ROW
// PATCHES — the vector rows, layer-0 / upper neighbour rows and SimHash rows the write touched, exactly the dirty set —
// and then advances the (generation, visible_seq) token the read-side guard compares.  Nothing is recomputed here: the
// rows are the reference's own post-write rows, so searches after a patch equal searches after a full re-hydration.
//
// Slots stay in ascending-id order (the (score, id) tie rule is a single 64-bit compare): an existing id is
// overwritten in place, a new id above the current maximum is appended (node ids come from a monotonic allocator,
// crates/db/src/id_allocator.rs), anything else returns HX_ERR_UNSUPPORTED and the caller re-hydrates that generation.

template <typename T>
static hx_status grow_dev(T** p, size_t old_n, size_t new_n, int fill) {
  T* q = nullptr;
  HX_CUDA(cudaMalloc((void**)&q, std::max<size_t>(new_n, 1) * sizeof(T)));
  cudaError_t e = cudaMemset(q, fill, std::max<size_t>(new_n, 1) * sizeof(T));
  if (e == cudaSuccess && *p && old_n) e = cudaMemcpy(q, *p, old_n * sizeof(T), cudaMemcpyDeviceToDevice);
  if (e != cudaSuccess) {
    cudaFree(q);
    hx_set_error("mirror growth failed: %s", cudaGetErrorString(e));
    return HX_ERR_CUDA;
  }
  if (*p) cudaFree(*p);
  *p = q;
  return HX_OK;
}

static void invalidate_bf16(hx_index* ix) {
  if (ix->d_vec_bf16) cudaFree(ix->d_vec_bf16);
  if (ix->d_sqnorm) cudaFree(ix->d_sqnorm);
  ix->d_vec_bf16 = nullptr;
  ix->d_sqnorm = nullptr;
}

// capacity for `want` rows in every per-row array (vectors, headers, ids, layer-0 rows, levels, fingerprints)
static hx_status reserve_rows(hx_index* ix, size_t want) {
  const size_t have = std::max(ix->cap_rows, ix->n);
  if (want <= have) return HX_OK;
  const size_t cap = std::max(want, have + have / 2 + 64);
  if (cap >= (1ull << 31)) {
    hx_set_error("a shard holds at most 2^31-1 rows");
    return HX_ERR_INVALID_PARAMETER;
  }
  hx_status rc;
  if ((rc = grow_dev(&ix->d_vec, ix->n * (size_t)ix->ld, cap * (size_t)ix->ld, 0))) return rc;
  if ((rc = grow_dev(&ix->d_hdr, ix->n, cap, 0))) return rc;
  if ((rc = grow_dev(&ix->d_ids, ix->n, cap, 0))) return rc;
  if (ix->d_nbr0) {
    if ((rc = grow_dev(&ix->d_nbr0, ix->n * (size_t)ix->stride0, cap * (size_t)ix->stride0, 0))) return rc;
    if ((rc = grow_dev(&ix->d_deg0, ix->n, cap, 0))) return rc;
    if ((rc = grow_dev(&ix->d_raw0, ix->n, cap, 0))) return rc;
    if ((rc = grow_dev(&ix->d_upper_off, ix->n, cap, 0xFF))) return rc;   // HX_ABSENT
    if ((rc = grow_dev(&ix->d_level, ix->n, cap, 0))) return rc;
  }
  if (ix->d_simhash) {
    if ((rc = grow_dev(&ix->d_simhash, ix->n, cap, 0))) return rc;
    if ((rc = grow_dev(&ix->d_has_simhash, ix->n, cap, 0))) return rc;
  }
  if (ix->d_deleted && (rc = grow_dev(&ix->d_deleted, ix->n, cap, 0))) return rc;
  ix->cap_rows = cap;
  return HX_OK;
}

static hx_status reserve_upper(hx_index* ix, size_t want) {
  const size_t have = std::max(ix->cap_upper, std::max<size_t>(ix->n_upper_rows, 1));
  if (want <= have && ix->d_upper_nbr) return HX_OK;
  const size_t cap = std::max(want, have + have / 2 + 16);
  hx_status rc;
  if ((rc = grow_dev(&ix->d_upper_nbr, ix->n_upper_rows * (size_t)ix->stride_u, cap * (size_t)ix->stride_u, 0))) return rc;
  if ((rc = grow_dev(&ix->d_upper_deg, ix->n_upper_rows, cap, 0))) return rc;
  ix->cap_upper = cap;
  return HX_OK;
}

// rows[i] (stride words) -> dst[row_idx[i]]; deg / raw likewise
static __global__ void k_patch_rows(uint32_t* __restrict__ dst, uint32_t stride, uint16_t* __restrict__ deg_dst,
                                    uint16_t* __restrict__ raw_dst, const uint32_t* __restrict__ row_idx,
                                    const uint32_t* __restrict__ rows, const uint16_t* __restrict__ deg,
                                    const uint16_t* __restrict__ raw, uint32_t n_rows) {
  const uint32_t r = blockIdx.x;
  if (r >= n_rows) return;
  const size_t d = row_idx[r];
  for (uint32_t i = threadIdx.x; i < stride; i += blockDim.x) dst[d * stride + i] = rows[(size_t)r * stride + i];
  if (threadIdx.x == 0) {
    deg_dst[d] = deg[r];
    if (raw_dst) raw_dst[d] = raw[r];
  }
}

extern "C" hx_status hx_index_set_version(hx_index* ix, uint64_t generation, uint64_t visible_seq) {
  if (!ix) return HX_ERR_INDEX_NOT_FOUND;
  ix->mirror_generation.store(generation);
  ix->mirror_visible_seq.store(visible_seq);
  return HX_OK;
}

extern "C" hx_status hx_index_get_version(hx_index* ix, uint64_t* generation, uint64_t* visible_seq, uint64_t* patches) {
  if (!ix) return HX_ERR_INDEX_NOT_FOUND;
  if (generation) *generation = ix->mirror_generation.load();
  if (visible_seq) *visible_seq = ix->mirror_visible_seq.load();
  if (patches) *patches = ix->mirror_patches.load();
  return HX_OK;
}

extern "C" hx_status hx_index_upsert_vectors(hx_index* ix, const uint64_t* ids, const float* rows, size_t n) {
  if (!ix) return HX_ERR_INDEX_NOT_FOUND;
  if (n == 0) return HX_OK;
  if (!ids || !rows) return HX_ERR_INVALID_PARAMETER;
  HX_CUDA(cudaSetDevice(ix->device));
  if (ix->n == 0) return hx_index_load_vectors(ix, ids, rows, n);
  hx_status rc = hx_finalize_graph(ix);
  if (rc) return rc;
  const uint32_t dim = ix->cfg.dimension;
  // classify: existing ids are overwritten; new ids must extend the id order at its upper end
  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0u);
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return ids[a] < ids[b]; });
  for (size_t i = 1; i < n; ++i)
    if (ids[order[i]] == ids[order[i - 1]]) {
      hx_set_error("duplicate node id %llu in one patch", (unsigned long long)ids[order[i]]);
      return HX_ERR_INVARIANT_VIOLATION;
    }
  const uint64_t max_id = ix->ids_sorted.back();
  std::vector<uint32_t> dst(n);
  size_t n_new = 0;
  for (size_t i = 0; i < n; ++i) {
    const uint64_t id = ids[order[i]];
    uint32_t slot;
    auto it = std::lower_bound(ix->ids_sorted.begin(), ix->ids_sorted.end(), id);
    if (it != ix->ids_sorted.end() && *it == id) {
      slot = (uint32_t)(it - ix->ids_sorted.begin());
      if (n_new) {   // cannot happen: order is ascending and every new id is above max_id
        hx_set_error("internal: existing id after a new id");
        return HX_ERR_INVARIANT_VIOLATION;
      }
    } else if (id > max_id) {
      slot = (uint32_t)(ix->n + n_new);
      n_new++;
    } else {
      hx_set_error("node id %llu falls inside the mirrored id range but has no row: in-place insertion would renumber the "
                   "slots; re-hydrate this generation", (unsigned long long)id);
      return HX_ERR_UNSUPPORTED;
    }
    dst[i] = slot;
  }
  if ((rc = reserve_rows(ix, ix->n + n_new))) return rc;
  // stage, validate like decode_item_borrowed (finite, magnitude bound, non-zero for cosine), compute headers
  float* d_stage = nullptr;
  float* d_shdr = nullptr;
  uint32_t* d_status = nullptr;
  std::vector<float> stage(n * (size_t)ix->ld, 0.f);
  for (size_t i = 0; i < n; ++i) memcpy(stage.data() + i * (size_t)ix->ld, rows + (size_t)order[i] * dim, dim * sizeof(float));
  cudaError_t e = cudaMalloc((void**)&d_stage, stage.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_shdr, n * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_status, n * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemcpy(d_stage, stage.data(), stage.size() * sizeof(float), cudaMemcpyHostToDevice);
  std::vector<uint32_t> st(n, 0);
  if (e == cudaSuccess) {
    float limit = 0.f;
    const bool has_limit = component_limit(ix->cfg.metric, dim, &limit);
    k_validate_and_header<<<(unsigned)((n + 7) / 8), 256>>>(d_stage, n, dim, ix->ld, ix->cfg.metric, limit, has_limit ? 1 : 0,
                                                           d_shdr, d_status);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(st.data(), d_status, n * sizeof(uint32_t), cudaMemcpyDeviceToHost);
  }
  hx_status out = HX_OK;
  if (e == cudaSuccess)
    for (size_t i = 0; i < n && !out; ++i)
      if (st[i] != HX_ST_OK) out = status_from_word(st[i], order[i], "row");
  if (e == cudaSuccess && !out) {
    for (size_t i = 0; i < n && e == cudaSuccess; ++i) {
      e = cudaMemcpyAsync(ix->d_vec + (size_t)dst[i] * ix->ld, d_stage + i * (size_t)ix->ld, (size_t)ix->ld * sizeof(float),
                          cudaMemcpyDeviceToDevice, 0);
      if (e == cudaSuccess) e = cudaMemcpyAsync(ix->d_hdr + dst[i], d_shdr + i, sizeof(float), cudaMemcpyDeviceToDevice, 0);
    }
    if (e == cudaSuccess && n_new) {
      std::vector<uint64_t> new_ids(n_new);
      for (size_t i = 0; i < n_new; ++i) new_ids[i] = ids[order[n - n_new + i]];
      e = cudaMemcpyAsync(ix->d_ids + ix->n, new_ids.data(), n_new * sizeof(uint64_t), cudaMemcpyHostToDevice, 0);
      if (e == cudaSuccess) e = cudaStreamSynchronize(0);
      if (e == cudaSuccess) {
        ix->ids_sorted.insert(ix->ids_sorted.end(), new_ids.begin(), new_ids.end());
        if (!ix->host_deleted.empty()) ix->host_deleted.resize(ix->n + n_new, 0);
        ix->n += n_new;
        ix->contiguous = (ix->ids_sorted.back() - ix->ids_sorted.front() == (uint64_t)(ix->n - 1));
      }
    } else if (e == cudaSuccess) {
      e = cudaStreamSynchronize(0);
    }
    // an overwritten id that had been deleted is alive again
    if (e == cudaSuccess && ix->d_deleted)
      for (size_t i = 0; i < n - n_new && e == cudaSuccess; ++i)
        if (ix->host_deleted[dst[i]]) {
          ix->host_deleted[dst[i]] = 0;
          ix->n_deleted--;
          e = cudaMemset(ix->d_deleted + dst[i], 0, 1);
        }
  }
  cudaFree(d_stage);
  cudaFree(d_shdr);
  cudaFree(d_status);
  if (e != cudaSuccess) {
    hx_set_error("vector patch failed: %s", cudaGetErrorString(e));
    return HX_ERR_CUDA;
  }
  if (out) return out;
  invalidate_bf16(ix);   // rebuilt lazily by the next dense search
  ix->mirror_patches.fetch_add(1);
  return HX_OK;
}

// Declare the HNSW level of (new) nodes: allocates their upper rows (empty) — insert_hnsw creates empty rows on every
// layer <= the node's layer (mutation.rs:706-739).  Levels never change after insertion.
extern "C" hx_status hx_index_set_levels(hx_index* ix, const uint64_t* ids, const uint16_t* levels, size_t n) {
  if (!ix) return HX_ERR_INDEX_NOT_FOUND;
  if (n == 0) return HX_OK;
  if (!ids || !levels) return HX_ERR_INVALID_PARAMETER;
  HX_CUDA(cudaSetDevice(ix->device));
  hx_status rc = hx_finalize_graph(ix);
  if (rc) return rc;
  if (!ix->d_nbr0) {
    hx_set_error("no graph image to patch: load or build the graph first");
    return HX_ERR_INVALID_PARAMETER;
  }
  for (size_t i = 0; i < n; ++i) {
    if (levels[i] > 63) {
      hx_set_error("layer %u above the maximum 63", levels[i]);
      return HX_ERR_INVALID_PARAMETER;
    }
    uint32_t slot;
    if (!hx_slot_of(ix, ids[i], &slot)) {
      hx_set_error("node %llu has no vector row in the device mirror", (unsigned long long)ids[i]);
      return HX_ERR_INVARIANT_VIOLATION;
    }
    uint8_t have = 0;
    HX_CUDA(cudaMemcpy(&have, ix->d_level + slot, 1, cudaMemcpyDeviceToHost));
    if (have == levels[i]) continue;
    if (have != 0) {
      hx_set_error("node %llu already has level %u (levels are fixed at insertion)", (unsigned long long)ids[i], have);
      return HX_ERR_UNSUPPORTED;
    }
    if ((rc = reserve_upper(ix, ix->n_upper_rows + levels[i]))) return rc;
    const uint32_t off = (uint32_t)ix->n_upper_rows;
    const uint8_t lv = (uint8_t)levels[i];
    HX_CUDA(cudaMemset(ix->d_upper_deg + off, 0, levels[i] * sizeof(uint16_t)));
    HX_CUDA(cudaMemcpy(ix->d_upper_off + slot, &off, 4, cudaMemcpyHostToDevice));
    HX_CUDA(cudaMemcpy(ix->d_level + slot, &lv, 1, cudaMemcpyHostToDevice));
    ix->n_upper_rows += levels[i];
  }
  ix->mirror_patches.fetch_add(1);
  return HX_OK;
}

// Replace the neighbour rows of `node_ids` on `layer` (decoded CSR, like hx_index_load_graph, but in place).
extern "C" hx_status hx_index_upsert_neighbor_rows(hx_index* ix, uint16_t layer, const uint64_t* node_ids,
                                                   const uint32_t* offsets, const uint64_t* neighbors, size_t n_nodes) {
  if (!ix) return HX_ERR_INDEX_NOT_FOUND;
  if (n_nodes == 0) return HX_OK;
  if (!node_ids || !offsets) return HX_ERR_INVALID_PARAMETER;
  if (layer > 63) {
    hx_set_error("layer %u above the maximum 63", layer);
    return HX_ERR_INVALID_PARAMETER;
  }
  HX_CUDA(cudaSetDevice(ix->device));
  hx_status rc = hx_finalize_graph(ix);
  if (rc) return rc;
  if (!ix->d_nbr0) {
    hx_set_error("no graph image to patch: load or build the graph first");
    return HX_ERR_INVALID_PARAMETER;
  }
  const uint32_t stride = layer == 0 ? ix->stride0 : ix->stride_u;
  std::vector<uint32_t> rows(n_nodes * (size_t)stride, 0), row_idx(n_nodes);
  std::vector<uint16_t> deg(n_nodes), raw(n_nodes);
  for (size_t i = 0; i < n_nodes; ++i) {
    uint32_t slot;
    if (!hx_slot_of(ix, node_ids[i], &slot)) {
      hx_set_error("layer %u row of node %llu: the node has no vector row in the device mirror", layer,
                   (unsigned long long)node_ids[i]);
      return HX_ERR_INVARIANT_VIOLATION;
    }
    const uint32_t b = offsets[i], e = offsets[i + 1];
    uint32_t kept = 0;
    uint64_t prev = 0;
    for (uint32_t j = b; j < e; ++j) {
      if (j > b && neighbors[j] <= prev) {
        hx_set_error("layer %u row of node %llu is not strictly ascending", layer, (unsigned long long)node_ids[i]);
        return HX_ERR_INVARIANT_VIOLATION;
      }
      prev = neighbors[j];
      if (neighbors[j] == node_ids[i]) {
        hx_set_error("layer %u row of node %llu links to itself", layer, (unsigned long long)node_ids[i]);
        return HX_ERR_INVARIANT_VIOLATION;
      }
      uint32_t ns;
      if (!hx_slot_of(ix, neighbors[j], &ns)) continue;   // skipped at fetch time by the reference (search.rs:909-913)
      if (kept >= stride) {
        hx_set_error("layer %u row of node %llu holds more than %u neighbours: re-hydrate this generation", layer,
                     (unsigned long long)node_ids[i], stride);
        return HX_ERR_UNSUPPORTED;
      }
      rows[i * (size_t)stride + kept++] = ns;
    }
    deg[i] = (uint16_t)kept;
    raw[i] = (uint16_t)std::min<uint32_t>(e - b, 65535);
    if (layer == 0) {
      row_idx[i] = slot;
    } else {
      uint8_t lv = 0;
      uint32_t off = HX_ABSENT;
      HX_CUDA(cudaMemcpy(&lv, ix->d_level + slot, 1, cudaMemcpyDeviceToHost));
      HX_CUDA(cudaMemcpy(&off, ix->d_upper_off + slot, 4, cudaMemcpyDeviceToHost));
      if (lv < layer || off == HX_ABSENT) {
        hx_set_error("node %llu has no row on layer %u (declare its level with hx_index_set_levels)",
                     (unsigned long long)node_ids[i], layer);
        return HX_ERR_INVARIANT_VIOLATION;
      }
      row_idx[i] = off + layer - 1u;
    }
  }
  uint32_t *d_rows = nullptr, *d_idx = nullptr;
  uint16_t *d_deg = nullptr, *d_raw = nullptr;
  cudaError_t e = cudaMalloc((void**)&d_rows, rows.size() * 4);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_idx, n_nodes * 4);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_deg, n_nodes * 2);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_raw, n_nodes * 2);
  if (e == cudaSuccess) e = cudaMemcpy(d_rows, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_idx, row_idx.data(), n_nodes * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_deg, deg.data(), n_nodes * 2, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_raw, raw.data(), n_nodes * 2, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    if (layer == 0)
      k_patch_rows<<<(unsigned)n_nodes, 64>>>(ix->d_nbr0, stride, ix->d_deg0, ix->d_raw0, d_idx, d_rows, d_deg, d_raw, (uint32_t)n_nodes);
    else
      k_patch_rows<<<(unsigned)n_nodes, 64>>>(ix->d_upper_nbr, stride, ix->d_upper_deg, nullptr, d_idx, d_rows, d_deg, d_raw, (uint32_t)n_nodes);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
  }
  cudaFree(d_rows); cudaFree(d_idx); cudaFree(d_deg); cudaFree(d_raw);
  if (e != cudaSuccess) {
    hx_set_error("neighbour-row patch failed: %s", cudaGetErrorString(e));
    return HX_ERR_CUDA;
  }
  ix->mirror_patches.fetch_add(1);
  return HX_OK;
}

// delete: the node's rows are gone (its vector can no longer be scored or returned); the rows of its former neighbours
// arrive as hx_index_upsert_neighbor_rows patches (delete_from_layer repairs them on the Rust side).  Deleting the entry
// point leaves the index unpopulated until hx_index_set_entry names the new one (configuration.rs).
extern "C" hx_status hx_index_delete_vectors(hx_index* ix, const uint64_t* ids, size_t n) {
  if (!ix) return HX_ERR_INDEX_NOT_FOUND;
  if (n == 0) return HX_OK;
  if (!ids) return HX_ERR_INVALID_PARAMETER;
  HX_CUDA(cudaSetDevice(ix->device));
  hx_status rc = hx_finalize_graph(ix);
  if (rc) return rc;
  if (!ix->d_deleted) {
    const size_t cap = std::max(ix->cap_rows, ix->n);
    HX_CUDA(cudaMalloc((void**)&ix->d_deleted, std::max<size_t>(cap, 1)));
    HX_CUDA(cudaMemset(ix->d_deleted, 0, std::max<size_t>(cap, 1)));
    ix->host_deleted.assign(ix->n, 0);
  }
  const uint8_t one = 1;
  const uint16_t zero16 = 0;
  for (size_t i = 0; i < n; ++i) {
    uint32_t slot;
    if (!hx_slot_of(ix, ids[i], &slot)) continue;   // deleting an absent id is a no-op (idempotent, like the reference)
    ix->host_deleted[slot] = 1;
    ix->n_deleted++;
    HX_CUDA(cudaMemcpy(ix->d_deleted + slot, &one, 1, cudaMemcpyHostToDevice));
    if (ix->d_nbr0) {
      HX_CUDA(cudaMemcpy(ix->d_deg0 + slot, &zero16, 2, cudaMemcpyHostToDevice));
      HX_CUDA(cudaMemcpy(ix->d_raw0 + slot, &zero16, 2, cudaMemcpyHostToDevice));
      uint8_t lv = 0;
      uint32_t off = HX_ABSENT;
      HX_CUDA(cudaMemcpy(&lv, ix->d_level + slot, 1, cudaMemcpyDeviceToHost));
      HX_CUDA(cudaMemcpy(&off, ix->d_upper_off + slot, 4, cudaMemcpyDeviceToHost));
      if (lv && off != HX_ABSENT) HX_CUDA(cudaMemset(ix->d_upper_deg + off, 0, lv * sizeof(uint16_t)));
    }
    if (ix->d_has_simhash) {
      HX_CUDA(cudaMemset(ix->d_has_simhash + slot, 0, 1));
      ix->simhash_count = ix->simhash_count ? ix->simhash_count - 1 : 0;
    }
    if (ix->populated && slot == ix->entry_slot) ix->populated = false;
  }
  ix->vector_generation++;   // cached candidate sets were mapped before the deletion
  invalidate_bf16(ix);
  ix->mirror_patches.fetch_add(1);
  return HX_OK;
}

// [0x13] upper-vector rows (keys/vectors.rs:1532-1603): the reference's hot-lane copy of the item row of every node that
// lives above layer 0 (encode_item bytes, search/vector/index.rs:3116).  The device image keeps ONE copy of each vector,
// so importing them means: a node the mirror does not hold yet is inserted from its hot-lane row; a node it holds must
// carry byte-identical data (header included) — a stale hot-lane row is corruption, not something to search with.
extern "C" hx_status hx_index_load_upper_vector_rows(hx_index* ix, const uint64_t* ids, const uint8_t* rows, size_t n) {
  if (!ix) return HX_ERR_INDEX_NOT_FOUND;
  if (n == 0) return HX_OK;
  if (!ids || !rows) return HX_ERR_INVALID_PARAMETER;
  HX_CUDA(cudaSetDevice(ix->device));
  const uint32_t dim = ix->cfg.dimension;
  const size_t rb = 4 + 4 * (size_t)dim;
  std::vector<uint64_t> new_ids;
  std::vector<float> new_rows;
  std::vector<uint32_t> new_hdr;
  std::vector<float> have(dim);
  for (size_t i = 0; i < n; ++i) {
    const uint8_t* r = rows + i * rb;
    uint32_t slot;
    if (hx_slot_of(ix, ids[i], &slot)) {
      uint32_t hdr_dev = 0, hdr_row = 0;
      memcpy(&hdr_row, r, 4);
      HX_CUDA(cudaMemcpy(&hdr_dev, ix->d_hdr + slot, 4, cudaMemcpyDeviceToHost));
      HX_CUDA(cudaMemcpy(have.data(), ix->d_vec + (size_t)slot * ix->ld, dim * sizeof(float), cudaMemcpyDeviceToHost));
      if (hdr_dev != hdr_row || memcmp(have.data(), r + 4, 4 * (size_t)dim) != 0) {
        hx_set_error("upper-vector row of node %llu differs from the canonical vector row", (unsigned long long)ids[i]);
        return HX_ERR_INVARIANT_VIOLATION;
      }
    } else {
      new_ids.push_back(ids[i]);
      uint32_t h;
      memcpy(&h, r, 4);
      new_hdr.push_back(h);
      const size_t o = new_rows.size();
      new_rows.resize(o + dim);
      memcpy(new_rows.data() + o, r + 4, 4 * (size_t)dim);
    }
  }
  if (new_ids.empty()) return HX_OK;
  hx_status rc = hx_index_upsert_vectors(ix, new_ids.data(), new_rows.data(), new_ids.size());
  if (rc) return rc;
  for (size_t i = 0; i < new_ids.size(); ++i) {   // decode_item_borrowed: the stored header must equal the recomputed one
    uint32_t slot, hdr_dev = 0;
    if (!hx_slot_of(ix, new_ids[i], &slot)) continue;
    HX_CUDA(cudaMemcpy(&hdr_dev, ix->d_hdr + slot, 4, cudaMemcpyDeviceToHost));
    if (hdr_dev != new_hdr[i]) {
      hx_set_error("upper-vector row of node %llu: stored header does not match the recomputed header",
                   (unsigned long long)new_ids[i]);
      return HX_ERR_INVARIANT_VIOLATION;
    }
  }
  return HX_OK;
}
