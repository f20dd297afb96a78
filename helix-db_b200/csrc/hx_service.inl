// hx_service.inl — query service: many concurrent one-query callers coalesced into shared launches (included by hx_api.cu).
//
// The reference's calling pattern on this path is ONE query per call, many concurrent tokio tasks, no intra-query
// parallelism (search/vector/read_index.rs:81-101 <- execution/interpreter/access/search/storage.rs:142-192).  A blocking
// B = 1 hx_search costs a launch, a stream sync and two PCIe round trips per query and keeps one SM busy; the service turns
// the same traffic into the regime the traversal kernels are built for:
//
//   caller    : ticket = tail++ (lock-free, Vyukov-style sequence per slot) -> validates + writes its query into slot
//               ticket % capacity of a pinned ring -> publishes the slot -> later polls / futex-waits the slot's done word
//   dispatcher: one thread; takes the contiguous run of published slots at `head` (<= max_batch, no wrap) and issues, on the
//               next idle stream of a pool: 1-2 small H2D copies + ONE k_hnsw_search_cta_ring launch with grid = run length.
//               It never synchronises: several launches are in flight, 2-3 CTAs resident per SM.
//   kernel    : one CTA per query (visited set in shared memory, beam in registers: the latency build); writes ids / scores /
//               count straight into HOST-MAPPED slot memory, then __threadfence_system + the slot's done word.
//   completer : one thread; wakes (futex) the callers that chose to block.  Pollers never need it.
//
// Results are bit-identical to hx_search: same kernel, same admission order (tests/test_gpu_service.py).
#include <linux/futex.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <chrono>
#include <thread>

#define HXF_SERVICE_FATAL 0x100u   // host-side: the dispatcher hit a CUDA error; every pending query fails with HX_ERR_CUDA

static inline long hx_futex(volatile uint32_t* addr, int op, uint32_t val, const struct timespec* ts) {
  return syscall(SYS_futex, (uint32_t*)addr, op, val, ts, nullptr, 0);
}
static inline void hx_cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#else
  std::this_thread::yield();
#endif
}

struct HxSvcStream {
  cudaStream_t st = nullptr;
  uint64_t first_ticket = 0;   // last batch issued on this stream
  uint32_t cnt = 0;
};

struct hx_service {
  hx_index* ix = nullptr;
  uint32_t k = 0, ef = 0, cap = 0, mask = 0, max_batch = 0, dim = 0;
  uint32_t min_batch = 16, busy_inflight = 24, window_ns = 30000;   // coalescing policy (see svc_dispatcher)
  bool zero_copy = false;                                          // the kernel reads the queries straight from the pinned ring
  HxCtaRingCfg cta{};
  int ctas_per_sm = 0;
  bool cosine = false;
  float limit = 0.f;
  bool has_limit = false;
  // pinned, host-mapped (the kernel writes results + done words here; the copy engine reads queries from here)
  float* h_q = nullptr;          // [cap][dim]
  float* h_qhdr = nullptr;       // [cap]
  uint64_t* h_ids = nullptr;     // [cap][k]
  float* h_scores = nullptr;     // [cap][k]
  uint32_t* h_counts = nullptr;  // [cap]
  uint32_t* h_done = nullptr;    // [cap]  0 = running, 0x80000000 | flags = results visible
  // device
  float* d_q = nullptr;
  float* d_qhdr = nullptr;       // [cap] (zeros for the non-cosine metrics)
  uint32_t* d_zero = nullptr;    // [cap] q_status = 0 (validation happened at submit)
  uint32_t* d_err = nullptr;     // [4]
  uint32_t* d_vpool = nullptr;
  uint32_t* d_vbusy = nullptr;
  uint64_t* d_tiepool = nullptr;
  uint32_t* d_tiebusy = nullptr;
  uint32_t pool_n = 32, pool_cap = 0;
  std::vector<HxSvcStream> streams;
  size_t rr = 0;
  // control (host only)
  std::atomic<uint64_t>* seq = nullptr;   // [cap]: == ticket: free for it; == ticket + 1: published / in flight
  std::atomic<uint8_t>* waiting = nullptr; // [cap]: a caller sleeps on this slot's done word
  std::atomic<uint64_t> tail{0};
  uint64_t head = 0;                       // dispatcher only
  std::atomic<uint64_t> head_pub{0};
  std::atomic<bool> stop{false};
  std::atomic<uint32_t> fatal{0};          // hx_status of a dispatcher failure
  volatile uint32_t disp_sleep = 0;        // futex word: 1 = dispatcher is sleeping
  volatile uint32_t comp_sleep = 0;        // futex word: 1 = completer is sleeping
  std::atomic<uint32_t> n_waiters{0};
  std::thread dispatcher, completer;
  // statistics
  std::atomic<uint64_t> st_submitted{0}, st_completed{0}, st_launches{0}, st_max_batch{0}, st_disp_sleeps{0}, st_wakes{0};
};

static void svc_free(hx_service* v) {
  if (!v) return;
  for (auto& s : v->streams)
    if (s.st) cudaStreamDestroy(s.st);
  if (v->h_q) cudaFreeHost(v->h_q);
  if (v->h_qhdr) cudaFreeHost(v->h_qhdr);
  if (v->h_ids) cudaFreeHost(v->h_ids);
  if (v->h_scores) cudaFreeHost(v->h_scores);
  if (v->h_counts) cudaFreeHost(v->h_counts);
  if (v->h_done) cudaFreeHost(v->h_done);
  if (v->d_q) cudaFree(v->d_q);
  if (v->d_qhdr) cudaFree(v->d_qhdr);
  if (v->d_zero) cudaFree(v->d_zero);
  if (v->d_err) cudaFree(v->d_err);
  if (v->d_vpool) cudaFree(v->d_vpool);
  if (v->d_vbusy) cudaFree(v->d_vbusy);
  if (v->d_tiepool) cudaFree(v->d_tiepool);
  if (v->d_tiebusy) cudaFree(v->d_tiebusy);
  delete[] v->seq;
  delete[] v->waiting;
  delete v;
}

// every slot of [first, first + cnt) has either delivered (done word set) or been consumed and recycled
static bool svc_batch_finished(const hx_service* v, uint64_t first, uint32_t cnt) {
  for (uint32_t i = 0; i < cnt; ++i) {
    const uint64_t t = first + i;
    const uint32_t slot = (uint32_t)t & v->mask;
    if (v->seq[slot].load(std::memory_order_acquire) == t + 1 && ((volatile uint32_t*)v->h_done)[slot] == 0u) return false;
  }
  return true;
}

static void svc_fail_pending(hx_service* v, hx_status why) {
  v->fatal.store((uint32_t)why);
  // everything published from `head` on will never be launched: deliver the failure through the done words
  const uint64_t tail = v->tail.load();
  for (uint64_t t = v->head; t < tail; ++t) {
    const uint32_t slot = (uint32_t)t & v->mask;
    for (int spin = 0; spin < 100000 && v->seq[slot].load(std::memory_order_acquire) != t + 1; ++spin) hx_cpu_relax();
    ((volatile uint32_t*)v->h_done)[slot] = 0x80000000u | HXF_SERVICE_FATAL;
    hx_futex(v->h_done + slot, FUTEX_WAKE_PRIVATE, 1, nullptr);
  }
  v->head = tail;
  v->head_pub.store(tail);
}

static hx_status svc_launch(hx_service* v, uint64_t first, uint32_t cnt) {
  const uint32_t slot0 = (uint32_t)first & v->mask;
  // an idle stream if there is one (a launch queued behind a running batch on the same stream would wait for all of it)
  size_t pick = v->rr;
  for (size_t i = 0; i < v->streams.size(); ++i) {
    const size_t c = (v->rr + i) % v->streams.size();
    if (v->streams[c].cnt == 0 || svc_batch_finished(v, v->streams[c].first_ticket, v->streams[c].cnt)) { pick = c; break; }
  }
  v->rr = (pick + 1) % v->streams.size();
  HxSvcStream& S = v->streams[pick];
  HxHnswArgs a{};
  if (v->zero_copy) {
    // dimension a multiple of 32: every warp loads the query into registers once and never touches it again, so it can
    // come straight from the host-mapped ring (3 KB over PCIe per warp) — the launch is the only driver call of the batch
    a.queries = v->h_q + (size_t)slot0 * v->dim;
    a.q_hdr = v->h_qhdr + slot0;
  } else {
    HX_CUDA(cudaMemcpyAsync(v->d_q + (size_t)slot0 * v->dim, v->h_q + (size_t)slot0 * v->dim,
                            (size_t)cnt * v->dim * sizeof(float), cudaMemcpyHostToDevice, S.st));
    if (v->cosine)
      HX_CUDA(cudaMemcpyAsync(v->d_qhdr + slot0, v->h_qhdr + slot0, cnt * sizeof(float), cudaMemcpyHostToDevice, S.st));
    a.queries = v->d_q + (size_t)slot0 * v->dim;
    a.q_hdr = v->d_qhdr + slot0;
  }
  a.q_status = v->d_zero;
  a.B = cnt;
  a.k = v->k;
  a.ef = v->ef;
  a.out_ids = v->h_ids + (size_t)slot0 * v->k;       // host-mapped: UVA makes the host pointer valid on the device
  a.out_scores = v->h_scores + (size_t)slot0 * v->k;
  a.out_counts = v->h_counts + slot0;
  a.err_flags = v->d_err;
  a.fr_cap = v->cta.fr_cap;
  a.done = v->h_done + slot0;
  HxRingArgs rg{};
  rg.vt_cap = v->cta.vt_cap;
  rg.pool = v->d_vpool;
  rg.pool_busy = v->d_vbusy;
  rg.pool_n = v->pool_n;
  rg.pool_cap = v->pool_cap;
  rg.counter = v->d_err + 1;
  rg.l2_hint = 1;
  rg.batch_admit = 1;
  rg.tie_pool = v->d_tiepool;
  rg.tie_busy = v->d_tiebusy;
  rg.tie_pool_n = HX_TIE_POOL_N;
  rg.tie_pool_cap = HX_TIE_POOL_CAP;
  hx_status rc = hx_launch_cta_ring(v->ix, v->cta, a, rg, cnt, S.st, nullptr);
  if (rc) return rc;
  S.first_ticket = first;
  S.cnt = cnt;
  v->st_launches.fetch_add(1, std::memory_order_relaxed);
  uint64_t mb = v->st_max_batch.load(std::memory_order_relaxed);
  while (cnt > mb && !v->st_max_batch.compare_exchange_weak(mb, cnt)) {}
  return HX_OK;
}

static inline uint64_t svc_now_ns() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

// Coalescing policy.  A launch costs the dispatcher a few microseconds whatever its size, so one launch per query caps
// the service at the driver's launch rate (measured: ~70-100 k launches/s, profiles/r02_service_sweep_v1.jsonl).  When
// the device is idle (few queries in flight) whatever is pending is launched at once — lowest latency; when it is busy
// (>= busy_inflight queries in flight) a small batch is held back for at most window_ns so that the callers arriving in
// that window share the launch: the wait is a few percent of a traversal and is hidden behind the queries already running.
static void svc_dispatcher(hx_service* v) {
  cudaSetDevice(v->ix->device);
  uint32_t idle_spins = 0;
  uint64_t first_seen = 0;
  while (!v->stop.load(std::memory_order_acquire)) {
    // contiguous run of published slots at head (ring order == ticket order; a run never wraps: the arrays are contiguous)
    uint32_t cnt = 0;
    const uint32_t slot0 = (uint32_t)v->head & v->mask;
    while (cnt < v->max_batch && slot0 + cnt < v->cap &&
           v->seq[(slot0 + cnt)].load(std::memory_order_acquire) == v->head + cnt + 1)
      cnt++;
    if (cnt == 0) {
      first_seen = 0;
      if (++idle_spins < 20000) { hx_cpu_relax(); continue; }
      // nothing for a while: sleep until a submitter wakes us (re-check after announcing, so no wake-up is lost)
      v->disp_sleep = 1;
      __sync_synchronize();
      if (v->seq[slot0].load(std::memory_order_acquire) != v->head + 1 && !v->stop.load()) {
        struct timespec ts = {0, 50 * 1000 * 1000};
        v->st_disp_sleeps.fetch_add(1, std::memory_order_relaxed);
        hx_futex(&v->disp_sleep, FUTEX_WAIT_PRIVATE, 1, &ts);
      }
      v->disp_sleep = 0;
      idle_spins = 0;
      continue;
    }
    idle_spins = 0;
    if (cnt < v->min_batch && cnt < v->max_batch && slot0 + cnt < v->cap) {
      const uint64_t inflight = v->head - v->st_completed.load(std::memory_order_relaxed);
      if (inflight >= v->busy_inflight) {
        const uint64_t now = svc_now_ns();
        if (!first_seen) first_seen = now;
        if (now - first_seen < v->window_ns) { hx_cpu_relax(); continue; }
      }
    }
    first_seen = 0;
    const hx_status rc = svc_launch(v, v->head, cnt);
    if (rc) {
      svc_fail_pending(v, rc);
      continue;
    }
    v->head += cnt;
    v->head_pub.store(v->head, std::memory_order_release);
  }
}

static void svc_completer(hx_service* v) {
  while (!v->stop.load(std::memory_order_acquire)) {
    if (v->n_waiters.load(std::memory_order_acquire) == 0) {
      v->comp_sleep = 1;
      __sync_synchronize();
      if (v->n_waiters.load(std::memory_order_acquire) == 0 && !v->stop.load()) {
        struct timespec ts = {0, 50 * 1000 * 1000};
        hx_futex(&v->comp_sleep, FUTEX_WAIT_PRIVATE, 1, &ts);
      }
      v->comp_sleep = 0;
      continue;
    }
    // scan the waiting bytes eight at a time; wake the sleepers whose done word has been published by the device
    const uint64_t* w8 = reinterpret_cast<const uint64_t*>(v->waiting);
    for (uint32_t g = 0; g < v->cap / 8; ++g) {
      if (__atomic_load_n(w8 + g, __ATOMIC_RELAXED) == 0) continue;
      for (uint32_t slot = g * 8; slot < g * 8 + 8; ++slot) {
        if (!v->waiting[slot].load(std::memory_order_acquire)) continue;
        if (((volatile uint32_t*)v->h_done)[slot] == 0u) continue;
        v->waiting[slot].store(0, std::memory_order_release);
        hx_futex(v->h_done + slot, FUTEX_WAKE_PRIVATE, 1, nullptr);
        v->st_wakes.fetch_add(1, std::memory_order_relaxed);
      }
    }
    hx_cpu_relax();
  }
}

extern "C" hx_status hx_service_create(hx_index* ix, const hx_service_config* cfg, hx_service** out) {
  if (!ix) {
    hx_set_error("null index handle");
    return HX_ERR_INDEX_NOT_FOUND;
  }
  if (!cfg || !out || cfg->k == 0) {
    hx_set_error("hx_service_create: null argument or k == 0");
    return HX_ERR_INVALID_PARAMETER;
  }
  *out = nullptr;
  const uint32_t ef = cfg->ef ? cfg->ef : std::max(cfg->k, 100u);
  if (ef < cfg->k || ef > 4096) {
    hx_set_error("search beam width must be in [k, 4096], got %u", ef);
    return HX_ERR_INVALID_PARAMETER;
  }
  HX_CUDA(cudaSetDevice(ix->device));
  hx_status rc = hx_finalize_graph(ix);
  if (rc) return rc;
  if (ix->n == 0 || !ix->populated || !ix->d_nbr0) {
    hx_set_error("the query service needs a populated index (vectors, graph and entry point loaded)");
    return HX_ERR_INVALID_PARAMETER;
  }
  hx_service* v = new hx_service();
  v->ix = ix;
  v->k = cfg->k;
  v->ef = ef;
  v->dim = ix->cfg.dimension;
  uint32_t cap = cfg->capacity ? cfg->capacity : 1024u;
  uint32_t p2 = 64;
  while (p2 < cap && p2 < (1u << 20)) p2 <<= 1;
  v->cap = p2;
  v->mask = p2 - 1;
  v->max_batch = std::min(cfg->max_batch ? cfg->max_batch : 128u, v->cap);
  v->cosine = ix->cfg.metric == HX_METRIC_COSINE;
  v->zero_copy = (v->dim % 32u) == 0u && !(cfg->flags & HX_SERVICE_STAGE_QUERIES);
  if (cfg->min_batch) v->min_batch = cfg->min_batch;
  if (cfg->batch_window_us) v->window_ns = cfg->batch_window_us * 1000u;
  if (cfg->flags & HX_SERVICE_NO_COALESCING) v->window_ns = 0;
  v->has_limit = component_limit(ix->cfg.metric, v->dim, &v->limit);
  // launch shape: by default two CTAs resident per SM (half the shared memory each, 6 warps so two fit the register file)
  const uint32_t target = cfg->ctas_per_sm ? std::min(cfg->ctas_per_sm, 4u) : 2u;
  const uint32_t want_warps = cfg->cta_warps ? cfg->cta_warps : (target == 1 ? 12u : target == 2 ? 6u : 4u);
  const size_t budget = (size_t)(227 * 1024) / target - (target > 1 ? 1024 : 0);
  if (!hx_cta_ring_config(ix, ef, want_warps, cfg->rows_in_flight, cfg->visited_log2, budget, &v->cta) &&
      !hx_cta_ring_config(ix, ef, want_warps, cfg->rows_in_flight, cfg->visited_log2, 227 * 1024, &v->cta)) {
    hx_set_error("query working set exceeds shared memory (dimension %u, ef %u)", v->dim, ef);
    svc_free(v);
    return HX_ERR_INVALID_PARAMETER;
  }
  cudaError_t e = cudaSuccess;
  const unsigned hf = cudaHostAllocMapped | cudaHostAllocPortable;
  auto halloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaHostAlloc(p, bytes, hf); if (e == cudaSuccess) memset(*p, 0, bytes); };
  auto dalloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); if (e == cudaSuccess) e = cudaMemset(*p, 0, bytes); };
  halloc((void**)&v->h_q, (size_t)v->cap * v->dim * sizeof(float));
  halloc((void**)&v->h_qhdr, (size_t)v->cap * sizeof(float));
  halloc((void**)&v->h_ids, (size_t)v->cap * v->k * sizeof(uint64_t));
  halloc((void**)&v->h_scores, (size_t)v->cap * v->k * sizeof(float));
  halloc((void**)&v->h_counts, (size_t)v->cap * sizeof(uint32_t));
  halloc((void**)&v->h_done, (size_t)v->cap * sizeof(uint32_t));
  dalloc((void**)&v->d_q, (size_t)v->cap * v->dim * sizeof(float));
  dalloc((void**)&v->d_qhdr, (size_t)v->cap * sizeof(float));
  dalloc((void**)&v->d_zero, (size_t)v->cap * sizeof(uint32_t));
  dalloc((void**)&v->d_err, 4 * sizeof(uint32_t));
  v->pool_cap = std::max<uint32_t>(v->cta.vt_cap * 16u, 65536u);
  dalloc((void**)&v->d_vpool, (size_t)v->pool_n * v->pool_cap * sizeof(uint32_t));
  dalloc((void**)&v->d_vbusy, v->pool_n * sizeof(uint32_t));
  dalloc((void**)&v->d_tiepool, (size_t)HX_TIE_POOL_N * HX_TIE_POOL_CAP * sizeof(uint64_t));
  dalloc((void**)&v->d_tiebusy, HX_TIE_POOL_N * sizeof(uint32_t));
  const uint32_t ns = std::max(1u, std::min(cfg->n_streams ? cfg->n_streams : 32u, 128u));
  v->streams.resize(ns);
  for (auto& s : v->streams)
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    hx_set_error("query service allocation failed: %s", cudaGetErrorString(e));
    svc_free(v);
    return e == cudaErrorMemoryAllocation ? HX_ERR_OUT_OF_MEMORY : HX_ERR_CUDA;
  }
  // function attributes + occupancy once, before any thread launches (grid 0 = prepare only)
  {
    HxHnswArgs a{};
    HxRingArgs rg{};
    if ((rc = hx_launch_cta_ring(ix, v->cta, a, rg, 0, nullptr, &v->ctas_per_sm))) {
      svc_free(v);
      return rc;
    }
  }
  v->seq = new std::atomic<uint64_t>[v->cap];
  v->waiting = new std::atomic<uint8_t>[v->cap];
  for (uint32_t i = 0; i < v->cap; ++i) { v->seq[i].store(i); v->waiting[i].store(0); }
  v->dispatcher = std::thread(svc_dispatcher, v);
  v->completer = std::thread(svc_completer, v);
  *out = v;
  return HX_OK;
}

extern "C" void hx_service_destroy(hx_service* v) {
  if (!v) return;
  // let what is in flight finish, then stop the threads
  const auto t0 = std::chrono::steady_clock::now();
  while (v->head_pub.load() < v->tail.load() && !v->fatal.load() &&
         std::chrono::steady_clock::now() - t0 < std::chrono::seconds(10))
    std::this_thread::sleep_for(std::chrono::microseconds(200));
  v->stop.store(true);
  hx_futex(&v->disp_sleep, FUTEX_WAKE_PRIVATE, 1, nullptr);
  hx_futex(&v->comp_sleep, FUTEX_WAKE_PRIVATE, 1, nullptr);
  if (v->dispatcher.joinable()) v->dispatcher.join();
  if (v->completer.joinable()) v->completer.join();
  cudaSetDevice(v->ix->device);
  for (auto& s : v->streams)
    if (s.st) cudaStreamSynchronize(s.st);
  svc_free(v);
}

extern "C" hx_status hx_service_submit(hx_service* v, const float* query, uint64_t* out_ticket) {
  if (!v || !query || !out_ticket) {
    hx_set_error("hx_service_submit: null argument");
    return HX_ERR_INVALID_PARAMETER;
  }
  if (const uint32_t f = v->fatal.load()) {
    hx_set_error("the query service stopped after a device error");
    return (hx_status)f;
  }
  // ValidatedMetricVector::try_new (domain.rs:113-154) before anything is queued: a bad query never reaches the device
  const uint32_t w = host_validate(query, v->dim, v->ix->cfg.metric, v->has_limit, v->limit);
  if (w != HX_ST_OK) return status_from_word(w, 0, "query");
  const float hdr = v->cosine ? host_cosine_norm(query, v->dim) : 0.0f;
  const uint64_t t = v->tail.fetch_add(1, std::memory_order_acq_rel);
  const uint32_t slot = (uint32_t)t & v->mask;
  // back-pressure: the slot is still owned by ticket t - capacity until that caller has consumed its results
  for (uint32_t spins = 0; v->seq[slot].load(std::memory_order_acquire) != t; ++spins) {
    if (spins < 2000) hx_cpu_relax();
    else std::this_thread::sleep_for(std::chrono::microseconds(20));
  }
  memcpy(v->h_q + (size_t)slot * v->dim, query, (size_t)v->dim * sizeof(float));
  v->h_qhdr[slot] = hdr;
  ((volatile uint32_t*)v->h_done)[slot] = 0u;
  v->seq[slot].store(t + 1, std::memory_order_release);   // published
  __sync_synchronize();
  if (v->disp_sleep) {
    v->disp_sleep = 0;
    hx_futex(&v->disp_sleep, FUTEX_WAKE_PRIVATE, 1, nullptr);
  }
  v->st_submitted.fetch_add(1, std::memory_order_relaxed);
  *out_ticket = t;
  return HX_OK;
}

// the slot's results are visible: copy them out, recycle the slot, map the flags to the query's own status
static hx_status svc_collect(hx_service* v, uint64_t t, uint32_t slot, uint32_t done, uint64_t* out_ids, float* out_scores,
                             uint32_t* out_count) {
  __sync_synchronize();
  const uint32_t flags = done & 0x7fffffffu;
  hx_status st = HX_OK;
  uint32_t cnt = 0;
  if (flags & HXF_SERVICE_FATAL) {
    hx_set_error("the query service stopped after a device error");
    st = v->fatal.load() ? (hx_status)v->fatal.load() : HX_ERR_CUDA;
  } else if (flags) {
    st = check_device_flags(flags);
  }
  if (!st) {
    cnt = v->h_counts[slot];
    if (cnt > v->k) cnt = v->k;
    memcpy(out_ids, v->h_ids + (size_t)slot * v->k, cnt * sizeof(uint64_t));
    memcpy(out_scores, v->h_scores + (size_t)slot * v->k, cnt * sizeof(float));
  }
  if (out_count) *out_count = cnt;
  v->seq[slot].store(t + v->cap, std::memory_order_release);   // free for the next lap
  v->st_completed.fetch_add(1, std::memory_order_relaxed);
  return st;
}

extern "C" hx_status hx_service_poll(hx_service* v, uint64_t ticket, int32_t* out_done, uint64_t* out_ids, float* out_scores,
                                     uint32_t* out_count) {
  if (!v || !out_done || !out_ids || !out_scores) {
    hx_set_error("hx_service_poll: null argument");
    return HX_ERR_INVALID_PARAMETER;
  }
  *out_done = 0;
  const uint32_t slot = (uint32_t)ticket & v->mask;
  if (v->seq[slot].load(std::memory_order_acquire) != ticket + 1) {
    hx_set_error("unknown or already consumed ticket %llu", (unsigned long long)ticket);
    return HX_ERR_INVALID_PARAMETER;
  }
  const uint32_t done = ((volatile uint32_t*)v->h_done)[slot];
  if (done == 0u) return HX_OK;
  *out_done = 1;
  return svc_collect(v, ticket, slot, done, out_ids, out_scores, out_count);
}

extern "C" hx_status hx_service_wait(hx_service* v, uint64_t ticket, uint64_t* out_ids, float* out_scores,
                                     uint32_t* out_count) {
  if (!v || !out_ids || !out_scores) {
    hx_set_error("hx_service_wait: null argument");
    return HX_ERR_INVALID_PARAMETER;
  }
  const uint32_t slot = (uint32_t)ticket & v->mask;
  if (v->seq[slot].load(std::memory_order_acquire) != ticket + 1) {
    hx_set_error("unknown or already consumed ticket %llu", (unsigned long long)ticket);
    return HX_ERR_INVALID_PARAMETER;
  }
  volatile uint32_t* dw = (volatile uint32_t*)v->h_done + slot;
  uint32_t done = *dw;
  if (done == 0u) {
    // a traversal takes a few hundred microseconds: sleep on the done word right away; the completer thread wakes us
    v->n_waiters.fetch_add(1, std::memory_order_acq_rel);
    if (v->comp_sleep) {
      v->comp_sleep = 0;
      hx_futex(&v->comp_sleep, FUTEX_WAKE_PRIVATE, 1, nullptr);
    }
    while ((done = *dw) == 0u) {
      v->waiting[slot].store(1, std::memory_order_release);
      struct timespec ts = {0, 2 * 1000 * 1000};   // safety net only; the normal exit is the completer's wake
      hx_futex(dw, FUTEX_WAIT_PRIVATE, 0u, &ts);
    }
    v->waiting[slot].store(0, std::memory_order_release);
    v->n_waiters.fetch_sub(1, std::memory_order_acq_rel);
  }
  return svc_collect(v, ticket, slot, done, out_ids, out_scores, out_count);
}

extern "C" hx_status hx_service_search(hx_service* v, const float* query, uint64_t* out_ids, float* out_scores,
                                       uint32_t* out_count) {
  uint64_t t = 0;
  hx_status rc = hx_service_submit(v, query, &t);
  if (rc) {
    if (out_count) *out_count = 0;
    return rc;
  }
  return hx_service_wait(v, t, out_ids, out_scores, out_count);
}

extern "C" hx_status hx_service_get_stats(hx_service* v, hx_service_stats* o) {
  if (!v || !o) return HX_ERR_INVALID_PARAMETER;
  memset(o, 0, sizeof(*o));
  o->submitted = v->st_submitted.load();
  o->completed = v->st_completed.load();
  o->launches = v->st_launches.load();
  o->max_batch_seen = v->st_max_batch.load();
  o->dispatcher_sleeps = v->st_disp_sleeps.load();
  o->completer_wakes = v->st_wakes.load();
  o->cta_warps = v->cta.warps;
  o->rows_in_flight = v->cta.RC;
  o->visited_cap = v->cta.vt_cap;
  o->smem_bytes = (uint32_t)v->cta.smem;
  o->ctas_per_sm = (uint32_t)v->ctas_per_sm;
  return HX_OK;
}
