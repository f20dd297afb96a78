// k_dense.cu — tensor-core batched distance path (configs C4/C5).  Placeholder until the tcgen05 kernel lands.
#include "hx_index.hpp"

hx_status hx_dense_impl(hx_index* ix, const float* queries, size_t B, const hx_search_params* p, uint64_t* out_ids,
                        float* out_scores, uint32_t* out_counts, hx_stats* stats) {
  (void)ix; (void)queries; (void)B; (void)p; (void)out_ids; (void)out_scores; (void)out_counts; (void)stats;
  hx_set_error("hx_search_dense: tensor-core path not available in this build");
  return HX_ERR_UNSUPPORTED;
}
