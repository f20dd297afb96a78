// k_build.cu — HNSW construction on the device.  Placeholder until the batched builder lands.
#include "hx_index.hpp"

hx_status hx_build_impl(hx_index* ix, const uint16_t* levels, uint64_t seed) {
  (void)ix; (void)levels; (void)seed;
  hx_set_error("hx_index_build: device construction not available in this build");
  return HX_ERR_UNSUPPORTED;
}
