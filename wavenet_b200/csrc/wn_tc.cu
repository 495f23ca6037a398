// tcgen05 / TMEM / TMA kernels (kind::tf32) -- placeholder until the fused layer kernel lands.
#include "wn_common.h"

bool tc_layer_supported(const wn_handle* h) {
  (void)h;
  return false;
}

int tc_forward_residual(wn_handle* h, const float* params, cudaStream_t s) {
  (void)h;
  (void)params;
  (void)s;
  wn_set_error("tcgen05 path not built");
  return WN_EINVAL;
}
