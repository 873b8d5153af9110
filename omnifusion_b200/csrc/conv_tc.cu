// tcgen05 / TMEM / TMA implicit-GEMM convolution engine (sm_100a).
#include "common.cuh"

namespace ofb {

bool conv_tc_supported(const ofb_conv_desc* d) { (void)d; return false; }

int conv_tc(const ofb_conv_desc* d, cudaStream_t s) {
  (void)d; (void)s;
  set_error("conv_tc: not built");
  return -1;
}

}  // namespace ofb
