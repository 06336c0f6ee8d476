// tcgen05 tensor-core path of the tap-list implicit GEMM (placeholder until the
// kernels land: reports every shape as unsupported so the dispatcher routes to
// the fp32 CUDA-core kernels).
#include "conv_common.cuh"

bool epb_conv_tc_supported(const epb_conv_geom*, bool) { return false; }
int epb_conv_fprop_tc(const epb_conv_geom*, const float*, const float*, const float*, const float*,
                      const float*, float*, double*, cudaStream_t) {
  epb_set_error("tcgen05 conv path not built");
  return EPB_EINVAL;
}
int epb_conv_wgrad_tc(const epb_conv_geom*, const float*, const float*, const float*, const float*,
                      float*, cudaStream_t) {
  epb_set_error("tcgen05 conv path not built");
  return EPB_EINVAL;
}
