// Dispatch of the conv family entry points (include/epb.h) onto the fp32
// CUDA-core kernels (precision 0) or the tcgen05 tensor-core kernels
// (precision 1 = TF32, 3 = 3xTF32 error-compensated).
#include "conv_common.cuh"

extern "C" __attribute__((visibility("default"))) int epb_conv_fprop(const epb_conv_geom* g, const float* in, const float* w,
                              const float* in_scale, const float* in_shift, const float* bias,
                              float* out, double* stats, epb_stream_t stream) {
  int rc = epb_conv_geom_check(g);
  if (rc) return rc;
  EPB_CHECK_ARG(in && w && out);
  EPB_CHECK_ARG((in_scale == nullptr) == (in_shift == nullptr));
  EPB_CHECK_ARG(g->precision == 0 || g->precision == 1 || g->precision == 3);
  EPB_CHECK_ARG(!(stats && g->accumulate));   // statistics describe a freshly written output
  if (g->precision != 0 && epb_conv_tc_supported(g, false))
    return epb_conv_fprop_tc(g, in, w, in_scale, in_shift, bias, out, stats, as_stream(stream));
  return epb_conv_fprop_simt(g, in, w, in_scale, in_shift, bias, out, stats, as_stream(stream));
}

extern "C" __attribute__((visibility("default"))) int epb_conv_wgrad(const epb_conv_geom* g, const float* in, const float* dout,
                              const float* in_scale, const float* in_shift, float* dw,
                              epb_stream_t stream) {
  int rc = epb_conv_geom_check(g);
  if (rc) return rc;
  EPB_CHECK_ARG(in && dout && dw);
  EPB_CHECK_ARG((in_scale == nullptr) == (in_shift == nullptr));
  EPB_CHECK_ARG(g->precision == 0 || g->precision == 1 || g->precision == 3);
  if (g->precision != 0 && epb_conv_tc_supported(g, true))
    return epb_conv_wgrad_tc(g, in, dout, in_scale, in_shift, dw, as_stream(stream));
  return epb_conv_wgrad_simt(g, in, dout, in_scale, in_shift, dw, as_stream(stream));
}
