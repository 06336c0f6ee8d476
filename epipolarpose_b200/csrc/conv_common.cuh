// Internal interfaces between the conv dispatch (conv.cu), the fp32 SIMT
// path (conv_simt.cu) and the tcgen05 tensor-core path (conv_tc.cu).
#pragma once
#include "common.cuh"

int epb_conv_geom_check(const epb_conv_geom* g);

int epb_conv_fprop_simt(const epb_conv_geom* g, const float* in, const float* w,
                        const float* in_scale, const float* in_shift, const float* bias,
                        float* out, double* stats, cudaStream_t st);
int epb_conv_wgrad_simt(const epb_conv_geom* g, const float* in, const float* dout,
                        const float* in_scale, const float* in_shift, float* dw, cudaStream_t st);

// tensor-core path; return EPB_EINVAL (without setting an error) if the shape
// is outside what the tcgen05 kernels take, so the dispatcher can route it.
bool epb_conv_tc_supported(const epb_conv_geom* g, bool wgrad);
int epb_conv_fprop_tc(const epb_conv_geom* g, const float* in, const float* w,
                      const float* in_scale, const float* in_shift, const float* bias,
                      float* out, double* stats, cudaStream_t st);
int epb_conv_wgrad_tc(const epb_conv_geom* g, const float* in, const float* dout,
                      const float* in_scale, const float* in_shift, float* dw, cudaStream_t st);
