// Parameter block shared by the SIMT and tcgen05 convolution kernels.
#pragma once
#include "common.cuh"

namespace cagc {

constexpr int kMaxTaps = 25;

struct Tap {
    int dy, dx, slab;
};

struct ConvP {
    const float* in;
    const float* w;
    const float* in_scale;
    const float* out_scale;
    const float* noise;
    const float* noise_w;
    const float* bias;
    const float* residual;   // optional tensor of the output's layout, added AFTER the activation (ResBlock skip)
    float* out;
    float* workspace;        // optional scratch for split-K of small layers (cagc_conv_workspace_bytes); null: never split
    int64_t workspace_bytes;
    int B, Hin, Win, in_pitch;
    int Ho, Wo, in_stride;
    int n_cols;  // weight slab leading dimension == out pitch
    int out_valid;
    int Hout, Wout, out_stride, out_oy, out_ox;
    int64_t noise_bstride;
    int w_bslabs;            // > 0: the weights are PER-SAMPLE slab sets (modulation folded in), sample b starts at slab
                             // b * w_bslabs; 0: one set shared by the batch
    int act;                 // 0: none, 1: leaky ReLU (slope 0.2) * act_gain
    float act_gain;          // sqrt2 unless stated (0 is read as sqrt2)
    int ntaps;
    Tap taps[kMaxTaps];
};

}  // namespace cagc

// tcgen05 / TMA implicit-GEMM path (conv_tc.cu)
int cagc_tc_conv(cudaStream_t stream, const cagc::ConvP& p, const char* what);
// 1 when cagc_tc_conv would run this shape on the halo-tile kernel (the only one that takes per-sample weight slabs)
int cagc_tc_conv_takes_sample_weights(const cagc::ConvP& p);
// several phases (output parities of the transposed convolution) in one persistent launch; 1 = handled (*rc)
int cagc_tc_conv_multi(cudaStream_t stream, const cagc::ConvP* phases, int nphase, const char* what, int* rc);
int cagc_tc_conv_multi_splitk(cudaStream_t stream, const cagc::ConvP* phases, int nphase, float* workspace,
                              int64_t workspace_bytes, const char* what, int* rc);
// weight gradient on the tensor pipe; `a` is the pre-modulated layer input; returns the number of splits used
int cagc_tc_wgrad_splits(int B, int H, int W, int a_pitch, int g_pitch, int ksize);
int cagc_tc_wgrad(cudaStream_t stream, const float* a, const float* g, float* partial, int* nsplits_io, int B, int H,
                  int W, int a_pitch, int g_pitch, int ksize, int mode);

// TMA-staged NHWC FIR; returns 1 when it handled the call (result code in *rc), 0 to use the plain kernel
int cagc_tc_fir_nhwc(cudaStream_t stream, const float* in, const float* fir, const float* out_scale, const float* noise,
                     const float* noise_w, const float* bias, float* out, int B, int in_h, int in_w, int out_h, int out_w,
                     int pitch, int valid, int pad_x0, int pad_y0, int64_t noise_bstride, int act, const float* taps_host,
                     int* rc, const float* mask_ref = nullptr, float mask_gain = 1.f);
