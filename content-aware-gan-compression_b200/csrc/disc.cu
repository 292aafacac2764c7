// Discriminator-side kernels (reference model.py:670-798) on NHWC-p tensors.  The 3x3 / 1x1 / stride-2
// convolutions themselves run on the convolution engines of conv_tc.cu / conv_simt.cu through cagc_conv2d
// (shared weights: no modulation, bias + leaky ReLU (+ residual add) in the epilogue); this file holds the
// bandwidth-class pieces the reference spreads over upfirdn2d + conv2d + fused_bias_act launches:
//
//   fir_resample_nhwc   Blur fused with the stride of the following convolution: the ResBlock skip branch is
//                       Blur(pad 1,1) -> 1x1 conv stride 2 (model.py:683-689, 723-728); three of four blurred
//                       pixels are never read, so the FIR is evaluated at the even output positions only
//                       (upfirdn2d with down = 2), and its adjoint (up = 2) carries the gradient back
//   from_rgb fwd / bwd  the first ConvLayer(3, C, 1): a K = 3 contraction is a bandwidth kernel, not a GEMM:
//                       y = lrelu(W x + b) * sqrt2 written straight to NHWC; backward produces the image
//                       gradient from (g, y) in one pass (activation mask from the sign of y)
//   act_mask_nhwc       gz = g * gain * (y > 0 ? 1 : 0.2)   (fused_bias_act_kernel.cu:43 semantics, no bias)
#include <algorithm>

#include "conv_params.cuh"

namespace cagc {

struct FirRsP {
    const float* in;
    float* out;
    int B, in_h, in_w, out_h, out_w, pitch, pad_x0, pad_y0;
    float kf[16];   // FLIPPED taps (true convolution, op/upfirdn2d.py:159-200), row major
};

// thread = one float4 (4 channels) of one output pixel; consecutive threads walk the channel axis, so every
// tap is a fully coalesced row segment; the overlapping 4x4 windows of neighbouring outputs hit L1/L2.
template <int UP, int DOWN>
__global__ void __launch_bounds__(256) fir_resample_nhwc_kernel(const __grid_constant__ FirRsP p) {
    // blockIdx.y = (b, oy) output row; threads walk (ox, c4) of that row: no 64-bit divisions, one 32-bit division
    const int c4n = p.pitch >> 2;
    const int row = blockIdx.y;
    const int b = row / p.out_h, oy = row - b * p.out_h;
    const int row_items = p.out_w * c4n;
    const float* src_b = p.in + (int64_t)b * p.in_h * p.in_w * p.pitch;
    float* dst_row = p.out + (int64_t)row * p.out_w * p.pitch;
    const int ay0 = oy * DOWN - p.pad_y0;
    for (int it = blockIdx.x * 256 + threadIdx.x; it < row_items; it += gridDim.x * 256) {
        const int ox = it / c4n, c4 = it - ox * c4n;
        const float* src = src_b + c4 * 4;
        const int ax0 = ox * DOWN - p.pad_x0;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int a = ay0 + i;
            if (a < 0 || (UP == 2 && (a & 1))) continue;
            const int iy = a / UP;
            if (iy >= p.in_h) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int bb = ax0 + j;
                if (bb < 0 || (UP == 2 && (bb & 1))) continue;
                const int ix = bb / UP;
                if (ix >= p.in_w) continue;
                const float k = p.kf[i * 4 + j];
                const float4 v = ldg4(src + ((int64_t)iy * p.in_w + ix) * p.pitch);
                acc.x = fmaf(k, v.x, acc.x);
                acc.y = fmaf(k, v.y, acc.y);
                acc.z = fmaf(k, v.z, acc.z);
                acc.w = fmaf(k, v.w, acc.w);
            }
        }
        st4(dst_row + (int64_t)it * 4, acc);
    }
}

// ------------------------------------------------------------------------------------------------
// from_rgb: y[b,p,o] = lrelu(sum_c w[o,c] * scale * img[b,c,p] + bias[o]) * gain, img through strides
// one thread = 4 output channels of one pixel; a warp covers 128 channels of one pixel (or several pixels)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) from_rgb_fwd_kernel(const float* __restrict__ img, int64_t sb, int64_t sc,
                                                           int64_t sh, int64_t sw, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ out,
                                                           int B, int H, int W, int cin, int cout, int pitch,
                                                           float wscale, int act, float gain) {
    extern __shared__ float sw_[];          // [cin][pitch] effective weights, then bias [pitch]
    float* sbias = sw_ + cin * pitch;
    for (int i = threadIdx.x; i < cin * pitch; i += 256) {
        const int c = i / pitch, o = i - c * pitch;
        sw_[i] = (o < cout) ? wscale * __ldg(w + o * cin + c) : 0.f;
    }
    for (int i = threadIdx.x; i < pitch; i += 256) sbias[i] = (bias && i < cout) ? __ldg(bias + i) : 0.f;
    __syncthreads();
    const int c4n = pitch >> 2;
    const int rows = B * H;
    for (int row = blockIdx.y; row < rows; row += gridDim.y) {
        const int b = row / H, y = row - b * H;
        const float* ip_row = img + b * sb + y * sh;
        float* out_row = out + (int64_t)row * W * pitch;
        const int row_items = W * c4n;
        for (int it = blockIdx.x * 256 + threadIdx.x; it < row_items; it += gridDim.x * 256) {
            const int x = it / c4n, c4 = it - x * c4n;
            const float* ip = ip_row + x * sw;
            float4 acc = ld4(sbias + c4 * 4);
            for (int c = 0; c < cin; ++c) {
                const float v = __ldg(ip + c * sc);
                const float4 wv = ld4(sw_ + c * pitch + c4 * 4);
                acc.x = fmaf(v, wv.x, acc.x);
                acc.y = fmaf(v, wv.y, acc.y);
                acc.z = fmaf(v, wv.z, acc.z);
                acc.w = fmaf(v, wv.w, acc.w);
            }
            if (act) {
                acc.x = lrelu_gain(acc.x, gain); acc.y = lrelu_gain(acc.y, gain);
                acc.z = lrelu_gain(acc.z, gain); acc.w = lrelu_gain(acc.w, gain);
            }
            st4(out_row + (int64_t)it * 4, acc);
        }
    }
}

// g_img[b,c,p] = sum_o w[o,c] * scale * gain * mask(y[b,p,o]) * g[b,p,o];  8 lanes per pixel, 4 pixels per warp step
__global__ void __launch_bounds__(256) from_rgb_bwd_kernel(const float* __restrict__ g, const float* __restrict__ yact,
                                                           const float* __restrict__ w, float* __restrict__ gimg,
                                                           int B, int HW, int cin, int cout, int pitch, float wscale,
                                                           int act, float gain) {
    extern __shared__ float sw_[];          // [cin][pitch]
    for (int i = threadIdx.x; i < cin * pitch; i += 256) {
        const int c = i / pitch, o = i - c * pitch;
        sw_[i] = (o < cout) ? wscale * gain * __ldg(w + o * cin + c) : 0.f;
    }
    __syncthreads();
    const int lane8 = threadIdx.x & 7;
    const int grp = threadIdx.x >> 3;       // 32 pixels per CTA step
    const int c4n = pitch >> 2;
    const int64_t npix = (int64_t)B * HW;
    for (int64_t base = (int64_t)blockIdx.x * 32; base < npix; base += (int64_t)gridDim.x * 32) {   // warp-uniform
        const int64_t pix = base + grp;
        const bool valid = pix < npix;
        const float* gp = g + (valid ? pix : 0) * pitch;
        const float* yp = yact + (valid ? pix : 0) * pitch;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};                // cin <= 4
        for (int c4 = lane8; c4 < c4n; c4 += 8) {
            float4 gv = ldg4(gp + c4 * 4);
            if (act) {
                const float4 yv = ldg4(yp + c4 * 4);
                gv.x *= yv.x > 0.f ? 1.f : kLreluSlope; gv.y *= yv.y > 0.f ? 1.f : kLreluSlope;
                gv.z *= yv.z > 0.f ? 1.f : kLreluSlope; gv.w *= yv.w > 0.f ? 1.f : kLreluSlope;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c < cin) {
                    const float4 wv = ld4(sw_ + c * pitch + c4 * 4);
                    acc[c] = fmaf(gv.x, wv.x, acc[c]);
                    acc[c] = fmaf(gv.y, wv.y, acc[c]);
                    acc[c] = fmaf(gv.z, wv.z, acc[c]);
                    acc[c] = fmaf(gv.w, wv.w, acc[c]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 4);
            acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 2);
            acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
        }
        if (valid && lane8 < cin) {
            float v = acc[0];
#pragma unroll
            for (int c = 1; c < 4; ++c)
                if (lane8 == c) v = acc[c];
            const int64_t b = pix / HW, pp = pix - b * HW;
            gimg[(b * cin + lane8) * HW + pp] = v;          // NCHW contiguous image gradient
        }
    }
}

__global__ void __launch_bounds__(256) act_mask_nhwc_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                            float* __restrict__ out, int64_t n4, float gain) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
        float4 gv = ldg4(g + i * 4);
        const float4 yv = ldg4(y + i * 4);
        gv.x *= yv.x > 0.f ? gain : gain * kLreluSlope; gv.y *= yv.y > 0.f ? gain : gain * kLreluSlope;
        gv.z *= yv.z > 0.f ? gain : gain * kLreluSlope; gv.w *= yv.w > 0.f ? gain : gain * kLreluSlope;
        st4(out + i * 4, gv);
    }
}

static inline unsigned grid_1d(int64_t work_items, int per_block) {
    int64_t blocks = ceil_div<int64_t>(work_items, per_block);
    const int64_t cap = (int64_t)kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace cagc

using namespace cagc;

// launch_conv of conv_simt.cu
int cagc_simt_conv(cudaStream_t stream, const cagc::ConvP& p, const char* what);

extern "C" {

int cagc_fir_resample_nhwc(cagc_stream_t stream_, const float* in, const float* taps_host, float* out, int B, int in_h,
                           int in_w, int pitch, int up, int down, int pad0, int pad1) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(in && taps_host && out, "fir_resample_nhwc: null pointer");
    CAGC_REQUIRE(pitch > 0 && pitch % 4 == 0, "fir_resample_nhwc: pitch must be a positive multiple of 4");
    CAGC_REQUIRE((up == 1 && down == 2) || (up == 2 && down == 1), "fir_resample_nhwc: (up, down) must be (1,2) or (2,1)");
    CAGC_REQUIRE(aligned16(in) && aligned16(out), "fir_resample_nhwc: pointers must be 16-byte aligned");
    const int num_h = in_h * up + pad0 + pad1 - 4, num_w = in_w * up + pad0 + pad1 - 4;
    if (B <= 0 || num_h < 0 || num_w < 0) return 0;
    FirRsP p;
    p.in = in; p.out = out; p.B = B; p.in_h = in_h; p.in_w = in_w; p.pitch = pitch; p.pad_x0 = pad0; p.pad_y0 = pad0;
    p.out_h = num_h / down + 1; p.out_w = num_w / down + 1;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) p.kf[i * 4 + j] = taps_host[(3 - i) * 4 + (3 - j)];
    const int64_t rows = (int64_t)B * p.out_h;
    CAGC_REQUIRE(rows <= 65535, "fir_resample_nhwc: too many output rows (B * out_h <= 65535)");
    const int row_items = p.out_w * (pitch / 4);
    dim3 grid((unsigned)std::min(8, ceil_div(row_items, 256)), (unsigned)rows);
    if (up == 1)
        fir_resample_nhwc_kernel<1, 2><<<grid, 256, 0, stream>>>(p);
    else
        fir_resample_nhwc_kernel<2, 1><<<grid, 256, 0, stream>>>(p);
    return launched("fir_resample_nhwc_kernel");
}

int cagc_from_rgb_fwd(cagc_stream_t stream_, const float* img, int64_t sb, int64_t sc, int64_t sh, int64_t sw,
                      const float* w, const float* bias, float* out, int B, int H, int W, int cin, int cout, int pitch,
                      float wscale, int act, float gain) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(img && w && out, "from_rgb_fwd: null pointer");
    CAGC_REQUIRE(cin >= 1 && cin <= 4, "from_rgb_fwd: 1..4 input channels (got %d)", cin);
    CAGC_REQUIRE(pitch % 4 == 0 && cout <= pitch && pitch <= 2048, "from_rgb_fwd: bad channel pitch %d", pitch);
    CAGC_REQUIRE(aligned16(out), "from_rgb_fwd: output must be 16-byte aligned");
    const int64_t total = (int64_t)B * H * W * (pitch / 4);
    if (total == 0) return 0;
    const size_t smem = (size_t)(cin + 1) * pitch * sizeof(float);
    dim3 grid((unsigned)std::min(8, ceil_div(W * (pitch / 4), 256)), (unsigned)std::min(B * H, kNumSMs * 8));
    from_rgb_fwd_kernel<<<grid, 256, smem, stream>>>(img, sb, sc, sh, sw, w, bias, out, B, H, W, cin, cout, pitch, wscale,
                                                     act, gain);
    return launched("from_rgb_fwd_kernel");
}

int cagc_from_rgb_bwd(cagc_stream_t stream_, const float* g, const float* yact, const float* w, float* gimg, int B,
                      int H, int W, int cin, int cout, int pitch, float wscale, int act, float gain) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(g && w && gimg && (yact || !act), "from_rgb_bwd: null pointer");
    CAGC_REQUIRE(cin >= 1 && cin <= 4, "from_rgb_bwd: 1..4 input channels (got %d)", cin);
    CAGC_REQUIRE(pitch % 4 == 0 && cout <= pitch && pitch <= 2048, "from_rgb_bwd: bad channel pitch %d", pitch);
    CAGC_REQUIRE(aligned16(g) && (!yact || aligned16(yact)), "from_rgb_bwd: inputs must be 16-byte aligned");
    const int64_t npix = (int64_t)B * H * W;
    if (npix == 0) return 0;
    const size_t smem = (size_t)cin * pitch * sizeof(float);
    from_rgb_bwd_kernel<<<grid_1d(npix, 64), 256, smem, stream>>>(g, yact, w, gimg, B, H * W, cin, cout, pitch, wscale,
                                                                  act, gain);
    return launched("from_rgb_bwd_kernel");
}

int cagc_act_mask_nhwc(cagc_stream_t stream_, const float* g, const float* y, float* out, int64_t n, float gain) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(g && y && out, "act_mask_nhwc: null pointer");
    CAGC_REQUIRE(n % 4 == 0 && aligned16(g) && aligned16(y) && aligned16(out), "act_mask_nhwc: needs 16-byte aligned, 4-divisible buffers");
    if (n == 0) return 0;
    act_mask_nhwc_kernel<<<grid_1d(n / 4, 1024), 256, 0, stream>>>(g, y, out, n / 4, gain);
    return launched("act_mask_nhwc_kernel");
}

// Plain (un-modulated) convolution on NHWC-p with fused epilogue: out = act(conv(in, W) + bias) [+ residual].
//   mode 0: stride 1, zero padding ksize/2 (EqualConv2d 3x3 / 1x1, model.py:99-126; also its data gradient with
//           flipped / transposed slabs)
//   mode 1: stride 2, no padding: H = 2*Ho + ksize - 2 (the convolution after Blur in a downsampling ConvLayer,
//           model.py:683-700)
// Weight slabs as produced by cagc_weight_prep (tensor pipe: K-major [tap][roundup16(out_pitch)][in_pitch], TF32;
// SIMT: [tap][in_pitch][out_pitch]).  The data gradient of mode 1 is cagc_conv_up (transposed convolution).
int cagc_conv2d(cagc_stream_t stream, const float* in, const float* w_slabs, const float* bias, const float* residual,
                float* out, int B, int Hin, int Win, int in_pitch, int out_pitch, int out_valid, int ksize, int mode,
                int act, float act_gain, int algo) {
    return cagc_conv2d_ws(stream, in, w_slabs, bias, residual, out, B, Hin, Win, in_pitch, out_pitch, out_valid, ksize, mode,
                          act, act_gain, algo, nullptr, 0);
}

static int conv2d_impl(cagc_stream_t stream_, const float* in, const float* w_slabs, const float* bias, const float* residual,
                       float* out, int B, int Hin, int Win, int in_pitch, int out_pitch, int out_valid, int ksize, int mode,
                       int act, float act_gain, int algo, float* workspace, int64_t workspace_bytes);

int cagc_conv2d_ws(cagc_stream_t stream_, const float* in, const float* w_slabs, const float* bias, const float* residual,
                   float* out, int B, int Hin, int Win, int in_pitch, int out_pitch, int out_valid, int ksize, int mode,
                   int act, float act_gain, int algo, float* workspace, int64_t workspace_bytes) {
    return conv2d_impl(stream_, in, w_slabs, bias, residual, out, B, Hin, Win, in_pitch, out_pitch, out_valid, ksize, mode,
                       act ? kActLrelu : kActNone, act_gain, algo, workspace, workspace_bytes);
}

// out = conv(in, W) * (mask_ref > 0): a same-size data-gradient convolution with the backward of the ReLU that produced
// its input activation (mask_ref, layout of out) applied in the epilogue
int cagc_conv2d_mask_ws(cagc_stream_t stream_, const float* in, const float* w_slabs, const float* mask_ref, float* out,
                        int B, int Hin, int Win, int in_pitch, int out_pitch, int out_valid, int ksize, int algo,
                        float* workspace, int64_t workspace_bytes) {
    CAGC_REQUIRE(mask_ref, "conv2d_mask: null mask reference");
    return conv2d_impl(stream_, in, w_slabs, nullptr, mask_ref, out, B, Hin, Win, in_pitch, out_pitch, out_valid, ksize, 0,
                       kActMaskRef, 1.f, algo, workspace, workspace_bytes);
}

static int conv2d_impl(cagc_stream_t stream_, const float* in, const float* w_slabs, const float* bias, const float* residual,
                       float* out, int B, int Hin, int Win, int in_pitch, int out_pitch, int out_valid, int ksize, int mode,
                       int act, float act_gain, int algo, float* workspace, int64_t workspace_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(in && w_slabs && out, "conv2d: null pointer");
    CAGC_REQUIRE(B >= 0 && Hin >= 0 && Win >= 0, "conv2d: negative size");
    CAGC_REQUIRE(in_pitch > 0 && in_pitch % 4 == 0 && out_pitch > 0 && out_pitch % 4 == 0,
                 "conv2d: channel pitches must be positive multiples of 4 (got %d, %d)", in_pitch, out_pitch);
    CAGC_REQUIRE(ksize >= 1 && ksize * ksize <= kMaxTaps, "conv2d: unsupported kernel size %d", ksize);
    CAGC_REQUIRE(mode == 0 || mode == 1, "conv2d: mode must be 0 (stride 1, same) or 1 (stride 2, valid)");
    CAGC_REQUIRE(mode == 1 || ksize % 2 == 1, "conv2d: same-size convolution needs an odd kernel size");
    CAGC_REQUIRE(aligned16(in) && aligned16(w_slabs) && aligned16(out) && (!residual || aligned16(residual)),
                 "conv2d: pointers must be 16-byte aligned");
    CAGC_REQUIRE(residual != out, "conv2d: residual must not alias the output");
    ConvP p{};
    p.in = in; p.w = w_slabs; p.bias = bias; p.residual = residual; p.out = out;
    p.workspace = workspace; p.workspace_bytes = workspace_bytes;
    p.B = B; p.Hin = Hin; p.Win = Win; p.in_pitch = in_pitch;
    p.n_cols = out_pitch; p.out_valid = out_valid; p.out_stride = 1; p.out_oy = 0; p.out_ox = 0;
    p.act = act; p.act_gain = act_gain; p.ntaps = ksize * ksize;
    if (mode == 0) {
        p.Ho = Hin; p.Wo = Win; p.in_stride = 1;
        for (int ky = 0; ky < ksize; ++ky)
            for (int kx = 0; kx < ksize; ++kx) p.taps[ky * ksize + kx] = Tap{ky - ksize / 2, kx - ksize / 2, ky * ksize + kx};
    } else {
        CAGC_REQUIRE(Hin >= ksize && Win >= ksize, "conv2d: input smaller than the kernel");
        p.Ho = (Hin - ksize) / 2 + 1; p.Wo = (Win - ksize) / 2 + 1; p.in_stride = 2;
        for (int ky = 0; ky < ksize; ++ky)
            for (int kx = 0; kx < ksize; ++kx) p.taps[ky * ksize + kx] = Tap{ky, kx, ky * ksize + kx};
    }
    p.Hout = p.Ho; p.Wout = p.Wo;
    if ((int64_t)B * p.Ho * p.Wo == 0) return 0;
    if (algo == 1) return cagc_tc_conv(stream, p, "conv2d[tc]");
    return cagc_simt_conv(stream, p, "conv2d[simt]");
}

}  // extern "C"
