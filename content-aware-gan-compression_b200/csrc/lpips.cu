// LPIPS (VGG16, 'net-lin') perceptual distance of the KD loss (reference train.py:172-182 ->
// lpips/__init__.py:13-41 -> lpips/networks_basic.py:26-92 -> lpips/pretrained_networks.py:97-137), frozen network:
// forward of both images and the data gradient towards the first one.
//
// The twelve 3x3 convolutions with 64..512 input channels run on the convolution engines of this library
// (cagc_conv2d with a ReLU epilogue: act_gain < 0); this file holds what is left, all of it bandwidth work on NHWC-p:
//
//   rgb_conv3x3_fwd/bwd   conv1_1 (3 -> 64, K = 27): ScalingLayer (networks_basic.py:94-101) on the operand load, bias +
//                         ReLU on the store; backward straight to the NCHW image gradient
//   maxpool2_nhwc         nn.MaxPool2d(2, 2) between the VGG slices
//   relu_pool_bwd         gz = ((a wins its 2x2 window ? g_pool : 0) + g_direct) * (a > 0): backward of MaxPool2d, the sum
//                         with the gradient that arrives at a tapped feature map, and the ReLU mask, in ONE pass
//   lpips_head_fwd/bwd    per tap: unit-normalise both feature maps over channels (lpips/__init__.py:43-45), squared
//                         difference, 1x1 'lin' layer, spatial mean (networks_basic.py:66-84); fixed-order reductions
#include <algorithm>

#include "common.cuh"

namespace cagc {

struct RgbAffine {
    float shift[3];
    float inv_scale[3];
};

// y[b,p,o] = relu(sum_{ky,kx,c} w[o,c,ky,kx] * xs[b,c,p + (ky-1, kx-1)] + bias[o]),  xs = (img - shift) / scale inside the
// image and 0 outside (the reference pads AFTER the scaling layer).  One thread = 4 output channels of one pixel.
__global__ void __launch_bounds__(256) rgb_conv3x3_fwd_kernel(const float* __restrict__ img, int64_t sb, int64_t sc,
                                                              int64_t sh, int64_t sw, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ out,
                                                              int B, int H, int W, int cout, int pitch, RgbAffine aff) {
    extern __shared__ float sw_[];          // [27][pitch]: row (ky*3+kx)*3+c, then bias [pitch]
    float* sbias = sw_ + 27 * pitch;
    for (int i = threadIdx.x; i < 27 * pitch; i += 256) {
        const int r = i / pitch, o = i - r * pitch;
        const int t = r / 3, c = r - t * 3;
        sw_[i] = (o < cout) ? __ldg(w + (o * 3 + c) * 9 + t) : 0.f;
    }
    for (int i = threadIdx.x; i < pitch; i += 256) sbias[i] = (bias && i < cout) ? __ldg(bias + i) : 0.f;
    __syncthreads();
    const int c4n = pitch >> 2;
    const int rows = B * H;
    for (int row = blockIdx.y; row < rows; row += gridDim.y) {
        const int b = row / H, y = row - b * H;
        const float* ip_img = img + b * sb;
        float* out_row = out + (int64_t)row * W * pitch;
        const int row_items = W * c4n;
        for (int it = blockIdx.x * 256 + threadIdx.x; it < row_items; it += gridDim.x * 256) {
            const int x = it / c4n, c4 = it - x * c4n;
            float4 acc = ld4(sbias + c4 * 4);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int yy = y + ky - 1;
                if (yy < 0 || yy >= H) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int xx = x + kx - 1;
                    if (xx < 0 || xx >= W) continue;
                    const float* ip = ip_img + yy * sh + xx * sw;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float v = (__ldg(ip + c * sc) - aff.shift[c]) * aff.inv_scale[c];
                        const float4 wv = ld4(sw_ + ((ky * 3 + kx) * 3 + c) * pitch + c4 * 4);
                        acc.x = fmaf(v, wv.x, acc.x);
                        acc.y = fmaf(v, wv.y, acc.y);
                        acc.z = fmaf(v, wv.z, acc.z);
                        acc.w = fmaf(v, wv.w, acc.w);
                    }
                }
            }
            acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
            st4(out_row + (int64_t)it * 4, acc);
        }
    }
}

// Register-blocked form for 32..256 output channels: one thread = 4 output channels of 8 ADJACENT pixels, so a weight
// vector is read from shared memory once per 8 pixels and an input value once per 3 taps; the scaled input rows of a
// segment (3 rows x (segment + 2) pixels x 3 channels, zero outside the image) are staged in shared memory.  The 16 (or
// 8, 32, 64) threads that share a pixel read the same input address (broadcast) and store 256 contiguous bytes.
template <int PX>
__global__ void __launch_bounds__(256) rgb_conv3x3_fwd_tiled_kernel(const float* __restrict__ img, int64_t sb, int64_t sc,
                                                                    int64_t sh, int64_t sw, const float* __restrict__ w,
                                                                    const float* __restrict__ bias, float* __restrict__ out,
                                                                    int B, int H, int W, int cout, int pitch, RgbAffine aff) {
    extern __shared__ float sw_[];          // [27][pitch] weights, [pitch] bias, [9][PW] input patch
    const int c4n = pitch >> 2;
    const int groups = 256 / c4n;           // pixel groups per CTA
    const int seg = groups * PX;            // pixels per segment
    const int PW = seg + 2;
    float* sbias = sw_ + 27 * pitch;
    float* s_in = sbias + pitch;
    for (int i = threadIdx.x; i < 27 * pitch; i += 256) {
        const int r = i / pitch, o = i - r * pitch;
        const int t = r / 3, c = r - t * 3;
        sw_[i] = (o < cout) ? __ldg(w + (o * 3 + c) * 9 + t) : 0.f;
    }
    for (int i = threadIdx.x; i < pitch; i += 256) sbias[i] = (bias && i < cout) ? __ldg(bias + i) : 0.f;
    const int c4 = threadIdx.x % c4n, xg = threadIdx.x / c4n;
    const int nseg = (W + seg - 1) / seg;
    const int64_t items = (int64_t)B * H * nseg;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
        const int sgi = (int)(it % nseg);
        const int64_t row = it / nseg;
        const int b = (int)(row / H), y = (int)(row - (int64_t)b * H);
        const int x0 = sgi * seg;
        __syncthreads();                    // previous patch fully consumed (and the weights visible on the first pass)
        for (int i = threadIdx.x; i < 9 * PW; i += 256) {
            const int r = i / PW, q = i - r * PW;
            const int c = r / 3, ky = r - c * 3;
            const int yy = y + ky - 1, xx = x0 + q - 1;
            float v = 0.f;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                const float sft = c == 0 ? aff.shift[0] : (c == 1 ? aff.shift[1] : aff.shift[2]);       // no local-memory copy
                const float isc = c == 0 ? aff.inv_scale[0] : (c == 1 ? aff.inv_scale[1] : aff.inv_scale[2]);
                v = (__ldg(img + b * sb + c * sc + yy * sh + xx * sw) - sft) * isc;
            }
            s_in[i] = v;
        }
        __syncthreads();
        float4 acc[PX];
        const float4 b4 = ld4(sbias + c4 * 4);
#pragma unroll
        for (int j = 0; j < PX; ++j) acc[j] = b4;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                float iv[PX + 2];
#pragma unroll
                for (int q = 0; q < PX + 2; ++q) iv[q] = s_in[(c * 3 + ky) * PW + xg * PX + q];
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float4 wv = ld4(sw_ + ((ky * 3 + kx) * 3 + c) * pitch + c4 * 4);
#pragma unroll
                    for (int j = 0; j < PX; ++j) {
                        acc[j].x = fmaf(iv[j + kx], wv.x, acc[j].x);
                        acc[j].y = fmaf(iv[j + kx], wv.y, acc[j].y);
                        acc[j].z = fmaf(iv[j + kx], wv.z, acc[j].z);
                        acc[j].w = fmaf(iv[j + kx], wv.w, acc[j].w);
                    }
                }
            }
        }
        float* orow = out + (row * W + x0 + xg * PX) * pitch + c4 * 4;
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            if (x0 + xg * PX + j < W) {
                float4 o = acc[j];
                o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                st4(orow + (int64_t)j * pitch, o);
            }
        }
    }
}

// g_img[b,c,Y,X] = inv_scale[c] * sum_{ky,kx,o} w[o,c,ky,kx] * gz[b, Y-ky+1, X-kx+1, o]   (gz already carries the ReLU
// mask of conv1_1).  8 lanes per pixel, 32 pixels per CTA step.
__global__ void __launch_bounds__(256) rgb_conv3x3_bwd_kernel(const float* __restrict__ gz, const float* __restrict__ w,
                                                              float* __restrict__ gimg, int B, int H, int W, int cout,
                                                              int pitch, RgbAffine aff) {
    extern __shared__ float sw_[];          // [27][pitch]
    for (int i = threadIdx.x; i < 27 * pitch; i += 256) {
        const int r = i / pitch, o = i - r * pitch;
        const int t = r / 3, c = r - t * 3;
        sw_[i] = (o < cout) ? __ldg(w + (o * 3 + c) * 9 + t) : 0.f;
    }
    __syncthreads();
    const int lane8 = threadIdx.x & 7;
    const int grp = threadIdx.x >> 3;
    const int c4n = pitch >> 2;
    const int HW = H * W;
    const int64_t npix = (int64_t)B * HW;
    for (int64_t base = (int64_t)blockIdx.x * 32; base < npix; base += (int64_t)gridDim.x * 32) {   // warp-uniform
        const int64_t pix = base + grp;
        const bool valid = pix < npix;
        const int64_t pv = valid ? pix : 0;
        const int b = (int)(pv / HW);
        const int r = (int)(pv - (int64_t)b * HW);
        const int Y = r / W, X = r - Y * W;
        float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = Y - ky + 1;
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int xx = X - kx + 1;
                if (xx < 0 || xx >= W) continue;
                const float* gp = gz + (((int64_t)b * H + yy) * W + xx) * pitch;
                const float* wp = sw_ + (ky * 3 + kx) * 3 * pitch;
                for (int c4 = lane8; c4 < c4n; c4 += 8) {
                    const float4 gv = ldg4(gp + c4 * 4);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float4 wv = ld4(wp + c * pitch + c4 * 4);
                        acc[c] = fmaf(gv.x, wv.x, acc[c]);
                        acc[c] = fmaf(gv.y, wv.y, acc[c]);
                        acc[c] = fmaf(gv.z, wv.z, acc[c]);
                        acc[c] = fmaf(gv.w, wv.w, acc[c]);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 4);
            acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 2);
            acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
        }
        if (valid && lane8 < 3) {
            const float v = lane8 == 0 ? acc[0] : (lane8 == 1 ? acc[1] : acc[2]);
            const float is = lane8 == 0 ? aff.inv_scale[0] : (lane8 == 1 ? aff.inv_scale[1] : aff.inv_scale[2]);
            gimg[((int64_t)b * 3 + lane8) * HW + r] = v * is;     // NCHW contiguous image gradient
        }
    }
}

// Same result, four adjacent pixels per 8-lane group (W % 4 == 0): per (ky, channel chunk) the six gradient vectors the
// four pixels touch are loaded once and each weight vector serves four pixels -- 2.3x fewer load/store-unit operations
// per pixel than the one-pixel form, which that unit (not DRAM) bounded.
__global__ void __launch_bounds__(256) rgb_conv3x3_bwd4_kernel(const float* __restrict__ gz, const float* __restrict__ w,
                                                               float* __restrict__ gimg, int B, int H, int W, int cout,
                                                               int pitch, RgbAffine aff) {
    extern __shared__ float sw_[];          // [27][pitch]
    for (int i = threadIdx.x; i < 27 * pitch; i += 256) {
        const int r = i / pitch, o = i - r * pitch;
        const int t = r / 3, c = r - t * 3;
        sw_[i] = (o < cout) ? __ldg(w + (o * 3 + c) * 9 + t) : 0.f;
    }
    __syncthreads();
    const int lane8 = threadIdx.x & 7;
    const int grp = threadIdx.x >> 3;
    const int c4n = pitch >> 2;
    const int W4 = W >> 2;
    const int64_t HW = (int64_t)H * W;
    const int64_t nquad = (int64_t)B * H * W4;
    for (int64_t base = (int64_t)blockIdx.x * 32; base < nquad; base += (int64_t)gridDim.x * 32) {   // warp-uniform
        const int64_t quad = base + grp;
        const bool valid = quad < nquad;
        const int64_t qv = valid ? quad : 0;
        const int X0 = (int)(qv % W4) * 4;
        const int64_t by = qv / W4;                    // b * H + Y
        const int Y = (int)(by % H);
        const int64_t brow = by - Y;                   // b * H
        float acc[4][3];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = Y - ky + 1;
            if (yy < 0 || yy >= H) continue;
            const float* grow = gz + (brow + yy) * W * pitch;
            for (int c4 = lane8; c4 < c4n; c4 += 8) {
                float4 gv[6];
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    const int xx = X0 - 1 + q;
                    gv[q] = (xx >= 0 && xx < W) ? ldg4(grow + (int64_t)xx * pitch + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float* wp = sw_ + (ky * 3 + kx) * 3 * pitch + c4 * 4;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float4 wv = ld4(wp + c * pitch);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {       // source pixel of output X0 + j under tap kx: X0 + j - kx + 1
                            const float4 g4 = gv[j - kx + 2];
                            acc[j][c] = fmaf(g4.x, wv.x, acc[j][c]);
                            acc[j][c] = fmaf(g4.y, wv.y, acc[j][c]);
                            acc[j][c] = fmaf(g4.z, wv.z, acc[j][c]);
                            acc[j][c] = fmaf(g4.w, wv.w, acc[j][c]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                acc[j][c] += __shfl_xor_sync(0xffffffffu, acc[j][c], 4);
                acc[j][c] += __shfl_xor_sync(0xffffffffu, acc[j][c], 2);
                acc[j][c] += __shfl_xor_sync(0xffffffffu, acc[j][c], 1);
            }
        if (valid && lane8 < 3) {
            const float is = lane8 == 0 ? aff.inv_scale[0] : (lane8 == 1 ? aff.inv_scale[1] : aff.inv_scale[2]);
            float4 o;
            o.x = (lane8 == 0 ? acc[0][0] : (lane8 == 1 ? acc[0][1] : acc[0][2])) * is;
            o.y = (lane8 == 0 ? acc[1][0] : (lane8 == 1 ? acc[1][1] : acc[1][2])) * is;
            o.z = (lane8 == 0 ? acc[2][0] : (lane8 == 1 ? acc[2][1] : acc[2][2])) * is;
            o.w = (lane8 == 0 ? acc[3][0] : (lane8 == 1 ? acc[3][1] : acc[3][2])) * is;
            const int64_t b = brow / H;
            st4(gimg + (b * 3 + lane8) * HW + (int64_t)Y * W + X0, o);      // NCHW contiguous image gradient
        }
    }
}

__device__ __forceinline__ float4 max4(float4 a, float4 b) {
    return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

// out[b,yo,xo,:] = max over the 2x2 window; H, W even
__global__ void __launch_bounds__(256) maxpool2_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int Ho,
                                                            int Wo, int c4n, int64_t items) {
    const int64_t in_row = (int64_t)2 * Wo * c4n * 4;        // floats per input row
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < items; i += (int64_t)gridDim.x * 256) {
        const int c4 = (int)(i % c4n);
        const int64_t q = i / c4n;
        const int xo = (int)(q % Wo);
        const int64_t by = q / Wo;                            // b * Ho + yo
        const float* p = in + by * 2 * in_row + ((int64_t)2 * xo * c4n + c4) * 4;
        const float4 m = max4(max4(ldg4(p), ldg4(p + c4n * 4)), max4(ldg4(p + in_row), ldg4(p + in_row + c4n * 4)));
        st4(out + i * 4, m);
    }
}

// one component of the 2x2 window: first maximum in scan order wins (ATen's max_pool2d: `val > maxval`), then the sum with
// the direct gradient and the ReLU mask of the activation itself
__device__ __forceinline__ void pool_bwd_1(float a00, float a01, float a10, float a11, float gp, float& o00, float& o01,
                                           float& o10, float& o11) {
    int idx = 0;
    float m = a00;
    if (a01 > m) { m = a01; idx = 1; }
    if (a10 > m) { m = a10; idx = 2; }
    if (a11 > m) { m = a11; idx = 3; }
    o00 = ((idx == 0 ? gp : 0.f) + o00) * (a00 > 0.f ? 1.f : 0.f);
    o01 = ((idx == 1 ? gp : 0.f) + o01) * (a01 > 0.f ? 1.f : 0.f);
    o10 = ((idx == 2 ? gp : 0.f) + o10) * (a10 > 0.f ? 1.f : 0.f);
    o11 = ((idx == 3 ? gp : 0.f) + o11) * (a11 > 0.f ? 1.f : 0.f);
}

__global__ void __launch_bounds__(256) relu_pool_bwd_kernel(const float* __restrict__ act, const float* __restrict__ g_pool,
                                                            const float* __restrict__ g_direct, float* __restrict__ out,
                                                            int Ho, int Wo, int c4n, int64_t items) {
    const int64_t in_row = (int64_t)2 * Wo * c4n * 4;
    const int px = c4n * 4;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < items; i += (int64_t)gridDim.x * 256) {
        const int c4 = (int)(i % c4n);
        const int64_t q = i / c4n;
        const int xo = (int)(q % Wo);
        const int64_t by = q / Wo;
        const int64_t off = by * 2 * in_row + ((int64_t)2 * xo * c4n + c4) * 4;
        const float4 a00 = ldg4(act + off), a01 = ldg4(act + off + px), a10 = ldg4(act + off + in_row),
                     a11 = ldg4(act + off + in_row + px);
        const float4 gp = ldg4(g_pool + i * 4);
        float4 o00 = make_float4(0.f, 0.f, 0.f, 0.f), o01 = o00, o10 = o00, o11 = o00;
        if (g_direct) {
            o00 = ldg4(g_direct + off); o01 = ldg4(g_direct + off + px);
            o10 = ldg4(g_direct + off + in_row); o11 = ldg4(g_direct + off + in_row + px);
        }
        pool_bwd_1(a00.x, a01.x, a10.x, a11.x, gp.x, o00.x, o01.x, o10.x, o11.x);
        pool_bwd_1(a00.y, a01.y, a10.y, a11.y, gp.y, o00.y, o01.y, o10.y, o11.y);
        pool_bwd_1(a00.z, a01.z, a10.z, a11.z, gp.z, o00.z, o01.z, o10.z, o11.z);
        pool_bwd_1(a00.w, a01.w, a10.w, a11.w, gp.w, o00.w, o01.w, o10.w, o11.w);
        st4(out + off, o00); st4(out + off + px, o01); st4(out + off + in_row, o10); st4(out + off + in_row + px, o11);
    }
}

// out = g * (a > 0)
__global__ void __launch_bounds__(256) relu_mask_kernel(const float* __restrict__ act, const float* __restrict__ g,
                                                        float* __restrict__ out, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
        float4 gv = ldg4(g + i * 4);
        const float4 a = ldg4(act + i * 4);
        gv.x = a.x > 0.f ? gv.x : 0.f; gv.y = a.y > 0.f ? gv.y : 0.f;
        gv.z = a.z > 0.f ? gv.z : 0.f; gv.w = a.w > 0.f ? gv.w : 0.f;
        st4(out + i * 4, gv);
    }
}

// ------------------------------------------------------------------------------------------------
// LPIPS head of one tapped layer.  C = 32 * NV channels, 8 lanes per pixel, each lane keeps NV float4 of both feature
// vectors in registers.  normalize_tensor (lpips/__init__.py:43-45): f = x / (sqrt(sum_c x^2) + 1e-10).
// ------------------------------------------------------------------------------------------------
constexpr float kLpipsEps = 1e-10f;

__device__ __forceinline__ float sum8(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

template <int NV>
__global__ void __launch_bounds__(256) lpips_head_fwd_kernel(const float* __restrict__ fs, const float* __restrict__ ft,
                                                             const float* __restrict__ lin_w, float* __restrict__ partial,
                                                             int HW) {
    constexpr int C = NV * 32;
    __shared__ float red[32];
    const int lane8 = threadIdx.x & 7, grp = threadIdx.x >> 3;
    const int b = blockIdx.y;
    float acc = 0.f;
    for (int base = blockIdx.x * 32; base < HW; base += gridDim.x * 32) {          // warp-uniform
        const int p = base + grp;
        const bool valid = p < HW;
        const int64_t off = ((int64_t)b * HW + (valid ? p : 0)) * C;
        float4 x[NV], t[NV];
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            x[v] = ldg4(fs + off + (lane8 + 8 * v) * 4);
            t[v] = ldg4(ft + off + (lane8 + 8 * v) * 4);
            s0 += dot4(x[v], x[v]);
            s1 += dot4(t[v], t[v]);
        }
        s0 = sum8(s0); s1 = sum8(s1);
        const float i0 = 1.f / (sqrtf(s0) + kLpipsEps), i1 = 1.f / (sqrtf(s1) + kLpipsEps);
        float d = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const float4 u = make_float4(x[v].x * i0 - t[v].x * i1, x[v].y * i0 - t[v].y * i1, x[v].z * i0 - t[v].z * i1,
                                         x[v].w * i0 - t[v].w * i1);
            const float4 w4 = ldg4(lin_w + (lane8 + 8 * v) * 4);        // 2 KB at most: L1 resident
            d += w4.x * u.x * u.x + w4.y * u.y * u.y + w4.z * u.z * u.z + w4.w * u.w * u.w;
        }
        d = sum8(d);
        if (valid) acc += d;           // identical in the 8 lanes of the pixel
    }
    if (lane8 == 0) red[grp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < 32; ++i) s += red[i];
        partial[(int64_t)b * gridDim.x + blockIdx.x] = s;
    }
}

// val[b] = (accumulate ? val[b] : 0) + sum_blk partial[b][blk] / HW, fixed order
__global__ void lpips_head_finalize_kernel(const float* __restrict__ partial, float* __restrict__ val, int B, int nblk,
                                           float inv_hw, int accumulate) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float s = 0.f;
    for (int i = 0; i < nblk; ++i) s += partial[(int64_t)b * nblk + i];
    val[b] = (accumulate ? val[b] : 0.f) + s * inv_hw;
}

// gs[b,p,c] = d val[b] / d fs[b,p,c] * gval[b]
template <int NV>
__global__ void __launch_bounds__(256) lpips_head_bwd_kernel(const float* __restrict__ fs, const float* __restrict__ ft,
                                                             const float* __restrict__ lin_w, const float* __restrict__ gval,
                                                             float* __restrict__ gs, int HW) {
    constexpr int C = NV * 32;
    const int lane8 = threadIdx.x & 7, grp = threadIdx.x >> 3;
    const int b = blockIdx.y;
    const float k = 2.f * __ldg(gval + b) / (float)HW;
    for (int base = blockIdx.x * 32; base < HW; base += gridDim.x * 32) {          // warp-uniform
        const int p = base + grp;
        const bool valid = p < HW;
        const int64_t off = ((int64_t)b * HW + (valid ? p : 0)) * C;
        float4 x[NV], t[NV];
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            x[v] = ldg4(fs + off + (lane8 + 8 * v) * 4);
            t[v] = ldg4(ft + off + (lane8 + 8 * v) * 4);
            s0 += dot4(x[v], x[v]);
            s1 += dot4(t[v], t[v]);
        }
        s0 = sum8(s0); s1 = sum8(s1);
        const float n0 = sqrtf(s0);
        const float i0 = 1.f / (n0 + kLpipsEps), i1 = 1.f / (sqrtf(s1) + kLpipsEps);
        float dot = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {      // t[v] <- gf = k * w * (x i0 - t i1): gradient w.r.t. the normalised feature
            const float4 w4 = ldg4(lin_w + (lane8 + 8 * v) * 4);
            t[v].x = k * w4.x * (x[v].x * i0 - t[v].x * i1);
            t[v].y = k * w4.y * (x[v].y * i0 - t[v].y * i1);
            t[v].z = k * w4.z * (x[v].z * i0 - t[v].z * i1);
            t[v].w = k * w4.w * (x[v].w * i0 - t[v].w * i1);
            dot += dot4(t[v], x[v]);
        }
        dot = sum8(dot);
        // d/dx of x / (|x| + eps): gf / (|x| + eps) - x * <gf, x> / (|x| (|x| + eps)^2); an all-zero feature vector (the
        // reference's sqrt backward gives 0/0 there) gets the first term only
        const float coef = n0 > 0.f ? dot * i0 * i0 / n0 : 0.f;
        if (valid) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const float4 o = make_float4(t[v].x * i0 - x[v].x * coef, t[v].y * i0 - x[v].y * coef,
                                             t[v].z * i0 - x[v].z * coef, t[v].w * i0 - x[v].w * coef);
                st4(gs + off + (lane8 + 8 * v) * 4, o);
            }
        }
    }
}

static inline unsigned lp_grid_1d(int64_t work_items, int per_block) {
    int64_t blocks = ceil_div<int64_t>(work_items, per_block);
    const int64_t cap = (int64_t)kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

static inline int head_blocks(int B, int HW) {
    // enough CTAs to fill the machine, at most one per 32 pixels, the same for forward and backward
    int per_img = std::max(1, (kNumSMs * 8 + std::max(B, 1) - 1) / std::max(B, 1));
    return std::max(1, std::min(ceil_div(HW, 32), std::min(per_img, 256)));
}

static inline RgbAffine make_affine(const float* shift_host, const float* scale_host) {
    RgbAffine a;
    for (int c = 0; c < 3; ++c) {
        a.shift[c] = shift_host ? shift_host[c] : 0.f;
        a.inv_scale[c] = scale_host ? 1.f / scale_host[c] : 1.f;
    }
    return a;
}

}  // namespace cagc

using namespace cagc;

extern "C" {

int cagc_rgb_conv3x3_fwd(cagc_stream_t stream_, const float* img, int64_t sb, int64_t sc, int64_t sh, int64_t sw,
                         const float* w, const float* bias, const float* shift_host, const float* scale_host, float* out,
                         int B, int H, int W, int cout, int pitch) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(img && w && out, "rgb_conv3x3_fwd: null pointer");
    CAGC_REQUIRE(pitch % 4 == 0 && cout >= 1 && cout <= pitch && pitch <= 256, "rgb_conv3x3_fwd: bad channel pitch %d", pitch);
    CAGC_REQUIRE(aligned16(out), "rgb_conv3x3_fwd: output must be 16-byte aligned");
    CAGC_REQUIRE(B >= 0 && H >= 0 && W >= 0, "rgb_conv3x3_fwd: negative size");
    if ((int64_t)B * H * W == 0) return 0;
    const int c4n = pitch / 4;
    if (c4n >= 8 && 256 % c4n == 0) {
        constexpr int PX = 8;
        const int seg = (256 / c4n) * PX;
        const size_t smem_t = ((size_t)28 * pitch + 9 * (seg + 2)) * sizeof(float);
        const int64_t items = (int64_t)B * H * ceil_div(W, seg);
        const unsigned blocks = (unsigned)std::min<int64_t>(items, (int64_t)kNumSMs * 6);
        rgb_conv3x3_fwd_tiled_kernel<PX><<<blocks, 256, smem_t, stream>>>(img, sb, sc, sh, sw, w, bias, out, B, H, W, cout,
                                                                          pitch, make_affine(shift_host, scale_host));
        return launched("rgb_conv3x3_fwd_tiled_kernel");
    }
    const size_t smem = (size_t)28 * pitch * sizeof(float);
    dim3 grid((unsigned)std::min(8, ceil_div(W * (pitch / 4), 256)), (unsigned)std::min(B * H, kNumSMs * 8));
    rgb_conv3x3_fwd_kernel<<<grid, 256, smem, stream>>>(img, sb, sc, sh, sw, w, bias, out, B, H, W, cout, pitch,
                                                        make_affine(shift_host, scale_host));
    return launched("rgb_conv3x3_fwd_kernel");
}

int cagc_rgb_conv3x3_bwd(cagc_stream_t stream_, const float* gz, const float* w, const float* scale_host, float* gimg, int B,
                         int H, int W, int cout, int pitch) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(gz && w && gimg, "rgb_conv3x3_bwd: null pointer");
    CAGC_REQUIRE(pitch % 4 == 0 && cout >= 1 && cout <= pitch && pitch <= 256, "rgb_conv3x3_bwd: bad channel pitch %d", pitch);
    CAGC_REQUIRE(aligned16(gz), "rgb_conv3x3_bwd: input must be 16-byte aligned");
    CAGC_REQUIRE(B >= 0 && H >= 0 && W >= 0, "rgb_conv3x3_bwd: negative size");
    const int64_t npix = (int64_t)B * H * W;
    if (npix == 0) return 0;
    const size_t smem = (size_t)27 * pitch * sizeof(float);
    if (W % 4 == 0 && aligned16(gimg)) {
        rgb_conv3x3_bwd4_kernel<<<lp_grid_1d(npix / 4, 32), 256, smem, stream>>>(gz, w, gimg, B, H, W, cout, pitch,
                                                                                 make_affine(nullptr, scale_host));
        return launched("rgb_conv3x3_bwd4_kernel");
    }
    rgb_conv3x3_bwd_kernel<<<lp_grid_1d(npix, 64), 256, smem, stream>>>(gz, w, gimg, B, H, W, cout, pitch,
                                                                        make_affine(nullptr, scale_host));
    return launched("rgb_conv3x3_bwd_kernel");
}

int cagc_maxpool2_nhwc(cagc_stream_t stream_, const float* in, float* out, int B, int H, int W, int pitch) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(in && out, "maxpool2_nhwc: null pointer");
    CAGC_REQUIRE(pitch > 0 && pitch % 4 == 0, "maxpool2_nhwc: pitch must be a positive multiple of 4");
    CAGC_REQUIRE(B >= 0 && H >= 0 && W >= 0 && H % 2 == 0 && W % 2 == 0, "maxpool2_nhwc: H and W must be even (got %d x %d)", H, W);
    CAGC_REQUIRE(aligned16(in) && aligned16(out), "maxpool2_nhwc: pointers must be 16-byte aligned");
    const int64_t items = (int64_t)B * (H / 2) * (W / 2) * (pitch / 4);
    if (items == 0) return 0;
    maxpool2_nhwc_kernel<<<lp_grid_1d(items, 512), 256, 0, stream>>>(in, out, H / 2, W / 2, pitch / 4, items);
    return launched("maxpool2_nhwc_kernel");
}

int cagc_relu_pool_bwd(cagc_stream_t stream_, const float* act, const float* g_pool, const float* g_direct, float* out,
                       int B, int H, int W, int pitch) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(act && out && (g_pool || g_direct), "relu_pool_bwd: null pointer");
    CAGC_REQUIRE(pitch > 0 && pitch % 4 == 0, "relu_pool_bwd: pitch must be a positive multiple of 4");
    CAGC_REQUIRE(B >= 0 && H >= 0 && W >= 0, "relu_pool_bwd: negative size");
    CAGC_REQUIRE(aligned16(act) && aligned16(out) && (!g_pool || aligned16(g_pool)) && (!g_direct || aligned16(g_direct)),
                 "relu_pool_bwd: pointers must be 16-byte aligned");
    if (g_pool) {
        CAGC_REQUIRE(H % 2 == 0 && W % 2 == 0, "relu_pool_bwd: H and W must be even (got %d x %d)", H, W);
        const int64_t items = (int64_t)B * (H / 2) * (W / 2) * (pitch / 4);
        if (items == 0) return 0;
        relu_pool_bwd_kernel<<<lp_grid_1d(items, 256), 256, 0, stream>>>(act, g_pool, g_direct, out, H / 2, W / 2, pitch / 4,
                                                                         items);
        return launched("relu_pool_bwd_kernel");
    }
    const int64_t n4 = (int64_t)B * H * W * (pitch / 4);
    if (n4 == 0) return 0;
    relu_mask_kernel<<<lp_grid_1d(n4, 1024), 256, 0, stream>>>(act, g_direct, out, n4);
    return launched("relu_mask_kernel");
}

int cagc_lpips_head_blocks(int B, int HW) { return head_blocks(B, HW); }

int cagc_lpips_head_fwd(cagc_stream_t stream_, const float* fs, const float* ft, const float* lin_w, float* partial,
                        float* val, int B, int HW, int C, int accumulate) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(fs && ft && lin_w && partial && val, "lpips_head_fwd: null pointer");
    CAGC_REQUIRE(C == 64 || C == 128 || C == 256 || C == 512, "lpips_head_fwd: 64, 128, 256 or 512 channels (got %d)", C);
    CAGC_REQUIRE(aligned16(fs) && aligned16(ft) && aligned16(lin_w), "lpips_head_fwd: pointers must be 16-byte aligned");
    CAGC_REQUIRE(B >= 0 && B <= 65535 && HW >= 0, "lpips_head_fwd: bad size");
    if (B == 0) return 0;
    const int nblk = head_blocks(B, HW);
    if (HW > 0) {
        dim3 grid((unsigned)nblk, (unsigned)B);
        switch (C) {
            case 64: lpips_head_fwd_kernel<2><<<grid, 256, 0, stream>>>(fs, ft, lin_w, partial, HW); break;
            case 128: lpips_head_fwd_kernel<4><<<grid, 256, 0, stream>>>(fs, ft, lin_w, partial, HW); break;
            case 256: lpips_head_fwd_kernel<8><<<grid, 256, 0, stream>>>(fs, ft, lin_w, partial, HW); break;
            default: lpips_head_fwd_kernel<16><<<grid, 256, 0, stream>>>(fs, ft, lin_w, partial, HW); break;
        }
        CAGC_TRY(launched("lpips_head_fwd_kernel"));
    }
    lpips_head_finalize_kernel<<<ceil_div(B, 128), 128, 0, stream>>>(partial, val, B, HW > 0 ? nblk : 0,
                                                                     HW > 0 ? 1.f / (float)HW : 0.f, accumulate);
    return launched("lpips_head_finalize_kernel");
}

int cagc_lpips_head_bwd(cagc_stream_t stream_, const float* fs, const float* ft, const float* lin_w, const float* gval,
                        float* gs, int B, int HW, int C) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(fs && ft && lin_w && gval && gs, "lpips_head_bwd: null pointer");
    CAGC_REQUIRE(C == 64 || C == 128 || C == 256 || C == 512, "lpips_head_bwd: 64, 128, 256 or 512 channels (got %d)", C);
    CAGC_REQUIRE(aligned16(fs) && aligned16(ft) && aligned16(lin_w) && aligned16(gs),
                 "lpips_head_bwd: pointers must be 16-byte aligned");
    CAGC_REQUIRE(B >= 0 && B <= 65535 && HW >= 0, "lpips_head_bwd: bad size");
    if (B == 0 || HW == 0) return 0;
    dim3 grid((unsigned)head_blocks(B, HW), (unsigned)B);
    switch (C) {
        case 64: lpips_head_bwd_kernel<2><<<grid, 256, 0, stream>>>(fs, ft, lin_w, gval, gs, HW); break;
        case 128: lpips_head_bwd_kernel<4><<<grid, 256, 0, stream>>>(fs, ft, lin_w, gval, gs, HW); break;
        case 256: lpips_head_bwd_kernel<8><<<grid, 256, 0, stream>>>(fs, ft, lin_w, gval, gs, HW); break;
        default: lpips_head_bwd_kernel<16><<<grid, 256, 0, stream>>>(fs, ft, lin_w, gval, gs, HW); break;
    }
    return launched("lpips_head_bwd_kernel");
}

}  // extern "C"
