// Bandwidth ops of the StyleGAN2 hot path: upfirdn2d, fused bias + leaky ReLU (fwd / bwd with the
// bias-gradient reduction), layout conversion and the fused Adam bucket step.
//
// Replaces op/upfirdn2d_kernel.cu and op/fused_bias_act_kernel.cu of the reference (behaviour
// restated in oracle/stylegan2_oracle.py); written from scratch for sm_100a: 128-bit accesses where
// the layout allows, register sliding windows instead of one shared-memory read per tap, no
// per-element integer division in the fast paths, 64-bit indexing throughout.
#include "common.cuh"

namespace cagc {

thread_local char g_err[512] = {0};
std::atomic<int64_t> g_launches{0};

// ------------------------------------------------------------------------------------------------
// upfirdn2d, generic gather form: any up/down/pad (negative = crop)/kernel/minor.
// out[oy,ox] = sum_{i,j} k[kh-1-i][kw-1-j] * U[oy*down_y + i - pad_y0][ox*down_x + j - pad_x0],
// U[a][b] = in[a/up_y][b/up_x] when both divide and are in range, else 0  (op/upfirdn2d.py:159-200).
// ------------------------------------------------------------------------------------------------
struct UpfirdnP {
    int64_t major;
    int in_h, in_w, minor, kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_y0, out_h, out_w;
};

__global__ void __launch_bounds__(256) upfirdn2d_generic_kernel(const float* __restrict__ in,
                                                                const float* __restrict__ kern,
                                                                float* __restrict__ out, UpfirdnP p) {
    extern __shared__ float sk[];  // flipped taps
    for (int i = threadIdx.x; i < p.kh * p.kw; i += blockDim.x) {
        int ky = i / p.kw, kx = i % p.kw;
        sk[i] = kern[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];
    }
    __syncthreads();
    const int64_t total = p.major * p.out_h * p.out_w * p.minor;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int mi = (int)(idx % p.minor);
        int64_t t = idx / p.minor;
        int ox = (int)(t % p.out_w);
        t /= p.out_w;
        int oy = (int)(t % p.out_h);
        int64_t mj = t / p.out_h;
        const float* src = in + mj * p.in_h * p.in_w * p.minor + mi;
        const int ay0 = oy * p.down_y - p.pad_y0;
        const int ax0 = ox * p.down_x - p.pad_x0;
        float acc = 0.f;
        for (int i = 0; i < p.kh; ++i) {
            int a = ay0 + i;
            if (a < 0 || (a % p.up_y) != 0) continue;
            int iy = a / p.up_y;
            if (iy >= p.in_h) break;
            for (int j = 0; j < p.kw; ++j) {
                int b = ax0 + j;
                if (b < 0 || (b % p.up_x) != 0) continue;
                int ix = b / p.up_x;
                if (ix >= p.in_w) break;
                acc = fmaf(sk[i * p.kw + j], __ldg(src + ((int64_t)iy * p.in_w + ix) * p.minor), acc);
            }
        }
        out[idx] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// Fast path: up = down = 1, minor = 1, KH x KW taps (the Blur of model.py:80-96; every large
// upfirdn2d call of the generator and the discriminator).  One CTA = 32 x 128 outputs of one plane;
// the (32+KH-1) x (128+KW-1) input window is staged once in shared memory with coalesced loads;
// each thread then produces a 4 x 4 output block from a register sliding window: two 128-bit
// shared loads feed 64 FMAs.
// ------------------------------------------------------------------------------------------------
template <int KH, int KW>
__global__ void __launch_bounds__(256) fir_planes_kernel(const float* __restrict__ in,
                                                         const float* __restrict__ kern,
                                                         float* __restrict__ out, int in_h, int in_w, int out_h,
                                                         int out_w, int pad_x0, int pad_y0, int tiles_x,
                                                         int tiles_y, int vec_store) {
    constexpr int TW = 128, TH = 32;
    constexpr int SH = TH + KH - 1;
    constexpr int SWL = ((TW + KW - 1 + 3) / 4) * 4;  // columns actually loaded
    constexpr int SP = SWL + 4;                       // pitch (multiple of 4 floats)
    __shared__ __align__(16) float s[SH][SP];
    __shared__ float skf[KH * KW];

    int64_t bid = blockIdx.x;
    const int tile_x = (int)(bid % tiles_x);
    bid /= tiles_x;
    const int tile_y = (int)(bid % tiles_y);
    const int64_t plane = bid / tiles_y;

    const int oy0 = tile_y * TH, ox0 = tile_x * TW;
    const int iy0 = oy0 - pad_y0, ix0 = ox0 - pad_x0;
    const float* src = in + plane * (int64_t)in_h * in_w;

    if (threadIdx.x < KH * KW) {
        int ky = threadIdx.x / KW, kx = threadIdx.x % KW;
        skf[threadIdx.x] = kern[(KH - 1 - ky) * KW + (KW - 1 - kx)];
    }
    for (int idx = threadIdx.x; idx < SH * SWL; idx += 256) {
        int r = idx / SWL, c = idx - r * SWL;
        int iy = iy0 + r, ix = ix0 + c;
        float v = 0.f;
        if (iy >= 0 && iy < in_h && ix >= 0 && ix < in_w) v = __ldg(src + (int64_t)iy * in_w + ix);
        s[r][c] = v;
    }
    __syncthreads();

    float kf[KH * KW];
#pragma unroll
    for (int i = 0; i < KH * KW; ++i) kf[i] = skf[i];

    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    float acc[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[q][c] = 0.f;

#pragma unroll
    for (int r = 0; r < 4 + KH - 1; ++r) {
        float v[8];
        float4 v0 = *reinterpret_cast<const float4*>(&s[ty * 4 + r][tx * 4]);
        float4 v1 = *reinterpret_cast<const float4*>(&s[ty * 4 + r][tx * 4 + 4]);
        v[0] = v0.x; v[1] = v0.y; v[2] = v0.z; v[3] = v0.w;
        v[4] = v1.x; v[5] = v1.y; v[6] = v1.z; v[7] = v1.w;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = r - q;
            if (i >= 0 && i < KH) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int j = 0; j < KW; ++j) acc[q][c] = fmaf(kf[i * KW + j], v[c + j], acc[q][c]);
            }
        }
    }

    float* dst = out + plane * (int64_t)out_h * out_w;
    const int ox = ox0 + tx * 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int oy = oy0 + ty * 4 + q;
        if (oy >= out_h || ox >= out_w) continue;
        float* o = dst + (int64_t)oy * out_w + ox;
        if (vec_store && ox + 3 < out_w) {
            st4(o, make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]));
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (ox + c < out_w) o[c] = acc[q][c];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// fused bias + activation, (act, grad) switch of op/fused_bias_act_kernel.cu:29-46
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float bias_act_apply(float x, float ref, int act, int grad, float alpha, float scale) {
    float y;
    if (act == 1) {  // linear
        y = (grad == 2) ? 0.f : x;
    } else {  // leaky relu
        if (grad == 0) y = x > 0.f ? x : x * alpha;
        else if (grad == 1) y = ref > 0.f ? x : x * alpha;
        else y = 0.f;
    }
    return y * scale;
}

// MODE 0: scalar; MODE 1: float4, one bias per 4 elements (step_b % 4 == 0); MODE 2: float4, bias
// contiguous (step_b == 1, size_b % 4 == 0)
template <int MODE>
__global__ void __launch_bounds__(256) bias_act_kernel(const float* __restrict__ in, const float* __restrict__ bias,
                                                       const float* __restrict__ refer, float* __restrict__ out,
                                                       int64_t n, int64_t step_b, int size_b, int act, int grad,
                                                       float alpha, float scale) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (MODE == 0) {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            float x = in[i];
            if (bias) x += __ldg(bias + (i / step_b) % size_b);
            float r = refer ? refer[i] : 0.f;
            out[i] = bias_act_apply(x, r, act, grad, alpha, scale);
        }
    } else {
        const int64_t n4 = n >> 2;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 x = ld4(in + 4 * i);
            if (bias) {
                if (MODE == 1) {
                    float b = __ldg(bias + ((4 * i) / step_b) % size_b);
                    x.x += b; x.y += b; x.z += b; x.w += b;
                } else {
                    float4 b = ldg4(bias + (4 * i) % size_b);
                    x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
                }
            }
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (refer) r = ld4(refer + 4 * i);
            float4 y;
            y.x = bias_act_apply(x.x, r.x, act, grad, alpha, scale);
            y.y = bias_act_apply(x.y, r.y, act, grad, alpha, scale);
            y.z = bias_act_apply(x.z, r.z, act, grad, alpha, scale);
            y.w = bias_act_apply(x.w, r.w, act, grad, alpha, scale);
            st4(out + 4 * i, y);
        }
    }
}

constexpr int kBiasChunk = 4096;

// backward with per-chunk bias partial sums; layout [outer*size_b planes][step_b]
__global__ void __launch_bounds__(256) bias_act_bwd_kernel(const float* __restrict__ g, const float* __restrict__ refer,
                                                           float* __restrict__ gin, float* __restrict__ partial,
                                                           int64_t step_b, int chunks, float alpha, float scale,
                                                           int vec) {
    const int64_t plane = blockIdx.y;
    const int chunk = blockIdx.x;
    const int64_t base = plane * step_b;
    const int64_t lo = (int64_t)chunk * kBiasChunk;
    const int64_t hi = min(step_b, lo + (int64_t)kBiasChunk);
    float sum = 0.f;
    if (vec) {
        for (int64_t i = lo + threadIdx.x * 4; i < hi; i += 256 * 4) {
            float4 gv = ld4(g + base + i), rv = ld4(refer + base + i), o;
            o.x = (rv.x > 0.f ? gv.x : gv.x * alpha) * scale;
            o.y = (rv.y > 0.f ? gv.y : gv.y * alpha) * scale;
            o.z = (rv.z > 0.f ? gv.z : gv.z * alpha) * scale;
            o.w = (rv.w > 0.f ? gv.w : gv.w * alpha) * scale;
            st4(gin + base + i, o);
            sum += (o.x + o.y) + (o.z + o.w);
        }
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += 256) {
            float gv = g[base + i], rv = refer[base + i];
            float o = (rv > 0.f ? gv : gv * alpha) * scale;
            gin[base + i] = o;
            sum += o;
        }
    }
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w];
        partial[plane * chunks + chunk] = t;
    }
}

// ------------------------------------------------------------------------------------------------
// strided NCHW-ish -> NHWC-p
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) to_nhwc_transpose_kernel(const float* __restrict__ src, int64_t sb, int64_t sc,
                                                                int64_t sh, int64_t sw, float* __restrict__ dst,
                                                                int C, int H, int W, int pitch) {
    // tile: 32 pixels x 32 channels of sample blockIdx.z; loads run along pixels, stores along channels
    __shared__ float t[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int HW = H * W;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;  // 32 x 8
    for (int cc = ly; cc < 32; cc += 8) {
        int c = c0 + cc, p = p0 + lx;
        float v = 0.f;
        if (c < C && p < HW) {
            int y = p / W, x = p - y * W;
            v = __ldg(src + b * sb + c * sc + y * sh + x * sw);
        }
        t[cc][lx] = v;
    }
    __syncthreads();
    for (int pp = ly; pp < 32; pp += 8) {
        int p = p0 + pp, c = c0 + lx;
        if (p < HW && c < pitch) dst[((int64_t)b * HW + p) * pitch + c] = t[lx][pp];
    }
}

__global__ void __launch_bounds__(256) to_nhwc_direct_kernel(const float* __restrict__ src, int64_t sb, int64_t sc,
                                                             int64_t sh, int64_t sw, float* __restrict__ dst,
                                                             int64_t total, int C, int H, int W, int pitch) {
    // channel-fastest source: thread per destination element
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(idx % pitch);
        int64_t t = idx / pitch;
        int x = (int)(t % W);
        t /= W;
        int y = (int)(t % H);
        int64_t b = t / H;
        dst[idx] = (c < C) ? __ldg(src + b * sb + c * sc + y * sh + x * sw) : 0.f;
    }
}

// ------------------------------------------------------------------------------------------------
// fused Adam on a flat bucket
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                                                   float b1, float b2, float eps, float gscale, float bc1, float bc2,
                                                   const float* __restrict__ step_dev, int vec,
                                                   float* __restrict__ ema, float ema_decay) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (step_dev) {  // step count lives on the device so that a captured CUDA graph stays valid across replays
        const float t = __ldg(step_dev);
        bc1 = 1.f - powf(b1, t);
        bc2 = 1.f - powf(b2, t);
    }
    const float step = lr / bc1;
    const float inv_sqrt_bc2 = rsqrtf(bc2);
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        gg *= gscale;
        mm = b1 * mm + (1.f - b1) * gg;
        vv = b2 * vv + (1.f - b2) * gg * gg;
        float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
        pp -= step * (mm / denom);
    };
    // exponential moving average of the updated parameters (train.py:124-129 `accumulate`, called right after
    // g_optim.step() at :398): ema = ema * decay + (1 - decay) * p, in the same pass over the bucket
    auto avg = [&](float& ee, float pp) { ee = ee * ema_decay + (1.f - ema_decay) * pp; };
    if (vec) {
        const int64_t n4 = n >> 2;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 pv = ld4(p + 4 * i), gv = ld4(g + 4 * i), mv = ld4(m + 4 * i), vv = ld4(v + 4 * i);
            upd(pv.x, gv.x, mv.x, vv.x);
            upd(pv.y, gv.y, mv.y, vv.y);
            upd(pv.z, gv.z, mv.z, vv.z);
            upd(pv.w, gv.w, mv.w, vv.w);
            st4(p + 4 * i, pv);
            st4(m + 4 * i, mv);
            st4(v + 4 * i, vv);
            if (ema) {
                float4 ev = ld4(ema + 4 * i);
                avg(ev.x, pv.x); avg(ev.y, pv.y); avg(ev.z, pv.z); avg(ev.w, pv.w);
                st4(ema + 4 * i, ev);
            }
        }
        for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            upd(p[i], g[i], m[i], v[i]);
            if (ema) avg(ema[i], p[i]);
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            upd(p[i], g[i], m[i], v[i]);
            if (ema) avg(ema[i], p[i]);
        }
    }
}

static inline int grid_for(int64_t work_items, int per_block = 256, int max_blocks = kNumSMs * 16) {
    int64_t b = ceil_div<int64_t>(work_items, per_block);
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (int)b;
}

}  // namespace cagc

using namespace cagc;

extern "C" {

int cagc_abi_version(void) { return CAGC_ABI_VERSION; }
const char* cagc_last_error(void) { return g_err; }
int64_t cagc_launch_count(void) { return g_launches.load(); }

int cagc_upfirdn2d(cagc_stream_t stream_, const float* input, const float* kernel, float* output, int64_t major,
                   int in_h, int in_w, int minor, int kh, int kw, int up_x, int up_y, int down_x, int down_y,
                   int pad_x0, int pad_x1, int pad_y0, int pad_y1) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(input && kernel && output, "upfirdn2d: null pointer");
    CAGC_REQUIRE(major >= 0 && in_h >= 0 && in_w >= 0 && minor >= 1, "upfirdn2d: bad input size");
    CAGC_REQUIRE(kh >= 1 && kw >= 1 && kh * kw <= 4096, "upfirdn2d: bad kernel size %dx%d", kh, kw);
    CAGC_REQUIRE(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1, "upfirdn2d: up/down must be >= 1");
    const int num_h = in_h * up_y + pad_y0 + pad_y1 - kh;
    const int num_w = in_w * up_x + pad_x0 + pad_x1 - kw;
    if (num_h < 0 || num_w < 0 || major == 0) return 0;  // empty output
    const int out_h = num_h / down_y + 1;
    const int out_w = num_w / down_x + 1;

    if (up_x == 1 && up_y == 1 && down_x == 1 && down_y == 1 && minor == 1 && kh == 4 && kw == 4 &&
        (int64_t)out_h * out_w >= 256) {
        const int tiles_x = ceil_div(out_w, 128), tiles_y = ceil_div(out_h, 32);
        const int64_t blocks = major * tiles_x * tiles_y;
        CAGC_REQUIRE(blocks <= 0x7fffffffLL, "upfirdn2d: too many tiles");
        const int vec = (out_w % 4 == 0) && aligned16(output);
        fir_planes_kernel<4, 4><<<(unsigned)blocks, 256, 0, stream>>>(input, kernel, output, in_h, in_w, out_h,
                                                                       out_w, pad_x0, pad_y0, tiles_x, tiles_y, vec);
        return launched("fir_planes_kernel");
    }
    UpfirdnP p;
    p.major = major; p.in_h = in_h; p.in_w = in_w; p.minor = minor; p.kh = kh; p.kw = kw;
    p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y;
    p.pad_x0 = pad_x0; p.pad_y0 = pad_y0; p.out_h = out_h; p.out_w = out_w;
    const int64_t total = major * out_h * out_w * minor;
    upfirdn2d_generic_kernel<<<grid_for(total), 256, kh * kw * sizeof(float), stream>>>(input, kernel, output, p);
    return launched("upfirdn2d_generic_kernel");
}

int cagc_fused_bias_act(cagc_stream_t stream_, const float* input, const float* bias, const float* refer,
                        float* output, int64_t n, int64_t step_b, int size_b, int act, int grad, float alpha,
                        float scale) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return 0;
    CAGC_REQUIRE(input && output && n > 0, "fused_bias_act: null pointer");
    CAGC_REQUIRE(act == 1 || act == 3, "fused_bias_act: act must be 1 (linear) or 3 (lrelu), got %d", act);
    CAGC_REQUIRE(grad >= 0 && grad <= 2, "fused_bias_act: grad must be 0..2");
    CAGC_REQUIRE(grad != 1 || act == 1 || refer, "fused_bias_act: grad=1 needs refer");
    if (bias) CAGC_REQUIRE(step_b >= 1 && size_b >= 1, "fused_bias_act: bad bias geometry");
    const bool al = aligned16(input) && aligned16(output) && (!refer || aligned16(refer)) && (n % 4 == 0);
    if (al && (!bias || step_b % 4 == 0)) {
        bias_act_kernel<1><<<grid_for(n / 4), 256, 0, stream>>>(input, bias, refer, output, n, bias ? step_b : 4,
                                                               bias ? size_b : 1, act, grad, alpha, scale);
    } else if (al && bias && step_b == 1 && size_b % 4 == 0 && aligned16(bias)) {
        bias_act_kernel<2><<<grid_for(n / 4), 256, 0, stream>>>(input, bias, refer, output, n, step_b, size_b, act,
                                                               grad, alpha, scale);
    } else {
        bias_act_kernel<0><<<grid_for(n), 256, 0, stream>>>(input, bias, refer, output, n, bias ? step_b : 1,
                                                           bias ? size_b : 1, act, grad, alpha, scale);
    }
    return launched("bias_act_kernel");
}

int cagc_bias_grad_chunks(int64_t step_b) {
    if (step_b < 256) return 0;  // tiny inner extent: caller reduces grad_in itself
    return (int)ceil_div<int64_t>(step_b, kBiasChunk);
}

int cagc_fused_bias_act_bwd(cagc_stream_t stream_, const float* grad_out, const float* refer, float* grad_in,
                            float* bias_partial, int64_t outer, int size_b, int64_t step_b, float alpha, float scale) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(grad_out && refer && grad_in && bias_partial, "fused_bias_act_bwd: null pointer");
    const int chunks = cagc_bias_grad_chunks(step_b);
    CAGC_REQUIRE(chunks > 0, "fused_bias_act_bwd: step_b too small for the fused reduction");
    const int64_t planes = outer * size_b;
    if (planes == 0) return 0;
    CAGC_REQUIRE(planes <= 65535, "fused_bias_act_bwd: too many planes (%lld)", (long long)planes);
    const int vec = aligned16(grad_out) && aligned16(refer) && aligned16(grad_in) && (step_b % 4 == 0);
    dim3 grid(chunks, (unsigned)planes);
    bias_act_bwd_kernel<<<grid, 256, 0, stream>>>(grad_out, refer, grad_in, bias_partial, step_b, chunks, alpha, scale,
                                                  vec);
    return launched("bias_act_bwd_kernel");
}

int cagc_to_nhwc(cagc_stream_t stream_, const float* src, int64_t sb, int64_t sc, int64_t sh, int64_t sw, float* dst,
                 int B, int C, int H, int W, int pitch) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(src && dst, "to_nhwc: null pointer");
    CAGC_REQUIRE(pitch >= C && pitch % 4 == 0, "to_nhwc: pitch %d must be >= C and a multiple of 4", pitch);
    if (B == 0 || H == 0 || W == 0) return 0;
    if (sc == 1 || C == 1) {
        const int64_t total = (int64_t)B * H * W * pitch;
        to_nhwc_direct_kernel<<<grid_for(total), 256, 0, stream>>>(src, sb, sc, sh, sw, dst, total, C, H, W, pitch);
        return launched("to_nhwc_direct_kernel");
    }
    CAGC_REQUIRE(B <= 65535, "to_nhwc: batch too large");
    dim3 grid(ceil_div(H * W, 32), ceil_div(pitch, 32), B);
    to_nhwc_transpose_kernel<<<grid, 256, 0, stream>>>(src, sb, sc, sh, sw, dst, C, H, W, pitch);
    return launched("to_nhwc_transpose_kernel");
}

int cagc_adam_ema_step(cagc_stream_t stream_, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                       int64_t n, float lr, float beta1, float beta2, float eps, float grad_scale, float bias_corr1,
                       float bias_corr2, const float* step_dev, float* ema, float ema_decay) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return 0;
    CAGC_REQUIRE(param && grad && exp_avg && exp_avg_sq, "adam_step: null pointer");
    CAGC_REQUIRE(!ema || (ema_decay >= 0.f && ema_decay <= 1.f), "adam_step: ema decay must lie in [0, 1]");
    const int vec = aligned16(param) && aligned16(grad) && aligned16(exp_avg) && aligned16(exp_avg_sq) &&
                    (!ema || aligned16(ema));
    adam_kernel<<<grid_for(vec ? n / 4 + 1 : n), 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1,
                                                                    beta2, eps, grad_scale, bias_corr1, bias_corr2, step_dev,
                                                                    vec, ema, ema_decay);
    return launched("adam_kernel");
}

int cagc_adam_step(cagc_stream_t stream, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                   int64_t n, float lr, float beta1, float beta2, float eps, float grad_scale, float bias_corr1,
                   float bias_corr2, const float* step_dev) {
    return cagc_adam_ema_step(stream, param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, grad_scale, bias_corr1,
                              bias_corr2, step_dev, nullptr, 0.f);
}

}  // extern "C"
