// Small "everything around the convolution" kernels of the modulated-convolution layer
// (model.py:241-289).  Each replaces a chain of 5-15 elementwise / reduction /
// tiny-GEMM library launches of the unfused formulation with ONE launch:
//
//   weight_prep        W[O,I,k,k] -> operand slabs for the forward and data-gradient convolutions
//                      (both layouts, zero padded, optionally TF32-rounded), Wsq[o,i] = sum_t (cW)^2 in
//                      both orientations, and the zero-padded activation bias
//   style_affine       s_l[b,i] = scale_l * <A_l[i,:], latent[b, idx_l, :]> + bias_l[i] for ALL modulated
//                      convolutions of a generator in one launch (model.py:248 x 20 layers), + backward
//   demod              d[b,o] = rsqrt(sum_i s^2 Wsq + eps)                       (model.py:251-253)
//   act_bwd_finalize   chunk partials of act_bwd -> g_bias, gq = -1/2 d^3 gd, noise-weight partials
//   style_grad_finalize  g_s = sum_chunks (gx~ . x) + 2 s * (gq @ Wsq)
//   wgrad_finalize     split-K partial slabs -> W.grad in the reference layout [O,I,k,k], including the
//                      demodulation term 2 c^2 W * (gq^T s^2)
//   torgb_bwd_finalize g_w, g_s of the 1x1 modulated ToRGB convolution from its chunk partials
//   linear_bias_act    epilogue of EqualLinear (model.py:156-166): act(acc*scale + bias*lr_mul), + backward
//
// All reductions run in a fixed order (deterministic).  Sizes are tiny (<= a few MB): the point of these
// kernels is launch count, not bandwidth.
#include "common.cuh"

namespace cagc {

__device__ __forceinline__ float round_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// ------------------------------------------------------------------------------------------------
struct WeightPrepP {
    const float* w;       // [O][I][kk]
    float wscale;
    int O, I, kk, ksize;
    float* outA;          // [kk][RA][CA], element (t, o, i)   (nullable)
    int RA, CA, flipA;
    float* outB;          // [kk][RB][CB], element (t, i, o)   (nullable)
    int RB, CB, flipB;
    int round;            // round slab values to TF32
    float* wsq_oi;        // [SO][SI] (nullable)
    float* wsq_io;        // [SI][SO] (nullable)
    int SO, SI;
    const float* bias_in; // [bias_n] (nullable)
    float* bias_out;      // [bias_np]
    int bias_n, bias_np;
    int Omax, Imax;
};

__global__ void __launch_bounds__(256) weight_prep_kernel(const WeightPrepP p) {
    const int64_t total = (int64_t)p.Omax * p.Imax;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int o = (int)(idx / p.Imax), i = (int)(idx - (int64_t)o * p.Imax);
        const bool real = o < p.O && i < p.I;
        const float* src = p.w + ((int64_t)o * p.I + i) * p.kk;
        float sq = 0.f;
        for (int t = 0; t < p.kk; ++t) {
            float v = real ? __ldg(src + t) * p.wscale : 0.f;
            sq = fmaf(v, v, sq);
            if (p.round) v = round_tf32(v);
            if (p.outA && o < p.RA && i < p.CA) {
                const int ta = p.flipA ? p.kk - 1 - t : t;
                p.outA[((int64_t)ta * p.RA + o) * p.CA + i] = v;
            }
            if (p.outB && i < p.RB && o < p.CB) {
                const int tb = p.flipB ? p.kk - 1 - t : t;
                p.outB[((int64_t)tb * p.RB + i) * p.CB + o] = v;
            }
        }
        if (o < p.SO && i < p.SI) {
            if (p.wsq_oi) p.wsq_oi[(int64_t)o * p.SI + i] = sq;
            if (p.wsq_io) p.wsq_io[(int64_t)i * p.SO + o] = sq;
        }
        if (p.bias_out && idx < p.bias_np) p.bias_out[idx] = (p.bias_in && idx < p.bias_n) ? __ldg(p.bias_in + idx) : 0.f;
    }
}

// ------------------------------------------------------------------------------------------------
constexpr int kMaxStyleLayers = 40;

struct StyleP {
    const float* A[kMaxStyleLayers];      // [I][D]
    const float* bias[kMaxStyleLayers];   // [I] or null
    float* out[kMaxStyleLayers];          // fwd: s [B][pin];  bwd: unused
    const float* gs[kMaxStyleLayers];     // bwd: g_s [B][pin] or null
    float* gA[kMaxStyleLayers];           // bwd: [I][D]
    float* gbias[kMaxStyleLayers];        // bwd: [I] or null
    int I[kMaxStyleLayers], pin[kMaxStyleLayers], lat[kMaxStyleLayers], blk0[kMaxStyleLayers + 1];
    float scale[kMaxStyleLayers], bias_mul[kMaxStyleLayers];
    int n_layers, B, D, n_latent;
    const float* latent;                   // [B][n_latent][D] with strides (lat_sb, lat_sl, 1)
    int64_t lat_sb, lat_sl;
    float* g_latent;                       // bwd: [B][n_latent][D] contiguous
    int param_blocks;                      // bwd: blocks [0, param_blocks) do parameter grads, the rest latent grads
};

__device__ __forceinline__ int style_find_layer(const StyleP& p, int blk) {
    int l = 0;
    while (l + 1 < p.n_layers && blk >= p.blk0[l + 1]) ++l;
    return l;
}

// one warp per output channel; 8 channels per block
__global__ void __launch_bounds__(256) style_affine_kernel(const __grid_constant__ StyleP p) {
    const int l = style_find_layer(p, blockIdx.x);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = (blockIdx.x - p.blk0[l]) * 8 + warp;
    const int pin = p.pin[l];
    if (i >= pin) return;
    float* out = p.out[l];
    if (i >= p.I[l]) {
        for (int b = lane; b < p.B; b += 32) out[(int64_t)b * pin + i] = 0.f;
        return;
    }
    const float* a = p.A[l] + (int64_t)i * p.D;
    const float bv = p.bias[l] ? __ldg(p.bias[l] + i) * p.bias_mul[l] : 0.f;
    const float* lat = p.latent + (int64_t)p.lat[l] * p.lat_sl;
    if (p.D <= 1024) {
        // the weight row lives in registers (all of its loads in flight at once); per sample the latent row is read
        // with independent loads as well -- a plain dot-product loop serialises on one load latency per element
        float ar[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) ar[k] = (lane + 32 * k < p.D) ? __ldg(a + lane + 32 * k) : 0.f;
        for (int b = 0; b < p.B; ++b) {
            const float* w = lat + (int64_t)b * p.lat_sb;
            float wv[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) wv[k] = (lane + 32 * k < p.D) ? __ldg(w + lane + 32 * k) : 0.f;
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 32; ++k) acc = fmaf(ar[k], wv[k], acc);
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
            if (lane == 0) out[(int64_t)b * pin + i] = fmaf(acc, p.scale[l], bv);
        }
        return;
    }
    for (int b = 0; b < p.B; ++b) {
        const float* w = lat + (int64_t)b * p.lat_sb;
        float acc = 0.f;
        for (int j = lane; j < p.D; j += 32) acc = fmaf(__ldg(a + j), __ldg(w + j), acc);
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
        if (lane == 0) out[(int64_t)b * pin + i] = fmaf(acc, p.scale[l], bv);
    }
}

// blocks [0, param_blocks): gA[i][:] = scale * sum_b gs[b,i] * latent[b, idx, :],  gbias[i] = bias_mul * sum_b gs[b,i]
// blocks [param_blocks, +B*n_latent): g_latent[b, idx, :] = sum_{l: lat_l == idx} scale_l * sum_i gs_l[b,i] * A_l[i,:]
__global__ void __launch_bounds__(256) style_affine_bwd_kernel(const __grid_constant__ StyleP p) {
    extern __shared__ float sg[];
    if ((int)blockIdx.x < p.param_blocks) {
        const int l = style_find_layer(p, blockIdx.x);
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int i = (blockIdx.x - p.blk0[l]) * 8 + warp;
        if (i >= p.I[l]) return;
        const float* gs = p.gs[l];
        float* ga = p.gA[l] + (int64_t)i * p.D;
        if (!gs) {
            for (int j = lane; j < p.D; j += 32) ga[j] = 0.f;
            if (lane == 0 && p.gbias[l]) p.gbias[l][i] = 0.f;
            return;
        }
        const int pin = p.pin[l];
        const float* lat = p.latent + (int64_t)p.lat[l] * p.lat_sl;
        for (int j0 = 0; j0 < p.D; j0 += 128) {      // 4 columns per lane per pass
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
            for (int b = 0; b < p.B; ++b) {
                const float g = __ldg(gs + (int64_t)b * pin + i);
                const float* w = lat + (int64_t)b * p.lat_sb;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = j0 + q * 32 + lane;
                    if (j < p.D) acc[q] = fmaf(g, __ldg(w + j), acc[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = j0 + q * 32 + lane;
                if (j < p.D) ga[j] = acc[q] * p.scale[l];
            }
        }
        if (lane == 0 && p.gbias[l]) {
            float t = 0.f;
            for (int b = 0; b < p.B; ++b) t += __ldg(gs + (int64_t)b * pin + i);
            p.gbias[l][i] = t * p.bias_mul[l];
        }
        return;
    }
    const int item = blockIdx.x - p.param_blocks;
    const int b = item / p.n_latent, idx = item - b * p.n_latent;
    float* dst = p.g_latent + ((int64_t)b * p.n_latent + idx) * p.D;
    for (int j0 = 0; j0 < p.D; j0 += 256) {
        const int j = j0 + threadIdx.x;
        float acc = 0.f;
        for (int l = 0; l < p.n_layers; ++l) {
            if (p.lat[l] != idx || !p.gs[l]) continue;   // block-uniform
            const int I = p.I[l];
            const int I8 = (I + 7) & ~7;                 // the staged row is zero padded to a multiple of 8 ...
            __syncthreads();
            for (int i = threadIdx.x; i < I8; i += 256)
                sg[i] = i < I ? __ldg(p.gs[l] + (int64_t)b * p.pin[l] + i) * p.scale[l] : 0.f;
            __syncthreads();
            if (j < p.D) {
                const float* a = p.A[l] + j;
                float t = 0.f;
                for (int i0 = 0; i0 < I8; i0 += 8) {     // ... so the 8-wide body (8 loads in flight) needs no remainder
                    float av[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) av[u] = __ldg(a + (int64_t)min(i0 + u, I - 1) * p.D);
#pragma unroll
                    for (int u = 0; u < 8; ++u) t = fmaf(sg[i0 + u], av[u], t);
                }
                acc += t;
            }
        }
        if (j < p.D) dst[j] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// d[b][o] = rsqrt(sum_i s[b][i]^2 * wsq_io[i][o] + eps)  (o < O), 0 for the padding channels
// block = (32 output channels, 8 slices of the input-channel sum); grid = (ceil(pout/32), B)
__global__ void __launch_bounds__(256) demod_kernel(const float* __restrict__ s, const float* __restrict__ wsq_io,
                                                    float* __restrict__ d, int pin, int O, int pout, float eps) {
    extern __shared__ float s2[];     // [pin] squares, then [8][32] partial sums
    float* red = s2 + pin;
    const int b = blockIdx.y;
    const int cx = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int o = blockIdx.x * 32 + cx;
    for (int i = threadIdx.x; i < pin; i += 256) {
        const float v = s[(int64_t)b * pin + i];
        s2[i] = v * v;
    }
    __syncthreads();
    float acc = 0.f;
    if (o < O) {
        const float* wp = wsq_io + o;
#pragma unroll 4
        for (int i = sl; i < pin; i += 8) acc = fmaf(s2[i], __ldg(wp + (int64_t)i * pout), acc);
    }
    red[sl * 32 + cx] = acc;
    __syncthreads();
    if (sl == 0 && o < pout) {
        float t = 0.f;
        if (o < O) {
#pragma unroll
            for (int k = 0; k < 8; ++k) t += red[k * 32 + cx];
            t = rsqrtf(t + eps);
        }
        d[(int64_t)b * pout + o] = t;
    }
}

// ------------------------------------------------------------------------------------------------
// partial [B][chunks][3][P] -> g_bias[c] = sum_{b,ch} p0;  gq[b][c] = -1/2 d^3 * sum_ch p1;  nw_part[block] = sum p2
// block = (32 channels, 8 batch lanes); grid = ceil(P / 32)
__global__ void __launch_bounds__(256) act_bwd_finalize_kernel(const float* __restrict__ partial,
                                                               const float* __restrict__ d, float* __restrict__ g_bias,
                                                               float* __restrict__ gq, float* __restrict__ nw_part,
                                                               int B, int chunks, int P) {
    __shared__ float red[2][8][32];
    const int cx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    float sb = 0.f, sn = 0.f;
    if (c < P) {
        for (int b = ty; b < B; b += 8) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
            const float* pp = partial + (int64_t)b * chunks * 3 * P + c;
            for (int ch = 0; ch < chunks; ++ch) {
                s0 += pp[((int64_t)ch * 3 + 0) * P];
                s1 += pp[((int64_t)ch * 3 + 1) * P];
                s2 += pp[((int64_t)ch * 3 + 2) * P];
            }
            sb += s0;
            sn += s2;
            if (gq) {
                const float dv = d[(int64_t)b * P + c];
                gq[(int64_t)b * P + c] = -0.5f * dv * dv * dv * s1;
            }
        }
    }
    red[0][ty][cx] = sb;
    red[1][ty][cx] = sn;
    __syncthreads();
    if (ty == 0) {
        float t0 = 0.f, t1 = 0.f;
        for (int k = 0; k < 8; ++k) { t0 += red[0][k][cx]; t1 += red[1][k][cx]; }
        if (g_bias && c < P) g_bias[c] = t0;
        if (nw_part) {
            // fixed-order sum over the 32 channels of this block
            red[1][0][cx] = t1;
            __syncwarp();
            if (cx == 0) {
                float t = 0.f;
                for (int k = 0; k < 32; ++k) t += red[1][0][k];
                nw_part[blockIdx.x] = t;
            }
        }
    }
}

// g_s[b][i] = sum_ch mpartial[b][ch][i] + 2 s[b][i] * sum_o gq[b][o] * wsq_oi[o][i]
// block = (32 input channels, 8 slices of the chunk / output-channel sums); grid = (ceil(pin/32), B)
__global__ void __launch_bounds__(256) style_grad_finalize_kernel(const float* __restrict__ mpartial,
                                                                  const float* __restrict__ gq,
                                                                  const float* __restrict__ s,
                                                                  const float* __restrict__ wsq_oi,
                                                                  float* __restrict__ g_s, int chunks, int pin,
                                                                  int O, int pout) {
    extern __shared__ float sq[];     // [O] gq row, then [2][8][32] partial sums
    float* red = sq + (gq ? O : 0);
    const int b = blockIdx.y;
    const int cx = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + cx;
    if (gq) {
        for (int o = threadIdx.x; o < O; o += 256) sq[o] = gq[(int64_t)b * pout + o];
        __syncthreads();
    }
    float acc = 0.f, t = 0.f;
    if (i < pin) {
        const float* mp = mpartial + (int64_t)b * chunks * pin + i;
        for (int ch = sl; ch < chunks; ch += 8) acc += mp[(int64_t)ch * pin];
        if (gq) {
            const float* wp = wsq_oi + i;
#pragma unroll 4
            for (int o = sl; o < O; o += 8) t = fmaf(sq[o], __ldg(wp + (int64_t)o * pin), t);
        }
    }
    red[sl * 32 + cx] = acc;
    red[256 + sl * 32 + cx] = t;
    __syncthreads();
    if (sl == 0 && i < pin) {
        float a = 0.f, q = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { a += red[k * 32 + cx]; q += red[256 + k * 32 + cx]; }
        if (gq) a = fmaf(2.f * s[(int64_t)b * pin + i], q, a);
        g_s[(int64_t)b * pin + i] = a;
    }
}

// out[o][i][t] = c * sum_sp wpart[sp][t][i][o] + 2 c^2 W[o][i][t] * sum_b gq[b][o] s[b][i]^2
// one thread per (t, i, o), o fastest: coalesced reads of the partial slabs
__global__ void __launch_bounds__(256) wgrad_finalize_kernel(const float* __restrict__ wpart, int nsp,
                                                             const float* __restrict__ w, float wscale,
                                                             const float* __restrict__ gq,
                                                             const float* __restrict__ s, int B, int O, int I, int kk,
                                                             int pin, int pout, float* __restrict__ out) {
    const int64_t per_tap = (int64_t)O * I;
    const int64_t total = per_tap * kk;
    const int64_t slab = (int64_t)pin * pout;
    const int64_t sp_stride = (int64_t)kk * slab;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(idx / per_tap);
        const int64_t r = idx - (int64_t)t * per_tap;
        const int i = (int)(r / O), o = (int)(r - (int64_t)i * O);
        const float* pp = wpart + (int64_t)t * slab + (int64_t)i * pout + o;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int sp = 0;
        for (; sp + 4 <= nsp; sp += 4) {      // fixed association: (sp%4) lanes, combined in order below
            a0 += pp[(int64_t)(sp + 0) * sp_stride];
            a1 += pp[(int64_t)(sp + 1) * sp_stride];
            a2 += pp[(int64_t)(sp + 2) * sp_stride];
            a3 += pp[(int64_t)(sp + 3) * sp_stride];
        }
        for (; sp < nsp; ++sp) a0 += pp[(int64_t)sp * sp_stride];
        float v = ((a0 + a1) + (a2 + a3)) * wscale;
        if (gq) {
            float dem = 0.f;
            for (int b = 0; b < B; ++b) {
                const float sv = s[(int64_t)b * pin + i];
                dem = fmaf(gq[(int64_t)b * pout + o], sv * sv, dem);
            }
            v = fmaf(2.f * wscale * wscale * dem, __ldg(w + ((int64_t)o * I + i) * kk + t), v);
        }
        out[((int64_t)o * I + i) * kk + t] = v;
    }
}

// ToRGB: t[b][o][i] = sum_ch partial[b][ch][o][i];  g_s[b][i] = c sum_o t w[o][i];  gw_part[b][o][i] = c t s[b][i]
// (the caller adds gw_part over b).  block = (32 channels, 8 chunk lanes); grid = (ceil(pin/32), B)
__global__ void __launch_bounds__(256) torgb_bwd_finalize_kernel(const float* __restrict__ partial,
                                                                 const float* __restrict__ s,
                                                                 const float* __restrict__ w, float wscale,
                                                                 float* __restrict__ gw_part, float* __restrict__ g_s,
                                                                 int chunks, int cin, int pin, int nout) {
    __shared__ float red[4][8][32];
    const int cx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + cx, b = blockIdx.y;
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    if (i < pin) {
        for (int ch = ty; ch < chunks; ch += 8) {
            const float* pp = partial + (((int64_t)b * chunks + ch) * nout) * pin + i;
            for (int o = 0; o < nout; ++o) t[o] += pp[(int64_t)o * pin];
        }
    }
    for (int o = 0; o < 4; ++o) red[o][ty][cx] = t[o];
    __syncthreads();
    if (ty == 0 && i < pin) {
        const float sv = s[(int64_t)b * pin + i];
        float gs = 0.f;
        for (int o = 0; o < nout; ++o) {
            float v = 0.f;
            for (int k = 0; k < 8; ++k) v += red[o][k][cx];
            if (i < cin) {
                gs = fmaf(v, __ldg(w + o * cin + i), gs);
                if (gw_part) gw_part[((int64_t)b * nout + o) * cin + i] = v * sv * wscale;
            }
        }
        if (g_s) g_s[(int64_t)b * pin + i] = gs * wscale;
    }
}

// ------------------------------------------------------------------------------------------------
// EqualLinear epilogue: out = act(acc*acc_scale + bias*bias_scale) (* gain when act)
__global__ void __launch_bounds__(256) linear_bias_act_kernel(const float* __restrict__ acc,
                                                              const float* __restrict__ bias, float* __restrict__ out,
                                                              int64_t total, int N, float acc_scale, float bias_scale,
                                                              int act, float alpha, float gain) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(idx % N);
        float v = acc[idx] * acc_scale;
        if (bias) v = fmaf(__ldg(bias + n), bias_scale, v);
        if (act) v = (v > 0.f ? v : v * alpha) * gain;
        out[idx] = v;
    }
}

// g_acc[m][n] = g[m][n] * (act ? gain*(out>0 ? 1 : alpha) : 1) * acc_scale;  g_bias[n] = bias_scale * sum_m (...)/acc_scale
__global__ void __launch_bounds__(256) linear_bias_act_bwd_kernel(const float* __restrict__ g,
                                                                  const float* __restrict__ out,
                                                                  float* __restrict__ g_acc, float* __restrict__ g_bias,
                                                                  int M, int N, float acc_scale, float bias_scale,
                                                                  int act, float alpha, float gain) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float sum = 0.f;
    for (int m = 0; m < M; ++m) {
        const int64_t idx = (int64_t)m * N + n;
        float v = g[idx];
        if (act) v *= (out[idx] > 0.f ? gain : gain * alpha);
        sum += v;
        g_acc[idx] = v * acc_scale;
    }
    if (g_bias) g_bias[n] = sum * bias_scale;
}

static unsigned grid_for(int64_t n, int cap = kNumSMs * 8) {
    int64_t b = ceil_div<int64_t>(n, 256);
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace cagc

using namespace cagc;

extern "C" {

int cagc_weight_prep(cagc_stream_t stream_, const float* w, float wscale, int O, int I, int ksize, float* outA, int RA,
                     int CA, int flipA, float* outB, int RB, int CB, int flipB, int round_tf32, float* wsq_oi,
                     float* wsq_io, int SO, int SI, const float* bias_in, float* bias_out, int bias_n, int bias_np) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(w && O > 0 && I > 0 && ksize >= 1 && ksize <= 5, "weight_prep: bad weight");
    CAGC_REQUIRE(!outA || (RA >= O && CA >= I), "weight_prep: slab A smaller than the weight");
    CAGC_REQUIRE(!outB || (RB >= I && CB >= O), "weight_prep: slab B smaller than the weight");
    CAGC_REQUIRE(!(wsq_oi || wsq_io) || (SO >= O && SI >= I), "weight_prep: wsq smaller than the weight");
    CAGC_REQUIRE(!bias_out || bias_np >= bias_n, "weight_prep: bad bias padding");
    WeightPrepP p{};
    p.w = w; p.wscale = wscale; p.O = O; p.I = I; p.ksize = ksize; p.kk = ksize * ksize;
    p.outA = outA; p.RA = RA; p.CA = CA; p.flipA = flipA;
    p.outB = outB; p.RB = RB; p.CB = CB; p.flipB = flipB;
    p.round = round_tf32;
    p.wsq_oi = wsq_oi; p.wsq_io = wsq_io; p.SO = SO; p.SI = SI;
    p.bias_in = bias_in; p.bias_out = bias_out; p.bias_n = bias_n; p.bias_np = bias_np;
    p.Omax = O; p.Imax = I;
    if (outA) { p.Omax = max(p.Omax, RA); p.Imax = max(p.Imax, CA); }
    if (outB) { p.Omax = max(p.Omax, CB); p.Imax = max(p.Imax, RB); }
    if (wsq_oi || wsq_io) { p.Omax = max(p.Omax, SO); p.Imax = max(p.Imax, SI); }
    CAGC_REQUIRE(!bias_out || (int64_t)p.Omax * p.Imax >= bias_np, "weight_prep: bias longer than the thread domain");
    weight_prep_kernel<<<grid_for((int64_t)p.Omax * p.Imax), 256, 0, stream>>>(p);
    return launched("weight_prep_kernel");
}

static int fill_style(StyleP& p, int n_layers, const void* const* A, const void* const* bias, const int* I,
                      const int* pin, const int* lat, const float* scale, const float* bias_mul, const float* latent,
                      int64_t lat_sb, int64_t lat_sl, int B, int D, int n_latent, const char* what) {
    CAGC_REQUIRE(n_layers >= 1 && n_layers <= kMaxStyleLayers, "%s: 1..%d layers per call", what, kMaxStyleLayers);
    CAGC_REQUIRE(latent && B >= 0 && D > 0 && n_latent > 0, "%s: bad latent", what);
    p.n_layers = n_layers; p.B = B; p.D = D; p.n_latent = n_latent;
    p.latent = latent; p.lat_sb = lat_sb; p.lat_sl = lat_sl;
    int blk = 0;
    for (int l = 0; l < n_layers; ++l) {
        CAGC_REQUIRE(A[l] && I[l] > 0 && pin[l] >= I[l], "%s: bad layer %d", what, l);
        CAGC_REQUIRE(lat[l] >= 0 && lat[l] < n_latent, "%s: layer %d latent index %d out of range", what, l, lat[l]);
        p.A[l] = (const float*)A[l];
        p.bias[l] = bias ? (const float*)bias[l] : nullptr;
        p.I[l] = I[l]; p.pin[l] = pin[l]; p.lat[l] = lat[l];
        p.scale[l] = scale[l]; p.bias_mul[l] = bias_mul[l];
        p.blk0[l] = blk;
        blk += ceil_div(pin[l], 8);
    }
    p.blk0[n_layers] = blk;
    return 0;
}

int cagc_style_affine(cagc_stream_t stream_, int n_layers, const void* const* A, const void* const* bias,
                      void* const* out, const int* I, const int* pin, const int* lat, const float* scale,
                      const float* bias_mul, const float* latent, int64_t lat_sb, int64_t lat_sl, int B, int D,
                      int n_latent) {
    cudaStream_t stream = (cudaStream_t)stream_;
    StyleP p{};
    CAGC_TRY(fill_style(p, n_layers, A, bias, I, pin, lat, scale, bias_mul, latent, lat_sb, lat_sl, B, D, n_latent,
                        "style_affine"));
    for (int l = 0; l < n_layers; ++l) {
        CAGC_REQUIRE(out[l], "style_affine: null output %d", l);
        p.out[l] = (float*)out[l];
    }
    if (B == 0) return 0;
    style_affine_kernel<<<p.blk0[n_layers], 256, 0, stream>>>(p);
    return launched("style_affine_kernel");
}

int cagc_style_affine_bwd(cagc_stream_t stream_, int n_layers, const void* const* A, const void* const* gs,
                          void* const* gA, void* const* gbias, const int* I, const int* pin, const int* lat,
                          const float* scale, const float* bias_mul, const float* latent, int64_t lat_sb,
                          int64_t lat_sl, float* g_latent, int B, int D, int n_latent) {
    cudaStream_t stream = (cudaStream_t)stream_;
    StyleP p{};
    CAGC_TRY(fill_style(p, n_layers, A, nullptr, I, pin, lat, scale, bias_mul, latent, lat_sb, lat_sl, B, D, n_latent,
                        "style_affine_bwd"));
    int max_i = 1;
    for (int l = 0; l < n_layers; ++l) {
        p.gs[l] = (const float*)gs[l];
        p.gA[l] = gA ? (float*)gA[l] : nullptr;
        p.gbias[l] = gbias ? (float*)gbias[l] : nullptr;
        max_i = max(max_i, I[l]);
    }
    const bool params = gA != nullptr;
    if (params)
        for (int l = 0; l < n_layers; ++l) CAGC_REQUIRE(p.gA[l], "style_affine_bwd: null gA %d", l);
    p.param_blocks = params ? p.blk0[n_layers] : 0;
    p.g_latent = g_latent;
    const int lat_blocks = g_latent ? B * n_latent : 0;
    if (p.param_blocks + lat_blocks == 0) return 0;
    const size_t smem = sizeof(float) * (((size_t)max_i + 7) & ~(size_t)7);
    CAGC_REQUIRE(smem <= 48 * 1024, "style_affine_bwd: layer too wide");
    style_affine_bwd_kernel<<<p.param_blocks + lat_blocks, 256, smem, stream>>>(p);
    return launched("style_affine_bwd_kernel");
}

int cagc_demod(cagc_stream_t stream_, const float* s, const float* wsq_io, float* d, int B, int pin, int O, int pout,
               float eps) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(s && wsq_io && d && pin > 0 && pout >= O && O > 0, "demod: bad arguments");
    CAGC_REQUIRE(sizeof(float) * (pin + 256) <= 48 * 1024, "demod: too many input channels");
    if (B == 0) return 0;
    CAGC_REQUIRE(B <= 65535, "demod: batch too large");
    demod_kernel<<<dim3(ceil_div(pout, 32), B), 256, sizeof(float) * (pin + 256), stream>>>(s, wsq_io, d, pin, O, pout,
                                                                                        eps);
    return launched("demod_kernel");
}

int cagc_act_bwd_finalize_blocks(int pitch) { return ceil_div(pitch, 32); }

int cagc_act_bwd_finalize(cagc_stream_t stream_, const float* partial, const float* d, float* g_bias, float* gq,
                          float* nw_part, int B, int chunks, int pitch) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(partial && B >= 0 && chunks >= 1 && pitch > 0, "act_bwd_finalize: bad arguments");
    CAGC_REQUIRE(!gq || d, "act_bwd_finalize: gq needs d");
    act_bwd_finalize_kernel<<<ceil_div(pitch, 32), 256, 0, stream>>>(partial, d, g_bias, gq, nw_part, B, chunks, pitch);
    return launched("act_bwd_finalize_kernel");
}

int cagc_style_grad_finalize(cagc_stream_t stream_, const float* mpartial, const float* gq, const float* s,
                             const float* wsq_oi, float* g_s, int B, int chunks, int pin, int O, int pout) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(mpartial && s && g_s && chunks >= 1 && pin > 0, "style_grad_finalize: bad arguments");
    CAGC_REQUIRE(!gq || (wsq_oi && O > 0 && pout >= O && sizeof(float) * (O + 512) <= 48 * 1024),
                 "style_grad_finalize: bad demod term");
    if (B == 0) return 0;
    CAGC_REQUIRE(B <= 65535, "style_grad_finalize: batch too large");
    style_grad_finalize_kernel<<<dim3(ceil_div(pin, 32), B), 256, sizeof(float) * ((gq ? O : 0) + 512), stream>>>(
        mpartial, gq, s, wsq_oi, g_s, chunks, pin, O, pout);
    return launched("style_grad_finalize_kernel");
}

int cagc_wgrad_finalize(cagc_stream_t stream_, const float* wpart, int nsplits, const float* w, float wscale,
                        const float* gq, const float* s, int B, int O, int I, int ksize, int pin, int pout, float* out) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(wpart && w && out && nsplits >= 1, "wgrad_finalize: bad arguments");
    CAGC_REQUIRE(pin >= I && pout >= O && O > 0 && I > 0, "wgrad_finalize: bad pitches");
    CAGC_REQUIRE(!gq || s, "wgrad_finalize: demodulation term needs s");
    wgrad_finalize_kernel<<<grid_for((int64_t)O * I * ksize * ksize, kNumSMs * 16), 256, 0, stream>>>(wpart, nsplits, w, wscale, gq, s, B, O, I,
                                                                       ksize * ksize, pin, pout, out);
    return launched("wgrad_finalize_kernel");
}

int cagc_torgb_bwd_finalize(cagc_stream_t stream_, const float* partial, const float* s, const float* w, float wscale,
                            float* g_w, float* g_s, int B, int chunks, int cin, int pin, int nout) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(partial && s && w && chunks >= 1 && pin >= cin && nout >= 1 && nout <= 4, "torgb_bwd_finalize: bad arguments");
    if (B == 0) return 0;
    CAGC_REQUIRE(B <= 65535, "torgb_bwd_finalize: batch too large");
    torgb_bwd_finalize_kernel<<<dim3(ceil_div(pin, 32), B), 256, 0, stream>>>(partial, s, w, wscale, g_w, g_s, chunks, cin,
                                                                             pin, nout);
    return launched("torgb_bwd_finalize_kernel");
}

int cagc_linear_bias_act(cagc_stream_t stream_, const float* acc, const float* bias, float* out, int M, int N,
                         float acc_scale, float bias_scale, int act, float alpha, float gain) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(acc && out && M >= 0 && N > 0, "linear_bias_act: bad arguments");
    if (M == 0) return 0;
    linear_bias_act_kernel<<<grid_for((int64_t)M * N), 256, 0, stream>>>(acc, bias, out, (int64_t)M * N, N, acc_scale,
                                                                        bias_scale, act, alpha, gain);
    return launched("linear_bias_act_kernel");
}

int cagc_linear_bias_act_bwd(cagc_stream_t stream_, const float* g, const float* out, float* g_acc, float* g_bias,
                             int M, int N, float acc_scale, float bias_scale, int act, float alpha, float gain) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(g && g_acc && M >= 0 && N > 0, "linear_bias_act_bwd: bad arguments");
    CAGC_REQUIRE(!act || out, "linear_bias_act_bwd: activation backward needs the forward output");
    linear_bias_act_bwd_kernel<<<ceil_div(N, 128), 128, 0, stream>>>(g, out, g_acc, g_bias, M, N, acc_scale, bias_scale,
                                                                    act, alpha, gain);
    return launched("linear_bias_act_bwd_kernel");
}

}  // extern "C"
