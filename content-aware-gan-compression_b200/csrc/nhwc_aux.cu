// Bandwidth kernels around the modulated convolution on NHWC-p tensors:
//   * fir_nhwc      : Blur after the transposed conv (model.py:270) + demod/noise/bias/lrelu epilogue
//   * act_bwd       : backward of the StyledConv epilogue with the bias / noise-weight / demod
//                     reductions folded into the same pass (SURVEY.md App. B "Backward")
//   * mod_bwd       : gs reduction + gx = s*gx~ (style-modulation backward)
//   * torgb fwd/bwd : 1x1 modulated conv to RGB + bias + upsampled skip (model.py:380-395)
// All of them are HBM-bound: one 128-bit access per thread per tensor element, threads contiguous
// along (pixel, channel), per-thread channel ownership so that the per-channel reductions live in
// registers, fixed-order (deterministic) partial sums.
#include <cstdlib>

#include "conv_params.cuh"

namespace cagc {

// ------------------------------------------------------------------------------------------------
// FIR on NHWC-p (up = down = 1) with fused epilogue
// thread = one float4 of the flattened output row (x, c4); a strip of TR output rows is produced
// from a ring of KH input rows kept in registers.
// ------------------------------------------------------------------------------------------------
template <int KH, int KW, int TR>
__global__ void __launch_bounds__(256) fir_nhwc_kernel(const float* __restrict__ in, const float* __restrict__ fir,
                                                       const float* __restrict__ out_scale,
                                                       const float* __restrict__ noise,
                                                       const float* __restrict__ noise_w,
                                                       const float* __restrict__ bias, float* __restrict__ out,
                                                       int in_h, int in_w, int out_h, int out_w, int pitch, int valid,
                                                       int pad_x0, int pad_y0, int64_t noise_bstride, int act) {
    __shared__ float skf[KH * KW];
    if (threadIdx.x < KH * KW) {
        int ky = threadIdx.x / KW, kx = threadIdx.x % KW;
        skf[threadIdx.x] = fir[(KH - 1 - ky) * KW + (KW - 1 - kx)];
    }
    __syncthreads();
    const int c4n = pitch >> 2;
    const int pos = blockIdx.x * 256 + threadIdx.x;
    if (pos >= out_w * c4n) return;
    const int ox = pos / c4n, c = (pos - ox * c4n) * 4;
    const int b = blockIdx.z;
    const int oy0 = blockIdx.y * TR;

    float kf[KH * KW];
#pragma unroll
    for (int i = 0; i < KH * KW; ++i) kf[i] = skf[i];

    float4 scale4 = make_float4(1.f, 1.f, 1.f, 1.f), bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (out_scale) scale4 = ldg4(out_scale + (int64_t)b * pitch + c);
    if (bias) bias4 = ldg4(bias + c);
    const float nw = noise ? __ldg(noise_w) : 0.f;

    const float* src = in + (int64_t)b * in_h * in_w * pitch + c;
    const int ix0 = ox - pad_x0;
    float4 ring[KH][KW];
#pragma unroll
    for (int r = 0; r < TR + KH - 1; ++r) {
        const int iy = oy0 + r - pad_y0;
        const bool rowok = iy >= 0 && iy < in_h && (oy0 + r - (KH - 1) < out_h);
#pragma unroll
        for (int j = 0; j < KW; ++j) {
            const int ix = ix0 + j;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rowok && ix >= 0 && ix < in_w) v = ldg4(src + ((int64_t)iy * in_w + ix) * pitch);
            ring[r % KH][j] = v;
        }
        if (r >= KH - 1) {
            const int q = r - (KH - 1);
            const int oy = oy0 + q;
            if (oy < out_h) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < KH; ++i)
#pragma unroll
                    for (int j = 0; j < KW; ++j) {
                        const float k = kf[i * KW + j];
                        const float4 v = ring[(q + i) % KH][j];
                        acc.x = fmaf(k, v.x, acc.x);
                        acc.y = fmaf(k, v.y, acc.y);
                        acc.z = fmaf(k, v.z, acc.z);
                        acc.w = fmaf(k, v.w, acc.w);
                    }
                float v[4] = {acc.x * scale4.x, acc.y * scale4.y, acc.z * scale4.z, acc.w * scale4.w};
                if (noise) {
                    const float nz = nw * __ldg(noise + (int64_t)b * noise_bstride + (int64_t)oy * out_w + ox);
                    v[0] += nz; v[1] += nz; v[2] += nz; v[3] += nz;
                }
                v[0] += bias4.x; v[1] += bias4.y; v[2] += bias4.z; v[3] += bias4.w;
                if (act) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = lrelu_sqrt2(v[j]);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c + j >= valid) v[j] = 0.f;
                st4(out + (((int64_t)b * out_h + oy) * out_w + ox) * pitch + c, make_float4(v[0], v[1], v[2], v[3]));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// epilogue backward.  block = (c4n, PY) threads: x = channel quad (fixed per thread), y = pixel lane
// ------------------------------------------------------------------------------------------------
constexpr int kChunkPixels = 1024;

__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ ga, int64_t sb, int64_t sc, int64_t sh,
                                                      int64_t sw, int ga_vec, const float* __restrict__ a,
                                                      const float* __restrict__ d, const float* __restrict__ noise,
                                                      const float* __restrict__ noise_w,
                                                      const float* __restrict__ bias, float* __restrict__ gu,
                                                      float* __restrict__ partial, int H, int W, int pitch, int valid,
                                                      int64_t noise_bstride, int act, int chunks) {
    extern __shared__ float4 red[];  // [PY][3][c4n]
    const int c4n = blockDim.x, PY = blockDim.y;
    const int cx = threadIdx.x, py = threadIdx.y;
    const int c = cx * 4;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int HW = H * W;
    const int per = ceil_div(HW, chunks);
    const int p_lo = chunk * per, p_hi = min(HW, p_lo + per);

    float4 d4 = make_float4(1.f, 1.f, 1.f, 1.f), bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (d) d4 = ldg4(d + (int64_t)b * pitch + c);
    if (bias) bias4 = ldg4(bias + c);
    const float nw = noise ? __ldg(noise_w) : 0.f;
    // guard the division for padding channels (d == 0 there)
    const float4 rd4 = make_float4(d4.x != 0.f ? 1.f / d4.x : 0.f, d4.y != 0.f ? 1.f / d4.y : 0.f,
                                   d4.z != 0.f ? 1.f / d4.z : 0.f, d4.w != 0.f ? 1.f / d4.w : 0.f);
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f}, s3[4] = {0.f, 0.f, 0.f, 0.f};
    const float dv[4] = {d4.x, d4.y, d4.z, d4.w}, rdv[4] = {rd4.x, rd4.y, rd4.z, rd4.w};
    const float bv[4] = {bias4.x, bias4.y, bias4.z, bias4.w};
    constexpr float kInvPos = 0.70710678118654752440f;            // 1/sqrt2
    constexpr float kInvNeg = 0.70710678118654752440f / 0.2f;     // 1/(0.2*sqrt2)

    const bool flat = (sh == (int64_t)W * sw);      // rows of ga follow each other: no per-pixel division
    // U pixels per trip: all of their loads are issued before the first use (the kernel is bound by load latency, not
    // by bytes, when a thread has a single pixel in flight); pixels past the chunk end re-read the thread's first one
    constexpr int U = 4;
    for (int p0 = p_lo + py; p0 < p_hi; p0 += U * PY) {
        float4 a4[U];
        float gav[U][4], nraw[U];
        int64_t off[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pu = p0 + u * PY;
            const int p = pu < p_hi ? pu : p0;
            off[u] = ((int64_t)b * HW + p) * pitch + c;
            a4[u] = ldg4(a + off[u]);
            const float* gp;
            if (flat) {
                gp = ga + b * sb + (int64_t)p * sw + (int64_t)c * sc;
            } else {
                const int y = p / W, x = p - y * W;
                gp = ga + b * sb + y * sh + x * sw + (int64_t)c * sc;
            }
            if (ga_vec) {
                const float4 t = ld4(gp);
                gav[u][0] = t.x; gav[u][1] = t.y; gav[u][2] = t.z; gav[u][3] = t.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) gav[u][j] = (c + j < valid) ? gp[j * sc] : 0.f;
            }
            nraw[u] = noise ? __ldg(noise + (int64_t)b * noise_bstride + p) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (p0 + u * PY >= p_hi) break;
            const float nz = nw * nraw[u];
            const float av[4] = {a4[u].x, a4[u].y, a4[u].z, a4[u].w};
            float guv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float gz, yv;
                if (act) {
                    const bool pos = av[j] > 0.f;
                    gz = gav[u][j] * (pos ? kSqrt2 : kSqrt2 * kLreluSlope);
                    yv = av[j] * (pos ? kInvPos : kInvNeg) - nz - bv[j];
                } else {
                    gz = gav[u][j];
                    yv = av[j] - nz - bv[j];
                }
                if (c + j >= valid) gz = 0.f;
                guv[j] = gz * dv[j];
                s1[j] += gz;
                s2[j] = fmaf(gz * yv, rdv[j], s2[j]);
                s3[j] = fmaf(gz, nraw[u], s3[j]);
            }
            st4(gu + off[u], make_float4(guv[0], guv[1], guv[2], guv[3]));
        }
    }
    red[(py * 3 + 0) * c4n + cx] = make_float4(s1[0], s1[1], s1[2], s1[3]);
    red[(py * 3 + 1) * c4n + cx] = make_float4(s2[0], s2[1], s2[2], s2[3]);
    red[(py * 3 + 2) * c4n + cx] = make_float4(s3[0], s3[1], s3[2], s3[3]);
    __syncthreads();
    for (int q = py; q < 3; q += PY) {  // thread (cx, q) reduces quantity q over the pixel lanes in a fixed order
        float4 t = red[(0 * 3 + q) * c4n + cx];
        for (int k = 1; k < PY; ++k) {
            const float4 v = red[(k * 3 + q) * c4n + cx];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        st4(partial + ((((int64_t)b * chunks + chunk) * 3 + q) * pitch) + c, t);
    }
}

__global__ void __launch_bounds__(256) mod_bwd_kernel(float* __restrict__ gxt, const float* __restrict__ x,
                                                      const float* __restrict__ s, float* __restrict__ partial,
                                                      int HW, int pitch, int chunks) {
    extern __shared__ float4 red[];  // [PY][c4n]
    const int c4n = blockDim.x, PY = blockDim.y;
    const int cx = threadIdx.x, py = threadIdx.y, c = cx * 4;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int per = ceil_div(HW, chunks);
    const int p_lo = chunk * per, p_hi = min(HW, p_lo + per);
    const float4 s4 = ldg4(s + (int64_t)b * pitch + c);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int U = 4;        // pixels in flight per thread (in-place update: the loads cannot be hoisted by the compiler)
    for (int p0 = p_lo + py; p0 < p_hi; p0 += U * PY) {
        float4 g[U], xv[U];
        int64_t off[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pu = p0 + u * PY;
            off[u] = ((int64_t)b * HW + (pu < p_hi ? pu : p0)) * pitch + c;
            g[u] = ld4(gxt + off[u]);
            xv[u] = ldg4(x + off[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (p0 + u * PY >= p_hi) break;
            acc.x = fmaf(g[u].x, xv[u].x, acc.x);
            acc.y = fmaf(g[u].y, xv[u].y, acc.y);
            acc.z = fmaf(g[u].z, xv[u].z, acc.z);
            acc.w = fmaf(g[u].w, xv[u].w, acc.w);
            st4(gxt + off[u], make_float4(g[u].x * s4.x, g[u].y * s4.y, g[u].z * s4.z, g[u].w * s4.w));
        }
    }
    red[py * c4n + cx] = acc;
    __syncthreads();
    if (py == 0) {
        float4 t = red[cx];
        for (int k = 1; k < PY; ++k) {
            const float4 v = red[k * c4n + cx];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        st4(partial + ((int64_t)b * chunks + chunk) * pitch + c, t);
    }
}

// ------------------------------------------------------------------------------------------------
// ToRGB forward: L lanes per pixel (a power of two, chosen so that a lane owns at most 8 float4 of its pixel), one
// pixel per lane group and step.  All of a lane's loads are issued before the first FMA (the op is bound by load
// latency, not by bytes or math); the effective weights (c * W * s, per sample) sit in shared memory and are read as
// broadcasts; the lane groups of a warp cover consecutive pixels, so the NCHW stores are coalesced.
// ------------------------------------------------------------------------------------------------
constexpr int kRgbMaxOut = 4;
constexpr int kRgbLoads = 8;    // float4 per lane and pixel

template <int L>
__global__ void __launch_bounds__(256) torgb_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ s, const float* __restrict__ bias,
                                                        const float* __restrict__ skip, const float* __restrict__ fir,
                                                        float* __restrict__ out, int H, int W, int pitch, int cin,
                                                        int nout, float wscale, int fh, int fw, int pad0, int chunk) {
    extern __shared__ __align__(16) float weff[];  // [nout][pitch]
    __shared__ float sfir[64];
    const int b = blockIdx.y;
    const int HW = H * W;
    for (int i = threadIdx.x; i < nout * pitch; i += 256) {
        const int o = i / pitch, c = i - o * pitch;
        weff[i] = (c < cin) ? wscale * __ldg(w + o * cin + c) * __ldg(s + (int64_t)b * pitch + c) : 0.f;
    }
    if (skip)
        for (int i = threadIdx.x; i < fh * fw; i += 256) {
            const int ky = i / fw, kx = i - ky * fw;
            sfir[i] = fir[(fh - 1 - ky) * fw + (fw - 1 - kx)];
        }
    __syncthreads();
    constexpr int G = 256 / L;                     // pixels per CTA step
    const int li = threadIdx.x & (L - 1), grp = threadIdx.x / L;
    const int c4n = pitch >> 2;
    const int p_lo = blockIdx.x * chunk;
    const int p_hi = min(HW, p_lo + chunk);
    const int Hs = H >> 1, Ws = W >> 1;
    constexpr int NO = (kRgbMaxOut + L - 1) / L;   // outputs finished by one lane: o = li + L * k
    const bool fast_skip = skip && fh <= 4 && fw <= 4;

    // everything one pixel needs from global memory
    struct Px {
        float4 r[kRgbLoads];
        float skw[4], skv[NO][4];
    };
    // request pixel p: the lane's float4s of the activation, and the 2 x 2 skip samples per output that meet a non-zero tap
    // (Upsample(skip): zero-insert x2, pad (pad0, .), true convolution with fir, model.py:38-56 -- only every other tap of
    // a 4 x 4 kernel meets a sample of the zero-inserted signal)
    auto request = [&](int p, Px& q) {
        const bool pv = p < p_hi;
        const float* xp = x + ((int64_t)b * HW + (pv ? p : p_lo)) * pitch;
#pragma unroll
        for (int u = 0; u < kRgbLoads; ++u) {
            const int c4 = li + L * u;
            q.r[u] = c4 < c4n ? ldg4(xp + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < NO; ++k) q.skv[k][0] = q.skv[k][1] = q.skv[k][2] = q.skv[k][3] = 0.f;
        q.skw[0] = q.skw[1] = q.skw[2] = q.skw[3] = 0.f;
        if (fast_skip && pv) {
            const int y = p / W, xx = p - y * W;
            const int i0 = (pad0 - y) & 1, j0 = (pad0 - xx) & 1;
            int idx[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int i = i0 + 2 * (t >> 1), j = j0 + 2 * (t & 1);
                const int sy = (y + i - pad0) >> 1, sx = (xx + j - pad0) >> 1;   // arithmetic shift: negative stays negative
                const bool ok = i < fh && j < fw && sy >= 0 && sy < Hs && sx >= 0 && sx < Ws;
                q.skw[t] = ok ? sfir[i * fw + j] : 0.f;
                idx[t] = ok ? sy * Ws + sx : 0;
            }
#pragma unroll
            for (int k = 0; k < NO; ++k) {
                const int o = li + L * k;
                if (o < nout) {
                    const float* sp = skip + ((int64_t)b * nout + o) * Hs * Ws;
#pragma unroll
                    for (int t = 0; t < 4; ++t) q.skv[k][t] = __ldg(sp + idx[t]);
                }
            }
        }
    };
    // dot products, reduction over the pixel's lanes, bias, skip, store
    auto finish = [&](int p, const Px& q) {
        float acc[kRgbMaxOut] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int u = 0; u < kRgbLoads; ++u) {
            const int c4 = li + L * u;
            if (c4 < c4n) {
#pragma unroll
                for (int o = 0; o < kRgbMaxOut; ++o) {
                    if (o < nout) {
                        const float4 wv = ld4(&weff[o * pitch + c4 * 4]);
                        acc[o] = fmaf(q.r[u].x, wv.x, acc[o]);
                        acc[o] = fmaf(q.r[u].y, wv.y, acc[o]);
                        acc[o] = fmaf(q.r[u].z, wv.z, acc[o]);
                        acc[o] = fmaf(q.r[u].w, wv.w, acc[o]);
                    }
                }
            }
        }
#pragma unroll
        for (int o = 0; o < kRgbMaxOut; ++o)
#pragma unroll
            for (int m = L >> 1; m >= 1; m >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], m);
        if (p >= p_hi) return;
#pragma unroll
        for (int k = 0; k < NO; ++k) {
            const int o = li + L * k;
            if (o >= nout) continue;
            float v = acc[0];
#pragma unroll
            for (int qo = 1; qo < kRgbMaxOut; ++qo)
                if (o == qo) v = acc[qo];
            if (bias) v += __ldg(bias + o);
            if (fast_skip) {
                float uacc = 0.f;
#pragma unroll
                for (int t = 0; t < 4; ++t) uacc = fmaf(q.skw[t], q.skv[k][t], uacc);
                v += uacc;
            } else if (skip) {
                const int y = p / W, xx = p - y * W;
                const float* sp = skip + ((int64_t)b * nout + o) * Hs * Ws;
                float uacc = 0.f;
                const int i0 = (pad0 - y) & 1, j0 = (pad0 - xx) & 1;
                for (int i = i0; i < fh; i += 2) {
                    const int sy = (y + i - pad0) >> 1;
                    if (sy < 0) continue;
                    if (sy >= Hs) break;
                    for (int j = j0; j < fw; j += 2) {
                        const int sx = (xx + j - pad0) >> 1;
                        if (sx < 0) continue;
                        if (sx >= Ws) break;
                        uacc = fmaf(sfir[i * fw + j], __ldg(sp + sy * Ws + sx), uacc);
                    }
                }
                v += uacc;
            }
            out[((int64_t)b * nout + o) * HW + p] = v;
        }
    };

    // two pixels in flight per lane group: the loads of step k+1 are requested before step k is finished, into the buffer
    // step k-1 has released (no register moves that would wait for a pending load); trip counts are warp-uniform
    Px qa, qb;
    request(p_lo + grp, qa);
    for (int pbase = p_lo; pbase < p_hi; pbase += 2 * G) {
        if (pbase + G < p_hi) request(pbase + G + grp, qb);
        finish(pbase + grp, qa);
        if (pbase + G >= p_hi) break;
        if (pbase + 2 * G < p_hi) request(pbase + 2 * G + grp, qa);
        finish(pbase + G + grp, qb);
    }
}

// ToRGB backward: block = (c4n, PY); gx written, per-sample T[o][c] partials
__global__ void __launch_bounds__(256) torgb_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                        const float* __restrict__ w, const float* __restrict__ s,
                                                        const float* __restrict__ gx_add, float* __restrict__ gx,
                                                        float* __restrict__ partial, int HW,
                                                        int pitch, int cin, int nout, float wscale, int chunks) {
    extern __shared__ float4 red[];  // [PY][nout][c4n]
    const int c4n = blockDim.x, PY = blockDim.y;
    const int cx = threadIdx.x, py = threadIdx.y, c = cx * 4;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int per = ceil_div(HW, chunks);
    const int p_lo = chunk * per, p_hi = min(HW, p_lo + per);
    float4 we[kRgbMaxOut], T[kRgbMaxOut];
    const float4 s4 = ldg4(s + (int64_t)b * pitch + c);
    const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
    for (int o = 0; o < kRgbMaxOut; ++o) {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        if (o < nout) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c + j < cin) t[j] = wscale * __ldg(w + o * cin + c + j) * sv[j];
        }
        we[o] = make_float4(t[0], t[1], t[2], t[3]);
        T[o] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    constexpr int U = 4;        // pixels in flight per thread
    for (int p0 = p_lo + py; p0 < p_hi; p0 += U * PY) {
        float4 xv[U], ev[U];
        float gv[U][kRgbMaxOut];
        int64_t off[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pu = p0 + u * PY;
            const int p = pu < p_hi ? pu : p0;
            off[u] = ((int64_t)b * HW + p) * pitch + c;
            xv[u] = ld4(x + off[u]);
            ev[u] = gx_add ? ld4(gx_add + off[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int o = 0; o < kRgbMaxOut; ++o) gv[u][o] = o < nout ? __ldg(g + ((int64_t)b * nout + o) * HW + p) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (p0 + u * PY >= p_hi) break;
            float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int o = 0; o < kRgbMaxOut; ++o) {
                if (o < nout) {
                    o4.x = fmaf(we[o].x, gv[u][o], o4.x);
                    o4.y = fmaf(we[o].y, gv[u][o], o4.y);
                    o4.z = fmaf(we[o].z, gv[u][o], o4.z);
                    o4.w = fmaf(we[o].w, gv[u][o], o4.w);
                    T[o].x = fmaf(gv[u][o], xv[u].x, T[o].x);
                    T[o].y = fmaf(gv[u][o], xv[u].y, T[o].y);
                    T[o].z = fmaf(gv[u][o], xv[u].z, T[o].z);
                    T[o].w = fmaf(gv[u][o], xv[u].w, T[o].w);
                }
            }
            // gradient that reached the same activation through its other consumer (the next conv)
            o4.x += ev[u].x; o4.y += ev[u].y; o4.z += ev[u].z; o4.w += ev[u].w;
            st4(gx + off[u], o4);
        }
    }
#pragma unroll
    for (int o = 0; o < kRgbMaxOut; ++o)
        if (o < nout) red[(py * nout + o) * c4n + cx] = T[o];
    __syncthreads();
    for (int o = py; o < nout; o += PY) {
        float4 t = red[o * c4n + cx];
        for (int k = 1; k < PY; ++k) {
            const float4 v = red[(k * nout + o) * c4n + cx];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        st4(partial + (((int64_t)b * chunks + chunk) * nout + o) * pitch + c, t);
    }
}


// fused leaky-ReLU backward on channel-contiguous storage [rows][C] (channels-last activations, 2-D
// [B, C] tensors) with the bias gradient reduced in the same pass: block = (C/4, PY) threads
__global__ void __launch_bounds__(256) bias_act_bwd_rows_kernel(const float* __restrict__ g, const float* __restrict__ refer,
                                                                float* __restrict__ gin, float* __restrict__ partial,
                                                                int64_t rows, int C, int chunks, float alpha, float scale) {
    extern __shared__ float4 red[];  // [PY][c4n]
    const int c4n = blockDim.x, PY = blockDim.y;
    const int cx = threadIdx.x, py = threadIdx.y, c = cx * 4;
    const int chunk = blockIdx.x;
    // each block owns a contiguous run of rows (an interleaved / grid-stride deal of the rows was measured 20% SLOWER)
    const int64_t per = ceil_div<int64_t>(rows, chunks);
    const int64_t r_lo = chunk * per, r_hi = min(rows, r_lo + per);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int64_t r = r_lo + py; r < r_hi; r += PY) {        const int64_t off = r * C + c;
        const float4 gv = ldg4(g + off), rv = ldg4(refer + off);
        float4 o;
        o.x = (rv.x > 0.f ? gv.x : gv.x * alpha) * scale;
        o.y = (rv.y > 0.f ? gv.y : gv.y * alpha) * scale;
        o.z = (rv.z > 0.f ? gv.z : gv.z * alpha) * scale;
        o.w = (rv.w > 0.f ? gv.w : gv.w * alpha) * scale;
        st4(gin + off, o);
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    red[py * c4n + cx] = acc;
    __syncthreads();
    if (py == 0) {
        float4 t = red[cx];
        for (int k = 1; k < PY; ++k) {
            const float4 v = red[k * c4n + cx];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        st4(partial + (int64_t)chunk * C + c, t);
    }
}

static inline int pixel_chunks(int HW) {
    // pixels per CTA: 1024 at >= 256^2, shrinking with the image so that the low-resolution layers still fill the GPU
    // (16 CTAs x 1024-pixel serial loops made the 4^2 .. 32^2 layers 30-50 us each: pure latency)
    const int per = HW >= 65536 ? kChunkPixels : HW >= 16384 ? 512 : HW >= 4096 ? 256 : 64;
    int c = ceil_div(HW, per);
    if (c < 1) c = 1;
    if (c > 128) c = 128;
    return c;
}

// (c4n, PY) block shape shared by the reduction kernels; PY >= need_py
static inline int block_shape(int pitch, int need_py, dim3* blk) {
    const int c4n = pitch / 4;
    if (c4n < 1 || c4n > 256) return -1;
    int PY = 256 / c4n;
    if (PY < need_py) return -1;
    if (PY > 64) PY = 64;
    *blk = dim3(c4n, PY);
    return 0;
}

}  // namespace cagc

using namespace cagc;

extern "C" {

static int fir_nhwc_impl(cagc_stream_t stream_, const float* in, const float* fir, const float* taps_host,
                         const float* out_scale, const float* noise, const float* noise_w, const float* bias, float* out,
                         int B, int in_h, int in_w, int pitch, int valid, int kh, int kw, int pad_x0, int pad_x1,
                         int pad_y0, int pad_y1, int64_t noise_bstride, int act) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(in && fir && out, "fir_nhwc: null pointer");
    CAGC_REQUIRE(pitch > 0 && pitch % 4 == 0, "fir_nhwc: pitch must be a positive multiple of 4");
    CAGC_REQUIRE(!noise || noise_w, "fir_nhwc: noise without noise weight");
    CAGC_REQUIRE(aligned16(in) && aligned16(out), "fir_nhwc: pointers must be 16-byte aligned");
    if (kh != 4 || kw != 4) return fail(CAGC_E_UNSUPPORTED, "fir_nhwc: only 4x4 FIR kernels are implemented (got %dx%d)", kh, kw);
    const int out_h = in_h + pad_y0 + pad_y1 - kh + 1, out_w = in_w + pad_x0 + pad_x1 - kw + 1;
    if (B == 0 || out_h <= 0 || out_w <= 0) return 0;
    {
        int rc = 0;
        if (cagc_tc_fir_nhwc(stream, in, fir, out_scale, noise, noise_w, bias, out, B, in_h, in_w, out_h, out_w, pitch,
                             valid, pad_x0, pad_y0, noise_bstride, act, taps_host, &rc))
            return rc;
    }
    CAGC_REQUIRE(B <= 65535, "fir_nhwc: batch too large");
    constexpr int TR = 8;
    dim3 grid(ceil_div(out_w * (pitch / 4), 256), ceil_div(out_h, TR), B);
    fir_nhwc_kernel<4, 4, TR><<<grid, 256, 0, stream>>>(in, fir, out_scale, noise, noise_w, bias, out, in_h, in_w, out_h,
                                                       out_w, pitch, valid, pad_x0, pad_y0, noise_bstride, act);
    return launched("fir_nhwc_kernel");
}

int cagc_fir_nhwc(cagc_stream_t stream, const float* in, const float* fir, const float* out_scale, const float* noise,
                  const float* noise_w, const float* bias, float* out, int B, int in_h, int in_w, int pitch, int valid,
                  int kh, int kw, int pad_x0, int pad_x1, int pad_y0, int pad_y1, int64_t noise_bstride, int act) {
    return fir_nhwc_impl(stream, in, fir, nullptr, out_scale, noise, noise_w, bias, out, B, in_h, in_w, pitch, valid, kh,
                         kw, pad_x0, pad_x1, pad_y0, pad_y1, noise_bstride, act);
}

int cagc_fir_nhwc_taps(cagc_stream_t stream, const float* in, const float* fir, const float* taps_host,
                       const float* out_scale, const float* noise, const float* noise_w, const float* bias, float* out,
                       int B, int in_h, int in_w, int pitch, int valid, int kh, int kw, int pad_x0, int pad_x1,
                       int pad_y0, int pad_y1, int64_t noise_bstride, int act) {
    CAGC_REQUIRE(taps_host, "fir_nhwc_taps: null host taps");
    return fir_nhwc_impl(stream, in, fir, taps_host, out_scale, noise, noise_w, bias, out, B, in_h, in_w, pitch, valid, kh,
                         kw, pad_x0, pad_x1, pad_y0, pad_y1, noise_bstride, act);
}

// FIR (4x4, up = down = 1) whose output is multiplied by the leaky-ReLU derivative of a reference activation:
// out = FIR(in) * (mask_ref > 0 ? gain : 0.2 * gain) -- Blur^T followed by the backward of FusedLeakyReLU in one pass
// (discriminator ResBlock backward).  CAGC_E_UNSUPPORTED when the TMA row-ring kernel cannot take the shape.
int cagc_fir_nhwc_mask(cagc_stream_t stream_, const float* in, const float* fir, const float* taps_host,
                       const float* mask_ref, float mask_gain, float* out, int B, int in_h, int in_w, int pitch, int valid,
                       int kh, int kw, int pad_x0, int pad_x1, int pad_y0, int pad_y1) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(in && fir && out && mask_ref && taps_host, "fir_nhwc_mask: null pointer");
    CAGC_REQUIRE(pitch > 0 && pitch % 8 == 0, "fir_nhwc_mask: pitch must be a positive multiple of 8");
    CAGC_REQUIRE(aligned16(in) && aligned16(out) && aligned16(mask_ref), "fir_nhwc_mask: pointers must be 16-byte aligned");
    if (kh != 4 || kw != 4) return fail(CAGC_E_UNSUPPORTED, "fir_nhwc_mask: only 4x4 FIR kernels are implemented");
    const int out_h = in_h + pad_y0 + pad_y1 - kh + 1, out_w = in_w + pad_x0 + pad_x1 - kw + 1;
    if (B == 0 || out_h <= 0 || out_w <= 0) return 0;
    int rc = 0;
    if (cagc_tc_fir_nhwc(stream, in, fir, nullptr, nullptr, nullptr, nullptr, out, B, in_h, in_w, out_h, out_w, pitch, valid,
                         pad_x0, pad_y0, 0, 0, taps_host, &rc, mask_ref, mask_gain))
        return rc;
    return fail(CAGC_E_UNSUPPORTED, "fir_nhwc_mask: shape not handled by the row-ring kernel");
}

int cagc_act_bwd_chunks(int H, int W) { return pixel_chunks(H * W); }

int cagc_bias_grad_rows_chunks(int64_t rows, int C) {
    if (C % 4 != 0 || C > 1024 || rows < 64) return 0;   // caller reduces grad_in itself
    // ~16 rows per block (a block covers PY >= 1 rows per step), capped at one resident wave.  (One block per
    // 512 rows left the low-resolution layers of the discriminator to a handful of blocks: 60-130 us each.)
    int64_t c = ceil_div<int64_t>(rows, 16);
    if (c > 8 * kNumSMs) c = 8 * kNumSMs;
    return (int)c;
}

int cagc_fused_bias_act_bwd_rows(cagc_stream_t stream_, const float* grad_out, const float* refer, float* grad_in,
                                 float* bias_partial, int64_t rows, int C, float alpha, float scale) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(grad_out && refer && grad_in && bias_partial, "fused_bias_act_bwd_rows: null pointer");
    const int chunks = cagc_bias_grad_rows_chunks(rows, C);
    CAGC_REQUIRE(chunks > 0, "fused_bias_act_bwd_rows: shape not eligible for the fused reduction");
    CAGC_REQUIRE(aligned16(grad_out) && aligned16(refer) && aligned16(grad_in) && aligned16(bias_partial),
                 "fused_bias_act_bwd_rows: pointers must be 16-byte aligned");
    dim3 blk;
    CAGC_REQUIRE(block_shape(C, 1, &blk) == 0, "fused_bias_act_bwd_rows: C %d unsupported", C);
    bias_act_bwd_rows_kernel<<<chunks, blk, sizeof(float4) * blk.x * blk.y, stream>>>(grad_out, refer, grad_in,
                                                                                  bias_partial, rows, C, chunks, alpha,
                                                                                  scale);
    return launched("bias_act_bwd_rows_kernel");
}

int cagc_act_bwd(cagc_stream_t stream_, const float* ga, int64_t sb, int64_t sc, int64_t sh, int64_t sw, const float* a,
                 const float* d, const float* noise, const float* noise_w, const float* bias, float* gu, float* partial,
                 int B, int H, int W, int pitch, int valid, int64_t noise_bstride, int act) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(ga && a && gu && partial, "act_bwd: null pointer");
    CAGC_REQUIRE(pitch > 0 && pitch % 4 == 0, "act_bwd: pitch must be a positive multiple of 4");
    CAGC_REQUIRE(!noise || noise_w, "act_bwd: noise without noise weight");
    if (B == 0 || H * W == 0) return 0;
    dim3 blk;
    CAGC_REQUIRE(block_shape(pitch, 1, &blk) == 0, "act_bwd: pitch %d unsupported (max 1024)", pitch);
    const int chunks = pixel_chunks(H * W);
    // a float4 read may touch channels >= valid; that is only in bounds when the source pixel pitch
    // covers our pitch (an NHWC-p buffer, or dense NHWC with C a multiple of 4)
    const int vec = (sc == 1) && (sb % 4 == 0) && (sh % 4 == 0) && (sw % 4 == 0) && aligned16(ga) && (sw >= pitch);
    dim3 grid(chunks, B);
    const size_t smem = sizeof(float4) * blk.y * 3 * blk.x;
    act_bwd_kernel<<<grid, blk, smem, stream>>>(ga, sb, sc, sh, sw, vec, a, d, noise, noise_w, bias, gu, partial, H, W,
                                                pitch, valid, noise_bstride, act, chunks);
    return launched("act_bwd_kernel");
}

int cagc_mod_bwd(cagc_stream_t stream_, float* gxt, const float* x, const float* s, float* partial, int B, int H, int W,
                 int pitch) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(gxt && x && s && partial, "mod_bwd: null pointer");
    CAGC_REQUIRE(pitch > 0 && pitch % 4 == 0, "mod_bwd: pitch must be a positive multiple of 4");
    if (B == 0 || H * W == 0) return 0;
    dim3 blk;
    CAGC_REQUIRE(block_shape(pitch, 1, &blk) == 0, "mod_bwd: pitch %d unsupported (max 1024)", pitch);
    const int chunks = pixel_chunks(H * W);
    dim3 grid(chunks, B);
    mod_bwd_kernel<<<grid, blk, sizeof(float4) * blk.x * blk.y, stream>>>(gxt, x, s, partial, H * W, pitch, chunks);
    return launched("mod_bwd_kernel");
}

int cagc_torgb_fwd(cagc_stream_t stream_, const float* x, const float* w, const float* s, const float* bias,
                   const float* skip, const float* fir, float* out, int B, int H, int W, int pitch, int cin, int nout,
                   float wscale, int fh, int fw, int pad0, int pad1) {
    cudaStream_t stream = (cudaStream_t)stream_;
    (void)pad1;
    CAGC_REQUIRE(x && w && s && out, "torgb_fwd: null pointer");
    CAGC_REQUIRE(pitch > 0 && pitch % 4 == 0 && cin <= pitch, "torgb_fwd: bad pitch");
    CAGC_REQUIRE(nout >= 1 && nout <= kRgbMaxOut, "torgb_fwd: nout must be 1..4");
    CAGC_REQUIRE(!skip || (fir && fh * fw <= 64 && H % 2 == 0 && W % 2 == 0), "torgb_fwd: bad skip/fir");
    if (B == 0 || H * W == 0) return 0;
    CAGC_REQUIRE(B <= 65535, "torgb_fwd: batch too large");
    const size_t smem = sizeof(float) * nout * pitch;
    CAGC_REQUIRE(smem <= 48 * 1024, "torgb_fwd: pitch too large");
    // lanes per pixel: at most kRgbLoads float4 per lane
    int L = 1;
    while (L < 32 && (pitch / 4 + L - 1) / L > kRgbLoads) L *= 2;
    // pixels per CTA: a multiple of the 256 / L pixels of one step, up to 16 steps (amortises the per-CTA effective-weight
    // set-up) while the grid keeps >= ~4 CTAs per SM
    const int G = 256 / L;
    int64_t steps = ((int64_t)H * W * B) / ((int64_t)G * 4 * kNumSMs);
    steps = steps < 1 ? 1 : steps > 16 ? 16 : steps;
    const int chunk = G * (int)steps;
    dim3 grid(ceil_div(H * W, chunk), B);
#define CAGC_RGB_LAUNCH(LL)                                                                                              \
    torgb_fwd_kernel<LL><<<grid, 256, smem, stream>>>(x, w, s, bias, skip, fir, out, H, W, pitch, cin, nout, wscale, fh, \
                                                      fw, pad0, chunk)
    switch (L) {
        case 1: CAGC_RGB_LAUNCH(1); break;
        case 2: CAGC_RGB_LAUNCH(2); break;
        case 4: CAGC_RGB_LAUNCH(4); break;
        case 8: CAGC_RGB_LAUNCH(8); break;
        case 16: CAGC_RGB_LAUNCH(16); break;
        default: CAGC_RGB_LAUNCH(32); break;
    }
#undef CAGC_RGB_LAUNCH
    return launched("torgb_fwd_kernel");
}

int cagc_torgb_bwd(cagc_stream_t stream_, const float* g, const float* x, const float* w, const float* s,
                   const float* gx_add, float* gx, float* partial, int B, int H, int W, int pitch, int cin, int nout,
                   float wscale) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(g && x && w && s && gx && partial, "torgb_bwd: null pointer");
    CAGC_REQUIRE(pitch > 0 && pitch % 4 == 0 && cin <= pitch, "torgb_bwd: bad pitch");
    CAGC_REQUIRE(nout >= 1 && nout <= kRgbMaxOut, "torgb_bwd: nout must be 1..4");
    if (B == 0 || H * W == 0) return 0;
    dim3 blk;
    CAGC_REQUIRE(block_shape(pitch, 1, &blk) == 0, "torgb_bwd: pitch %d unsupported (max 1024)", pitch);
    const int chunks = pixel_chunks(H * W);
    dim3 grid(chunks, B);
    const size_t smem = sizeof(float4) * blk.y * nout * blk.x;
    torgb_bwd_kernel<<<grid, blk, smem, stream>>>(g, x, w, s, gx_add, gx, partial, H * W, pitch, cin, nout, wscale, chunks);
    return launched("torgb_bwd_kernel");
}

}  // extern "C"
