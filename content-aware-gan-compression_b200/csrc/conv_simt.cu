// fp32 SIMT implicit-GEMM kernels for the modulated convolution (model.py:241-289) on NHWC-p tensors:
// forward / data-gradient (one generic gather-conv kernel), transposed stride-2 conv as four phase
// convolutions, and the weight gradient with a deterministic split-K reduction.
//
// These are the exact-fp32 engines: the content-aware saliency pass needs fp32-accurate weight
// gradients with a fixed reduction order (SURVEY.md §0 finding 7), and every tensor-pipe (tcgen05)
// kernel is validated against them.  GEMM view: M = pixels (linearised over b, y, x), N = output
// channels, K = taps x input channels; style modulation is applied to the A operand while it is
// staged (no per-sample weights are ever materialised, unlike model.py:249-257).
#include <algorithm>

#include "conv_params.cuh"

namespace cagc {

constexpr int TM = 128, TN = 64, KC = 16;

__global__ void __launch_bounds__(256, 2) conv_simt_kernel(const ConvP p) {
    __shared__ __align__(16) float As[2][KC][TM + 4];
    __shared__ __align__(16) float Bs[2][KC][TN];

    const int tid = threadIdx.x;
    const int64_t M = (int64_t)p.B * p.Ho * p.Wo;
    const int64_t m0 = (int64_t)blockIdx.x * TM;
    const int n0 = blockIdx.y * TN;

    // ---- A loader: one pixel per thread, two channel quads
    const int lm = tid & 127;
    const int lkq = tid >> 7;  // 0/1 -> quads lkq, lkq+2
    const int64_t gm = m0 + lm;
    const bool mvalid = gm < M;
    int lb = 0, loy = 0, lox = 0;
    if (mvalid) {
        lb = (int)(gm / ((int64_t)p.Ho * p.Wo));
        int rem = (int)(gm - (int64_t)lb * p.Ho * p.Wo);
        loy = rem / p.Wo;
        lox = rem - loy * p.Wo;
    }
    const float* scale_row = p.in_scale ? p.in_scale + (int64_t)lb * p.in_pitch : nullptr;
    // ---- B loader
    const int brow = tid >> 4, bcol = (tid & 15) * 4;

    const int nk = ceil_div(p.in_pitch, KC);
    const int total = p.ntaps * nk;

    float4 ra[2], rb;
    auto load_regs = [&](int it) {
        const int tap = it / nk, kc = it - tap * nk;
        const Tap tp = p.taps[tap];
        const int iy = loy * p.in_stride + tp.dy, ix = lox * p.in_stride + tp.dx;
        const bool ok = mvalid && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
        const float* src = p.in + (((int64_t)lb * p.Hin + iy) * p.Win + ix) * p.in_pitch;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int k = kc * KC + (lkq + 2 * r) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok && k < p.in_pitch) {
                v = ldg4(src + k);
                if (scale_row) {
                    float4 s = ldg4(scale_row + k);
                    v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
                }
            }
            ra[r] = v;
        }
        const int krow = kc * KC + brow, col = n0 + bcol;
        rb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (krow < p.in_pitch && col < p.n_cols)
            rb = ldg4(p.w + ((int64_t)tp.slab * p.in_pitch + krow) * p.n_cols + col);
    };
    auto store_smem = [&](int buf) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int kq = (lkq + 2 * r) * 4;
            As[buf][kq + 0][lm] = ra[r].x;
            As[buf][kq + 1][lm] = ra[r].y;
            As[buf][kq + 2][lm] = ra[r].z;
            As[buf][kq + 3][lm] = ra[r].w;
        }
        st4(&Bs[buf][brow][bcol], rb);
    };

    const int tm = tid >> 4, tn = tid & 15;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    load_regs(0);
    store_smem(0);
    __syncthreads();
    for (int it = 0; it < total; ++it) {
        const int buf = it & 1;
        if (it + 1 < total) load_regs(it + 1);
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const float4 a0 = ld4(&As[buf][k][tm * 8]);
            const float4 a1 = ld4(&As[buf][k][tm * 8 + 4]);
            const float4 b = ld4(&Bs[buf][k][tn * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (it + 1 < total) store_smem(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue
    const int n = n0 + tn * 4;
    if (n >= p.n_cols) return;
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) bias4 = ldg4(p.bias + n);
    const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t g = m0 + tm * 8 + i;
        if (g >= M) break;
        const int b = (int)(g / ((int64_t)p.Ho * p.Wo));
        const int rem = (int)(g - (int64_t)b * p.Ho * p.Wo);
        const int oy = rem / p.Wo, ox = rem - oy * p.Wo;
        const int yy = oy * p.out_stride + p.out_oy, xx = ox * p.out_stride + p.out_ox;
        float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
        if (p.out_scale) {
            const float4 s = ldg4(p.out_scale + (int64_t)b * p.n_cols + n);
            v[0] *= s.x; v[1] *= s.y; v[2] *= s.z; v[3] *= s.w;
        }
        if (p.noise) {
            const float nz = nw * __ldg(p.noise + (int64_t)b * p.noise_bstride + (int64_t)yy * p.Wout + xx);
            v[0] += nz; v[1] += nz; v[2] += nz; v[3] += nz;
        }
        v[0] += bias4.x; v[1] += bias4.y; v[2] += bias4.z; v[3] += bias4.w;
        const int64_t off = (((int64_t)b * p.Hout + yy) * p.Wout + xx) * p.n_cols + n;
        if (p.act == kActLrelu) {
            const float gain = (p.act_gain != 0.f) ? p.act_gain : kSqrt2;
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = lrelu_gain(v[j], gain);
        }
        if (p.residual) {
            const float4 r4 = ldg4(p.residual + off);
            if (p.act == kActMaskRef) {
                v[0] = r4.x > 0.f ? v[0] : 0.f; v[1] = r4.y > 0.f ? v[1] : 0.f;
                v[2] = r4.z > 0.f ? v[2] : 0.f; v[3] = r4.w > 0.f ? v[3] : 0.f;
            } else {
                v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (n + j >= p.out_valid) v[j] = 0.f;
        st4(p.out + off, make_float4(v[0], v[1], v[2], v[3]));
    }
}

static int launch_conv(cudaStream_t stream, const ConvP& p, const char* what) {
    const int64_t M = (int64_t)p.B * p.Ho * p.Wo;
    if (M == 0) return 0;
    const int64_t gx = ceil_div<int64_t>(M, TM);
    CAGC_REQUIRE(gx <= 0x7fffffffLL, "%s: too many pixel tiles", what);
    dim3 grid((unsigned)gx, ceil_div(p.n_cols, TN));
    conv_simt_kernel<<<grid, 256, 0, stream>>>(p);
    return launched(what);
}

// ------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------
struct WgradP {
    const float* a;
    const float* a_scale;
    const float* g;
    float* partial;
    int B, H, W;  // base pixel domain
    int a_pitch, g_pitch;
    int Ha, Wa, Hg, Wg, sa, sg;
    int ntaps, nsplits;
    int64_t per_split;  // base pixels per split (multiple of WK)
    int dya[kMaxTaps], dxa[kMaxTaps], dyg[kMaxTaps], dxg[kMaxTaps];
};

constexpr int WT = 64, WK = 16;

__global__ void __launch_bounds__(256, 2) wgrad_simt_kernel(const WgradP p) {
    __shared__ __align__(16) float As[2][WK][WT];
    __shared__ __align__(16) float Bs[2][WK][WT];
    const int tid = threadIdx.x;
    const int i0 = blockIdx.x * WT, o0 = blockIdx.y * WT;
    const int tap = blockIdx.z % p.ntaps, split = blockIdx.z / p.ntaps;
    const int64_t Mtot = (int64_t)p.B * p.H * p.W;
    const int64_t k_lo = (int64_t)split * p.per_split;
    const int64_t k_hi = min(Mtot, k_lo + p.per_split);
    const int lp = tid >> 4, lc = (tid & 15) * 4;
    const int dya = p.dya[tap], dxa = p.dxa[tap], dyg = p.dyg[tap], dxg = p.dxg[tap];

    float4 ra, rb;
    auto load_regs = [&](int64_t k0) {
        const int64_t pix = k0 + lp;
        ra = make_float4(0.f, 0.f, 0.f, 0.f);
        rb = ra;
        if (pix < k_hi) {
            const int b = (int)(pix / ((int64_t)p.H * p.W));
            const int rem = (int)(pix - (int64_t)b * p.H * p.W);
            const int y = rem / p.W, x = rem - y * p.W;
            const int ya = y * p.sa + dya, xa = x * p.sa + dxa;
            const int yg = y * p.sg + dyg, xg = x * p.sg + dxg;
            const bool oka = ya >= 0 && ya < p.Ha && xa >= 0 && xa < p.Wa;
            const bool okg = yg >= 0 && yg < p.Hg && xg >= 0 && xg < p.Wg;
            if (oka && okg) {
                if (i0 + lc < p.a_pitch) {
                    ra = ldg4(p.a + (((int64_t)b * p.Ha + ya) * p.Wa + xa) * p.a_pitch + i0 + lc);
                    if (p.a_scale) {
                        const float4 s = ldg4(p.a_scale + (int64_t)b * p.a_pitch + i0 + lc);
                        ra.x *= s.x; ra.y *= s.y; ra.z *= s.z; ra.w *= s.w;
                    }
                }
                if (o0 + lc < p.g_pitch) rb = ldg4(p.g + (((int64_t)b * p.Hg + yg) * p.Wg + xg) * p.g_pitch + o0 + lc);
            }
        }
    };
    auto store_smem = [&](int buf) {
        st4(&As[buf][lp][lc], ra);
        st4(&Bs[buf][lp][lc], rb);
    };

    const int ti = tid >> 4, to = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    if (k_lo < k_hi) {
        load_regs(k_lo);
        store_smem(0);
        __syncthreads();
        int buf = 0;
        for (int64_t k0 = k_lo; k0 < k_hi; k0 += WK) {
            const bool more = k0 + WK < k_hi;
            if (more) load_regs(k0 + WK);
#pragma unroll
            for (int k = 0; k < WK; ++k) {
                const float4 a = ld4(&As[buf][k][ti * 4]);
                const float4 b = ld4(&Bs[buf][k][to * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w};
                const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            if (more) store_smem(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
    const int o = o0 + to * 4;
    if (o >= p.g_pitch) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ii = i0 + ti * 4 + i;
        if (ii >= p.a_pitch) break;
        float* dst = p.partial + (((int64_t)split * p.ntaps + tap) * p.a_pitch + ii) * p.g_pitch + o;
        st4(dst, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    }
}

__global__ void __launch_bounds__(256) split_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                           int64_t n4, int nsplits) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 s = ld4(partial + 4 * i);
        for (int k = 1; k < nsplits; ++k) {
            const float4 v = ld4(partial + 4 * ((int64_t)k * n4 + i));
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        st4(out + 4 * i, s);
    }
}

}  // namespace cagc

using namespace cagc;

// the SIMT engine for callers in other translation units (disc.cu)
int cagc_simt_conv(cudaStream_t stream, const cagc::ConvP& p, const char* what) { return launch_conv(stream, p, what); }

static int check_nhwc(const char* what, int B, int H, int W, int in_pitch, int out_pitch, int ksize) {
    CAGC_REQUIRE(B >= 0 && H >= 0 && W >= 0, "%s: negative size", what);
    CAGC_REQUIRE(in_pitch > 0 && in_pitch % 4 == 0 && out_pitch > 0 && out_pitch % 4 == 0,
                 "%s: channel pitches must be positive multiples of 4 (got %d, %d)", what, in_pitch, out_pitch);
    CAGC_REQUIRE(ksize >= 1 && ksize * ksize <= kMaxTaps, "%s: unsupported kernel size %d", what, ksize);
    return 0;
}

extern "C" {

int64_t cagc_conv_workspace_bytes(int B, int Ho, int Wo, int out_pitch) {
    // split-K scratch of the tensor-pipe engine: up to 16 partial slabs of the output (at most 64 MB); only layers too small to fill the
    // machine ever use it (<= 256 pixel tiles), so anything large answers 0 and is never split
    const int64_t pixels = (int64_t)B * Ho * Wo;
    if (pixels <= 0 || pixels > 256 * 128) return 0;
    const int64_t slab = pixels * out_pitch * (int64_t)sizeof(float);
    return std::min<int64_t>(16 * slab, std::max<int64_t>(2 * slab, 64ll << 20));
}

int cagc_conv_same(cagc_stream_t stream, const float* in, const float* w_slabs, const float* in_scale,
                   const float* out_scale, const float* noise, const float* noise_w, const float* bias, float* out,
                   int B, int H, int W, int in_pitch, int out_pitch, int out_valid, int ksize, int64_t noise_bstride,
                   int act, int algo) {
    return cagc_conv_same_ws(stream, in, w_slabs, in_scale, out_scale, noise, noise_w, bias, out, B, H, W, in_pitch,
                             out_pitch, out_valid, ksize, noise_bstride, act, algo, nullptr, 0);
}

int cagc_conv_same_ws(cagc_stream_t stream_, const float* in, const float* w_slabs, const float* in_scale,
                      const float* out_scale, const float* noise, const float* noise_w, const float* bias, float* out,
                      int B, int H, int W, int in_pitch, int out_pitch, int out_valid, int ksize, int64_t noise_bstride,
                      int act, int algo, float* workspace, int64_t workspace_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_TRY(check_nhwc("conv_same", B, H, W, in_pitch, out_pitch, ksize));
    CAGC_REQUIRE(in && w_slabs && out, "conv_same: null pointer");
    CAGC_REQUIRE(ksize % 2 == 1, "conv_same: kernel size must be odd");
    CAGC_REQUIRE(!noise || noise_w, "conv_same: noise without noise weight");
    CAGC_REQUIRE(aligned16(in) && aligned16(w_slabs) && aligned16(out), "conv_same: pointers must be 16-byte aligned");
    ConvP p{};
    p.in = in; p.w = w_slabs; p.in_scale = in_scale; p.out_scale = out_scale; p.noise = noise; p.noise_w = noise_w;
    p.bias = bias; p.out = out; p.workspace = workspace; p.workspace_bytes = workspace_bytes;
    p.B = B; p.Hin = H; p.Win = W; p.in_pitch = in_pitch; p.Ho = H; p.Wo = W; p.in_stride = 1;
    p.n_cols = out_pitch; p.out_valid = out_valid; p.Hout = H; p.Wout = W; p.out_stride = 1; p.out_oy = 0; p.out_ox = 0;
    p.noise_bstride = noise_bstride; p.act = act ? kActLrelu : kActNone; p.ntaps = ksize * ksize;
    for (int ky = 0; ky < ksize; ++ky)
        for (int kx = 0; kx < ksize; ++kx) p.taps[ky * ksize + kx] = Tap{ky - ksize / 2, kx - ksize / 2, ky * ksize + kx};
    if (algo == 1) return cagc_tc_conv(stream, p, "conv_same[tc]");
    return launch_conv(stream, p, "conv_same[simt]");
}

int cagc_conv_up(cagc_stream_t stream, const float* in, const float* w_slabs, const float* in_scale, float* out_t,
                 int B, int H, int W, int in_pitch, int out_pitch, int ksize, int algo) {
    return cagc_conv_up_ws(stream, in, w_slabs, in_scale, out_t, B, H, W, in_pitch, out_pitch, ksize, algo, nullptr, 0);
}

int cagc_conv_up_ws(cagc_stream_t stream_, const float* in, const float* w_slabs, const float* in_scale, float* out_t,
                    int B, int H, int W, int in_pitch, int out_pitch, int ksize, int algo, float* workspace,
                    int64_t workspace_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_TRY(check_nhwc("conv_up", B, H, W, in_pitch, out_pitch, ksize));
    CAGC_REQUIRE(in && w_slabs && out_t, "conv_up: null pointer");
    CAGC_REQUIRE(ksize >= 2, "conv_up: kernel size must be >= 2");
    const int Hu = 2 * H + ksize - 2, Wu = 2 * W + ksize - 2;
    ConvP phases[4];
    int nph = 0;
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            ConvP& p = phases[nph++];
            p = ConvP{};
            p.in = in; p.w = w_slabs; p.in_scale = in_scale; p.out = out_t;
            p.B = B; p.Hin = H; p.Win = W; p.in_pitch = in_pitch; p.in_stride = 1;
            p.Ho = (Hu - py + 1) / 2; p.Wo = (Wu - px + 1) / 2;
            p.n_cols = out_pitch; p.out_valid = out_pitch; p.Hout = Hu; p.Wout = Wu; p.out_stride = 2;
            p.out_oy = py; p.out_ox = px; p.act = 0; p.ntaps = 0;
            for (int ky = py; ky < ksize; ky += 2)
                for (int kx = px; kx < ksize; kx += 2)
                    p.taps[p.ntaps++] = Tap{-(ky - py) / 2, -(kx - px) / 2, ky * ksize + kx};
        }
    // wide layers on the tensor pipe: all four phases as one persistent launch; narrow ones (halo-tile kernel) and
    // the SIMT engine: phase by phase
    if (algo == 1) {
        int rc = 0;
        // low-resolution layers with a workspace: all phases and a split of their K loops in one launch
        if (workspace && cagc_tc_conv_multi_splitk(stream, phases, nph, workspace, workspace_bytes, "conv_up[tc,splitk]", &rc))
            return rc;
        if (out_pitch > 80 && cagc_tc_conv_multi(stream, phases, nph, "conv_up[tc,multi]", &rc)) return rc;
    }
    for (int i = 0; i < nph; ++i) {
        if (phases[i].ntaps == 0) continue;
        if (algo == 1) {
            CAGC_TRY(cagc_tc_conv(stream, phases[i], "conv_up[tc]"));
        } else {
            CAGC_TRY(launch_conv(stream, phases[i], "conv_up[simt]"));
        }
    }
    return 0;
}

int cagc_conv_up_dgrad(cagc_stream_t stream, const float* g_t, const float* w_slabs, float* g_in, int B, int H, int W,
                       int g_pitch, int in_pitch, int ksize, int algo) {
    return cagc_conv_up_dgrad_ws(stream, g_t, w_slabs, g_in, B, H, W, g_pitch, in_pitch, ksize, algo, nullptr, 0);
}

int cagc_conv_up_dgrad_ws(cagc_stream_t stream_, const float* g_t, const float* w_slabs, float* g_in, int B, int H, int W,
                          int g_pitch, int in_pitch, int ksize, int algo, float* workspace, int64_t workspace_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_TRY(check_nhwc("conv_up_dgrad", B, H, W, g_pitch, in_pitch, ksize));
    CAGC_REQUIRE(g_t && w_slabs && g_in, "conv_up_dgrad: null pointer");
    ConvP p{};
    p.in = g_t; p.w = w_slabs; p.out = g_in; p.workspace = workspace; p.workspace_bytes = workspace_bytes;
    p.B = B; p.Hin = 2 * H + ksize - 2; p.Win = 2 * W + ksize - 2; p.in_pitch = g_pitch; p.in_stride = 2;
    p.Ho = H; p.Wo = W; p.n_cols = in_pitch; p.out_valid = in_pitch; p.Hout = H; p.Wout = W; p.out_stride = 1;
    p.act = 0; p.ntaps = ksize * ksize;
    for (int ky = 0; ky < ksize; ++ky)
        for (int kx = 0; kx < ksize; ++kx) p.taps[ky * ksize + kx] = Tap{ky, kx, ky * ksize + kx};
    if (algo == 1) return cagc_tc_conv(stream, p, "conv_up_dgrad[tc]");
    return launch_conv(stream, p, "conv_up_dgrad[simt]");
}

int cagc_conv_wgrad_splits(int B, int H, int W, int a_pitch, int g_pitch, int ksize, int algo) {
    if (algo == 1) return cagc_tc_wgrad_splits(B, H, W, a_pitch, g_pitch, ksize);
    const int64_t M = (int64_t)B * H * W;
    const int64_t tiles = (int64_t)ceil_div(a_pitch, WT) * ceil_div(g_pitch, WT) * ksize * ksize;
    int64_t want = ceil_div<int64_t>(4 * kNumSMs, tiles);       // aim for ~4 CTAs per SM
    const int64_t max_by_k = ceil_div<int64_t>(M, 8 * WK);      // at least 128 pixels per split
    if (want > max_by_k) want = max_by_k;
    if (want > 256) want = 256;
    if (want < 1) want = 1;
    return (int)want;
}

// gw == nullptr: leave the split partials unreduced (the caller folds the reduction into
// cagc_wgrad_finalize); *used receives the number of partial slab sets actually written
static int conv_wgrad_impl(cudaStream_t stream, const float* a, const float* a_scale, const float* g, float* gw,
                           float* partial, int nsplits, int B, int H, int W, int a_pitch, int g_pitch, int ksize,
                           int mode, int algo, int* used) {
    CAGC_TRY(check_nhwc("conv_wgrad", B, H, W, a_pitch, g_pitch, ksize));
    CAGC_REQUIRE(a && g && partial, "conv_wgrad: null pointer");
    CAGC_REQUIRE(nsplits >= 1 && nsplits <= 4096, "conv_wgrad: bad nsplits %d", nsplits);
    CAGC_REQUIRE(mode == 0 || mode == 1, "conv_wgrad: mode must be 0 (same) or 1 (up)");
    if (algo == 1) {
        CAGC_REQUIRE(a_scale == nullptr, "conv_wgrad: the tcgen05 path takes a pre-modulated input (cagc_modulate)");
        const int64_t n1 = (int64_t)ksize * ksize * a_pitch * g_pitch;
        if ((int64_t)B * H * W == 0) {
            *used = 1;
            cudaError_t e = cudaMemsetAsync(gw ? gw : partial, 0, n1 * sizeof(float), stream);
            return e == cudaSuccess ? 0 : fail((int)e, "conv_wgrad: memset failed");
        }
        const int tiles_cap = cagc_tc_wgrad_splits(B, H, W, a_pitch, g_pitch, ksize);
        if (nsplits > tiles_cap) nsplits = tiles_cap;
        CAGC_TRY(cagc_tc_wgrad(stream, a, g, partial, &nsplits, B, H, W, a_pitch, g_pitch, ksize, mode));
        *used = nsplits;
        if (!gw) return 0;
        int64_t blocks1 = ceil_div<int64_t>(n1 / 4, 256);
        if (blocks1 > kNumSMs * 8) blocks1 = kNumSMs * 8;
        split_reduce_kernel<<<(unsigned)blocks1, 256, 0, stream>>>(partial, gw, n1 / 4, nsplits);
        return launched("split_reduce_kernel");
    }
    WgradP p{};
    p.a = a; p.a_scale = a_scale; p.g = g; p.partial = partial;
    p.B = B; p.H = H; p.W = W; p.a_pitch = a_pitch; p.g_pitch = g_pitch;
    p.ntaps = ksize * ksize; p.nsplits = nsplits;
    const int64_t M = (int64_t)B * H * W;
    p.per_split = ceil_div<int64_t>(ceil_div<int64_t>(M, nsplits), WK) * WK;
    if (p.per_split < WK) p.per_split = WK;
    if (mode == 0) {
        p.Ha = H; p.Wa = W; p.Hg = H; p.Wg = W; p.sa = 1; p.sg = 1;
        for (int ky = 0; ky < ksize; ++ky)
            for (int kx = 0; kx < ksize; ++kx) {
                const int t = ky * ksize + kx;
                p.dya[t] = ky - ksize / 2; p.dxa[t] = kx - ksize / 2; p.dyg[t] = 0; p.dxg[t] = 0;
            }
    } else {
        p.Ha = H; p.Wa = W; p.Hg = 2 * H + ksize - 2; p.Wg = 2 * W + ksize - 2; p.sa = 1; p.sg = 2;
        for (int ky = 0; ky < ksize; ++ky)
            for (int kx = 0; kx < ksize; ++kx) {
                const int t = ky * ksize + kx;
                p.dya[t] = 0; p.dxa[t] = 0; p.dyg[t] = ky; p.dxg[t] = kx;
            }
    }
    const int64_t n = (int64_t)p.ntaps * a_pitch * g_pitch;
    if (M == 0) {
        *used = 1;
        cudaError_t e = cudaMemsetAsync(gw ? gw : partial, 0, n * sizeof(float), stream);
        if (e != cudaSuccess) return fail((int)e, "conv_wgrad: memset failed");
        return 0;
    }
    dim3 grid(ceil_div(a_pitch, WT), ceil_div(g_pitch, WT), p.ntaps * nsplits);
    wgrad_simt_kernel<<<grid, 256, 0, stream>>>(p);
    CAGC_TRY(launched("wgrad_simt_kernel"));
    *used = nsplits;
    if (!gw) return 0;
    int64_t blocks = ceil_div<int64_t>(n / 4, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    split_reduce_kernel<<<(unsigned)blocks, 256, 0, stream>>>(partial, gw, n / 4, nsplits);
    return launched("split_reduce_kernel");
}

int cagc_conv_wgrad(cagc_stream_t stream_, const float* a, const float* a_scale, const float* g, float* gw,
                    float* partial, int nsplits, int B, int H, int W, int a_pitch, int g_pitch, int ksize, int mode,
                    int algo) {
    CAGC_REQUIRE(gw, "conv_wgrad: null pointer");
    int used = 0;
    return conv_wgrad_impl((cudaStream_t)stream_, a, a_scale, g, gw, partial, nsplits, B, H, W, a_pitch, g_pitch, ksize,
                           mode, algo, &used);
}

int cagc_conv_wgrad_partial(cagc_stream_t stream_, const float* a, const float* a_scale, const float* g,
                            float* partial, int nsplits, int B, int H, int W, int a_pitch, int g_pitch, int ksize,
                            int mode, int algo, int* nsplits_used) {
    CAGC_REQUIRE(nsplits_used, "conv_wgrad_partial: null pointer");
    return conv_wgrad_impl((cudaStream_t)stream_, a, a_scale, g, nullptr, partial, nsplits, B, H, W, a_pitch, g_pitch,
                           ksize, mode, algo, nsplits_used);
}

}  // extern "C"
