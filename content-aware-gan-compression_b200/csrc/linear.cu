// EqualLinear (reference model.py:137-166) as own kernels: y = act(scale * x W^T + bias * lr_mul) [* gain].
//
// The reference calls F.linear (a cuBLAS sgemm) followed by fused_bias_act; with M = batch (16 .. 64 rows) the
// product is a skinny GEMM whose cost is reading W once (1 MB for the 512 x 512 mapping layers, 16.8 MB for the
// discriminator head) plus launch latency, so it is written as bandwidth kernels with the epilogue fused:
//   linear_fwd      one warp per output feature n: the warp streams W[n, :] with 128-bit loads, every lane keeps
//                   ROWS accumulators (x rows come from L1), warp-shuffle reduction, bias + leaky ReLU in place
//   linear_bwd_x    g_x[m, k] = sum_n g_acc[m, n] W[n, k]: thread = (k, group of 4 rows), W read coalesced along k
//   linear_bwd_w    g_W[n, k] = sum_m g_acc[m, n] x[m, k]: thread = (n, k) element, fixed summation order
// fp32 FMA accumulation in a fixed order (deterministic; the saliency path needs that).
#include "common.cuh"

namespace cagc {

constexpr int kLinRows = 16;     // rows per pass of the forward kernel (accumulators per lane)
constexpr int kLinKTile = 1024;  // K elements of x staged in shared memory per pass (kLinRows x 1024 x 4 B = 64 KB)

// Block = 8 warps = 8 output features; the block stages a [rows x K-tile] slab of x in shared memory once and every
// warp streams its own W row against it: no dependent global round trip per k step (the first version re-read x
// through L1/L2 inside the k loop and was latency-bound: 57 us per 512 x 512 layer under ncu, cuBLAS 18 us).
__global__ void __launch_bounds__(256) linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ out, int M,
                                                         int N, int K, float acc_scale, float bias_scale, int act,
                                                         float alpha, float gain, int vec) {
    extern __shared__ __align__(16) float xs[];          // [kLinRows][kt]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + warp;
    const bool nvalid = n < N;
    const float* wr = w + (int64_t)(nvalid ? n : 0) * K;
    for (int m0 = 0; m0 < M; m0 += kLinRows) {
        const int rows = min(kLinRows, M - m0);
        float acc[kLinRows];
#pragma unroll
        for (int r = 0; r < kLinRows; ++r) acc[r] = 0.f;
        for (int k0 = 0; k0 < K; k0 += kLinKTile) {
            const int kt = min(kLinKTile, K - k0);
            __syncthreads();                              // previous tile fully consumed
            if (vec) {
                const int kt4 = kt >> 2;
                for (int i = threadIdx.x; i < rows * kt4; i += 256) {
                    const int r = i / kt4, c = i - r * kt4;
                    st4(xs + r * kLinKTile + c * 4, ldg4(x + (int64_t)(m0 + r) * K + k0 + c * 4));
                }
            } else {
                for (int i = threadIdx.x; i < rows * kt; i += 256) {
                    const int r = i / kt, c = i - r * kt;
                    xs[r * kLinKTile + c] = __ldg(x + (int64_t)(m0 + r) * K + k0 + c);
                }
            }
            __syncthreads();
            if (vec) {
#pragma unroll 2
                for (int k = lane * 4; k < kt; k += 128) {
                    const float4 wv = ldg4(wr + k0 + k);
#pragma unroll
                    for (int r = 0; r < kLinRows; ++r) {
                        if (r < rows) {
                            const float4 xv = ld4(xs + r * kLinKTile + k);
                            acc[r] = fmaf(wv.x, xv.x, acc[r]);
                            acc[r] = fmaf(wv.y, xv.y, acc[r]);
                            acc[r] = fmaf(wv.z, xv.z, acc[r]);
                            acc[r] = fmaf(wv.w, xv.w, acc[r]);
                        }
                    }
                }
            } else {
                for (int k = lane; k < kt; k += 32) {
                    const float wv = __ldg(wr + k0 + k);
#pragma unroll
                    for (int r = 0; r < kLinRows; ++r)
                        if (r < rows) acc[r] = fmaf(wv, xs[r * kLinKTile + k], acc[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kLinRows; ++r) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
        }
        if (lane == 0 && nvalid) {
            const float b = bias ? __ldg(bias + n) * bias_scale : 0.f;
#pragma unroll
            for (int r = 0; r < kLinRows; ++r) {
                if (r < rows) {
                    float v = acc[r] * acc_scale + b;
                    if (act) v = (v > 0.f ? v : v * alpha) * gain;
                    out[(int64_t)(m0 + r) * N + n] = v;
                }
            }
        }
    }
}

// g_x[m, k] = sum_n ga[m, n] * w[n, k]: block = 256 consecutive k x 4 rows; the 4 x N slab of ga sits in shared memory
// (broadcast reads), W is read coalesced along k with 8 independent loads in flight per thread.
__global__ void __launch_bounds__(256) linear_bwd_x_kernel(const float* __restrict__ ga, const float* __restrict__ w,
                                                           float* __restrict__ gx, int M, int N, int K) {
    extern __shared__ __align__(16) float gs[];           // [4][N]
    const int k = blockIdx.x * 256 + threadIdx.x;
    const int m0 = blockIdx.y * 4;
    for (int i = threadIdx.x; i < 4 * N; i += 256) {
        const int r = i / N, nn = i - r * N;
        gs[i] = (m0 + r < M) ? __ldg(ga + (int64_t)(m0 + r) * N + nn) : 0.f;
    }
    __syncthreads();
    if (k >= K) return;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int n = 0;
    for (; n + 8 <= N; n += 8) {
        float wv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) wv[u] = __ldg(w + (int64_t)(n + u) * K + k);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = fmaf(gs[r * N + n + u], wv[u], acc[r]);
        }
    }
    for (; n < N; ++n) {
        const float wv = __ldg(w + (int64_t)n * K + k);
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r] = fmaf(gs[r * N + n], wv, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
        if (m0 + r < M) gx[(int64_t)(m0 + r) * K + k] = acc[r];
}

// thread = element (n, k): g_w[n, k] = sum_m ga[m, n] * x[m, k]
__global__ void __launch_bounds__(256) linear_bwd_w_kernel(const float* __restrict__ ga, const float* __restrict__ x,
                                                           float* __restrict__ gw, int M, int N, int K) {
    const int k = blockIdx.x * 256 + threadIdx.x;
    const int n = blockIdx.y;
    if (k >= K) return;
    float acc = 0.f;
    for (int m = 0; m < M; ++m) acc = fmaf(__ldg(ga + (int64_t)m * N + n), __ldg(x + (int64_t)m * K + k), acc);
    gw[(int64_t)n * K + k] = acc;
}

}  // namespace cagc

using namespace cagc;

extern "C" {

int cagc_linear_fwd(cagc_stream_t stream_, const float* x, const float* w, const float* bias, float* out, int M, int N,
                    int K, float acc_scale, float bias_scale, int act, float alpha, float gain) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(x && w && out, "linear_fwd: null pointer");
    CAGC_REQUIRE(M >= 0 && N > 0 && K > 0, "linear_fwd: bad sizes");
    if (M == 0) return 0;
    const int vec = (K % 4 == 0) && aligned16(w) && aligned16(x);
    static DeviceOnce attr_set{0};
    const size_t smem = (size_t)kLinRows * kLinKTile * sizeof(float);
    if (device_once_needed(attr_set)) {
        cudaError_t e = cudaFuncSetAttribute(linear_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail((int)e, "linear_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        device_once_done(attr_set);
    }
    linear_fwd_kernel<<<ceil_div(N, 8), 256, smem, stream>>>(x, w, bias, out, M, N, K, acc_scale, bias_scale, act, alpha,
                                                             gain, vec);
    return launched("linear_fwd_kernel");
}

int cagc_linear_bwd(cagc_stream_t stream_, const float* g_acc, const float* x, const float* w, float* g_x, float* g_w,
                    int M, int N, int K) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(g_acc && x && w, "linear_bwd: null pointer");
    CAGC_REQUIRE(M > 0 && N > 0 && K > 0 && N <= 65535, "linear_bwd: bad sizes");
    if (g_x) {
        dim3 grid(ceil_div(K, 256), ceil_div(M, 4));
        CAGC_REQUIRE((size_t)4 * N * sizeof(float) <= 48 * 1024, "linear_bwd: more than 3072 output features");
        linear_bwd_x_kernel<<<grid, 256, (size_t)4 * N * sizeof(float), stream>>>(g_acc, w, g_x, M, N, K);
        CAGC_TRY(launched("linear_bwd_x_kernel"));
    }
    if (g_w) {
        dim3 grid(ceil_div(K, 256), N);
        linear_bwd_w_kernel<<<grid, 256, 0, stream>>>(g_acc, x, g_w, M, N, K);
        CAGC_TRY(launched("linear_bwd_w_kernel"));
    }
    return 0;
}

}  // extern "C"
