// Shared helpers for the cagc_b200 native library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "cagc_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "cagc_b200 is written for sm_100a (B200) only"
#endif

namespace cagc {

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// call right after a kernel launch
inline int launched(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, "%s: launch failed: %s", what, cudaGetErrorString(e));
    return 0;
}

#define CAGC_REQUIRE(cond, ...)                                   \
    do {                                                          \
        if (!(cond)) return ::cagc::fail(CAGC_E_INVALID, __VA_ARGS__); \
    } while (0)

#define CAGC_TRY(expr)            \
    do {                          \
        int _rc = (expr);         \
        if (_rc != 0) return _rc; \
    } while (0)

// One-time-per-DEVICE flags.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device (per-context)
// property: a single process that drives several GPUs (nn.DataParallel, the reference's only live multi-GPU
// path: train.py:522-525, get_fid.py:27) must opt in on each of them.  One bit per device ordinal; threads racing on the
// same device both set the (idempotent) attribute.
using DeviceOnce = std::atomic<unsigned long long>;
inline bool device_once_needed(const DeviceOnce& m) {
    int d = 0;
    cudaGetDevice(&d);
    return !((m.load(std::memory_order_acquire) >> (d & 63)) & 1ull);
}
inline void device_once_done(DeviceOnce& m) {
    int d = 0;
    cudaGetDevice(&d);
    m.fetch_or(1ull << (d & 63), std::memory_order_release);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

constexpr float kSqrt2 = 1.41421356237309504880f;
constexpr float kLreluSlope = 0.2f;
constexpr int kNumSMs = 148;  // B200
// ConvP::act of the convolution epilogues: none / leaky ReLU (or ReLU, act_gain < 0) / "the residual pointer is a ReLU
// mask reference": out = acc * (ref > 0), the backward of a ReLU fused into the data-gradient convolution that feeds it
constexpr int kActNone = 0, kActLrelu = 1, kActMaskRef = 3;

__device__ __forceinline__ float lrelu_sqrt2(float v) { return (v > 0.f ? v : v * kLreluSlope) * kSqrt2; }
// leaky ReLU with an explicit output gain (sqrt2 for StyleGAN2's scaled activation; 1 where a following 1/sqrt2 is folded in)
// gain < 0 selects a plain ReLU scaled by |gain| (the VGG layers of the LPIPS loss); the branch is uniform per launch
__device__ __forceinline__ float lrelu_gain(float v, float gain) {
    if (gain < 0.f) return v > 0.f ? v * -gain : 0.f;
    return (v > 0.f ? v : v * kLreluSlope) * gain;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

}  // namespace cagc
