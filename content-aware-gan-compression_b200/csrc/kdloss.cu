// Content-mask glue of the KD loss on the device (reference Util/content_aware_pruning.py:61-117, called from
// train.py:154-158): the reference rescales / resizes / normalises the teacher image with seven ATen launches, takes an
// int64 argmax of the parser's 19 class maps, builds the mask on the HOST (`.type(torch.FloatTensor)`: a device->host
// copy and a sync in the middle of the step) and multiplies image by image in a Python loop.  Two bandwidth kernels:
//
//   parse_preprocess   out = (bilinear_resize(clamp((img + 1) / 2, 0, 1)) - mean) / std          (:71-82)
//   parsing_mask       mask = bilinear_resize(float(argmax_k logits > 0 and != 16)) > 0.5          (:87, :103-109)
//
// Both resizes are F.interpolate(mode='bilinear', align_corners=False, scale_factor=s): source coordinate
// (dst + 0.5) / s - 0.5 clamped at 0, neighbours clamped at the border.
#include "common.cuh"

namespace cagc {

struct Bilerp {
    int i0, i1;
    float l1;   // weight of i1
};

__device__ __forceinline__ Bilerp bilerp_index(int dst, float inv_scale, int in_size) {
    float src = ((float)dst + 0.5f) * inv_scale - 0.5f;
    if (src < 0.f) src = 0.f;
    Bilerp r;
    r.i0 = min((int)src, in_size - 1);
    r.i1 = min(r.i0 + 1, in_size - 1);
    r.l1 = src - (float)r.i0;
    return r;
}

struct PreP {
    const float* img;
    int64_t sb, sc, sh, sw;
    float* out;
    int N, S, P;
    int64_t on, oc, oh, ow;     // output element strides (NCHW-contiguous or channels-last)
    float inv_scale;
    float mean[3], inv_std[3];
};

__global__ void __launch_bounds__(256) parse_preprocess_kernel(const __grid_constant__ PreP p) {
    const int64_t total = (int64_t)p.N * p.P * p.P;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int x = (int)(i % p.P);
        const int64_t q = i / p.P;
        const int y = (int)(q % p.P);
        const int n = (int)(q / p.P);
        const Bilerp by = bilerp_index(y, p.inv_scale, p.S), bx = bilerp_index(x, p.inv_scale, p.S);
        const float* ip = p.img + n * p.sb;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* ic = ip + c * p.sc;
            auto px = [&](int yy, int xx) {
                const float v = (__ldg(ic + yy * p.sh + xx * p.sw) + 1.f) * 0.5f;
                return fminf(fmaxf(v, 0.f), 1.f);
            };
            const float top = px(by.i0, bx.i0) * (1.f - bx.l1) + px(by.i0, bx.i1) * bx.l1;
            const float bot = px(by.i1, bx.i0) * (1.f - bx.l1) + px(by.i1, bx.i1) * bx.l1;
            const float v = top * (1.f - by.l1) + bot * by.l1;
            p.out[n * p.on + c * p.oc + y * p.oh + x * p.ow] = (v - p.mean[c]) * p.inv_std[c];
        }
    }
}

// foreground bit of one parser pixel: argmax over K class maps (first maximum wins, as torch.argmax), then
// `(label > 0) * (label != 16)` (Util/content_aware_pruning.py:103)
__device__ __forceinline__ float fg_bit(const float* __restrict__ lg, int64_t plane, int K) {
    if (K == 1) {                      // a single plane holds the LABELS themselves (Batch_Img_Parsing's return value)
        const int lab = (int)__ldg(lg);
        return (lab > 0 && lab != 16) ? 1.f : 0.f;
    }
    float m = __ldg(lg);
    int arg = 0;
    for (int k = 1; k < K; ++k) {
        const float v = __ldg(lg + k * plane);
        if (v > m) { m = v; arg = k; }
    }
    return (arg > 0 && arg != 16) ? 1.f : 0.f;
}

__global__ void __launch_bounds__(256) parsing_mask_kernel(const float* __restrict__ logits, float* __restrict__ mask, int N,
                                                           int K, int P, int S, float inv_scale) {
    const int64_t total = (int64_t)N * S * S;
    const int64_t plane = (int64_t)P * P;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int x = (int)(i % S);
        const int64_t q = i / S;
        const int y = (int)(q % S);
        const int n = (int)(q / S);
        const Bilerp by = bilerp_index(y, inv_scale, P), bx = bilerp_index(x, inv_scale, P);
        const float* lg = logits + (int64_t)n * K * plane;
        const float b00 = fg_bit(lg + (int64_t)by.i0 * P + bx.i0, plane, K);
        const float b01 = (bx.i1 != bx.i0) ? fg_bit(lg + (int64_t)by.i0 * P + bx.i1, plane, K) : b00;
        float b10 = b00, b11 = b01;
        if (by.i1 != by.i0) {
            b10 = fg_bit(lg + (int64_t)by.i1 * P + bx.i0, plane, K);
            b11 = (bx.i1 != bx.i0) ? fg_bit(lg + (int64_t)by.i1 * P + bx.i1, plane, K) : b10;
        }
        const float top = b00 * (1.f - bx.l1) + b01 * bx.l1;
        const float bot = b10 * (1.f - bx.l1) + b11 * bx.l1;
        const float v = top * (1.f - by.l1) + bot * by.l1;
        mask[i] = v > 0.5f ? 1.f : 0.f;
    }
}

// The same mask from the parser's LOW-RESOLUTION class scores [N,K,h,w] (element strides: any layout): the final
// F.interpolate(..., (P, P), mode='bilinear', align_corners=True) of BiSeNet.forward (Util/face_parsing/BiSeNet.py:247)
// is evaluated per needed grid point inside the kernel, so the [N,K,P,P] tensor (318 MB at N = 16) is never written.
struct LrP {
    const float* lr;
    int64_t sn, sk, sh, sw;
    float* mask;
    int N, K, h, w, P, S;
    float inv_scale, ry, rx;      // P/S; (h-1)/(P-1), (w-1)/(P-1)
};

__device__ __forceinline__ float fg_bit_lowres(const LrP& p, const float* __restrict__ base, int Y, int X) {
    // align_corners=True source coordinates (ATen area_pixel_compute_source_index): src = dst * (in - 1) / (out - 1)
    const float fy = p.ry * (float)Y, fx = p.rx * (float)X;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < p.h - 1 ? 1 : 0), x1 = x0 + (x0 < p.w - 1 ? 1 : 0);
    const float ly1 = fy - (float)y0, lx1 = fx - (float)x0;
    const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    const float* p00 = base + y0 * p.sh + x0 * p.sw;
    const float* p01 = base + y0 * p.sh + x1 * p.sw;
    const float* p10 = base + y1 * p.sh + x0 * p.sw;
    const float* p11 = base + y1 * p.sh + x1 * p.sw;
    float m = 0.f;
    int arg = 0;
    for (int k = 0; k < p.K; ++k) {
        const int64_t o = k * p.sk;
        const float v = ly0 * (lx0 * __ldg(p00 + o) + lx1 * __ldg(p01 + o)) + ly1 * (lx0 * __ldg(p10 + o) + lx1 * __ldg(p11 + o));
        if (k == 0 || v > m) { m = v; arg = k; }
    }
    return (arg > 0 && arg != 16) ? 1.f : 0.f;
}

__global__ void __launch_bounds__(256) parsing_mask_lowres_kernel(const __grid_constant__ LrP p) {
    const int64_t total = (int64_t)p.N * p.S * p.S;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int x = (int)(i % p.S);
        const int64_t q = i / p.S;
        const int y = (int)(q % p.S);
        const int n = (int)(q / p.S);
        const Bilerp by = bilerp_index(y, p.inv_scale, p.P), bx = bilerp_index(x, p.inv_scale, p.P);
        const float* base = p.lr + n * p.sn;
        const float b00 = fg_bit_lowres(p, base, by.i0, bx.i0);
        const float b01 = (bx.i1 != bx.i0) ? fg_bit_lowres(p, base, by.i0, bx.i1) : b00;
        float b10 = b00, b11 = b01;
        if (by.i1 != by.i0) {
            b10 = fg_bit_lowres(p, base, by.i1, bx.i0);
            b11 = (bx.i1 != bx.i0) ? fg_bit_lowres(p, base, by.i1, bx.i1) : b10;
        }
        const float top = b00 * (1.f - bx.l1) + b01 * bx.l1;
        const float bot = b10 * (1.f - bx.l1) + b11 * bx.l1;
        p.mask[i] = (top * (1.f - by.l1) + bot * by.l1) > 0.5f ? 1.f : 0.f;
    }
}

static inline unsigned kd_grid_1d(int64_t work_items, int per_block) {
    int64_t blocks = ceil_div<int64_t>(work_items, per_block);
    const int64_t cap = (int64_t)kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace cagc

using namespace cagc;

extern "C" {

int cagc_parse_preprocess(cagc_stream_t stream_, const float* img, int64_t sb, int64_t sc, int64_t sh, int64_t sw,
                          float* out, int N, int S, int P, int channels_last) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(img && out, "parse_preprocess: null pointer");
    CAGC_REQUIRE(N >= 0 && S >= 1 && P >= 1, "parse_preprocess: bad size");
    if (N == 0) return 0;
    PreP p;
    p.img = img; p.sb = sb; p.sc = sc; p.sh = sh; p.sw = sw; p.out = out; p.N = N; p.S = S; p.P = P;
    if (channels_last) { p.on = (int64_t)P * P * 3; p.oc = 1; p.oh = (int64_t)P * 3; p.ow = 3; }
    else { p.on = (int64_t)3 * P * P; p.oc = (int64_t)P * P; p.oh = P; p.ow = 1; }
    p.inv_scale = 1.f / ((float)P / (float)S);          // F.interpolate uses 1 / scale_factor
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    for (int c = 0; c < 3; ++c) { p.mean[c] = mean[c]; p.inv_std[c] = 1.f / stdv[c]; }
    parse_preprocess_kernel<<<kd_grid_1d((int64_t)N * P * P, 256), 256, 0, stream>>>(p);
    return launched("parse_preprocess_kernel");
}

int cagc_parsing_mask(cagc_stream_t stream_, const float* logits, float* mask, int N, int K, int P, int S) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(logits && mask, "parsing_mask: null pointer");
    CAGC_REQUIRE(N >= 0 && K >= 1 && S >= 1 && P >= 1, "parsing_mask: bad size");
    if (N == 0) return 0;
    parsing_mask_kernel<<<kd_grid_1d((int64_t)N * S * S, 256), 256, 0, stream>>>(logits, mask, N, K, P, S,
                                                                                 1.f / ((float)S / (float)P));
    return launched("parsing_mask_kernel");
}

int cagc_parsing_mask_lowres(cagc_stream_t stream_, const float* scores, int64_t sn, int64_t sk, int64_t sh, int64_t sw,
                             float* mask, int N, int K, int h, int w, int P, int S) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(scores && mask, "parsing_mask_lowres: null pointer");
    CAGC_REQUIRE(N >= 0 && K >= 1 && h >= 1 && w >= 1 && P >= 2 && S >= 1, "parsing_mask_lowres: bad size");
    if (N == 0) return 0;
    LrP p;
    p.lr = scores; p.sn = sn; p.sk = sk; p.sh = sh; p.sw = sw; p.mask = mask;
    p.N = N; p.K = K; p.h = h; p.w = w; p.P = P; p.S = S;
    p.inv_scale = 1.f / ((float)S / (float)P);
    p.ry = (float)(h - 1) / (float)(P - 1);
    p.rx = (float)(w - 1) / (float)(P - 1);
    parsing_mask_lowres_kernel<<<kd_grid_1d((int64_t)N * S * S, 256), 256, 0, stream>>>(p);
    return launched("parsing_mask_lowres_kernel");
}

}  // extern "C"
