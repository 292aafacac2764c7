// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (B200).
//
// GEMM view of the modulated convolution (model.py:241-289 in the fused form of SURVEY.md App. B):
//     M = 128 output pixels per CTA (a bb x bh x bw box of the NHWC-p activation tensor),
//     N = output channels (<= 256 per CTA, multiple of 16),
//     K = taps x input channels, consumed as (tap, 32-channel chunk) stages.
// A operand: one TMA 4-D box load per stage from the *shifted* activation tensor
//     coords (c0, x0 + dx, y0 + dy, b0); the halo of the convolution is TMA out-of-bounds zero fill.
// B operand: one TMA 3-D box load per stage from the K-major weight slabs [tap][n][k].
// Both land in shared memory in the 128-byte swizzled K-major layout that tcgen05.mma reads.
// One elected thread issues tcgen05.mma.kind::tf32 (M=128, N, K=8) into a TMEM accumulator;
// four epilogue warps read it back with tcgen05.ld, apply demodulation / noise / bias / leaky-ReLU
// and store NHWC-p.  Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = epilogue.
// Two CTAs are resident per SM (<= 110 KB of shared memory, <= 256 TMEM columns each), so one CTA's
// epilogue overlaps the other's main loop.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "conv_params.cuh"

namespace cagc {
namespace tc {

constexpr int kThreads = 192;
constexpr int kThreadsPersist = 320;             // halo-tile conv kernel: TMA + MMA + 8 epilogue warps
constexpr int kPersistEpiWarps = 16;             // persistent conv kernel: at most 16 epilogue warps (4 per TMEM lane quarter)
constexpr int kThreadsPersistP = 64 + 32 * kPersistEpiWarps;
// Epilogue warps of one persistent launch: the epilogue is a dependent chain per 16-column chunk (TMEM load -> transpose ->
// math -> store, ~300 instructions), so its throughput scales with the number of warps running it.  Layers with few K
// iterations per work item (1x1 convolutions, the up-convolutions of the narrow end) are bound by it and take 16 warps
// (measured: 128->256 1x1 @128^2 with residual 216 -> 150 us); the compute-bound ones keep 8 and the deeper ring
// (512ch @64^2: 386 us with 8 warps / 4 stages, 440 us with 16 warps / 3 stages).
static inline int persist_epi_warps(int k_iters) {
    static const int env = [] { const char* e = getenv("CAGC_TC_EPI_WARPS"); return e ? atoi(e) : 0; }();
    if (env == 8 || env == 16) return env;
    return k_iters <= 24 ? 16 : 8;
}
constexpr int kTileM = 128;
constexpr int kChunkK = 32;                      // fp32 elements per 128-byte swizzle row
constexpr int kABytes = kTileM * kChunkK * 4;    // 16 KB
constexpr int kMaxStages = 12;
constexpr int kMaxPhases = 4;
constexpr int kSmemBudget = 110 * 1024;          // two CTAs per SM (weight-gradient kernels)
constexpr int kSmemBudget1 = 220 * 1024;         // one CTA per SM
// convolution kernels: 16 KB of static shared memory per CTA belong to the epilogue's transpose staging
constexpr int kEpiStageBytes = 4 * 32 * 32 * 4;
constexpr int kConvBudget1 = kSmemBudget1 - kEpiStageBytes;   // one CTA per SM (two M sub-tiles, > 256 TMEM columns)

struct TcParams {
    uint32_t epi_off;           // byte offset of the epilogue staging tiles inside dynamic smem (transposing epilogue only)
    const float* out_scale;
    const float* noise;
    const float* noise_w;
    const float* bias;
    const float* residual;
    float act_gain;
    float* out;
    int B, Ho, Wo;              // iteration domain
    int bw, bh, bb;             // box: bw*bh*bb == 128
    int tiles_x, tiles_y;       // tiles per sample group
    int k_valid;                // in_pitch
    int n_pitch, out_valid;     // output pitch / logical channels
    int n_tile;                 // MMA N of this launch's full tiles (multiple of 16, <= 256)
    int n_rows;                 // rows per weight slab
    int Hout, Wout, out_stride, out_oy, out_ox;
    int64_t noise_bstride;
    int act, ntaps, stages, in_stride;
    int mt, tiles_total;        // M sub-tiles (128 pixels each) per CTA sharing one weight tile; number of pixel tiles
    int it_per_split;           // conv_tc_kernel split-K: (tap, chunk) iterations per blockIdx.z (0: no split)
    float* ws;                  // split-K: raw partial sums, slab z at ws + z * ws_slab (layout of out)
    int64_t ws_slab;
    uint32_t b_bytes;           // bytes of one B stage
    Tap taps[kMaxTaps];
    // persistent kernel only: several "phases" (the four output parities of the stride-2 transposed convolution)
    // in ONE launch; a phase has its own taps, output lattice and pixel-tile grid.  Work items are enumerated
    // phase-major (most taps first) and dealt round-robin to the CTAs, which balances the mixed costs.
    int nphase;
    int ph_item0[kMaxPhases + 1], ph_tap0[kMaxPhases], ph_ntaps[kMaxPhases];
    int ph_gmin;                // smallest item count of a phase (tile-group-major enumeration, see decode)
    int ph_Ho[kMaxPhases], ph_Wo[kMaxPhases], ph_oy[kMaxPhases], ph_ox[kMaxPhases];
    int ph_tx[kMaxPhases], ph_ty[kMaxPhases], ph_tiles[kMaxPhases];
    int ph_z0[kMaxPhases + 1], ph_itper[kMaxPhases];   // conv_tc_kernel, phased split-K: blockIdx.z range and iterations per split of each phase
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps (sticky launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("cagc conv_tc: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y,
                   threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tc_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------------------------
// Epilogue.  tcgen05.ld hands every lane one accumulator ROW (= one output pixel): storing from that layout makes
// each 16-byte store of a warp touch 32 different cache lines -- measured, that request rate (not the tensor pipe,
// not L2->SMEM operand traffic) is what held every HBM-heavy layer at ~0.8 TB/s of output.  So each warp transposes
// its 32 rows x CW columns through a swizzled shared-memory tile: afterwards 8 (CW = 32) or 4 (CW = 16) consecutive
// lanes own the 16-byte segments of ONE pixel, a warp store covers 4 / 8 pixels x 128 / 64 contiguous bytes (whole
// sectors), the residual is read with the same pattern, and a lane's four channels -- hence its bias -- are fixed.
// ------------------------------------------------------------------------------------------------
struct EpiArgs {
    const float* out_scale;   // [B][n_pitch] or null
    const float* bias;        // [n_pitch] or null
    const float* residual;    // layout of out, or null
    float* out;
    int n_pitch, out_valid, act, has_noise;
    float act_gain;
};

template <int CW>
__device__ __forceinline__ void epi_chunk(const EpiArgs& e, float* st, uint32_t taddr, int lane, int n, bool rvalid,
                                          int64_t roff, float rnz, int rb) {
    constexpr int NSEG = CW / 4;          // 16-byte segments per row
    constexpr int RI = 32 / NSEG;         // rows covered by one warp-wide store
    float v[CW];
    if (CW == 32) tc_ld32(taddr, v); else tc_ld16(taddr, v);
    // row `lane`, segment k -> physical segment k ^ (lane % NSEG): conflict-free 128-bit shared stores and loads
#pragma unroll
    for (int k = 0; k < NSEG; ++k)
        st4(st + (lane * NSEG + (k ^ (lane & (NSEG - 1)))) * 4, make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
    __syncwarp();
    const int seg = lane & (NSEG - 1), rsub = lane / NSEG;
    const int cn = n + seg * 4;
    const bool cvalid = cn < e.n_pitch;
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e.bias && cvalid) b4 = ldg4(e.bias + cn);
    // rows of one warp almost always belong to one sample: then the demodulation scale is per lane, loaded once
    const int b0 = __shfl_sync(0xffffffffu, rb, 0);
    const bool uniform_b = __all_sync(0xffffffffu, rb == b0 || !rvalid);
    float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f);
    if (e.out_scale && uniform_b && cvalid) s4 = ldg4(e.out_scale + (int64_t)b0 * e.n_pitch + cn);
    // pass 1: row descriptors through shuffles, and ALL residual loads of the chunk in flight at once (the stores
    // of pass 2 may alias e.residual as far as the compiler knows, so loads issued inside pass 2 would each expose a
    // full HBM round trip: measured 2x on the residual-carrying 1x1 / data-gradient convolutions)
    int64_t offs[NSEG];
    float nzs[NSEG];
    int bs[NSEG];
    bool oks[NSEG];
    float4 rs[NSEG];
#pragma unroll
    for (int i = 0; i < NSEG; ++i) {
        const int row = i * RI + rsub;
        const bool valid = __shfl_sync(0xffffffffu, (int)rvalid, row) != 0;
        const int off_lo = __shfl_sync(0xffffffffu, (int)(uint32_t)(roff & 0xffffffffll), row);
        const int off_hi = __shfl_sync(0xffffffffu, (int)(roff >> 32), row);
        nzs[i] = e.has_noise ? __shfl_sync(0xffffffffu, rnz, row) : 0.f;
        bs[i] = uniform_b ? b0 : __shfl_sync(0xffffffffu, rb, row);
        oks[i] = valid && cvalid;
        offs[i] = ((int64_t)off_hi << 32) | (uint32_t)off_lo;
        rs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e.residual && oks[i]) rs[i] = ldg4(e.residual + offs[i] + cn);
    }
#pragma unroll
    for (int i = 0; i < NSEG; ++i) {
        if (!oks[i]) continue;
        const int row = i * RI + rsub;
        const float4 x = ld4(st + (row * NSEG + (seg ^ (row & (NSEG - 1)))) * 4);
        float o[4] = {x.x, x.y, x.z, x.w};
        if (e.out_scale) {
            const float4 sc = uniform_b ? s4 : ldg4(e.out_scale + (int64_t)bs[i] * e.n_pitch + cn);
            o[0] *= sc.x; o[1] *= sc.y; o[2] *= sc.z; o[3] *= sc.w;
        }
        if (e.has_noise) { o[0] += nzs[i]; o[1] += nzs[i]; o[2] += nzs[i]; o[3] += nzs[i]; }
        if (e.bias) { o[0] += b4.x; o[1] += b4.y; o[2] += b4.z; o[3] += b4.w; }
        if (e.act == kActLrelu) {
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = lrelu_gain(o[j], e.act_gain);
        }
        if (e.residual) {
            if (e.act == kActMaskRef) {
                o[0] = rs[i].x > 0.f ? o[0] : 0.f; o[1] = rs[i].y > 0.f ? o[1] : 0.f;
                o[2] = rs[i].z > 0.f ? o[2] : 0.f; o[3] = rs[i].w > 0.f ? o[3] : 0.f;
            } else {
                o[0] += rs[i].x; o[1] += rs[i].y; o[2] += rs[i].z; o[3] += rs[i].w;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (cn + j >= e.out_valid) o[j] = 0.f;
        st4(e.out + offs[i] + cn, make_float4(o[0], o[1], o[2], o[3]));
    }
    __syncwarp();          // the staging tile is rewritten by the next chunk
}

constexpr int kEpiTransposeMinN = 128;
// Narrow layers (N < 128: the 39 / 77-channel student layers and their up-conv phases): a pixel's whole channel vector is only 160-320 bytes, the
// per-lane row layout already writes it as a few adjacent 16-byte stores, and the transposing form's extra
// shuffles / shared-memory round trip cost more issue slots than they save (measured: 39ch@256^2 222 -> 287 us, 154->77 up-conv 135 -> 165 us; 128ch@256^2 674 -> 506 us the other way).
__device__ __forceinline__ void epi_rows_direct(const EpiArgs& e, uint32_t taddr, int n0, int n_mma, bool rvalid, int64_t roff,
                                                float rnz, int rb) {
    float* dst = e.out + roff;
    const float* res = e.residual ? e.residual + roff : nullptr;
    const float* sc = e.out_scale ? e.out_scale + (int64_t)rb * e.n_pitch : nullptr;
    for (int c = 0; c < n_mma; c += 16) {
        float v[16];
        tc_ld16(taddr + (uint32_t)c, v);   // warp-collective
        if (!rvalid) continue;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int n = n0 + c + g * 4;
            if (n >= e.n_pitch) break;
            float o[4] = {v[g * 4 + 0], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]};
            if (sc) {
                const float4 s4 = ldg4(sc + n);
                o[0] *= s4.x; o[1] *= s4.y; o[2] *= s4.z; o[3] *= s4.w;
            }
            if (e.has_noise) { o[0] += rnz; o[1] += rnz; o[2] += rnz; o[3] += rnz; }
            if (e.bias) {
                const float4 b4 = ldg4(e.bias + n);
                o[0] += b4.x; o[1] += b4.y; o[2] += b4.z; o[3] += b4.w;
            }
            if (e.act == kActLrelu) {
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = lrelu_gain(o[j], e.act_gain);
            }
            if (res) {
                const float4 r4 = ldg4(res + n);
                if (e.act == kActMaskRef) {
                    o[0] = r4.x > 0.f ? o[0] : 0.f; o[1] = r4.y > 0.f ? o[1] : 0.f;
                    o[2] = r4.z > 0.f ? o[2] : 0.f; o[3] = r4.w > 0.f ? o[3] : 0.f;
                } else {
                    o[0] += r4.x; o[1] += r4.y; o[2] += r4.z; o[3] += r4.w;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j >= e.out_valid) o[j] = 0.f;
            st4(dst + n, make_float4(o[0], o[1], o[2], o[3]));
        }
    }
}

// all column chunks of one 128-row sub-tile for this warp's 32 rows
__device__ __forceinline__ void epi_rows(const EpiArgs& e, float* st, uint32_t taddr, int lane, int n0, int n_mma, bool rvalid,
                                         int64_t roff, float rnz, int rb) {
    if (n_mma < kEpiTransposeMinN || st == nullptr) {
        epi_rows_direct(e, taddr, n0, n_mma, rvalid, roff, rnz, rb);
        return;
    }
    int c = 0;
    for (; c + 32 <= n_mma; c += 32) epi_chunk<32>(e, st, taddr + (uint32_t)c, lane, n0 + c, rvalid, roff, rnz, rb);
    if (c < n_mma) epi_chunk<16>(e, st, taddr + (uint32_t)c, lane, n0 + c, rvalid, roff, rnz, rb);
}

// Two epilogue warps per TMEM lane quarter (persistent kernel): warp `part` of the pair takes every other 16-column
// chunk.  One warp per scheduler runs the epilogue's dependent chain (TMEM load -> transpose -> math -> store) at
// a few instructions per 10 cycles; layers with little K per output (1x1 convolutions, the up-convolutions' large
// outputs) were bound by it.  16-column chunks keep the staging at 2 KB per warp (8 warps: the same 16 KB).
__device__ __forceinline__ void epi_rows_pair(const EpiArgs& e, float* st, uint32_t taddr, int lane, int n0, int n_mma,
                                              bool rvalid, int64_t roff, float rnz, int rb, int part, int nparts = 2) {
    if (n_mma < kEpiTransposeMinN || st == nullptr) {
        // direct form: the warps of a lane quarter split the column range in pieces (multiples of 16)
        const int per = (((n_mma >> 4) + nparts - 1) / nparts) << 4;
        const int c0 = part * per, c1 = min(n_mma, c0 + per);
        if (c1 > c0) epi_rows_direct(e, taddr + (uint32_t)c0, n0 + c0, c1 - c0, rvalid, roff, rnz, rb);
        return;
    }
    for (int c = part * 16; c < n_mma; c += 16 * nparts) epi_chunk<16>(e, st, taddr + (uint32_t)c, lane, n0 + c, rvalid, roff, rnz, rb);
}

// One elected lane of a converged warp.  The TMA / MMA warps keep their control flow warp-uniform and put only
// the asynchronous instruction itself under the election: every address / descriptor computation then stays on
// the uniform datapath, and the single-thread issue loop (the real bound of narrow-N tcgen05 tiles: ~150 cycles
// per MMA when the descriptors were rebuilt in vector registers by a lone lane) shrinks to a few instructions.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// low / high words of the K-major SWIZZLE_128B descriptor (see make_desc_sw128): only the 14-bit start-address
// field of the low word changes between MMAs, so stepping along K (+32 bytes) or M (+128 rows) is a 32-bit add
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo_sw128(uint32_t saddr) { return ((saddr >> 4) & 0x3fffu) | (1u << 16); }
__device__ __forceinline__ void tc_mma_tf32_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the (up to) four K = 8 steps of one 32-channel stage for one (A tile, B tile) pair
__device__ __forceinline__ void mma_stage_k(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, int kk, bool fresh) {
    if (kk == 4) {
        tc_mma_tf32_lh(d, a_lo, kDescHiSw128, b_lo, kDescHiSw128, idesc, fresh ? 0u : 1u);
        tc_mma_tf32_lh(d, a_lo + 2, kDescHiSw128, b_lo + 2, kDescHiSw128, idesc, 1u);
        tc_mma_tf32_lh(d, a_lo + 4, kDescHiSw128, b_lo + 4, kDescHiSw128, idesc, 1u);
        tc_mma_tf32_lh(d, a_lo + 6, kDescHiSw128, b_lo + 6, kDescHiSw128, idesc, 1u);
    } else {
        for (int k = 0; k < kk; ++k)
            tc_mma_tf32_lh(d, a_lo + 2 * k, kDescHiSw128, b_lo + 2 * k, kDescHiSw128, idesc, (fresh && k == 0) ? 0u : 1u);
    }
}

// Warp-uniform issue: all 32 lanes execute the (uniform) descriptor arithmetic, the lane election happens INSIDE the
// asm block and only predicates the asynchronous instruction.  The C++ around it has no divergent region, so the
// compiler keeps descriptors / barrier addresses on the uniform datapath (no R2UR round trips, no BSSY/BSYNC):
// ncu showed ~200 issue-warp instructions per stage with the `if (elect_one())` form, i.e. the issuing warp -- not
// the tensor pipe -- bounded every layer whose MMAs are short (N <= 128).
__device__ __forceinline__ void mma_k4_elect(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b64 da, db;\n\t.reg .b32 ta, tb;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "setp.ne.b32 pa, %5, 0;\n\t"
        "setp.eq.b32 pt, 0, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, pa;\n\t"
        "add.u32 ta, %1, 2;\n\tadd.u32 tb, %2, 2;\n\t"
        "mov.b64 da, {ta, %3};\n\tmov.b64 db, {tb, %3};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, pt;\n\t"
        "add.u32 ta, %1, 4;\n\tadd.u32 tb, %2, 4;\n\t"
        "mov.b64 da, {ta, %3};\n\tmov.b64 db, {tb, %3};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, pt;\n\t"
        "add.u32 ta, %1, 6;\n\tadd.u32 tb, %2, 6;\n\t"
        "mov.b64 da, {ta, %3};\n\tmov.b64 db, {tb, %3};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, pt;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(kDescHiSw128), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_k1_elect(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred pe, pa;\n\t.reg .b64 da, db;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "setp.ne.b32 pa, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, pa;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(kDescHiSw128), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the (up to) four K = 8 steps of one 32-channel stage for one (A tile, B tile) pair; call with the warp converged
__device__ __forceinline__ void mma_stage_k_uniform(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, int kk, bool fresh) {
    if (kk == 4) {
        mma_k4_elect(d, a_lo, b_lo, idesc, fresh ? 0u : 1u);
    } else {
        for (int k = 0; k < kk; ++k) mma_k1_elect(d, a_lo + 2 * k, b_lo + 2 * k, idesc, (fresh && k == 0) ? 0u : 1u);
    }
}
__device__ __forceinline__ void tc_commit_elect(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(bar) : "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor: 8-row groups of 1024 bytes
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64))
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;            // leading byte offset (unused for swizzled K-major; canonical value 1)
    d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;            // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;            // SWIZZLE_128B
    return d;
}

__global__ void __launch_bounds__(kThreads, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 1];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    float* const epi_stage_base = (p.epi_off == 0xffffffffu) ? nullptr
        : reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + p.epi_off);
    const int S = p.stages;
    const int MT = p.mt;
    const uint32_t stage_bytes = (uint32_t)MT * kABytes + p.b_bytes;
    auto a_addr = [&](int s, int j) { return smem_base + (uint32_t)s * stage_bytes + (uint32_t)j * kABytes; };
    auto b_addr = [&](int s) { return smem_base + (uint32_t)s * stage_bytes + (uint32_t)MT * kABytes; };
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
    const uint32_t acc_bar = bar0 + 8u * (2 * kMaxStages);

    // ---- phased split-K (all output parities of a small transposed convolution in one launch): blockIdx.z selects the
    // phase -- its taps, output lattice and extent -- and the split of that phase's K loop
    int f_tap0 = 0, f_ntaps = p.ntaps, f_Ho = p.Ho, f_Wo = p.Wo, f_oy = p.out_oy, f_ox = p.out_ox;
    int f_tx = p.tiles_x, f_ty = p.tiles_y, f_tiles = p.tiles_total, f_itper = p.it_per_split, zs = blockIdx.z;
    if (p.nphase > 1) {
        int f = 0;
        while (f + 1 < p.nphase && (int)blockIdx.z >= p.ph_z0[f + 1]) ++f;
        zs = (int)blockIdx.z - p.ph_z0[f];
        f_tap0 = p.ph_tap0[f]; f_ntaps = p.ph_ntaps[f];
        f_Ho = p.ph_Ho[f]; f_Wo = p.ph_Wo[f]; f_oy = p.ph_oy[f]; f_ox = p.ph_ox[f];
        f_tx = p.ph_tx[f]; f_ty = p.ph_ty[f]; f_tiles = p.ph_tiles[f]; f_itper = p.ph_itper[f];
        if ((int)blockIdx.x * MT >= f_tiles) return;      // the grid is sized for the largest phase
    }

    // ---- tile coordinates of the (up to two) 128-pixel sub-tiles; a sub-tile past the end is parked
    // at batch index B (fully out of bounds: TMA zero-fills, the epilogue stores nothing)
    int x0s[2], y0s[2], b0s[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        int t = blockIdx.x * MT + j;
        if (j < MT && t < f_tiles) {
            x0s[j] = (t % f_tx) * p.bw;
            t /= f_tx;
            y0s[j] = (t % f_ty) * p.bh;
            b0s[j] = (t / f_ty) * p.bb;
        } else {
            x0s[j] = 0; y0s[j] = 0; b0s[j] = p.B;
        }
    }
    const int n0 = blockIdx.y * p.n_tile;
    int n_mma = p.n_tile;
    {
        const int rem = ((p.n_pitch - n0) + 15) & ~15;
        if (rem < n_mma) n_mma = rem;
    }
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < MT * n_mma) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base_slot;

    const int nk = (p.k_valid + kChunkK - 1) / kChunkK;
    // split-K (small layers): blockIdx.z owns the (tap, chunk) iterations [it0, total) of the full K loop and leaves raw
    // partial sums in its workspace slab; splitk_epilogue_kernel adds the slabs in a fixed order and applies the epilogue
    const int it0 = f_itper ? zs * f_itper : 0;
    const int total = f_itper ? min(f_ntaps * nk, it0 + f_itper) : f_ntaps * nk;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        for (int it = it0; it < total; ++it) {
            const int s = (it - it0) % S;
            const uint32_t ph = (uint32_t)((it - it0) / S) & 1u;
            mbar_wait(empty_bar(s), ph ^ 1u);
            const int tap = it / nk, kc = it - tap * nk;
            const Tap tp = p.taps[f_tap0 + tap];
            if (elect_one()) {
                mbar_expect_tx(full_bar(s), stage_bytes);
                for (int j = 0; j < MT; ++j)
                    tma_load_4d(a_addr(s, j), &map_a, full_bar(s), kc * kChunkK, x0s[j] * p.in_stride + tp.dx,
                                y0s[j] * p.in_stride + tp.dy, b0s[j]);
                tma_load_3d(b_addr(s), &map_b, full_bar(s), kc * kChunkK, n0, tp.slab);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // =============================== MMA issuer =================================
        // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, K-major both, N>>3, M>>4
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_mma >> 3) << 17) |
                               ((uint32_t)(kTileM >> 4) << 24);
        for (int it = it0; it < total; ++it) {
            const int s = (it - it0) % S;
            const uint32_t ph = (uint32_t)((it - it0) / S) & 1u;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const int kc = it % nk;
            int kk = (p.k_valid - kc * kChunkK + 7) >> 3;   // 8-wide tf32 MMAs with real data
            if (kk > 4) kk = 4;
            const uint32_t b_lo = desc_lo_sw128(b_addr(s));
            if (elect_one()) {
                for (int j = 0; j < MT; ++j)
                    mma_stage_k(tmem_acc + (uint32_t)(j * n_mma), desc_lo_sw128(a_addr(s, j)), b_lo, idesc, kk, it == it0);
                tc_commit(empty_bar(s));   // frees the smem stage once these MMAs have read it
                if (it == total - 1) tc_commit(acc_bar);            // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // =============================== epilogue ===================================
        const int q = warp & 3;            // TMEM lane quarter this warp may access
        const int m = q * 32 + lane;       // accumulator row == pixel of the box
        const int lx = m % p.bw;
        const int ly = (m / p.bw) % p.bh;
        const int lb = m / (p.bw * p.bh);
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        const EpiArgs ea = f_itper
            ? EpiArgs{nullptr, nullptr, nullptr, p.ws + (int64_t)zs * p.ws_slab, p.n_pitch, p.n_pitch, 0, 0, 1.f}
            : EpiArgs{p.out_scale, p.bias, p.residual, p.out, p.n_pitch, p.out_valid, p.act, p.noise != nullptr, p.act_gain};
      for (int j = 0; j < MT; ++j) {
        const int ox = x0s[j] + lx, oy = y0s[j] + ly, b = b0s[j] + lb;
        const bool pvalid = (ox < f_Wo) && (oy < f_Ho) && (b < p.B);
        const int yy = oy * p.out_stride + f_oy, xx = ox * p.out_stride + f_ox;
        float nz = 0.f;
        if (p.noise && pvalid)
            nz = __ldg(p.noise_w) * __ldg(p.noise + (int64_t)b * p.noise_bstride + (int64_t)yy * p.Wout + xx);
        const int64_t roff = (((int64_t)b * p.Hout + yy) * p.Wout + xx) * p.n_pitch;
        epi_rows(ea, (epi_stage_base ? epi_stage_base + q * 1024 : nullptr), tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * n_mma), lane, n0, n_mma, pvalid,
                 roff, nz, b);
      }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(tmem_cols) : "memory");
    }
}


// ================================================================================================
// Persistent variant: one CTA per SM loops over (n-tile, pixel super-tile) work items; the TMEM
// accumulator is double buffered, so the four epilogue warps drain tile i while the MMA warp already
// accumulates tile i+1 and the TMA warp runs further ahead through a deep (up to 8-stage, ~200 KB)
// shared-memory ring.  TMEM allocation, barrier setup and tensor-map fetch are paid once per SM, which
// is what the narrow, HBM-bound student layers (K = 9 taps x 2 chunks) were dominated by.
// ================================================================================================
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(kThreadsPersistP, 1)
conv_tc_persist_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 4];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    float* const epi_stage_base = (p.epi_off == 0xffffffffu) ? nullptr
        : reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + p.epi_off);
    const int S = p.stages, MT = p.mt;
    const uint32_t stage_bytes = (uint32_t)MT * kABytes + p.b_bytes;
    // stage s: [MT A tiles of 16 KB][B tile] at smem_base + s * stage_bytes
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * kMaxStages + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * kMaxStages + 2 + a); };

    const int acc_stride = MT * p.n_tile;           // TMEM columns per accumulator buffer
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 2 * acc_stride) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), (blockDim.x >> 5) - 2);   // one arrival per epilogue warp (8 or 16, chosen by the host)            // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base_slot;

    const int nk = (p.k_valid + kChunkK - 1) / kChunkK;
    const int items = p.ph_item0[p.nphase];

    // phase, weight tile and sub-tile coordinates of work item `item`
    auto decode = [&](int item, int* x0s, int* y0s, int* b0s, int& n0, int& n_mma) {
        int f = 0, local, n_idx, m_idx;
        if (p.nphase > 1) {
            // Several phases (output parities of the transposed convolution) read the SAME input tiles: enumerate the
            // items tile-group-major -- the k-th item of every phase back to back, N tiles innermost -- so that the
            // phases (and N tiles) of a pixel tile run at about the same time on different CTAs and share the input
            // through L2.  Phase-major order re-read the input from DRAM once per phase (ncu: 1.08 GB for a 268 MB input).
            const int head = p.ph_gmin * p.nphase;
            if (item < head) {
                local = item / p.nphase;
                f = item - local * p.nphase;
                // the phases carry 1/2/2/4 taps: with a grid that is a multiple of nphase a CTA would keep the same
                // phase on every round (item += gridDim.x) -- rotate the phase by the round number
                if (gridDim.x % p.nphase == 0) f = (f + item / (int)gridDim.x) % p.nphase;
            } else {
                int t = item - head;
                for (f = 0; f + 1 < p.nphase; ++f) {
                    const int extra = p.ph_item0[f + 1] - p.ph_item0[f] - p.ph_gmin;
                    if (t < extra) break;
                    t -= extra;
                }
                local = p.ph_gmin + t;
            }
            const int n_tiles_ = (p.n_rows + p.n_tile - 1) / p.n_tile;
            m_idx = local / n_tiles_;
            n_idx = local - m_idx * n_tiles_;
        } else {
            local = item;
            const int m_super = (p.ph_tiles[0] + MT - 1) / MT;
            n_idx = local / m_super;
            m_idx = local - n_idx * m_super;
        }
        n0 = n_idx * p.n_tile;
        n_mma = p.n_tile;
        const int rem = ((p.n_pitch - n0) + 15) & ~15;
        if (rem < n_mma) n_mma = rem;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            int t = m_idx * MT + j;
            if (j < MT && t < p.ph_tiles[f]) {
                x0s[j] = (t % p.ph_tx[f]) * p.bw;
                t /= p.ph_tx[f];
                y0s[j] = (t % p.ph_ty[f]) * p.bh;
                b0s[j] = (t / p.ph_ty[f]) * p.bb;
            } else {
                x0s[j] = 0; y0s[j] = 0; b0s[j] = p.B;
            }
        }
        return f;
    };

    // Both issue loops keep RUNNING state (stage address / barrier address / chunk and tap counters advanced by adds
    // with a wrap) instead of re-deriving it from the iteration index: with 64-cycle N = 128 MMAs the single issuing
    // thread is the bound of the kernel (ncu: ~200 instructions, among them two integer divisions, per 8-MMA stage).
    if (warp == 0) {
        int s = 0;
        uint32_t ph = 0;
        uint32_t st_addr = smem_base, fb = full_bar(0);
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            int x0s[2], y0s[2], b0s[2], n0, n_mma;
            const int f = decode(item, x0s, y0s, b0s, n0, n_mma);
            const int ntap = p.ph_ntaps[f], tap0 = p.ph_tap0[f];
            const int ax0 = x0s[0] * p.in_stride, ay0 = y0s[0] * p.in_stride;
            const int ax1 = x0s[1] * p.in_stride, ay1 = y0s[1] * p.in_stride;
            for (int tap = 0; tap < ntap; ++tap) {
                const Tap tp = p.taps[tap0 + tap];
                for (int kc = 0; kc < nk; ++kc) {
                    mbar_wait(fb + 8u * kMaxStages, ph ^ 1u);
                    if (elect_one()) {
                        mbar_expect_tx(fb, stage_bytes);
                        tma_load_4d(st_addr, &map_a, fb, kc * kChunkK, ax0 + tp.dx, ay0 + tp.dy, b0s[0]);
                        if (MT == 2) tma_load_4d(st_addr + kABytes, &map_a, fb, kc * kChunkK, ax1 + tp.dx, ay1 + tp.dy, b0s[1]);
                        tma_load_3d(st_addr + (uint32_t)MT * kABytes, &map_b, fb, kc * kChunkK, n0, tp.slab);
                    }
                    __syncwarp();
                    if (++s == S) { s = 0; ph ^= 1u; st_addr = smem_base; fb = bar0; }
                    else { st_addr += stage_bytes; fb += 8u; }
                }
            }
        }
    } else if (warp == 1) {
        int s = 0, acc = 0;
        uint32_t ph = 0, aph = 0;
        const uint32_t lo_base = desc_lo_sw128(smem_base);
        const uint32_t st_step = stage_bytes >> 4, a_step = (uint32_t)kABytes >> 4, b_off = (uint32_t)MT * a_step;
        uint32_t lo_s = lo_base, fb = full_bar(0);
        int kk_last = (p.k_valid - (nk - 1) * kChunkK + 7) >> 3;
        if (kk_last > 4) kk_last = 4;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            int x0s[2], y0s[2], b0s[2], n0, n_mma;
            const int f = decode(item, x0s, y0s, b0s, n0, n_mma);
            const int per_item = p.ph_ntaps[f] * nk;
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_mma >> 3) << 17) |
                                   ((uint32_t)(kTileM >> 4) << 24);
            mbar_wait(tempty_bar(acc), aph ^ 1u);        // epilogue has drained this accumulator buffer
            tc_fence_after();
            const uint32_t d0 = tmem_acc + (uint32_t)(acc * acc_stride);
            int kc = 0;
            for (int it = 0; it < per_item; ++it) {
                mbar_wait(fb, ph);
                tc_fence_after();
                const int kk = (kc == nk - 1) ? kk_last : 4;
                mma_stage_k_uniform(d0, lo_s, lo_s + b_off, idesc, kk, it == 0);
                if (MT == 2) mma_stage_k_uniform(d0 + (uint32_t)p.n_tile, lo_s + a_step, lo_s + b_off, idesc, kk, it == 0);
                tc_commit_elect(fb + 8u * kMaxStages);
                if (it == per_item - 1) tc_commit_elect(tfull_bar(acc));
                if (++kc == nk) kc = 0;
                if (++s == S) { s = 0; ph ^= 1u; lo_s = lo_base; fb = bar0; }
                else { lo_s += st_step; fb += 8u; }
            }
            if (++acc == 2) { acc = 0; aph ^= 1u; }
        }
    } else {
        const int q = warp & 3;                   // TMEM lane quarter; warps w and w + 4 share it
        const int part = (warp - 2) >> 2;         // which 16-column chunks this warp takes among the warps of its quarter
        const int nparts = ((int)(blockDim.x >> 5) - 2) >> 2;
        float* const my_stage = epi_stage_base ? epi_stage_base + (warp - 2) * 512 : nullptr;     // 2 KB per warp
        const int m = q * 32 + lane;
        const int lx = m % p.bw;
        const int ly = (m / p.bw) % p.bh;
        const int lb = m / (p.bw * p.bh);
        const float nwv = p.noise ? __ldg(p.noise_w) : 0.f;
        const EpiArgs ea{p.out_scale, p.bias, p.residual, p.out, p.n_pitch, p.out_valid, p.act, p.noise != nullptr, p.act_gain};
        int acc = 0;
        uint32_t aph = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            int x0s[2], y0s[2], b0s[2], n0, n_mma;
            const int f = decode(item, x0s, y0s, b0s, n0, n_mma);
            // noise values requested before the accumulator wait (latency hidden behind the tile's MMAs)
            float nzs[2] = {0.f, 0.f};
            if (p.noise) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int ox = x0s[j] + lx, oy = y0s[j] + ly, b = b0s[j] + lb;
                    if (j < MT && (ox < p.ph_Wo[f]) && (oy < p.ph_Ho[f]) && (b < p.B))
                        nzs[j] = nwv * __ldg(p.noise + (int64_t)b * p.noise_bstride +
                                             (int64_t)(oy * p.out_stride + p.ph_oy[f]) * p.Wout + (ox * p.out_stride + p.ph_ox[f]));
                }
            }
            mbar_wait(tfull_bar(acc), aph);
            tc_fence_after();
            const uint32_t d0 = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * acc_stride);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (j >= MT) break;
                const int ox = x0s[j] + lx, oy = y0s[j] + ly, b = b0s[j] + lb;
                const bool pvalid = (ox < p.ph_Wo[f]) && (oy < p.ph_Ho[f]) && (b < p.B);
                const int yy = oy * p.out_stride + p.ph_oy[f], xx = ox * p.out_stride + p.ph_ox[f];
                const float nz = nzs[j];
                const int64_t roff = (((int64_t)b * p.Hout + yy) * p.Wout + xx) * p.n_pitch;
                epi_rows_pair(ea, my_stage, d0 + (uint32_t)(j * p.n_tile), lane, n0, n_mma, pvalid, roff, nz, b, part, nparts);
            }
            // this warp has finished reading the accumulator buffer: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; aph ^= 1u; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(tmem_cols) : "memory");
    }
}

// ================================================================================================
// Halo-tile variant: the activation window of a (R rows x TW pixels) output tile -- including the halo the
// taps reach into -- is loaded ONCE per 32-channel chunk as a single TMA box of (R + dyspan) x (TW + dxspan)
// pixels; its pixels sit in shared memory in raster order, one 128-byte swizzled row each.  Every tap is then
// the SAME tile seen through a descriptor whose start address is shifted by (dy*Wt + dx) rows: tcgen05 derives
// the 128-byte swizzle from absolute shared-memory address bits, so a start row that is not aligned to the
// 8-row swizzle atom is legal (measured: scripts/probe/umma_shift_probe.cu, exact for every shift).  The M
// dimension of the GEMM is the raster range of the tile (Wt = TW + dxspan positions per row; the dxspan
// positions per row that fall on halo columns produce accumulator rows nobody stores).
// Operand traffic per output pixel drops from ntaps x (once per tap) to (R+dyspan)(TW+dxspan)/(R*TW) ~ 1.3x,
// which is what bounds the narrow, HBM-class student layers (K = 9 taps x 2..5 chunks).
// Weights stream per (chunk, tap) through their own ring.  Persistent, one CTA per SM, TMEM accumulator
// double buffered when 2*MT*N <= 512 columns.  Warp roles as above.
// ================================================================================================
struct HaloParams {
    uint32_t epi_off;           // byte offset of the epilogue staging tiles inside dynamic smem (transposing epilogue only)
    const float* out_scale;
    const float* noise;
    const float* noise_w;
    const float* bias;
    const float* residual;
    float act_gain;
    float* out;
    int B, Ho, Wo;                  // iteration domain (output pixels of this launch)
    int TW, R, Wt, rows_box;        // tile width / rows, raster pitch, box rows (R + dyspan)
    int dx_min, dy_min;
    int w_bslabs;                   // per-sample weight slab sets: sample b reads slab tap_slab + b * w_bslabs (0: shared)
    int strips, ytiles;             // tiles per sample
    int mt;                         // 128-row M sub-tiles per tile: ceil(R*Wt / 128) (raster mode) or R (row mode)
    int row_mode;                   // 1: sub-tile j = output row j of the tile (TW <= 128 pixels, no halo columns in M)
    int sub_rows;                   // raster rows between consecutive sub-tiles: 128 (raster mode) or Wt (row mode)
    int k_valid, n_pitch, out_valid, n_tile, n_rows;
    int Hout, Wout, out_stride, out_oy, out_ox;
    int64_t noise_bstride;
    int act, ntaps, b_stages, acc_bufs;
    uint32_t a_bytes, a_stride, b_bytes;   // box bytes, A buffer stride, bytes of one B stage
    int tap_row[kMaxTaps];          // row shift of tap t inside the halo tile
    int tap_slab[kMaxTaps];
};

constexpr int kHaloMaxB = 8;

__global__ void __launch_bounds__(kThreadsPersist, 1)
conv_tc_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ HaloParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kHaloMaxB + 8];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    float* const epi_stage_base = (p.epi_off == 0xffffffffu) ? nullptr
        : reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + p.epi_off);
    const uint32_t b_base = smem_base + 2u * p.a_stride;
    const int SB = p.b_stages, MT = p.mt;
    const uint32_t bar0 = smem_u32(bars);
    auto bfull = [&](int s) { return bar0 + 8u * s; };
    auto bempty = [&](int s) { return bar0 + 8u * (kHaloMaxB + s); };
    auto afull = [&](int a) { return bar0 + 8u * (2 * kHaloMaxB + a); };
    auto aempty = [&](int a) { return bar0 + 8u * (2 * kHaloMaxB + 2 + a); };
    auto tfull = [&](int a) { return bar0 + 8u * (2 * kHaloMaxB + 4 + a); };
    auto tempty = [&](int a) { return bar0 + 8u * (2 * kHaloMaxB + 6 + a); };

    const int acc_stride = MT * p.n_tile;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < p.acc_bufs * acc_stride) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < SB; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
        for (int a = 0; a < 2; ++a) {
            mbar_init(afull(a), 1); mbar_init(aempty(a), 1);
            mbar_init(tfull(a), 1); mbar_init(tempty(a), 8);      // 8 epilogue warps
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base_slot;

    const int nk = (p.k_valid + kChunkK - 1) / kChunkK;
    const int tiles_per_sample = p.strips * p.ytiles;
    const int m_tiles = tiles_per_sample * p.B;
    const int n_tiles = (p.n_rows + p.n_tile - 1) / p.n_tile;
    const int items = m_tiles * n_tiles;

    auto decode = [&](int item, int& x0, int& y0, int& b, int& n0, int& n_mma) {
        const int n_idx = item / m_tiles;
        int t = item - n_idx * m_tiles;
        n0 = n_idx * p.n_tile;
        n_mma = p.n_tile;
        const int rem = ((p.n_pitch - n0) + 15) & ~15;
        if (rem < n_mma) n_mma = rem;
        b = t / tiles_per_sample;
        t -= b * tiles_per_sample;
        y0 = (t / p.strips) * p.R;
        x0 = (t % p.strips) * p.TW;
    };

    if (warp == 0) {
        int s = 0, ab = 0;
        uint32_t ph = 0, aph = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            int x0, y0, b, n0, n_mma;
            decode(item, x0, y0, b, n0, n_mma);
            for (int kc = 0; kc < nk; ++kc) {
                mbar_wait(aempty(ab), aph ^ 1u);
                // one TMA op per image row of the window (a single large box is no faster, and the row ops can
                // overlap); destinations are 128-byte aligned, which is all the absolute-address swizzle needs
                if (elect_one()) {
                    mbar_expect_tx(afull(ab), p.a_bytes);
                    for (int r = 0; r < p.rows_box; ++r)
                        tma_load_4d(smem_base + (uint32_t)ab * p.a_stride + (uint32_t)(r * p.Wt) * 128u, &map_a, afull(ab),
                                    kc * kChunkK, x0 + p.dx_min, y0 + p.dy_min + r, b);
                }
                __syncwarp();
                if (++ab == 2) { ab = 0; aph ^= 1u; }
                for (int tap = 0; tap < p.ntaps; ++tap) {
                    mbar_wait(bempty(s), ph ^ 1u);
                    if (elect_one()) {
                        mbar_expect_tx(bfull(s), p.b_bytes);
                        tma_load_3d(b_base + (uint32_t)s * p.b_bytes, &map_b, bfull(s), kc * kChunkK, n0,
                                    p.tap_slab[tap] + b * p.w_bslabs);
                    }
                    __syncwarp();
                    if (++s == SB) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // Issue loop with running state: the single issuing thread is the bound of the narrow layers (a 128 x 48 x 8
        // MMA occupies the tensor pipe for 24 cycles), so nothing is re-derived per stage that an add can carry along.
        int s = 0, ab = 0, acc = 0;
        uint32_t ph = 0, aph = 0, tph = 0;
        const uint32_t sub_step = ((uint32_t)p.sub_rows * 128u) >> 4;      // descriptor address units (16 bytes)
        const uint32_t n_tile = (uint32_t)p.n_tile, b_step = p.b_bytes >> 4;
        const uint32_t b_lo0 = desc_lo_sw128(b_base), a_lo0 = desc_lo_sw128(smem_base), a_step = p.a_stride >> 4;
        uint32_t b_lo = b_lo0, bf = bfull(0);
        const int ntaps = p.ntaps;
        int kk_last = (p.k_valid - (nk - 1) * kChunkK + 7) >> 3;
        if (kk_last > 4) kk_last = 4;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            int x0, y0, b, n0, n_mma;
            decode(item, x0, y0, b, n0, n_mma);
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_mma >> 3) << 17) |
                                   ((uint32_t)(kTileM >> 4) << 24);
            mbar_wait(tempty(acc), tph ^ 1u);
            tc_fence_after();
            const uint32_t d0 = tmem_acc + (uint32_t)(acc * acc_stride);
            for (int kc = 0; kc < nk; ++kc) {
                mbar_wait(afull(ab), aph);
                tc_fence_after();
                const int kk = (kc == nk - 1) ? kk_last : 4;
                const uint32_t a_buf_lo = a_lo0 + (uint32_t)ab * a_step;
                for (int tap = 0; tap < ntaps; ++tap) {
                    mbar_wait(bf, ph);
                    tc_fence_after();
                    const uint32_t a_lo = a_buf_lo + (uint32_t)p.tap_row[tap] * 8u;      // 128-byte rows, >> 4
                    // (measured A/B on one box: in this kernel the divergent elect_one() region issues faster than the
                    //  in-asm election that the persistent kernel uses -- 39ch@256^2 229 vs 264 us)
                    if (elect_one()) {
                        uint32_t d = d0, a = a_lo;
                        for (int j = 0; j < MT; ++j) {          // next sub-tile: +sub_rows raster rows
                            mma_stage_k(d, a, b_lo, idesc, kk, kc == 0 && tap == 0);
                            d += n_tile;
                            a += sub_step;
                        }
                        tc_commit(bf + 8u * kHaloMaxB);
                        if (tap == ntaps - 1) {
                            tc_commit(aempty(ab));
                            if (kc == nk - 1) tc_commit(tfull(acc));
                        }
                    }
                    __syncwarp();
                    if (++s == SB) { s = 0; ph ^= 1u; b_lo = b_lo0; bf = bar0; }
                    else { b_lo += b_step; bf += 8u; }
                }
                if (++ab == 2) { ab = 0; aph ^= 1u; }
            }
            if (++acc == p.acc_bufs) { acc = 0; tph ^= 1u; }
        }
    } else {
        const int q = warp & 3;                   // TMEM lane quarter; warps w and w + 4 share it (see epi_rows_pair)
        const int part = (warp - 2) >> 2;
        float* const my_stage = epi_stage_base ? epi_stage_base + (warp - 2) * 512 : nullptr;
        const int m = q * 32 + lane;
        const float nwv = p.noise ? __ldg(p.noise_w) : 0.f;
        const EpiArgs ea{p.out_scale, p.bias, p.residual, p.out, p.n_pitch, p.out_valid, p.act, p.noise != nullptr, p.act_gain};
        int acc = 0;
        uint32_t tph = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            int x0, y0, b, n0, n_mma;
            decode(item, x0, y0, b, n0, n_mma);
            // this thread's noise values for all sub-tiles are requested BEFORE waiting for the accumulator: their
            // latency hides behind the MMAs of the tile instead of adding ~1 us per sub-tile to the epilogue
            float nzs[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                nzs[j] = 0.f;
                if (p.noise && j < MT) {
                    const int pos = j * kTileM + m;
                    const int ry = p.row_mode ? j : pos / p.Wt, rx = p.row_mode ? m : pos - ry * p.Wt;
                    const int ox = x0 + rx, oy = y0 + ry;
                    if ((rx < p.TW) && (ry < p.R) && (ox < p.Wo) && (oy < p.Ho))
                        nzs[j] = nwv * __ldg(p.noise + (int64_t)b * p.noise_bstride +
                                             (int64_t)(oy * p.out_stride + p.out_oy) * p.Wout + (ox * p.out_stride + p.out_ox));
                }
            }
            mbar_wait(tfull(acc), tph);
            tc_fence_after();
            const uint32_t d0 = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * acc_stride);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j >= MT) break;
                const int pos = j * kTileM + m;              // raster position inside the tile
                const int ry = p.row_mode ? j : pos / p.Wt, rx = p.row_mode ? m : pos - ry * p.Wt;
                const int ox = x0 + rx, oy = y0 + ry;
                const bool pvalid = (rx < p.TW) && (ry < p.R) && (ox < p.Wo) && (oy < p.Ho);
                const int yy = oy * p.out_stride + p.out_oy, xx = ox * p.out_stride + p.out_ox;
                const float nz = nzs[j];
                const int64_t roff = (((int64_t)b * p.Hout + yy) * p.Wout + xx) * p.n_pitch;
                epi_rows_pair(ea, my_stage, d0 + (uint32_t)(j * p.n_tile), lane, n0, n_mma, pvalid, roff, nz, b, part);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(acc));
            if (++acc == p.acc_bufs) { acc = 0; tph ^= 1u; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(tmem_cols) : "memory");
    }
}

// elementwise x~ = round_tf32(x * s[b, c])  (s == nullptr: plain rounding); NHWC-p, float4.  blockIdx.y = sample;
// the channel-quad index of a thread advances by a fixed amount per grid stride (no division in the loop)
__global__ void __launch_bounds__(256) modulate_kernel(const float* __restrict__ x, const float* __restrict__ s,
                                                       float* __restrict__ out, int per_sample4, int c4n) {
    const int b = blockIdx.y;
    const int stride = gridDim.x * 256;
    const int step_c = stride % c4n;
    int i = blockIdx.x * 256 + threadIdx.x;
    int c4 = i % c4n;
    const float* xs = x + (int64_t)b * per_sample4 * 4;
    float* os = out + (int64_t)b * per_sample4 * 4;
    const float* sb = s ? s + (int64_t)b * c4n * 4 : nullptr;
#pragma unroll 4
    for (; i < per_sample4; i += stride) {
        float4 v = ldg4(xs + 4 * (int64_t)i);
        if (sb) {
            const float4 sv = ldg4(sb + c4 * 4);
            v.x *= sv.x; v.y *= sv.y; v.z *= sv.z; v.w *= sv.w;
        }
        uint32_t r0, r1, r2, r3;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r0) : "f"(v.x));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r1) : "f"(v.y));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r2) : "f"(v.z));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r3) : "f"(v.w));
        st4(os + 4 * (int64_t)i, make_float4(__uint_as_float(r0), __uint_as_float(r1), __uint_as_float(r2), __uint_as_float(r3)));
        c4 += step_c;
        if (c4 >= c4n) c4 -= c4n;
    }
}

// Per-sample weight slabs: out[b][row][i] = tf32(w[row][i] * s[b][i]) for the K-major slab rows [taps * n_rows][in_pitch]
// of the tensor-pipe forward operand -- the modulation of model.py:248-250 folded into the B operand instead of a pass
// over the activation (a frozen layer without gradients: the teacher's 128^2 / 256^2 layers move 16 x 0.6 .. 1.2 MB of
// weights instead of 0.5 .. 1.1 GB of activations).
__global__ void __launch_bounds__(256) modulate_weights_kernel(const float* __restrict__ w, const float* __restrict__ s,
                                                               float* __restrict__ out, int per_sample4, int c4n) {
    const int b = blockIdx.y;
    const float* sb = s + (int64_t)b * c4n * 4;
    float* os = out + (int64_t)b * per_sample4 * 4;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < per_sample4; i += gridDim.x * 256) {
        float4 v = ldg4(w + 4 * (int64_t)i);
        const float4 sv = ldg4(sb + (i % c4n) * 4);
        v.x *= sv.x; v.y *= sv.y; v.z *= sv.z; v.w *= sv.w;
        uint32_t r0, r1, r2, r3;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r0) : "f"(v.x));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r1) : "f"(v.y));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r2) : "f"(v.z));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r3) : "f"(v.w));
        st4(os + 4 * (int64_t)i, make_float4(__uint_as_float(r0), __uint_as_float(r1), __uint_as_float(r2), __uint_as_float(r3)));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn lookup_encode() {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
        return (EncodeTiledFn)sym;
    return nullptr;
}

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = lookup_encode();      // thread-safe one-time initialisation (C++11 magic static)
    return fn;
}


// ================================================================================================
// Weight gradient on the tensor pipe:  gw[t][i][o] = sum_pixels a[pix + off_a(t)][i] * g[pix*sg + off_g(t)][o]
// GEMM with M = input channels (128 per CTA), N = output channels (<= 256), K = pixels.  Both operands
// are channel-contiguous in memory, i.e. MN-major for the MMA: each stage is a 32-pixel box, loaded as
// 32-channel x 32-pixel TMA boxes (4 KB: 32 K-rows of 128 bytes).  MN-major tf32 operands exist only
// in the "128B swizzle, 32-byte atom" layout (TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B <-> UMMA
// SWIZZLE_128B_BASE32B, Swizzle<2,5,2>: 4-row x 128-byte atoms): chunks of 32 channels sit 4 KB apart
// (descriptor LBO), groups of 4 K-rows 512 bytes apart (SBO); each tf32 MMA (K = 8 pixels) reads two.  One tap per CTA, split-K over pixel tiles with a fixed-order reduction afterwards.
// ================================================================================================
struct WgTcParams {
    float* partial;
    int a_pitch, g_pitch, n_mma, b_boxes;
    int ntaps, nsplits, tiles_total, tiles_per_split;
    int tiles_x, tiles_y, bw, bh, bb;
    int g_stride, stages;
    int dya[kMaxTaps], dxa[kMaxTaps], dyg[kMaxTaps], dxg[kMaxTaps];
};

constexpr int kWgPix = 32;           // pixels (GEMM K) per stage
constexpr int kBoxBytes = 32 * kWgPix * 4;   // 4 KB

__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)(kBoxBytes >> 4) << 16;   // LBO: next 32-channel chunk along M / N
    d |= (uint64_t)(512 >> 4) << 32;         // SBO: next group of 4 K-rows (one 32B-atom swizzle period)
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                  // SWIZZLE_128B_BASE32B
    return d;
}

__global__ void __launch_bounds__(kThreads, 2)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_g, const WgTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 1];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int S = p.stages;
    const uint32_t a_bytes = 4 * kBoxBytes, b_bytes = (uint32_t)p.b_boxes * kBoxBytes;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    auto a_addr = [&](int s) { return smem_base + (uint32_t)s * stage_bytes; };
    auto b_addr = [&](int s) { return smem_base + (uint32_t)s * stage_bytes + a_bytes; };
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
    const uint32_t acc_bar = bar0 + 8u * (2 * kMaxStages);

    const int i0 = blockIdx.x * kTileM;
    const int tap = blockIdx.y, split = blockIdx.z;
    const int t_lo = split * p.tiles_per_split;
    const int t_hi = min(p.tiles_total, t_lo + p.tiles_per_split);
    const int total = t_hi - t_lo;            // host guarantees >= 1

    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < p.n_mma) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base_slot;

    if (warp == 0) {
        const int dya = p.dya[tap], dxa = p.dxa[tap], dyg = p.dyg[tap], dxg = p.dxg[tap];
        const int a_boxes = min(4, (p.a_pitch - i0 + 31) / 32);
        for (int it = 0; it < total; ++it) {
            const int s = it % S;
            const uint32_t ph = (uint32_t)(it / S) & 1u;
            mbar_wait(empty_bar(s), ph ^ 1u);
            int t = t_lo + it;
            const int tx = t % p.tiles_x;
            t /= p.tiles_x;
            const int ty = t % p.tiles_y;
            const int tb = t / p.tiles_y;
            const int x0 = tx * p.bw, y0 = ty * p.bh, b0 = tb * p.bb;
            // only the 32-channel boxes that hold real input channels are loaded; accumulator rows fed
            // from the untouched shared memory are never stored
            if (elect_one()) {
                mbar_expect_tx(full_bar(s), (uint32_t)(a_boxes + p.b_boxes) * kBoxBytes);
                for (int c = 0; c < a_boxes; ++c)
                    tma_load_4d(a_addr(s) + c * kBoxBytes, &map_a, full_bar(s), i0 + 32 * c, x0 + dxa, y0 + dya, b0);
                for (int c = 0; c < p.b_boxes; ++c)
                    tma_load_4d(b_addr(s) + c * kBoxBytes, &map_g, full_bar(s), 32 * c, x0 * p.g_stride + dxg,
                                y0 * p.g_stride + dyg, b0);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // D=f32, A=B=tf32, both MN-major (bits 15, 16), N>>3, M>>4
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(p.n_mma >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
        // MN-major descriptor (make_desc_mn_sw128) split in words: LBO = 4 KB box pitch, SBO = 512, 32B-atom swizzle
        constexpr uint32_t kHi = (512u >> 4) | (1u << 14) | (1u << 29);
        constexpr uint32_t kLoFlags = (uint32_t)(kBoxBytes >> 4) << 16;
        for (int it = 0; it < total; ++it) {
            const int s = it % S;
            const uint32_t ph = (uint32_t)(it / S) & 1u;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint32_t a_lo = ((a_addr(s) >> 4) & 0x3fffu) | kLoFlags, b_lo = ((b_addr(s) >> 4) & 0x3fffu) | kLoFlags;
            if (elect_one()) {
                // 8 pixels (K) per MMA = 1024 bytes = +64 in the 16-byte address field
                tc_mma_tf32_lh(tmem_acc, a_lo, kHi, b_lo, kHi, idesc, it > 0 ? 1u : 0u);
                tc_mma_tf32_lh(tmem_acc, a_lo + 64, kHi, b_lo + 64, kHi, idesc, 1u);
                tc_mma_tf32_lh(tmem_acc, a_lo + 128, kHi, b_lo + 128, kHi, idesc, 1u);
                tc_mma_tf32_lh(tmem_acc, a_lo + 192, kHi, b_lo + 192, kHi, idesc, 1u);
                tc_commit(empty_bar(s));
                if (it == total - 1) tc_commit(acc_bar);
            }
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int ii = i0 + q * 32 + lane;     // accumulator row = input channel
        float* dst = p.partial + (((int64_t)split * p.ntaps + tap) * p.a_pitch + ii) * p.g_pitch;
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        for (int c = 0; c < p.n_mma; c += 16) {
            float v[16];
            tc_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            if (ii >= p.a_pitch) continue;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int o = c + g * 4;
                if (o < p.g_pitch) st4(dst + o, make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(tmem_cols) : "memory");
    }
}


// ================================================================================================
// Weight gradient, all taps of a 3x3 same-resolution convolution in ONE CTA: per 32-pixel stage the gradient tile
// g[32 px][N] is loaded once and the layer-input window a[3 rows][34 px][M] once; tap (dy, dx) is the SAME window
// seen through an MN-major descriptor shifted by (dy+1)*36 + (dx+1) K rows (absolute-address swizzle, see the
// probe), accumulating into its own TMEM tile.  Operand traffic per stage: 3*a_boxes + b_boxes TMA boxes instead
// of 9*(a_boxes + b_boxes) -- the one-tap-per-CTA kernel above is L2-bandwidth bound on exactly that re-read.
// One CTA per SM (taps_per * n_mma <= 512 TMEM columns), persistent over its split of the pixel tiles.
// ================================================================================================
constexpr int kWgRowPitch = 36;                                  // K rows reserved per image row (34 used): 9 x 512 B
constexpr int kWgABoxBytes = 3 * kWgRowPitch * 128;              // 13824: one 32-channel box of the 3-row window

struct WgAllParams {
    float* partial;
    int a_pitch, g_pitch, n_mma, a_boxes, b_boxes;
    int taps_per, nsplits, tiles_total, tiles_per_split;
    int tiles_x, tiles_y, stages;
    uint32_t stage_bytes;
};

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_alltaps_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_g,
                        const __grid_constant__ WgAllParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 1];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int S = p.stages;
    const uint32_t a_bytes = (uint32_t)p.a_boxes * kWgABoxBytes;
    auto a_addr = [&](int s) { return smem_base + (uint32_t)s * p.stage_bytes; };
    auto b_addr = [&](int s) { return smem_base + (uint32_t)s * p.stage_bytes + a_bytes; };
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
    const uint32_t acc_bar = bar0 + 8u * (2 * kMaxStages);

    const int i0 = blockIdx.x * kTileM;
    const int tap0 = blockIdx.y * p.taps_per;
    const int ntap = min(p.taps_per, 9 - tap0);
    const int split = blockIdx.z;
    const int t_lo = split * p.tiles_per_split;
    const int t_hi = min(p.tiles_total, t_lo + p.tiles_per_split);
    const int total = t_hi - t_lo;            // host guarantees >= 1

    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < p.taps_per * p.n_mma) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base_slot;

    if (warp == 0) {
        const int a_ld = min(p.a_boxes, (p.a_pitch - i0 + 31) / 32);      // boxes that hold real input channels
        const uint32_t tx_bytes = (uint32_t)a_ld * 3u * 34u * 128u + (uint32_t)p.b_boxes * kBoxBytes;
        for (int it = 0; it < total; ++it) {
            const int s = it % S;
            const uint32_t ph = (uint32_t)(it / S) & 1u;
            mbar_wait(empty_bar(s), ph ^ 1u);
            int t = t_lo + it;
            const int tx = t % p.tiles_x;
            t /= p.tiles_x;
            const int y0 = t % p.tiles_y;
            const int b0 = t / p.tiles_y;
            const int x0 = tx * kWgPix;
            if (elect_one()) {
                mbar_expect_tx(full_bar(s), tx_bytes);
                for (int c = 0; c < a_ld; ++c)
                    for (int r = 0; r < 3; ++r)
                        tma_load_4d(a_addr(s) + c * kWgABoxBytes + r * (kWgRowPitch * 128), &map_a, full_bar(s), i0 + 32 * c,
                                    x0 - 1, y0 + r - 1, b0);
                for (int c = 0; c < p.b_boxes; ++c)
                    tma_load_4d(b_addr(s) + c * kBoxBytes, &map_g, full_bar(s), 32 * c, x0, y0, b0);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(p.n_mma >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
        constexpr uint32_t kHi = (512u >> 4) | (1u << 14) | (1u << 29);
        constexpr uint32_t kLoA = (uint32_t)(kWgABoxBytes >> 4) << 16, kLoB = (uint32_t)(kBoxBytes >> 4) << 16;
        for (int it = 0; it < total; ++it) {
            const int s = it % S;
            const uint32_t ph = (uint32_t)(it / S) & 1u;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint32_t a_lo = ((a_addr(s) >> 4) & 0x3fffu) | kLoA, b_lo = ((b_addr(s) >> 4) & 0x3fffu) | kLoB;
            if (elect_one()) {
                for (int tl = 0; tl < ntap; ++tl) {
                    const int tap = tap0 + tl;
                    const int dy = tap / 3, dx = tap - dy * 3;                    // (dy+1, dx+1) of the conv tap
                    const uint32_t a_t = a_lo + (uint32_t)((dy * kWgRowPitch + dx) * 8);   // K rows are 128 B = 8 address units
                    const uint32_t d = tmem_acc + (uint32_t)(tl * p.n_mma);
                    const uint32_t acc = it > 0 ? 1u : 0u;
                    tc_mma_tf32_lh(d, a_t, kHi, b_lo, kHi, idesc, acc);
                    tc_mma_tf32_lh(d, a_t + 64, kHi, b_lo + 64, kHi, idesc, 1u);
                    tc_mma_tf32_lh(d, a_t + 128, kHi, b_lo + 128, kHi, idesc, 1u);
                    tc_mma_tf32_lh(d, a_t + 192, kHi, b_lo + 192, kHi, idesc, 1u);
                }
                tc_commit(empty_bar(s));
                if (it == total - 1) tc_commit(acc_bar);
            }
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int ii = i0 + q * 32 + lane;     // accumulator row = input channel
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        for (int tl = 0; tl < ntap; ++tl) {
            float* dst = p.partial + (((int64_t)split * 9 + tap0 + tl) * p.a_pitch + ii) * p.g_pitch;
            for (int c = 0; c < p.n_mma; c += 16) {
                float v[16];
                tc_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(tl * p.n_mma + c), v);
                if (ii >= p.a_pitch) continue;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int o = c + g * 4;
                    if (o < p.g_pitch) st4(dst + o, make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(tmem_cols) : "memory");
    }
}

// ================================================================================================
// All-taps weight gradient of the stride-2 TRANSPOSED 3x3 convolution (mode 1):
//     gw[(ky,kx)][i][o] = sum_{b,y,x} a[b,y,x,i] * g_T[b, 2y+ky, 2x+kx, o]
// Per 32-pixel stage the input tile a[32 px][M] is loaded once; of the gradient, for every needed row 2y+ky, the
// even-column samples E[j] = g_T[2(x0+j)] (33 of them) and the odd-column samples O[j] = g_T[2(x0+j)+1] (32) arrive
// through TMA boxes with element stride 2.  Tap (ky, kx) is then a row-shifted MN-major descriptor view of those
// rows: kx = 0 -> E, kx = 1 -> O, kx = 2 -> E shifted by one sample.  6 gradient boxes per channel box and stage
// instead of 9, and the input once instead of 9 times.
// ================================================================================================
constexpr int kWgUpKy = 2 * kWgRowPitch;                         // K rows reserved per gradient row: E (33) | O (32)

struct WgUpParams {
    float* partial;
    int a_pitch, g_pitch, n_mma, b_boxes;
    int n_ky, nsplits, tiles_total, tiles_per_split;             // n_ky gradient rows (= 3 taps each) per CTA
    int tiles_x, tiles_y, stages;
    uint32_t b_box_bytes, stage_bytes;
};

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_alltaps_up_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_ge,
                           const __grid_constant__ CUtensorMap map_go, const __grid_constant__ WgUpParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 1];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int S = p.stages;
    constexpr uint32_t a_bytes = 4 * kBoxBytes;                   // M = 128 always addresses four channel boxes
    auto a_addr = [&](int s) { return smem_base + (uint32_t)s * p.stage_bytes; };
    auto b_addr = [&](int s) { return smem_base + (uint32_t)s * p.stage_bytes + a_bytes; };
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
    const uint32_t acc_bar = bar0 + 8u * (2 * kMaxStages);

    const int i0 = blockIdx.x * kTileM;
    const int ky0 = blockIdx.y * p.n_ky;
    const int nky = min(p.n_ky, 3 - ky0);
    const int split = blockIdx.z;
    const int t_lo = split * p.tiles_per_split;
    const int t_hi = min(p.tiles_total, t_lo + p.tiles_per_split);
    const int total = t_hi - t_lo;            // host guarantees >= 1

    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 3 * p.n_ky * p.n_mma) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base_slot;

    if (warp == 0) {
        const int a_ld = min(4, (p.a_pitch - i0 + 31) / 32);
        const uint32_t tx_bytes = (uint32_t)a_ld * kBoxBytes + (uint32_t)(p.b_boxes * nky) * (33u + 32u) * 128u;
        for (int it = 0; it < total; ++it) {
            const int s = it % S;
            const uint32_t ph = (uint32_t)(it / S) & 1u;
            mbar_wait(empty_bar(s), ph ^ 1u);
            int t = t_lo + it;
            const int tx = t % p.tiles_x;
            t /= p.tiles_x;
            const int y0 = t % p.tiles_y;
            const int b0 = t / p.tiles_y;
            const int x0 = tx * kWgPix;
            if (elect_one()) {
                mbar_expect_tx(full_bar(s), tx_bytes);
                for (int c = 0; c < a_ld; ++c)
                    tma_load_4d(a_addr(s) + c * kBoxBytes, &map_a, full_bar(s), i0 + 32 * c, x0, y0, b0);
                for (int c = 0; c < p.b_boxes; ++c)
                    for (int r = 0; r < nky; ++r) {
                        const uint32_t dst = b_addr(s) + (uint32_t)c * p.b_box_bytes + (uint32_t)(r * kWgUpKy) * 128u;
                        tma_load_4d(dst, &map_ge, full_bar(s), 32 * c, 2 * x0, 2 * y0 + ky0 + r, b0);
                        tma_load_4d(dst + kWgRowPitch * 128, &map_go, full_bar(s), 32 * c, 2 * x0 + 1, 2 * y0 + ky0 + r, b0);
                    }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(p.n_mma >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
        constexpr uint32_t kHi = (512u >> 4) | (1u << 14) | (1u << 29);
        constexpr uint32_t kLoA = (uint32_t)(kBoxBytes >> 4) << 16;
        const uint32_t lo_b = (p.b_box_bytes >> 4) << 16;
        for (int it = 0; it < total; ++it) {
            const int s = it % S;
            const uint32_t ph = (uint32_t)(it / S) & 1u;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint32_t a_lo = ((a_addr(s) >> 4) & 0x3fffu) | kLoA, b_lo = ((b_addr(s) >> 4) & 0x3fffu) | lo_b;
            if (elect_one()) {
                for (int r = 0; r < nky; ++r)
                    for (int kx = 0; kx < 3; ++kx) {
                        // kx = 0: E rows 0..31, kx = 1: O rows (36 rows further), kx = 2: E rows 1..32
                        const int row = r * kWgUpKy + (kx == 1 ? kWgRowPitch : (kx == 2 ? 1 : 0));
                        const uint32_t b_t = b_lo + (uint32_t)(row * 8);
                        const uint32_t d = tmem_acc + (uint32_t)((r * 3 + kx) * p.n_mma);
                        tc_mma_tf32_lh(d, a_lo, kHi, b_t, kHi, idesc, it > 0 ? 1u : 0u);
                        tc_mma_tf32_lh(d, a_lo + 64, kHi, b_t + 64, kHi, idesc, 1u);
                        tc_mma_tf32_lh(d, a_lo + 128, kHi, b_t + 128, kHi, idesc, 1u);
                        tc_mma_tf32_lh(d, a_lo + 192, kHi, b_t + 192, kHi, idesc, 1u);
                    }
                tc_commit(empty_bar(s));
                if (it == total - 1) tc_commit(acc_bar);
            }
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int ii = i0 + q * 32 + lane;     // accumulator row = input channel
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        for (int tl = 0; tl < 3 * nky; ++tl) {
            float* dst = p.partial + (((int64_t)split * 9 + ky0 * 3 + tl) * p.a_pitch + ii) * p.g_pitch;
            for (int c = 0; c < p.n_mma; c += 16) {
                float v[16];
                tc_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(tl * p.n_mma + c), v);
                if (ii >= p.a_pitch) continue;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int o = c + g * 4;
                    if (o < p.g_pitch) st4(dst + o, make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(tmem_cols) : "memory");
    }
}

// ================================================================================================
// Streaming NHWC-p FIR: a CTA owns a (pixel-column strip x channel chunk x row segment) of one sample and
// marches down the rows.  A producer warp streams ONE input row per stage through a TMA/mbarrier ring
// (box = chunk channels x (strip + KW - 1) pixels, padding = out-of-bounds zero fill); the 256 consumer
// threads keep the last KH rows of their KW-pixel window in registers, so every input element is read from
// shared memory KW times and from L2/HBM once (plus the 3-pixel strip halo).  No vertical halo re-read
// inside a segment, loads run up to `stages` rows ahead of the math, stores are 128-bit and coalesced.
// Channel pitches that are not multiples of 32 (the pruned widths 154/77/39 -> 160/80/40) are covered by
// 32-channel chunks plus one narrower tail chunk with a proportionally longer pixel strip, so all 256
// threads stay busy there too.
// ================================================================================================
constexpr int kFirMaxStages = 8;

struct FirSP {
    const float* fir;
    const float* out_scale;
    const float* noise;
    const float* noise_w;
    const float* bias;
    float* out;
    const float* mask_ref;       // optional, layout of out: v *= mask_ref > 0 ? mask_gain : 0.2 * mask_gain (activation backward)
    float mask_gain;
    int out_h, out_w, pitch, valid, pad_x0, pad_y0, act;
    int64_t noise_bstride;
    int main_w, main_tx;         // channel width of a main chunk and its pixel strip (main_w * main_tx == PXT * 1024)
    int n_main_chunks, n_main;   // full main chunks; CTAs (x) that work on them
    int tail_w, tail_tx;         // width of the tail chunk (0: none) and its pixel strip
    int rows_per_seg, stages, stage_stride;
    float2 taps2[16];            // PTAPS: flipped taps, each duplicated (k, k) -- live in the constant bank / uniform registers
};

// two fp32 FMAs per issue slot (sm_100 FFMA2): a.xy += k.xy * v.xy
__device__ __forceinline__ void fma2(float2& a, const float2 k, const float2 v) {
    unsigned long long ua = *reinterpret_cast<unsigned long long*>(&a);
    const unsigned long long uk = *reinterpret_cast<const unsigned long long*>(&k);
    const unsigned long long uv = *reinterpret_cast<const unsigned long long*>(&v);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(ua) : "l"(uk), "l"(uv));
    a = *reinterpret_cast<float2*>(&ua);
}

// PXT: adjacent output pixels per thread (sliding window of PXT + KW - 1 shared-memory loads per row);
// PTAPS: taps passed by value in the parameter block (uniform registers) instead of being read from `fir`.
// 256 threads, all consumers; thread 0 also runs the TMA ring (it refills the stage of row r-1 with row
// r+S-1 before it waits for row r), so two CTAs fit the register file of an SM.
template <int KH, int KW, int PXT, bool PTAPS>
__global__ void __launch_bounds__(256, 2) fir_nhwc_stream_kernel(const __grid_constant__ CUtensorMap map_main,
                                                                 const __grid_constant__ CUtensorMap map_tail,
                                                                 const __grid_constant__ FirSP p) {
    static_assert(KH * KW <= 16 && KH == 4, "tap table holds 16 entries; the row loop is unrolled by KH == 4");
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kFirMaxStages];
    __shared__ float skf[KH * KW];
    const uint32_t base_u32 = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint8_t* base_ptr = smem_raw + (base_u32 - smem_u32(smem_raw));
    const int S = p.stages;
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kFirMaxStages + s); };

    const bool tail = (int)blockIdx.x >= p.n_main;
    int cw, tx_n, c0, x0;
    if (!tail) {
        cw = p.main_w; tx_n = p.main_tx;
        c0 = ((int)blockIdx.x % p.n_main_chunks) * cw;
        x0 = ((int)blockIdx.x / p.n_main_chunks) * tx_n;
    } else {
        cw = p.tail_w; tx_n = p.tail_tx;
        c0 = p.n_main_chunks * p.main_w;
        x0 = ((int)blockIdx.x - p.n_main) * tx_n;
    }
    const CUtensorMap* map = tail ? &map_tail : &map_main;
    const int y0 = blockIdx.y * p.rows_per_seg;
    const int rows_out = min(p.rows_per_seg, p.out_h - y0);
    const int rows_in = rows_out + KH - 1;
    const int b = blockIdx.z;
    const uint32_t stage_bytes = (uint32_t)(tx_n + KW - 1) * cw * 4;
    const int tx_c = x0 - p.pad_x0, ty_c = y0 - p.pad_y0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 8);      // one arrival per warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int pre = min(S - 1, rows_in);
        for (int r = 0; r < pre; ++r) {      // fill the ring: rows 0 .. S-2
            mbar_expect_tx(full_bar(r), stage_bytes);
            tma_load_4d(base_u32 + (uint32_t)r * p.stage_stride, map, full_bar(r), c0, tx_c, ty_c + r, b);
        }
    }
    if (!PTAPS && threadIdx.x < KH * KW) {
        const int ky = threadIdx.x / KW, kx = threadIdx.x % KW;
        skf[threadIdx.x] = p.fir[(KH - 1 - ky) * KW + (KW - 1 - kx)];
    }
    __syncthreads();

    float2 kf2[KH * KW];
#pragma unroll
    for (int i = 0; i < KH * KW; ++i) kf2[i] = PTAPS ? p.taps2[i] : make_float2(skf[i], skf[i]);
    const int q4 = cw >> 2;
    const int g = threadIdx.x / q4, c4 = threadIdx.x - g * q4;
    const int px0 = g * PXT;
    const bool active = px0 < tx_n;
    const int c = c0 + c4 * 4;
    const int lane = threadIdx.x & 31;
    float4 scale4 = make_float4(1.f, 1.f, 1.f, 1.f), bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float nw = 0.f;
    if (active) {
        if (p.out_scale) scale4 = ldg4(p.out_scale + (int64_t)b * p.pitch + c);
        if (p.bias) bias4 = ldg4(p.bias + c);
        if (p.noise) nw = __ldg(p.noise_w);
    }
    // epilogue in 3 instructions per element: lrelu(v)*sqrt2 == max(v', 0.2 v') with v' = sqrt2*v, and the gain is
    // folded into the demodulation scale, the bias and the noise weight once per thread
    if (p.act) {
        scale4.x *= kSqrt2; scale4.y *= kSqrt2; scale4.z *= kSqrt2; scale4.w *= kSqrt2;
        bias4.x *= kSqrt2; bias4.y *= kSqrt2; bias4.z *= kSqrt2; bias4.w *= kSqrt2;
        nw *= kSqrt2;
    }
    bool ok[PXT];
#pragma unroll
    for (int pp = 0; pp < PXT; ++pp) ok[pp] = active && (px0 + pp < tx_n) && (x0 + px0 + pp < p.out_w);
    const bool edge = c + 3 >= p.valid;
    const float* sbase = reinterpret_cast<const float*>(base_ptr) + (active ? (px0 * cw + c4 * 4) : 0);
    const int sstride = p.stage_stride >> 2;
    // running output / noise pointers of this thread's first pixel (advanced one image row per emitted row)
    float* optr = p.out + (((int64_t)b * p.out_h + y0) * p.out_w + (x0 + px0)) * p.pitch + c;
    const int64_t orow = (int64_t)p.out_w * p.pitch;
    const float* nptr = p.noise ? p.noise + (int64_t)b * p.noise_bstride + (int64_t)y0 * p.out_w + (x0 + px0) : nullptr;

    // acc[k][pp] = output row (k mod KH), pixel pp, in flight: input row r adds its vertical tap i = r - oy to
    // rows oy = r-KH+1 .. r (same fma order per output as a plain (i, j) double loop)
    float2 acc[KH][PXT][2];
#pragma unroll
    for (int i = 0; i < KH; ++i)
#pragma unroll
        for (int pp = 0; pp < PXT; ++pp) acc[i][pp][0] = acc[i][pp][1] = make_float2(0.f, 0.f);
    int s = 0, prev_s = 0;
    uint32_t ph = 0, prev_ph = 0;

    // register rings of the per-row side inputs: output row o (completed by input row r = o + KH - 1, Q = r mod KH) uses
    // noise slot Q and mask slot Q mod 2
    float nzr[4][PXT];
    float4 mrr[2][PXT];
#pragma unroll
    for (int pp = 0; pp < PXT; ++pp) {
        nzr[0][pp] = nzr[1][pp] = nzr[2][pp] = nzr[3][pp] = 0.f;
        mrr[0][pp] = mrr[1][pp] = make_float4(1.f, 1.f, 1.f, 1.f);
    }
    int nz_left = p.noise ? rows_out : 0, mr_left = p.mask_ref ? rows_out : 0;
    const float* mptr = p.mask_ref ? p.mask_ref + (optr - p.out) : nullptr;
    auto nz_req = [&](float (&dst)[PXT]) {
        if (nz_left > 0) {
#pragma unroll
            for (int pp = 0; pp < PXT; ++pp)
                if (ok[pp]) dst[pp] = __ldg(nptr + pp);
            nptr += p.out_w;
            --nz_left;
        }
    };
    auto mr_req = [&](float4 (&dst)[PXT]) {
        if (mr_left > 0) {
#pragma unroll
            for (int pp = 0; pp < PXT; ++pp)
                if (ok[pp]) dst[pp] = ldg4(mptr + (int64_t)pp * p.pitch);
            mptr += orow;
            --mr_left;
        }
    };
    nz_req(nzr[3]); nz_req(nzr[0]); nz_req(nzr[1]);       // output rows 0, 1, 2
    mr_req(mrr[1]);                                       // output row 0

    // one input row; Q = r mod KH (compile time), FIRST = the first KH rows of the segment (r == Q)
    auto row = [&](int r, auto qc, auto firstc) {
        constexpr int Q = decltype(qc)::value;
        constexpr bool FIRST = decltype(firstc)::value;
        if (threadIdx.x == 0 && r + S - 1 < rows_in) {
            const int st = FIRST && Q == 0 ? S - 1 : prev_s;
            if (!(FIRST && Q == 0)) mbar_wait(empty_bar(st), prev_ph);     // every warp is done with row r-1
            mbar_expect_tx(full_bar(st), stage_bytes);
            tma_load_4d(base_u32 + (uint32_t)st * p.stage_stride, map, full_bar(st), c0, tx_c, ty_c + r + S - 1, b);
        }
        // noise / activation-mask reference of LATER output rows: a row of this loop lasts about one DRAM latency, so
        // the loads are requested three rows (noise) / one row (mask) before the epilogue that consumes them
        constexpr bool EMIT = !FIRST || Q == KH - 1;
        if (EMIT) {
            nz_req(nzr[(Q + 3) % 4]);
            mr_req(mrr[(Q + 1) % 2]);
        }
        mbar_wait(full_bar(s), ph);
        const float* src = sbase + s * sstride;
        float4 win[PXT + KW - 1];
#pragma unroll
        for (int j = 0; j < PXT + KW - 1; ++j) win[j] = ld4(src + j * cw);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(s));
        prev_s = s; prev_ph = ph;
        if (++s == S) { s = 0; ph ^= 1u; }
#pragma unroll
        for (int i = 0; i < KH; ++i) {
            if (FIRST && Q - i < 0) continue;           // output row above this segment
#pragma unroll
            for (int pp = 0; pp < PXT; ++pp)
#pragma unroll
                for (int j = 0; j < KW; ++j) {
                    const float4 v = win[pp + j];
                    fma2(acc[(Q - i + KH) % KH][pp][0], kf2[i * KW + j], make_float2(v.x, v.y));
                    fma2(acc[(Q - i + KH) % KH][pp][1], kf2[i * KW + j], make_float2(v.z, v.w));
                }
        }
        if (!FIRST || Q == KH - 1) {
#pragma unroll
            for (int pp = 0; pp < PXT; ++pp) {
                const float2 a0 = acc[(Q + 1) % KH][pp][0], a1 = acc[(Q + 1) % KH][pp][1];
                acc[(Q + 1) % KH][pp][0] = acc[(Q + 1) % KH][pp][1] = make_float2(0.f, 0.f);
                const float nzv = nzr[Q][pp];
                float v[4] = {fmaf(a0.x, scale4.x, fmaf(nzv, nw, bias4.x)), fmaf(a0.y, scale4.y, fmaf(nzv, nw, bias4.y)),
                              fmaf(a1.x, scale4.z, fmaf(nzv, nw, bias4.z)), fmaf(a1.y, scale4.w, fmaf(nzv, nw, bias4.w))};
                if (p.act) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], kLreluSlope * v[j]);
                }
                if (p.mask_ref) {
                    const float4 m = mrr[Q % 2][pp];
                    v[0] *= m.x > 0.f ? p.mask_gain : p.mask_gain * kLreluSlope;
                    v[1] *= m.y > 0.f ? p.mask_gain : p.mask_gain * kLreluSlope;
                    v[2] *= m.z > 0.f ? p.mask_gain : p.mask_gain * kLreluSlope;
                    v[3] *= m.w > 0.f ? p.mask_gain : p.mask_gain * kLreluSlope;
                }
                if (edge) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (c + j >= p.valid) v[j] = 0.f;
                }
                if (ok[pp]) st4(optr + (int64_t)pp * p.pitch, make_float4(v[0], v[1], v[2], v[3]));
            }
            optr += orow;
        }
    };
    using std::integral_constant;
    row(0, integral_constant<int, 0>{}, integral_constant<bool, true>{});      // rows_in >= KH always
    row(1, integral_constant<int, 1>{}, integral_constant<bool, true>{});
    row(2, integral_constant<int, 2>{}, integral_constant<bool, true>{});
    row(3, integral_constant<int, 3>{}, integral_constant<bool, true>{});
    for (int r0 = KH; r0 < rows_in; r0 += KH) {
        row(r0, integral_constant<int, 0>{}, integral_constant<bool, false>{});
        if (r0 + 1 >= rows_in) break;
        row(r0 + 1, integral_constant<int, 1>{}, integral_constant<bool, false>{});
        if (r0 + 2 >= rows_in) break;
        row(r0 + 2, integral_constant<int, 2>{}, integral_constant<bool, false>{});
        if (r0 + 3 >= rows_in) break;
        row(r0 + 3, integral_constant<int, 3>{}, integral_constant<bool, false>{});
    }
}

// ================================================================================================
// Register-streaming NHWC-p FIR without shared memory: one thread per float4 of the flattened (x, c) output
// row, marching down a row segment.  The KW-pixel window of the next input row is loaded (128-bit,
// read-only path; the KW-fold horizontal overlap is served by L1) while the current row is accumulated
// into the KH output rows it contributes to.  A warp touches 512 contiguous bytes per load.
// ================================================================================================
struct FirLP {
    const float* in;
    const float* fir;
    const float* out_scale;
    const float* noise;
    const float* noise_w;
    const float* bias;
    float* out;
    int in_h, in_w, out_h, out_w, pitch, valid, pad_x0, pad_y0, act, rows_per_seg;
    int64_t noise_bstride;
};

template <int KH, int KW>
__global__ void __launch_bounds__(256, 2) fir_nhwc_ldg_kernel(const FirLP p) {
    __shared__ float skf[KH * KW];
    if (threadIdx.x < KH * KW) {
        const int ky = threadIdx.x / KW, kx = threadIdx.x % KW;
        skf[threadIdx.x] = p.fir[(KH - 1 - ky) * KW + (KW - 1 - kx)];
    }
    __syncthreads();
    const int c4n = p.pitch >> 2;
    const int pos = blockIdx.x * 256 + threadIdx.x;
    if (pos >= p.out_w * c4n) return;
    const int ox = pos / c4n, c = (pos - ox * c4n) * 4;
    const int b = blockIdx.z;
    const int y0 = blockIdx.y * p.rows_per_seg;
    const int rows_out = min(p.rows_per_seg, p.out_h - y0);
    const int rows_in = rows_out + KH - 1;
    float kf[KH * KW];
#pragma unroll
    for (int i = 0; i < KH * KW; ++i) kf[i] = skf[i];
    float4 scale4 = make_float4(1.f, 1.f, 1.f, 1.f), bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.out_scale) scale4 = ldg4(p.out_scale + (int64_t)b * p.pitch + c);
    if (p.bias) bias4 = ldg4(p.bias + c);
    const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
    const int ix0 = ox - p.pad_x0;
    const float* src = p.in + (int64_t)b * p.in_h * p.in_w * p.pitch + (int64_t)ix0 * p.pitch + c;
    bool xok[KW];
#pragma unroll
    for (int j = 0; j < KW; ++j) xok[j] = (ix0 + j >= 0) && (ix0 + j < p.in_w);
    auto load_row = [&](int r, float4* w) {
        const int iy = y0 - p.pad_y0 + r;
        const bool rowok = (r < rows_in) && iy >= 0 && iy < p.in_h;
        const float* rp = src + (int64_t)iy * p.in_w * p.pitch;
#pragma unroll
        for (int j = 0; j < KW; ++j)
            w[j] = (rowok && xok[j]) ? ldg4(rp + (int64_t)j * p.pitch) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float4 acc4[KH];
#pragma unroll
    for (int i = 0; i < KH; ++i) acc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 nxt[KW];
    load_row(0, nxt);
    for (int r0 = 0; r0 < rows_in; r0 += KH) {
#pragma unroll
        for (int q = 0; q < KH; ++q) {
            const int r = r0 + q;
            if (r >= rows_in) break;
            float4 win[KW];
#pragma unroll
            for (int j = 0; j < KW; ++j) win[j] = nxt[j];
            load_row(r + 1, nxt);
#pragma unroll
            for (int i = 0; i < KH; ++i) {
                if (r - i < 0) continue;
                float4& a = acc4[(q - i + KH) % KH];
#pragma unroll
                for (int j = 0; j < KW; ++j) {
                    const float k = kf[i * KW + j];
                    a.x = fmaf(k, win[j].x, a.x);
                    a.y = fmaf(k, win[j].y, a.y);
                    a.z = fmaf(k, win[j].z, a.z);
                    a.w = fmaf(k, win[j].w, a.w);
                }
            }
            if (r >= KH - 1) {
                const int oy = y0 + r - (KH - 1);
                const float4 acc = acc4[(q + 1) % KH];
                acc4[(q + 1) % KH] = make_float4(0.f, 0.f, 0.f, 0.f);
                float v[4] = {acc.x * scale4.x, acc.y * scale4.y, acc.z * scale4.z, acc.w * scale4.w};
                if (p.noise) {
                    const float nz = nw * __ldg(p.noise + (int64_t)b * p.noise_bstride + (int64_t)oy * p.out_w + ox);
                    v[0] += nz; v[1] += nz; v[2] += nz; v[3] += nz;
                }
                v[0] += bias4.x; v[1] += bias4.y; v[2] += bias4.z; v[3] += bias4.w;
                if (p.act) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = lrelu_sqrt2(v[j]);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c + j >= p.valid) v[j] = 0.f;
                st4(p.out + (((int64_t)b * p.out_h + oy) * p.out_w + ox) * p.pitch + c, make_float4(v[0], v[1], v[2], v[3]));
            }
        }
    }
}

static int encode_act_map(EncodeTiledFn encode, CUtensorMap* map, const float* ptr, int B, int H, int W, int pitch,
                          int bw, int bh, int bb, int stride) {
    cuuint64_t dims[4] = {(cuuint64_t)pitch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)pitch * 4, (cuuint64_t)W * pitch * 4, (cuuint64_t)H * W * pitch * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)(bw * stride), (cuuint32_t)(bh * stride), (cuuint32_t)bb};
    cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    return (int)encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// Split-K second pass: out = epilogue(sum_z ws[z]) -- slabs added in the fixed order z = 0 .. nsplit-1 (deterministic).
struct SplitKEpiP {
    const float* ws;
    int64_t slab;              // floats per slab = B * H * W * n_pitch
    int nsplit;
    const float* out_scale;
    const float* noise;
    const float* noise_w;
    const float* bias;
    const float* residual;
    float* out;
    int HW, n_pitch, out_valid, act;
    int64_t noise_bstride;
    float act_gain;
    int W;                     // phased form (W > 0): pixel (y, x) was produced by ns_par[(y & 1) * 2 + (x & 1)] splits
    int ns_par[4];
};

__global__ void __launch_bounds__(256) splitk_epilogue_kernel(const __grid_constant__ SplitKEpiP p) {
    const int c4n = p.n_pitch >> 2;
    const int64_t n4 = p.slab >> 2;
    const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
        const int64_t pix = i / c4n;
        const int n = (int)(i - pix * c4n) * 4;
        const int b = (int)(pix / p.HW);
        int ns = p.nsplit;
        if (p.W > 0) {
            const int r = (int)(pix - (int64_t)b * p.HW);
            const int y = r / p.W, x = r - y * p.W;
            ns = p.ns_par[((y & 1) << 1) | (x & 1)];
        }
        float4 a = ld4(p.ws + i * 4);
        for (int z = 1; z < ns; ++z) {
            const float4 v = ld4(p.ws + (int64_t)z * p.slab + i * 4);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        float o[4] = {a.x, a.y, a.z, a.w};
        if (p.out_scale) {
            const float4 s4 = ldg4(p.out_scale + (int64_t)b * p.n_pitch + n);
            o[0] *= s4.x; o[1] *= s4.y; o[2] *= s4.z; o[3] *= s4.w;
        }
        if (p.noise) {
            const float nz = nw * __ldg(p.noise + (int64_t)b * p.noise_bstride + (pix - (int64_t)b * p.HW));
            o[0] += nz; o[1] += nz; o[2] += nz; o[3] += nz;
        }
        if (p.bias) {
            const float4 b4 = ldg4(p.bias + n);
            o[0] += b4.x; o[1] += b4.y; o[2] += b4.z; o[3] += b4.w;
        }
        if (p.act == kActLrelu) {
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = lrelu_gain(o[j], p.act_gain);
        }
        if (p.residual) {
            const float4 r4 = ldg4(p.residual + i * 4);
            if (p.act == kActMaskRef) {
                o[0] = r4.x > 0.f ? o[0] : 0.f; o[1] = r4.y > 0.f ? o[1] : 0.f;
                o[2] = r4.z > 0.f ? o[2] : 0.f; o[3] = r4.w > 0.f ? o[3] : 0.f;
            } else {
                o[0] += r4.x; o[1] += r4.y; o[2] += r4.z; o[3] += r4.w;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (n + j >= p.out_valid) o[j] = 0.f;
        st4(p.out + i * 4, make_float4(o[0], o[1], o[2], o[3]));
    }
}

static int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace tc
}  // namespace cagc

using namespace cagc;

// Halo-tile launch (see conv_tc_halo_kernel).  Returns 1 when it took the call (result in *rc).
static int try_halo_conv(cudaStream_t stream, const ConvP& c, const char* what, cagc::tc::EncodeTiledFn encode, int* rc,
                         bool dry_run = false) {
    using namespace cagc::tc;
    static const int halo_env = [] { const char* e = getenv("CAGC_TC_HALO"); return e ? atoi(e) : 1; }();
    if (!halo_env || c.in_stride != 1 || c.ntaps < 2 || c.B > 65535) return 0;
    // default: the narrow (operand-traffic-bound) layers; CAGC_TC_HALO=2 forces every eligible shape
    // (measured per layer, scripts/layer_times.py: wins for N <= 80 -- 39/77-channel layers and their up-conv phases;
    //  at N >= 128 the plain kernels' larger MMAs and two-CTA overlap are faster)
    if (halo_env == 1 && c.n_cols > 80 && !(c.n_cols <= 256 && c.Wo >= 128 && c.Wo % 128 == 0)) return 0;
    const int min_w = [] { const char* e = getenv("CAGC_TC_HALO_MINW"); return e ? atoi(e) : 32; }();
    if (c.Wo < min_w) return 0;
    int dxmin = 0, dxmax = 0, dymin = 0, dymax = 0;
    for (int i = 0; i < c.ntaps; ++i) {
        dxmin = std::min(dxmin, c.taps[i].dx); dxmax = std::max(dxmax, c.taps[i].dx);
        dymin = std::min(dymin, c.taps[i].dy); dymax = std::max(dymax, c.taps[i].dy);
    }
    const int dxs = dxmax - dxmin, dys = dymax - dymin;
    HaloParams p{};
    p.n_rows = (c.n_cols + 15) & ~15;
    p.n_tile = std::min(256, p.n_rows);
    const int n_tiles = ceil_div(p.n_rows, p.n_tile);
    const int mt_max = std::min(8, 512 / p.n_tile);
    p.b_bytes = (uint32_t)p.n_tile * kChunkK * 4;
    const int epi_bytes = (p.n_tile >= 128) ? kEpiStageBytes : 0;      // staging tiles of the transposing epilogue
    const int smem_budget = 222 * 1024 - epi_bytes;
    // the weight ring must cover the TMA round trip of its per-tap stages: 8 small stages for narrow N, at least 4
    // of the large (>= 16 KB) stages of wide N
    const int b_min = (p.b_bytes >= 32768) ? 3 : (p.b_bytes >= 16384) ? 4 : std::min(kHaloMaxB, c.ntaps);
    // Row mode (wide layers, W >= 128): a tile is R full 128-pixel output rows; sub-tile j is output row j, its taps
    // are the R+dys input rows of the window shifted by dx -- no halo column ever enters the M dimension, weights are
    // shared by the R rows, every input row is loaded once per 32-channel chunk instead of once per tap.
    p.row_mode = (c.n_cols > 80 && c.Wo >= 128 && c.Wo % 128 == 0) ? 1 : 0;
    if (const char* e = getenv("CAGC_TC_HALO_ROWMODE")) p.row_mode = atoi(e) && c.Wo >= 128;
    double best = 0;
    int bestR = 0;
    if (p.row_mode) {
        p.TW = 128;
        p.Wt = p.TW + dxs;
        p.strips = ceil_div(c.Wo, p.TW);
        for (int R = 1; R <= std::min(c.Ho, mt_max); ++R) {
            const int rows = (R + dys) * p.Wt;
            const uint32_t a_stride = ((uint32_t)rows * 128u + 1023u) & ~1023u;
            if (2 * (int64_t)a_stride + (int64_t)b_min * p.b_bytes + 1024 > smem_budget) break;
            const bool dbl = 2 * R * p.n_tile <= 512;
            // operand bytes per output row (A rows + the weights of 9 taps shared by R rows), single-buffered
            // accumulators pay an un-overlapped epilogue
            const double bytes = ((double)(R + dys) * p.Wt * 128.0 + (double)c.ntaps * p.b_bytes) / R;
            const double score = 1.0 / (bytes * (dbl ? 1.0 : 1.12));
            if (score > best) { best = score; bestR = R; }
        }
        if (const char* e = getenv("CAGC_TC_HALO_R")) bestR = std::min(bestR ? mt_max : 0, atoi(e));
        if (bestR <= 0) return 0;
        p.R = bestR;
        p.mt = p.R;
        p.sub_rows = p.Wt;
    } else {
        p.TW = std::min(c.Wo, 64);
        if (const char* e = getenv("CAGC_TC_HALO_TW")) p.TW = std::min(c.Wo, std::max(8, atoi(e)));
        p.Wt = p.TW + dxs;
        p.strips = ceil_div(c.Wo, p.TW);
        // choose the tile height: best MMA-row / operand-byte efficiency that fits shared memory and TMEM
        for (int R = 1; R <= std::min(c.Ho, 32); ++R) {
            const int mt = ceil_div(R * p.Wt, kTileM);
            if (mt > mt_max) break;
            const int rows = std::max((R + dys) * p.Wt, mt * kTileM + dys * p.Wt + dxs);
            const uint32_t a_stride = ((uint32_t)rows * 128u + 1023u) & ~1023u;
            if (2 * (int64_t)a_stride + (int64_t)b_min * p.b_bytes + 1024 > smem_budget) break;
            const int ytiles = ceil_div(c.Ho, R);
            // cost per useful output pixel: MMA rows + (weighted) operand rows, including the ragged last row tile
            const double useful = (double)c.Ho * c.Wo;
            const double mma = (double)ytiles * p.strips * mt * kTileM;
            const double load = (double)ytiles * p.strips * (R + dys) * p.Wt;
            const bool dbl = 2 * mt * p.n_tile <= 512;
            const double score = useful / (mma * (dbl ? 1.0 : 1.15) + 0.5 * load);
            if (score > best) { best = score; bestR = R; }
        }
        if (const char* e = getenv("CAGC_TC_HALO_R")) bestR = std::min(bestR ? 32 : 0, atoi(e));
        if (bestR <= 0) return 0;
        p.R = bestR;
        p.mt = ceil_div(p.R * p.Wt, kTileM);
        p.sub_rows = kTileM;
    }
    if (p.mt > mt_max) return 0;
    p.rows_box = p.R + dys;
    p.ytiles = ceil_div(c.Ho, p.R);
    p.acc_bufs = (2 * p.mt * p.n_tile <= 512) ? 2 : 1;
    p.a_bytes = (uint32_t)p.rows_box * p.Wt * 128u;
    const int rows = p.row_mode ? p.rows_box * p.Wt : std::max(p.rows_box * p.Wt, p.mt * kTileM + dys * p.Wt + dxs);
    p.a_stride = ((uint32_t)rows * 128u + 1023u) & ~1023u;
    const int64_t left = smem_budget - 1024 - 2 * (int64_t)p.a_stride;
    p.b_stages = (int)std::min<int64_t>(kHaloMaxB, left / p.b_bytes);
    if (p.b_stages < 2) return 0;
    if (dry_run) return 1;          // eligibility only (cagc_tc_conv_takes_sample_weights)
    p.w_bslabs = c.w_bslabs;
    p.out_scale = c.out_scale; p.noise = c.noise; p.noise_w = c.noise_w; p.bias = c.bias; p.out = c.out;
    p.residual = c.residual; p.act_gain = (c.act_gain != 0.f) ? c.act_gain : kSqrt2;
    p.B = c.B; p.Ho = c.Ho; p.Wo = c.Wo;
    p.dx_min = dxmin; p.dy_min = dymin;
    p.k_valid = c.in_pitch; p.n_pitch = c.n_cols; p.out_valid = c.out_valid;
    p.Hout = c.Hout; p.Wout = c.Wout; p.out_stride = c.out_stride; p.out_oy = c.out_oy; p.out_ox = c.out_ox;
    p.noise_bstride = c.noise_bstride; p.act = c.act; p.ntaps = c.ntaps;
    int max_slab = 0;
    for (int i = 0; i < c.ntaps; ++i) {
        p.tap_row[i] = (c.taps[i].dy - dymin) * p.Wt + (c.taps[i].dx - dxmin);
        p.tap_slab[i] = c.taps[i].slab;
        max_slab = std::max(max_slab, c.taps[i].slab);
    }
    CUtensorMap map_a, map_b;
    {
        cuuint64_t dims[4] = {(cuuint64_t)c.in_pitch, (cuuint64_t)c.Win, (cuuint64_t)c.Hin, (cuuint64_t)c.B};
        cuuint64_t strides[3] = {(cuuint64_t)c.in_pitch * 4, (cuuint64_t)c.Win * c.in_pitch * 4,
                                 (cuuint64_t)c.Hin * c.Win * c.in_pitch * 4};
        cuuint32_t box[4] = {(cuuint32_t)kChunkK, (cuuint32_t)p.Wt, 1, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        if (p.Wt > 256) return 0;
        CUresult r = encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(c.in), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { *rc = fail(CAGC_E_INVALID, "%s: cuTensorMapEncodeTiled(A, halo) failed with %d", what, (int)r); return 1; }
    }
    {
        const int nslabs = c.w_bslabs > 0 ? c.w_bslabs * c.B : max_slab + 1;
        cuuint64_t dims[3] = {(cuuint64_t)c.in_pitch, (cuuint64_t)p.n_rows, (cuuint64_t)nslabs};
        cuuint64_t strides[2] = {(cuuint64_t)c.in_pitch * 4, (cuuint64_t)p.n_rows * c.in_pitch * 4};
        cuuint32_t box[3] = {(cuuint32_t)kChunkK, (cuuint32_t)p.n_tile, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(c.w), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { *rc = fail(CAGC_E_INVALID, "%s: cuTensorMapEncodeTiled(B, halo) failed with %d", what, (int)r); return 1; }
    }
    static DeviceOnce attr_set{0};
    if (device_once_needed(attr_set)) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024);
        if (e != cudaSuccess) { *rc = fail((int)e, "%s: cudaFuncSetAttribute failed: %s", what, cudaGetErrorString(e)); return 1; }
        device_once_done(attr_set);
    }
    p.epi_off = 2u * p.a_stride + (uint32_t)p.b_stages * p.b_bytes;
    const size_t smem = 2 * (size_t)p.a_stride + (size_t)p.b_stages * p.b_bytes + 1024 + epi_bytes;
    const int64_t items = (int64_t)p.strips * p.ytiles * c.B * n_tiles;
    const unsigned grid = (unsigned)std::min<int64_t>(items, kNumSMs);
    conv_tc_halo_kernel<<<grid, kThreadsPersist, smem, stream>>>(map_a, map_b, p);
    *rc = launched(what);
    return 1;
}

int cagc_tc_conv(cudaStream_t stream, const ConvP& c, const char* what) {
    using namespace cagc::tc;
    CAGC_REQUIRE(c.in_scale == nullptr, "%s: the tcgen05 path takes a pre-modulated input (cagc_modulate)", what);
    CAGC_REQUIRE(c.in_stride == 1 || c.in_stride == 2, "%s: input stride must be 1 or 2", what);
    CAGC_REQUIRE(c.in_pitch % 8 == 0 && c.n_cols % 8 == 0, "%s: tcgen05 path needs channel pitches that are multiples of 8",
                 what);
    EncodeTiledFn encode = get_encode();
    if (!encode) return fail(CAGC_E_UNSUPPORTED, "%s: cuTensorMapEncodeTiled not available from the driver", what);
    if ((int64_t)c.B * c.Ho * c.Wo == 0) return 0;

    {
        int rc = 0;
        if (try_halo_conv(stream, c, what, encode, &rc)) return rc;
        if (c.w_bslabs > 0)
            return fail(CAGC_E_UNSUPPORTED, "%s: per-sample weight slabs need the halo-tile kernel, which does not take this shape", what);
    }
    TcParams p{};
    p.out_scale = c.out_scale; p.noise = c.noise; p.noise_w = c.noise_w; p.bias = c.bias; p.out = c.out;
    p.residual = c.residual; p.act_gain = (c.act_gain != 0.f) ? c.act_gain : kSqrt2;
    p.B = c.B; p.Ho = c.Ho; p.Wo = c.Wo;
    p.bw = std::min(16, next_pow2(c.Wo));
    p.bh = std::min(kTileM / p.bw, next_pow2(c.Ho));
    p.bb = kTileM / (p.bw * p.bh);
    p.tiles_x = ceil_div(c.Wo, p.bw);
    p.tiles_y = ceil_div(c.Ho, p.bh);
    const int tiles_b = ceil_div(c.B, p.bb);
    p.k_valid = c.in_pitch;
    p.n_pitch = c.n_cols; p.out_valid = c.out_valid;
    p.n_rows = (c.n_cols + 15) & ~15;
    p.n_tile = std::min(256, p.n_rows);
    p.tiles_total = p.tiles_x * p.tiles_y * tiles_b;
    // Tile-shape heuristic (measured per layer on B200, scripts/layer_times.py):
    //  * plenty of pixel tiles, N tile <= 128: two 128-pixel sub-tiles per CTA share each weight tile (the kernel
    //    is bound by L2->SMEM operand traffic there; +20..27%);
    //  * plenty of pixel tiles, N tile > 128: one sub-tile, two CTAs per SM (epilogue overlap matters more);
    //  * few pixel tiles (4x4 .. 16x16 layers): narrower N tiles -> more CTAs and a deeper TMA pipeline per CTA
    //    (those layers are bound by the serial K loop's load latency).
    p.mt = 1;
    if ((int64_t)p.tiles_total * ceil_div(p.n_rows, p.n_tile) >= 4 * kNumSMs) {
        if (p.n_tile <= 128) p.mt = 2;
    } else {
        // (with a split-K workspace the K loop is dealt to up to 8 CTAs per tile: keep wider N tiles, fewer A re-reads)
        const int want_ctas = (c.workspace && c.out_stride == 1) ? kNumSMs / 8 : kNumSMs;
        while (p.n_tile > 32 && (int64_t)p.tiles_total * ceil_div(p.n_rows, p.n_tile) < want_ctas) {
            int nt = (p.n_tile / 2 + 15) & ~15;
            if (nt < 32) nt = 32;
            p.n_tile = nt;
        }
    }
    p.Hout = c.Hout; p.Wout = c.Wout; p.out_stride = c.out_stride; p.out_oy = c.out_oy; p.out_ox = c.out_ox;
    p.noise_bstride = c.noise_bstride; p.act = c.act; p.ntaps = c.ntaps; p.in_stride = c.in_stride;
    for (int i = 0; i < c.ntaps; ++i) p.taps[i] = c.taps[i];
    p.b_bytes = (uint32_t)p.n_tile * kChunkK * 4;
    if (const char* e = getenv("CAGC_TC_MT")) p.mt = (atoi(e) == 2 && p.tiles_total >= 2) ? 2 : 1;   // tuning knob
    if (const char* e = getenv("CAGC_TC_NTILE")) p.n_tile = std::min(p.n_tile, std::max(32, atoi(e) & ~15));
    p.b_bytes = (uint32_t)p.n_tile * kChunkK * 4;
    const uint32_t stage_bytes = (uint32_t)p.mt * kABytes + p.b_bytes;
    int tmem_need = 32;
    while (tmem_need < p.mt * p.n_tile) tmem_need <<= 1;
    // two CTAs per SM (110 KB each) when two stages fit; the transposing epilogue's staging tiles are dropped there if
    // they do not fit beside the ring (n_tile = 256: 2 x 48 KB stages) -- the second CTA hides the epilogue instead
    const bool two_ctas = tmem_need <= 256 && 2 * stage_bytes + 1024 <= (uint32_t)kSmemBudget;
    const bool stage_fits = !two_ctas || 2 * stage_bytes + 1024 + kEpiStageBytes <= (uint32_t)kSmemBudget;
    int budget = two_ctas ? (stage_fits ? kSmemBudget - kEpiStageBytes : kSmemBudget) : kConvBudget1;
    // grids smaller than the machine (4x4 .. 32x32 layers): one CTA per SM anyway, and the launch is bound by the
    // serial K loop's TMA latency -- give each CTA the whole shared memory for a deeper ring
    if ((int64_t)ceil_div(p.tiles_total, p.mt) * ceil_div(p.n_rows, p.n_tile) <= kNumSMs) budget = kConvBudget1;
    p.stages = std::max(2, std::min(kMaxStages, (int)((budget - 1024) / stage_bytes)));
    bool with_stage = stage_fits;
    if ((int64_t)ceil_div(p.tiles_total, p.mt) * ceil_div(p.n_rows, p.n_tile) <= kNumSMs) with_stage = true;   // one CTA per SM
    p.epi_off = with_stage ? (uint32_t)p.stages * stage_bytes : 0xffffffffu;
    const size_t smem = (size_t)p.stages * stage_bytes + 1024 + (with_stage ? kEpiStageBytes : 0);

    // A: activations [B, Hin, Win, in_pitch] fp32, box (32 ch, bw, bh, bb), 128-byte swizzle, OOB -> 0
    CUtensorMap map_a, map_b;
    {
        cuuint64_t dims[4] = {(cuuint64_t)c.in_pitch, (cuuint64_t)c.Win, (cuuint64_t)c.Hin, (cuuint64_t)c.B};
        cuuint64_t strides[3] = {(cuuint64_t)c.in_pitch * 4, (cuuint64_t)c.Win * c.in_pitch * 4,
                                 (cuuint64_t)c.Hin * c.Win * c.in_pitch * 4};
        // strided gather (data gradient of the transposed conv): TMA traverses the box with element
        // stride 2 and delivers ceil(box / stride) elements per dimension
        const cuuint32_t st = (cuuint32_t)c.in_stride;
        cuuint32_t box[4] = {(cuuint32_t)kChunkK, (cuuint32_t)p.bw * st, (cuuint32_t)p.bh * st, (cuuint32_t)p.bb};
        cuuint32_t es[4] = {1, st, st, 1};
        CUresult r = encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(c.in), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(CAGC_E_INVALID, "%s: cuTensorMapEncodeTiled(A) failed with %d", what, (int)r);
    }
    // B: weight slabs [taps][n_rows][in_pitch] fp32 (K contiguous), box (32 k, n_tile, 1)
    {
        cuuint64_t dims[3] = {(cuuint64_t)c.in_pitch, (cuuint64_t)p.n_rows, (cuuint64_t)kMaxTaps};
        cuuint64_t strides[2] = {(cuuint64_t)c.in_pitch * 4, (cuuint64_t)p.n_rows * c.in_pitch * 4};
        cuuint32_t box[3] = {(cuuint32_t)kChunkK, (cuuint32_t)p.n_tile, 1};
        cuuint32_t es[3] = {1, 1, 1};
        int max_slab = 0;
        for (int i = 0; i < c.ntaps; ++i) max_slab = std::max(max_slab, c.taps[i].slab);
        dims[2] = (cuuint64_t)(max_slab + 1);
        CUresult r = encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(c.w), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(CAGC_E_INVALID, "%s: cuTensorMapEncodeTiled(B) failed with %d", what, (int)r);
    }
    static DeviceOnce attr_set{0};
    if (device_once_needed(attr_set)) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget1);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(conv_tc_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget1);
        if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute failed: %s", what, cudaGetErrorString(e));
        device_once_done(attr_set);
    }
    // persistent form: whenever a double-buffered accumulator fits in TMEM (MT * n_tile <= 256 columns)
    static const int persist_env = [] { const char* e = getenv("CAGC_TC_PERSIST"); return e ? atoi(e) : 1; }();
    // (measured: +9% on the large teacher layers; with fewer than ~6 work items per SM the static
    // round-robin schedule loses more to wave imbalance than the overlap gains)
    const int64_t items = ceil_div<int64_t>(p.tiles_total, p.mt) * ceil_div(p.n_rows, p.n_tile);
    if (persist_env && p.mt * p.n_tile <= 256 && (items >= 6 * kNumSMs || persist_env == 2)) {
        TcParams q = p;
        q.nphase = 1;
        q.ph_item0[0] = 0; q.ph_item0[1] = (int)items;
        q.ph_tap0[0] = 0; q.ph_ntaps[0] = p.ntaps;
        q.ph_Ho[0] = p.Ho; q.ph_Wo[0] = p.Wo; q.ph_oy[0] = p.out_oy; q.ph_ox[0] = p.out_ox;
        q.ph_tx[0] = p.tiles_x; q.ph_ty[0] = p.tiles_y; q.ph_tiles[0] = p.tiles_total;
        const uint32_t sb = (uint32_t)q.mt * kABytes + q.b_bytes;
        const int ew = persist_epi_warps(c.ntaps * ceil_div(c.in_pitch, kChunkK));
        const int epi_bytes = ew * 2048;            // 2 KB (32 rows x 16 columns) of staging per epilogue warp
        q.stages = std::max(2, std::min(kMaxStages, (int)((kSmemBudget1 - epi_bytes - 1024) / sb)));
        q.epi_off = (uint32_t)q.stages * sb;
        const size_t smem_p = (size_t)q.stages * sb + 1024 + epi_bytes;
        const unsigned gridp = (unsigned)std::min<int64_t>(items, kNumSMs);
        conv_tc_persist_kernel<<<gridp, 64 + 32 * ew, smem_p, stream>>>(map_a, map_b, q);
        return launched(what);
    }
    const int64_t gx = ceil_div<int64_t>(p.tiles_total, p.mt);
    // split-K for layers that cannot fill the machine (4x4 .. 16x16 images, per-rank batch 2): every CTA walks the whole
    // K loop (9 taps x 16 chunks = 144 stages at ~0.5 us) on a handful of SMs; deal the (tap, chunk) iterations to
    // up to 16 CTAs per tile instead and add their partial sums in a second, elementwise pass
    {
        const int nk_h = ceil_div(c.in_pitch, kChunkK);
        const int total_it = c.ntaps * nk_h;
        const int64_t ctas = gx * ceil_div(p.n_rows, p.n_tile);
        const int64_t slab = (int64_t)c.B * c.Ho * c.Wo * c.n_cols;
        static const int split_env = [] { const char* e = getenv("CAGC_TC_SPLITK"); return e ? atoi(e) : 1; }();
        int ns = (int)std::min<int64_t>(std::min<int64_t>(16, kNumSMs / std::max<int64_t>(ctas, 1)), total_it / 4);
        if (slab > 0) ns = (int)std::min<int64_t>(ns, c.workspace_bytes / (slab * 4));
        if (split_env && c.workspace && c.out_stride == 1 && c.Hout == c.Ho && c.Wout == c.Wo && ns >= 2 &&
            (int64_t)ns * slab * 4 <= c.workspace_bytes) {
            p.it_per_split = ceil_div(total_it, ns);
            ns = ceil_div(total_it, p.it_per_split);
            p.ws = c.workspace; p.ws_slab = slab;
            dim3 grid((unsigned)gx, (unsigned)ceil_div(p.n_rows, p.n_tile), (unsigned)ns);
            conv_tc_kernel<<<grid, kThreads, smem, stream>>>(map_a, map_b, p);
            CAGC_TRY(launched(what));
            SplitKEpiP q{};
            q.ws = c.workspace; q.slab = slab; q.nsplit = ns;
            q.out_scale = c.out_scale; q.noise = c.noise; q.noise_w = c.noise_w; q.bias = c.bias; q.residual = c.residual;
            q.out = c.out; q.HW = c.Ho * c.Wo; q.n_pitch = c.n_cols; q.out_valid = c.out_valid; q.act = c.act;
            q.noise_bstride = c.noise_bstride; q.act_gain = (c.act_gain != 0.f) ? c.act_gain : kSqrt2;
            const int64_t blocks = std::min<int64_t>(ceil_div<int64_t>(slab / 4, 256), kNumSMs * 4);
            splitk_epilogue_kernel<<<(unsigned)std::max<int64_t>(blocks, 1), 256, 0, stream>>>(q);
            return launched("splitk_epilogue_kernel");
        }
    }
    CAGC_REQUIRE(gx <= 0x7fffffffLL, "%s: too many tiles", what);
    dim3 grid((unsigned)gx, ceil_div(p.n_rows, p.n_tile));
    conv_tc_kernel<<<grid, kThreads, smem, stream>>>(map_a, map_b, p);
    return launched(what);
}


// All phases of the stride-2 transposed convolution in ONE persistent launch (see TcParams::nphase).  Phases must
// share input, weights, pitches and output tensor; they differ in taps, output lattice offset and extent.
// Returns 1 when it took the call (result in *rc), 0 when the caller should launch the phases one by one.
int cagc_tc_conv_multi(cudaStream_t stream, const ConvP* ph, int nphase, const char* what, int* rc) {
    using namespace cagc::tc;
    *rc = 0;
    static const int multi_env = [] { const char* e = getenv("CAGC_TC_MULTI"); return e ? atoi(e) : 1; }();
    if (!multi_env || nphase < 2 || nphase > kMaxPhases) return 0;
    const ConvP& c = ph[0];
    if (c.in_scale != nullptr || c.in_stride != 1 || c.in_pitch % 8 != 0 || c.n_cols % 8 != 0) return 0;
    EncodeTiledFn encode = get_encode();
    if (!encode) return 0;
    int Hmax = 0, Wmax = 0, taps_total = 0;
    for (int f = 0; f < nphase; ++f) {
        Hmax = std::max(Hmax, ph[f].Ho); Wmax = std::max(Wmax, ph[f].Wo);
        taps_total += ph[f].ntaps;
    }
    if ((int64_t)c.B * Hmax * Wmax == 0 || taps_total > kMaxTaps) return 0;
    TcParams p{};
    p.out_scale = c.out_scale; p.noise = c.noise; p.noise_w = c.noise_w; p.bias = c.bias; p.out = c.out;
    p.residual = c.residual; p.act_gain = (c.act_gain != 0.f) ? c.act_gain : kSqrt2;
    p.B = c.B; p.Ho = Hmax; p.Wo = Wmax;
    p.bw = std::min(16, next_pow2(Wmax));
    p.bh = std::min(kTileM / p.bw, next_pow2(Hmax));
    p.bb = kTileM / (p.bw * p.bh);
    const int tiles_b = ceil_div(c.B, p.bb);
    p.k_valid = c.in_pitch;
    p.n_pitch = c.n_cols; p.out_valid = c.out_valid;
    p.n_rows = (c.n_cols + 15) & ~15;
    p.n_tile = std::min(256, p.n_rows);
    const int n_tiles = ceil_div(p.n_rows, p.n_tile);
    p.mt = (p.n_tile <= 128) ? 2 : 1;
    if (p.mt * p.n_tile > 256) return 0;                 // needs a double-buffered TMEM accumulator
    p.Hout = c.Hout; p.Wout = c.Wout; p.out_stride = c.out_stride;
    p.noise_bstride = c.noise_bstride; p.act = c.act; p.in_stride = 1;
    p.b_bytes = (uint32_t)p.n_tile * kChunkK * 4;
    // phases in order of decreasing tap count (round-robin dealing then balances the CTAs)
    int order[kMaxPhases];
    for (int f = 0; f < nphase; ++f) order[f] = f;
    std::sort(order, order + nphase, [&](int a, int b) { return ph[a].ntaps > ph[b].ntaps; });
    int64_t items = 0;
    int tap0 = 0, max_slab = 0;
    p.nphase = nphase;
    for (int k = 0; k < nphase; ++k) {
        const ConvP& q = ph[order[k]];
        p.ph_item0[k] = (int)items;
        p.ph_tap0[k] = tap0; p.ph_ntaps[k] = q.ntaps;
        for (int i = 0; i < q.ntaps; ++i) {
            p.taps[tap0 + i] = q.taps[i];
            max_slab = std::max(max_slab, q.taps[i].slab);
        }
        tap0 += q.ntaps;
        p.ph_Ho[k] = q.Ho; p.ph_Wo[k] = q.Wo; p.ph_oy[k] = q.out_oy; p.ph_ox[k] = q.out_ox;
        p.ph_tx[k] = ceil_div(q.Wo, p.bw); p.ph_ty[k] = ceil_div(q.Ho, p.bh);
        p.ph_tiles[k] = p.ph_tx[k] * p.ph_ty[k] * tiles_b;
        items += (int64_t)ceil_div(p.ph_tiles[k], p.mt) * n_tiles;
    }
    p.ph_item0[nphase] = (int)items;
    p.ph_gmin = 0x7fffffff;
    for (int k = 0; k < nphase; ++k) p.ph_gmin = std::min(p.ph_gmin, p.ph_item0[k + 1] - p.ph_item0[k]);
    p.ntaps = taps_total;
    if (items < 2 * kNumSMs || items > 0x7fffffff) return 0;   // tiny layers: the per-phase launches are latency-bound anyway
    const uint32_t sb = (uint32_t)p.mt * kABytes + p.b_bytes;
    const int ew = persist_epi_warps(ceil_div(taps_total * ceil_div(c.in_pitch, kChunkK), nphase));
    const int epi_bytes = ew * 2048;
    p.stages = std::max(2, std::min(kMaxStages, (int)((kSmemBudget1 - epi_bytes - 1024) / sb)));
    p.epi_off = (uint32_t)p.stages * sb;
    const size_t smem = (size_t)p.stages * sb + 1024 + epi_bytes;
    CUtensorMap map_a, map_b;
    {
        cuuint64_t dims[4] = {(cuuint64_t)c.in_pitch, (cuuint64_t)c.Win, (cuuint64_t)c.Hin, (cuuint64_t)c.B};
        cuuint64_t strides[3] = {(cuuint64_t)c.in_pitch * 4, (cuuint64_t)c.Win * c.in_pitch * 4,
                                 (cuuint64_t)c.Hin * c.Win * c.in_pitch * 4};
        cuuint32_t box[4] = {(cuuint32_t)kChunkK, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bb};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(c.in), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { *rc = fail(CAGC_E_INVALID, "%s: cuTensorMapEncodeTiled(A) failed with %d", what, (int)r); return 1; }
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)c.in_pitch, (cuuint64_t)p.n_rows, (cuuint64_t)(max_slab + 1)};
        cuuint64_t strides[2] = {(cuuint64_t)c.in_pitch * 4, (cuuint64_t)p.n_rows * c.in_pitch * 4};
        cuuint32_t box[3] = {(cuuint32_t)kChunkK, (cuuint32_t)p.n_tile, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(c.w), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { *rc = fail(CAGC_E_INVALID, "%s: cuTensorMapEncodeTiled(B) failed with %d", what, (int)r); return 1; }
    }
    static DeviceOnce attr_set{0};
    if (device_once_needed(attr_set)) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget1);
        if (e != cudaSuccess) { *rc = fail((int)e, "%s: cudaFuncSetAttribute failed: %s", what, cudaGetErrorString(e)); return 1; }
        device_once_done(attr_set);
    }
    conv_tc_persist_kernel<<<kNumSMs, 64 + 32 * ew, smem, stream>>>(map_a, map_b, p);
    *rc = launched(what);
    return 1;
}

// Small transposed convolutions (4x4 .. 16x16 inputs; per-rank batch 2): all phases AND a split of each phase's K loop in
// one conv_tc_kernel launch (blockIdx.z = (phase, split)), raw partial sums into workspace slabs laid out like the
// output, then one elementwise pass that adds the slabs of every pixel's phase in a fixed order.  Launched phase by
// phase these layers cost 4 x ~25 us of serial TMA latency for a few MFLOP.
// Returns 1 when it took the call (result in *rc), 0 when the caller should use another form.
int cagc_tc_conv_multi_splitk(cudaStream_t stream, const ConvP* ph, int nphase, float* workspace, int64_t workspace_bytes,
                              const char* what, int* rc) {
    using namespace cagc::tc;
    *rc = 0;
    static const int env = [] { const char* e = getenv("CAGC_TC_MULTI_SPLITK"); return e ? atoi(e) : 1; }();
    if (!env || !workspace || nphase < 2 || nphase > kMaxPhases) return 0;
    const ConvP& c = ph[0];
    if (c.in_scale != nullptr || c.in_stride != 1 || c.out_stride != 2 || c.in_pitch % 8 != 0 || c.n_cols % 8 != 0) return 0;
    if (c.out_scale || c.noise || c.bias || c.residual || c.act) return 0;      // the transposed convolution has a plain epilogue
    EncodeTiledFn encode = get_encode();
    if (!encode) return 0;
    int Hmax = 0, Wmax = 0, taps_total = 0, par_seen = 0;
    for (int f = 0; f < nphase; ++f) {
        if (ph[f].ntaps == 0) return 0;
        Hmax = std::max(Hmax, ph[f].Ho); Wmax = std::max(Wmax, ph[f].Wo);
        taps_total += ph[f].ntaps;
        par_seen |= 1 << (((ph[f].out_oy & 1) << 1) | (ph[f].out_ox & 1));
    }
    if (nphase != 4 || par_seen != 15 || taps_total > kMaxTaps) return 0;       // every output pixel written by exactly one phase
    const int64_t slab = (int64_t)c.B * c.Hout * c.Wout * c.n_cols;
    if ((int64_t)c.B * Hmax * Wmax == 0 || slab <= 0 || workspace_bytes < 2 * slab * 4) return 0;
    TcParams p{};
    p.out = c.out;
    p.B = c.B; p.Ho = Hmax; p.Wo = Wmax;
    p.bw = std::min(16, next_pow2(Wmax));
    p.bh = std::min(kTileM / p.bw, next_pow2(Hmax));
    p.bb = kTileM / (p.bw * p.bh);
    const int tiles_b = ceil_div(c.B, p.bb);
    p.k_valid = c.in_pitch;
    p.n_pitch = c.n_cols; p.out_valid = c.out_valid;
    p.n_rows = (c.n_cols + 15) & ~15;
    p.n_tile = std::min(256, p.n_rows);
    p.mt = 1;
    int tiles_max = 0;
    p.nphase = nphase;
    int tap0 = 0, max_slab = 0;
    for (int k = 0; k < nphase; ++k) {
        const ConvP& q = ph[k];
        p.ph_tap0[k] = tap0; p.ph_ntaps[k] = q.ntaps;
        for (int i = 0; i < q.ntaps; ++i) {
            p.taps[tap0 + i] = q.taps[i];
            max_slab = std::max(max_slab, q.taps[i].slab);
        }
        tap0 += q.ntaps;
        p.ph_Ho[k] = q.Ho; p.ph_Wo[k] = q.Wo; p.ph_oy[k] = q.out_oy; p.ph_ox[k] = q.out_ox;
        p.ph_tx[k] = ceil_div(q.Wo, p.bw); p.ph_ty[k] = ceil_div(q.Ho, p.bh);
        p.ph_tiles[k] = p.ph_tx[k] * p.ph_ty[k] * tiles_b;
        tiles_max = std::max(tiles_max, p.ph_tiles[k]);
    }
    // only layers that cannot fill the machine phase by phase
    if ((int64_t)tiles_max * ceil_div(p.n_rows, p.n_tile) * nphase >= 3 * kNumSMs) return 0;
    // narrower N tiles -> more CTAs and a deeper ring per CTA (as in cagc_tc_conv)
    while (p.n_tile > 32 && (int64_t)tiles_max * ceil_div(p.n_rows, p.n_tile) < kNumSMs / 8) {
        int nt = (p.n_tile / 2 + 15) & ~15;
        if (nt < 32) nt = 32;
        p.n_tile = nt;
    }
    const int n_tiles = ceil_div(p.n_rows, p.n_tile);
    // iterations ((tap, 32-channel chunk) steps) per CTA: about two CTAs per SM over all phases, at least 4
    const int nk_h = ceil_div(c.in_pitch, kChunkK);
    int64_t work = 0;
    for (int k = 0; k < nphase; ++k) work += (int64_t)p.ph_tiles[k] * n_tiles * p.ph_ntaps[k] * nk_h;
    int ipc = (int)std::max<int64_t>(4, ceil_div<int64_t>(work, 2 * kNumSMs));
    const int max_slabs = (int)std::min<int64_t>(16, workspace_bytes / (slab * 4));
    int z = 0, ns_max = 0;
    for (;;) {
        z = 0; ns_max = 0;
        for (int k = 0; k < nphase; ++k) {
            const int total_it = p.ph_ntaps[k] * nk_h;
            p.ph_itper[k] = std::min(ipc, total_it);
            const int ns = ceil_div(total_it, p.ph_itper[k]);
            p.ph_z0[k] = z;
            z += ns;
            ns_max = std::max(ns_max, ns);
        }
        p.ph_z0[nphase] = z;
        if (ns_max <= max_slabs) break;
        ipc += std::max(1, ipc / 4);
    }
    p.ws = workspace; p.ws_slab = slab;
    p.Hout = c.Hout; p.Wout = c.Wout; p.out_stride = c.out_stride;
    p.in_stride = 1; p.ntaps = taps_total;
    p.b_bytes = (uint32_t)p.n_tile * kChunkK * 4;
    const uint32_t stage_bytes = (uint32_t)kABytes + p.b_bytes;
    p.stages = std::max(2, std::min(kMaxStages, (int)((kConvBudget1 - 1024) / stage_bytes)));
    p.epi_off = (uint32_t)p.stages * stage_bytes;
    const size_t smem = (size_t)p.stages * stage_bytes + 1024 + kEpiStageBytes;
    CUtensorMap map_a, map_b;
    {
        cuuint64_t dims[4] = {(cuuint64_t)c.in_pitch, (cuuint64_t)c.Win, (cuuint64_t)c.Hin, (cuuint64_t)c.B};
        cuuint64_t strides[3] = {(cuuint64_t)c.in_pitch * 4, (cuuint64_t)c.Win * c.in_pitch * 4,
                                 (cuuint64_t)c.Hin * c.Win * c.in_pitch * 4};
        cuuint32_t box[4] = {(cuuint32_t)kChunkK, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bb};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(c.in), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { *rc = fail(CAGC_E_INVALID, "%s: cuTensorMapEncodeTiled(A) failed with %d", what, (int)r); return 1; }
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)c.in_pitch, (cuuint64_t)p.n_rows, (cuuint64_t)(max_slab + 1)};
        cuuint64_t strides[2] = {(cuuint64_t)c.in_pitch * 4, (cuuint64_t)p.n_rows * c.in_pitch * 4};
        cuuint32_t box[3] = {(cuuint32_t)kChunkK, (cuuint32_t)p.n_tile, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(c.w), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { *rc = fail(CAGC_E_INVALID, "%s: cuTensorMapEncodeTiled(B) failed with %d", what, (int)r); return 1; }
    }
    static DeviceOnce attr_set{0};
    if (device_once_needed(attr_set)) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget1);
        if (e != cudaSuccess) { *rc = fail((int)e, "%s: cudaFuncSetAttribute failed: %s", what, cudaGetErrorString(e)); return 1; }
        device_once_done(attr_set);
    }
    dim3 grid((unsigned)tiles_max, (unsigned)n_tiles, (unsigned)z);
    conv_tc_kernel<<<grid, kThreads, smem, stream>>>(map_a, map_b, p);
    *rc = launched(what);
    if (*rc) return 1;
    SplitKEpiP q{};
    q.ws = workspace; q.slab = slab; q.nsplit = 1;
    q.out = c.out; q.HW = c.Hout * c.Wout; q.n_pitch = c.n_cols; q.out_valid = c.out_valid; q.act = 0;
    q.act_gain = kSqrt2;
    q.W = c.Wout;
    for (int k = 0; k < nphase; ++k)
        q.ns_par[((p.ph_oy[k] & 1) << 1) | (p.ph_ox[k] & 1)] = p.ph_z0[k + 1] - p.ph_z0[k];
    const int64_t blocks = std::min<int64_t>(ceil_div<int64_t>(slab / 4, 256), kNumSMs * 4);
    splitk_epilogue_kernel<<<(unsigned)std::max<int64_t>(blocks, 1), 256, 0, stream>>>(q);
    *rc = launched("splitk_epilogue_kernel");
    return 1;
}

// all-taps plan (wgrad_tc_alltaps_kernel): taps per CTA, tap groups, splits; returns false when not applicable
static bool wgrad_alltaps_plan(int B, int H, int W, int a_pitch, int g_pitch, int ksize, int mode, int* taps_per,
                               int* groups, int* nsplits) {
    using namespace cagc::tc;
    static const int env = [] { const char* e = getenv("CAGC_TC_WGRAD_ALL"); return e ? atoi(e) : 1; }();
    if (!env || ksize != 3 || W < kWgPix || W % kWgPix != 0) return false;
    const int n_mma = (g_pitch + 15) & ~15;
    int tp = std::min(9, 512 / n_mma);
    if (tp < 3) return false;
    if (mode == 1) tp = tp / 3 * 3;                   // up-conv form: whole gradient rows (3 taps each) per CTA
    const int gr = ceil_div(9, tp);
    const int64_t tiles = (int64_t)(W / kWgPix) * H * B;
    const int64_t fixed = (int64_t)ceil_div(a_pitch, kTileM) * gr;
    int64_t sp = std::max<int64_t>(1, kNumSMs / fixed);
    sp = std::min(sp, std::max<int64_t>(1, tiles / 8));
    *taps_per = tp; *groups = gr; *nsplits = (int)sp;
    return true;
}

static int wgrad_onetap_splits(int B, int H, int W, int a_pitch, int g_pitch, int ksize);

int cagc_tc_wgrad_splits(int B, int H, int W, int a_pitch, int g_pitch, int ksize) {
    int all_sp = 0, tp, gr, sp;   // the caller sizes its partial buffer before it knows the mode: cover all kernels
    if (wgrad_alltaps_plan(B, H, W, a_pitch, g_pitch, ksize, 0, &tp, &gr, &sp)) all_sp = sp;
    if (wgrad_alltaps_plan(B, H, W, a_pitch, g_pitch, ksize, 1, &tp, &gr, &sp)) all_sp = std::max(all_sp, sp);
    return std::max(all_sp, wgrad_onetap_splits(B, H, W, a_pitch, g_pitch, ksize));
}

static int wgrad_onetap_splits(int B, int H, int W, int a_pitch, int g_pitch, int ksize) {
    using namespace cagc::tc;
    const int bw = std::min(kWgPix, next_pow2(W)), bh = std::min(kWgPix / bw, next_pow2(H)), bb = kWgPix / (bw * bh);
    const int64_t tiles = (int64_t)ceil_div(W, bw) * ceil_div(H, bh) * ceil_div(B, bb);
    const int64_t ctas = (int64_t)ceil_div(a_pitch, kTileM) * ksize * ksize;
    int64_t want = ceil_div<int64_t>(2 * 2 * kNumSMs, ctas);   // two waves of two CTAs per SM
    const int64_t max_by_k = std::max<int64_t>(1, tiles / 8);  // at least 8 pixel tiles per split
    want = std::min(want, max_by_k);
    want = std::min<int64_t>(want, 512);
    return (int)std::max<int64_t>(1, want);
}

int cagc_tc_wgrad(cudaStream_t stream, const float* a, const float* g, float* partial, int* nsplits_io, int B, int H,
                  int W, int a_pitch, int g_pitch, int ksize, int mode) {
    int nsplits = *nsplits_io;
    using namespace cagc::tc;
    const char* what = "conv_wgrad[tc]";
    CAGC_REQUIRE(a_pitch % 8 == 0 && g_pitch % 8 == 0 && a_pitch >= 32 && g_pitch >= 32,
                 "%s: channel pitches must be multiples of 8 and >= 32", what);
    CAGC_REQUIRE(g_pitch <= 256, "%s: more than 256 gradient channels per call (caller splits)", what);
    EncodeTiledFn encode = get_encode();
    if (!encode) return fail(CAGC_E_UNSUPPORTED, "%s: cuTensorMapEncodeTiled not available from the driver", what);
    {
        int tp, gr, sp;
        static const int up_env = [] { const char* e = getenv("CAGC_TC_WGRAD_UP"); return e ? atoi(e) : 1; }();
        if (mode == 1 && up_env && wgrad_alltaps_plan(B, H, W, a_pitch, g_pitch, ksize, mode, &tp, &gr, &sp) &&
            sp <= nsplits) {
            WgUpParams q{};
            q.partial = partial; q.a_pitch = a_pitch; q.g_pitch = g_pitch;
            q.n_mma = (g_pitch + 15) & ~15;
            q.b_boxes = ceil_div(q.n_mma, 32);
            q.n_ky = tp / 3;
            q.tiles_x = W / kWgPix; q.tiles_y = H;
            q.tiles_total = q.tiles_x * H * B;
            q.tiles_per_split = ceil_div(q.tiles_total, sp);
            sp = ceil_div(q.tiles_total, q.tiles_per_split);
            q.nsplits = sp;
            q.b_box_bytes = (uint32_t)(q.n_ky * kWgUpKy) * 128u;
            q.stage_bytes = 4u * kBoxBytes + (uint32_t)q.b_boxes * q.b_box_bytes;
            q.stages = std::min(kMaxStages, (int)((kSmemBudget1 - 1024) / q.stage_bytes));
            if (q.stages >= 2) {
                const size_t smem_up = (size_t)q.stages * q.stage_bytes + 1024;
                const int Hg = 2 * H + 1, Wg = 2 * W + 1;
                CUtensorMap map_a, map_ge, map_go;
                if (encode_act_map(encode, &map_a, a, B, H, W, a_pitch, kWgPix, 1, 1, 1) != 0)
                    return fail(CAGC_E_INVALID, "%s: cuTensorMapEncodeTiled(a, up all taps) failed", what);
                cuuint64_t dims[4] = {(cuuint64_t)g_pitch, (cuuint64_t)Wg, (cuuint64_t)Hg, (cuuint64_t)B};
                cuuint64_t strides[3] = {(cuuint64_t)g_pitch * 4, (cuuint64_t)Wg * g_pitch * 4, (cuuint64_t)Hg * Wg * g_pitch * 4};
                cuuint32_t es[4] = {1, 2, 1, 1};
                cuuint32_t box_e[4] = {32, 66, 1, 1}, box_o[4] = {32, 64, 1, 1};     // element stride 2: 33 / 32 samples
                if (encode(&map_ge, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(g), dims, strides, box_e, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
                    encode(&map_go, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(g), dims, strides, box_o, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                    return fail(CAGC_E_INVALID, "%s: cuTensorMapEncodeTiled(g, up all taps) failed", what);
                static DeviceOnce attr_up{0};
                if (device_once_needed(attr_up)) {
                    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_alltaps_up_kernel,
                                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget1);
                    if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute failed: %s", what, cudaGetErrorString(e));
                    device_once_done(attr_up);
                }
                *nsplits_io = sp;
                dim3 grid(ceil_div(a_pitch, kTileM), gr, sp);
                wgrad_tc_alltaps_up_kernel<<<grid, kThreads, smem_up, stream>>>(map_a, map_ge, map_go, q);
                return launched(what);
            }
        }
        if (mode == 0 && wgrad_alltaps_plan(B, H, W, a_pitch, g_pitch, ksize, mode, &tp, &gr, &sp) && sp <= nsplits) {
            WgAllParams q{};
            q.partial = partial; q.a_pitch = a_pitch; q.g_pitch = g_pitch;
            q.n_mma = (g_pitch + 15) & ~15;
            q.b_boxes = ceil_div(q.n_mma, 32);
            q.a_boxes = std::min(4, ceil_div(a_pitch, 32));
            q.taps_per = tp;
            q.tiles_x = W / kWgPix; q.tiles_y = H;
            q.tiles_total = q.tiles_x * H * B;
            q.tiles_per_split = ceil_div(q.tiles_total, sp);
            sp = ceil_div(q.tiles_total, q.tiles_per_split);
            q.nsplits = sp;
            q.stage_bytes = (uint32_t)q.a_boxes * kWgABoxBytes + (uint32_t)q.b_boxes * kBoxBytes;
            // M = 128 always addresses four channel boxes: the last stage's reads may run into the tail padding
            const int tail = (4 - q.a_boxes) * kWgABoxBytes;
            q.stages = std::max(2, std::min(kMaxStages, (int)((kSmemBudget1 - 1024 - tail) / q.stage_bytes)));
            const size_t smem_all = (size_t)q.stages * q.stage_bytes + tail + 1024;
            CUtensorMap map_a, map_g;
            {
                cuuint64_t dims[4] = {(cuuint64_t)a_pitch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
                cuuint64_t strides[3] = {(cuuint64_t)a_pitch * 4, (cuuint64_t)W * a_pitch * 4, (cuuint64_t)H * W * a_pitch * 4};
                cuuint32_t box[4] = {32, 34, 1, 1};
                cuuint32_t es[4] = {1, 1, 1, 1};
                if (encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                    return fail(CAGC_E_INVALID, "%s: cuTensorMapEncodeTiled(a, all taps) failed", what);
            }
            if (encode_act_map(encode, &map_g, g, B, H, W, g_pitch, kWgPix, 1, 1, 1) != 0)
                return fail(CAGC_E_INVALID, "%s: cuTensorMapEncodeTiled(g, all taps) failed", what);
            static DeviceOnce attr_all{0};
            if (device_once_needed(attr_all)) {
                cudaError_t e = cudaFuncSetAttribute(wgrad_tc_alltaps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     kSmemBudget1);
                if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute failed: %s", what, cudaGetErrorString(e));
                device_once_done(attr_all);
            }
            *nsplits_io = sp;
            dim3 grid(ceil_div(a_pitch, kTileM), gr, sp);
            wgrad_tc_alltaps_kernel<<<grid, kThreads, smem_all, stream>>>(map_a, map_g, q);
            return launched(what);
        }
    }
    nsplits = std::min(nsplits, wgrad_onetap_splits(B, H, W, a_pitch, g_pitch, ksize));
    WgTcParams p{};
    p.partial = partial; p.a_pitch = a_pitch; p.g_pitch = g_pitch;
    p.n_mma = (g_pitch + 15) & ~15;
    p.b_boxes = ceil_div(p.n_mma, 32);
    p.ntaps = ksize * ksize;
    p.bw = std::min(kWgPix, next_pow2(W));
    p.bh = std::min(kWgPix / p.bw, next_pow2(H));
    p.bb = kWgPix / (p.bw * p.bh);
    p.tiles_x = ceil_div(W, p.bw); p.tiles_y = ceil_div(H, p.bh);
    const int tiles_b = ceil_div(B, p.bb);
    p.tiles_total = p.tiles_x * p.tiles_y * tiles_b;
    if (nsplits > p.tiles_total) nsplits = p.tiles_total;
    p.tiles_per_split = ceil_div(p.tiles_total, nsplits);
    nsplits = ceil_div(p.tiles_total, p.tiles_per_split);      // no empty split
    *nsplits_io = nsplits;
    p.nsplits = nsplits;
    p.g_stride = (mode == 1) ? 2 : 1;
    for (int ky = 0; ky < ksize; ++ky)
        for (int kx = 0; kx < ksize; ++kx) {
            const int t = ky * ksize + kx;
            if (mode == 0) { p.dya[t] = ky - ksize / 2; p.dxa[t] = kx - ksize / 2; p.dyg[t] = 0; p.dxg[t] = 0; }
            else           { p.dya[t] = 0; p.dxa[t] = 0; p.dyg[t] = ky; p.dxg[t] = kx; }
        }
    const uint32_t stage_bytes = (4 + p.b_boxes) * kBoxBytes;
    p.stages = std::max(2, std::min(kMaxStages, (int)((kSmemBudget - 1024) / stage_bytes)));
    const size_t smem = (size_t)p.stages * stage_bytes + 1024;
    CUtensorMap map_a, map_g;
    int r = encode_act_map(encode, &map_a, a, B, H, W, a_pitch, p.bw, p.bh, p.bb, 1);
    if (r != 0) return fail(CAGC_E_INVALID, "%s: cuTensorMapEncodeTiled(a) failed with %d", what, r);
    const int Hg = mode == 1 ? 2 * H + ksize - 2 : H, Wg = mode == 1 ? 2 * W + ksize - 2 : W;
    r = encode_act_map(encode, &map_g, g, B, Hg, Wg, g_pitch, p.bw, p.bh, p.bb, p.g_stride);
    if (r != 0) return fail(CAGC_E_INVALID, "%s: cuTensorMapEncodeTiled(g) failed with %d", what, r);
    static DeviceOnce attr_set{0};
    if (device_once_needed(attr_set)) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
        if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute failed: %s", what, cudaGetErrorString(e));
        device_once_done(attr_set);
    }
    dim3 grid(ceil_div(a_pitch, kTileM), p.ntaps, nsplits);
    wgrad_tc_kernel<<<grid, kThreads, smem, stream>>>(map_a, map_g, p);
    return launched(what);
}

// returns 1 if the TMA path took the call, 0 if the caller should use the plain kernel, <0 / >1 on error
int cagc_tc_fir_nhwc(cudaStream_t stream, const float* in, const float* fir, const float* out_scale, const float* noise,
                     const float* noise_w, const float* bias, float* out, int B, int in_h, int in_w, int out_h, int out_w,
                     int pitch, int valid, int pad_x0, int pad_y0, int64_t noise_bstride, int act, const float* taps_host,
                     int* rc, const float* mask_ref, float mask_gain) {
    using namespace cagc::tc;
    *rc = 0;
    if (pitch % 4 != 0 || B > 65535) return 0;
    const char* impl = getenv("CAGC_FIR_IMPL");          // tuning knobs (development)
    const bool use_ldg = (impl && impl[0] == 'l') || pitch % 8 != 0;
    if (use_ldg && mask_ref) return 0;      // the fused activation mask lives in the TMA row-ring kernel only
    auto pick_segs = [&](int64_t per_seg, int min_rows) {
        // enough CTAs for ~4 waves of 2 CTAs per SM, but at least `min_rows` output rows per segment (3 halo rows each)
        int want = 8 * kNumSMs;
        if (const char* e = getenv("CAGC_FIR_CTAS")) want = atoi(e);
        int sg = (int)std::min<int64_t>(ceil_div<int64_t>(want, per_seg), std::max(1, out_h / min_rows));
        return sg < 1 ? 1 : sg;
    };
    if (use_ldg) {
        FirLP q{};
        q.in = in; q.fir = fir; q.out_scale = out_scale; q.noise = noise; q.noise_w = noise_w; q.bias = bias; q.out = out;
        q.in_h = in_h; q.in_w = in_w; q.out_h = out_h; q.out_w = out_w; q.pitch = pitch; q.valid = valid;
        q.pad_x0 = pad_x0; q.pad_y0 = pad_y0; q.act = act; q.noise_bstride = noise_bstride;
        const int gx = ceil_div(out_w * (pitch / 4), 256);
        int segs = pick_segs((int64_t)gx * B, 16);
        q.rows_per_seg = ceil_div(out_h, segs);
        segs = ceil_div(out_h, q.rows_per_seg);
        if (segs > 65535) return 0;
        fir_nhwc_ldg_kernel<4, 4><<<dim3(gx, segs, B), 256, 0, stream>>>(q);
        *rc = launched("fir_nhwc_ldg_kernel");
        return 1;
    }
    EncodeTiledFn encode = get_encode();
    if (!encode) return 0;
    FirSP p{};
    p.fir = fir; p.out_scale = out_scale; p.noise = noise; p.noise_w = noise_w; p.bias = bias; p.out = out;
    p.out_h = out_h; p.out_w = out_w; p.pitch = pitch; p.valid = valid;
    p.pad_x0 = pad_x0; p.pad_y0 = pad_y0; p.act = act; p.noise_bstride = noise_bstride;
    p.mask_ref = mask_ref; p.mask_gain = mask_gain;
    constexpr int PXT = 2;
    // chunk width: the widest of 128 / 64 / 32 channels that divides the pitch (measured: +5..40% on 256..512
    // channel tensors, longer contiguous runs per TMA row); ragged pitches use 32 + a tail chunk
    p.main_w = pitch % 128 == 0 ? 128 : pitch % 64 == 0 ? 64 : 32;
    if (const char* e = getenv("CAGC_FIR_CW")) p.main_w = atoi(e);
    p.main_tx = PXT * 1024 / p.main_w;
    p.n_main_chunks = pitch / p.main_w;
    p.tail_w = pitch % p.main_w;
    p.tail_tx = p.tail_w ? std::min(252, PXT * 1024 / p.tail_w) / PXT * PXT : 0;   // TMA box <= 256 pixels
    p.n_main = p.n_main_chunks * ceil_div(out_w, p.main_tx);
    const int n_tail = p.tail_w ? ceil_div(out_w, p.tail_tx) : 0;
    int segs = pick_segs((int64_t)(p.n_main + n_tail) * B, 16);
    p.rows_per_seg = ceil_div(out_h, segs);
    segs = ceil_div(out_h, p.rows_per_seg);
    if (segs > 65535) return 0;
    p.stages = std::min(6, p.rows_per_seg + 3);        // 6 x 8.5 KB: two or three CTAs per SM
    if (const char* e = getenv("CAGC_FIR_STAGES")) p.stages = std::max(2, std::min(kFirMaxStages, atoi(e)));
    int sbytes = p.n_main_chunks ? (p.main_tx + 3) * p.main_w * 4 : 0;
    if (p.tail_w) sbytes = std::max(sbytes, (p.tail_tx + 3) * p.tail_w * 4);
    p.stage_stride = (sbytes + 127) & ~127;
    if (taps_host)
        for (int i = 0; i < 16; ++i) {
            const float k = taps_host[15 - i];            // flipped: true convolution
            p.taps2[i] = make_float2(k, k);
        }

    CUtensorMap map_main, map_tail;
    cuuint64_t dims[4] = {(cuuint64_t)pitch, (cuuint64_t)in_w, (cuuint64_t)in_h, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)pitch * 4, (cuuint64_t)in_w * pitch * 4, (cuuint64_t)in_h * in_w * pitch * 4};
    cuuint32_t es[4] = {1, 1, 1, 1};
    auto enc = [&](CUtensorMap* m, int cw, int tx) {
        cuuint32_t box[4] = {(cuuint32_t)cw, (cuuint32_t)(tx + 3), 1, 1};
        return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), dims, strides, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    if (p.n_main_chunks && !enc(&map_main, p.main_w, p.main_tx)) return 0;
    if (p.tail_w && !enc(&map_tail, p.tail_w, p.tail_tx)) return 0;
    if (!p.n_main_chunks) map_main = map_tail;
    if (!p.tail_w) map_tail = map_main;
    const size_t smem = (size_t)p.stages * p.stage_stride + 128;
    static DeviceOnce attr_set{0};
    if (device_once_needed(attr_set)) {
        cudaError_t e = cudaFuncSetAttribute(fir_nhwc_stream_kernel<4, 4, PXT, true>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(fir_nhwc_stream_kernel<4, 4, PXT, false>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        // ask for the largest shared-memory carve-out: the default heuristic sized it for ONE resident CTA
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(fir_nhwc_stream_kernel<4, 4, PXT, true>,
                                     cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(fir_nhwc_stream_kernel<4, 4, PXT, false>,
                                     cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) { *rc = fail((int)e, "fir_nhwc[stream]: cudaFuncSetAttribute failed"); return 1; }
        device_once_done(attr_set);
    }
    if (smem > 100 * 1024) return 0;
    dim3 grid(p.n_main + n_tail, segs, B);
    if (taps_host)
        fir_nhwc_stream_kernel<4, 4, PXT, true><<<grid, 256, smem, stream>>>(map_main, map_tail, p);
    else
        fir_nhwc_stream_kernel<4, 4, PXT, false><<<grid, 256, smem, stream>>>(map_main, map_tail, p);
    *rc = launched("fir_nhwc_stream_kernel");
    return 1;
}

extern "C" {

int cagc_tc_available(void) {
    // the tcgen05 / TMA kernels exist only for sm_100: answer for the CURRENT device
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    return (major == 10 && tc::get_encode() != nullptr) ? 1 : 0;
}

int cagc_modulate(cagc_stream_t stream_, const float* x, const float* s, float* out, int B, int H, int W, int pitch) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(x && out, "modulate: null pointer");
    CAGC_REQUIRE(pitch > 0 && pitch % 4 == 0, "modulate: pitch must be a positive multiple of 4");
    const int64_t per4 = (int64_t)H * W * pitch / 4;
    if (per4 == 0 || B == 0) return 0;
    CAGC_REQUIRE(per4 < (1LL << 30) && B <= 65535, "modulate: tensor too large");
    int64_t bx = ceil_div<int64_t>(per4, 256);
    const int64_t cap = std::max<int64_t>(1, (int64_t)kNumSMs * 16 / B);
    if (bx > cap) bx = cap;
    cagc::tc::modulate_kernel<<<dim3((unsigned)bx, (unsigned)B), 256, 0, stream>>>(x, s, out, (int)per4, pitch / 4);
    return launched("modulate_kernel");
}

static void same_conv_params(ConvP& p, int B, int H, int W, int in_pitch, int out_pitch, int out_valid, int ksize) {
    p.B = B; p.Hin = H; p.Win = W; p.in_pitch = in_pitch; p.Ho = H; p.Wo = W; p.in_stride = 1;
    p.n_cols = out_pitch; p.out_valid = out_valid; p.Hout = H; p.Wout = W; p.out_stride = 1; p.out_oy = 0; p.out_ox = 0;
    p.ntaps = ksize * ksize;
    for (int ky = 0; ky < ksize; ++ky)
        for (int kx = 0; kx < ksize; ++kx) p.taps[ky * ksize + kx] = Tap{ky - ksize / 2, kx - ksize / 2, ky * ksize + kx};
}

int64_t cagc_conv_same_psw_bytes(int B, int H, int W, int in_pitch, int out_pitch, int ksize) {
    if (B <= 0 || H <= 0 || W <= 0 || in_pitch <= 0 || out_pitch <= 0 || in_pitch % 8 || out_pitch % 8 || ksize < 1 ||
        ksize % 2 == 0 || ksize * ksize > kMaxTaps)
        return 0;
    ConvP p{};
    same_conv_params(p, B, H, W, in_pitch, out_pitch, out_pitch, ksize);
    p.w_bslabs = ksize * ksize;
    if (!cagc_tc_conv_takes_sample_weights(p)) return 0;
    return (int64_t)B * ksize * ksize * ((out_pitch + 15) & ~15) * in_pitch * 4;
}

int cagc_conv_same_psw(cagc_stream_t stream_, const float* in, const float* w_slabs, const float* in_scale,
                       const float* out_scale, const float* noise, const float* noise_w, const float* bias, float* out,
                       int B, int H, int W, int in_pitch, int out_pitch, int out_valid, int ksize, int64_t noise_bstride,
                       int act, float* w_ps, int64_t w_ps_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CAGC_REQUIRE(in && w_slabs && in_scale && out && w_ps, "conv_same_psw: null pointer");
    CAGC_REQUIRE(!noise || noise_w, "conv_same_psw: noise without noise weight");
    CAGC_REQUIRE(aligned16(in) && aligned16(w_slabs) && aligned16(out) && aligned16(w_ps) && aligned16(in_scale),
                 "conv_same_psw: pointers must be 16-byte aligned");
    const int64_t need = cagc_conv_same_psw_bytes(B, H, W, in_pitch, out_pitch, ksize);
    if (need == 0) return fail(CAGC_E_UNSUPPORTED, "conv_same_psw: shape not taken by the per-sample-weight path");
    CAGC_REQUIRE(w_ps_bytes >= need, "conv_same_psw: weight workspace too small (%lld < %lld bytes)", (long long)w_ps_bytes,
                 (long long)need);
    const int64_t per4 = (int64_t)ksize * ksize * ((out_pitch + 15) & ~15) * in_pitch / 4;
    CAGC_REQUIRE(per4 < (1LL << 30) && B <= 65535, "conv_same_psw: tensor too large");
    int64_t bx = std::min<int64_t>(ceil_div<int64_t>(per4, 256), std::max<int64_t>(1, (int64_t)kNumSMs * 8 / B));
    cagc::tc::modulate_weights_kernel<<<dim3((unsigned)bx, (unsigned)B), 256, 0, stream>>>(w_slabs, in_scale, w_ps, (int)per4,
                                                                                       in_pitch / 4);
    CAGC_TRY(launched("modulate_weights_kernel"));
    ConvP p{};
    same_conv_params(p, B, H, W, in_pitch, out_pitch, out_valid, ksize);
    p.in = in; p.w = w_ps; p.out_scale = out_scale; p.noise = noise; p.noise_w = noise_w; p.bias = bias; p.out = out;
    p.noise_bstride = noise_bstride; p.act = act ? kActLrelu : kActNone; p.w_bslabs = ksize * ksize;
    return cagc_tc_conv(stream, p, "conv_same_psw[tc]");
}

}  // extern "C"

int cagc_tc_conv_takes_sample_weights(const cagc::ConvP& c) {
    int rc = 0;
    return try_halo_conv(nullptr, c, "conv", nullptr, &rc, true);
}
