// Hardware probe (development aid): can a tcgen05 shared-memory matrix descriptor (K-major, 128-byte swizzle) start
// at a 128-byte row that is NOT aligned to the 1024-byte swizzle atom?  This is what an implicit-GEMM convolution
// needs to read the taps of a 3x3 window as shifted views of ONE halo tile instead of nine TMA loads.
// For every row shift d = 0..9 and both settings of the descriptor's base-offset field it prints the max error of
// D[m][n] = sum_k A[d + m][k] * B[n][k] against the host result (integer-valued operands: exact in TF32).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_shift_probe umma_shift_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (clock64() - t0 > 2000000000LL) { printf("probe: mbarrier timeout\n"); __trap(); }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t base_off, int mn_major) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    if (!mn_major) {
        d |= (uint64_t)1 << 16;
        d |= (uint64_t)(1024 >> 4) << 32;
        d |= (uint64_t)2 << 61;            // SWIZZLE_128B
    } else {
        d |= (uint64_t)(4096 >> 4) << 16;  // LBO: next 32-element chunk along MN
        d |= (uint64_t)(512 >> 4) << 32;   // SBO: next group of 4 K rows
        d |= (uint64_t)1 << 61;            // SWIZZLE_128B_BASE32B
    }
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)1 << 46;
    return d;
}

constexpr int kRows = 160, kN = 32, kShifts = 10;

// mode 0: K-major A (rows = M, 32 floats of K per row): shift = rows.  mode 1: MN-major A ([K rows][32 M] boxes,
// 4 boxes side by side for M = 128): shift = K rows (pixels) -- the weight-gradient case.
__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap map_a,
                                                    const __grid_constant__ CUtensorMap map_b, float* out, int mode_in) {
    // mode 2: the math of mode 0, but A arrives as five 33-row TMA boxes whose shared-memory destinations are only
    // 128-byte aligned (33 * 128 = 4224 bytes apart): is the TMA swizzle a function of the absolute address as well?
    const int mode = mode_in == 2 ? 0 : mode_in;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_addr = base, b_addr = base + 32 * 1024;
    const uint32_t bar_ld = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar_ld, 1);
        mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        if (mode_in == 2) {
            mbar_expect_tx(bar_ld, 5 * 33 * 128 + kN * 128);
            for (int c = 0; c < 5; ++c) tma_load_2d(a_addr + c * 33 * 128, &map_a, bar_ld, 0, 33 * c);
            tma_load_2d(b_addr, &map_b, bar_ld, 0, 0);
        } else if (mode == 0) {
            mbar_expect_tx(bar_ld, kRows * 128 + kN * 128);
            tma_load_2d(a_addr, &map_a, bar_ld, 0, 0);            // kRows rows of 128 bytes, SW128
            tma_load_2d(b_addr, &map_b, bar_ld, 0, 0);            // kN rows of 128 bytes, SW128
        } else {
            // A: 4 boxes [48 K rows][32 M] (SW128 32B-atom), 6 KB apart -> LBO below is 6144; B: 1 box [48][32]
            mbar_expect_tx(bar_ld, 5 * 48 * 128);
            for (int c = 0; c < 4; ++c) tma_load_2d(a_addr + c * 6144, &map_a, bar_ld, 32 * c, 0);
            tma_load_2d(b_addr, &map_b, bar_ld, 0, 0);
        }
    }
    mbar_wait(bar_ld, 0);
    uint32_t ph = 0;
    for (int bo_mode = 0; bo_mode < 2; ++bo_mode)
        for (int d = 0; d < kShifts; ++d) {
            if (threadIdx.x == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                if (mode == 1) idesc |= (1u << 15) | (1u << 16);
                for (int k = 0; k < 4; ++k) {
                    uint64_t ad, bd;
                    if (mode == 0) {
                        const uint32_t sa = a_addr + d * 128;
                        ad = make_desc(sa, bo_mode ? (sa >> 7) : 0, 0) + (uint64_t)(2 * k);
                        bd = make_desc(b_addr, 0, 0) + (uint64_t)(2 * k);
                    } else {
                        // K step = 8 pixels = 8 rows of 128 bytes = two 4-row atoms (SBO 512); shifted by d rows
                        const uint32_t sa = a_addr + (d + 8 * k) * 128, sb = b_addr + (8 * k) * 128;
                        uint64_t da = make_desc(sa, bo_mode ? (sa >> 7) : 0, 1);
                        da &= ~((uint64_t)0x3fff << 16);
                        da |= (uint64_t)(6144 >> 4) << 16;
                        ad = da;
                        bd = make_desc(sb, 0, 1);
                    }
                    uint32_t accum = k > 0;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(accum) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_mma) : "memory");
            }
            mbar_wait(bar_mma, ph);
            ph ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t r[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                           "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                           "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                           "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float* dst = out + (((size_t)bo_mode * kShifts + d) * 128 + warp * 32 + lane) * kN;
            for (int n = 0; n < kN; ++n) dst[n] = __uint_as_float(r[n]);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
        }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    EncodeTiledFn encode = (EncodeTiledFn)sym;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int mode_in = 0; mode_in < 3; ++mode_in) {
        const int mode = mode_in == 2 ? 0 : mode_in;
        // mode 0: A [kRows][32] (K contiguous), B [kN][32].  mode 1: A [48 pixels][128 ch] (M contiguous), B [48][32]
        const int a_rows = mode == 0 ? kRows : 48, a_cols = mode == 0 ? 32 : 128;
        const int b_rows = mode == 0 ? kN : 48, b_cols = 32;
        std::vector<float> ha((size_t)a_rows * a_cols), hb((size_t)b_rows * b_cols);
        for (auto& v : ha) v = (float)(rand() % 9 - 4);
        for (auto& v : hb) v = (float)(rand() % 7 - 3);
        float *da, *db, *dout;
        cudaMalloc(&da, ha.size() * 4); cudaMalloc(&db, hb.size() * 4);
        cudaMalloc(&dout, 2 * kShifts * 128 * kN * 4);
        cudaMemcpy(da, ha.data(), ha.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(dout, 0, 2 * kShifts * 128 * kN * 4);
        CUtensorMap ma, mb;
        cuuint32_t es[2] = {1, 1};
        {
            cuuint64_t dims[2] = {(cuuint64_t)a_cols, (cuuint64_t)a_rows};
            cuuint64_t strides[1] = {(cuuint64_t)a_cols * 4};
            cuuint32_t box[2] = {32, (cuuint32_t)(mode_in == 2 ? 33 : mode == 0 ? kRows : 48)};
            CUresult r = encode(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, da, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                mode == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r) { printf("encode A failed %d\n", (int)r); return 1; }
        }
        {
            cuuint64_t dims[2] = {(cuuint64_t)b_cols, (cuuint64_t)b_rows};
            cuuint64_t strides[1] = {(cuuint64_t)b_cols * 4};
            cuuint32_t box[2] = {32, (cuuint32_t)b_rows};
            CUresult r = encode(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, db, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                mode == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r) { printf("encode B failed %d\n", (int)r); return 1; }
        }
        probe_kernel<<<1, 128, 64 * 1024>>>(ma, mb, dout, mode_in);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: kernel failed: %s\n", mode, cudaGetErrorString(e)); return 1; }
        std::vector<float> ho((size_t)2 * kShifts * 128 * kN);
        cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
        for (int bo = 0; bo < 2; ++bo)
            for (int d = 0; d < kShifts; ++d) {
                double maxerr = 0;
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < kN; ++n) {
                        double ref = 0;
                        if (mode == 0) {
                            for (int k = 0; k < 32; ++k) ref += (double)ha[(size_t)(d + m) * 32 + k] * hb[(size_t)n * 32 + k];
                        } else {   // D[m][n] = sum_{pix < 32} A[d + pix][m] * B[pix][n]
                            for (int k = 0; k < 32; ++k) ref += (double)ha[(size_t)(d + k) * 128 + m] * hb[(size_t)k * 32 + n];
                        }
                        maxerr = fmax(maxerr, fabs(ref - ho[(((size_t)bo * kShifts + d) * 128 + m) * kN + n]));
                    }
                printf("mode %d (%s) base_offset=%s shift %d rows: max err %g %s\n", mode_in, mode_in == 2 ? "K-major A, 33-row TMA boxes" : mode == 0 ? "K-major A" : "MN-major A",
                       bo ? "(addr>>7)&7" : "0", d, maxerr, maxerr == 0 ? "EXACT" : "WRONG");
            }
        cudaFree(da); cudaFree(db); cudaFree(dout);
    }
    return 0;
}
