// mlp_tcgen05.cu -- tensor-core Linear layers for the policy / value nets: rollout-time forward and the whole autograd of the PPO update.
//   y[M][N] = act(x[M][Kp] W[N][Kp]^T + b),  bf16 operands (K-major, zero padded to Kp % 64 == 0), fp32 accumulation in TMEM.  sm_100a only.
// Two kernels share the TMA maps, the operand layout (128-byte swizzle) and the epilogue building blocks:
//   k_linear_tc   one CTA per 128 x 128 tile: TMA (cp.async.bulk.tensor) -> 3-stage smem ring -> tcgen05.mma cta_group::1 (one elected thread, UMMA 128x128x16,
//                 128 TMEM columns, 2 CTAs/SM so that one CTA's epilogue overlaps the other's MMAs) -> tcgen05.ld epilogue in 8 warps.  Rollout-time forward
//                 (M = 4096), small shapes, split-K by fp32 atomics, and the fallback of every fused path.
//   k_linear_tc2  CTA pairs (cta_group::2, 256 x 256 tiles), persistent, two TMEM accumulators, 5-stage ring, 16 epilogue warps, split-K work items whose
//                 partial tiles the TMA engine adds into the output: the training GEMMs (M >= 16384 rows, or a reduction over >= 16384 rows).
// Epilogues (both kernels): bias + activation; outputs staged in swizzled shared-memory tiles and stored by the TMA engine (fp32 pre-activation z, bf16 / fp32 y,
// the TRANSPOSED bf16 y for the backward pass' dW GEMM); DACT variant = the backward pass' dX GEMM with the previous layer's activation backward fused in
// (z tile fetched by TMA, dz / dz^T stored by TMA, bias gradient by a warp transpose-reduce).  Host entry points at the end of the file (C ABI: include/uhc_nn.h).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string>
#include "../../include/uhc_nn.h"

namespace {
#ifndef UHC_TC_BN
#define UHC_TC_BN 128
#endif
#ifndef UHC_TC_STAGES
#define UHC_TC_STAGES 3
#endif
constexpr int BM = 128, BN = UHC_TC_BN, BK = 64, UK = 16, STAGES = UHC_TC_STAGES;   // 3 stages x 32 KB -> two CTAs per SM: one CTA's epilogue overlaps the other's MMAs
constexpr int STAGE_BYTES = (BM * BK + BN * BK) * 2;             // 32 KB
constexpr int NTHREADS = 320;                                    // warp 0: TMA producer, warp 1: MMA issuer, warps 2..9: epilogue (two per TMEM lane quarter, half the columns each)
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;    // + alignment slack + barriers
// epilogue staging inside the (then idle) operand ring, one block per epilogue warp: [fp32 32x32 tile A | fp32 32x32 tile B | bf16 32x32 tile | bias row]
constexpr int EPI_F32_A = 0, EPI_F32_B = 4096, EPI_BF16 = 8192, EPI_BIAS = 10240, EPI_BLOCK = 11264;
static_assert(8 * EPI_BLOCK <= STAGES * STAGE_BYTES, "epilogue staging must fit the operand ring");
#ifndef UHC_TC_TMA_STORE
#define UHC_TC_TMA_STORE 1       /* outputs leave through the TMA engine (cp.async.bulk.tensor shared -> global) where their row pitch allows it */
#endif
thread_local std::string g_tc_err;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(cnt)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// one 32 x 32 tile, shared -> global, clipped at the tensor's bounds by the TMA engine
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
// this thread's row of 32 fp32 values into a 32 x 128 B tile laid out for a SWIZZLE_128B tensor map (16-byte chunk j of row r sits at chunk j ^ (r & 7))
__device__ __forceinline__ void stage_row_f32_sw128(uint32_t tile, int r, const float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(tile + r * 128 + ((j ^ (r & 7)) << 4)), "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
}
// the same as bf16 into a 32 x 64 B tile for a SWIZZLE_64B map (chunk j of row r sits at chunk j ^ ((r >> 1) & 3))
__device__ __forceinline__ void stage_row_bf16_sw64(uint32_t tile, int r, const float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]), p1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]), p3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile + r * 64 + ((j ^ ((r >> 1) & 3)) << 4)), "r"(*(uint32_t *)&p0), "r"(*(uint32_t *)&p1),
                     "r"(*(uint32_t *)&p2), "r"(*(uint32_t *)&p3) : "memory");
    }
}
// this thread's row m (= lane) of 32 values as COLUMN m of the transposed 32 x 64 B bf16 tile (row j = output column nb + j), SWIZZLE_64B layout
__device__ __forceinline__ void stage_col_bf16_sw64(uint32_t tile, int m, const float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const unsigned short h = __bfloat16_as_ushort(__float2bfloat16_rn(v[j]));
        asm volatile("st.shared.b16 [%0], %1;" ::"r"(tile + j * 64 + ((((m >> 3) ^ ((j >> 1) & 3))) << 4) + ((m & 7) << 1)), "h"(h) : "memory");
    }
}
// K-major, 128B-swizzled operand tile: 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor, version 1 = Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset
    d |= (uint64_t)1 << 46;                 // descriptor version
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// erf by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7, i.e. below the bf16 / fp32 noise of this path) on the fast exp / rcp units:
// the epilogue is instruction-latency bound, and erff() alone costs more than the tile's tensor-core time
__device__ __forceinline__ float gelu_fast(float z) {
    const float x = fabsf(z) * 0.70710678118654752f;
    const float t = __fdividef(1.0f, fmaf(0.3275911f, x, 1.0f));
    const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
    const float e = 1.0f - poly * __expf(-x * x);            // erf(|z| / sqrt 2)
    return 0.5f * z + 0.5f * fabsf(z) * e;                  // 0.5 z (1 + sign(z) e)
}
__device__ __forceinline__ float act_f(float z, int act) {
    switch (act) {
    case UHC_ACT_GELU: return gelu_fast(z);
    case UHC_ACT_TANH: return tanhf(z);
    case UHC_ACT_RELU: return z > 0.f ? z : 0.f;
    case UHC_ACT_SIGMOID: return 1.0f / (1.0f + expf(-z));
    }
    return z;
}

// act'(z) for the fused activation backward; GELU through the same erf approximation (one exp shared by the erf tail and the density)
__device__ __forceinline__ float act_b_fast(float z, int act) {
    switch (act) {
    case UHC_ACT_GELU: {
        const float x = fabsf(z) * 0.70710678118654752f;
        const float t = __fdividef(1.0f, fmaf(0.3275911f, x, 1.0f));
        const float E = __expf(-x * x);                                   // exp(-z^2 / 2)
        const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
        const float e = 1.0f - poly * E;                                  // erf(|z| / sqrt 2)
        return 0.5f + copysignf(0.5f * e, z) + z * E * 0.3989422804014327f;
    }
    case UHC_ACT_TANH: { const float th = tanhf(z); return 1.0f - th * th; }
    case UHC_ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case UHC_ACT_SIGMOID: { const float sg = 1.0f / (1.0f + __expf(-z)); return sg * (1.0f - sg); }
    }
    return 1.f;
}
// one warp writes its 32 x 32 block (thread = row, vals = that row's 32 columns) to dst[row0 + rr][col0 + lane], rows in order
__device__ __forceinline__ void stage_store(float *stg, float *__restrict__ dst, const float (&vals)[32], int row0, int col0, int M, int N, int lane, bool atomic) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = vals[j];
    __syncwarp();
    const int n = col0 + lane;
    if (n < N) {
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) {
            const int grow = row0 + rr;
            if (grow < M) {
                const float x = stg[rr * 33 + lane];
                if (atomic) atomicAdd(dst + (size_t)grow * N + n, x); else dst[(size_t)grow * N + n] = x;
            }
        }
    }
}

#ifndef UHC_TC_MINB
#define UHC_TC_MINB (STAGES * STAGE_BYTES <= 100 * 1024 ? 2 : 1)     /* resident CTAs per SM the register budget is cut for (experiment knob) */
#endif
// DACT = the backward pass' dX GEMM with the activation backward fused into its epilogue: the accumulator (dh of the previous layer) is multiplied by act'(z_prev)
// (z tile fetched by TMA through mapZ) and leaves as bf16 dz (mapYb), its transpose (mapYT) and 32-row partial column sums added to dbias (the bias gradient);
// no fp32 dh is ever written.
template <bool DACT>
__global__ void __launch_bounds__(NTHREADS, UHC_TC_MINB)
k_linear_tc(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const float *__restrict__ bias,
            __nv_bfloat16 *__restrict__ ybf, float *__restrict__ yf, float *__restrict__ zf, int M, int N, int Kp, int ldy, int act, int ksplit,
            const __grid_constant__ CUtensorMap mapZ, const __grid_constant__ CUtensorMap mapYf, const __grid_constant__ CUtensorMap mapYb, int tma_mask,
            const __grid_constant__ CUtensorMap mapYT, float *__restrict__ dbias) {
    // tma_mask: bit 0 = zf, bit 1 = yf, bit 2 = ybf leave through their tensor map (32 x 32 boxes from swizzled staging tiles) instead of per-thread stores;
    // bit 3 = the TRANSPOSE of the bf16 activation ([N][M pitch], what the backward pass' dW = dz^T h GEMM reads as its K-major operand) is emitted as well
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = (uint64_t *)(smem + STAGES * STAGE_BYTES);   // full[STAGES], empty[STAGES], tmem_full
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * STAGES + 1);
    uint64_t *zbars = bars + 16;                                   // DACT: one barrier per epilogue warp for its z tiles
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    // split-K (gridDim.z slices of the reduction, fp32 atomic accumulation into yf): used for dW = dZ^T X whose output has few tiles
    const int nkb_all = Kp / BK, kb0 = (int)(((long)nkb_all * blockIdx.z) / ksplit), nkb = (int)(((long)nkb_all * (blockIdx.z + 1)) / ksplit) - kb0;
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), tfull = smem_u32(bars + 2 * STAGES);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB));
        for (int s = 0; s < STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull, 1);
        if (DACT) for (int w = 0; w < 8; w++) mbar_init(smem_u32(zbars + w), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the first ring fill does not depend on anything the other warps set up: start it before the CTA-wide barrier
        for (int kb = 0; kb < STAGES && kb < nkb; ++kb) {
            mbar_expect_tx(full0 + 8 * kb, STAGE_BYTES);
            const uint32_t a = smem_u32(smem + kb * STAGE_BYTES), b = a + BM * BK * 2;
            tma_load_2d(a, &mapA, full0 + 8 * kb, (kb0 + kb) * BK, m0);
            tma_load_2d(b, &mapB, full0 + 8 * kb, (kb0 + kb) * BK, n0);
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = STAGES; kb < nkb; ++kb) {
                const int s = kb % STAGES; const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(empty0 + 8 * s, ph ^ 1);
                mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);
                const uint32_t a = smem_u32(smem + s * STAGE_BYTES), b = a + BM * BK * 2;
                tma_load_2d(a, &mapA, full0 + 8 * s, (kb0 + kb) * BK, m0);
                tma_load_2d(b, &mapB, full0 + 8 * s, (kb0 + kb) * BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D = F32, A = B = BF16, both K-major, N = 128, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES; const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(full0 + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a = smem_u32(smem + s * STAGE_BYTES), b = a + BM * BK * 2;
                const uint64_t ad = make_desc(a), bd = make_desc(b);
#pragma unroll
                for (int k = 0; k < BK / UK; ++k) umma_bf16(tmem, ad + (uint64_t)(k * UK * 2 >> 4), bd + (uint64_t)(k * UK * 2 >> 4), idesc, (kb | k) != 0);
                umma_commit(empty0 + 8 * s);             // frees the smem stage when these MMAs retire
            }
            umma_commit(tfull);                          // accumulator complete
        }
    } else {
        const int q = warp & 3;                          // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;                // which half of the tile's columns this warp drains
        mbar_wait(tfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = m0 + 32 * q + lane;
        if constexpr (DACT) {
            uint8_t *blk = smem + (warp - 2) * EPI_BLOCK;
            const uint32_t blk_s = smem_u32(blk), zbar = smem_u32(zbars + (warp - 2));
            const int r0 = m0 + 32 * q, cfirst = half * (BN / 64);
            static_assert(BN / 64 == 2, "the z tiles of a warp's two chunks live in the two fp32 staging areas");
            const bool zl0 = r0 < M && n0 + cfirst * 32 < N, zl1 = r0 < M && n0 + cfirst * 32 + 32 < N;
            if (lane == 0 && zl0) {      // the operand ring is idle now: fetch this warp's z tiles while the accumulator is read from TMEM
                mbar_expect_tx(zbar, (zl1 ? 2u : 1u) * 4096u);
                tma_load_2d(blk_s + EPI_F32_A, &mapZ, zbar, n0 + cfirst * 32, r0);
                if (zl1) tma_load_2d(blk_s + EPI_F32_B, &mapZ, zbar, n0 + cfirst * 32 + 32, r0);
            }
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                float v[32];
                const int c = cfirst + cc, nb = n0 + c * 32;
                const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(c * 32);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                             : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]), "=f"(v[10]),
                               "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]),
                               "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]), "+f"(v[16]), "+f"(v[17]), "+f"(v[18]), "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]), "+f"(v[24]), "+f"(v[25]), "+f"(v[26]), "+f"(v[27]), "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31]) :: "memory");
                if (cc == 0 && zl0) mbar_wait(zbar, 0);
                if (cc == 1 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the bf16 tile of the first chunk has been read
                __syncwarp();
                const uint32_t zarea = blk_s + (cc ? EPI_F32_B : EPI_F32_A);
                if (cc ? zl1 : zl0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {        // this thread's row of the 128B-swizzled z tile
                        float a, b, cq, d;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(cq), "=f"(d) : "r"(zarea + lane * 128 + ((j ^ (lane & 7)) << 4)) : "memory");
                        v[4 * j] *= act_b_fast(a, act); v[4 * j + 1] *= act_b_fast(b, act); v[4 * j + 2] *= act_b_fast(cq, act); v[4 * j + 3] *= act_b_fast(d, act);
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = (row < M && nb + j < N) ? v[j] : 0.f;
                __syncwarp();                             // every lane has read its z row: the area now stages the transposed tile
                if (nb < ldy) stage_row_bf16_sw64(blk_s + EPI_BF16, lane, v);
                if (nb < N) stage_col_bf16_sw64(zarea, lane, v);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if (nb < ldy) tma_store_2d(&mapYb, blk_s + EPI_BF16, nb, r0);
                    if (nb < N) tma_store_2d(&mapYT, zarea, r0, nb);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (dbias && nb < N) {                    // column sums over the warp's 32 rows: transpose-reduce (31 shuffles), lane l ends with column l
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) {
                        const bool up = (lane & o) != 0;
#pragma unroll
                        for (int k = 0; k < o; ++k) {
                            const float send = up ? v[k] : v[k + o], keep = up ? v[k + o] : v[k];
                            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                        }
                    }
                    if (nb + lane < N) atomicAdd(dbias + nb + lane, v[0]);
                }
            }
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the staging tiles have been read (the writes drain after the CTA)
        } else {
#pragma unroll 1
        for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
            float v[32];
            const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(c * 32);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                         "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]), "=f"(v[10]),
                           "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]),
                           "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]), "+f"(v[16]), "+f"(v[17]), "+f"(v[18]), "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]), "+f"(v[24]), "+f"(v[25]), "+f"(v[26]), "+f"(v[27]), "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31]) :: "memory");   // (operands: the loaded registers may not be read before the wait)
            // every output of this 32 x 32 chunk is staged in the warp's block of the (now idle) operand ring.  Where the output's row pitch allows a tensor map
            // (tma_mask) the tile is written in the map's swizzled layout and ONE elected lane hands it to the TMA engine (bounds are clipped by the hardware);
            // otherwise fp32 goes through a 32 x 33 tile so that every store instruction writes one full 128-byte row segment, bf16 as 16-byte stores of the row.
            const int nb = n0 + c * 32;
            uint8_t *blk = smem + (warp - 2) * EPI_BLOCK;
            float *stg = reinterpret_cast<float *>(blk);                      // legacy 32 x 33 tile over the two fp32 areas
            float *sbias = reinterpret_cast<float *>(blk + EPI_BIAS);
            const uint32_t blk_s = smem_u32(blk);
            if (tma_mask) { if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }      // the previous chunk's tiles have been read
            __syncwarp();
            sbias[lane] = (bias && blockIdx.z == 0 && nb + lane < N) ? __ldg(bias + nb + lane) : 0.f;   // this chunk's 32 biases, read back as broadcasts
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (row < M && nb + j < N) ? v[j] + sbias[j] : 0.f;
            if (zf) {                                                            // pre-activation (for the backward pass)
                if (tma_mask & 1) stage_row_f32_sw128(blk_s + EPI_F32_A, lane, v);
                else stage_store(stg, zf, v, m0 + 32 * q, nb, M, N, lane, false);
            }
            if (act == UHC_ACT_GELU) {        // act(0) = 0 for every supported activation except sigmoid, so padded entries stay 0
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
            } else if (act != UHC_ACT_NONE) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = (row < M && nb + j < N) ? act_f(v[j], act) : 0.f;
            }
            if (yf) {
                if (tma_mask & 2) stage_row_f32_sw128(blk_s + EPI_F32_B, lane, v);
                else stage_store(stg, yf, v, m0 + 32 * q, nb, M, N, lane, ksplit > 1);
            }
            if ((tma_mask & 8) && nb < N) stage_col_bf16_sw64(blk_s + EPI_F32_B, lane, v);     // (the second fp32 area is free: a transposed copy is only asked for next to a bf16 y)
            if (ybf && nb < ldy) {
                if (tma_mask & 4) stage_row_bf16_sw64(blk_s + EPI_BF16, lane, v);
                else if (row < M) {
                    if (nb + 32 <= ldy) {
                        uint4 *dst = (uint4 *)(ybf + (size_t)row * ldy + nb);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]), p1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
                            __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]), p3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
                            uint4 u; u.x = *(uint32_t *)&p0; u.y = *(uint32_t *)&p1; u.z = *(uint32_t *)&p2; u.w = *(uint32_t *)&p3;
                            dst[j] = u;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (nb + j < ldy) ybf[(size_t)row * ldy + nb + j] = __float2bfloat16_rn(v[j]);      // (static indices: v stays in registers)
                    }
                }
            }
            if (tma_mask) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this lane's tile rows -> visible to the async proxy
                __syncwarp();
                if (lane == 0) {
                    const int r0 = m0 + 32 * q;
                    if (zf && (tma_mask & 1)) tma_store_2d(&mapZ, blk_s + EPI_F32_A, nb, r0);
                    if (yf && (tma_mask & 2)) tma_store_2d(&mapYf, blk_s + EPI_F32_B, nb, r0);
                    if (ybf && (tma_mask & 4) && nb < ldy) tma_store_2d(&mapYb, blk_s + EPI_BF16, nb, r0);
                    if ((tma_mask & 8) && nb < N) tma_store_2d(&mapYT, blk_s + EPI_F32_B, r0, nb);          // box = 32 rows (n) x 32 m-values
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (tma_mask && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the staging tiles have been read before the CTA's shared memory goes away (the global writes drain on their own)
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(BN) : "memory");
    }
}


// ---------------------------------------------------------------------------------------------------------------------------------------------------------
// k_linear_tc2: the same Linear / fused-backward GEMMs on CTA PAIRS (tcgen05 cta_group::2), persistent, with two accumulators in TMEM.
//   * a cluster of two CTAs (one per SM of a TPC) owns a 256 x 256 output tile: each CTA stages ITS 128 rows of A and ITS 128 of the tile's 256 B rows per k-block
//     (32 KB / stage / CTA for twice the flops of the single-CTA 128 x 128 tile: the operand traffic L2 -> shared memory per flop halves -- the single-CTA
//     kernel's main loop is bound by exactly that traffic); the leader CTA's elected thread issues UMMA 256 x 256 x 16 for the pair;
//   * persistent: each pair walks tiles pair, pair + npairs, ...; the accumulator is double buffered (2 x 256 TMEM columns), so the main loop of tile t + 1 runs
//     under the epilogue of tile t; the epilogue staging has its own shared memory (the operand ring never idles);
//   * 16 epilogue warps (four per TMEM lane quarter, 64 of the tile's 256 columns each): the epilogue is instruction-latency bound, so it gets the warps;
//   * every output leaves through the TMA engine; DACT as in k_linear_tc, with both z tiles of a warp prefetched before the accumulator is ready.
// Barriers: full[s] (leader's: both CTAs' TMA bytes), empty[s] (each CTA's: multicast commit), tmem_full[a] (each CTA's: multicast commit),
// tmem_empty[a] (leader's: one arrive per epilogue warp of both CTAs), z[warp] (DACT).
constexpr int P_STAGES = 5, P_TILE_N = 256, P_EPI_WARPS = 16, P_THREADS = (2 + P_EPI_WARPS) * 32, P_BLOCK = 4096;
constexpr int P_EPI_OFF = P_STAGES * STAGE_BYTES, P_BARS_OFF = P_EPI_OFF + P_EPI_WARPS * P_BLOCK;
constexpr int P_SMEM_BYTES = P_BARS_OFF + 512 + 1024;
// staging block of one epilogue warp: 4 KB = one fp32 32 x 32 tile or two bf16 ones.  Forward: z (fp32), then -- after the TMA engine has read it -- y / y^T
// (bf16) or the fp32 y.  DACT: the chunk's z tile lands here and, once in registers, makes room for dz^T and dz; the next chunk's z tile follows their store.
// (the small blocks buy a 5-stage operand ring: 160 KB in flight per CTA is what the TMA latency needs at this tile's consumption rate)
static_assert(STAGE_BYTES == 32768 && BM == 128 && BN == 128, "the pair kernel reuses the single-CTA operand boxes: 128 rows of A, 128 rows of B per CTA and k-block");
static_assert(P_SMEM_BYTES <= 227 * 1024, "pair kernel shared memory");
static_assert((2 * P_STAGES + 4 + P_EPI_WARPS) * 8 + 4 <= 512, "pair kernel barrier area");

__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {      // global[tile] += shared tile (element type from the map: fp32)
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {      // arrives on the barrier at this offset in BOTH CTAs when the MMAs issued so far retire
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((unsigned short)3) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
#define UHC_LDTM32(v, taddr)                                                                                                                                  \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                                                    \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]), "=f"(v[10]),      \
                   "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]),          \
                   "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31]) \
                 : "r"(taddr));                                                                                                                               \
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]), "+f"(v[16]), "+f"(v[17]), "+f"(v[18]), "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]), "+f"(v[24]), "+f"(v[25]), "+f"(v[26]), "+f"(v[27]), "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31]) :: "memory")

template <bool DACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
k_linear_tc2(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const float *__restrict__ bias, int M, int N, int Kp, int ldy, int act,
             const __grid_constant__ CUtensorMap mapZ, const __grid_constant__ CUtensorMap mapYf, const __grid_constant__ CUtensorMap mapYb,
             const __grid_constant__ CUtensorMap mapYT, int out_mask, float *__restrict__ dbias, int ksplit) {
    // ksplit > 1 (dW = dz^T h: few output tiles, a very long reduction): a work item is (tile, slice of the k-blocks); the fp32 y tiles are ADDED to the zeroed
    // output by the TMA engine (cp.reduce.async.bulk.tensor .add) -- only out_mask == 2 and no activation then.
    // out_mask (forward): bit 0 = fp32 pre-activation z (mapZ), bit 1 = fp32 y (mapYf; then no bf16 outputs: the host sends those shapes to k_linear_tc),
    // bit 2 = bf16 y (mapYb, ldy columns incl. padding), bit 3 = bf16 y^T (mapYT).
    // DACT: mapZ is the INPUT z_prev; outputs bf16 dz (mapYb), dz^T (mapYT), bias gradient sums (dbias).
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = (uint64_t *)(smem + P_BARS_OFF);              // full[S] | empty[S] | tmem_full[2] | tmem_empty[2] | zbar[warps]
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * P_STAGES + 4 + P_EPI_WARPS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    // work item w = tile * ksplit + slice; "ntiles" counts work items, w / ksplit is the tile every index computation below uses
    const int tiles_n = (N + P_TILE_N - 1) / P_TILE_N, ntiles = ((M + 2 * BM - 1) / (2 * BM)) * tiles_n * ksplit, nkb_all = Kp / BK;
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + P_STAGES), tfull0 = smem_u32(bars + 2 * P_STAGES), tempty0 = smem_u32(bars + 2 * P_STAGES + 2);
    const uint32_t zbar0 = smem_u32(bars + 2 * P_STAGES + 4);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB));
        for (int s = 0; s < P_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; a++) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 2 * P_EPI_WARPS); }
        for (int w = 0; w < P_EPI_WARPS; w++) mbar_init(zbar0 + 8 * w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {      // the same warp of both CTAs: one allocation of all 512 columns in each SM of the pair
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                            // the peer's barriers exist before anything can arrive on them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;                                            // k-blocks issued so far (ring position)
            uint32_t full_leader0;                                 // the leader CTA's full[] in the cluster's shared window
            asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(full_leader0) : "r"(full0));
            for (int w = pair; w < ntiles; w += npairs) {
                const int t = w / ksplit, sl = w - t * ksplit;
                const int kb0 = (int)(((long)nkb_all * sl) / ksplit), kb1 = (int)(((long)nkb_all * (sl + 1)) / ksplit);
                const int m0 = (t / tiles_n) * 2 * BM + (int)rank * BM, nrow0 = (t % tiles_n) * P_TILE_N + (int)rank * BN;
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % P_STAGES; const uint32_t ph = (it / P_STAGES) & 1;
                    mbar_wait(empty0 + 8 * s, ph ^ 1);
                    if (rank == 0) mbar_expect_tx(full0 + 8 * s, 2 * STAGE_BYTES);          // the pair's bytes land on the leader's barrier
                    const uint32_t a = smem_u32(smem + s * STAGE_BYTES), b = a + BM * BK * 2, lb = full_leader0 + 8 * s;
                    tma_load_2d_pair(a, &mapA, lb, kb * BK, m0);
                    tma_load_2d_pair(b, &mapB, lb, kb * BK, nrow0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            // instruction descriptor: D = F32, A = B = BF16, both K-major, N = 256, M = 256 (128 rows per CTA)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P_TILE_N >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
            int it = 0, ti = 0;
            for (int w = pair; w < ntiles; w += npairs, ++ti) {
                const int sl = w % ksplit, nkb = (int)(((long)nkb_all * (sl + 1)) / ksplit) - (int)(((long)nkb_all * sl) / ksplit);
                const int acc = ti & 1; const uint32_t aph = (ti >> 1) & 1;
                mbar_wait(tempty0 + 8 * acc, aph ^ 1);                                      // both CTAs' epilogues have drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem + (uint32_t)(acc * P_TILE_N);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % P_STAGES; const uint32_t ph = (it / P_STAGES) & 1;
                    mbar_wait(full0 + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a = smem_u32(smem + s * STAGE_BYTES), b = a + BM * BK * 2;
                    const uint64_t ad = make_desc(a), bd = make_desc(b);
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k) umma_bf16_pair(tacc, ad + (uint64_t)(k * UK * 2 >> 4), bd + (uint64_t)(k * UK * 2 >> 4), idesc, (kb | k) != 0);
                    umma_commit_pair(empty0 + 8 * s);
                }
                umma_commit_pair(tfull0 + 8 * acc);
            }
        }
    } else {
        const int q = warp & 3, ew = warp - 2, colq = ew >> 2;          // TMEM lane quarter (fixed by the hardware: warp id % 4); which 64 of the tile's 256 columns
        const uint32_t blk_s = smem_u32(smem + P_EPI_OFF + ew * P_BLOCK), zb = zbar0 + 8 * ew;
        uint32_t tempty_leader0;
        asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(tempty_leader0) : "r"(tempty0));
        int ti = 0;
        uint32_t zuse = 0;                                 // completed uses of this warp's z barrier (a chunk outside the matrix has neither a load nor a wait)
        if constexpr (DACT) {
            if (lane == 0 && pair < ntiles) {              // the first chunk's z tile: in flight before the first accumulator exists (DACT never splits k)
                const int r0 = (pair / tiles_n) * 2 * BM + (int)rank * BM + 32 * q, n0 = (pair % tiles_n) * P_TILE_N + 64 * colq;
                if (r0 < M && n0 < N) { mbar_expect_tx(zb, 4096u); tma_load_2d(blk_s, &mapZ, zb, n0, r0); }
            }
        }
        for (int w = pair; w < ntiles; w += npairs, ++ti) {
            const int t = w / ksplit;
            const bool first_slice = w == t * ksplit;
            const int acc = ti & 1; const uint32_t aph = (ti >> 1) & 1;
            const int r0 = (t / tiles_n) * 2 * BM + (int)rank * BM + 32 * q, n0 = (t % tiles_n) * P_TILE_N + 64 * colq;
            const int row = r0 + lane;
            mbar_wait(tfull0 + 8 * acc, aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                float v[32];
                const int nb = n0 + 32 * cc;
                const bool interior = r0 + 32 <= M && nb + 32 <= N;      // (warp uniform) no element of this chunk needs masking
                const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * P_TILE_N + 64 * colq + 32 * cc);
                UHC_LDTM32(v, taddr);
                if (cc == 1) {                          // the accumulator is in registers: hand it back to the MMA warp of the leader
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    // (relaxed: the TMEM reads are ordered by the tcgen05 fence above; a release at cluster scope would drain every outstanding global access of the lane)
                    if (lane == 0) asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(tempty_leader0 + 8 * acc) : "memory");
                }
                if constexpr (DACT) {
                    if (r0 < M && nb < N) {
                        mbar_wait(zb, zuse & 1); ++zuse;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float a, b, cq, d;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(cq), "=f"(d) : "r"(blk_s + lane * 128 + ((j ^ (lane & 7)) << 4)) : "memory");
                            if (act == UHC_ACT_GELU) {       // (the common case without the per-element switch)
                                v[4 * j] *= act_b_fast(a, UHC_ACT_GELU); v[4 * j + 1] *= act_b_fast(b, UHC_ACT_GELU); v[4 * j + 2] *= act_b_fast(cq, UHC_ACT_GELU); v[4 * j + 3] *= act_b_fast(d, UHC_ACT_GELU);
                            } else {
                                v[4 * j] *= act_b_fast(a, act); v[4 * j + 1] *= act_b_fast(b, act); v[4 * j + 2] *= act_b_fast(cq, act); v[4 * j + 3] *= act_b_fast(d, act);
                            }
                        }
                    }
                    if (!interior) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = (row < M && nb + j < N) ? v[j] : 0.f;
                    }
                    __syncwarp();                       // every lane has its z row in registers: the block now stages this chunk's dz^T and dz
                    if (nb < N) stage_col_bf16_sw64(blk_s, lane, v);
                    if (nb < ldy) stage_row_bf16_sw64(blk_s + 2048, lane, v);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        if (nb < N) tma_store_2d(&mapYT, blk_s, r0, nb);
                        if (nb < ldy) tma_store_2d(&mapYb, blk_s + 2048, nb, r0);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");          // the block is free again: fetch the NEXT chunk's z tile under this chunk's tail
                        int nr0 = r0, nnb = nb + 32;
                        if (cc == 1) { const int tn = w + npairs; nr0 = tn < ntiles ? (tn / tiles_n) * 2 * BM + (int)rank * BM + 32 * q : M; nnb = (tn % tiles_n) * P_TILE_N + 64 * colq; }
                        if (nr0 < M && nnb < N) { mbar_expect_tx(zb, 4096u); tma_load_2d(blk_s, &mapZ, zb, nnb, nr0); }
                    }
                    if (dbias && nb < N) {              // column sums over the warp's 32 rows: transpose-reduce (31 shuffles), lane l ends with column l
#pragma unroll
                        for (int o = 16; o >= 1; o >>= 1) {
                            const bool up = (lane & o) != 0;
#pragma unroll
                            for (int k = 0; k < o; ++k) {
                                const float send = up ? v[k] : v[k + o], keep = up ? v[k + o] : v[k];
                                v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                            }
                        }
                        if (nb + lane < N) atomicAdd(dbias + nb + lane, v[0]);
                    }
                    __syncwarp();
                } else {
                    if (bias && first_slice) {
                        if (interior && (((uintptr_t)(bias + nb)) & 15) == 0) {      // 8 uniform 16-byte loads instead of 32 shuffles
#pragma unroll
                            for (int j = 0; j < 8; ++j) { const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias + nb) + j); v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w; }
                        } else {
                            const float bl = nb + lane < N ? __ldg(bias + nb + lane) : 0.f;
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, bl, j);
                        }
                    }
                    if (!interior) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = (row < M && nb + j < N) ? v[j] : 0.f;
                    }
                    if ((out_mask & 1) && nb < N) {     // phase 1: the fp32 pre-activation through the warp's 4 KB block
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        __syncwarp();
                        stage_row_f32_sw128(blk_s, lane, v);
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) { tma_store_2d(&mapZ, blk_s, nb, r0); asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
                    }
                    if (act == UHC_ACT_GELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
                    } else if (act != UHC_ACT_NONE) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = act_f(v[j], act);
                        if (!interior) {                 // (sigmoid(0) != 0)
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = (row < M && nb + j < N) ? v[j] : 0.f;
                        }
                    }
                    if (out_mask & 14) {                // phase 2: fp32 y, or bf16 y (first 2 KB) and y^T (second 2 KB), once the block has been read
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        __syncwarp();
                        if (out_mask & 2) { if (nb < N) stage_row_f32_sw128(blk_s, lane, v); }
                        else {
                            if ((out_mask & 4) && nb < ldy) stage_row_bf16_sw64(blk_s, lane, v);
                            if ((out_mask & 8) && nb < N) stage_col_bf16_sw64(blk_s + 2048, lane, v);
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
                            if (out_mask & 2) { if (nb < N) { if (ksplit > 1) tma_reduce_add_2d(&mapYf, blk_s, nb, r0); else tma_store_2d(&mapYf, blk_s, nb, r0); } }
                            else {
                                if ((out_mask & 4) && nb < ldy) tma_store_2d(&mapYb, blk_s, nb, r0);
                                if ((out_mask & 8) && nb < N) tma_store_2d(&mapYT, blk_s + 2048, r0, nb);
                            }
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    }
                }
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                            // both CTAs are done with TMEM and with each other's barriers
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

__global__ void k_f32_to_bf16_padded(const float *__restrict__ x, __nv_bfloat16 *__restrict__ y, int M, int K, int Kp) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)M * Kp; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % Kp); const size_t r = i / Kp;
        y[i] = __float2bfloat16_rn(k < K ? x[r * K + k] : 0.f);
    }
}


// d/dz gelu(z) = Phi(z) + z phi(z), Phi through the same A&S erf as gelu_fast (shares the exp)
__device__ __forceinline__ float dgelu_fast(float z) {
    const float x = fabsf(z) * 0.70710678118654752f;
    const float t = __fdividef(1.0f, fmaf(0.3275911f, x, 1.0f));
    const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
    const float ex = __expf(-x * x);                       // exp(-z^2 / 2)
    const float e = 1.0f - poly * ex;                      // erf(|z| / sqrt 2)
    return 0.5f + copysignf(0.5f * e, z) + z * 0.3989422804014327f * ex;
}
__device__ __forceinline__ float act_b(float z, int act) {
    switch (act) {
    case UHC_ACT_GELU: return dgelu_fast(z);
    case UHC_ACT_TANH: { float t = tanhf(z); return 1.0f - t * t; }
    case UHC_ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case UHC_ACT_SIGMOID: { float s = 1.0f / (1.0f + expf(-z)); return s * (1.0f - s); }
    }
    return 1.f;
}
// out[c][r] = in[r][c]  (bf16, 64x64 tiles through shared memory); rows r >= R of `in` read as zero so out is zero padded to ld_out
__global__ void k_transpose_bf16(const __nv_bfloat16 *__restrict__ in, __nv_bfloat16 *__restrict__ out, int R, int Cc, int ld_in, int ld_out) {
    __shared__ __nv_bfloat16 t[64][66];
    const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
    for (int i = threadIdx.y; i < 64; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        t[i][threadIdx.x] = (r < R && c < Cc) ? in[(size_t)r * ld_in + c] : __float2bfloat16_rn(0.f);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 64; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < Cc && r < ld_out) out[(size_t)c * ld_out + r] = t[threadIdx.x][i];
    }
}
// vectorised variants (4 elements per thread along the contiguous dimension, all loads of a thread issued before use): used when the
// row lengths are multiples of 4
__device__ __forceinline__ uint2 pack_bf16x4(float a, float b, float c, float d) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a, b), p1 = __floats2bfloat162_rn(c, d);
    uint2 u; u.x = *(uint32_t *)&p0; u.y = *(uint32_t *)&p1; return u;
}
__global__ void __launch_bounds__(256) k_transpose_bf16_v4(const __nv_bfloat16 *__restrict__ in, __nv_bfloat16 *__restrict__ out, int R, int Cc, int ld_in, int ld_out) {
    __shared__ float t[64][65];
    const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64, tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    uint2 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = r0 + ty + 16 * k, c = c0 + 4 * tx;
        v[k] = (r < R && c < Cc) ? *reinterpret_cast<const uint2 *>(in + (size_t)r * ld_in + c) : make_uint2(0u, 0u);   // Cc % 4 == 0: a group is all in or all out
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 a = *(__nv_bfloat162 *)&v[k].x, b = *(__nv_bfloat162 *)&v[k].y;
        float *row = t[ty + 16 * k] + 4 * tx;
        row[0] = __low2float(a); row[1] = __high2float(a); row[2] = __low2float(b); row[3] = __high2float(b);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 16 * k, r = r0 + 4 * tx;
        if (c < Cc && r < ld_out)
            *reinterpret_cast<uint2 *>(out + (size_t)c * ld_out + r) = pack_bf16x4(t[4 * tx][ty + 16 * k], t[4 * tx + 1][ty + 16 * k], t[4 * tx + 2][ty + 16 * k], t[4 * tx + 3][ty + 16 * k]);
    }
}
__global__ void __launch_bounds__(256) k_dact_bf16_v4(const float *__restrict__ dh, const float *__restrict__ z, __nv_bfloat16 *__restrict__ dz, __nv_bfloat16 *__restrict__ dzT,
                                                     float *__restrict__ db, int M, int N, int ld_dz, int ld_dzT, int act) {
    __shared__ float t[64][65];
    __shared__ float cs[64];
    const int n0 = blockIdx.x * 64, m0 = blockIdx.y * 64, tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    if (threadIdx.x < 64) cs[threadIdx.x] = 0.f;
    float4 g[4], zz[4];
    const int n = n0 + 4 * tx;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int m = m0 + ty + 16 * k;
        const bool in = m < M && n < N;
        g[k] = in ? *reinterpret_cast<const float4 *>(dh + (size_t)m * N + n) : make_float4(0.f, 0.f, 0.f, 0.f);
        zz[k] = (in && z) ? *reinterpret_cast<const float4 *>(z + (size_t)m * N + n) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int m = m0 + ty + 16 * k;
        float4 v = g[k];
        if (z) { v.x *= act_b(zz[k].x, act); v.y *= act_b(zz[k].y, act); v.z *= act_b(zz[k].z, act); v.w *= act_b(zz[k].w, act); }
        float *row = t[ty + 16 * k] + 4 * tx;
        row[0] = v.x; row[1] = v.y; row[2] = v.z; row[3] = v.w;
        c0 += v.x; c1 += v.y; c2 += v.z; c3 += v.w;
        if (dz && m < M && n < ld_dz) *reinterpret_cast<uint2 *>(dz + (size_t)m * ld_dz + n) = pack_bf16x4(v.x, v.y, v.z, v.w);
    }
    if (db) { atomicAdd(&cs[4 * tx], c0); atomicAdd(&cs[4 * tx + 1], c1); atomicAdd(&cs[4 * tx + 2], c2); atomicAdd(&cs[4 * tx + 3], c3); }
    __syncthreads();
    if (db && threadIdx.x < 64 && n0 + threadIdx.x < N) atomicAdd(db + n0 + threadIdx.x, cs[threadIdx.x]);
    if (dzT) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int nn = n0 + ty + 16 * k, m = m0 + 4 * tx;
            if (nn < N && m < ld_dzT)
                *reinterpret_cast<uint2 *>(dzT + (size_t)nn * ld_dzT + m) = pack_bf16x4(t[4 * tx][ty + 16 * k], t[4 * tx + 1][ty + 16 * k], t[4 * tx + 2][ty + 16 * k], t[4 * tx + 3][ty + 16 * k]);
        }
    }
}
// dz = dh * act'(z): writes dz (bf16, [M][ld_dz]) and its transpose ([N][ld_dzT], zero padded in M) and accumulates column sums (bias grads)
__global__ void k_dact_bf16(const float *__restrict__ dh, const float *__restrict__ z, __nv_bfloat16 *__restrict__ dz, __nv_bfloat16 *__restrict__ dzT,
                            float *__restrict__ db, int M, int N, int ld_dz, int ld_dzT, int act) {
    __shared__ float t[64][65];
    const int n0 = blockIdx.x * 64, m0 = blockIdx.y * 64;
    float colsum = 0.f;
    for (int i = threadIdx.y; i < 64; i += 8) {
        const int m = m0 + i, n = n0 + threadIdx.x;
        float v = 0.f;
        if (m < M && n < N) { v = dh[(size_t)m * N + n]; if (z) v *= act_b(z[(size_t)m * N + n], act); }
        t[i][threadIdx.x] = v;
        colsum += v;
        if (dz && m < M && n < ld_dz) dz[(size_t)m * ld_dz + n] = __float2bfloat16_rn(v);
    }
    if (db && n0 + threadIdx.x < N) atomicAdd(db + n0 + threadIdx.x, colsum);
    __syncthreads();
    if (dzT) for (int i = threadIdx.y; i < 64; i += 8) {
        const int n = n0 + i, m = m0 + threadIdx.x;
        if (n < N && m < ld_dzT) dzT[(size_t)n * ld_dzT + m] = __float2bfloat16_rn(t[threadIdx.x][i]);
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode() {
    static EncodeFn fn = nullptr;
    if (!fn) {
        void *p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeFn)p;
    }
    return fn;
}
int make_map(CUtensorMap *m, const void *base, int rows, int Kp, int box_rows = BM) {  // row-major [rows][Kp] bf16, box 64 x 128, 128B swizzle
    EncodeFn enc = get_encode();
    if (!enc) { g_tc_err = "cuTensorMapEncodeTiled unavailable"; return -1; }
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)Kp * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows}, estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { g_tc_err = "cuTensorMapEncodeTiled failed: " + std::to_string((int)r); return -1; }
    return 0;
}
// output tensor map: row-major [rows][cols] with a row pitch of `pitch` bytes, 32 x 32 boxes; the staging tiles use the 128-byte (fp32) / 64-byte (bf16) swizzle
int make_map_out(CUtensorMap *m, const void *base, int rows, int cols, size_t pitch, bool bf16) {
    EncodeFn enc = get_encode();
    if (!enc) { g_tc_err = "cuTensorMapEncodeTiled unavailable"; return -1; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {32, 32}, estr[2] = {1, 1};
    CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     bf16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { g_tc_err = "cuTensorMapEncodeTiled (output) failed: " + std::to_string((int)r); return -1; }
    return 0;
}
bool tma_store_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("UHC_TC_TMA_STORE"); on = e ? (e[0] != '0') : (UHC_TC_TMA_STORE != 0); }
    return on != 0;
}
// the CTA-pair kernel takes the big training GEMMs (many 256 x 256 tiles, every output through a tensor map); UHC_TC_PAIR=0 keeps everything on k_linear_tc
bool pair_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("UHC_TC_PAIR"); on = e ? (e[0] != '0') : 1; }
    return on != 0;
}
int pair_min_rows() {      // smallest M the pair kernel takes (UHC_TC_PAIR_MINM): below it the 128 x 128 tiles of k_linear_tc spread over more SMs
    static int v = -1;
    if (v < 0) { const char *e = getenv("UHC_TC_PAIR_MINM"); v = e ? atoi(e) : 16384; if (v < 256) v = 256; }
    return v;
}
int pair_probe();
bool pair_shape(int M, int N) { return pair_enabled() && M >= pair_min_rows() && N >= 256 && pair_probe() == 1; }
// per device: 0 = not probed, 1 = the pair kernel can run (attribute set, at least one 2-CTA cluster with its shared memory fits), 2 = it cannot (then every shape
// stays on k_linear_tc: a partitioned or smaller GPU is a slower path, not an error)
int g_pair_state[64] = {0};
int pair_probe() {
    int dev = 0; cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 2;
    if (g_pair_state[dev]) return g_pair_state[dev];
    int st = 1;
    if (cudaFuncSetAttribute(k_linear_tc2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_linear_tc2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES) != cudaSuccess) st = 2;
    if (st == 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2, 1, 1); cfg.blockDim = dim3(P_THREADS, 1, 1); cfg.dynamicSmemBytes = P_SMEM_BYTES;
        cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, k_linear_tc2<false>, &cfg) != cudaSuccess || nclusters < 1) st = 2;
    }
    (void)cudaGetLastError();
    g_pair_state[dev] = st;
    return st;
}
int pair_attr() { if (pair_probe() != 1) { g_tc_err = "the CTA-pair kernel cannot run on this device"; return -1; } return 0; }
int pair_grid(int M, int N) {
    int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int tiles = ((M + 255) / 256) * ((N + 255) / 256), pairs = sms / 2;
    return 2 * (tiles < pairs ? tiles : pairs);
}
}  // namespace

extern "C" {
const char *uhc_tc_last_error(void) { return g_tc_err.c_str(); }

int uhc_f32_to_bf16_padded(const float *x, void *y_bf16, int M, int K, int Kp, void *stream) {
    k_f32_to_bf16_padded<<<1184, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16 *)y_bf16, M, K, Kp);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

static int linear_tc_impl(const void *x_bf16, const void *W_bf16, const float *b, void *y_bf16_or_null, float *y_f32_or_null, float *z_f32_or_null,
                          int M, int N, int Kp, int ldy_bf16, int act, void *stream, void *yT_bf16_or_null = nullptr, int ld_yT = 0, int ld_yf = 0) {
    if (ld_yf <= 0) ld_yf = N;      // row pitch of the fp32 y in floats (only the long-reduction pair path takes a pitch other than N)
    if (Kp % BK != 0 || M <= 0 || N <= 0) { g_tc_err = "uhc_linear_forward_tc: Kp must be a positive multiple of 64"; return -2; }
    if (y_bf16_or_null && (ldy_bf16 % 8 != 0)) { g_tc_err = "uhc_linear_forward_tc: ldy must be a multiple of 8"; return -2; }
    static bool attr_set[64] = {false};   // per device: the attribute belongs to the function on the CURRENT device
    int dev = 0; cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        if (cudaFuncSetAttribute(k_linear_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(k_linear_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) { g_tc_err = "cudaFuncSetAttribute failed"; return -1; }
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    CUtensorMap ma, mb;
    if (make_map(&ma, x_bf16, M, Kp, BM) || make_map(&mb, W_bf16, N, Kp, BN)) return -1;
    const int tiles = ((N + BN - 1) / BN) * ((M + BM - 1) / BM), nkb = Kp / BK;
    // a plain fp32 product with a long reduction and few output tiles (dW = dz^T h): CTA pairs, the reduction split into as many slices as fill the pairs evenly,
    // partial tiles added into the zeroed output by the TMA engine
    if (pair_enabled() && pair_probe() == 1 && tma_store_enabled() && !y_bf16_or_null && !z_f32_or_null && !yT_bf16_or_null && act == UHC_ACT_NONE && y_f32_or_null && M >= 256 && N >= 256 &&
        nkb >= 256 && ((uintptr_t)y_f32_or_null & 15) == 0 && ld_yf % 4 == 0 && ld_yf >= N) {
        int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int pairs = sms / 2, pt = ((M + 255) / 256) * ((N + 255) / 256);
        int ks = 1; double best = 0.0;
        for (int c = 1; c <= 16 && c <= nkb / 32; ++c) {
            const int items = pt * c; const double eff = (double)items / (double)(((items + pairs - 1) / pairs) * pairs);
            if (eff > best + 0.02) { best = eff; ks = c; }
        }
        CUtensorMap my;
        if (make_map_out(&my, y_f32_or_null, M, N, (size_t)ld_yf * sizeof(float), false) || pair_attr()) return -1;
        if ((ks > 1 || ld_yf != N) && cudaMemsetAsync(y_f32_or_null, 0, (size_t)M * ld_yf * sizeof(float), (cudaStream_t)stream) != cudaSuccess) { g_tc_err = "memset failed"; return -1; }   // (a padded pitch: the padding columns read as zeros)
        const int items = pt * ks;
        k_linear_tc2<false><<<2 * (items < pairs ? items : pairs), P_THREADS, P_SMEM_BYTES, (cudaStream_t)stream>>>(ma, mb, b, M, N, Kp, 0, UHC_ACT_NONE, ma, my, ma, ma, 2, nullptr, ks);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { g_tc_err = cudaGetErrorString(e); return -1; }
        return 0;
    }
    if (ld_yf != N) { g_tc_err = "uhc_linear_forward_tc_f32_pitched: this shape does not run on the CTA-pair path (needs M, N >= 256, Kp >= 16384, a pitch that is a multiple of 4)"; return -2; }
    // split the reduction when the output has too few tiles to fill the GPU (only legal for a plain fp32 accumulate output)
    int ksplit = 1;
    if (!y_bf16_or_null && !z_f32_or_null && act == UHC_ACT_NONE && y_f32_or_null && tiles < 148 && nkb >= 32) {
        ksplit = (296 + tiles - 1) / tiles;
        if (ksplit > nkb / 8) ksplit = nkb / 8;
        if (ksplit < 1) ksplit = 1;
        if (ksplit > 1 && cudaMemsetAsync(y_f32_or_null, 0, (size_t)M * N * sizeof(float), (cudaStream_t)stream) != cudaSuccess) { g_tc_err = "memset failed"; return -1; }
    }
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, ksplit);
    // outputs through the TMA engine: a tensor map needs a 16-byte aligned base and row pitch.  The fp32 outputs share one staging decision (their legacy
    // 32 x 33 tile spans both fp32 areas of the warp's block); split-K accumulation keeps the atomic path.
    CUtensorMap mz = ma, myf = ma, myb = ma;
    int tma_mask = 0;
    if (tma_store_enabled()) {
        auto ok = [](const void *p, size_t pitch) { return p && ((uintptr_t)p & 15) == 0 && pitch % 16 == 0; };
        const size_t pf = (size_t)N * sizeof(float), pb = (size_t)ldy_bf16 * 2;
        const bool f32_ok = ksplit == 1 && (z_f32_or_null || y_f32_or_null) && (!z_f32_or_null || ok(z_f32_or_null, pf)) && (!y_f32_or_null || ok(y_f32_or_null, pf));
        if (f32_ok) {
            if (z_f32_or_null) { if (make_map_out(&mz, z_f32_or_null, M, N, pf, false)) return -1; tma_mask |= 1; }
            if (y_f32_or_null) { if (make_map_out(&myf, y_f32_or_null, M, N, pf, false)) return -1; tma_mask |= 2; }
        }
        if (y_bf16_or_null && ok(y_bf16_or_null, pb)) { if (make_map_out(&myb, y_bf16_or_null, M, ldy_bf16, pb, true)) return -1; tma_mask |= 4; }
    }
    CUtensorMap myt = ma;
    if (yT_bf16_or_null) {      // transposed activation [N][ld_yT] (ld_yT >= M rounded up to 64, the padding columns receive the tile's zero rows)
        if (!tma_store_enabled() || y_f32_or_null || ((uintptr_t)yT_bf16_or_null & 15) || ld_yT % 8 != 0 || ld_yT < M) { g_tc_err = "uhc_linear_forward_tc_train_t: the transposed output needs the TMA store path, no fp32 y, and a pitch >= M that is a multiple of 8"; return -2; }
        if (make_map_out(&myt, yT_bf16_or_null, N, ld_yT, (size_t)ld_yT * 2, true)) return -1;
        tma_mask |= 8;
    }
    {   // every requested output has a tensor map and the shape is big: CTA pairs
        const int want = (z_f32_or_null ? 1 : 0) | (y_f32_or_null ? 2 : 0) | (y_bf16_or_null ? 4 : 0) | (yT_bf16_or_null ? 8 : 0);
        if (ksplit == 1 && want == tma_mask && !((want & 2) && (want & 12)) && pair_shape(M, N)) {
            if (pair_attr()) return -1;
            k_linear_tc2<false><<<pair_grid(M, N), P_THREADS, P_SMEM_BYTES, (cudaStream_t)stream>>>(ma, mb, b, M, N, Kp, ldy_bf16, act, mz, myf, myb, myt, tma_mask, nullptr, 1);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { g_tc_err = cudaGetErrorString(e); return -1; }
            return 0;
        }
    }
    k_linear_tc<false><<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(ma, mb, b, (__nv_bfloat16 *)y_bf16_or_null, y_f32_or_null, z_f32_or_null, M, N, Kp, ldy_bf16, act, ksplit,
                                                                            mz, myf, myb, tma_mask, myt, nullptr);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_tc_err = cudaGetErrorString(e); return -1; }
    return 0;
}

int uhc_linear_forward_tc(const void *x_bf16, const void *W_bf16, const float *b, void *y_bf16_or_null, float *y_f32_or_null, int M, int N, int Kp,
                          int ldy_bf16, int act, void *stream) {
    return linear_tc_impl(x_bf16, W_bf16, b, y_bf16_or_null, y_f32_or_null, nullptr, M, N, Kp, ldy_bf16, act, stream);
}
int uhc_linear_forward_tc_train(const void *x_bf16, const void *W_bf16, const float *b, void *y_bf16_or_null, float *y_f32_or_null, float *z_f32,
                                int M, int N, int Kp, int ldy_bf16, int act, void *stream) {
    return linear_tc_impl(x_bf16, W_bf16, b, y_bf16_or_null, y_f32_or_null, z_f32, M, N, Kp, ldy_bf16, act, stream);
}
/* training forward that also emits the transpose of the bf16 activation, yT [N][ld_yT] (zero padded to ld_yT): the K-major operand of the backward pass' dW GEMM,
 * written by the same epilogue through the TMA engine instead of by a separate transpose kernel */
int uhc_linear_forward_tc_train_t(const void *x_bf16, const void *W_bf16, const float *b, void *y_bf16, void *yT_bf16, int ld_yT, float *z_f32_or_null,
                                  int M, int N, int Kp, int ldy_bf16, int act, void *stream) {
    if (!y_bf16 || !yT_bf16) { g_tc_err = "uhc_linear_forward_tc_train_t: y and yT are required"; return -2; }
    return linear_tc_impl(x_bf16, W_bf16, b, y_bf16, nullptr, z_f32_or_null, M, N, Kp, ldy_bf16, act, stream, yT_bf16, ld_yT);
}
/* plain fp32 product y[M][ld_y] = x W^T with a row pitch ld_y >= N (floats, multiple of 4): lets an output whose own pitch no tensor map accepts (N % 4 != 0) be
 * computed on the CTA-pair path into a padded scratch.  Returns -2 when the shape does not qualify (then use uhc_linear_forward_tc). */
int uhc_linear_forward_tc_f32_pitched(const void *x_bf16, const void *W_bf16, float *y_f32, int ld_y, int M, int N, int Kp, void *stream) {
    if (!y_f32 || ld_y < N || ld_y % 4 != 0 || !pair_enabled() || pair_probe() != 1 || !tma_store_enabled() || M < 256 || N < 256 || Kp / BK < 256) { g_tc_err = "uhc_linear_forward_tc_f32_pitched: shape not eligible"; return -2; }
    return linear_tc_impl(x_bf16, W_bf16, nullptr, nullptr, y_f32, nullptr, M, N, Kp, 0, UHC_ACT_NONE, stream, nullptr, 0, ld_y);
}
int uhc_tc_tma_store_enabled(void) { return tma_store_enabled() ? 1 : 0; }
/* backward through one Linear + the previous layer's activation in ONE kernel:  dz_prev = (dz W) * act'(z_prev)  as bf16 [M][ld_dz] and transposed [K][ld_dzT],
 * db_prev[k] += sum_m dz_prev[m][k].  dz [M][Np] and WT [K][Np] are the K-major bf16 operands (Np = N rounded up to 64).  Needs the TMA-store path, K % 4 == 0
 * and 16-byte aligned buffers (returns -2 otherwise: run uhc_linear_forward_tc + uhc_dact_bf16 instead). */
int uhc_linear_dx_dact_tc(const void *dz_bf16, const void *WT_bf16, const float *z_prev, void *dzp_bf16, void *dzpT_bf16, float *db_prev_or_null,
                          int M, int K, int Np, int ld_dz, int ld_dzT, int act, void *stream) {
    if (Np % BK != 0 || M <= 0 || K <= 0) { g_tc_err = "uhc_linear_dx_dact_tc: Np must be a positive multiple of 64"; return -2; }
    auto al = [](const void *p) { return p && ((uintptr_t)p & 15) == 0; };
    if (!tma_store_enabled() || K % 4 != 0 || ld_dz % 8 != 0 || ld_dzT % 8 != 0 || ld_dz < K || ld_dzT < M || !al(z_prev) || !al(dzp_bf16) || !al(dzpT_bf16)) {
        g_tc_err = "uhc_linear_dx_dact_tc: needs the TMA store path, K % 4 == 0, pitches that are multiples of 8 elements and 16-byte aligned buffers"; return -2;
    }
    int dev = 0; cudaGetDevice(&dev);
    static bool attr_set[64] = {false};
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        if (cudaFuncSetAttribute(k_linear_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) { g_tc_err = "cudaFuncSetAttribute failed"; return -1; }
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    CUtensorMap ma, mb, mz, mdz, mdzT;
    if (make_map(&ma, dz_bf16, M, Np, BM) || make_map(&mb, WT_bf16, K, Np, BN)) return -1;
    if (make_map_out(&mz, z_prev, M, K, (size_t)K * 4, false) || make_map_out(&mdz, dzp_bf16, M, ld_dz, (size_t)ld_dz * 2, true) ||
        make_map_out(&mdzT, dzpT_bf16, K, ld_dzT, (size_t)ld_dzT * 2, true)) return -1;
    if (db_prev_or_null && cudaMemsetAsync(db_prev_or_null, 0, (size_t)K * sizeof(float), (cudaStream_t)stream) != cudaSuccess) { g_tc_err = "uhc_linear_dx_dact_tc: memset failed"; return -1; }
    if (pair_shape(M, K)) {
        if (pair_attr()) return -1;
        k_linear_tc2<true><<<pair_grid(M, K), P_THREADS, P_SMEM_BYTES, (cudaStream_t)stream>>>(ma, mb, nullptr, M, K, Np, ld_dz, act, mz, mz, mdz, mdzT, 0, db_prev_or_null, 1);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { g_tc_err = cudaGetErrorString(e); return -1; }
        return 0;
    }
    dim3 grid((K + BN - 1) / BN, (M + BM - 1) / BM, 1);
    k_linear_tc<true><<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(ma, mb, nullptr, nullptr, nullptr, nullptr, M, K, Np, ld_dz, act, 1, mz, mz, mdz, 0, mdzT, db_prev_or_null);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_tc_err = cudaGetErrorString(e); return -1; }
    return 0;
}
int uhc_transpose_bf16(const void *in, void *out, int R, int Cc, int ld_in, int ld_out, void *stream) {
    dim3 grid((Cc + 63) / 64, (R + 63) / 64);
    if (Cc % 4 == 0 && ld_in % 4 == 0 && ld_out % 4 == 0 && ((uintptr_t)in & 7) == 0 && ((uintptr_t)out & 7) == 0)
        k_transpose_bf16_v4<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)in, (__nv_bfloat16 *)out, R, Cc, ld_in, ld_out);
    else
        k_transpose_bf16<<<grid, dim3(64, 8), 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)in, (__nv_bfloat16 *)out, R, Cc, ld_in, ld_out);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
int uhc_dact_bf16(const float *dh, const float *z_or_null, void *dz_bf16, void *dzT_bf16, float *db_or_null, int M, int N, int ld_dz, int ld_dzT, int act,
                  void *stream) {
    if (db_or_null && cudaMemsetAsync(db_or_null, 0, N * sizeof(float), (cudaStream_t)stream) != cudaSuccess) { g_tc_err = "uhc_dact_bf16: memset failed"; return -1; }
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    const bool v4 = N % 4 == 0 && (!dz_bf16 || (ld_dz % 4 == 0 && ((uintptr_t)dz_bf16 & 7) == 0)) && (!dzT_bf16 || (ld_dzT % 4 == 0 && ((uintptr_t)dzT_bf16 & 7) == 0)) &&
                    ((uintptr_t)dh & 15) == 0 && (!z_or_null || ((uintptr_t)z_or_null & 15) == 0);
    if (v4) k_dact_bf16_v4<<<grid, 256, 0, (cudaStream_t)stream>>>(dh, z_or_null, (__nv_bfloat16 *)dz_bf16, (__nv_bfloat16 *)dzT_bf16, db_or_null, M, N, ld_dz, ld_dzT, act);
    else k_dact_bf16<<<grid, dim3(64, 8), 0, (cudaStream_t)stream>>>(dh, z_or_null, (__nv_bfloat16 *)dz_bf16, (__nv_bfloat16 *)dzT_bf16, db_or_null, M, N, ld_dz, ld_dzT, act);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
}
