// neg_cent_tc.cu -- Gaussian-prior score matrix on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   neg_cent[b,x,y] = rowterm[x] + sum_k A[x,k] * Bm[y,k]        (see neg_cent.cu for the k = 2c / 2c+1 operands)
//
// One CTA computes a 128 (text) x 256 (mel) tile.  The accumulator lives in tensor memory (256 fp32 columns x 128
// lanes); operands are produced by the CTA itself: the raw m_p / logs_p / z values are read in their native layouts
// (t_text, t_mel contiguous -> coalesced), transformed (exp(-2 logs), m s2, -0.5 z^2, z), split into a TF32 "hi" part
// and a TF32 "lo" remainder, and stored straight into the canonical K-major SWIZZLE_128B shared-memory layout that
// tcgen05.mma descriptors address (16-byte chunk index XOR (row % 8); 8-row groups 1024 bytes apart).
// fp32 accuracy comes from the 3xTF32 split:  A B ~= Ahi Bhi + Alo Bhi + Ahi Blo,  accumulated in fp32 in TMEM
// (the dropped Alo Blo term is 2^-22 relative).  A single TF32 pass would miss the 1e-5 bound by two orders.
//
// Pipeline per 16-channel K chunk (32 k-values = one 128-byte swizzled row per operand row), two stages:
//   all 256 threads : wait until the MMAs that read this stage two chunks ago have committed (mbarrier),
//                     produce Ahi/Alo (128 x 32) and Bhi/Blo (256 x 32), fence.proxy.async, __syncthreads
//   thread 0        : 12 x tcgen05.mma.cta_group::1.kind::tf32 (3 products x 4 k-steps of 8), tcgen05.commit -> mbarrier
// so chunk i+1 is produced on the CUDA cores while the tensor pipe works on chunk i.
// Epilogue: 8 warps read their TMEM lane quarter with tcgen05.ld 32x32b.x32, add rowterm[x], store fp32.
#include "../../include/aligner_b200.h"

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace alb {
extern thread_local char g_err[512];
extern thread_local uint64_t g_launches;
}

namespace albtc {

constexpr int BM = 128, BN = 256, KCH = 16;                 // tile rows, tile cols, channels per chunk (32 k)
constexpr int A_TILE = BM * 128, B_TILE = BN * 128;         // bytes: rows x 128-byte swizzled K row
constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;              // Ahi, Alo, Bhi, Blo = 96 KB
constexpr int SMEM_BYTES = 2 * STAGE + 1024 /*align slack*/ + 2048 /*rowterm + barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void sts128u(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    // no "memory" clobber on purpose: the global loads of the next chunk must be free to move above these stores
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // bounded: a wrong descriptor must surface as a launch error, never as a hung GPU
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 24)) __trap();
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, LBO (unused for
// swizzled K-major, 1), SBO = 1024 bytes between 8-row groups, version 1 (Blackwell), layout type 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(256, 1) gaussian_tc_kernel(const float* __restrict__ z, const float* __restrict__ m,
                                                             const float* __restrict__ logs, float* __restrict__ out, int C, int Tx, int Ty)
{
    extern __shared__ unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int b = blockIdx.z, x0 = blockIdx.y * BM, y0 = blockIdx.x * BN;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // swizzle atoms need 1024-byte alignment
    unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
    float* rowpart = reinterpret_cast<float*>(gbase + 2 * STAGE);         // [2][128]
    const uint32_t bars = base + 2 * STAGE + 1024;                        // empty[0], empty[1], accum, tmem slot
    const uint32_t bar_empty0 = bars, bar_accum = bars + 16, tmem_slot = bars + 32;

    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(bar_empty0, 1); mbar_init(bar_empty0 + 8, 1); mbar_init(bar_accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    // producer roles: A row = tid % 128, channels 8*(tid/128) .. +7 of the chunk;  B row = tid, all 16 channels.
    // Rows past t_text / t_mel are clamped to a valid row: their products land in accumulator rows / columns the
    // epilogue never stores, so no masking is needed on the fast path (only the K tail must be zero).
    const int arow = tid & 127, ahalf = tid >> 7;
    const int ax = min(x0 + arow, Tx - 1), by = min(y0 + tid, Ty - 1);
    const float* pm = m + (size_t)b * C * Tx + (size_t)(8 * ahalf) * Tx + ax;
    const float* pl = logs + (size_t)b * C * Tx + (size_t)(8 * ahalf) * Tx + ax;
    const float* pz = z + (size_t)b * C * Ty + by;
    const uint32_t a_off = (uint32_t)(arow >> 3) * 1024 + (uint32_t)(arow & 7) * 128;
    const uint32_t b_off = (uint32_t)(tid >> 3) * 1024 + (uint32_t)(tid & 7) * 128;
    const uint32_t a_x = (uint32_t)(arow & 7), b_x = (uint32_t)(tid & 7);
    float rsum = 0.f;

    const int nchunks = (C + KCH - 1) / KCH;
    // raw operands of one chunk, fetched one chunk ahead so the global-load latency hides behind the previous chunk's work
    float am[8], al[8], bz[16];
    auto fetch = [&](int c0) {
        if (c0 + KCH <= C) {                                   // whole chunk: 32 unconditional strided loads
#pragma unroll
            for (int q = 0; q < 8; ++q) { am[q] = pm[(size_t)q * Tx]; al[q] = pl[(size_t)q * Tx]; }
#pragma unroll
            for (int q = 0; q < 16; ++q) bz[q] = pz[(size_t)q * Ty];
        } else {                                               // K tail: channels >= C contribute exact zeros
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const bool ok = c0 + 8 * ahalf + q < C;
                am[q] = ok ? pm[(size_t)q * Tx] : 0.f;
                al[q] = ok ? pl[(size_t)q * Tx] : __int_as_float(0x7f800000);   // +inf marks padding: s2 = 0, no row term
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) bz[q] = (c0 + q < C) ? pz[(size_t)q * Ty] : 0.f;
        }
        pm += (size_t)KCH * Tx; pl += (size_t)KCH * Tx; pz += (size_t)KCH * Ty;
    };
    // TF32 split: hi keeps the top 19 bits (what the tensor core reads), lo = v - hi is exact in fp32 and is itself read
    // truncated, so hi + lo carries ~21 mantissa bits of v
    auto split4 = [&](const float (&v)[4], uint32_t (&hi)[4], uint32_t (&lo)[4]) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            hi[q] = __float_as_uint(v[q]) & 0xffffe000u;
            lo[q] = __float_as_uint(v[q] - __uint_as_float(hi[q]));
        }
    };
    fetch(0);
    for (int i = 0; i < nchunks; ++i) {
        const int st = i & 1;
        if (i >= 2) mbar_wait(bar_empty0 + 8 * st, (uint32_t)(((i >> 1) - 1) & 1));
        const uint32_t sA_hi = base + st * STAGE, sA_lo = sA_hi + A_TILE, sB_hi = sA_lo + A_TILE, sB_lo = sB_hi + B_TILE;
        // ---- A: 4 chunks of 16 bytes (2 channels -> s2, m s2, s2', m' s2')
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v[4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float mm = am[2 * j + h], lg = al[2 * j + h];
                const bool pad = (lg == __int_as_float(0x7f800000));
                const float s2 = pad ? 0.f : __expf(-2.f * lg);
                const float ms2 = mm * s2;
                rsum += pad ? 0.f : ((-0.9189385332046727f - lg) - 0.5f * mm * ms2);
                v[2 * h] = s2; v[2 * h + 1] = ms2;
            }
            uint32_t hi[4], lo[4];
            split4(v, hi, lo);
            const uint32_t off = a_off + ((((uint32_t)(4 * ahalf + j)) ^ a_x) << 4);   // 16-byte chunk index along K, swizzled
            sts128u(sA_hi + off, hi[0], hi[1], hi[2], hi[3]);
            sts128u(sA_lo + off, lo[0], lo[1], lo[2], lo[3]);
        }
        // ---- B: 8 chunks of 16 bytes (2 channels -> -0.5 z^2, z, -0.5 z'^2, z')
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v[4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float zz = bz[2 * j + h];
                v[2 * h] = -0.5f * zz * zz; v[2 * h + 1] = zz;
            }
            uint32_t hi[4], lo[4];
            split4(v, hi, lo);
            const uint32_t off = b_off + (((uint32_t)j ^ b_x) << 4);
            sts128u(sB_hi + off, hi[0], hi[1], hi[2], hi[3]);
            sts128u(sB_lo + off, lo[0], lo[1], lo[2], lo[3]);
        }
        if (i + 1 < nchunks) fetch((i + 1) * KCH);                             // in flight across the barrier and the MMA issue
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int prod = 0; prod < 3; ++prod) {
                const uint32_t sa = (prod == 1) ? sA_lo : sA_hi;                // hi*hi, lo*hi, hi*lo
                const uint32_t sb = (prod == 2) ? sB_lo : sB_hi;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)                                  // 4 k-steps of 8 tf32 = 32 bytes along the swizzled row
                    mma_tf32(tmem_base, make_desc(sa + ks * 32), make_desc(sb + ks * 32), (i | prod | ks) ? 1u : 0u);
            }
            mma_commit(bar_empty0 + 8 * st);                                    // stage reusable when these MMAs have read it
            if (i == nchunks - 1) mma_commit(bar_accum);                        // accumulator complete
        }
    }
    rowpart[ahalf * 128 + arow] = rsum;
    mbar_wait(bar_accum, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();

    // ---- epilogue: warp w owns TMEM lanes 32*(w%4)..+31 (tile rows) and columns 128*(w/4)..+127
    {
        const int q = wid & 3, half = wid >> 2;
        const int row = 32 * q + lane, x = x0 + row;
        const float rt = rowpart[row] + rowpart[128 + row];
        float* orow = out + ((size_t)b * Tx + (x < Tx ? x : 0)) * Ty;
#pragma unroll 1
        for (int it = 0; it < 4; ++it) {
            const int col0 = 128 * half + 32 * it;
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)col0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (x < Tx) {
                const int yb = y0 + col0;
                float* o = orow + yb;
                if (yb + 32 <= Ty && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        reinterpret_cast<float4*>(o)[j] = make_float4(__uint_as_float(r[4 * j]) + rt, __uint_as_float(r[4 * j + 1]) + rt,
                                                                       __uint_as_float(r[4 * j + 2]) + rt, __uint_as_float(r[4 * j + 3]) + rt);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (yb + j < Ty) o[j] = __uint_as_float(r[j]) + rt;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
}

}  // namespace albtc

extern "C" int alb200_neg_cent_gaussian_tc(const float* z, const float* m_p, const float* logs_p, float* out, int b, int c, int tx, int ty,
                                           void* stream)
{
    using namespace albtc;
    static thread_local bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gaussian_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) { snprintf(alb::g_err, sizeof(alb::g_err), "neg_cent_gaussian_tc: %s", cudaGetErrorString(e)); return ALB200_E_CUDA; }
        configured = true;
    }
    dim3 grid((ty + BN - 1) / BN, (tx + BM - 1) / BM, b);
    gaussian_tc_kernel<<<grid, 256, SMEM_BYTES, (cudaStream_t)stream>>>(z, m_p, logs_p, out, c, tx, ty);
    ++alb::g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(alb::g_err, sizeof(alb::g_err), "neg_cent_gaussian_tc: %s", cudaGetErrorString(e)); return ALB200_E_CUDA; }
    return 0;
}
