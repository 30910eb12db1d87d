// neg_cent_tc.cu -- both score matrices on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// One CTA owns 128 mel frames (the M axis = TMEM lanes) and up to 512 text tokens (the N axis = TMEM columns):
//
//   gaussian  D[y,x] = sum_k A[y,k] B[x,k],  A = (-0.5 z^2, z) per channel,  B = (s2, m s2) per channel,  K = 2C
//             out[b,x,y] = D[y,x] + colterm[x],   colterm = sum_c(-0.5 log 2pi - logs - 0.5 m^2 s2)
//   ota       D[y,x] = sum_c q[c,y] k[c,x],  K = C
//             d = -T (|q_y|^2 + |k_x|^2 - 2 D),   out[b,x,y] = d - logsumexp_x(d) + log(prior + 1e-8)
//
// Why this orientation: the 128-lane M axis is filled by the long mel axis (1000 = 7.8 tiles) instead of the short
// text axis (200 = 1.56 tiles); a whole text column sits in ONE thread's TMEM lane, so the OTA log-softmax over the
// text axis is a per-thread reduction; and for a fixed token the 32 lanes of a warp hold 32 consecutive frames, so
// every store of the [b, t_text, t_mel] result is a full 128-byte line.
//
// Operands are produced by the CTA itself: raw values are read coalesced in their native layouts, transformed, split
// into a TF32 "hi" part and a TF32 "lo" remainder (3xTF32: hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM -- a
// single TF32 pass misses the 1e-5 bound by two orders of magnitude) and stored straight into the canonical K-major
// SWIZZLE_128B layout the UMMA descriptors address (16-byte chunk index XOR (row % 8), 8-row groups 1024 bytes apart).
// Per 32-k chunk, two stages: all 256 threads produce chunk i+1 while the tensor pipe multiplies chunk i
// (tcgen05.mma.cta_group::1.kind::tf32 issued by one thread, tcgen05.commit -> mbarrier releases the stage).
// Text blocks of 256 columns are separate passes over K into disjoint TMEM column ranges.
#include "../../include/aligner_b200.h"

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <type_traits>

namespace alb {
extern thread_local char g_err[512];
extern thread_local uint64_t g_launches;
}

namespace albtc {

constexpr int BM = 128;                                     // mel frames per CTA (TMEM lanes)
constexpr int NPASS = 256;                                  // text tokens per pass (max N of one tcgen05.mma)
constexpr int NMAX = 512;                                   // text tokens per CTA (TMEM columns)
constexpr int A_TILE = BM * 128, B_TILE = NPASS * 128;      // bytes: rows x one 128-byte swizzled K row (32 tf32)
constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;              // Ahi, Alo, Bhi, Blo = 96 KB
constexpr int AUX = 8192;                                   // colterm/knorm[512], qnorm[2][128], softmax exchange, barriers
constexpr int SMEM_BYTES = 2 * STAGE + 1024 /*align slack*/ + AUX;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4, LBO (unused for
// swizzled K-major, 1), SBO = 1024 bytes between 8-row groups, version 1 (Blackwell), layout type 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
__device__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// exp(x) for the variance term: one FMUL + one MUFU.EX2 (2^-22 relative error, far inside the 1e-5 budget)
__device__ __forceinline__ float fast_exp(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
}
// TF32 split: hi keeps the top 19 bits (what the tensor core reads), lo = v - hi is exact in fp32 and is itself read
// truncated, so hi + lo carries ~21 mantissa bits of v
__device__ __forceinline__ void split4(const float (&v)[4], uint32_t (&hi)[4], uint32_t (&lo)[4]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        hi[q] = __float_as_uint(v[q]) & 0xffffe000u;
        lo[q] = __float_as_uint(v[q] - __uint_as_float(hi[q]));
    }
}

struct TcParams {
    const float* a_src;        // mel side:  z [b,C,Ty]   or queries [b,C,Ty]
    const float* b_src0;       // text side: m_p [b,C,Tx] or keys [b,C,Tx]
    const float* b_src1;       // text side: logs_p [b,C,Tx] (gaussian only)
    const float* prior;        // ota: optional [b,Tx,Ty]
    const int32_t* x_lengths;  // ota: optional [b]
    float* out;                // [b,Tx,Ty]
    float temperature;
    int C, Tx, Ty;
};

// MODE 0 = gaussian (16 channels per 32-k chunk), MODE 1 = ota (32 channels per chunk)
template <int MODE>
__global__ void __launch_bounds__(256, 1) nc_tc_kernel(const TcParams p)
{
    constexpr int CPC = MODE == 0 ? 16 : 32;                 // channels per chunk
    constexpr int ACH = CPC / 2;                             // channels per A-producer thread (two threads per mel row)
    extern __shared__ unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int b = blockIdx.z, y0 = blockIdx.x * BM, xt0 = blockIdx.y * NMAX;
    const int C = p.C, Tx = p.Tx, Ty = p.Ty;
    const int ntext = min(NMAX, Tx - xt0);                   // text tokens of this CTA
    const int NT = (ntext + 15) & ~15;                       // accumulator columns (N must be a multiple of 16)
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // swizzle atoms need 1024-byte alignment
    unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
    float* colv = reinterpret_cast<float*>(gbase + 2 * STAGE);            // [512] colterm (gaussian) / |k|^2 (ota)
    float* rowv = colv + NMAX;                                            // [2][128] |q|^2 halves (ota)
    float* xch = rowv + 2 * BM;                                           // [2][128][2] softmax exchange (ota)
    const uint32_t bars = base + 2 * STAGE + AUX - 64;                    // empty[0], empty[1], accum, tmem slot
    const uint32_t bar_empty0 = bars, bar_accum = bars + 16, tmem_slot = bars + 32;

    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(NMAX) : "memory");
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

    // producer roles.  A (mel): row = tid % 128, channels ACH*(tid/128) .. of the chunk.  B (text): row = tid, all channels.
    // Rows past the tensor are clamped to a valid row: their products land in accumulator rows / columns that are never
    // stored, so the fast path needs no masking (only the K tail must be exact zeros).
    const int arow = tid & 127, ahalf = tid >> 7;
    const int ay = min(y0 + arow, Ty - 1);
    const float* pa0 = p.a_src + (size_t)b * C * Ty + (size_t)(ACH * ahalf) * Ty + ay;
    const uint32_t a_off = (uint32_t)(arow >> 3) * 1024 + (uint32_t)(arow & 7) * 128;
    const uint32_t b_off = (uint32_t)(tid >> 3) * 1024 + (uint32_t)(tid & 7) * 128;
    const uint32_t a_x = (uint32_t)(arow & 7), b_x = (uint32_t)(tid & 7);
    const int nchunks = (C + CPC - 1) / CPC;
    const int npass = (NT + NPASS - 1) / NPASS;
    float qn = 0.f;                                          // ota: this thread's share of |q_y|^2

    int it = 0;                                              // global chunk counter: stage = it & 1
    for (int pass = 0; pass < npass; ++pass) {
        const int ncols = min(NPASS, NT - pass * NPASS);
        const uint32_t idesc = make_idesc(ncols);
        const int bx = min(xt0 + pass * NPASS + tid, Tx - 1);
        const float* pa = pa0;
        const float* pb0 = p.b_src0 + (size_t)b * C * Tx + bx;
        const float* pb1 = MODE == 0 ? p.b_src1 + (size_t)b * C * Tx + bx : nullptr;
        float cacc = 0.f;                                    // colterm (gaussian) / |k_x|^2 (ota) of text row bx

        // raw operands of two chunks: chunk i+1 is fetched BEFORE chunk i is transformed, so a full chunk of work (and the
        // MMA wait) hides the global-load latency; two register sets, selected at compile time (loop unrolled by two)
        float av[2][ACH], bv0[2][CPC], bv1[2][MODE == 0 ? CPC : 1];
        auto fetch = [&](auto bufc, int c0) {
            constexpr int B_ = decltype(bufc)::value;
            if (c0 + CPC <= C) {                             // whole chunk: unconditional strided loads
#pragma unroll
                for (int q = 0; q < ACH; ++q) av[B_][q] = pa[(size_t)q * Ty];
#pragma unroll
                for (int q = 0; q < CPC; ++q) { bv0[B_][q] = pb0[(size_t)q * Tx]; if (MODE == 0) bv1[B_][q] = pb1[(size_t)q * Tx]; }
            } else {                                         // K tail: channels >= C contribute exact zeros
#pragma unroll
                for (int q = 0; q < ACH; ++q) av[B_][q] = (c0 + ACH * ahalf + q < C) ? pa[(size_t)q * Ty] : 0.f;
#pragma unroll
                for (int q = 0; q < CPC; ++q) {
                    const bool ok = c0 + q < C;
                    bv0[B_][q] = ok ? pb0[(size_t)q * Tx] : 0.f;
                    if (MODE == 0) bv1[B_][q] = ok ? pb1[(size_t)q * Tx] : __int_as_float(0x7f800000);   // +inf marks padding
                }
            }
            pa += (size_t)CPC * Ty; pb0 += (size_t)CPC * Tx;
            if (MODE == 0) pb1 += (size_t)CPC * Tx;
        };
        auto chunk = [&](auto bufc, auto nextc, int i) {
            constexpr int B_ = decltype(bufc)::value;
            if (i + 1 < nchunks) fetch(nextc, (i + 1) * CPC);
            const int st = it & 1;
            if (it >= 2) mbar_wait(bar_empty0 + 8 * st, (uint32_t)(((it >> 1) - 1) & 1));
            const uint32_t sA_hi = base + st * STAGE, sA_lo = sA_hi + A_TILE, sB_hi = sA_lo + A_TILE, sB_lo = sB_hi + B_TILE;
            // ---- A (mel side): 4 chunks of 16 bytes per thread
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v[4];
                if (MODE == 0) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) { const float zz = av[B_][2 * j + h]; v[2 * h] = -0.5f * zz * zz; v[2 * h + 1] = zz; }
                } else {
#pragma unroll
                    for (int h = 0; h < 4; ++h) { v[h] = av[B_][4 * j + h]; if (pass == 0) qn = fmaf(v[h], v[h], qn); }
                }
                uint32_t hi[4], lo[4];
                split4(v, hi, lo);
                const uint32_t off = a_off + ((((uint32_t)(4 * ahalf + j)) ^ a_x) << 4);     // chunk index along K, swizzled
                sts128u(sA_hi + off, hi[0], hi[1], hi[2], hi[3]);
                sts128u(sA_lo + off, lo[0], lo[1], lo[2], lo[3]);
            }
            // ---- B (text side): 8 chunks of 16 bytes per thread; rows >= NT are never read by the MMA
            if (tid < ncols) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float v[4];
                    if (MODE == 0) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const float mm = bv0[B_][2 * j + h], lg = bv1[B_][2 * j + h];
                            const bool pad = (lg == __int_as_float(0x7f800000));
                            const float s2 = pad ? 0.f : fast_exp(-2.f * lg);
                            const float ms2 = mm * s2;
                            cacc += pad ? 0.f : ((-0.9189385332046727f - lg) - 0.5f * mm * ms2);
                            v[2 * h] = s2; v[2 * h + 1] = ms2;
                        }
                    } else {
#pragma unroll
                        for (int h = 0; h < 4; ++h) { v[h] = bv0[B_][4 * j + h]; cacc = fmaf(v[h], v[h], cacc); }
                    }
                    uint32_t hi[4], lo[4];
                    split4(v, hi, lo);
                    const uint32_t off = b_off + (((uint32_t)j ^ b_x) << 4);
                    sts128u(sB_hi + off, hi[0], hi[1], hi[2], hi[3]);
                    sts128u(sB_lo + off, lo[0], lo[1], lo[2], lo[3]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // generic-proxy stores -> visible to the tensor core
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t dcol = tmem_base + (uint32_t)(pass * NPASS);
#pragma unroll
                for (int prod = 0; prod < 3; ++prod) {
                    const uint32_t sa = (prod == 1) ? sA_lo : sA_hi;                // hi*hi, lo*hi, hi*lo
                    const uint32_t sb = (prod == 2) ? sB_lo : sB_hi;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)                                  // 4 k-steps of 8 tf32 = 32 bytes along the swizzled row
                        mma_tf32(dcol, make_desc(sa + ks * 32), make_desc(sb + ks * 32), idesc, (i | prod | ks) ? 1u : 0u);
                }
                mma_commit(bar_empty0 + 8 * st);                                    // stage reusable when these MMAs have read it
                if (pass == npass - 1 && i == nchunks - 1) mma_commit(bar_accum);   // accumulator complete
            }
            ++it;
        };
        using I0 = std::integral_constant<int, 0>;
        using I1 = std::integral_constant<int, 1>;
        fetch(I0{}, 0);
        for (int i = 0; i < nchunks; i += 2) {
            chunk(I0{}, I1{}, i);
            if (i + 1 < nchunks) chunk(I1{}, I0{}, i + 1);
        }
        if (pass * NPASS + tid < NMAX) colv[pass * NPASS + tid] = cacc;
    }
    if (MODE == 1) rowv[ahalf * BM + arow] = qn;
    mbar_wait(bar_accum, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();

    // ---- epilogue: warp w owns TMEM lanes 32*(w%4)..+31 (mel frames) and one half of the 16-column groups (text tokens)
    {
        const int q = wid & 3, half = wid >> 2;
        const int row = 32 * q + lane, y = y0 + row;
        const bool y_ok = y < Ty;
        const int ngroups = NT >> 4;
        const int g_lo = half == 0 ? 0 : (ngroups + 1) / 2, g_hi = half == 0 ? (ngroups + 1) / 2 : ngroups;
        const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
        float* ob = p.out + (size_t)b * Tx * Ty + (size_t)xt0 * Ty + (y_ok ? y : 0);
        if (MODE == 0) {
            for (int g = g_lo; g < g_hi; ++g) {
                uint32_t r[16];
                tmem_ld16(tlane + (uint32_t)(16 * g), r);
                if (y_ok) {
                    float* o = ob + (size_t)(16 * g) * Ty;
                    const float4* cv = reinterpret_cast<const float4*>(colv + 16 * g);
                    if (16 * g + 16 <= ntext) {                  // whole group: no per-token checks; 32 lanes = 32 consecutive frames
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const float4 c4 = cv[j4];
                            o[0] = __uint_as_float(r[4 * j4]) + c4.x; o += Ty;
                            o[0] = __uint_as_float(r[4 * j4 + 1]) + c4.y; o += Ty;
                            o[0] = __uint_as_float(r[4 * j4 + 2]) + c4.z; o += Ty;
                            o[0] = __uint_as_float(r[4 * j4 + 3]) + c4.w; o += Ty;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (16 * g + j < ntext) o[(size_t)j * Ty] = __uint_as_float(r[j]) + colv[16 * g + j];
                    }
                }
            }
        } else {
            const int tlen = p.x_lengths ? min(max(p.x_lengths[b], 0), Tx) : Tx;
            const float qn2 = rowv[row] + rowv[BM + row];
            const float T = p.temperature;
            float mx = -INFINITY, sm = 0.f;                   // online log-sum-exp over this warp's half of the text axis
            for (int g = g_lo; g < g_hi; ++g) {
                uint32_t r[16];
                tmem_ld16(tlane + (uint32_t)(16 * g), r);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int xl = 16 * g + j;
                    if (xl < tlen) {
                        const float d = -T * (qn2 + colv[xl] - 2.f * __uint_as_float(r[j]));
                        const float nm = fmaxf(mx, d);
                        sm = sm * __expf(mx - nm) + __expf(d - nm);
                        mx = nm;
                    }
                }
            }
            xch[(half * BM + row) * 2] = mx; xch[(half * BM + row) * 2 + 1] = sm;
            __syncthreads();
            const float m0 = xch[row * 2], s0 = xch[row * 2 + 1], m1 = xch[(BM + row) * 2], s1 = xch[(BM + row) * 2 + 1];
            const float gm = fmaxf(m0, m1);
            const float tot = (s0 > 0.f ? s0 * __expf(m0 - gm) : 0.f) + (s1 > 0.f ? s1 * __expf(m1 - gm) : 0.f);
            const float lse = gm + __logf(tot);
            const float* pr = p.prior ? p.prior + (size_t)b * Tx * Ty + (y_ok ? y : 0) : nullptr;
            for (int g = g_lo; g < g_hi; ++g) {
                uint32_t r[16];
                tmem_ld16(tlane + (uint32_t)(16 * g), r);
                if (y_ok) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int xl = 16 * g + j;
                        if (xl < ntext) {
                            float v = -INFINITY;                // text padding is excluded from the softmax
                            if (xl < tlen) {
                                v = -T * (qn2 + colv[xl] - 2.f * __uint_as_float(r[j])) - lse;
                                if (pr) v += __logf(pr[(size_t)xl * Ty] + 1e-8f);
                            }
                            ob[(size_t)xl * Ty] = v;
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(NMAX) : "memory");
}

template <int MODE>
static int launch(const TcParams& p, int b, void* stream, const char* who)
{
    // the attribute is per device: remember it per (thread, device)
    static thread_local bool configured[64] = {false};
    int dev = 0;
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 != cudaSuccess) { snprintf(alb::g_err, sizeof(alb::g_err), "%s: %s", who, cudaGetErrorString(e0)); return ALB200_E_NO_DEVICE; }
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(nc_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) { snprintf(alb::g_err, sizeof(alb::g_err), "%s: %s", who, cudaGetErrorString(e)); return ALB200_E_CUDA; }
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    dim3 grid((p.Ty + BM - 1) / BM, (p.Tx + NMAX - 1) / NMAX, b);
    nc_tc_kernel<MODE><<<grid, 256, SMEM_BYTES, (cudaStream_t)stream>>>(p);
    ++alb::g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(alb::g_err, sizeof(alb::g_err), "%s: %s", who, cudaGetErrorString(e)); return ALB200_E_CUDA; }
    return 0;
}

}  // namespace albtc

extern "C" int alb200_neg_cent_gaussian_tc(const float* z, const float* m_p, const float* logs_p, float* out, int b, int c, int tx, int ty,
                                           void* stream)
{
    albtc::TcParams p{z, m_p, logs_p, nullptr, nullptr, out, 0.f, c, tx, ty};
    return albtc::launch<0>(p, b, stream, "neg_cent_gaussian_tc");
}

// t_text <= 512 only: the log-softmax needs the whole text column in one CTA's tensor memory
extern "C" int alb200_neg_cent_ota_tc(const float* queries, const float* keys, const float* prior, const int32_t* x_lengths, float* out,
                                      float temperature, int b, int c, int tx, int ty, void* stream)
{
    albtc::TcParams p{queries, keys, nullptr, prior, x_lengths, out, temperature, c, tx, ty};
    return albtc::launch<1>(p, b, stream, "neg_cent_ota_tc");
}
