// neg_cent_v2.cu -- score matrices on tcgen05, second generation: TMA-fed, warp-specialised, persistent.
//
//   gaussian  out[b,x,y] = colterm[x] + sum_c ( s2[c,x] * (-0.5 z[c,y]^2) + (m s2)[c,x] * z[c,y] )        K = 2C
//   ota       out[b,x,y] = log_softmax_x( -T (|q_y|^2 + |k_x|^2 - 2 q_y.k_x) ) + log(prior + 1e-8)         K = C
//
// What changed against neg_cent_tc.cu (which stays as the fallback for rows that are not 16-byte aligned and t_x > 512):
//   * fp16 x 3 instead of tf32 x 3.  Every operand element v is split as hi = fp16(v), lo = fp16(v - hi) and the product
//     is hi*hi + lo*hi + hi*lo with fp32 accumulation in tensor memory -- the same 22 mantissa bits as the tf32 split (both
//     formats carry 11 significant bits), at twice the tensor rate and half the shared-memory bytes.  fp16's narrow exponent
//     is handled by exact power-of-two scaling: text-side rows are scaled to [2^14, 2^15) by the prep kernel (the inverse is
//     applied to the accumulator in the epilogue), the mel side uses fixed factors (z * 2^5, -0.5 z^2 * 2^-2; q * 2^5) that
//     are exact for |z| < 500 (|q| < 1000).  A tile that sees a larger or non-finite mel-side value is recomputed by its own
//     CTA in plain fp32 from the raw inputs (slow, exact, never silent).  numpy model: within 1e-7 of fp64 (tools/sim_f16x3.py).
//   * The text-side operand (exp, products, split, scaling, per-token terms) is produced ONCE per utterance by nc_prep_kernel
//     into an L2-resident [b, t_x, K] fp16 hi/lo pair; the first generation recomputed it in every mel tile (8x for C2).
//   * All global->shared traffic is TMA (cp.async.bulk.tensor.3d): raw z / q boxes [32 channels x 128 frames] into a staging
//     ring, text-side boxes [<=256 tokens x 64 k] straight into the K-major SWIZZLE_128B layout the UMMA descriptors address.
//   * Warp roles: 1 TMA producer, 1 MMA issuer (single thread, tcgen05.mma.kind::f16), 8 transform warps (raw box -> scaled
//     hi/lo fp16 A tiles in the swizzled layout), 4 epilogue warps (tcgen05.ld -> scale, per-token term / log-softmax -> coalesced
//     stores).  Two 256-column accumulators in tensor memory: the epilogue of tile i overlaps the MMAs of tile i+1
//     (t_x <= 256; longer texts use both halves for one tile).  Persistent grid, one CTA per SM, static round-robin over
//     (utterance, mel tile) so that neighbouring CTAs share an utterance's text operand in L2.
#include "../../include/aligner_b200.h"
#include "alb_opts.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

namespace alb {
extern thread_local char g_err[512];
extern thread_local uint64_t g_launches;
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmapEncodeFn tmap_encode_fn();          // mas_api.cu
}

namespace albv2 {

constexpr int BM = 128;                       // mel frames per tile (TMEM lanes)
constexpr int KC = 64;                        // fp16 k per chunk: one 128-byte swizzled row
constexpr int RAW_CH = 32;                    // channels per raw box
constexpr int RAW_BYTES = RAW_CH * BM * 4;    // 16 KB
constexpr int A_TILE = BM * 128;              // 16 KB: 128 rows x one swizzled 128-byte row (hi or lo)
constexpr int SA = 2;                         // A / B stages
constexpr int NTHREADS = 15 * 32;             // warp 0 raw-box TMA, warp 1 MMA, warp 2 text-operand TMA, warps 3-10 transform, warps 11-14 epilogue
constexpr int N_TRANSFORM = 256;
constexpr int AUX_BYTES = 256 + 2 * 512 * 4 + 3 * 128 * 8;  // barriers, tmem slot, per-tile flags (256 B), the tile's per-token terms [2][512] fp32,
                                                            // and the log-sum-exp partials of the three warps of a TMEM quadrant (ota, long texts)
constexpr float kZLimit = 500.f, kQLimit = 1000.f;
// fixed mel-side factors (exact powers of two) and their inverses on the text side
constexpr float kA1 = -0.5f * 0.25f;          // -0.5 z^2 * 2^-2
constexpr float kA2 = 32.f;                   //  z * 2^5        (also q * 2^5)
constexpr float kB1 = 4.f, kB2 = 1.f / 32.f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // bounded: a wrong descriptor or tensor map must surface as a launch error, never as a hung GPU
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 26)) __trap();
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4, LBO (unused for
// swizzled K-major, 1), SBO = 1024 bytes between 8-row groups, version 1 (Blackwell), layout type 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bit 4), A = B = F16 (format 0), both K-major, N >> 3, M >> 4
__device__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// same, accumulating unconditionally (the predicate folds to PT: no SETP on the issuing thread's critical path)
__device__ __forceinline__ void mma_f16_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
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
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float fast_exp(float x) {                  // one FMUL + MUFU.EX2 (2^-22 relative)
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
}
__device__ __forceinline__ float fast_exp2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_log2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// (v0, v1) -> packed fp16 hi pair and packed fp16 lo pair;  hi = rn(v), lo = rn(v - hi)
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ------------------------------------------------------------------ text-side operand, once per utterance
// grid (ceil(Tx / 8), b), 256 threads: thread = (token tid % 8, channel group tid / 8); a CTA owns 8 tokens so that many CTAs share
// an SM and hide each other's DRAM round trips.  MODE 0: rows (s2 * 4, m s2 / 32) per channel, colterm; MODE 1: rows k / 32, |k|^2.
// Each row is scaled by a power of two to [2^14, 2^15), split into fp16 hi / lo and written K-contiguous ([b, Tx, Kpad], what the
// SWIZZLE_128B tensor map of the main kernel reads); k >= K is zero.
constexpr int PT = 8;                         // tokens per prep CTA
constexpr int PG = 16;                        // channel groups per prep CTA (PT * PG = 128 threads: 16 CTAs per SM, one wave for 64 x 200 tokens)
constexpr int PCH = 12;                       // channels a prep thread loads per batch (all in flight together: one DRAM round trip up to C = 192)
template <int MODE>
__global__ void __launch_bounds__(PT * PG) nc_prep_kernel(const float* __restrict__ src0, const float* __restrict__ src1, __half* __restrict__ b_hi,
                                                          __half* __restrict__ b_lo, float* __restrict__ colv, float* __restrict__ inv_sb, int C, int Tx,
                                                          int K, int Kpad, int TxS, float temperature)
{
    extern __shared__ __align__(16) float vals[];              // [PT][Kpad + 4]
    __shared__ float part_max[PG / 4][PT], part_sum[PG / 4][PT], row_scale[PT];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the score kernel may start its prologue and its mel-side work now
    const int tid = threadIdx.x, tok = tid & (PT - 1), grp = tid >> 3;      // PG channel groups; a warp holds 4 of them for all 8 tokens
    const int b = blockIdx.y, x = blockIdx.x * PT + tok;
    const int VS = Kpad + 4;
    const bool ok = x < Tx;
    const size_t cstep = (size_t)PG * Tx;                        // this thread's channels are grp, grp + PG, ...
    const float* p0 = src0 + (size_t)b * C * Tx + (size_t)grp * Tx + (ok ? x : 0);
    const float* p1 = MODE == 0 ? src1 + (size_t)b * C * Tx + (size_t)grp * Tx + (ok ? x : 0) : nullptr;
    float* vrow = vals + tok * VS + (MODE == 0 ? 2 * grp : grp);
    float mx = 0.f, sum = 0.f;
    for (int c0 = grp; c0 < C; c0 += PG * PCH) {
        float a[PCH], l[PCH];
#pragma unroll
        for (int u = 0; u < PCH; ++u) {                          // all loads of the batch first: one DRAM round trip
            const bool in = ok && (c0 + PG * u < C);
            a[u] = in ? p0[u * cstep] : 0.f;
            if (MODE == 0) l[u] = in ? p1[u * cstep] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < PCH; ++u) {
            if (c0 + PG * u < C) {
                if (MODE == 0) {
                    const float mm = a[u], lg = l[u];
                    const float s2 = fast_exp(-2.f * lg), ms2 = mm * s2;
                    if (ok) sum += (-0.9189385332046727f - lg) - 0.5f * mm * ms2;      // -0.5 log(2 pi) - logs - 0.5 m^2 s2
                    const float v0 = ok ? s2 * kB1 : 0.f, v1 = ok ? ms2 * kB2 : 0.f;
                    *reinterpret_cast<float2*>(vrow + 2 * PG * u) = make_float2(v0, v1);
                    mx = fmaxf(mx, fmaxf(fabsf(v0), fabsf(v1)));
                } else {
                    sum = fmaf(a[u], a[u], sum);
                    const float v0 = a[u] * kB2;
                    vrow[PG * u] = v0;
                    mx = fmaxf(mx, fabsf(v0));
                }
            }
        }
        p0 += PCH * cstep; vrow += (MODE == 0 ? 2 * PG : PG) * PCH;
        if (MODE == 0) p1 += PCH * cstep;
    }
    for (int k = K + grp; k < Kpad; k += PG) vals[tok * VS + k] = 0.f;
    // per-token maximum and sum: the four channel groups of a warp by shuffles (lanes 8 apart hold the same token), fixed order
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8)); sum += __shfl_xor_sync(0xffffffffu, sum, 8);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16)); sum += __shfl_xor_sync(0xffffffffu, sum, 16);
    if ((tid & 31) < PT) { part_max[tid >> 5][tok] = mx; part_sum[tid >> 5][tok] = sum; }
    __syncthreads();
    if (tid < PT) {
        float m8 = part_max[0][tid], s8 = part_sum[0][tid];
#pragma unroll
        for (int g = 1; g < PG / 4; ++g) { m8 = fmaxf(m8, part_max[g][tid]); s8 += part_sum[g][tid]; }      // fixed order: deterministic
        // power-of-two scale that puts the row maximum into [2^14, 2^15); rows of zeros (or non-finite rows) keep scale 1
        float sc = 1.f;
        if (m8 > 0.f && m8 < 3.0e38f) {
            const int e = (int)((__float_as_uint(m8) >> 23) & 0xff) - 127;       // floor(log2(m8)) for normal numbers
            int sh = 14 - e;
            sh = sh > 100 ? 100 : (sh < -100 ? -100 : sh);
            sc = __uint_as_float((uint32_t)(127 + sh) << 23);
        }
        row_scale[tid] = sc;
        const int xr = blockIdx.x * PT + tid;
        if (xr < Tx) {
            if (MODE == 0) { colv[(size_t)b * TxS + xr] = s8; inv_sb[(size_t)b * TxS + xr] = 1.f / sc; }
            else {
                // d[x,y] = -T (|q_y|^2 + |k_x|^2 - 2 q.k); |q_y|^2 is constant along the softmax axis and cancels in the log-softmax,
                // so the epilogue needs d' = acc * (2 T / scale) - T |k_x|^2 only: one FFMA per cell
                colv[(size_t)b * TxS + xr] = -temperature * s8;
                inv_sb[(size_t)b * TxS + xr] = 2.f * temperature / sc;
            }
        }
    }
    __syncthreads();
    // one warp per pair of token rows, lanes over a row's 16-byte chunks (8 fp16): coalesced 512-byte stores
    const int cpr = Kpad >> 3;
    for (int r = tid >> 5; r < PT; r += PT * PG / 32) {
        const int xr = blockIdx.x * PT + r;
        if (xr >= Tx) continue;
        const float sc = row_scale[r];
        const float* vr = vals + r * VS;
        __half* oh = b_hi + ((size_t)b * Tx + xr) * Kpad;
        __half* ol = b_lo + ((size_t)b * Tx + xr) * Kpad;
        for (int j = tid & 31; j < cpr; j += 32) {
            const float4 a = *reinterpret_cast<const float4*>(vr + 8 * j);
            const float4 c4 = *reinterpret_cast<const float4*>(vr + 8 * j + 4);
            uint32_t h[4], l[4];
            split2(a.x * sc, a.y * sc, h[0], l[0]); split2(a.z * sc, a.w * sc, h[1], l[1]);
            split2(c4.x * sc, c4.y * sc, h[2], l[2]); split2(c4.z * sc, c4.w * sc, h[3], l[3]);
            *reinterpret_cast<uint4*>(oh + 8 * j) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(ol + 8 * j) = make_uint4(l[0], l[1], l[2], l[3]);
        }
    }
}

// ------------------------------------------------------------------ main kernel
#ifndef ALB200_DBG_BUILD
#define ALB200_DBG_BUILD 0
#endif
constexpr bool kDbg = ALB200_DBG_BUILD != 0;   // developer aid: clock64 stamps of CTA 0's roles (tools/nc_timeline.py)
constexpr int kDbgSlots = 64;                 // events kept per role
// role r, event e (chunk / tile counter), field f:  dbg[(r * kDbgSlots + e) * 4 + f]
#define NC_STAMP(role, ev, field) do { if (kDbg && p.dbg != nullptr && blockIdx.x == 0 && (ev) < (uint32_t)kDbgSlots) \
        p.dbg[((role) * kDbgSlots + (ev)) * 4 + (field)] = clock64(); } while (0)

struct V2Params {
    long long* dbg;
    int* ready;                // pipelined with the search: ready[b * n_mtiles + mel tile] = epoch once the tile is in global memory
    int epoch;
    int tile_major;            // work order: all utterances' mel tile 0, then tile 1, ... (what a concurrent search consumes first)
    const float* a_src;        // raw mel side: z / queries [b, C, Ty]   (slow exact path only; the fast path reads it through TMA)
    const float* b_src0;       // raw text side: m_p / keys [b, C, Tx]    (slow exact path only)
    const float* b_src1;       // logs_p (gaussian)
    const float* colv;         // [b, NT] colterm (gaussian) / -T |k|^2 (ota), from nc_prep_kernel (row stride NT keeps float4 loads aligned)
    const float* inv_sb;       // [b, NT] inverse row scale (gaussian) / 2 T / row scale (ota)
    const float* prior;        // ota: optional [b, Tx, Ty]
    const int32_t* x_lengths;  // ota: optional [b]
    const int32_t* y_lengths;  // ota, generated prior: optional [b] mel lengths
    float prior_scaling;       // ota: > 0 = beta-binomial prior generated in the epilogue instead of read from `prior`
    float* out;                // [b, Tx, Ty]
    float temperature;
    int B, C, Tx, Ty;
    int NT;                    // accumulator columns: Tx rounded up to 16
    int NB;                    // text rows per B box / pass: min(NT, 256)
    int npass, nchunks, n_mtiles, n_items;
    int sr;                    // raw stages
    uint32_t off_a, off_b, off_aux, b_stage_bytes;   // raw ring at offset 0
    int sb;                    // text-operand stages: 2, or 1 for texts longer than 256 tokens (a stage then holds both token blocks)
};

__device__ __forceinline__ void item_to_tile(const V2Params& p, int item, int& b, int& mt) {
    if (p.tile_major) { mt = item / p.B; b = item - mt * p.B; }
    else { b = item / p.n_mtiles; mt = item - b * p.n_mtiles; }
}

// OTA epilogue, fast path, for one thread = one mel frame (TMEM lane).  d'[x] = acc * (2T / scale_x) - T |k_x|^2  (|q_y|^2 is
// constant along the softmax axis and cancels).  Pass 1: log-sum-exp over the text axis, one dependent step per 16 tokens (group
// maximum first, then 16 independent exponentials); pass 2: normalise (+ prior tensor) and store.  nparts > 1: the token groups
// of the tile are dealt round-robin to the `nparts` warps that share this TMEM quadrant (texts longer than 256 tokens fill both
// accumulator halves, nothing overlaps the epilogue, so the transform warps help); their partial (max, sum) pairs meet in xch.
__device__ __forceinline__ void ota_fast_part(uint32_t tcol, const float4* colv4, const float4* isb4, float* xch, int part, int nparts, int row,
                                              bool y_ok, int Tx, int Ty, int NT, int tlen, const float* pr, float* ob)
{
    constexpr float L2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
#define NC_LOAD_TERMS2(g)                                                                                                     \
    float cc[16], ss[16];                                                                                                      \
    _Pragma("unroll") for (int j4 = 0; j4 < 4; ++j4) {                                                                         \
        const float4 c4 = colv4[4 * (g) + j4], s4 = isb4[4 * (g) + j4];                                                        \
        cc[4 * j4] = c4.x; cc[4 * j4 + 1] = c4.y; cc[4 * j4 + 2] = c4.z; cc[4 * j4 + 3] = c4.w;                                \
        ss[4 * j4] = s4.x; ss[4 * j4 + 1] = s4.y; ss[4 * j4 + 2] = s4.z; ss[4 * j4 + 3] = s4.w;                                \
    }
    float mx = -INFINITY, sm = 0.f;
    const int gl = (tlen + 15) >> 4, ngroups = NT >> 4;
    for (int g = part; g < gl; g += nparts) {
        NC_LOAD_TERMS2(g)
        uint32_t r[16];
        tmem_ld16(tcol + (uint32_t)(16 * g), r);
        float d[16], gm = -INFINITY;
        if (16 * g + 16 <= tlen) {
#pragma unroll
            for (int j = 0; j < 16; ++j) { d[j] = fmaf(__uint_as_float(r[j]), ss[j], cc[j]); gm = fmaxf(gm, d[j]); }
        } else {
            const int nv = tlen - 16 * g;                          // >= 1
#pragma unroll
            for (int j = 0; j < 16; ++j) { d[j] = (j < nv) ? fmaf(__uint_as_float(r[j]), ss[j], cc[j]) : -INFINITY; gm = fmaxf(gm, d[j]); }
        }
        const float nm = fmaxf(mx, gm);
        const float nml = nm * L2E;
        float psum = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) psum += fast_exp2(fmaf(d[j], L2E, -nml));   // 2^-inf = 0 for the padding
        sm = fmaf(sm, fast_exp2((mx - nm) * L2E), psum);
        mx = nm;
    }
    if (nparts > 1) {
        xch[(part * BM + row) * 2] = mx; xch[(part * BM + row) * 2 + 1] = sm;
        asm volatile("barrier.sync 2, 384;" ::: "memory");
        float gm = -INFINITY;
        for (int q = 0; q < nparts; ++q) gm = fmaxf(gm, xch[(q * BM + row) * 2]);
        float tot = 0.f;
        for (int q = 0; q < nparts; ++q) {
            const float mq = xch[(q * BM + row) * 2], sq = xch[(q * BM + row) * 2 + 1];
            if (sq > 0.f) tot = fmaf(sq, fast_exp2((mq - gm) * L2E), tot);      // a part without tokens holds (-inf, 0)
        }
        mx = gm; sm = tot;
    }
    const float lse = mx + fast_log2(sm) * LN2;
    for (int g = part; g < ngroups; g += nparts) {
        NC_LOAD_TERMS2(g)
        uint32_t r[16];
        tmem_ld16(tcol + (uint32_t)(16 * g), r);
        if (y_ok) {
            float* o = ob + (size_t)(16 * g) * Ty;
            if (16 * g + 16 <= tlen && pr == nullptr) {
#pragma unroll
                for (int j = 0; j < 16; ++j) { *o = fmaf(__uint_as_float(r[j]), ss[j], cc[j]) - lse; o += Ty; }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int xl = 16 * g + j;
                    if (xl < Tx) {
                        float v = -INFINITY;        // text padding is excluded from the softmax
                        if (xl < tlen) {
                            v = fmaf(__uint_as_float(r[j]), ss[j], cc[j]) - lse;
                            if (pr) v += __logf(pr[(size_t)xl * Ty] + 1e-8f);
                        }
                        *o = v; o += Ty;
                    }
                }
            }
        }
    }
#undef NC_LOAD_TERMS2
}

// BB: OTA with the beta-binomial prior generated in the epilogue -- its own instance, so that the fp64 lgamma code does not weigh on
// the registers of the plain OTA kernel
template <int MODE, bool BB = false>
__global__ void __launch_bounds__(NTHREADS, 1) nc_v2_kernel(const V2Params p, const __grid_constant__ CUtensorMap map_a,
                                                            const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
                                                            const __grid_constant__ CUtensorMap map_bhi2, const __grid_constant__ CUtensorMap map_blo2)
{
    constexpr int RPC = MODE == 0 ? 1 : 2;                   // raw boxes per 64-k chunk (gaussian: 2 k per channel)
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t base = smem_u32(smem);
    const uint32_t raw0 = base, a0 = base + p.off_a, b0 = base + p.off_b, aux = base + p.off_aux;
    // aux: barriers (8 bytes each), then the tmem slot, then per-tile flags and |q|^2 rows
    const uint32_t bar_raw_full = aux, bar_raw_empty = aux + 32, bar_a_full = aux + 64, bar_b_full = aux + 80, bar_a_empty = aux + 96,
                   bar_t_full = aux + 112, bar_t_empty = aux + 128, tmem_slot = aux + 144, bar_b_empty = aux + 192;
    volatile int* ovf = reinterpret_cast<volatile int*>(smem + p.off_aux + 160);          // [8] per-tile "mel side out of fp16 range"
    const int SR = p.sr;
    const uint32_t SB = (uint32_t)p.sb;

    if ((base & 1023u) != 0u) __trap();                      // the swizzle atoms need 1024-byte alignment
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // a search launched behind us may start beside us (it polls p.ready)
    if (kDbg && p.ready != nullptr && tid == 0) {
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        atomicMin(reinterpret_cast<unsigned long long*>(p.ready + (size_t)p.B * p.n_mtiles + 64) + 0, gt);
    }
    if (wid == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < 4; ++s) { mbar_init(bar_raw_full + 8 * s, 1); mbar_init(bar_raw_empty + 8 * s, N_TRANSFORM / 32); }
        for (int s = 0; s < SA; ++s) { mbar_init(bar_a_full + 8 * s, N_TRANSFORM); mbar_init(bar_b_full + 8 * s, 1); mbar_init(bar_a_empty + 8 * s, 1); mbar_init(bar_b_empty + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(bar_t_full + 8 * s, 1); mbar_init(bar_t_empty + 8 * s, (MODE == 1 && p.npass > 1 && !BB) ? 384 : 128); }
        for (int s = 0; s < 8; ++s) ovf[s] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    const int K = MODE == 0 ? 2 * p.C : p.C;
    const int two_slots = (p.npass == 1);
    // OTA with a text longer than 256 tokens: one accumulator spanning both TMEM halves, so the epilogue overlaps nothing and the
    // eight transform warps (idle by then) take two thirds of it (not with the generated prior: its recurrence walks every token)
    const bool helpers = (MODE == 1) && !two_slots && !BB;
    float* xch = reinterpret_cast<float*>(smem + p.off_aux + 256 + 2 * 512 * 4);

    if (wid == 0) {
        // ================= TMA producer, mel side: raw z / q boxes (HBM latency: runs as far ahead as the ring allows) =================
        if (lane == 0) {
            uint32_t rc = 0, rs = 0, rph = 0;                 // ring cursor (no divisions by the runtime stage count)
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                int b, mt;
                item_to_tile(p, item, b, mt);
                const int y0 = mt * BM;
                // (a text longer than 256 tokens is two token blocks of the SAME mel-side operand: it is loaded and converted once)
                    for (int ch = 0; ch < p.nchunks * RPC; ++ch, ++rc) {
                        NC_STAMP(0, rc, 0);
                        mbar_wait(bar_raw_empty + 8 * rs, rph ^ 1u);
                        NC_STAMP(0, rc, 1);
                        mbar_expect_tx(bar_raw_full + 8 * rs, RAW_BYTES);
                        tma_load_3d(raw0 + rs * RAW_BYTES, &map_a, y0, ch * RAW_CH, b, bar_raw_full + 8 * rs);
                        if (++rs == (uint32_t)SR) { rs = 0; rph ^= 1u; }
                    }
            }
        }
    } else if (wid == 2) {
        // ================= TMA producer, text side: scaled fp16 hi / lo boxes from the L2-resident operand =================
        if (lane == 0) {
            // programmatic dependent launch: this kernel may have started while nc_prep_kernel was still running (its prologue and
            // the mel-side loads / conversion do not depend on it); the text-side operand must be complete before it is read
            asm volatile("griddepcontrol.wait;" ::: "memory");
            uint32_t cc = 0, s = 0, ph = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                int b, mt;
                item_to_tile(p, item, b, mt);
                    for (int ch = 0; ch < p.nchunks; ++ch, ++cc) {
                        NC_STAMP(1, cc, 0);
                        mbar_wait(bar_b_empty + 8 * s, ph ^ 1u);
                        NC_STAMP(1, cc, 1);
                        mbar_expect_tx(bar_b_full + 8 * s, 2u * (uint32_t)p.NT * 128u);
                        // stage layout: hi rows [0, NB) | hi rows [256, NT) | lo rows [0, NB) | lo rows [256, NT)
                        const uint32_t bs = b0 + s * p.b_stage_bytes, lo = (uint32_t)p.NT * 128u;
                        tma_load_3d(bs, &map_bhi, ch * KC, 0, b, bar_b_full + 8 * s);
                        tma_load_3d(bs + lo, &map_blo, ch * KC, 0, b, bar_b_full + 8 * s);
                        if (p.npass > 1) {
                            tma_load_3d(bs + 256u * 128u, &map_bhi2, ch * KC, 256, b, bar_b_full + 8 * s);
                            tma_load_3d(bs + lo + 256u * 128u, &map_blo2, ch * KC, 256, b, bar_b_full + 8 * s);
                        }
                        if (++s == SB) { s = 0; ph ^= 1u; }
                    }
            }
        }
    } else if (wid == 1) {
        // ================= MMA issuer (one thread) =================
        if (lane == 0) {
            uint32_t cc = 0, it = 0;
            uint32_t s = 0, ph = 0, sbs = 0, bph = 0;         // A and text-operand ring cursors (no divisions on this thread's critical path)
            bool ready = false;
            // operand descriptors of every stage, formed once (a k-step is +2 on the 16-byte-unit address field)
            const uint64_t dA0h = make_desc(a0), dA0l = make_desc(a0 + A_TILE), dA1h = make_desc(a0 + 2u * A_TILE), dA1l = make_desc(a0 + 3u * A_TILE);
            const uint64_t dB0h = make_desc(b0), dB0l = make_desc(b0 + (uint32_t)p.NT * 128u);
            const uint64_t dB1h = make_desc(b0 + p.b_stage_bytes), dB1l = make_desc(b0 + p.b_stage_bytes + (uint32_t)p.NT * 128u);
            const uint32_t idesc0 = make_idesc(min(256, p.NT)), idesc1 = make_idesc(p.NT > 256 ? p.NT - 256 : 16);
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
                const uint32_t slot = two_slots ? (it & 1u) : 0u, use = two_slots ? (it >> 1) : it;
                mbar_wait(bar_t_empty + 8 * slot, (use & 1u) ^ 1u);          // the epilogue has drained this accumulator
                fence_after();
                for (int ch = 0; ch < p.nchunks; ++ch, ++cc) {
                    NC_STAMP(2, cc, 0);
                    if (!ready) {                                                   // (normally already waited for, see below)
                        mbar_wait(bar_a_full + 8 * s, ph);
                        NC_STAMP(2, cc, 1);
                        mbar_wait(bar_b_full + 8 * sbs, bph);
                    }
                    ready = false;
                    NC_STAMP(2, cc, 2);
                    fence_after();
                    const int nk = min(4, (K - ch * KC + 15) >> 4);                 // 16-k steps that hold real channels
                    const uint64_t dAh = s ? dA1h : dA0h, dAl = s ? dA1l : dA0l;
                    // a text longer than 256 tokens: two token blocks = two MMA series into the two TMEM halves, same A stage
                    for (int pass = 0; pass < p.npass; ++pass) {
                        const uint32_t idesc = pass ? idesc1 : idesc0;
                        const uint32_t dcol = tmem_base + slot * 256u + (uint32_t)pass * 256u;
                        // (second token block: 256 rows = 2048 sixteen-byte units further into the stage)
                        const uint64_t dBh = (sbs ? dB1h : dB0h) + (pass ? 2048u : 0u), dBl = (sbs ? dB1l : dB0l) + (pass ? 2048u : 0u);
                        if (nk == 4) {
                            mma_f16(dcol, dAh, dBh, idesc, ch ? 1u : 0u);                                  // hi * hi
                            mma_f16_acc(dcol, dAh + 2, dBh + 2, idesc); mma_f16_acc(dcol, dAh + 4, dBh + 4, idesc); mma_f16_acc(dcol, dAh + 6, dBh + 6, idesc);
                            mma_f16_acc(dcol, dAl, dBh, idesc); mma_f16_acc(dcol, dAl + 2, dBh + 2, idesc);  // lo * hi
                            mma_f16_acc(dcol, dAl + 4, dBh + 4, idesc); mma_f16_acc(dcol, dAl + 6, dBh + 6, idesc);
                        } else {
                            for (int ks = 0; ks < nk; ++ks) mma_f16(dcol, dAh + 2 * ks, dBh + 2 * ks, idesc, (ch | ks) ? 1u : 0u);
                            for (int ks = 0; ks < nk; ++ks) mma_f16_acc(dcol, dAl + 2 * ks, dBh + 2 * ks, idesc);
                        }
#ifndef NC_NO_LOOKAHEAD
                        if (pass == p.npass - 1 && ch != p.nchunks - 1 && SB > 1) {
                            // the tensor pipe still has this chunk's first two products queued: the barrier round trips of the
                            // NEXT chunk (same tile) hide behind them instead of opening a gap between the chunks
                            const uint32_t s2 = (s + 1 == (uint32_t)SA) ? 0u : s + 1, ph2 = (s + 1 == (uint32_t)SA) ? ph ^ 1u : ph;
                            const uint32_t sb2 = (sbs + 1 == SB) ? 0u : sbs + 1, bph2 = (sbs + 1 == SB) ? bph ^ 1u : bph;
                            mbar_wait(bar_a_full + 8 * s2, ph2);
                            mbar_wait(bar_b_full + 8 * sb2, bph2);
                            ready = true;
                        }
#endif
                        if (nk == 4) {                                                                     // hi * lo
                            mma_f16_acc(dcol, dAh, dBl, idesc); mma_f16_acc(dcol, dAh + 2, dBl + 2, idesc);
                            mma_f16_acc(dcol, dAh + 4, dBl + 4, idesc); mma_f16_acc(dcol, dAh + 6, dBl + 6, idesc);
                        } else {
                            for (int ks = 0; ks < nk; ++ks) mma_f16_acc(dcol, dAh + 2 * ks, dBl + 2 * ks, idesc);
                        }
                    }
                    mma_commit(bar_a_empty + 8 * s);                                // both stages reusable when these MMAs have read them
                    mma_commit(bar_b_empty + 8 * sbs);
                    NC_STAMP(2, cc, 3);
                    if (++s == (uint32_t)SA) { s = 0; ph ^= 1u; }
                    if (++sbs == SB) { sbs = 0; bph ^= 1u; }
                }
                mma_commit(bar_t_full + 8 * slot);                                  // accumulator complete
            }
        }
    } else if (wid < 11) {
        // ================= transform warps: raw box -> scaled fp16 hi / lo A tiles =================
        const int t = tid - 96;                              // 0..255
        const int row = t & 127, half = t >> 7;              // frame row of the tile; which 16 channels of a raw box
        const uint32_t a_row = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u, a_x = (uint32_t)(row & 7);
        // The converted chunk lives in registers (hi / lo, eight 16-byte pieces) while its shared-memory stage is still being read by
        // the tensor core: what remains on the critical path after "stage free" is eight stores, a proxy fence and an arrival.
        uint32_t rc = 0, cc = 0, it = 0;
        uint32_t rs = 0, rph = 0, s = 0, ph = 0;             // ring cursors
        uint32_t H[RPC * (MODE == 0 ? 4 : 2)][4], L[RPC * (MODE == 0 ? 4 : 2)][4];
        const uint32_t total_chunks = (uint32_t)p.nchunks;          // (the mel-side operand of a tile is produced once, whatever the text length)
        bool bad = false;
        auto produce = [&]() {                               // raw boxes of the next chunk -> H / L
#pragma unroll
            for (int r = 0; r < RPC; ++r, ++rc) {
                if (t == 0) NC_STAMP(3, cc, 0);
                mbar_wait(bar_raw_full + 8 * rs, rph);
                if (t == 0) NC_STAMP(3, cc, 1);
                float v[16];
                const uint32_t src = raw0 + rs * RAW_BYTES + (uint32_t)(16 * half) * (BM * 4) + (uint32_t)row * 4u;
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = lds32(src + q * (BM * 4));
#pragma unroll
                for (int q = 0; q < 16; ++q) bad = bad || !(fabsf(v[q]) < (MODE == 0 ? kZLimit : kQLimit));   // also catches NaN
                // The stage may be handed back only when the loads have RETURNED, not merely been issued: the barrier arrival is not
                // ordered behind outstanding shared-memory loads, and under a saturated shared-memory pipe (the tensor core reads its
                // operands there) a load can still be queued when the next TMA box lands -- seen as sporadic stale frames with a
                // one-stage ring.  The range check above consumes every loaded value, the warp issues in order, and the arrival takes
                // the check's result as an operand so that the compiler cannot sink the check below it.
                __syncwarp();
                if (lane == 0) asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.shared::cta.b64 t, [%0];\n\t}" ::"r"(bar_raw_empty + 8 * rs), "r"((int)bad) : "memory");
                if (MODE == 0) {
                    // 16 channels -> 32 k = four 16-byte pieces (4 channels each) of hi and of lo; piece index 4*half + j
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float zz = v[4 * j + e];
                            split2(kA1 * zz * zz, kA2 * zz, H[j][e], L[j][e]);
                        }
                } else {
                    // 16 channels -> 16 k = two 16-byte pieces; piece index 4*r + 2*half + j
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int e = 0; e < 4; ++e) split2(kA2 * v[8 * j + 2 * e], kA2 * v[8 * j + 2 * e + 1], H[2 * r + j][e], L[2 * r + j][e]);
                }
                if (++rs == (uint32_t)SR) { rs = 0; rph ^= 1u; }
            }
        };
        if (blockIdx.x < p.n_items) produce();
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
            if (t == 0) ovf[(it + 4) & 7] = 0;                // nobody is within four tiles of that slot
            for (uint32_t ci = 0; ci < total_chunks; ++ci, ++cc) {
                const uint32_t sA_hi = a0 + s * 2u * A_TILE + a_row, sA_lo = sA_hi + A_TILE;
                mbar_wait(bar_a_empty + 8 * s, ph ^ 1u);
                if (t == 0) NC_STAMP(3, cc, 2);
                if (MODE == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t off = (((uint32_t)(4 * half + j)) ^ a_x) << 4;
                        sts128u(sA_hi + off, H[j][0], H[j][1], H[j][2], H[j][3]);
                        sts128u(sA_lo + off, L[j][0], L[j][1], L[j][2], L[j][3]);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < RPC; ++r)
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint32_t off = (((uint32_t)(4 * r + 2 * half + j)) ^ a_x) << 4;
                            sts128u(sA_hi + off, H[2 * r + j][0], H[2 * r + j][1], H[2 * r + j][2], H[2 * r + j][3]);
                            sts128u(sA_lo + off, L[2 * r + j][0], L[2 * r + j][1], L[2 * r + j][2], L[2 * r + j][3]);
                        }
                }
                // per-tile side band for the epilogue, published BEFORE the arrival that lets the MMAs of the chunk start
                // (ordered by the mbarrier chain a_full -> MMA -> commit -> t_full)
                if (bad) ovf[it & 7] = 1;
                fence_proxy_async();                                          // generic-proxy stores -> visible to the tensor core
                mbar_arrive(bar_a_full + 8 * s);
                if (kDbg && p.dbg != nullptr && blockIdx.x == 0 && cc < (uint32_t)kDbgSlots)
                    atomicMax(reinterpret_cast<unsigned long long*>(p.dbg) + (3 * kDbgSlots + cc) * 4 + 3, (unsigned long long)clock64());   // LAST arrival
                if (++s == (uint32_t)SA) { s = 0; ph ^= 1u; }
                const bool last_of_item = (ci + 1 == total_chunks);
                if (last_of_item) bad = false;                                // the next chunk belongs to the next tile
                if (!last_of_item || item + (int)gridDim.x < p.n_items) produce();
            }
            if (MODE == 1 && helpers) {
                // this tile's operand is complete (and the next tile's first chunk sits converted in registers): help with its epilogue
                int b, mt;
                item_to_tile(p, item, b, mt);
                const int q4 = wid & 3, hrow = 32 * q4 + lane, y = mt * BM + hrow;
                const bool y_ok = y < p.Ty;
                const float* tokc = reinterpret_cast<const float*>(smem + p.off_aux + 256);
                // (barrier.sync, not bar.sync: bar.sync is the .aligned form, and compute-sanitizer synccheck reports these warps as not
                //  converged here after the chunk loop's lane-dependent paths; every named barrier of the 384 threads uses this form)
                asm volatile("barrier.sync 2, 384;" ::: "memory");                 // the epilogue warps have staged the per-token terms
                mbar_wait(bar_t_full, it & 1u);
                fence_after();
                const bool slow = ovf[it & 7] != 0;
                const int tlen = p.x_lengths ? min(max(p.x_lengths[b], 0), p.Tx) : p.Tx;
                if (!slow)
                    ota_fast_part(tmem_base + ((uint32_t)(32 * q4) << 16), reinterpret_cast<const float4*>(tokc), reinterpret_cast<const float4*>(tokc + 512), xch,
                                  (wid - 3) >> 2, 3, hrow, y_ok, p.Tx, p.Ty, p.NT, tlen,
                                  p.prior ? p.prior + (size_t)b * p.Tx * p.Ty + (y_ok ? y : 0) : nullptr, p.out + (size_t)b * p.Tx * p.Ty + (y_ok ? y : 0));
                else
                    asm volatile("barrier.sync 2, 384;" ::: "memory");             // (the exact path is the epilogue warps' alone; keep the barrier count)
                fence_before();
                mbar_arrive(bar_t_empty);
                if (p.ready != nullptr) __threadfence();
                asm volatile("barrier.sync 2, 384;" ::: "memory");                 // everybody is done with the per-token terms and the partials
            }
        }
    } else {
        // ================= epilogue warps: TMEM -> registers -> global =================
        const int q = wid & 3;                               // TMEM lane quadrant this warp may read
        const int row = 32 * q + lane;
        const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
        float* tokc = reinterpret_cast<float*>(smem + p.off_aux + 256);        // [512] colterm / -T |k|^2 of the tile's utterance
        float* toks = tokc + 512;                                              // [512] inverse row scale / 2 T / row scale
        asm volatile("griddepcontrol.wait;" ::: "memory");                     // the per-token terms come from nc_prep_kernel too
        uint32_t it = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
            int b, mt;
            item_to_tile(p, item, b, mt);
            const int y0 = mt * BM;
            const uint32_t slot = two_slots ? (it & 1u) : 0u, use = two_slots ? (it >> 1) : it;
            const int y = y0 + row;
            const bool y_ok = y < p.Ty;
            const int Tx = p.Tx, Ty = p.Ty;
            // the tile's per-token terms -> shared memory, while the tensor core is still working on the tile
            {
                const int et = tid - 11 * 32;                                  // 0..127
                const float* gc = p.colv + (size_t)b * p.NT;
                const float* gs = p.inv_sb + (size_t)b * p.NT;
                for (int i = et; i < p.NT; i += 128) { tokc[i] = __ldg(gc + i); toks[i] = __ldg(gs + i); }
                if (helpers) asm volatile("barrier.sync 2, 384;" ::: "memory");
                else asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            const float4* colv4 = reinterpret_cast<const float4*>(tokc);
            const float4* isb4 = reinterpret_cast<const float4*>(toks);
            float* ob = p.out + (size_t)b * Tx * Ty + (y_ok ? y : 0);
            if (wid == 11 && lane == 0) NC_STAMP(4, it, 0);
            mbar_wait(bar_t_full + 8 * slot, use & 1u);
            if (wid == 11 && lane == 0) NC_STAMP(4, it, 1);
            fence_after();
            const uint32_t tcol = tlane + slot * 256u;
            const bool slow = ovf[it & 7] != 0;              // the mel side of this tile left the fp16 range: exact fp32 path below
            const int ngroups = p.NT >> 4;
#ifdef NC_CHECK_TMEM
            {   // developer check: the accumulator must not change while the epilogue owns it
                uint32_t h1 = 0, h2 = 0;
                for (int g = 0; g < ngroups; ++g) { uint32_t r[16]; tmem_ld16(tcol + (uint32_t)(16 * g), r); for (int j = 0; j < 16; ++j) h1 = h1 * 31u + r[j]; }
                for (int spin = 0; spin < 2000; ++spin) asm volatile("nanosleep.u32 20;");
                for (int g = 0; g < ngroups; ++g) { uint32_t r[16]; tmem_ld16(tcol + (uint32_t)(16 * g), r); for (int j = 0; j < 16; ++j) h2 = h2 * 31u + r[j]; }
                if (h1 != h2) { printf("TMEM changed under the epilogue: block %d item %d row %d\n", (int)blockIdx.x, item, row); }
            }
#endif
            // Full groups of 16 tokens take a branch-free path (one FFMA [+ one add] and one 128-byte-per-warp store per cell); the
            // ragged last group takes the predicated one.
#define NC_LOAD_TERMS(g)                                                                                                      \
            float cc[16], ss[16];                                                                                              \
            _Pragma("unroll") for (int j4 = 0; j4 < 4; ++j4) {                                                                 \
                const float4 c4 = colv4[4 * (g) + j4], s4 = isb4[4 * (g) + j4];                                                \
                cc[4 * j4] = c4.x; cc[4 * j4 + 1] = c4.y; cc[4 * j4 + 2] = c4.z; cc[4 * j4 + 3] = c4.w;                        \
                ss[4 * j4] = s4.x; ss[4 * j4 + 1] = s4.y; ss[4 * j4 + 2] = s4.z; ss[4 * j4 + 3] = s4.w;                        \
            }
            if (MODE == 0) {
                if (!slow) {
                    float* o = ob;
                    for (int g = 0; g < ngroups; ++g) {
                        NC_LOAD_TERMS(g)
                        uint32_t r[16];
                        if (wid == 11 && lane == 0 && it == 1) NC_STAMP(5, (uint32_t)g, 0);
                        tmem_ld16(tcol + (uint32_t)(16 * g), r);
                        if (wid == 11 && lane == 0 && it == 1) NC_STAMP(5, (uint32_t)g, 1);
                        if (y_ok) {
                            if (16 * g + 16 <= Tx) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) { *o = fmaf(__uint_as_float(r[j]), ss[j], cc[j]); o += Ty; }
                            } else {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    if (16 * g + j < Tx) { *o = fmaf(__uint_as_float(r[j]), ss[j], cc[j]); o += Ty; }
                            }
                        }
                        if (wid == 11 && lane == 0 && it == 1) NC_STAMP(5, (uint32_t)g, 2);
                    }
                } else if (y_ok) {
                    // exact fp32 from the raw inputs, fixed order (same expression as the CUDA-core kernel in neg_cent.cu)
                    const float* zb = p.a_src + (size_t)b * p.C * Ty + y;
                    for (int x = 0; x < Tx; ++x) {
                        const float* mb = p.b_src0 + (size_t)b * p.C * Tx + x;
                        const float* lb = p.b_src1 + (size_t)b * p.C * Tx + x;
                        float acc = 0.f, ct = 0.f;
                        for (int c = 0; c < p.C; ++c) {
                            const float zz = zb[(size_t)c * Ty], mm = __ldg(mb + (size_t)c * Tx), lg = __ldg(lb + (size_t)c * Tx);
                            const float s2 = expf(-2.f * lg), ms2 = mm * s2;
                            ct += (-0.9189385332046727f - lg) - 0.5f * mm * ms2;
                            acc = fmaf(s2, -0.5f * zz * zz, acc);
                            acc = fmaf(ms2, zz, acc);
                        }
                        ob[(size_t)x * Ty] = acc + ct;
                    }
                }
            } else {
                const int tlen = p.x_lengths ? min(max(p.x_lengths[b], 0), Tx) : Tx;
                const float* pr = p.prior ? p.prior + (size_t)b * Tx * Ty + (y_ok ? y : 0) : nullptr;
                constexpr float L2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
                if (!slow && !BB) {
                    ota_fast_part(tcol, colv4, isb4, xch, helpers ? 2 : 0, helpers ? 3 : 1, row, y_ok, Tx, Ty, p.NT, tlen, pr, ob);
                } else if (!slow) {
                    // generated beta-binomial prior: this thread walks every token in order
                    float mx = -INFINITY, sm = 0.f;
                    const int gl = (tlen + 15) >> 4;
                    for (int g = 0; g < gl; ++g) {
                        NC_LOAD_TERMS(g)
                        uint32_t r[16];
                        tmem_ld16(tcol + (uint32_t)(16 * g), r);
                        float d[16], gm = -INFINITY;
                        const int nv = tlen - 16 * g;                              // >= 1
#pragma unroll
                        for (int j = 0; j < 16; ++j) { d[j] = (j < nv) ? fmaf(__uint_as_float(r[j]), ss[j], cc[j]) : -INFINITY; gm = fmaxf(gm, d[j]); }
                        const float nm = fmaxf(mx, gm);
                        const float nml = nm * L2E;
                        float psum = 0.f;
#pragma unroll
                        for (int j = 0; j < 16; ++j) psum += fast_exp2(fmaf(d[j], L2E, -nml));
                        sm = fmaf(sm, fast_exp2((mx - nm) * L2E), psum);
                        mx = nm;
                    }
                    const float lse = mx + fast_log2(sm) * LN2;
                    float* o = ob;
                    {
                        // Beta-binomial alignment prior of the OTA paper, generated here instead of read from a [b, t_x, t_y] tensor:
                        //   prior[x, y] = BetaBinom(x; n = t_x - 1, a = s (y + 1), b = s (t_y - y))        (oracle/neg_cent.py:beta_binomial_prior)
                        // This thread owns frame y and walks the tokens in order, so the pmf is a recurrence along x,
                        //   pmf(x + 1) / pmf(x) = (n - x)(x + a) / ((x + 1)(n - x - 1 + b)),   pmf(0) = B(a, n + b) / B(a, b),
                        // one division and one logarithm per cell; the start value needs four lgamma in fp64 per frame (their sizes,
                        // ~1e4, cancel to ~1e1: fp32 would lose the 1e-5 bound).  Frames past t_y get prior 0, i.e. log(1e-8), what a
                        // zero-padded prior tensor gives.
                        const int tyl = p.y_lengths ? min(max(p.y_lengths[b], 0), Ty) : Ty;
                        const int n = tlen - 1;
                        const bool inb = y < tyl;
                        const float af = p.prior_scaling * (float)(y + 1), bf = p.prior_scaling * (float)(tyl - y);
                        double lp = 0.0;                                          // fp64 running sum: 300 fp32 roundings at |lp| ~ 50 would cost the 1e-5 bound
                        if (inb && n > 0) {
                            const double a = (double)p.prior_scaling * (double)(y + 1), bq = (double)p.prior_scaling * (double)(tyl - y);
                            lp = lgamma((double)n + bq) + lgamma(a + bq) - lgamma((double)n + a + bq) - lgamma(bq);
                        }
                        const float lzero = -18.420680743952367f;                 // log(1e-8)
                        for (int g = 0; g < ngroups; ++g) {
                            NC_LOAD_TERMS(g)
                            uint32_t r[16];
                            tmem_ld16(tcol + (uint32_t)(16 * g), r);
                            if (y_ok) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    const int xl = 16 * g + j;
                                    if (xl < Tx) {
                                        float v = -INFINITY;        // text padding is excluded from the softmax
                                        if (xl < tlen) {
                                            v = fmaf(__uint_as_float(r[j]), ss[j], cc[j]) - lse + (inb ? logf(expf((float)lp) + 1e-8f) : lzero);
                                            const float num = (float)(n - xl) * ((float)xl + af), den = (float)(xl + 1) * ((float)(n - xl - 1) + bf);
                                            lp += (double)logf(num / den);        // x = n: log(0) = -inf, never used again
                                        }
                                        *o = v; o += Ty;
                                    }
                                }
                            }
                        }
                    }
                } else {
                  if (helpers) asm volatile("barrier.sync 2, 384;" ::: "memory");   // the helpers skip an exact tile; keep the barrier count
                  if (y_ok) {
                    // exact fp32 from the raw inputs: squared distance from differences, two passes over the text axis
                    const float T = p.temperature;
                    const float* qb = p.a_src + (size_t)b * p.C * Ty + y;
                    float mx = -INFINITY, sm = 0.f;
                    for (int ps = 0; ps < 2; ++ps) {
                        const float lse = ps ? mx + logf(sm) : 0.f;
                        for (int x = 0; x < (ps ? Tx : tlen); ++x) {
                            float v = -INFINITY;
                            if (x < tlen) {
                                const float* kb = p.b_src0 + (size_t)b * p.C * Tx + x;
                                float d2 = 0.f;
                                for (int c = 0; c < p.C; ++c) { const float e = qb[(size_t)c * Ty] - __ldg(kb + (size_t)c * Tx); d2 = fmaf(e, e, d2); }
                                v = -T * d2;
                            }
                            if (!ps) { const float nm = fmaxf(mx, v); sm = sm * expf(mx - nm) + expf(v - nm); mx = nm; }
                            else {
                                if (x < tlen) {
                                    v -= lse;
                                    if (pr) v += logf(pr[(size_t)x * Ty] + 1e-8f);
                                    if (BB) {                                        // generated prior, straight from the definition (exact path: speed is irrelevant)
                                        const int tyl = p.y_lengths ? min(max(p.y_lengths[b], 0), Ty) : Ty;
                                        const int n = tlen - 1;
                                        double pm = 0.0;
                                        if (y < tyl) {
                                            const double a = (double)p.prior_scaling * (y + 1), bq = (double)p.prior_scaling * (tyl - y);
                                            pm = exp(lgamma(n + 1.0) - lgamma(x + 1.0) - lgamma(n - x + 1.0) + lgamma(x + a) + lgamma(n - x + bq) - lgamma(n + a + bq)
                                                     - (lgamma(a) + lgamma(bq) - lgamma(a + bq)));
                                        }
                                        v += (float)log(pm + 1e-8);
                                    }
                                }
                                ob[(size_t)x * Ty] = v;
                            }
                        }
                    }
                  }
                }
            }
#undef NC_LOAD_TERMS
            fence_before();
            mbar_arrive(bar_t_empty + 8 * slot);               // 128 arrivals: the accumulator may be overwritten
            if (wid == 11 && lane == 0) NC_STAMP(4, it, 2);
            if (p.ready != nullptr) __threadfence();          // this thread's part of the tile is visible device-wide ...
            if (helpers) asm volatile("barrier.sync 2, 384;" ::: "memory");
            else asm volatile("bar.sync 1, 128;" ::: "memory");    // every epilogue thread is done with this tile's per-token terms
            if (p.ready != nullptr && wid == 11 && lane == 0)   // ... before the tile is published to the search running beside us
                asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.ready + (size_t)b * p.n_mtiles + mt), "r"(p.epoch) : "memory");
        }
    }
    if (kDbg && p.ready != nullptr && tid == 0) {
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        atomicMax(reinterpret_cast<unsigned long long*>(p.ready + (size_t)p.B * p.n_mtiles + 64) + 1, gt);
    }
    fence_before();
    __syncthreads();
    if (wid == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// ------------------------------------------------------------------ host side
static int fail(int code, const char* who, const char* msg)
{
    snprintf(alb::g_err, sizeof(alb::g_err), "%s: %s", who, msg);
    return code;
}

struct Plan { int K, Kpad, NT, NB, npass, nchunks, sr, sb; uint32_t off_a, off_b, off_aux, b_stage, smem; size_t ws_bhi, ws_blo, ws_colv, ws_isb, ws_total; };

static bool make_plan(int mode, int b, int c, int tx, int ty, Plan* pl)
{
    if (tx > 512 || (ty & 3) != 0 || b < 1) return false;
    pl->K = mode == 0 ? 2 * c : c;
    pl->Kpad = (pl->K + KC - 1) / KC * KC;
    pl->NT = (tx + 15) & ~15;
    pl->NB = pl->NT < 256 ? pl->NT : 256;
    pl->npass = (pl->NT + 255) / 256;
    pl->nchunks = pl->Kpad / KC;
    pl->b_stage = 2u * pl->NT * 128u;                    // hi + lo, every token row of the tile (both 256-token blocks of a long text)
    pl->sb = pl->npass == 1 ? SA : 1;
    // raw ring first (16 KB stages keep everything after it 1024-byte aligned), then A, then B, then aux
    const int budget = 232448 - AUX_BYTES - SA * 2 * A_TILE - pl->sb * (int)pl->b_stage;
    int sr = budget / RAW_BYTES;
    if (sr > 4) sr = 4;
    if (sr < 1) return false;                 // (one stage is enough to run: the transform warps hold the next chunk in registers)
    pl->sr = sr;
    pl->off_a = sr * RAW_BYTES;
    pl->off_b = pl->off_a + SA * 2 * A_TILE;
    pl->off_aux = pl->off_b + pl->sb * pl->b_stage;
    pl->smem = pl->off_aux + AUX_BYTES;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    pl->ws_bhi = 0;
    pl->ws_blo = up((size_t)b * tx * pl->Kpad * 2);
    pl->ws_colv = pl->ws_blo + up((size_t)b * tx * pl->Kpad * 2);
    pl->ws_isb = pl->ws_colv + up((size_t)b * pl->NT * 4);
    pl->ws_total = pl->ws_isb + up((size_t)b * pl->NT * 4);
    return true;
}

static int encode3(CUtensorMap* m, CUtensorMapDataType dt, int esize, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                   CUtensorMapSwizzle sw, const char* who)
{
    alb::TmapEncodeFn enc = alb::tmap_encode_fn();
    if (!enc) return fail(ALB200_E_CUDA, who, "cuTensorMapEncodeTiled is not available in this driver");
    cuuint64_t dims[3] = { d0, d1, d2 };
    cuuint64_t strides[2] = { d0 * (uint64_t)esize, d0 * d1 * (uint64_t)esize };
    cuuint32_t box[3] = { b0, b1, 1 }, es[3] = { 1, 1, 1 };
    CUresult r = enc(m, dt, 3, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(alb::g_err, sizeof(alb::g_err), "%s: cuTensorMapEncodeTiled failed with %d", who, (int)r); return ALB200_E_CUDA; }
    return 0;
}

// cuTensorMapEncodeTiled costs 1-2 us of host time each and the score call needs up to five: the last few are remembered per
// thread (same pointer, shape and box -> same map), and all of them are encoded BEFORE the prep kernel is launched so that the
// two launches of a call go out back to back.
static int encode3_cached(CUtensorMap* m, CUtensorMapDataType dt, int esize, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                          CUtensorMapSwizzle sw, const char* who)
{
    struct Key { const void* ptr; uint64_t d0, d1, d2; uint32_t b0, b1; int dt, sw; };
    static thread_local Key keys[32];
    static thread_local CUtensorMap maps[32];
    static thread_local int used = 0, next = 0;
    const Key k = { ptr, d0, d1, d2, b0, b1, (int)dt, (int)sw };
    for (int i = 0; i < used; ++i)
        if (keys[i].ptr == k.ptr && keys[i].d0 == d0 && keys[i].d1 == d1 && keys[i].d2 == d2 && keys[i].b0 == b0 && keys[i].b1 == b1 && keys[i].dt == k.dt && keys[i].sw == k.sw) {
            *m = maps[i];
            return 0;
        }
    const int rc = encode3(m, dt, esize, ptr, d0, d1, d2, b0, b1, sw, who);
    if (rc) return rc;
    keys[next] = k; maps[next] = *m;
    next = (next + 1) % 32; if (used < 32) ++used;
    return 0;
}

template <int MODE>
static int run(const float* a_src, const float* b_src0, const float* b_src1, const float* prior, const int32_t* x_lengths, float* out, float temperature,
               int b, int c, int tx, int ty, void* workspace, size_t workspace_bytes, cudaStream_t stream, const char* who,
               int* ready = nullptr, int epoch = 0, int max_ctas = 0, const int32_t* y_lengths = nullptr, float prior_scaling = 0.f)
{
    Plan pl;
    if (!make_plan(MODE, b, c, tx, ty, &pl)) return ALB200_E_UNSUPPORTED;
    if (!workspace || workspace_bytes < pl.ws_total) return fail(ALB200_E_INVALID, who, "workspace too small (alb200_neg_cent_workspace_bytes)");
    if ((reinterpret_cast<uintptr_t>(a_src) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 255)) return ALB200_E_UNSUPPORTED;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail(ALB200_E_NO_DEVICE, who, "no CUDA device");
    static thread_local int sm_count[64] = {0};
    static thread_local bool configured[64] = {false};
    if (dev < 0 || dev >= 64) return ALB200_E_UNSUPPORTED;
    if (!configured[dev]) {
        cudaError_t e = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(nc_v2_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e == cudaSuccess && MODE == 1) e = cudaFuncSetAttribute(nc_v2_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(nc_prep_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return fail(ALB200_E_CUDA, who, cudaGetErrorString(e));
        configured[dev] = true;
    }
    sms = sm_count[dev];
    char* ws = reinterpret_cast<char*>(workspace);
    __half* b_hi = reinterpret_cast<__half*>(ws + pl.ws_bhi);
    __half* b_lo = reinterpret_cast<__half*>(ws + pl.ws_blo);
    float* colv = reinterpret_cast<float*>(ws + pl.ws_colv);
    float* isb = reinterpret_cast<float*>(ws + pl.ws_isb);
    const size_t prep_smem = (size_t)PT * (pl.Kpad + 4) * 4;
    if (prep_smem > 200 * 1024) return ALB200_E_UNSUPPORTED;
    CUtensorMap map_a, map_bhi, map_blo, map_bhi2, map_blo2;
    int rc = encode3_cached(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a_src, (uint64_t)ty, (uint64_t)c, (uint64_t)b, BM, RAW_CH, CU_TENSOR_MAP_SWIZZLE_NONE, who);
    if (!rc) rc = encode3_cached(&map_bhi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, b_hi, (uint64_t)pl.Kpad, (uint64_t)tx, (uint64_t)b, KC, (uint32_t)pl.NB, CU_TENSOR_MAP_SWIZZLE_128B, who);
    if (!rc) rc = encode3_cached(&map_blo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, b_lo, (uint64_t)pl.Kpad, (uint64_t)tx, (uint64_t)b, KC, (uint32_t)pl.NB, CU_TENSOR_MAP_SWIZZLE_128B, who);
    map_bhi2 = map_bhi; map_blo2 = map_blo;                   // (one token block: never dereferenced)
    if (pl.npass > 1) {                                        // second token block of a long text: its own, shorter box
        const uint32_t rows2 = (uint32_t)(pl.NT - 256);
        if (!rc) rc = encode3_cached(&map_bhi2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, b_hi, (uint64_t)pl.Kpad, (uint64_t)tx, (uint64_t)b, KC, rows2, CU_TENSOR_MAP_SWIZZLE_128B, who);
        if (!rc) rc = encode3_cached(&map_blo2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, b_lo, (uint64_t)pl.Kpad, (uint64_t)tx, (uint64_t)b, KC, rows2, CU_TENSOR_MAP_SWIZZLE_128B, who);
    }
    if (rc) return rc;
    nc_prep_kernel<MODE><<<dim3((tx + PT - 1) / PT, b), PT * PG, prep_smem, stream>>>(b_src0, b_src1, b_hi, b_lo, colv, isb, c, tx, pl.K, pl.Kpad, pl.NT, temperature);
    ++alb::g_launches;
    V2Params p;
    memset(&p, 0, sizeof(p));
    p.a_src = a_src; p.b_src0 = b_src0; p.b_src1 = b_src1; p.colv = colv; p.inv_sb = isb; p.prior = prior; p.x_lengths = x_lengths; p.out = out;
    p.temperature = temperature; p.B = b; p.C = c; p.Tx = tx; p.Ty = ty; p.NT = pl.NT; p.NB = pl.NB; p.npass = pl.npass; p.nchunks = pl.nchunks;
    p.n_mtiles = (ty + BM - 1) / BM; p.n_items = b * p.n_mtiles; p.sr = pl.sr;
    p.off_a = pl.off_a; p.off_b = pl.off_b; p.off_aux = pl.off_aux; p.b_stage_bytes = pl.b_stage; p.sb = pl.sb;
    p.ready = ready; p.epoch = epoch; p.tile_major = ready != nullptr;
    p.y_lengths = y_lengths; p.prior_scaling = prior_scaling;
    int grid = p.n_items < sms ? p.n_items : sms;
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;     // the rest of the machine belongs to the search running beside us
    static long long* d_dbg = nullptr;
    const size_t dbg_n = 6 * kDbgSlots * 4;
    if (kDbg && alb::opts().dbg == 1) {
        if (!d_dbg) cudaMalloc(&d_dbg, dbg_n * 8);
        cudaMemsetAsync(d_dbg, 0, dbg_n * 8, stream);
        p.dbg = d_dbg;
    }
    {
        // programmatic stream serialization: the score kernel is allowed to start once every CTA of nc_prep_kernel has passed its
        // griddepcontrol.launch_dependents; the two consumers of the prep output wait with griddepcontrol.wait
        cudaLaunchConfig_t lc;
        memset(&lc, 0, sizeof(lc));
        lc.gridDim = dim3(grid); lc.blockDim = dim3(NTHREADS); lc.dynamicSmemBytes = pl.smem; lc.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at; lc.numAttrs = alb::opts().nc_no_pdl ? 0 : 1;
        cudaError_t le = (MODE == 1 && prior_scaling > 0.f) ? cudaLaunchKernelEx(&lc, nc_v2_kernel<1, true>, p, map_a, map_bhi, map_blo, map_bhi2, map_blo2)
                                                             : cudaLaunchKernelEx(&lc, nc_v2_kernel<MODE, false>, p, map_a, map_bhi, map_blo, map_bhi2, map_blo2);
        if (le != cudaSuccess) return fail(ALB200_E_CUDA, who, cudaGetErrorString(le));
    }
    ++alb::g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ALB200_E_CUDA, who, cudaGetErrorString(e));
    if (kDbg && p.dbg) {
        static long long h[6 * kDbgSlots * 4];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, d_dbg, dbg_n * 8, cudaMemcpyDeviceToHost);
        long long t0 = 0x7fffffffffffffffLL;
        for (size_t i = 0; i < dbg_n; ++i) if (h[i] > 0 && h[i] < t0) t0 = h[i];
        static const char* names[6] = { "rawtma", "btma", "mma", "xform", "epi", "epigrp" };
        for (int r = 0; r < 6; ++r)
            for (int ev = 0; ev < kDbgSlots; ++ev) {
                const long long* q = h + (r * kDbgSlots + ev) * 4;
                if (!q[0] && !q[1]) continue;
                fprintf(stderr, "[nc dbg] %-6s %3d :", names[r], ev);
                for (int f = 0; f < 4; ++f) fprintf(stderr, " %8lld", q[f] ? q[f] - t0 : -1);
                fprintf(stderr, "\n");
            }
    }
    return 0;
}

}  // namespace albv2

// Internal entries (neg_cent.cu dispatches here); ALB200_E_UNSUPPORTED = shape / alignment outside this generation, take the next path.
extern "C" size_t alb200_neg_cent_workspace_bytes(int mode, int b, int c, int tx, int ty)
{
    albv2::Plan pl;
    if (mode < 0 || mode > 1 || b <= 0 || c <= 0 || tx <= 0 || ty <= 0 || !albv2::make_plan(mode, b, c, tx, ty, &pl)) return 0;
    return pl.ws_total;
}

extern "C" int alb200_neg_cent_gaussian_v2(const float* z, const float* m_p, const float* logs_p, float* out, int b, int c, int tx, int ty, void* workspace,
                                           size_t workspace_bytes, void* stream)
{
    return albv2::run<0>(z, m_p, logs_p, nullptr, nullptr, out, 0.f, b, c, tx, ty, workspace, workspace_bytes, (cudaStream_t)stream, "neg_cent_gaussian");
}

// Pipelined form for the fused neg_cent -> search entry (mas_api.cu): tiles in tile-major order on at most max_ctas SMs, each
// published in ready[] (value = epoch) as soon as it is in global memory.
extern "C" int alb200_neg_cent_gaussian_v2_pipelined(const float* z, const float* m_p, const float* logs_p, float* out, int b, int c, int tx, int ty,
                                                     void* workspace, size_t workspace_bytes, void* stream, int* ready, int epoch, int max_ctas)
{
    return albv2::run<0>(z, m_p, logs_p, nullptr, nullptr, out, 0.f, b, c, tx, ty, workspace, workspace_bytes, (cudaStream_t)stream, "neg_cent_gaussian",
                         ready, epoch, max_ctas);
}

extern "C" int alb200_neg_cent_ota_v2_pipelined(const float* queries, const float* keys, const float* prior, const int32_t* x_lengths, float* out,
                                                float temperature, int b, int c, int tx, int ty, void* workspace, size_t workspace_bytes, void* stream,
                                                int* ready, int epoch, int max_ctas)
{
    return albv2::run<1>(queries, keys, nullptr, prior, x_lengths, out, temperature, b, c, tx, ty, workspace, workspace_bytes, (cudaStream_t)stream,
                         "neg_cent_ota", ready, epoch, max_ctas);
}

// OTA score with the beta-binomial prior generated in the epilogue (no [b, t_x, t_y] prior read).  ALB200_E_UNSUPPORTED for
// shapes this generation does not take: the caller materialises the prior and uses alb200_neg_cent_ota_ws.
extern "C" int alb200_neg_cent_ota_bb(const float* queries, const float* keys, const int32_t* x_lengths, const int32_t* y_lengths, float prior_scaling,
                                      float* out, float temperature, int b, int c, int tx, int ty, void* workspace, size_t workspace_bytes, void* stream)
{
    if (!queries || !keys || !out || b < 0 || c <= 0 || tx <= 0 || ty <= 0 || !(prior_scaling > 0.f)) {
        snprintf(alb::g_err, sizeof(alb::g_err), "neg_cent_ota_bb: null pointer, bad shape or non-positive prior scaling");
        return ALB200_E_INVALID;
    }
    if (b == 0) return 0;
    if (!workspace || alb::opts().nc_ffma || alb::opts().nc_v1) return ALB200_E_UNSUPPORTED;
    return albv2::run<1>(queries, keys, nullptr, nullptr, x_lengths, out, temperature, b, c, tx, ty, workspace, workspace_bytes, (cudaStream_t)stream,
                         "neg_cent_ota_bb", nullptr, 0, 0, y_lengths, prior_scaling);
}

extern "C" int alb200_neg_cent_ota_v2(const float* queries, const float* keys, const float* prior, const int32_t* x_lengths, float* out, float temperature,
                                      int b, int c, int tx, int ty, void* workspace, size_t workspace_bytes, void* stream)
{
    return albv2::run<1>(queries, keys, nullptr, prior, x_lengths, out, temperature, b, c, tx, ty, workspace, workspace_bytes, (cudaStream_t)stream, "neg_cent_ota");
}
