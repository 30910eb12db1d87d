// mas_kernel.cuh -- monotonic alignment search for sm_100a.
//
// One CTA per utterance (persistent, work-stealing over the batch).  The text
// axis is spread over threads: lane l of compute warp w owns R consecutive token
// rows and keeps their running score column in registers; the DP marches along
// the mel axis.  What the reference does per item (monotonic_align/core.pyx:17-35):
//
//   forward   V[x,y] = max(V[x,y-1], V[x-1,y-1]) + value[x,y]                     (core.pyx:19-30)
//   backtrack walk y = t_y-1 .. 0 choosing x-1 iff V[x,y-1] < V[x-1,y-1]          (core.pyx:32-35)
//
// is restated as: running fp32 column + ONE direction bit per cell (bit =
// v_prev > v_cur, the very predicate the backtrack re-evaluates), then a
// bit-driven backtrack.  oracle/mas_oracle.c:mas_oracle_bits is the CPU twin of
// this formulation and is proven equal to the table form by tests/test_oracle.py.
//
// Structure of a CTA: NW compute warps + NW loader warps (warp specialisation).
//   * loader warp w streams compute warp w's rows: 128-bit asynchronous copies
//     (cp.async.cg, SASS LDGSTS.128) of TF frames per row land in that warp's
//     ring of NS stages; completion is an mbarrier the copies arrive on
//     (cp.async.mbarrier.arrive), release is a second mbarrier the compute lanes
//     arrive on.  Only chunks inside the reference's band (core.pyx:18) are
//     fetched.  Rows sit with a 16-byte skew per lane so the lanes' 128-bit
//     shared loads are bank-conflict free for any R.  The loader also issues the
//     zero fill of the dense output as bulk shared->global stores (TMA engine,
//     SASS UBLKCP) paced along the forward pass -- a fused memset.
//     (Measured history, profiles/r01_notes.md: per-row cp.async.bulk loads cost
//     60-90 cycles of TMA issue per 128-byte row, 8x slower end to end; loads
//     issued by the compute warps themselves doubled their per-frame latency.)
//   * compute warps form a dataflow pipeline: warp w consumes the last row of
//     warp w-1 through a 128-frame shared ring plus a progress flag, 16 or 32
//     frames at a time.  The diagonal band provides the pipeline skew for free (warp w
//     starts at frame 32*R*w); there is no CTA-wide barrier in the forward pass.
//   * forward forms, chosen by the host:
//       lock-step  every lane works on the same frame; the neighbour exchange
//                  (SHFL) sits on the per-frame dependency chain.  No fill cost;
//                  used when several CTAs share an SM and hide each other.
//       skewed     lane l runs ONE frame behind lane l-1 (systolic): the shuffle is
//                  issued two frames before its result is used.  Tiles arrive as
//                  one 2-D TMA box per 32 frames; costs 31 frames of fill per warp
//                  and 32 more per hand-off; used when an utterance has an SM to
//                  itself (forward_unit).
//       skewed, 4-frame lag on pre-skewed 3-D boxes (forward_unit4): fewer
//                  instructions per frame, four times the fill; one long warp only.
//   * direction bits: shifted into a 32-frame word per row, flushed to shared
//     memory when it fits, else to an L2-resident per-CTA slot of the workspace.
//   * backtrack: one warp walks, 32 frames (one direction word per row) per block,
//     one add and one 3-input logic op per row visited; the other warps turn the
//     published (token, step mask) pairs into stores.  The skewed forms store the
//     words ready to walk and the walker reads them in place when they are in
//     shared memory; otherwise a software pipeline copies windows of words ahead.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace alb {

constexpr int kMaxWarps = 8;       // compute warps per CTA (an equal number of loader warps rides along)
constexpr int kRing = 128;         // frames in a warp-boundary ring
constexpr int kRing4 = 256;        // ... of the 4-frame-lag form (its consumer trails the producer by >= 160 steps)
constexpr int kLag4 = 4;           // frames lane l trails lane l-1 in the pre-skewed form (the granularity TMA can shift a row by: 16 bytes)
constexpr int kZeroChunk = 7168;   // bytes per zero-fill bulk store (56 x 128; sized so 2 CTAs x 3 stages still fit an SM at t_x = 400)
constexpr int kLanePad = 16;       // bytes of skew per lane inside a tile stage
constexpr int kProgDone = 0x3fffffff;
#ifndef ALB_SPEC_ADD
#define ALB_SPEC_ADD 1
#endif
#ifndef ALB200_DBG_BUILD
#define ALB200_DBG_BUILD 0
#endif
#ifndef ALB_ABL
#define ALB_ABL 0        // timing ablations of the skewed unit body (wrong results; build_lib.py --variant): 1 no shuffle, 2 no tile loads, 4 no bits, 8 no boundary ring
#endif
constexpr bool kDbgBuild = ALB200_DBG_BUILD != 0;   // per-warp clock64 stamps (developer aid)

struct WsHeader {       // first 64 bytes of the workspace
    int counter;        // work-stealing cursor (self-resetting)
    int done;           // CTAs that left the item loop
    int status;         // bit0: invalid lengths seen
    int zero_cursor;    // shared zero fill: next unclaimed chunk of the dense output (self-resetting, like the two below)
    int zero_done;      // ... chunks whose zeros are in global memory
    int zero_exit;      // ... CTAs (search + filler) that have left the kernel
    int pad[10];
};

struct MasParams {
    const void* values;         // fp32, or fp16 / bf16 (promoted exactly on load; kernel instances are compiled per element type)
    void* paths;
    const int32_t* t_xs;
    const int32_t* t_ys;
    const void* mask;
    int64_t msb, msx, msy;
    int32_t* frame_tok;
    int32_t* durations;
    int32_t* lens_out;
    WsHeader* ws;
    uint32_t* bits_ws;          // global direction-bit slots (nullptr when bits are in smem)
    long long* dbg;             // optional [grid][2*kMaxWarps+2][2] clock64 stamps (ALB200_DBG)
    // Pipelined with the score kernel (alb200_gaussian_mas_fused): `values` is being written by a kernel running beside this one;
    // tile_ready[item * ready_tiles + k] == ready_epoch once frames [128 k, 128 k + 128) of every token row of `item` are in global memory.
    const int* tile_ready;
    int ready_epoch, ready_tiles;
    int pdl_wait;               // launched programmatically dependent on the kernel that writes `values` (fused entry, back-to-back form):
                                // everything up to the first score load overlaps that kernel's tail, the loaders then wait for it
    uint64_t one;
    int64_t bits_slot_words;
    int B, Tx, Ty;
    int esize;
    int mask_dtype;
    int zero_fill;
    int nw;                     // compute warps (blockDim = 2 * nw * 32)
    int ns;                     // ring stages per warp
    int nblk;                   // ceil(Ty/32)
    int aligned;                // values base and Ty allow 16-byte copies
    int nc;                     // CTAs per utterance (thread-block cluster; 1 = one CTA per utterance)
    int tail_rows;              // skewed form: token rows in the box of the LAST compute warp of the utterance (second tensor map)
    int search_ctas;            // shared zero fill (latency regime, fewer utterances than SMs): CTAs [0, search_ctas) align, the rest of the
    int zero_chunks;            //   grid only zero-fills; zero_chunks = kZeroChunk-sized pieces of the whole dense output (0 = every CTA fills its own utterance)
    float neg;
    uint32_t off_full, off_empty, off_xbar, off_flags, off_misc, off_bnd, off_zero, off_ring, off_bits, off_dur, off_bt, stage_bytes;   // make_layout(), done on the host
};

struct SmemLayout {
    uint32_t off_full, off_empty, off_xbar, off_flags, off_misc, off_bnd, off_zero, off_ring, off_bits, off_dur, off_bt, total;
    uint32_t stage_bytes;
};

__host__ __device__ inline uint32_t alb_align(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

// dense: the stages are written by 2-D TMA box loads (skewed form): rows back to back, no per-lane skew
__host__ __device__ inline SmemLayout make_layout(int NW, int NS, int R, int TF, int bits_smem, int nblk, int want_dur, int dense = 0, int nc = 1,
                                                  int elem_bytes = 4, int lag = 1)
{
    SmemLayout L;
    const uint32_t RW = 32u * R;
    L.stage_bytes = RW * TF * elem_bytes + (dense ? 0 : 32 * kLanePad);
    uint32_t o = 0;
    L.off_full = o;  o += NW * NS * 8;
    L.off_empty = o; o += NW * NS * 8;
    L.off_xbar = o;  o += (nc > 1 ? 4 * 8 : 0);
    L.off_flags = alb_align(o, 16); o = L.off_flags + (2 * NW + 2) * 4;    // tail (lane 31) and head (lane 0) progress per warp, the cross-CTA slot, a scratch word
    L.off_misc = alb_align(o, 16);  o = L.off_misc + 64 + 2 * kMaxWarps * 16; // item/lengths + per-warp partial mask sums
    L.off_bnd = alb_align(o, 16);   o = L.off_bnd + (NW + 1) * (lag == kLag4 ? kRing4 : kRing) * 4;   // ring 0: constant sentinel (the row above token 0), ring w+1: last row of warp w
    L.off_zero = alb_align(o, 128); o = L.off_zero + kZeroChunk;
    L.off_ring = alb_align(o, lag == kLag4 ? 1024 : 128); o = L.off_ring + NW * NS * L.stage_bytes;      // 128-byte-swizzled boxes: 1024-byte atoms
    L.off_bits = alb_align(o, 16);  o = L.off_bits + (bits_smem ? (uint32_t)nblk * NW * RW * 4 : 0);
    L.off_dur = alb_align(o, 16);   o = L.off_dur + (want_dur ? nc * NW * RW * 4 : 0);
    L.off_bt = alb_align(o, 16);    o = L.off_bt + (uint32_t)nblk * 8 + 16;                 // backtrack hand-off: (token, step mask) per 32-frame block + cursor
    L.total = alb_align(o, 16);
    return L;
}

#ifdef __CUDACC__
// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Every wait in this file is bounded: a wrong tensor map or a broken hand-off must surface as a launch error
// (cudaErrorLaunchFailure at the next synchronisation), never as a hung GPU.  2^26 polls of >= 20 cycles is more than half a
// second; the longest legitimate wait is a fraction of one utterance's forward pass (milliseconds).
#ifndef ALB_SPIN_LIMIT_LOG2
#define ALB_SPIN_LIMIT_LOG2 26
#endif
constexpr uint32_t kSpinLimit = 1u << ALB_SPIN_LIMIT_LOG2;
#ifdef ALB_TRAP_PRINT          // developer aid: say which wait gave up (build_lib.py --variant x -DALB_TRAP_PRINT=1 -DALB_SPIN_LIMIT_LOG2=20)
#include <cstdio>
#define spin_guard(n) if (++(n) > kSpinLimit) { if ((threadIdx.x & 31) == 0) printf("[alb200] wait at mas_kernel.cuh:%d gave up: block %d thread %d\n", __LINE__, (int)blockIdx.x, (int)threadIdx.x); break; }
#else
__device__ __forceinline__ void spin_guard(uint32_t& n) { if (++n > kSpinLimit) __trap(); }
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t n = 0;
    while (!mbar_try_wait(bar, parity)) spin_guard(n);
}
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {     // one non-blocking probe
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// 16-byte asynchronous global -> shared copy (LDGSTS.128), L2 only
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// the mbarrier receives one arrival when all of this thread's earlier cp.async have landed
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 2-D TMA box load (SASS UTMALDG): rows [c1, c1+box_rows) x frames [c0, c0+box_frames) land densely at dst, bytes complete on bar
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
// shared -> global bulk store (TMA engine, UBLKCP)
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
// Score element types: VT = 0 fp32, 1 fp16, 2 bf16.  Half types are promoted to fp32 on load -- exactly what the reference's
// `.astype(np.float32)` does (__init__.py:14), so paths are bit-identical to the reference run on the promoted values.
template <int VT> struct ValT { static constexpr int bytes = (VT == 0) ? 4 : 2; };
__device__ __forceinline__ float half_bits_to_float(uint32_t h) { return __half2float(__ushort_as_half((unsigned short)h)); }
template <int VT> __device__ __forceinline__ float ld_val(uint32_t a) {                 // one score
    if (VT == 0) return lds32(a);
    uint32_t h;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(h) : "r"(a));
    return VT == 1 ? half_bits_to_float(h) : __uint_as_float(h << 16);
}
template <int VT> __device__ __forceinline__ float4 ld_val4(uint32_t a) {               // four consecutive frames of a row
    if (VT == 0) return lds128(a);
    uint32_t lo, hi;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(a));
    if (VT == 1) return make_float4(half_bits_to_float(lo & 0xffffu), half_bits_to_float(lo >> 16), half_bits_to_float(hi & 0xffffu), half_bits_to_float(hi >> 16));
    return make_float4(__uint_as_float(lo << 16), __uint_as_float(lo & 0xffff0000u), __uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u));
}
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w));
}
// ---- thread-block cluster helpers (an utterance too long for one CTA's fast form is split over the CTAs of a cluster)
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {       // same shared-memory offset in CTA `rank` of the cluster
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// shared -> remote shared bulk copy (async proxy); the bytes complete on an mbarrier of the destination CTA.  This is the
// fence-free way to hand data to another CTA: a release store at cluster scope compiles to MEMBAR.ALL.GPU (~2000 cycles).
__device__ __forceinline__ void bulk_s2c(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t mbar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(mbar_cluster) : "memory");
}
// 16-byte store from registers into another CTA's shared memory whose bytes complete on an mbarrier over there (SASS STAS.128):
// data and "it has arrived" travel together, no fence on either side.
__device__ __forceinline__ void st_async_v4(uint32_t dst_cluster, float4 v, uint32_t mbar_cluster) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];"
                 ::"r"(dst_cluster), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)), "r"(mbar_cluster)
                 : "memory");
}
__device__ __forceinline__ void st_cluster_flag(uint32_t a, int v) {
    asm volatile("st.relaxed.cluster.shared::cluster.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Progress flags between neighbouring warps of one CTA.  Producer: the SAME lane stores the boundary
// values and then the flag; consumer: reads the flag, then the values.  Shared-memory accesses of one
// thread are performed in program order by the SM's in-order LSU pipe, so plain volatile accesses are
// sufficient; ld.acquire/st.release compile to MEMBAR.ALL.CTA (profiles/r01_notes.md).
__device__ __forceinline__ int ld_flag(uint32_t a) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void st_flag(uint32_t a, int v) {
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// bounded polls (see kSpinLimit): returns the first flag value that is >= need / <= need
__device__ __forceinline__ int wait_flag_ge(uint32_t a, int need, int seen) {
    uint32_t n = 0;
    while (seen < need) { seen = ld_flag(a); spin_guard(n); }
    return seen;
}
__device__ __forceinline__ void wait_flag_le(uint32_t a, int need) {
    uint32_t n = 0;
    while (ld_flag(a) > need) spin_guard(n);
}

// Waits until producer tiles <= need of this utterance are published (see MasParams::tile_ready).  Every lane polls the same word
// (one broadcast transaction); the acquire orders this lane's later loads, the proxy fence the TMA loads it issues, behind the
// producer's stores.  Bounded like every other wait in this file.
__device__ __forceinline__ void wait_tiles_ready(const int* flags, int epoch, int& known, int need) {
    uint32_t n = 0;
    while (known < need) {
        int v;
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + known + 1) : "memory");
        if (v == epoch) ++known; else { __nanosleep(64); spin_guard(n); }
    }
    asm volatile("fence.proxy.async.global;" ::: "memory");
}

// ------------------------------------------------------------------ mask -> length
// reference: t_x = mask.sum(1)[:,0], t_y = mask.sum(2)[:,0], astype(int32)  (__init__.py:18-19)
// Every thread of the CTA takes elements tid, tid+n, ... of the concatenation [mask[b,:,0] ; mask[b,0,:]] with eight
// independent loads in flight (walking them one dependent load at a time cost ~15% of the kernel).
template <typename T> __device__ __forceinline__ double mask_to_double(T v) { return (double)v; }
template <> __device__ __forceinline__ double mask_to_double<__half>(__half v) { return (double)__half2float(v); }
struct bf16_raw { unsigned short u; };
template <> __device__ __forceinline__ double mask_to_double<bf16_raw>(bf16_raw v) { return (double)__uint_as_float(((uint32_t)v.u) << 16); }

template <typename T>
__device__ __forceinline__ void mask_partial(const void* mv, int64_t base, int64_t sx, int64_t sy, int Tx, int Ty, int tid, int nthr,
                                             double& ax, double& ay)
{
    const T* m = reinterpret_cast<const T*>(mv) + base;
    const int n = Tx + Ty;
    for (int i0 = tid; i0 < n; i0 += 8 * nthr) {
        T v[8];
        int idx[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            idx[q] = i0 + q * nthr;
            if (idx[q] < n) v[q] = m[idx[q] < Tx ? (int64_t)idx[q] * sx : (int64_t)(idx[q] - Tx) * sy];
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (idx[q] < n) { if (idx[q] < Tx) ax += mask_to_double(v[q]); else ay += mask_to_double(v[q]); }
        }
    }
}
__device__ __forceinline__ void mask_partial_any(const void* m, int dtype, int64_t base, int64_t sx, int64_t sy, int Tx, int Ty,
                                                 int tid, int nthr, double& ax, double& ay)
{
    switch (dtype) {
        case 0: mask_partial<float>(m, base, sx, sy, Tx, Ty, tid, nthr, ax, ay); break;
        case 1: mask_partial<__half>(m, base, sx, sy, Tx, Ty, tid, nthr, ax, ay); break;
        case 2: mask_partial<bf16_raw>(m, base, sx, sy, Tx, Ty, tid, nthr, ax, ay); break;
        case 3: mask_partial<double>(m, base, sx, sy, Tx, Ty, tid, nthr, ax, ay); break;
        case 4: mask_partial<uint8_t>(m, base, sx, sy, Tx, Ty, tid, nthr, ax, ay); break;
        case 5: mask_partial<int8_t>(m, base, sx, sy, Tx, Ty, tid, nthr, ax, ay); break;
        case 6: mask_partial<int16_t>(m, base, sx, sy, Tx, Ty, tid, nthr, ax, ay); break;
        case 7: mask_partial<int32_t>(m, base, sx, sy, Tx, Ty, tid, nthr, ax, ay); break;
        default: mask_partial<long long>(m, base, sx, sy, Tx, Ty, tid, nthr, ax, ay); break;
    }
}

__device__ __forceinline__ void store_one(void* paths, int64_t elem, int esize, uint64_t one) {
    switch (esize) {
        case 1: ((uint8_t*)paths)[elem] = (uint8_t)one; break;
        case 2: ((uint16_t*)paths)[elem] = (uint16_t)one; break;
        case 4: ((uint32_t*)paths)[elem] = (uint32_t)one; break;
        default: ((unsigned long long*)paths)[elem] = (unsigned long long)one; break;
    }
}

// ------------------------------------------------------------------ forward state of one compute lane
template <int R>
struct Fwd {
    float old[R];        // running column: value of each of our rows at the previous frame
    uint32_t wbits[R];   // direction word per row, shifted in from the top 4 frames at a time
    uint32_t wprev[R];   // skewed: the 32 frames before those in wbits (a lane's words straddle two units)
    float up;            // lock-step: neighbour's last row at the previous frame
    float u1, u2;        // skewed: neighbour's last row for the next frame and the one after (shuffles two frames in flight)
    float u3, u4, u5;    // 4-frame lag: five shuffles in flight
    float bprev;         // lane 0: value of the row above our first row at the frame before this group
    float bprev1;        // skewed, lane 0: ... and at the first frame of this group (the boundary ring is step-indexed)
};

// UNIT consecutive frames for this lane's R rows.
//   Y   frame of lane 0 at the start of the unit;  yl = this lane's frame (Y, or Y - lane when skewed)
//   Reference semantics per cell: core.pyx:19-30 (see inline notes).
// Scheduling notes (one warp per scheduler: every exposed latency is paid in full):
//   * the tile values of group g+1 and all boundary values of the unit are fetched before they are needed;
//   * the body is branch-free: a completed direction word is snapshotted with predicated moves and stored once,
//     after the unit, so the four groups stay one basic block.
template <int R, int TF, int UNIT, bool SKEW, bool DIAG, int VT, bool VL>
__device__ __forceinline__ void forward_unit(Fwd<R>& S, uint32_t tile_addr, uint32_t tile_prev, uint32_t bin_addr, uint32_t bout_addr, int Y, int yl,
                                             bool lane0, bool lane31, float neg, int dxy, uint32_t* bits_row, int TXS,
                                             int y_lo, unsigned span, uint32_t* bits_unit, bool row0)
{
    constexpr int NG = UNIT / 4;
    static_assert(!SKEW || UNIT == 32, "the skewed form assembles one direction word per 32-frame unit");
    // Skewed form: lane l is l frames behind lane 0, so lane 31 of the warp above us finishes frame f at ITS step f + 31.  The
    // boundary ring is indexed by the producer's step (aligned 128-bit stores); frame f sits in slot f + 31.
    float4 bin[NG];
    // Y is a multiple of UNIT and the ring a multiple of 32 slots: the unit's slots never wrap, one base + immediates
    const uint32_t bin_base = bin_addr + (((Y + (SKEW ? 32 : 0)) & (kRing - 1)) << 2);
    const uint32_t bout_base = bout_addr + ((Y & (kRing - 1)) << 2);
#pragma unroll
    for (int g = 0; g < NG; ++g) bin[g] = (ALB_ABL & 8) ? make_float4(neg, neg, neg, neg) : lds128(bin_base + 16 * g);
    float4 vn[R];
    // Skewed form: the tiles in shared memory are NOT skewed (same 128-bit asynchronous copies as the lock-step form); lane l
    // reads frame Y + k - l, which is tile position k - l of this tile, or 32 + k - l of the previous one while k < l.  One
    // compare + select per frame picks the base; the loads are scalar and run two frames ahead.
    const int lane = Y - yl;
    constexpr int ES = ValT<VT>::bytes;
    // VL (VITS layout, scores stored [t_mel][t_text]): a tile is [TF frames][32*R tokens], so consecutive frames of a token
    // are one tile row apart and a lane's R tokens are adjacent; otherwise [32*R tokens][TF frames].
    constexpr int FSTR = VL ? 32 * R * ES : ES;              // bytes between consecutive frames of one token
    constexpr int RSTR = VL ? ES : TF * ES;                  // bytes between a lane's consecutive tokens
    static_assert(!VL || SKEW, "the VITS layout exists for the skewed/TMA form only");
    const uint32_t curA = tile_addr - (uint32_t)(FSTR * lane), prevA = tile_prev + (uint32_t)(FSTR * (TF - lane));
    float vq[2][R];
    if (SKEW) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const uint32_t a = (lane <= q) ? curA : prevA;
#pragma unroll
            for (int r = 0; r < R; ++r) vq[q][r] = ld_val<VT>(a + q * FSTR + r * RSTR);
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) vn[r] = ld_val4<VT>(tile_addr + r * (TF * ES));
    }
    uint32_t wdone[R];                                                // the word this lane completes inside this unit, if any
#pragma unroll
    for (int r = 0; r < R; ++r) wdone[r] = 0u;
    const int gdone = ((28 - yl) & 31) >> 2;                          // group after which our 32-frame word is complete (>= NG: not in this unit)
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        float4 v[R];
        if (!SKEW) {
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = vn[r];
            if (g + 1 < NG) {
#pragma unroll
                for (int r = 0; r < R; ++r) vn[r] = ld_val4<VT>(tile_addr + r * (TF * ES) + (g + 1) * 4 * ES);
            }
        }
        // row above us at frames Y+4g-1 .. Y+4g+2 (skewed: slots Y+4g+30 .. Y+4g+33, i.e. .zw of the previous load, .xy of this one)
        const float b0 = S.bprev, b1 = SKEW ? S.bprev1 : bin[g].x, b2 = SKEW ? bin[g].x : bin[g].y, b3 = SKEW ? bin[g].y : bin[g].z;
        S.bprev = SKEW ? bin[g].z : bin[g].w;
        if (SKEW) S.bprev1 = bin[g].w;
        uint32_t hb[R];
#pragma unroll
        for (int r = 0; r < R; ++r) hb[r] = 0u;
        float o4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int kk = 4 * g + k;
            const float bk = (k == 0) ? b0 : (k == 1) ? b1 : (k == 2) ? b2 : b3;
            const float upv = lane0 ? bk : (SKEW ? S.u1 : S.up);
            float nv[R];
#pragma unroll
            for (int r = R - 1; r >= 0; --r) {
                const float stay = S.old[r];                        // v_cur  (core.pyx:22; == neg on the diagonal, rows above it are held)
                const float move = (r == 0) ? upv : S.old[r - 1];   // v_prev (core.pyx:29)
                const bool take = move > stay;                      // core.c:19384
                const float vr = SKEW ? vq[kk & 1][r] : ((k == 0) ? v[r].x : (k == 1) ? v[r].y : (k == 2) ? v[r].z : v[r].w);
                float res;
                if (ALB_ABL & 16) {                                 // timing ablation: max + add (not the reference's NaN behaviour)
                    res = fmaxf(stay, move) + vr;
                } else if (ALB_ABL & 32) {                          // timing ablation: both sums formed beside the compare, select last
                    float rs, rm;
                    asm("add.f32 %0, %2, %4;\n\tadd.f32 %1, %3, %4;" : "=f"(rs), "=f"(rm) : "f"(stay), "f"(move), "f"(vr));
                    asm("" : "+f"(rs), "+f"(rm));
                    res = take ? rm : rs;
                } else if (ALB_SPEC_ADD && SKEW) {
                    // latency regime (one warp per scheduler, chain-bound): both candidate sums are formed while the compare runs and
                    // the select comes last -- the dependent chain per frame is {FADD | FSETP} -> FSEL instead of FSETP -> FSEL -> FADD.
                    // Bit-identical: the selected operand meets the same single fp32 add (core.pyx:30).
                    const float rs = stay + vr, rm = move + vr;
                    res = take ? rm : rs;
                } else {
                    res = (take ? move : stay) + vr;                // core.pyx:30
                }
                if (DIAG) res = (dxy + r > kk) ? neg : res;         // rows above the diagonal stay at the sentinel
                nv[r] = res;
                if (take && !(ALB_ABL & 4)) hb[r] |= (1u << k);
            }
            if (SKEW) {
                // lane l-1 is one frame ahead: what it finishes now is what we need the frame after next
                S.u1 = S.u2;
                S.u2 = (ALB_ABL & 1) ? nv[R - 1] * 0.5f : __shfl_up_sync(0xffffffffu, nv[R - 1], 1);
            } else {
                S.up = __shfl_up_sync(0xffffffffu, nv[R - 1], 1);
            }
            o4[k] = nv[R - 1];
#pragma unroll
            for (int r = 0; r < R; ++r) S.old[r] = nv[r];
            if (SKEW && kk + 2 < UNIT && !(ALB_ABL & 2)) {
                const uint32_t a = (lane <= kk + 2) ? curA : prevA;
#pragma unroll
                for (int r = 0; r < R; ++r) vq[kk & 1][r] = ld_val<VT>(a + (kk + 2) * FSTR + r * RSTR);
            }
        }
        if (lane31 && !(ALB_ABL & 8)) sts128(bout_base + 16 * g, o4[0], o4[1], o4[2], o4[3]);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            S.wbits[r] = __funnelshift_r(S.wbits[r], hb[r], 4);
            if (!SKEW) wdone[r] = (gdone == g) ? S.wbits[r] : wdone[r];
        }
    }
    if (SKEW) {
        // wbits now holds our frames [Y - lane, Y + 32 - lane), wprev the 32 before: the aligned word of block [Y - 32, Y)
        // is the 64-bit window shifted right by `lane`
        // The skewed forms store the word in the form the backtrack walks it (the lock-step form leaves this to the walker's
        // window copy): the forced step of the diagonal cell (core.pyx:34, index == y) set -- it can only fall into a DIAG unit --,
        // token 0 never steps down (index != 0), bit-reversed (frame k at bit 31-k).
        const int yw = Y - 32;
        if ((unsigned)(yw - y_lo) < span) {
            uint32_t* brow = bits_unit;                                   // == bits_row + ((Y - 32) >> 5) * TXS, advanced by the caller
#pragma unroll
            for (int r = 0; r < R; ++r) {
                uint32_t wv = __funnelshift_r(S.wprev[r], S.wbits[r], lane);
                if (DIAG) { const int d = dxy + r + 32 - lane; if ((unsigned)d < 32u) wv |= 1u << d; }     // d = token - first frame of the block
                if (r == 0 && row0) wv = 0u;
                brow[r] = __brev(wv);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) S.wprev[r] = S.wbits[r];
    } else {
        const int yw = yl + 4 * gdone;                               // first frame of the last group of the completed word
        if (gdone < NG && (unsigned)(yw - y_lo) < span) {
            uint32_t* brow = bits_row + (int64_t)(yw >> 5) * TXS;
#pragma unroll
            for (int r = 0; r < R; ++r) brow[r] = wdone[r];
        }
    }
}

// ------------------------------------------------------------------ forward unit, 4-frame lag on pre-skewed tiles
// Lane l runs kLag4 = 4 frames behind lane l-1.  Four fp32 frames are 16 bytes -- the granularity by which a tensor map can shift
// one box row against the next (row stride R*t_y*4 - 16 bytes; box start coordinates themselves must be 16-byte aligned, so a
// 1-frame skew cannot be had this way: tools/micro/tma_skew_test.cu).  ONE 3-D box load per unit then delivers every lane's own 32
// frames, 128-byte swizzled: lane l's row r is box row r*32 + l, its frames 4g..4g+3 the 16-byte chunk g ^ (l & 7).  That is one
// conflict-free LDS.128 per row and four frames, no current/previous-tile select, nothing kept from the previous tile; and the
// neighbour's value is needed five steps after it was produced, so no shuffle latency is ever exposed.
// Measured (tools/micro/body2_bench.cu, a lone warp, cycles per frame, R = 2 / 3 / 4): dense tiles with a 1-frame lag 36.6 / 50.2 /
// 59.6; this form 21.3 / 25.9 / 28.6.  The price is 4*31 frames of fill per warp instead of 31.
template <int R, bool DIAG>
__device__ __forceinline__ void forward_unit4(Fwd<R>& S, uint32_t tile_lane, uint32_t xs, uint32_t bin_addr, uint32_t bout_addr, int Y, int yl,
                                              bool lane0, bool lane31, float neg, int dxy, uint32_t* bits_unit, int y_lo, unsigned span, bool row0)
{
    constexpr int NG = 8;
    // Boundary ring, indexed by the producer's step: its lane 31 finishes frame f at step f + 124.  Our step Y + kk needs frame
    // Y + kk - 1 of the row above = slot Y + kk + 123: kk = 0 is the last slot of the previous unit's loads (S.bprev), then
    // slots Y + 124 .. Y + 154.  The unit's slots may straddle the ring's end: one mask per 16-byte group.
    float4 bin[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) bin[g] = lds128(bin_addr + (uint32_t)(((Y + 31 * kLag4 + 4 * g) & (kRing4 - 1)) << 2));
    const uint32_t bout_base = bout_addr + (uint32_t)((Y & (kRing4 - 1)) << 2);     // Y is a multiple of 32: our 32 slots never wrap
    float4 vg[2][R];
#pragma unroll
    for (int r = 0; r < R; ++r) vg[0][r] = lds128(tile_lane + r * 4096 + xs);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        if (g + 1 < NG) {
            const uint32_t a = tile_lane + (((uint32_t)(g + 1) << 4) ^ xs);
#pragma unroll
            for (int r = 0; r < R; ++r) vg[(g + 1) & 1][r] = lds128(a + r * 4096);
        }
        const float b0 = S.bprev, b1 = bin[g].x, b2 = bin[g].y, b3 = bin[g].z;
        S.bprev = bin[g].w;
        uint32_t hb[R];
#pragma unroll
        for (int r = 0; r < R; ++r) hb[r] = 0u;
        float o4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int kk = 4 * g + k;
            const float bk = (k == 0) ? b0 : (k == 1) ? b1 : (k == 2) ? b2 : b3;
            const float upv = lane0 ? bk : S.u1;
            float nv[R];
#pragma unroll
            for (int r = R - 1; r >= 0; --r) {
                const float stay = S.old[r];                        // v_cur  (core.pyx:22)
                const float move = (r == 0) ? upv : S.old[r - 1];   // v_prev (core.pyx:29)
                const bool take = move > stay;                      // core.c:19384
                const float4 v4 = vg[g & 1][r];
                const float vr = (k == 0) ? v4.x : (k == 1) ? v4.y : (k == 2) ? v4.z : v4.w;
                float res = (take ? move : stay) + vr;              // core.pyx:30
                if (DIAG) res = (dxy + r > kk) ? neg : res;         // rows above the diagonal stay at the sentinel
                nv[r] = res;
                if (take) hb[r] |= (1u << k);
            }
            // lane l-1 is four frames ahead: what it finishes now is what we need five steps from now
            S.u1 = S.u2; S.u2 = S.u3; S.u3 = S.u4; S.u4 = S.u5;
            S.u5 = __shfl_up_sync(0xffffffffu, nv[R - 1], 1);
            o4[k] = nv[R - 1];
#pragma unroll
            for (int r = 0; r < R; ++r) S.old[r] = nv[r];
        }
        if (lane31) sts128(bout_base + 16 * g, o4[0], o4[1], o4[2], o4[3]);
#pragma unroll
        for (int r = 0; r < R; ++r) S.wbits[r] = __funnelshift_r(S.wbits[r], hb[r], 4);
    }
    // wbits holds our frames [Y - lag, Y - lag + 32), wprev the 32 before: the aligned word this lane completes now is the block
    // that starts at Y - 32 - (lag & ~31); it is the 64-bit window shifted right by lag & 31
    const int lagf = Y - yl;
    const int yw = Y - 32 - (lagf & ~31);
    if ((unsigned)(yw - y_lo) < span) {
        uint32_t* brow = bits_unit;                                       // == bits_row + (yw >> 5) * TXS: per-lane base, advanced by the caller
#pragma unroll
        for (int r = 0; r < R; ++r) {                                     // final form, as in forward_unit
            uint32_t wv = __funnelshift_r(S.wprev[r], S.wbits[r], lagf & 31);
            if (DIAG) { const int d = dxy + r + 32 - (lagf & 31); if ((unsigned)d < 32u) wv |= 1u << d; }
            if (r == 0 && row0) wv = 0u;
            brow[r] = __brev(wv);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) S.wprev[r] = S.wbits[r];
}

// ------------------------------------------------------------------ backtrack walker (one warp)
// Walks 32 frames (one direction word per row) per step and publishes (token at the block's last frame, mask of the
// frames at which the path steps down) per block; the other warps turn those into stores.  What the reference does
// per frame (core.pyx:32-35: index -= 1 iff index != 0 and (index == y or value[index, y-1] < value[index-1, y-1]))
// becomes, per ROW visited: "the next step down is at the highest frame not above the current one whose bit is set".
//   * Words are stored bit-REVERSED (frame k at bit 31-k) so that "highest frame" is "lowest set bit":
//         t     = W & below                 remaining step candidates on this row
//         below'= ~(t ^ (t - 1))            frames strictly before the step (0 when t == 0: the block is finished)
//     so the next row's t is ONE add and ONE three-input logic op after this row's: t' = W' & ~(t ^ (t - 1)).
//   * Software pipeline over blocks, so that only [window loads -> walk] is on the serial path:
//       iteration blk, start token known:   (1) issue the direction-word loads of block blk-LD, rows anchor-0 .. anchor-32(LD+1)+1
//                                               with anchor = the current token (the path drops at most 32 rows per block);
//                                           (2) patch the forced diagonal steps into the words loaded LD-1 iterations ago and
//                                               store them as the shared-memory window of block blk-1;
//                                           (3) walk block blk through the window stored one iteration ago (broadcast loads).
//     LD = 2 when the bits live in shared memory, 4 when they come from L2.
template <int LD, bool FINAL>
__device__ __forceinline__ void backtrack_walk(const uint32_t* bits, int TXS, int t_x, int t_y, int top, int lane, uint32_t win_a,
                                               volatile int* btTok, volatile uint32_t* btMov, uint32_t bt_cur_a, long long* dbg_e)
{
    constexpr int NWD = LD + 1;                              // words per lane in a window
    constexpr uint32_t WIN_BYTES = (NWD * 32 + 16) * 4;        // plus the walk's read-ahead
    if (top < 0) return;
    uint32_t fifo[LD][NWD];                                  // [0] = next block to store
    int anchor[LD];
    auto load_win = [&](int blk, int a, uint32_t (&wv)[NWD]) {
#pragma unroll
        for (int j = 0; j < NWD; ++j) {
            const int r = a - 32 * j - lane;
            wv[j] = (blk >= 0 && r > 0) ? bits[(int64_t)blk * TXS + r] : 0u;   // row 0 can never step down (core.pyx:34 index != 0)
        }
    };
    auto store_win = [&](int blk, int a, const uint32_t (&wv)[NWD]) {
        const int yb = blk << 5;
        const uint32_t wbuf = win_a + (uint32_t)(blk & 1) * WIN_BYTES;
#pragma unroll
        for (int j = 0; j < NWD; ++j) {
            const int r = a - 32 * j - lane, d = r - yb;
            uint32_t v = wv[j];
            if (!FINAL) {                                                        // (the skewed forms store their words patched and reversed)
                if (r > 0 && d >= 0 && d < 32) v |= (1u << d);                   // diagonal cell: forced step (index == y)
                v = __brev(v);
            }
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(wbuf + 4u * (uint32_t)(32 * j + lane)), "r"(v) : "memory");
        }
    };
    int tok0 = t_x - 1;
    // prologue: blocks top .. top-LD+1 are all anchored at the final token
    int anc_cur = tok0;
#pragma unroll
    for (int j = 0; j < LD; ++j) { load_win(top - j, tok0, fifo[j]); anchor[j] = tok0; }
    store_win(top, tok0, fifo[0]);
#pragma unroll
    for (int j = 0; j + 1 < LD; ++j) {
#pragma unroll
        for (int q = 0; q < NWD; ++q) fifo[j][q] = fifo[j + 1][q];
        anchor[j] = anchor[j + 1];
    }
    __syncwarp();
    long long bB = 0, bC = 0, q0 = 0, q2 = 0;
    for (int blk = top; blk >= 0; --blk) {
        if (dbg_e) q0 = clock64();
        const int yb = blk << 5;
        const int nvalid = (t_y - yb < 32) ? t_y - yb : 32;
        // (3a) first window words of this block: the only loads on the serial path
        uint32_t wadr = win_a + (uint32_t)(blk & 1) * WIN_BYTES + 4u * (uint32_t)(anc_cur - tok0);
        uint32_t w0, w1, w2, w3;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w0) : "r"(wadr) : "memory");      // "memory": nothing is hoisted above these
        asm volatile("ld.shared.b32 %0, [%1+4];" : "=r"(w1) : "r"(wadr) : "memory");
        asm volatile("ld.shared.b32 %0, [%1+8];" : "=r"(w2) : "r"(wadr) : "memory");
        asm volatile("ld.shared.b32 %0, [%1+12];" : "=r"(w3) : "r"(wadr) : "memory");
        // (1) loads for block blk-LD, anchored here;  (2) window of block blk-1 from the words loaded LD-1 iterations ago
        load_win(blk - LD, tok0, fifo[LD - 1]);
        anchor[LD - 1] = tok0;
        const int anc_next = anchor[0];
        // (3b) the walk.  Rows are left strictly in the order tok0, tok0-1, ...; every lane replays the same scalar walk.
        uint32_t moves = 0u;                                                           // bit 31-k: the path steps down going from frame k to k-1
        const uint32_t below0 = (nvalid < 32) ? ~((1u << (32 - nvalid)) - 1u) : 0xffffffffu;   // (reversed) frames of this block
        uint32_t t = w0 & below0;
#define ALB_BT_STEP(T_IN, W_NEXT, T_OUT)                                                                                \
    {                                                                                                                   \
        const uint32_t tm = (T_IN) - 1u;                                                                                \
        moves |= (T_IN) & ~tm;                        /* the step: highest remaining frame = lowest reversed bit */     \
        T_OUT = (W_NEXT) & ~((T_IN) ^ tm);            /* earlier frames go to the rows further down; 0 ends the block */ \
    }
        // branch-free steps (a GPU does not speculate: a per-step exit test would put the branch latency on the chain); once a
        // row has no step left, t is 0 and the remaining steps of the group are no-ops.  Reads may run a few words past the
        // window (the buffer is padded); those values meet t == 0 and are never used.  The first eight steps are straight-line
        // code in the same basic block as (1) and (2), so the scheduler hides those in the shadow of the chain.
#define ALB_BT_QUAD                                                                                                     \
        {                                                                                                               \
            uint32_t t1, t2, t3;                                                                                        \
            asm volatile("ld.shared.b32 %0, [%1+16];" : "=r"(w0) : "r"(wadr));                                          \
            ALB_BT_STEP(t, w1, t1)                                                                                      \
            asm volatile("ld.shared.b32 %0, [%1+20];" : "=r"(w1) : "r"(wadr));                                          \
            ALB_BT_STEP(t1, w2, t2)                                                                                     \
            asm volatile("ld.shared.b32 %0, [%1+24];" : "=r"(w2) : "r"(wadr));                                          \
            ALB_BT_STEP(t2, w3, t3)                                                                                     \
            asm volatile("ld.shared.b32 %0, [%1+28];" : "=r"(w3) : "r"(wadr));                                          \
            ALB_BT_STEP(t3, w0, t)                                                                                      \
            wadr += 16u;                                                                                                \
        }
        ALB_BT_QUAD
        ALB_BT_QUAD
        // (2) in program order AFTER the straight-line steps: the stores are fire-and-forget, and the arithmetic that feeds
        // them is free to move up into the idle issue slots of the chain above
        store_win(blk - 1, anchor[0], fifo[0]);               // unconditional (block -1 lands in a buffer nobody reads): no branch, one basic block
        __syncwarp();                                         // the window of block blk-1 is visible to every lane before its walk
#pragma unroll
        for (int j = 0; j + 1 < LD; ++j) {
#pragma unroll
            for (int q = 0; q < NWD; ++q) fifo[j][q] = fifo[j + 1][q];
            anchor[j] = anchor[j + 1];
        }
        while (t != 0u) ALB_BT_QUAD
#undef ALB_BT_QUAD
#undef ALB_BT_STEP
        const int nmove = __popc(moves);
        if (dbg_e) q2 = clock64();
        if (lane == 0) {
            btTok[blk] = tok0;
            btMov[blk] = moves;
            st_flag(bt_cur_a, blk);                           // same lane, after the payload: in-order shared-memory pipe
        }
        anc_cur = anc_next;
        tok0 -= nmove;
        if (dbg_e) { const long long q3 = clock64(); bB += q2 - q0; bC += q3 - q2; }
    }
    if (dbg_e && lane == 0) { dbg_e[0] = 0; dbg_e[1] = bB / (top + 1); dbg_e[2] = bC / (top + 1); dbg_e[3] = -(top + 1); }
}

// Direction words in shared memory and already in their final form (skewed forms): the walk reads them where the forward pass put
// them -- word (block, token) at bits[block * TXS + token], the rows it leaves strictly downwards at descending addresses.  No window
// copy, no software pipeline: the copy's ~100 instructions per block were what the walker's ~340 cycles per block went to (a lone
// warp issues one instruction per ~2.4 cycles), not the step chain.  Reads may run a few words below token 0 or into rows whose
// words were never written; those meet t == 0 (token 0's word is stored as 0) and are never used.
__device__ __forceinline__ void backtrack_walk_direct(uint32_t bits_a, int TXS, int t_x, int t_y, int top, int lane,
                                                      volatile int* btTok, volatile uint32_t* btMov, uint32_t bt_cur_a, long long* dbg_e)
{
    if (top < 0) return;
    int tok0 = t_x - 1;
    long long bB = 0, bC = 0, q0 = 0, q2 = 0;
    for (int blk = top; blk >= 0; --blk) {
        if (dbg_e) q0 = clock64();
        const int yb = blk << 5;
        const int nvalid = (t_y - yb < 32) ? t_y - yb : 32;
        uint32_t wadr = bits_a + 4u * (uint32_t)(blk * TXS + tok0);
        uint32_t w0, w1, w2, w3;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w0) : "r"(wadr) : "memory");
        asm volatile("ld.shared.b32 %0, [%1+-4];" : "=r"(w1) : "r"(wadr) : "memory");
        asm volatile("ld.shared.b32 %0, [%1+-8];" : "=r"(w2) : "r"(wadr) : "memory");
        asm volatile("ld.shared.b32 %0, [%1+-12];" : "=r"(w3) : "r"(wadr) : "memory");
        uint32_t moves = 0u;                                                           // bit 31-k: the path steps down going from frame k to k-1
        const uint32_t below0 = (nvalid < 32) ? ~((1u << (32 - nvalid)) - 1u) : 0xffffffffu;   // (reversed) frames of this block
        uint32_t t = w0 & below0;
#define ALB_BT_STEP(T_IN, W_NEXT, T_OUT)                                                                                \
    {                                                                                                                   \
        const uint32_t tm = (T_IN) - 1u;                                                                                \
        moves |= (T_IN) & ~tm;                        /* the step: highest remaining frame = lowest reversed bit */     \
        T_OUT = (W_NEXT) & ~((T_IN) ^ tm);            /* earlier frames go to the rows further down; 0 ends the block */ \
    }
#define ALB_BT_QUAD                                                                                                     \
        {                                                                                                               \
            uint32_t t1, t2, t3;                                                                                        \
            asm volatile("ld.shared.b32 %0, [%1+-16];" : "=r"(w0) : "r"(wadr));                                         \
            ALB_BT_STEP(t, w1, t1)                                                                                      \
            asm volatile("ld.shared.b32 %0, [%1+-20];" : "=r"(w1) : "r"(wadr));                                         \
            ALB_BT_STEP(t1, w2, t2)                                                                                     \
            asm volatile("ld.shared.b32 %0, [%1+-24];" : "=r"(w2) : "r"(wadr));                                         \
            ALB_BT_STEP(t2, w3, t3)                                                                                     \
            asm volatile("ld.shared.b32 %0, [%1+-28];" : "=r"(w3) : "r"(wadr));                                         \
            ALB_BT_STEP(t3, w0, t)                                                                                      \
            wadr -= 16u;                                                                                                \
        }
        ALB_BT_QUAD
        ALB_BT_QUAD
        while (t != 0u) ALB_BT_QUAD
#undef ALB_BT_QUAD
#undef ALB_BT_STEP
        const int nmove = __popc(moves);
        if (dbg_e) q2 = clock64();
        if (lane == 0) {
            btTok[blk] = tok0;
            btMov[blk] = moves;
            st_flag(bt_cur_a, blk);                           // same lane, after the payload: in-order shared-memory pipe
        }
        tok0 -= nmove;
        if (dbg_e) { const long long q3 = clock64(); bB += q2 - q0; bC += q3 - q2; }
    }
    if (dbg_e && lane == 0) { dbg_e[0] = 0; dbg_e[1] = bB / (top + 1); dbg_e[2] = bC / (top + 1); dbg_e[3] = -(top + 1); }
}

// ------------------------------------------------------------------ the kernel
// NWMAX bounds the compute warps of an instance (4 -> 256 threads, 8 -> 512 threads).  MINB = 2 holds an instance to 128
// registers so two CTAs share an SM (throughput regime); MINB = 1 lets an utterance that owns its SM use up to 255.
template <int R, int TF, bool SKEW, int NWMAX, int MINB, bool CL = false, int VT = 0, bool VL = false, int LAG = 1>
__global__ void __launch_bounds__(2 * NWMAX * 32, MINB) mas_kernel(const MasParams p, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_tail)
{
    constexpr int RW = 32 * R;
    // hand-off granularity between warps: a whole 32-frame tile when an utterance owns its SM (fewer flag/barrier round trips per
    // frame), 16 frames in the register-capped throughput instances
    constexpr int UNIT = (MINB == 1 && TF == 32) ? 32 : (TF < 16 ? TF : 16);
    constexpr int ES = ValT<VT>::bytes;                  // bytes per score
    constexpr int LANE_STRIDE = R * TF * ES + (SKEW ? 0 : kLanePad);   // skewed: dense TMA tiles, the scalar reads of lanes l and frames k - l hit 32 banks
    constexpr bool L4 = (LAG == kLag4);                  // 4-frame lag on pre-skewed, swizzled boxes (forward_unit4)
    static_assert(LAG == 1 || (L4 && SKEW && !CL && VT == 0 && !VL && TF == 32 && MINB == 1), "the 4-frame lag exists for the fp32 skewed/TMA form only");
    constexpr int LAG31 = SKEW ? 31 * LAG : 0;           // frames lane 31 trails lane 0
    constexpr int RING = L4 ? kRing4 : kRing;            // slots of a warp-boundary ring

    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int NW = p.nw;
    const int NS = p.ns;
    const bool is_loader = wid >= NW;
    const int w = is_loader ? wid - NW : wid;            // the compute warp this warp is, or serves
    // Cluster mode (p.nc > 1, latency regime, long utterances): the CTAs of a cluster split one utterance's token rows; CTA c
    // owns compute warps c*NW .. c*NW+NW-1 of one long pipeline.  The only cross-CTA traffic is the boundary ring between the
    // last warp of CTA c and warp 0 of CTA c+1 (remote shared-memory stores, polls stay local), bits go to the L2 slot, CTA 0
    // backtracks.
    const int NC = CL ? p.nc : 1;                        // CL instances only: the single-CTA instances carry none of this
    const int crank = NC > 1 ? (int)cluster_ctarank() : 0;
    const int unit_id = NC > 1 ? (int)blockIdx.x / NC : (int)blockIdx.x;      // cluster index, or CTA index
    const int gw = crank * NW + w;                       // position of this warp in the utterance's pipeline
    const int TXS = NC * NW * RW;
    const int nthr = blockDim.x;
    const bool bits_smem = (p.bits_ws == nullptr);
    struct { uint32_t off_full, off_empty, off_flags, off_misc, off_bnd, off_zero, off_ring, off_bits, off_dur, off_bt, stage_bytes; } L =
        { p.off_full, p.off_empty, p.off_flags, p.off_misc, p.off_bnd, p.off_zero, p.off_ring, p.off_bits, p.off_dur, p.off_bt, p.stage_bytes };

    const uint32_t smem0 = smem_u32(smem);
    const uint32_t full0 = smem0 + L.off_full + w * NS * 8;
    const uint32_t empty0 = smem0 + L.off_empty + w * NS * 8;
    const uint32_t tail_a = smem0 + L.off_flags;                    // frames finished by lane 31 of warp w at +4*w
    const uint32_t head_a = tail_a + 4 * NW;                        // frames finished by lane 0  of warp w at +4*w
    const uint32_t xout_head_a = head_a + 4 * NW;                   // cluster: progress of the next CTA's warp 0 (written remotely)
    const uint32_t xbar_a = smem0 + p.off_xbar;                     // cluster: 4 mbarriers, one per 32-slot quarter of the incoming boundary ring
    int* misc = reinterpret_cast<int*>(smem + L.off_misc);          // [0]=item [1]=t_x [2]=t_y
    double* msum = reinterpret_cast<double*>(smem + L.off_misc + 64);   // [warp][2] partial mask sums
    const uint32_t bnd_a = smem0 + L.off_bnd;
    const uint32_t zero_a = smem0 + L.off_zero;
    const uint32_t ring_a = smem0 + L.off_ring + w * NS * L.stage_bytes;
    uint32_t* bits = bits_smem ? reinterpret_cast<uint32_t*>(smem + L.off_bits)
                               : p.bits_ws + (int64_t)unit_id * p.bits_slot_words;
    int* durS = reinterpret_cast<int*>(smem + L.off_dur);
    volatile int* btTok = reinterpret_cast<volatile int*>(smem + L.off_bt);          // [nblk] token at the last frame of each block
    volatile uint32_t* btMov = reinterpret_cast<volatile uint32_t*>(smem + L.off_bt) + p.nblk;   // [nblk] frames at which the path steps down (bit 31-k)
    const uint32_t bt_cur_a = smem0 + L.off_bt + 8 * (uint32_t)p.nblk;              // lowest block the walker has published

    // ---- one-time setup
    if (CL && tid == 0)
        for (int q = 0; q < 4; ++q) mbar_init(xbar_a + 8 * q, 1);
    if (!is_loader && lane == 0)
        for (int s = 0; s < NS; ++s) { mbar_init(full0 + 8 * s, SKEW ? 1 : 32); mbar_init(empty0 + 8 * s, 32); }   // one arrival per lane (full, skewed: the TMA issuer)
    for (int i = tid; i < kZeroChunk / 16; i += nthr)
        reinterpret_cast<int4*>(smem + L.off_zero)[i] = make_int4(0, 0, 0, 0);
    if (crank == 0)
        for (int i = tid; i < RING; i += nthr) reinterpret_cast<float*>(smem + L.off_bnd)[i] = p.neg;
    fence_mbar_init();
    fence_proxy_async_smem();
    __syncthreads();

    if (kDbgBuild && p.tile_ready != nullptr && tid == 0) {   // developer aid: when did the pipelined search start / reach its first wait
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        atomicMin(reinterpret_cast<unsigned long long*>(const_cast<int*>(p.tile_ready) + (int64_t)p.B * p.ready_tiles + 64) + 2, gt);
    }
    uint32_t stage = 0, phase = 0;          // ring cursor of this warp (consumer side for compute, producer side for loaders)
    const int64_t item_elems = (int64_t)p.Tx * p.Ty;
    const int Ty = p.Ty;
    const float neg = p.neg;

    int item = unit_id;
    // One box load of tile t (frames [t*TF, t*TF+TF)) of the rows of pipeline warp `g` of utterance `it` into stage address st (skewed forms).
    // The last compute warp of the (padded) text axis is usually only partly filled: it fetches a shorter box through a second
    // tensor map -- every box row costs TMA engine time (profiles/r01_notes.md).
    auto issue_tile = [&](int it, int g, int t, uint32_t st, uint32_t bar) {
        const int xg = g * RW;
        if (L4) {
            // one pre-skewed box per tile; the utterance's last compute warp goes through the second map, whose lane and frame
            // extents stop at the tensor's last row (lanes past it are zero-filled, never fetched)
            mbar_expect_tx(bar, RW * TF * ES);
            tma_load_3d(st, (g == NC * NW - 1) ? &tmap_tail : &tmap, t * TF, 0, it * p.Tx + xg, bar);
        } else {
            const bool tail = !VL && (g == NC * NW - 1) && p.tail_rows < RW;
            mbar_expect_tx(bar, (tail ? p.tail_rows : RW) * TF * ES);
            if (VL) tma_load_2d(st, &tmap, xg, it * p.Ty + t * TF, bar);      // box = RW tokens x TF frames of [b*t_mel, t_text]
            else tma_load_2d(st, tail ? &tmap_tail : &tmap, t * TF, it * p.Tx + xg, bar);
        }
    };
    // Latency regime, one utterance per CTA: the first two tiles of the first compute warp are requested NOW, before the utterance's
    // lengths are known (the lengths take one or two dependent global round trips, a first tile ~3000 cycles more): a warp's first
    // tile only depends on its row offset, and an active warp always consumes at least two.  An inactive warp (empty utterance)
    // just waits for the two boxes to land before the CTA may exit.
    int pre = 0;
    // Shared zero fill.  In the latency regime the dense output is B x t_text x t_mel zeros written by B of the 148 SMs, and one SM's bulk
    // stores drain at ~30 GB/s: on C3 (32 utterances, 57.6 MB) that is as long as the whole search, and the scatter of the ones has to
    // wait for it (measured without the dense output: C2 41.4 -> 37.3 us, C3 70.1 -> 55.8, C4 293 -> 263).  So the grid is padded with
    // filler CTAs on the idle SMs that only zero-fill: every loader warp of every CTA claims batches of 7 KB chunks of the WHOLE
    // output from one cursor in the workspace header, and a search CTA scatters its ones once all chunks are reported done.
    // Deadlock-free without any co-residency assumption: a CTA that never starts never claims, and the search CTAs' own loaders keep
    // claiming until the cursor is exhausted.
    const bool sz = p.zero_chunks > 0;
    int sz_mine = 0;                                                  // chunks this thread issued
    bool sz_over = false;                                             // the cursor was seen exhausted
    auto sz_claim = [&](int n) {                                     // (lane 0 of a warp) n chunks with one atomic
        unsigned char* zb = reinterpret_cast<unsigned char*>(p.paths);
        const int64_t total = (int64_t)p.B * p.Tx * p.Ty * p.esize;
        const int first = atomicAdd(&p.ws->zero_cursor, n);
        if (first >= p.zero_chunks) { sz_over = true; return false; }
        for (int k = first; k < first + n && k < p.zero_chunks; ++k) {
            const int64_t off = (int64_t)k * kZeroChunk, left = total - off;
            bulk_s2g(zb + off, smem_u32(smem + p.off_zero), (uint32_t)(left < kZeroChunk ? left : kZeroChunk));
            ++sz_mine;
        }
        return true;
    };
    auto sz_report = [&]() {                                          // (lane 0 of a warp) everything this thread issued is in global memory
        if (sz_mine) {
            bulk_commit(); bulk_wait_all(); fence_proxy_async_global();
            __threadfence();
            atomicAdd(&p.ws->zero_done, sz_mine);
            sz_mine = 0;
        }
    };
    auto sz_leave = [&]() {                                           // (one thread per CTA) the last CTA out re-arms the counters
        const int d = atomicAdd(&p.ws->zero_exit, 1);
        if (d == (int)gridDim.x - 1) { p.ws->zero_cursor = 0; p.ws->zero_done = 0; p.ws->zero_exit = 0; __threadfence(); }
    };
    if (sz && (int)blockIdx.x >= p.search_ctas) {                     // ---- filler CTA
        for (int i = threadIdx.x; i < kZeroChunk / 16; i += blockDim.x)
            reinterpret_cast<int4*>(smem + p.off_zero)[i] = make_int4(0, 0, 0, 0);
        fence_proxy_async_smem();
        __syncthreads();
        if ((threadIdx.x & 31) == 0) {
            while (sz_claim(4)) { }
            sz_report();
        }
        __syncthreads();
        if (threadIdx.x == 0) sz_leave();
        return;
    }
#ifndef ALB_EARLY_ALL
#define ALB_EARLY_ALL 0             // A/B builds: 1 = every loader warp, not only the first (measured: C2 +0.6 %, the later warps have time)
#endif
#ifndef ALB_NO_EARLY_TILES
#define ALB_NO_EARLY_TILES 0        // A/B builds: 1 = off (measured with warp 0 only: C1 25.4 -> 24.4 us, C2 and C3 unchanged)
#endif
    if (!ALB_NO_EARLY_TILES && SKEW && NC == 1 && is_loader && p.B <= (int)gridDim.x && p.tile_ready == nullptr && !p.pdl_wait && item < p.B && w * RW < p.Tx && (ALB_EARLY_ALL || w == 0)) {
        pre = NS < 2 ? NS : 2;
        if (lane == 0)
            for (int q = 0; q < pre; ++q) issue_tile(item, w, (w * RW) / TF + q, ring_a + q * L.stage_bytes, full0 + 8 * q);
    }
    const bool dbg_on = kDbgBuild && (p.dbg != nullptr);     // compiled out of the shipped library (tools/dbg_timing.py builds its own)
    long long* dbg = dbg_on ? p.dbg + (int64_t)blockIdx.x * (2 * kMaxWarps + 2) * 2 : nullptr;
    bool first_item = true;
    while (item < p.B) {
        if (dbg_on && first_item && tid == 0) dbg[2 * kMaxWarps * 2] = clock64();
        // ---- lengths
        if (p.mask != nullptr) {
            double ax = 0.0, ay = 0.0;
            mask_partial_any(p.mask, p.mask_dtype, (int64_t)item * p.msb, p.msx, p.msy, p.Tx, Ty, tid, nthr, ax, ay);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ax += __shfl_xor_sync(0xffffffffu, ax, o);
                ay += __shfl_xor_sync(0xffffffffu, ay, o);
            }
            if (lane == 0) { msum[2 * wid] = ax; msum[2 * wid + 1] = ay; }
        } else if (tid == 0) {
            misc[1] = p.t_xs[item];
            misc[2] = p.t_ys[item];
        }
        if (tid < NW) {
            st_flag(tail_a + 4 * tid, -(1 << 28));
            st_flag(head_a + 4 * tid, (crank * NW + tid) * RW);
        }
        if (tid == 0) st_flag(xout_head_a, (crank + 1) * NW * RW);
        if (p.durations != nullptr)
            for (int i = tid; i < TXS; i += nthr) durS[i] = 0;
        if (NC > 1) cluster_sync_all();      // every CTA's flags and rings are initialised before any remote store can land
        else __syncthreads();
        int t_x, t_y;
        if (p.mask != nullptr) {          // sum the per-warp partials; truncation like astype(np.int32) (__init__.py:18-19)
            double ax = 0.0, ay = 0.0;
            for (int q = 0; q < 2 * NW; ++q) { ax += msum[2 * q]; ay += msum[2 * q + 1]; }
            t_x = (int)ax; t_y = (int)ay;
        } else {
            t_x = misc[1]; t_y = misc[2];
        }
        if (p.lens_out != nullptr && tid == 0) {
            p.lens_out[item] = t_x;
            p.lens_out[p.B + item] = t_y;
        }
        bool valid = true;
        if (t_x <= 0 || t_y <= 0) valid = false;                       // empty item: all-zero path
        else if (t_x > t_y || t_x > p.Tx || t_y > Ty) {                // reference reads out of bounds here
            valid = false;
            if (tid == 0) atomicOr(&p.ws->status, 1);
        }
        if (!valid) { t_x = 0; t_y = 0; }
        if (dbg_on && first_item && lane == 0) dbg[wid * 2] = clock64();
        if (tid == 0) st_flag(bt_cur_a, ((t_y - 1) >> 5) + 1);     // nothing published yet (ordered by the barrier after the forward pass)

        // ---- geometry of this warp's slice of the band
        const int x0 = gw * RW;
        const bool active = valid && x0 < t_x;
        const int x1 = (x0 + RW < t_x) ? x0 + RW : t_x;
        const int nrows = x1 - x0;
        const int y_start = x0;                                 // first frame where any of our rows is on/below the diagonal
        const int y_last = t_y - t_x + x1 - 1;                  // last frame where our last row is inside the band (core.pyx:18)
        const int span = (y_last + 1 - y_start + 31) & ~31;     // whole 32-frame direction words
        // lane-0 frames; skewed: lane 31 needs 31 more -- with the 4-frame lag the last live lane needs 4 * (lanes - 1) more, and a
        // lane emits an aligned direction word (lag / 32) + 1 units after lane 0 would
        const int live_lanes = nrows > 0 ? (nrows + R - 1) / R : 1;
        const int y_end = y_start + span + (SKEW ? (L4 ? 32 + ((kLag4 * (live_lanes - 1)) & ~31) : 32) : 0);
        const int t_s = y_start / TF, t_e = y_end / TF;         // tiles [t_s, t_e), all whole

        if (is_loader) {
            // ================= loader warp: tile stream + fused zero fill =================
            const bool zf = p.zero_fill && p.paths != nullptr && !sz;        // (shared zero fill: claimed from the global cursor instead)
            unsigned char* pbase = reinterpret_cast<unsigned char*>(p.paths) + item * item_elems * p.esize;
            unsigned char* zA = nullptr;
            int64_t zbytes = 0;
            const int ltid = crank * NW * 32 + (tid - NW * 32), lthr = NC * NW * 32;   // loader thread index over the whole cluster
            if (zf) {
                unsigned char* pend = pbase + item_elems * p.esize;
                zA = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(pbase) + 15) & ~uintptr_t(15));
                unsigned char* zE = reinterpret_cast<unsigned char*>(reinterpret_cast<uintptr_t>(pend) & ~uintptr_t(15));
                if (zA >= zE) {   // tiny or fully misaligned item: plain byte stores
                    for (int64_t i = ltid; i < item_elems * p.esize; i += lthr) pbase[i] = 0;
                } else {
                    const int head = (int)(zA - pbase), tail = (int)(pend - zE);
                    if (ltid < head) pbase[ltid] = 0;
                    if (ltid < tail) zE[ltid] = 0;
                    zbytes = zE - zA;
                }
            }
            const int nact = valid ? (t_x + RW - 1) / RW : NC * NW;   // loader warps that take part in the zero fill
            const int64_t nchunks = (zbytes + kZeroChunk - 1) / kZeroChunk;
            int64_t zc = gw;                                       // next chunk this warp issues
            auto issue_zero = [&](int n) {
                if (lane == 0) {
                    for (int q = 0; q < n && zc < nchunks; ++q, zc += nact) {
                        const int64_t off = zc * kZeroChunk;
                        const int64_t left = zbytes - off;
                        bulk_s2g(zA + off, zero_a, (uint32_t)(left < kZeroChunk ? left : kZeroChunk));
                    }
                }
            };
            if (active) {
                const unsigned char* vrow = reinterpret_cast<const unsigned char*>(p.values) + (item * item_elems + (int64_t)x0 * Ty) * ES;
                const int band_hi0 = t_y - t_x + x0;               // last live frame of row i is band_hi0 + i
                constexpr int FPC = 16 / ES;                        // frames per 16-byte chunk
                constexpr int CPR = TF / FPC;                       // 16-byte chunks per row
                constexpr int RPI = 32 / CPR;                       // row segments per warp instruction
                const int ck = lane % CPR, q0 = lane / CPR;
                const int my_tiles = t_e - t_s;
                const int64_t my_chunks = (nchunks > gw) ? (nchunks - gw + nact - 1) / nact : 0;
                const int zq = (int)((my_chunks + my_tiles - 1) / (my_tiles > 0 ? my_tiles : 1));
                long long l_e = 0, l_c = 0, l_z = 0, l0 = 0, l1 = 0, l2 = 0;
                int known_ready = -1;                               // highest producer tile known to be published (pipelined mode)
                if (p.pdl_wait) asm volatile("griddepcontrol.wait;" ::: "memory");   // the scores are complete and visible from here on
                for (int t = t_s; t < t_e; ++t) {
                    if (dbg_on) l0 = clock64();
                    // the compute lanes have released this stage.  (Not for the tiles requested before the lengths were known: their
                    // stages were free by construction, and the compute warp may consume such a tile and release its stage before this
                    // loop gets here -- a parity wait would then wait for a phase that never comes.)
                    if (t - t_s >= pre) mbar_wait(empty0 + 8 * stage, phase ^ 1u);
                    if (p.tile_ready != nullptr) {                 // scores still being produced: wait for the frames of this tile
                        int need = (t * TF + TF - 1) >> 7;
                        need = need < p.ready_tiles ? need : p.ready_tiles - 1;
                        if (known_ready < need) wait_tiles_ready(p.tile_ready + (int64_t)item * p.ready_tiles, p.ready_epoch, known_ready, need);
                    }
                    if (dbg_on) l1 = clock64();
                    const uint32_t st = ring_a + stage * L.stage_bytes;
                    if (SKEW) {
                        // one box load per tile: all 32*R rows of this warp x TF frames (the host only picks this form for 16-byte
                        // aligned inputs); rows past t_x and frames past T_mel are fetched or zero-filled, never used
                        if (lane == 0 && t - t_s >= pre) issue_tile(item, gw, t, st, full0 + 8 * stage);      // (the first `pre` are already on their way)
                    } else if (p.aligned) {
                        // Loader lane = (16-byte chunk ck of a row, row group q0); it walks the owner lanes li = q0, q0+RPI, ...
                        // and their R rows, so one warp-wide LDGSTS.128 moves RPI whole row segments.
                        const int f0 = t * TF + ck * FPC;
                        for (int li = q0; li * R < nrows; li += RPI) {
                            const int f = f0;                                   // frame of this chunk
                            const int lo = (x0 + li * R) & ~(FPC - 1);          // chunks wholly above the diagonal are never read
                            const uint32_t d = st + li * LANE_STRIDE + ck * 16;
                            const unsigned char* src = vrow + ((int64_t)(li * R) * Ty + f) * ES;
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                const int i = li * R + r;
                                if (i < nrows && f >= lo && f <= band_hi0 + i && f + FPC <= Ty)
                                    cp_async16(d + r * (TF * ES), src + (int64_t)r * Ty * ES);
                            }
                        }
                        cp_async_arrive(full0 + 8 * stage);
                    } else {
                        // unaligned base pointer or row pitch: plain element loads, same placement
                        unsigned char* sp = smem + L.off_ring + (size_t)(w * NS + stage) * L.stage_bytes;
                        for (int idx = lane; idx < nrows * TF; idx += 32) {
                            const int i = idx / TF, fl = idx - i * TF, li = i / R;
                            const int f = t * TF + fl;
                            if (f >= 0 && f <= band_hi0 + i && f < Ty) {
                                const size_t so = (size_t)i * (TF * ES) + (size_t)li * kLanePad + (size_t)fl * ES;
                                const size_t go = ((int64_t)i * Ty + f) * ES;
                                if (ES == 4) *reinterpret_cast<float*>(sp + so) = *reinterpret_cast<const float*>(vrow + go);
                                else *reinterpret_cast<unsigned short*>(sp + so) = *reinterpret_cast<const unsigned short*>(vrow + go);
                            }
                        }
                        mbar_arrive(full0 + 8 * stage);
                    }
                    if (++stage == (uint32_t)NS) { stage = 0; phase ^= 1u; }
                    if (dbg_on) l2 = clock64();
                    if (zf) issue_zero(zq);
                    // shared zero fill: the filler CTAs do most of it meanwhile; should they not be running (SMs taken by another
                    // kernel), 16 chunks every eighth tile keep this CTA's share of the zeros under its forward pass as before.  One
                    // global atomic round trip per claim: every tile was measured (C2 38.3 -> 43.4 us), every fourth costs C4 3 %.
                    if (sz && lane == 0 && !sz_over && ((t - t_s) & 7) == 7) sz_claim(16);
                    if (dbg_on) { const long long l3 = clock64(); l_e += l1 - l0; l_c += l2 - l1; l_z += l3 - l2; }
                }
                if (dbg_on && first_item && lane == 0) {
                    long long* e = p.dbg + (int64_t)gridDim.x * ((2 * kMaxWarps + 2) * 2 + kMaxWarps * 4) + ((int64_t)blockIdx.x * kMaxWarps + w) * 4;
                    e[0] = l_e; e[1] = l_c; e[2] = l_z; e[3] = t_e - t_s;
                }
            }
            if (!active && pre) {                                  // early boxes nobody consumes: landed before this CTA can exit
                for (int q = 0; q < pre; ++q) mbar_wait(full0 + 8 * q, 0u);
            }
            pre = 0;
            if (zf && gw < nact) {
                issue_zero(0x7fffffff);
                if (lane == 0) { bulk_commit(); bulk_wait_all(); fence_proxy_async_global(); }
            }
            if (sz && lane == 0) {                                 // whatever nobody has taken by now, then report
                while (!sz_over && sz_claim(16)) { }
                sz_report();
            }
        } else if (active) {
            // ================= compute warp: forward pass for rows [x0, x1) =================
            const bool has_in = (gw > 0);
            const bool has_consumer = (x1 < t_x);
            const bool remote_in = (w == 0 && crank > 0);                        // our producer is the last warp of the previous CTA
            const bool remote_out = (NC > 1 && w == NW - 1 && has_consumer);               // our consumer is warp 0 of the next CTA (has_consumer => there is one)
            const bool lane0 = (lane == 0), lane31 = (lane == 31);
            const int xl0 = x0 + lane * R;
            const bool row0 = (xl0 == 0);                      // this lane's first row is token 0
            const int lag = SKEW ? LAG * lane : 0;
            const uint32_t bin_addr = bnd_a + w * RING * 4;            // warp 0 reads the constant sentinel ring: x == 0, y > 0: v_prev = max_neg_val (core.pyx:27)
            const uint32_t bout_addr = bnd_a + (w + 1) * RING * 4;
            const uint32_t my_tail = tail_a + 4 * w, my_head = head_a + 4 * w;
            const uint32_t prog_addr = lane31 ? my_tail : (lane0 ? my_head : xout_head_a + 4);   // where this lane publishes the unit's end
            const uint32_t in_tail = tail_a + 4 * (w > 0 ? w - 1 : 0);
            const uint32_t out_head = remote_out ? xout_head_a : head_a + 4 * (has_consumer ? w + 1 : w);
            // Cluster hand-off: the producer (last warp of CTA c) writes its local ring as usual and, once per unit, lane 31
            // reads its own 32 boundary values back and sends them to ring 0 of CTA c+1 with eight st.async (STAS.128) whose
            // bytes complete on an mbarrier over there; the consumer arms (expect_tx 128) and waits on that mbarrier.  No
            // fences on either side.  (Inside the unit body the stores turn into one branch per group, ptxas will not
            // predicate STAS; a shared->remote bulk copy needs fence.proxy.async, ~400 cycles -- profiles/r01_notes.md.)
            const uint32_t r_ring = remote_out ? mapa_u32(bnd_a, (uint32_t)crank + 1) : 0u;         // next CTA's ring 0
            const uint32_t r_xbar = remote_out ? mapa_u32(xbar_a, (uint32_t)crank + 1) : 0u;
            const uint32_t r_head = remote_in ? mapa_u32(xout_head_a, (uint32_t)crank - 1) : 0u;    // where the previous CTA polls our progress
            // producer geometry as seen by a remote consumer (the producer is a full warp of rows [x0 - RW, x0))
            const int p_start = x0 - RW;
            const int p_units = (((t_y - t_x + x0 - p_start + 31) & ~31) + 32) >> 5;                  // its (y_end - y_start) / 32
            int xk = 0;                                          // producer units received so far
            auto wait_remote = [&](int progress_needed) {        // producer progress is counted in its lane-0 steps
                int k_need = ((progress_needed - p_start) >> 5) - 1;
                if (k_need > p_units - 1) k_need = p_units - 1;
                while (xk <= k_need) {
                    if (lane0) mbar_expect_tx(xbar_a + 8 * (xk & 3), 128);
                    mbar_wait(xbar_a + 8 * (xk & 3), (uint32_t)(xk >> 2) & 1u);
                    ++xk;
                }
            };
            const int diag_end = SKEW ? x0 + RW + LAG31 : x1;      // lane-0 frame from which no lane holds a row any more

            Fwd<R> S;
#pragma unroll
            for (int r = 0; r < R; ++r) { S.old[r] = neg; S.wbits[r] = 0u; }
#pragma unroll
            for (int r = 0; r < R; ++r) S.wprev[r] = 0u;
            S.up = neg; S.u1 = neg; S.u2 = neg; S.u3 = neg; S.u4 = neg; S.u5 = neg;
            S.bprev = neg; S.bprev1 = neg;
            if (!has_in) {
                S.bprev = 0.f;                                      // x == 0, y == 0: v_prev = 0 (core.pyx:25)
            } else {
                if (CL && remote_in) wait_remote(y_start + UNIT + 32);
                else wait_flag_ge(in_tail, y_start + UNIT + (SKEW ? (L4 ? 32 * kLag4 : 32) : 0), -(1 << 30));
                if (L4) {
                    S.bprev = lds32(bin_addr + (uint32_t)(((y_start + 31 * kLag4 - 1) & (RING - 1)) << 2));   // frame y_start - 1 of the row above
                } else if (SKEW) {
                    const float4 b = lds128(bin_addr + (((y_start + 28) & (kRing - 1)) << 2));   // frames y_start-4 .. y_start-1 (+31)
                    S.bprev = b.z; S.bprev1 = b.w;
                } else {
                    S.bprev = lds32(bin_addr + (((y_start - 4) & (kRing - 1)) << 2) + 12);   // V[x0-1, x0-1], the diagonal cell above us
                }
            }
            int seen_cons = has_consumer ? 0 : kProgDone;
            uint32_t prev_stage = 0;
            constexpr int IN_LEAD = SKEW ? (L4 ? 32 * kLag4 : 32) : 0;   // skewed: the producer's lane 31 trails its lane 0 by 31 (124) steps
            int seen_in = (has_in && !remote_in) ? y_start + UNIT + IN_LEAD : kProgDone;   // producer progress (in its steps) as last read
            uint32_t* bits_row = bits + xl0;
            // skewed forms: the block whose aligned word this lane completes in the unit that starts at y_start (a block before the
            // band for the first unit(s): never dereferenced), advanced by one block row per unit
            uint32_t* bits_unit = bits_row + ((int64_t)((y_start - 32) >> 5) - (L4 ? (lag >> 5) : 0)) * TXS;
            bool tile_ok = false;                              // next tile's "full" barrier already seen complete
            long long c_full = 0, c_poll = 0, c_unit = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0;   // ALB200_DBG cycle breakdown

            for (int y = y_start; y < y_end; y += UNIT) {           // y = frame of lane 0
                const int fin = y & (TF - 1);
                if (dbg_on) c0 = clock64();
                // our lane 31 is about to overwrite these ring slots (4-frame lag: slots [y - 256, y - 224) were last read by the
                // consumer's unit y - 256 - 96, finished once its head flag reads y - 320)
                const int need_in = y + UNIT + IN_LEAD;
                const int need_cons = L4 ? y + UNIT - RING - 96 : (SKEW ? y + UNIT - 32 - kRing : y + UNIT - (kRing - 4));
                // In steady state the tile has landed (probe below) and both neighbours were seen far enough along after the last
                // unit: ONE branch guards the three waits (each wait as its own branch region cost ~40 cycles of the ~260-cycle head).
                if (!((fin != 0 || tile_ok) && seen_in >= need_in && seen_cons >= need_cons)) {
                    if (fin == 0 && !tile_ok) mbar_wait(full0 + 8 * stage, phase);
                    seen_in = wait_flag_ge(in_tail, need_in, seen_in);
                    seen_cons = wait_flag_ge(out_head, need_cons, seen_cons);
                }
                if (dbg_on) c1 = clock64();
                if (CL && remote_in) wait_remote(y + UNIT + IN_LEAD);
                // read now, needed after this unit (latency hidden): both neighbours' progress.  (Probing the next tile's barrier
                // here with mbarrier.test_wait was measured: the probe itself costs ~800 cycles per unit -- profiles/r01_notes.md.)
                const int next_in = (has_in && !remote_in) ? ld_flag(in_tail) : kProgDone;
                const int next_cons = has_consumer ? ld_flag(out_head) : kProgDone;
                if (dbg_on) c2 = clock64();
                constexpr int LSTR = VL ? R * ES : LANE_STRIDE;  // where a lane's first token starts inside a tile
                const uint32_t tile_addr = ring_a + stage * L.stage_bytes + lane * LSTR + fin * ES;
                const uint32_t tile_prev = (SKEW && y > y_start) ? ring_a + prev_stage * L.stage_bytes + lane * LSTR : tile_addr;
                const int yl = y - lag;
                if (L4) {
                    const uint32_t tile_lane = ring_a + stage * L.stage_bytes + lane * 128;
                    const uint32_t xs = (uint32_t)(lane & 7) << 4;
                    if (y < diag_end)
                        forward_unit4<R, true>(S, tile_lane, xs, bin_addr, bout_addr, y, yl, lane0, lane31, neg, xl0 - yl, bits_unit, y_start, (unsigned)span, row0);
                    else
                        forward_unit4<R, false>(S, tile_lane, xs, bin_addr, bout_addr, y, yl, lane0, lane31, neg, 0, bits_unit, y_start, (unsigned)span, row0);
                } else if (y < diag_end)
                    forward_unit<R, TF, UNIT, SKEW, true, VT, VL>(S, tile_addr, tile_prev, bin_addr, bout_addr, y, yl, lane0, lane31, neg,
                                                          xl0 - yl, bits_row, TXS, y_start, (unsigned)span, bits_unit, row0);
                else
                    forward_unit<R, TF, UNIT, SKEW, false, VT, VL>(S, tile_addr, tile_prev, bin_addr, bout_addr, y, yl, lane0, lane31, neg,
                                                           0, bits_row, TXS, y_start, (unsigned)span, bits_unit, row0);
                if (SKEW) bits_unit += TXS;
                seen_in = next_in;
                seen_cons = next_cons;
                if (dbg_on) { c3 = clock64(); c_full += c1 - c0; c_poll += c2 - c1; c_unit += c3 - c2; }
                st_flag(prog_addr, y + UNIT);        // lane 31 -> tail, lane 0 -> head, the rest -> scratch: one store, no branch
                if (CL) {
                    if (remote_out && lane31) {                  // ship this unit's 32 boundary values (written by this very lane)
                        const uint32_t off = (uint32_t)(y & (kRing - 1)) << 2;
                        const int k = (y - y_start) >> 5;
#pragma unroll
                        for (int g = 0; g < 8; ++g)              // same thread wrote these slots: program order, no fence
                            st_async_v4(r_ring + off + 16 * g, lds128(bout_addr + off + 16 * g), r_xbar + 8 * (k & 3));
                    }
                    if (remote_in && lane0) st_cluster_flag(r_head, y + UNIT);
                }
                if (((y + UNIT) & (TF - 1)) == 0) {                 // tile consumed: hand the stage back to the loader
                    if (SKEW && !L4) {                              // the trailing lanes still read this tile during the next unit
                        if (y > y_start) mbar_arrive(empty0 + 8 * prev_stage);
                        prev_stage = stage;
                    } else {
                        mbar_arrive(empty0 + 8 * stage);
                    }
                    if (++stage == (uint32_t)NS) { stage = 0; phase ^= 1u; }
                    // The next tile has usually landed long ago: one non-blocking probe here, in the shadow of the flag stores, instead of
                    // a barrier round trip at the head of the next unit (C1 29.4 -> 28.6 us, C2 42.2 -> 41.6 us).  (Round 1 probed BEFORE
                    // the unit body and lost 800 cycles per unit: the volatile asm sat in the middle of the frame arithmetic.)
                    tile_ok = mbar_test_wait(full0 + 8 * stage, phase);
                }
            }
            if (SKEW && !L4 && y_end > y_start) mbar_arrive(empty0 + 8 * prev_stage);
            if (lane31) st_flag(my_tail, kProgDone);
            if (CL && remote_in) wait_remote(0x3fffffff);          // every incoming copy has landed before this CTA may leave the cluster barrier
            if (dbg_on && first_item && lane == 0) {
                long long* e = p.dbg + (int64_t)gridDim.x * (2 * kMaxWarps + 2) * 2 + ((int64_t)blockIdx.x * kMaxWarps + w) * 4;
                e[0] = c_full; e[1] = c_poll; e[2] = c_unit; e[3] = (y_end - y_start) / UNIT;
            }
        }
        if (dbg_on && first_item && lane == 0) dbg[wid * 2 + 1] = clock64();
        if (NC > 1) cluster_sync_all();      // all CTAs' direction bits (L2 slot) and zero fill are complete and visible
        else __syncthreads();
        if (sz) {                            // shared zero fill: the ones may only be scattered once every chunk of zeros has landed
            if (tid == 0) {
                uint32_t n = 0;
                for (;;) {
                    int done;
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(done) : "l"(&p.ws->zero_done) : "memory");
                    if (done >= p.zero_chunks) break;
                    __nanosleep(100);
                    spin_guard(n);
                }
            }
            __syncthreads();
        }

        // ================= backtrack =================
        // Warp 0 walks: 32 frames per step, one find-leading-one chain per step DOWN (about t_x/t_y of the frames).  It only
        // publishes (token at the block's last frame, mask of step frames) per block; the other warps turn those into
        // stores, so address arithmetic and memory traffic are off the serial chain.
        const int top = (crank == 0) ? (t_y - 1) >> 5 : -1;      // cluster: CTA 0 backtracks, the others are done
        if (wid == 0) {
            long long* dbg_bt = kDbgBuild && dbg_on && first_item ? p.dbg + (int64_t)gridDim.x * (2 * kMaxWarps + 2) * 2 + ((int64_t)blockIdx.x * kMaxWarps + (kMaxWarps - 1)) * 4 : nullptr;
            if (SKEW) {     // words stored in their final form
                if (bits_smem) backtrack_walk_direct(smem0 + L.off_bits, TXS, t_x, t_y, top, lane, btTok, btMov, bt_cur_a, dbg_bt);
                else           backtrack_walk<4, true>(bits, TXS, t_x, t_y, top, lane, smem0 + L.off_ring, btTok, btMov, bt_cur_a, dbg_bt);
            } else {
                if (bits_smem) backtrack_walk<2, false>(bits, TXS, t_x, t_y, top, lane, smem0 + L.off_ring, btTok, btMov, bt_cur_a, dbg_bt);
                else           backtrack_walk<4, false>(bits, TXS, t_x, t_y, top, lane, smem0 + L.off_ring, btTok, btMov, bt_cur_a, dbg_bt);
            }
            fence_proxy_async_smem();   // the row windows went through the generic proxy into ring memory that TMA writes next
        } else {
            if (p.frame_tok != nullptr && crank == 0)
                for (int yy = t_y + (tid - 32); yy < Ty; yy += nthr - 32) p.frame_tok[(int64_t)item * Ty + yy] = -1;
            const int nemit = (nthr >> 5) - 1;                    // every warp but the walker
            for (int blk = top - (wid - 1); blk >= 0; blk -= nemit) {
                wait_flag_le(bt_cur_a, blk);
                const int tokb = btTok[blk];
                const uint32_t moves = btMov[blk];
                const int yb = blk << 5;
                const int nvalid = (t_y - yb < 32) ? t_y - yb : 32;
                if (lane < nvalid) {
                    const int tok = tokb - __popc(moves & ((1u << (31 - lane)) - 1u));   // steps at frames above ours (mask is bit-reversed)
                    const int yy = yb + lane;
                    if (p.paths != nullptr) store_one(p.paths, item * item_elems + (VL ? (int64_t)yy * p.Tx + tok : (int64_t)tok * Ty + yy), p.esize, p.one);
                    if (p.frame_tok != nullptr) p.frame_tok[(int64_t)item * Ty + yy] = tok;
                    if (p.durations != nullptr) atomicAdd(&durS[tok], 1);
                }
            }
        }
        __syncthreads();
        if (p.durations != nullptr && crank == 0)
            for (int i = tid; i < p.Tx; i += nthr) p.durations[(int64_t)item * p.Tx + i] = (i < t_x) ? durS[i] : 0;

        if (dbg_on && first_item && tid == 0) dbg[2 * kMaxWarps * 2 + 1] = clock64();
        first_item = false;
        // ---- next item
        if (NC > 1 || p.B <= (int)gridDim.x) break;
        if (tid == 0) {
            misc[0] = atomicAdd(&p.ws->counter, 1) + (int)gridDim.x;
        }
        __syncthreads();
        item = misc[0];
    }

    if (kDbgBuild && p.tile_ready != nullptr && tid == 0) {
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        atomicMax(reinterpret_cast<unsigned long long*>(const_cast<int*>(p.tile_ready) + (int64_t)p.B * p.ready_tiles + 64) + 3, gt);
    }
    if (sz && tid == 0) sz_leave();
    if (NC == 1 && p.B > (int)gridDim.x && tid == 0) {
        const int d = atomicAdd(&p.ws->done, 1);
        if (d == (int)gridDim.x - 1) {      // last CTA out re-arms the counters for the next launch
            p.ws->counter = 0;
            p.ws->done = 0;
            __threadfence();
        }
    }
}
#endif  // __CUDACC__

}  // namespace alb
