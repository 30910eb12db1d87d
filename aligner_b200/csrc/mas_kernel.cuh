// mas_kernel.cuh -- monotonic alignment search for sm_100a.
//
// One CTA per utterance (persistent, work-stealing over the batch).  The text
// axis is spread over threads: lane l of warp w owns R consecutive token rows
// and keeps their running score column in registers; the DP marches along the
// mel axis.  What the reference does per item (monotonic_align/core.pyx:17-35):
//
//   forward   V[x,y] = max(V[x,y-1], V[x-1,y-1]) + value[x,y]   (core.pyx:19-30)
//   backtrack walk y = t_y-1 .. 0 choosing x-1 iff V[x,y-1] < V[x-1,y-1] (core.pyx:32-35)
//
// is restated as: running fp32 column + ONE direction bit per cell
// (bit = v_prev > v_cur, the very predicate the backtrack re-evaluates), then a
// bit-driven backtrack.  oracle/mas_oracle.c:mas_oracle_bits is the CPU twin of
// this formulation and is proven equal to the table form by tests/test_oracle.py.
//
// Data movement
//   * every warp streams ITS OWN rows: 128-bit asynchronous copies (cp.async.cg,
//     SASS LDGSTS.128) of TF frames per row land in a per-warp ring of NS stages
//     and complete on a per-stage mbarrier (cp.async.mbarrier.arrive).  Rows are
//     placed with a 16-byte skew per lane so the lanes' 128-bit shared loads are
//     bank-conflict free for any R.  (Per-row cp.async.bulk copies were measured
//     first: ~60-90 cycles of TMA issue per 128-byte row made the loader the
//     bottleneck, 8x slower end to end -- profiles/r01_notes.md.)
//   * only tiles inside the reference's band (core.pyx:18) are fetched.
//   * warps are a dataflow pipeline: warp w consumes the last row of warp w-1
//     through a small shared ring + progress flag, 16 frames at a time.  The
//     diagonal band gives the pipeline skew for free (warp w starts at frame
//     32*R*w), so there is no CTA-wide barrier inside the forward pass.
//   * the dense 0/1 output is zero-filled with bulk shared->global stores issued
//     along the forward pass (fused memset), then the backtrack drops the ones.
//   * direction bits live in shared memory when they fit, else in an L2-resident
//     per-CTA slot of the workspace.
//   * backtrack: one warp, 32 frames per step.  Lane l fetches the direction
//     word of row (tok - l); 32 ballots transpose the 32x32 bit block; the walk
//     itself is two dependent integer ops per frame on a one-hot position.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace alb {

constexpr int kMaxWarps = 16;
constexpr int kRing = 64;          // floats in a warp-boundary ring (4 units of 16 frames)
constexpr int kZeroChunk = 8192;   // bytes per zero-fill bulk store
constexpr int kLanePad = 16;       // bytes of skew per lane inside a tile stage
constexpr int kProgDone = 0x3fffffff;

struct WsHeader {       // first 64 bytes of the workspace
    int counter;        // work-stealing cursor (self-resetting)
    int done;           // CTAs that left the item loop
    int status;         // bit0: invalid lengths seen
    int pad[13];
};

struct MasParams {
    const float* values;
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
    long long* dbg;             // optional [grid][kMaxWarps+2][2] clock64 stamps (ALB200_DBG)
    uint64_t one;
    int64_t bits_slot_words;
    int B, Tx, Ty;
    int esize;
    int mask_dtype;
    int zero_fill;
    int ns;                     // ring stages per warp
    int nblk;                   // ceil(Ty/32)
    int aligned;                // values base and Ty allow 16-byte bulk copies
    float neg;
};

struct SmemLayout {
    uint32_t off_bar, off_prog, off_misc, off_bnd, off_zero, off_ring, off_bits, off_dur, total;
    uint32_t stage_bytes;
};

__host__ __device__ inline uint32_t alb_align(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline SmemLayout make_layout(int NW, int NS, int R, int TF, int bits_smem, int nblk, int want_dur)
{
    SmemLayout L;
    const uint32_t RW = 32u * R;
    L.stage_bytes = RW * TF * 4 + 32 * kLanePad;
    uint32_t o = 0;
    L.off_bar = o;  o += NW * NS * 8;
    L.off_prog = alb_align(o, 16); o = L.off_prog + (2 * NW + 2) * 4;   // tail (lane 31) and head (lane 0) progress per warp
    L.off_misc = alb_align(o, 16); o = L.off_misc + 64 + 2 * 16 * 8;     // item/lengths + per-warp partial mask sums
    L.off_bnd = alb_align(o, 16);  o = L.off_bnd + NW * kRing * 4;
    L.off_zero = alb_align(o, 128); o = L.off_zero + kZeroChunk;
    L.off_ring = alb_align(o, 128); o = L.off_ring + NW * NS * L.stage_bytes;
    L.off_bits = alb_align(o, 16); o = L.off_bits + (bits_smem ? (uint32_t)nblk * NW * RW * 4 : 0);
    L.off_dur = alb_align(o, 16);  o = L.off_dur + (want_dur ? NW * RW * 4 : 0);
    L.total = alb_align(o, 16);
    return L;
}

#ifdef __CUDACC__
// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) { }
}
// global -> shared bulk copy (TMA engine), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// 16-byte asynchronous global -> shared copy (LDGSTS), L2 only
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// the mbarrier receives one arrival when all of this thread's earlier cp.async have landed
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// shared -> global bulk store
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
__device__ __forceinline__ void sts32(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v));
}
// Progress flags between neighbouring warps of one CTA.  Producer: the SAME lane stores the boundary
// values and then the flag; consumer: reads the flag, then the values.  Shared-memory accesses of one
// thread are performed in program order by the SM's in-order LSU pipe, so plain volatile accesses are
// sufficient; ld.acquire/st.release compile to MEMBAR.ALL.CTA, which also drains this thread's in-flight
// LDGSTS tile loads and cost ~1 us per 16-frame unit (measured, profiles/r01_notes.md).
__device__ __forceinline__ int ld_flag(uint32_t a) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void st_flag(uint32_t a, int v) {
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// ------------------------------------------------------------------ mask -> length
// reference: t_x = mask.sum(1)[:,0], t_y = mask.sum(2)[:,0], astype(int32)  (__init__.py:18-19)
// Every thread of the CTA takes elements tid, tid+n, ... of the concatenation [mask[b,:,0] ; mask[b,0,:]] with four
// independent loads in flight (the first version walked them one dependent load at a time: ~15% of the kernel).
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
    for (int i0 = tid; i0 < n; i0 += 4 * nthr) {
        T v[4];
        int idx[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            idx[q] = i0 + q * nthr;
            if (idx[q] < n) v[q] = m[idx[q] < Tx ? (int64_t)idx[q] * sx : (int64_t)(idx[q] - Tx) * sy];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
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

// ------------------------------------------------------------------ forward unit
// UNIT consecutive frames starting at frame y for this lane's R rows.
//   old[r]   running column (value at frame y-1 on entry, y+UNIT-1 on exit)
//   up       value of row (first row - 1) at the previous frame, for lanes > 0
//   hb[r]    direction bits of this unit, bit kk = frame y+kk
template <int R, int TF, int UNIT, bool DIAG>
__device__ __forceinline__ void mas_unit(float (&old)[R], float& up, uint32_t (&hb)[R], uint32_t tile_addr,
                                         uint32_t bin_addr, uint32_t bout_addr, int y, int w, int lane,
                                         float neg, int dxy)
{
#pragma unroll
    for (int g = 0; g < UNIT / 4; ++g) {
        float4 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = lds128(tile_addr + r * (TF * 4) + g * 16);
        float4 bin;
        if (w > 0) {
            bin = lds128(bin_addr + (((y + 4 * g) & (kRing - 1)) << 2));
        } else {
            bin = make_float4(neg, neg, neg, neg);   // x == 0, y > 0: v_prev = max_neg_val (core.pyx:27)
            if (y + 4 * g == 0) bin.x = 0.f;         // x == 0, y == 0: v_prev = 0          (core.pyx:25)
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int kk = 4 * g + k;
            const float bk = (k == 0) ? bin.x : (k == 1) ? bin.y : (k == 2) ? bin.z : bin.w;
            const float upv = (lane == 0) ? bk : up;
            float nv[R];
#pragma unroll
            for (int r = R - 1; r >= 0; --r) {
                const float stay = old[r];                          // v_cur  (core.pyx:22; == neg on the diagonal)
                const float move = (r == 0) ? upv : old[r - 1];     // v_prev (core.pyx:29)
                const bool take = move > stay;                      // core.c:19384
                const float vr = (k == 0) ? v[r].x : (k == 1) ? v[r].y : (k == 2) ? v[r].z : v[r].w;
                float res = (take ? move : stay) + vr;              // core.pyx:30
                if (DIAG) res = (dxy + r > kk) ? neg : res;         // rows above the diagonal stay at the sentinel
                nv[r] = res;
                if (take) hb[r] |= (1u << kk);
            }
            up = __shfl_up_sync(0xffffffffu, nv[R - 1], 1);
            if (lane == 31) sts32(bout_addr + (((y + kk + 1) & (kRing - 1)) << 2), nv[R - 1]);
#pragma unroll
            for (int r = 0; r < R; ++r) old[r] = nv[r];
        }
    }
}

// ------------------------------------------------------------------ forward unit, lane-skewed (systolic) form
// Lane l runs 4 frames behind lane l-1: at the same instruction it works on frame (Y - 4*l).  The neighbour's value
// a lane needs was produced five frames earlier, so the shuffle that fetches it is issued four frames before its use
// and its latency never sits on the per-frame dependency chain (in the lock-step form above it does, every frame).
// Cost: 4 frames of pipeline fill per lane.  Tiles are loaded with the same per-lane skew, so in lane-local terms
// the shared-memory addressing is identical to the lock-step form.
//   upn[k]   neighbour value for frame k of the NEXT group of four (fetched during this group)
//   lastp    this lane's last-row value of the previous frame (what the next shuffle ships)
//   wbits[r] 32-frame direction word being shifted in from the top, 4 bits per group
template <int R, int TF, int UNIT, bool DIAG>
__device__ __forceinline__ void mas_unit_skew(float (&old)[R], float (&upn)[4], float& lastp, uint32_t (&wbits)[R],
                                              uint32_t tile_addr, uint32_t bin_addr, uint32_t bout_addr, int Y, int yl,
                                              int w, int lane, float neg, int dxy, uint32_t* bits_row, int TXS, int y_lo,
                                              unsigned span)
{
#pragma unroll
    for (int g = 0; g < UNIT / 4; ++g) {
        float4 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = lds128(tile_addr + r * (TF * 4) + g * 16);
        float4 bin;
        if (w > 0) {
            bin = lds128(bin_addr + (((Y + 4 * g) & (kRing - 1)) << 2));
        } else {
            bin = make_float4(neg, neg, neg, neg);
            if (Y + 4 * g == 0) bin.x = 0.f;
        }
        uint32_t hb[R];
#pragma unroll
        for (int r = 0; r < R; ++r) hb[r] = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int kk = 4 * g + k;
            const float bk = (k == 0) ? bin.x : (k == 1) ? bin.y : (k == 2) ? bin.z : bin.w;
            const float upv = (lane == 0) ? bk : upn[k];
            float nv[R];
#pragma unroll
            for (int r = R - 1; r >= 0; --r) {
                const float stay = old[r];
                const float move = (r == 0) ? upv : old[r - 1];
                const bool take = move > stay;
                const float vr = (k == 0) ? v[r].x : (k == 1) ? v[r].y : (k == 2) ? v[r].z : v[r].w;
                float res = (take ? move : stay) + vr;
                if (DIAG) res = (dxy + r > kk) ? neg : res;
                nv[r] = res;
                if (take) hb[r] |= (1u << k);
            }
            upn[k] = __shfl_up_sync(0xffffffffu, lastp, 1);     // consumed at frame k of the next group
            lastp = nv[R - 1];
            if (lane == 31) sts32(bout_addr + (((yl + kk + 1) & (kRing - 1)) << 2), nv[R - 1]);
#pragma unroll
            for (int r = 0; r < R; ++r) old[r] = nv[r];
        }
        const int yg = yl + 4 * g;
#pragma unroll
        for (int r = 0; r < R; ++r) wbits[r] = __funnelshift_r(wbits[r], hb[r], 4);
        if ((yg & 31) == 28 && (unsigned)(yg - y_lo) < span) {
            uint32_t* brow = bits_row + (int64_t)(yg >> 5) * TXS;
#pragma unroll
            for (int r = 0; r < R; ++r) brow[r] = wbits[r];
        }
    }
}

// ------------------------------------------------------------------ the kernel
template <int R, int TF, bool SKEW>
__global__ void __launch_bounds__(kMaxWarps * 32) mas_kernel(const MasParams p)
{
    constexpr int RW = 32 * R;
    constexpr int UNIT = TF < 16 ? TF : 16;
    constexpr int LANE_STRIDE = R * TF * 4 + kLanePad;

    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int NW = blockDim.x >> 5;
    const int NS = p.ns;
    const int TXS = NW * RW;
    const bool bits_smem = (p.bits_ws == nullptr);
    const SmemLayout L = make_layout(NW, NS, R, TF, bits_smem, p.nblk, p.durations != nullptr);

    const uint32_t bar0 = smem_u32(smem + L.off_bar) + w * NS * 8;
    const uint32_t prog_a = smem_u32(smem + L.off_prog);            // tail progress of warp w at +4*w, head progress at +4*(NW+w)
    const uint32_t head_a = prog_a + 4 * NW;
    double* msum = reinterpret_cast<double*>(smem + L.off_misc + 64);   // [w][2] partial mask sums
    int* misc = reinterpret_cast<int*>(smem + L.off_misc);          // [0]=item [1]=t_x [2]=t_y
    const uint32_t bnd_a = smem_u32(smem + L.off_bnd);
    const uint32_t zero_a = smem_u32(smem + L.off_zero);
    const uint32_t ring_a = smem_u32(smem + L.off_ring) + w * NS * L.stage_bytes;
    uint32_t* bits = bits_smem ? reinterpret_cast<uint32_t*>(smem + L.off_bits)
                               : p.bits_ws + (int64_t)blockIdx.x * p.bits_slot_words;
    int* durS = reinterpret_cast<int*>(smem + L.off_dur);

    // ---- one-time setup
    if (lane == 0)
        for (int s = 0; s < NS; ++s) mbar_init(bar0 + 8 * s, 32);      // one arrival per lane
    for (int i = tid; i < kZeroChunk / 16; i += blockDim.x)
        reinterpret_cast<int4*>(smem + L.off_zero)[i] = make_int4(0, 0, 0, 0);
    fence_mbar_init();
    fence_proxy_async_smem();
    __syncthreads();

    uint32_t cstage = 0, cphase = 0, pstage = 0;   // per-warp ring cursors, persist across items
    const int64_t item_elems = (int64_t)p.Tx * p.Ty;

    int item = blockIdx.x;
    const bool dbg_on = (p.dbg != nullptr);
    long long* dbg = dbg_on ? p.dbg + (int64_t)blockIdx.x * (kMaxWarps + 2) * 2 : nullptr;
    bool first_item = true;
    while (item < p.B) {
        if (dbg_on && first_item && lane == 0) dbg[w * 2] = clock64();
        // ---- lengths
        if (p.mask != nullptr) {
            double ax = 0.0, ay = 0.0;
            mask_partial_any(p.mask, p.mask_dtype, (int64_t)item * p.msb, p.msx, p.msy, p.Tx, p.Ty, tid, blockDim.x, ax, ay);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ax += __shfl_xor_sync(0xffffffffu, ax, o);
                ay += __shfl_xor_sync(0xffffffffu, ay, o);
            }
            if (lane == 0) { msum[2 * w] = ax; msum[2 * w + 1] = ay; }
        } else if (tid == 0) {
            misc[1] = p.t_xs[item];
            misc[2] = p.t_ys[item];
        }
        if (tid < NW) {
            st_flag(prog_a + 4 * tid, SKEW ? -(1 << 28) : tid * RW);
            st_flag(head_a + 4 * tid, tid * RW);
        }
        if (p.durations != nullptr)
            for (int i = tid; i < TXS; i += blockDim.x) durS[i] = 0;
        __syncthreads();
        int t_x, t_y;
        if (p.mask != nullptr) {          // sum the per-warp partials; truncation like astype(np.int32) (__init__.py:18-19)
            double ax = 0.0, ay = 0.0;
            for (int q = 0; q < NW; ++q) { ax += msum[2 * q]; ay += msum[2 * q + 1]; }
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
        else if (t_x > t_y || t_x > p.Tx || t_y > p.Ty) {              // reference reads out of bounds here
            valid = false;
            if (tid == 0) atomicOr(&p.ws->status, 1);
        }
        if (!valid) { t_x = 0; t_y = 0; }

        // ---- zero-fill bookkeeping for this item (dense output, fused memset)
        const bool zf = p.zero_fill && p.paths != nullptr;
        unsigned char* pbase = reinterpret_cast<unsigned char*>(p.paths) + item * item_elems * p.esize;
        unsigned char* zA = nullptr;
        int64_t zbytes = 0;
        if (zf) {
            unsigned char* pend = pbase + item_elems * p.esize;
            zA = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(pbase) + 15) & ~uintptr_t(15));
            unsigned char* zE = reinterpret_cast<unsigned char*>(reinterpret_cast<uintptr_t>(pend) & ~uintptr_t(15));
            if (zA >= zE) {   // tiny or fully misaligned item: plain byte stores
                for (int64_t i = tid; i < item_elems * p.esize; i += blockDim.x) pbase[i] = 0;
                zbytes = 0;
            } else {
                const int head = (int)(zA - pbase), tail = (int)(pend - zE);
                if (tid < head) pbase[tid] = 0;
                if (tid < tail) zE[tid] = 0;
                zbytes = zE - zA;
            }
        }
        const int nact = valid ? (t_x + RW - 1) / RW : NW;     // warps that take part in the zero fill
        const int64_t nchunks = (zbytes + kZeroChunk - 1) / kZeroChunk;
        int64_t zc = w;                                        // next chunk this warp issues
        auto issue_zero = [&](int n) {
            if (lane == 0) {
                for (int q = 0; q < n && zc < nchunks; ++q, zc += nact) {
                    const int64_t off = zc * kZeroChunk;
                    const int64_t left = zbytes - off;
                    bulk_s2g(zA + off, zero_a, (uint32_t)(left < kZeroChunk ? left : kZeroChunk));
                }
            }
        };

        const int x0 = w * RW;
        if (valid && x0 < t_x) {
            // ================= forward pass for this warp's rows [x0, x1) =================
            const int x1 = (x0 + RW < t_x) ? x0 + RW : t_x;
            const int nrows = x1 - x0;
            const int y_start = x0;                                 // first frame where any of our rows is on/below the diagonal
            const int y_last = t_y - t_x + x1 - 1;                  // last frame where our last row is inside the band (core.pyx:18)
            // lock-step: whole units up to y_last.  skewed: whole 32-frame words, plus 124 frames so lane 31 finishes too.
            const int span = SKEW ? ((y_last + 1 - y_start + 31) & ~31) : 0;
            const int y_end = SKEW ? y_start + span + 128 : (y_last + UNIT) / UNIT * UNIT;
            const int t_s = y_start / TF;
            const int y_cap = SKEW ? y_end : (y_end < p.Ty ? y_end : p.Ty);
            const int t_e = (y_cap + TF - 1) / TF;                  // tiles [t_s, t_e)
            const bool has_consumer = (x1 < t_x);
            const float* vrow = p.values + item * item_elems + (int64_t)x0 * p.Ty;
            const int band_hi0 = t_y - t_x + x0;                    // last live frame of row i is band_hi0 + i

            // One tile = our rows x TF frames (per-lane skewed by 4 frames in SKEW mode).  Loader lane = (16-byte chunk
            // ck of a row, row group q); it walks the owner lanes li = q, q+RPI, ... and their R rows, so one warp-wide
            // LDGSTS.128 moves RPI whole row segments: full 32..128-byte global segments, conflict-free shared writes.
            constexpr int CPR = TF / 4;            // 16-byte chunks per row
            constexpr int RPI = 32 / CPR;          // row segments per warp instruction
            const int ck = lane % CPR, q0 = lane / CPR;
            auto issue_tile = [&](int t) {
                const int f0 = t * TF + ck * 4;
                const uint32_t bar = bar0 + 8 * pstage;
                const uint32_t st = ring_a + pstage * L.stage_bytes + ck * 16;
                for (int li = q0; li * R < nrows; li += RPI) {
                    const int f = SKEW ? f0 - 4 * li : f0;                       // frame of this chunk for owner lane li
                    const int lo = SKEW ? ((x0 + li * R) & ~3) : 0;              // chunks wholly above the diagonal are never read
                    const uint32_t d = st + li * LANE_STRIDE;
                    const float* src = vrow + (int64_t)(li * R) * p.Ty + f;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int i = li * R + r;
                        const bool in = SKEW ? (f >= lo && f <= band_hi0 + i) : true;
                        if (i < nrows && in && f + 4 <= p.Ty) cp_async16(d + r * (TF * 4), src + (int64_t)r * p.Ty);
                    }
                }
                cp_async_arrive(bar);
                if (++pstage == (uint32_t)NS) pstage = 0;
            };
            auto load_tile_sync = [&](int t) {   // unaligned inputs: plain 4-byte loads into stage 0
                const int f0 = t * TF;
                float* st = reinterpret_cast<float*>(smem + L.off_ring + (size_t)w * NS * L.stage_bytes);
                __syncwarp();
                for (int idx = lane; idx < nrows * TF; idx += 32) {
                    const int i = idx / TF, fl = idx - i * TF, li = i / R;
                    const int f = SKEW ? f0 + fl - 4 * li : f0 + fl;
                    if (f >= 0 && f < p.Ty) st[(i * (TF * 4) + li * kLanePad) / 4 + fl] = vrow[(int64_t)i * p.Ty + f];
                }
                __syncwarp();
            };

            const int my_tiles = t_e - t_s;
            int64_t my_chunks = (nchunks > w) ? (nchunks - w + nact - 1) / nact : 0;
            const int zq = (int)((my_chunks + my_tiles - 1) / (my_tiles > 0 ? my_tiles : 1));

            int t_next = t_s;
            if (p.aligned) {
                for (int s = 0; s < NS && t_next < t_e; ++s, ++t_next) issue_tile(t_next);
            }

            float old[R];
            uint32_t wbits[R];
#pragma unroll
            for (int r = 0; r < R; ++r) { old[r] = p.neg; wbits[r] = 0u; }
            float up = p.neg;                                       // lock-step form
            float upn[4] = { p.neg, p.neg, p.neg, p.neg };          // skewed form
            float lastp = p.neg;
            const int xl0 = x0 + lane * R;
            const int lag = SKEW ? 4 * lane : 0;
            const uint32_t bin_addr = bnd_a + (w > 0 ? (w - 1) : 0) * kRing * 4;
            const uint32_t bout_addr = bnd_a + w * kRing * 4;
            const int diag_end = SKEW ? x0 + RW + 124 : x1;
            int seen_cons = 0;

            for (int y = y_start; y < y_end; y += UNIT) {           // y = frame of lane 0
                const int fin = y & (TF - 1);
                if (fin == 0) {
                    if (p.aligned) mbar_wait(bar0 + 8 * cstage, cphase);
                    else load_tile_sync(y / TF);
                }
                if (w > 0) {
                    while (ld_flag(prog_a + 4 * (w - 1)) < y + UNIT) { }
                }
                if (has_consumer) {
                    const int need = y - (SKEW ? 124 : 0) + UNIT - (kRing - 1);   // our lane 31 is about to overwrite these ring slots
                    while (seen_cons < need) seen_cons = ld_flag(head_a + 4 * (w + 1));
                }
                const uint32_t tile_addr = (p.aligned ? ring_a + cstage * L.stage_bytes : ring_a) + lane * LANE_STRIDE + fin * 4;
                if constexpr (SKEW) {
                    const int yl = y - lag;
                    if (y < diag_end)
                        mas_unit_skew<R, TF, UNIT, true>(old, upn, lastp, wbits, tile_addr, bin_addr, bout_addr, y, yl, w, lane, p.neg,
                                                         xl0 - yl, bits + xl0, TXS, y_start, (unsigned)span);
                    else
                        mas_unit_skew<R, TF, UNIT, false>(old, upn, lastp, wbits, tile_addr, bin_addr, bout_addr, y, yl, w, lane, p.neg,
                                                          0, bits + xl0, TXS, y_start, (unsigned)span);
                    if (lane == 31) st_flag(prog_a + 4 * w, y + UNIT - 124);
                    if (lane == 0) st_flag(head_a + 4 * w, y + UNIT);
                } else {
                    uint32_t hb[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) hb[r] = 0u;
                    if (y < diag_end)
                        mas_unit<R, TF, UNIT, true>(old, up, hb, tile_addr, bin_addr, bout_addr, y, w, lane, p.neg, xl0 - y);
                    else
                        mas_unit<R, TF, UNIT, false>(old, up, hb, tile_addr, bin_addr, bout_addr, y, w, lane, p.neg, 0);
                    const int pos = y & 31;
#pragma unroll
                    for (int r = 0; r < R; ++r) wbits[r] |= hb[r] << pos;
                    if (lane == 31) { st_flag(prog_a + 4 * w, y + UNIT); st_flag(head_a + 4 * w, y + UNIT); }
                }

                const int yn = y + UNIT;
                if ((yn & (TF - 1)) == 0 || yn >= y_end) {          // tile consumed
                    if (p.aligned) {
                        __syncwarp();
                        if (t_next < t_e) { issue_tile(t_next); }
                        ++t_next;
                        if (++cstage == (uint32_t)NS) { cstage = 0; cphase ^= 1u; }
                    }
                    if (zf) issue_zero(zq);
                }
                if (!SKEW && ((yn & 31) == 0 || yn >= y_end)) {     // direction word complete (lock-step form)
                    uint32_t* brow = bits + (int64_t)(y >> 5) * TXS + xl0;
#pragma unroll
                    for (int r = 0; r < R; ++r) { brow[r] = wbits[r]; wbits[r] = 0u; }
                }
            }
            if (lane == 31) st_flag(prog_a + 4 * w, kProgDone);
            if (dbg_on && first_item && lane == 0) dbg[w * 2 + 1] = clock64();
        }
        if (zf && w < nact) {
            issue_zero(0x7fffffff);
            if (lane == 0) { bulk_commit(); bulk_wait_all(); fence_proxy_async_global(); }
        }
        __syncthreads();

        // ================= backtrack (warp 0) =================
        if (dbg_on && first_item && tid == 0) dbg[kMaxWarps * 2] = clock64();
        if (w == 0) {
            int tok0 = t_x - 1;
            for (int blk = (t_y - 1) >> 5; blk >= 0; --blk) {
                const int yb = blk << 5;
                const int nvalid = (t_y - yb < 32) ? t_y - yb : 32;
                const int row = tok0 - lane;
                uint32_t wd = (row > 0) ? bits[(int64_t)blk * TXS + row] : 0u;      // row 0 can never step down (core.pyx:34 index != 0)
                const int dg = row - yb;                                            // diagonal cell of this row: forced step (index == y)
                if (row > 0 && dg >= 0 && dg < 32) wd |= (1u << dg);
                if (nvalid < 32) wd &= (1u << nvalid) - 1u;
                // transpose the 32x32 bit block with 32 ballots FIRST (independent, pipelined), then walk: the walk's
                // dependent chain is one AND and one ADD per frame on a one-hot position.
                uint32_t m[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) m[k] = __ballot_sync(0xffffffffu, (wd & (1u << k)) != 0u);
                uint32_t pos = 1u, mine = 1u;
#pragma unroll
                for (int k = 31; k >= 0; --k) {
                    if (lane == k) mine = pos;
                    pos = pos + (pos & m[k]);
                }
                if (lane < nvalid) {
                    const int tok = tok0 - (__ffs(mine) - 1);
                    const int yy = yb + lane;
                    if (p.paths != nullptr) store_one(p.paths, item * item_elems + (int64_t)tok * p.Ty + yy, p.esize, p.one);
                    if (p.frame_tok != nullptr) p.frame_tok[(int64_t)item * p.Ty + yy] = tok;
                    if (p.durations != nullptr) atomicAdd(&durS[tok], 1);
                }
                tok0 -= (pos != 0u) ? (__ffs(pos) - 1) : 32;
            }
        } else if (p.frame_tok != nullptr) {
            for (int yy = t_y + (tid - 32); yy < p.Ty; yy += blockDim.x - 32) p.frame_tok[(int64_t)item * p.Ty + yy] = -1;
        }
        if (p.frame_tok != nullptr && NW == 1)
            for (int yy = t_y + lane; yy < p.Ty; yy += 32) p.frame_tok[(int64_t)item * p.Ty + yy] = -1;
        __syncthreads();
        if (p.durations != nullptr)
            for (int i = tid; i < p.Tx; i += blockDim.x) p.durations[(int64_t)item * p.Tx + i] = (i < t_x) ? durS[i] : 0;

        if (dbg_on && first_item && tid == 0) dbg[kMaxWarps * 2 + 1] = clock64();
        first_item = false;
        // ---- next item
        if (p.B <= (int)gridDim.x) break;
        if (tid == 0) misc[0] = atomicAdd(&p.ws->counter, 1) + (int)gridDim.x;
        __syncthreads();
        item = misc[0];
    }

    if (p.B > (int)gridDim.x && tid == 0) {
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
