// Micro-benchmark (round 2): ALU-pipe relief for the skewed per-frame body of a lone warp, R rows per lane.
//   bit0  select by predicated add:  res = stay + v;  @take res = move + v      (FSEL leaves the ALU pipe, the add is on the FMA pipe)
//   bit1  direction bits accumulated in a float: @take hbf += 2^k               (predicated FADD instead of predicated integer add)
//   bit3  values from a per-lane skewed tile: one aligned LDS.128 per row per 4 frames, no current/previous select
//   bit2  lane-to-lane value through shared memory (STS + LDS two frames ahead) instead of SHFL + lane-0 select + staged STS.128
// All variants must print the same checksum.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ float lds32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ float lds32m(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }

template <int R, int FLAGS, int POLL = 0, int SLEEP = 0, int NCW = 1>
__global__ void body(float* out, long long* cyc, int nframes)
{
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31;
    const uint32_t pbar = (uint32_t)__cvta_generic_to_shared(sm) + 88000;
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pbar) : "memory");
    __syncthreads();
    if (threadIdx.x >= 32 * NCW) {
        if ((threadIdx.x >> 5) == POLL) {      // the warp on warp 0's scheduler (4), or on another one (1)
            uint32_t ok = 0;
            while (!ok) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(pbar), "r"(0) : "memory");
                if (SLEEP) __nanosleep(SLEEP);
            }
        }
        return;
    }
    for (int i = lane; i < 20000; i += 32) sm[i] = 0.001f * (float)((i * 7919) % 1000) - 0.5f;
    __syncthreads();
    float old[R]; uint32_t hb[R], hbp[R]; float hf[R][2];
    for (int r = 0; r < R; ++r) { old[r] = -1e9f; hb[r] = 0; hbp[r] = 0; }
    float u1 = -1e9f, u2 = -1e9f, u3 = -1e9f, u4 = -1e9f, u5 = -1e9f;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + (NCW > 1 ? (threadIdx.x >> 5) * 1024 : 0);
    const uint32_t bnd = base + 60000;          // incoming boundary ring (128 slots)
    const uint32_t bout = base + 61440;         // outgoing boundary ring
    const uint32_t xbuf = base + 63488;         // lane exchange: 32 lanes x 33 words
    const bool lane0 = lane == 0, lane31 = lane == 31;
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int y = 0; y < nframes; y += 32) {
        const uint32_t curA = base + ((y >> 5) & 1) * 8192 + lane * (R * 128) - 4 * lane;
        const uint32_t prevA = base + (((y >> 5) + 1) & 1) * 8192 + lane * (R * 128) + 4 * (32 - lane);
        float4 bin[8];
        float o4[4];
        // exchange addresses: we write slot k of our own row (lane 31: the outgoing ring), read slot k of the lane below (lane 0: the incoming ring)
        const uint32_t wr = lane31 ? bout + ((y & 127) << 2) : xbuf + lane * 132;
        const uint32_t rd = lane0 ? bnd + (((y + 32) & 127) << 2) : xbuf + (lane - 1) * 132;
        if (!(FLAGS & 4))
            for (int g = 0; g < 8; ++g) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bin[g].x), "=f"(bin[g].y), "=f"(bin[g].z), "=f"(bin[g].w) : "r"(bnd + ((y & 127) * 4) + 16 * g));
        float vq[2][R];
        float4 vg[2][R];                      // bit3: group g in vg[g & 1]
        const uint32_t skA = base + ((y >> 5) & 1) * 8192 + lane * 128;      // row r of this lane at + r * 4096; chunk g at ((g ^ (lane & 7)) << 4)
        if (FLAGS & 8) {
            for (int r = 0; r < R; ++r) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(vg[0][r].x), "=f"(vg[0][r].y), "=f"(vg[0][r].z), "=f"(vg[0][r].w) : "r"(skA + r * 4096 + (((0 ^ lane) & 7) << 4)));
        } else
        for (int q = 0; q < 2; ++q) for (int r = 0; r < R; ++r) vq[q][r] = lds32((lane <= q ? curA : prevA) + q * 4 + r * 128);
        for (int r = 0; r < R; ++r) { hf[r][0] = 8388608.f; hf[r][1] = 8388608.f; hb[r] = 0; }
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            float upv;
            if (FLAGS & 4) upv = u1;
            else {
                const float4 b4 = bin[k >> 2];
                const float bk = (k & 3) == 0 ? b4.x : (k & 3) == 1 ? b4.y : (k & 3) == 2 ? b4.z : b4.w;
                upv = lane0 ? bk : u1;
            }
            float nv[R];
#pragma unroll
            for (int r = R - 1; r >= 0; --r) {
                const float stay = old[r];
                const float move = (r == 0) ? upv : old[r - 1];
                const float4 v4 = vg[(k >> 2) & 1][r];
                const float v = (FLAGS & 8) ? ((k & 3) == 0 ? v4.x : (k & 3) == 1 ? v4.y : (k & 3) == 2 ? v4.z : v4.w) : vq[k & 1][r];
                if (FLAGS & 16) {       // speculative: both sums formed during the compare, select last
                    float res; uint32_t h = hb[r];
                    asm("{\n\t.reg .pred p;\n\t.reg .f32 a, b;\n\tsetp.gt.f32 p, %3, %2;\n\tadd.f32 a, %2, %4;\n\tadd.f32 b, %3, %4;\n\tselp.f32 %0, b, a, p;\n\t@p or.b32 %1, %1, %5;\n\t}"
                        : "=f"(res), "+r"(h) : "f"(stay), "f"(move), "f"(v), "r"(1u << k));
                    nv[r] = res; hb[r] = h;
                } else if (FLAGS & 32) {   // not the reference's NaN semantics: lower bound only
                    nv[r] = fmaxf(stay, move) + v;
                    if (move > stay) hb[r] |= 1u << k;
                } else if ((FLAGS & 3) == 3) {
                    float res;
                    asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %3, %2;\n\tadd.f32 %0, %2, %4;\n\t@p add.f32 %0, %3, %4;\n\t@p add.f32 %1, %1, %5;\n\t}"
                        : "=&f"(res), "+f"(hf[r][k >> 4]) : "f"(stay), "f"(move), "f"(v), "f"((float)(1u << (k & 15))));
                    nv[r] = res;
                } else if (FLAGS & 1) {
                    float res; uint32_t h = hb[r];
                    asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %3, %2;\n\tadd.f32 %0, %2, %4;\n\t@p add.f32 %0, %3, %4;\n\t@p or.b32 %1, %1, %5;\n\t}"
                        : "=&f"(res), "+r"(h) : "f"(stay), "f"(move), "f"(v), "r"(1u << k));
                    nv[r] = res; hb[r] = h;
                } else if (FLAGS & 2) {
                    float res;
                    asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %3, %2;\n\tselp.f32 %0, %3, %2, p;\n\tadd.f32 %0, %0, %4;\n\t@p add.f32 %1, %1, %5;\n\t}"
                        : "=&f"(res), "+f"(hf[r][k >> 4]) : "f"(stay), "f"(move), "f"(v), "f"((float)(1u << (k & 15))));
                    nv[r] = res;
                } else {
                    const bool take = move > stay;
                    nv[r] = (take ? move : stay) + v;
                    if (!(FLAGS & 128)) { if (take) hb[r] |= 1u << k; }
                }
            }
            if (FLAGS & 4) {
                if (!(FLAGS & 256)) sts32(wr + 4 * k, nv[R - 1]);
                u1 = u2;
                u2 = (FLAGS & 512) ? nv[R - 1] * 0.5f : lds32m(((FLAGS & 64) ? rd + 8192 : rd) + 4 * k);
            } else {
                if (FLAGS & 2048) { u1 = u2; u2 = u3; u3 = u4; u4 = u5; u5 = __shfl_up_sync(0xffffffffu, nv[R - 1], 1); }
                else { u1 = u2; u2 = __shfl_up_sync(0xffffffffu, nv[R - 1], 1); }
                o4[k & 3] = nv[R - 1];
                if ((k & 3) == 3 && lane31) asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(bout + ((y & 127) << 2) + 4 * (k - 3)), "f"(o4[0]), "f"(o4[1]), "f"(o4[2]), "f"(o4[3]) : "memory");
            }
#pragma unroll
            for (int r = 0; r < R; ++r) old[r] = nv[r];
            if (FLAGS & 8) {
                if ((k & 3) == 0 && k + 4 < 32 && !(FLAGS & 1024)) {
                    const int g = (k >> 2) + 1;
                    for (int r = 0; r < R; ++r) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(vg[g & 1][r].x), "=f"(vg[g & 1][r].y), "=f"(vg[g & 1][r].z), "=f"(vg[g & 1][r].w) : "r"(skA + r * 4096 + (((g ^ lane) & 7) << 4)));
                }
            } else if (k + 2 < 32) {
                const uint32_t a = (lane <= k + 2) ? curA : prevA;
                for (int r = 0; r < R; ++r) vq[k & 1][r] = lds32(a + (k + 2) * 4 + r * 128);
            }
        }
        for (int r = 0; r < R; ++r) {
            if (FLAGS & 2) hb[r] = (__float_as_uint(hf[r][0]) & 0xffffu) | (__float_as_uint(hf[r][1]) << 16);
            acc ^= __funnelshift_r(hbp[r], hb[r], lane); hbp[r] = hb[r];
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int r = 0; r < R; ++r) s += old[r];
    out[lane] = s + (float)(acc & 0xffff);
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pbar) : "memory");
}
template <int R, int FLAGS, int POLL = 0, int SLEEP = 0, int NCW = 1> void run()
{
    const int n = 4096;
    float* out; long long* cyc;
    cudaMalloc(&out, 128); cudaMalloc(&cyc, 8);
    cudaFuncSetAttribute(body<R, FLAGS, POLL, SLEEP, NCW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 90000);
    for (int it = 0; it < 2; ++it) body<R, FLAGS, POLL, SLEEP, NCW><<<1, POLL ? 160 : 32 * NCW, 90000>>>(out, cyc, n);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    float ho[32]; cudaMemcpy(ho, out, 128, cudaMemcpyDeviceToHost);
    double cs = 0; for (int i = 0; i < 32; ++i) cs += ho[i] * (i + 1);
    printf("warps %d poll-warp %d sleep %d  R=%d flags=%d [%s%s%s%s%s%s%s%s]  %7.2f cycles/frame  checksum %.6e (%s)\n", NCW, POLL, SLEEP, R, FLAGS, FLAGS & 1 ? "predadd " : "", FLAGS & 2 ? "floatbits " : "", FLAGS & 4 ? "smemxchg " : "", FLAGS & 8 ? "skewtile " : "", FLAGS & 16 ? "specadd " : "", FLAGS & 32 ? "fmax " : "", FLAGS & 64 ? "xchg-nodep " : "", FLAGS & 2048 ? "lag4 " : "",
           (double)h / n, cs, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    run<2, 2056>(); run<2, 2056, 0, 0, 2>(); run<2, 2056, 0, 0, 4>(); run<4, 2056>(); run<4, 2056, 0, 0, 2>(); run<4, 2056, 0, 0, 4>(); run<2, 8>(); run<2, 8, 0, 0, 4>();
    return 0;
}
