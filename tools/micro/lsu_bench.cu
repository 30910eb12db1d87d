// Micro-benchmark: what does one shared-memory / shuffle instruction cost a LONE warp whose arithmetic is a short dependent chain?
// Per iteration: a 3-op dependent fp chain (as one MAS cell) x NC independent chains, plus a configurable mix of LSU-side ops whose
// results are consumed DIST iterations later.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ float lds32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v)); }

// NL plain loads, NS stores, NSH shuffles per iteration; NC arithmetic chains; DIST = iterations between a load and its use
template <int NL, int NS, int NSH, int NC, int DIST>
__global__ void k(float* out, long long* cyc, int iters)
{
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x;
    for (int i = lane; i < 8192; i += 32) sm[i] = 0.001f * (float)((i * 7919) % 1000) - 0.5f;
    __syncwarp();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + lane * 4;
    float c[NC > 0 ? NC : 1];
    for (int j = 0; j < NC; ++j) c[j] = 0.1f * j;
    float pend[DIST][NL + NSH > 0 ? NL + NSH : 1];
    for (int d = 0; d < DIST; ++d) for (int j = 0; j < NL + NSH; ++j) pend[d][j] = 0.f;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            // consume what was loaded DIST iterations ago
            float inj = 0.f;
            for (int j = 0; j < NL + NSH; ++j) inj += pend[0][j];
            for (int d = 0; d + 1 < DIST; ++d) for (int j = 0; j < NL + NSH; ++j) pend[d][j] = pend[d + 1][j];
            for (int j = 0; j < NC; ++j) {
                const float a = c[j], b = (j == 0) ? inj : c[j - 1];
                c[j] = (b > a ? b : a) + 0.25f;          // FSETP -> FSEL -> FADD
            }
            for (int j = 0; j < NS; ++j) sts32(base + 16384 + 128 * (u + 8 * j), c[NC - 1]);
            for (int j = 0; j < NL; ++j) pend[DIST - 1][j] = lds32(base + 128 * (u + 8 * j));
            for (int j = 0; j < NSH; ++j) pend[DIST - 1][NL + j] = __shfl_up_sync(0xffffffffu, c[NC - 1], 1 + j);
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int j = 0; j < NC; ++j) s += c[j];
    out[lane] = s;
    if (lane == 0) cyc[0] = t1 - t0;
}
template <int NL, int NS, int NSH, int NC, int DIST> void run()
{
    const int n = 8192;
    float* out; long long* cyc;
    cudaMalloc(&out, 128); cudaMalloc(&cyc, 8);
    cudaFuncSetAttribute(k<NL, NS, NSH, NC, DIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int it = 0; it < 2; ++it) k<NL, NS, NSH, NC, DIST><<<1, 32, 65536>>>(out, cyc, n);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("loads %d stores %d shuffles %d chains %d dist %d : %6.2f cycles/iter (%s)\n", NL, NS, NSH, NC, DIST, (double)h / n, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    run<0, 0, 0, 2, 2>();
    run<1, 0, 0, 2, 2>(); run<2, 0, 0, 2, 2>(); run<3, 0, 0, 2, 2>(); run<4, 0, 0, 2, 2>();
    run<1, 0, 0, 2, 3>(); run<2, 0, 0, 2, 3>(); run<2, 0, 0, 2, 4>(); run<4, 0, 0, 2, 4>();
    run<0, 1, 0, 2, 2>(); run<0, 2, 0, 2, 2>(); run<1, 1, 0, 2, 2>(); run<1, 1, 0, 2, 3>();
    run<0, 0, 1, 2, 2>(); run<0, 0, 2, 2, 2>(); run<1, 0, 1, 2, 2>(); run<2, 0, 1, 2, 2>(); run<0, 0, 1, 2, 3>();
    run<1, 0, 0, 1, 2>(); run<0, 0, 1, 1, 2>(); run<1, 0, 0, 4, 2>(); run<0, 0, 1, 4, 2>(); run<2, 0, 1, 4, 2>();
    return 0;
}
