// Micro-benchmark: cycles per frame of the MAS recurrence inner loop for one warp alone on an SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_bench chain_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

// MODE: 0 lockstep, 1 skewed (use 4 frames later), 2 skewed no bits, 3 skewed fmax (no select), 4 skewed + plain C++ smem loads,
//       5 skewed, loads for next group prefetched, 6 lockstep fmax no bits
template <int R, int MODE>
__global__ void chain(float* out, long long* cyc, int nframes)
{
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x;
    for (int i = lane; i < 32 * R * 32 + 64; i += 32) sm[i] = 0.001f * (float)((i * 7919) % 1000) - 0.5f;
    __syncwarp();
    float old[R], upn[4] = {-1e9f, -1e9f, -1e9f, -1e9f}, lastp = -1e9f, up = -1e9f;
    uint32_t wb[R];
    for (int r = 0; r < R; ++r) { old[r] = -1e9f; wb[r] = 0; }
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + lane * (R * 128 + 16);
    const float4* basep = reinterpret_cast<const float4*>(reinterpret_cast<const char*>(sm) + lane * (R * 128 + 16));
    const bool lane0 = lane == 0;
    long long t0 = clock64();
    float4 vn[R];
    if (MODE == 5) for (int r = 0; r < R; ++r) vn[r] = lds128(base + r * 128);
    for (int y = 0; y < nframes; y += 4) {
        const int g = (y >> 2) & 7;
        float4 v[R];
        if (MODE == 4) { for (int r = 0; r < R; ++r) v[r] = basep[r * 8 + g]; }
        else if (MODE == 5) { for (int r = 0; r < R; ++r) v[r] = vn[r]; const int g2 = (g + 1) & 7; for (int r = 0; r < R; ++r) vn[r] = lds128(base + r * 128 + g2 * 16); }
        else { for (int r = 0; r < R; ++r) v[r] = lds128(base + r * 128 + g * 16); }
        uint32_t hb[R];
        for (int r = 0; r < R; ++r) hb[r] = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float upv = lane0 ? 0.f : ((MODE == 0 || MODE == 6) ? up : upn[k]);
            float nv[R];
#pragma unroll
            for (int r = R - 1; r >= 0; --r) {
                const float stay = old[r];
                const float move = (r == 0) ? upv : old[r - 1];
                const float vr = (k == 0) ? v[r].x : (k == 1) ? v[r].y : (k == 2) ? v[r].z : v[r].w;
                if (MODE == 3 || MODE == 6) {
                    nv[r] = fmaxf(move, stay) + vr;
                    if (MODE == 3 && move > stay) hb[r] |= 1u << k;
                } else {
                    const bool take = move > stay;
                    nv[r] = (take ? move : stay) + vr;
                    if (MODE != 2 && take) hb[r] |= 1u << k;
                }
            }
            if (MODE == 0 || MODE == 6) up = __shfl_up_sync(0xffffffffu, nv[R - 1], 1);
            else { upn[k] = __shfl_up_sync(0xffffffffu, lastp, 1); lastp = nv[R - 1]; }
#pragma unroll
            for (int r = 0; r < R; ++r) old[r] = nv[r];
        }
        for (int r = 0; r < R; ++r) wb[r] = __funnelshift_r(wb[r], hb[r], 4);
    }
    long long t1 = clock64();
    float s = 0; uint32_t b = 0;
    for (int r = 0; r < R; ++r) { s += old[r]; b ^= wb[r]; }
    out[blockIdx.x * 32 + lane] = s + (float)b;
    if (lane == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int R, int MODE>
void run(const char* name, int nframes)
{
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 32 * 4); cudaMalloc(&cyc, 148 * 8);
    size_t smem = (32 * R * 32 + 64) * 4 + 32 * 16;
    cudaFuncSetAttribute(chain<R, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int it = 0; it < 2; ++it) chain<R, MODE><<<8, 32, smem>>>(out, cyc, nframes);
    cudaDeviceSynchronize();
    long long h[8];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-44s R=%d  %7.2f cycles/frame  (%s)\n", name, R, (double)h[0] / nframes, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    const int n = 8192;
    run<2, 0>("lockstep select+bits", n);
    run<2, 6>("lockstep fmax no bits", n);
    run<2, 1>("skewed select+bits", n);
    run<2, 2>("skewed select, no bits", n);
    run<2, 3>("skewed fmax + bits", n);
    run<2, 4>("skewed select+bits, C++ smem loads", n);
    run<2, 5>("skewed select+bits, prefetched loads", n);
    run<1, 0>("lockstep select+bits", n);
    run<1, 1>("skewed select+bits", n);
    run<1, 5>("skewed select+bits, prefetched loads", n);
    run<4, 0>("lockstep select+bits", n);
    run<4, 1>("skewed select+bits", n);
    run<4, 5>("skewed select+bits, prefetched loads", n);
    run<8, 1>("skewed select+bits", n);
    run<8, 5>("skewed select+bits, prefetched loads", n);
    return 0;
}
