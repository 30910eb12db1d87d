// Micro-benchmark: which part of the skewed per-frame body limits a lone warp?  R = 2, lag 1, variants switch pieces off.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ float lds32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
// FLAGS bit0: fmaxf instead of select; bit1: no direction bits; bit2: no tile loads (register values); bit3: no shuffle; bit4: no cur/prev select
template <int R, int FLAGS>
__global__ void body(float* out, long long* cyc, int nframes)
{
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x;
    for (int i = lane; i < 16384; i += 32) sm[i] = 0.001f * (float)((i * 7919) % 1000) - 0.5f;
    __syncwarp();
    float old[R]; uint32_t hb[R], hbp[R];
    for (int r = 0; r < R; ++r) { old[r] = -1e9f; hb[r] = 0; hbp[r] = 0; }
    float u1 = -1e9f, u2 = -1e9f;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    const uint32_t bnd = base + 60000;
    const bool lane0 = lane == 0;
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int y = 0; y < nframes; y += 32) {
        const uint32_t curA = base + ((y >> 5) & 1) * 8192 + lane * (R * 128) - 4 * lane;
        const uint32_t prevA = base + (((y >> 5) + 1) & 1) * 8192 + lane * (R * 128) + 4 * (32 - lane);
        float4 bin[8];
        for (int g = 0; g < 8; ++g) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bin[g].x), "=f"(bin[g].y), "=f"(bin[g].z), "=f"(bin[g].w) : "r"(bnd + ((y & 127) * 4) + 16 * g));
        float vq[2][R];
        for (int q = 0; q < 2; ++q) for (int r = 0; r < R; ++r) vq[q][r] = (FLAGS & 4) ? 0.25f * (q + r) : lds32(((lane <= q || (FLAGS & 16)) ? curA : prevA) + q * 4 + r * 128);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const float4 b4 = bin[k >> 2];
            const float bk = (k & 3) == 0 ? b4.x : (k & 3) == 1 ? b4.y : (k & 3) == 2 ? b4.z : b4.w;
            const float upv = lane0 ? bk : u1;
            float nv[R];
#pragma unroll
            for (int r = R - 1; r >= 0; --r) {
                const float stay = old[r];
                const float move = (r == 0) ? upv : old[r - 1];
                const bool take = move > stay;
                nv[r] = ((FLAGS & 1) ? fmaxf(stay, move) : (take ? move : stay)) + vq[k & 1][r];
                if (!(FLAGS & 2)) { if (take) hb[r] |= 1u << k; }
            }
            u1 = u2;
            u2 = (FLAGS & 8) ? nv[R - 1] * 0.5f : __shfl_up_sync(0xffffffffu, nv[R - 1], 1);
#pragma unroll
            for (int r = 0; r < R; ++r) old[r] = nv[r];
            if (k + 2 < 32 && !(FLAGS & 4)) {
                const uint32_t a = (lane <= k + 2 || (FLAGS & 16)) ? curA : prevA;
                for (int r = 0; r < R; ++r) vq[k & 1][r] = lds32(a + (k + 2) * 4 + r * 128);
            }
        }
        for (int r = 0; r < R; ++r) { acc ^= __funnelshift_r(hbp[r], hb[r], lane); hbp[r] = hb[r]; hb[r] = 0; }
    }
    long long t1 = clock64();
    float s = 0;
    for (int r = 0; r < R; ++r) s += old[r];
    out[lane] = s + (float)acc;
    if (lane == 0) cyc[0] = t1 - t0;
}
template <int R, int FLAGS> void run()
{
    const int n = 4096;
    float* out; long long* cyc;
    cudaMalloc(&out, 128); cudaMalloc(&cyc, 8);
    cudaFuncSetAttribute(body<R, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    for (int it = 0; it < 2; ++it) body<R, FLAGS><<<1, 32, 70000>>>(out, cyc, n);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("R=%d flags=%2d [%s%s%s%s%s]  %7.2f cycles/frame (%s)\n", R, FLAGS, FLAGS & 1 ? "fmax " : "", FLAGS & 2 ? "nobits " : "", FLAGS & 4 ? "noload " : "",
           FLAGS & 8 ? "noshfl " : "", FLAGS & 16 ? "nosel " : "", (double)h / n, cudaGetErrorString(cudaGetLastError()));
}
int main() { run<2, 0>(); run<2, 1>(); run<2, 2>(); run<2, 3>(); run<2, 4>(); run<2, 8>(); run<2, 16>(); run<2, 6>(); run<2, 7>(); run<2, 14>(); run<2, 15>(); run<2, 31>(); run<1, 0>(); run<1, 15>(); run<3, 0>(); run<4, 0>(); return 0; }
