// Micro-benchmark: lag-1 systolic form (lane l one frame behind lane l-1), scalar shared loads at per-lane ring offsets.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ float lds32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }

template <int R, int HASIN, int LAG>
__global__ void lag1(float* out, long long* cyc, int nframes)
{
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x;
    constexpr int RL = 128;                       // ring frames per row
    constexpr int PITCH = RL * 4;                 // bytes
    constexpr int LSTR = R * PITCH + 16;          // lane stride
    for (int i = lane; i < (32 * LSTR) / 4 + 256; i += 32) sm[i] = 0.001f * (float)((i * 7919) % 1000) - 0.5f;
    __syncwarp();
    float old[R]; uint32_t hb[R], hbp[R];
    for (int r = 0; r < R; ++r) { old[r] = -1e9f; hb[r] = 0; hbp[r] = 0; }
    float uq[LAG + 1];
    for (int i = 0; i <= LAG; ++i) uq[i] = -1e9f;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + lane * LSTR;
    const uint32_t bnd = (uint32_t)__cvta_generic_to_shared(sm) + 32 * LSTR;
    const bool lane0 = lane == 0;
    uint32_t foff = (uint32_t)((0 - LAG * lane) & (RL - 1)) * 4;
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int y = 0; y < nframes; y += 32) {
        float4 bin[8];
        if (HASIN) for (int g = 0; g < 8; ++g) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bin[g].x), "=f"(bin[g].y), "=f"(bin[g].z), "=f"(bin[g].w) : "r"(bnd + ((y + 4 * g) & 127) * 4));
        float vn[R];
        for (int r = 0; r < R; ++r) vn[r] = lds32(base + r * PITCH + foff);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            float v[R];
            for (int r = 0; r < R; ++r) v[r] = vn[r];
            foff = (foff + 4) & (PITCH - 1);
            for (int r = 0; r < R; ++r) vn[r] = lds32(base + r * PITCH + foff);
            float bk = -1e9f;
            if (HASIN) { const float4 b4 = bin[k >> 2]; bk = (k & 3) == 0 ? b4.x : (k & 3) == 1 ? b4.y : (k & 3) == 2 ? b4.z : b4.w; }
            const float upv = lane0 ? bk : uq[0];
            float nv[R];
#pragma unroll
            for (int r = R - 1; r >= 0; --r) {
                const float stay = old[r];
                const float move = (r == 0) ? upv : old[r - 1];
                const bool take = move > stay;
                nv[r] = (take ? move : stay) + v[r];
                if (take) hb[r] |= 1u << k;
            }
            for (int i = 0; i < LAG; ++i) uq[i] = uq[i + 1];
            uq[LAG] = __shfl_up_sync(0xffffffffu, nv[R - 1], 1);
#pragma unroll
            for (int r = 0; r < R; ++r) old[r] = nv[r];
        }
        for (int r = 0; r < R; ++r) { acc ^= __funnelshift_r(hbp[r], hb[r], LAG * lane); hbp[r] = hb[r]; hb[r] = 0; }
    }
    long long t1 = clock64();
    float s = 0;
    for (int r = 0; r < R; ++r) s += old[r];
    out[lane] = s + (float)acc;
    if (lane == 0) cyc[0] = t1 - t0;
}
template <int R, int HASIN, int LAG> void run(const char* name)
{
    const int n = 4096;
    float* out; long long* cyc;
    cudaMalloc(&out, 128); cudaMalloc(&cyc, 8);
    size_t smem = 32 * (R * 512 + 16) + 1024 + 64;
    cudaFuncSetAttribute(lag1<R, HASIN, LAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int it = 0; it < 2; ++it) lag1<R, HASIN, LAG><<<1, 32, smem>>>(out, cyc, n);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-20s lag=%d R=%d in=%d  %7.2f cycles/frame (%s)\n", name, LAG, R, HASIN, (double)h / n, cudaGetErrorString(cudaGetLastError()));
}
int main() { run<1,1,1>("x"); run<1,1,2>("x"); run<1,1,3>("x"); run<2,1,1>("x"); run<2,1,2>("x"); run<2,1,3>("x"); run<2,1,4>("x"); run<3,1,2>("x"); run<3,1,3>("x"); run<4,1,2>("x"); return 0; }
