// Micro-benchmark of the real forward_unit from mas_kernel.cuh: one warp alone, fake tile data.
#include <cstdio>
#include "../../aligner_b200/csrc/mas_kernel.cuh"
using namespace alb;

template <int R, bool SKEW, bool DIAG>
__global__ void ub(float* out, long long* cyc, int nframes, int has_in_i, uint32_t* bits_g)
{
    extern __shared__ __align__(128) unsigned char sm[];
    const int lane = threadIdx.x;
    float* f = reinterpret_cast<float*>(sm);
    const int nfl = (32 * R * 32 * 4 + 32 * 16 + 1024 + 8192) / 4;
    for (int i = lane; i < nfl; i += 32) f[i] = 0.001f * (float)((i * 7919) % 1000) - 0.5f;
    __syncwarp();
    const uint32_t s0 = smem_u32(sm);
    const uint32_t tile = s0 + lane * (R * 128 + 16);
    const uint32_t bnd_in = s0 + 32 * R * 128 + 512, bnd_out = bnd_in + 256;
    uint32_t* bits = bits_g ? bits_g : reinterpret_cast<uint32_t*>(sm + 32 * R * 128 + 512 + 1024);
    Fwd<R> S;
    for (int r = 0; r < R; ++r) { S.old[r] = -1e9f; S.wbits[r] = 0; }
    S.up = S.lastp = S.bprev = -1e9f;
    for (int k = 0; k < 4; ++k) S.upn[k] = -1e9f;
    const bool has_in = has_in_i != 0;
    const int lag = SKEW ? 4 * lane : 0;
    long long t0 = clock64();
    for (int y = 0; y < nframes; y += 16) {
        const int yl = y - lag;
        forward_unit<R, 32, 16, SKEW, DIAG>(S, tile + (y & 16) * 4, bnd_in, bnd_out, y, yl, has_in, lane == 0, lane == 31, -1e9f,
                                            DIAG ? (lane * R - yl) : 0, bits + lane * R, 64, 0, 1u << 30);
    }
    long long t1 = clock64();
    float s = 0; uint32_t b = 0;
    for (int r = 0; r < R; ++r) { s += S.old[r]; b ^= S.wbits[r]; }
    out[lane] = s + (float)b;
    if (lane == 0) cyc[0] = t1 - t0;
}

template <int R, bool SKEW, bool DIAG>
void run(const char* name, int has_in, bool gbits)
{
    const int n = 1024;
    float* out; long long* cyc; uint32_t* gb = nullptr;
    cudaMalloc(&out, 128); cudaMalloc(&cyc, 8);
    if (gbits) cudaMalloc(&gb, 64 * 64 * 4);
    size_t smem = 32 * R * 128 + 512 + 1024 + 8192;
    cudaFuncSetAttribute(ub<R, SKEW, DIAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int it = 0; it < 2; ++it) ub<R, SKEW, DIAG><<<1, 32, smem>>>(out, cyc, n, has_in, gb);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-40s R=%d skew=%d diag=%d in=%d gbits=%d  %7.2f cycles/frame (%s)\n", name, R, SKEW, DIAG, has_in, gbits, (double)h / n,
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    run<2, false, false>("lockstep", 0, false);
    run<2, false, false>("lockstep", 1, false);
    run<2, true, false>("skewed", 0, false);
    run<2, true, false>("skewed", 1, false);
    run<2, true, true>("skewed diag", 0, false);
    run<2, true, false>("skewed gbits", 0, true);
    run<4, true, false>("skewed", 1, false);
    run<4, false, false>("lockstep", 1, false);
    return 0;
}
