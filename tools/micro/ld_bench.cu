// Micro-benchmark: per-SM global->shared streaming rate of (a) LDGSTS.128 issued by W warps, (b) 2-D TMA box loads.
// Access pattern of the MAS loader: 64 rows x 32 frames (128 B per row) per tile, rows 4000 B apart, tiles march along the row.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory"); } while (!ok);
}

constexpr int ROWS = 64, TF = 32, NS = 4;

template <int MODE>   // 0: LDGSTS.128, 1: TMA 2D
__global__ void __launch_bounds__(256) k(const float* v, int Ty, int tiles, long long* cyc, const __grid_constant__ CUtensorMap tm)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t ring = s32(smem) + 1024 + w * NS * (ROWS * TF * 4);
    const uint32_t bar0 = s32(smem) + w * NS * 8;
    if (lane == 0) for (int s = 0; s < NS; ++s) mbar_init(bar0 + 8 * s, MODE == 0 ? 32 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int row0 = (blockIdx.x * (blockDim.x >> 5) + w) * ROWS;
    const float* base = v + (size_t)row0 * Ty;
    long long t0 = clock64();
    int stage = 0, phase = 0;
    for (int t = 0; t < tiles + NS; ++t) {
        if (t >= NS) { mbar_wait(bar0 + 8 * ((t - NS) % NS), ((t - NS) / NS) & 1); }   // tile t-NS landed -> its stage is free
        if (t < tiles) {
            const uint32_t st = ring + stage * (ROWS * TF * 4);
            if (MODE == 0) {
                const int ck = lane & 7, q = lane >> 3;
#pragma unroll 4
                for (int i = q; i < ROWS; i += 4)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(st + i * 128 + ck * 16), "l"(base + (size_t)i * Ty + t * TF + ck * 4) : "memory");
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar0 + 8 * stage) : "memory");
            } else if (lane == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * stage), "r"(ROWS * TF * 4) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(st), "l"(&tm), "r"(t * TF), "r"(row0), "r"(bar0 + 8 * stage) : "memory");
            }
            if (++stage == NS) stage = 0;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    const int Ty = 1000, tiles = Ty / TF;
    const int maxrows = 148 * 8 * ROWS;
    float* v; cudaMalloc(&v, (size_t)maxrows * Ty * 4); cudaMemset(v, 0, (size_t)maxrows * Ty * 4);
    long long* cyc; cudaMalloc(&cyc, 148 * 8);
    EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
    CUtensorMap tm;
    cuuint64_t dims[2] = { (cuuint64_t)Ty, (cuuint64_t)maxrows }, strides[1] = { (cuuint64_t)Ty * 4 };
    cuuint32_t box[2] = { TF, ROWS }, es[2] = { 1, 1 };
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, v, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode)
        for (int grid : { 64, 148 })
            for (int W : { 1, 2, 4, 8 }) {
                size_t smem = 1024 + (size_t)W * NS * ROWS * TF * 4;
                if (smem > 227 * 1024) continue;
                auto fn = mode == 0 ? k<0> : k<1>;
                cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                fn<<<grid, W * 32, smem>>>(v, Ty, tiles, cyc, tm);
                cudaEventRecord(e0);
                fn<<<grid, W * 32, smem>>>(v, Ty, tiles, cyc, tm);
                cudaEventRecord(e1);
                cudaDeviceSynchronize();
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
                double bytes = (double)grid * W * ROWS * tiles * TF * 4;
                printf("%s grid=%3d warps=%d: %8.1f cycles/tile/warp  %6.2f B/clk/SM  %7.1f GB/s total  (%s)\n", mode ? "TMA2D " : "LDGSTS", grid, W,
                       (double)h / tiles, (double)W * ROWS * TF * 4 * tiles / h, bytes / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
            }
    return 0;
}
