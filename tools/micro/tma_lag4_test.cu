// Feasibility test for the 4-frame-skew tile: ONE 3-D box per 32-frame tile delivers lane l's frames F0 - 4l .. F0 - 4l + 31.
//   dims (frames [Ty + 4(nl-1)], lane [nl, stride R*Ty*4 - 16 B], row [nrows, stride Ty*4]); box (32, 32, R); 128-byte swizzle.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_lag4_test tma_lag4_test.cu
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
__device__ __forceinline__ void mbar_init(uint32_t a, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t a, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t ph) {
    uint32_t ok = 0, n = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(ph) : "memory");
        if (++n > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void tma3(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
template <int R>
__global__ void k(const __grid_constant__ CUtensorMap map, const float* g, int Ty, int nrows, int nl, int row0, int F0, int* bad, long long* cyc, int reps)
{
    extern __shared__ __align__(1024) unsigned char sm[];
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(sm);
    const uint32_t bar = s0 + R * 4096;
    const int lane = threadIdx.x;
    if (lane == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    long long t0 = clock64();
    for (int it = 0; it < reps; ++it) {
        if (lane == 0) { mbar_expect(bar, R * 4096); tma3(s0, &map, F0, 0, row0, bar); }
        mbar_wait(bar, it & 1);
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = (t1 - t0) / reps;
    int nbad = 0;
    for (int r = 0; r < R; ++r)
        for (int kk = 0; kk < 32; ++kk) {
            const uint32_t off = (r * 32 + lane) * 128 + ((((kk >> 2) ^ lane) & 7) << 4) + (kk & 3) * 4;
            const float v = *reinterpret_cast<const float*>(sm + off);
            const int row = row0 + lane * R + r, f = F0 + kk - 4 * lane, x = F0 + kk;
            float want;
            if (x < 0 || x >= Ty + 4 * (nl - 1) || lane >= nl || row0 + r >= nrows) want = 0.f;
            else want = g[(long long)row * Ty + f];
            if (v != want) { ++nbad; if (nbad < 3) printf("lane %d r %d k %d: got %f want %f (row %d frame %d)\n", lane, r, kk, v, want, row, f); }
        }
    atomicAdd(bad, nbad);
}
typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncFn enc;
template <int R> int run(const float* d, int Ty, int nrows, int nl)
{
    int* bad; long long* cyc;
    cudaMalloc(&bad, 4); cudaMalloc(&cyc, 8);
    CUtensorMap map;
    cuuint64_t dims[3] = { (cuuint64_t)Ty + 4 * (nl - 1), (cuuint64_t)nl, (cuuint64_t)nrows };
    cuuint64_t strides[2] = { (cuuint64_t)R * Ty * 4 - 16, (cuuint64_t)Ty * 4 };
    cuuint32_t box[3] = { 32, 32, (cuuint32_t)R }, es[3] = { 1, 1, 1 };
    CUresult rr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(d), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode R=%d nl=%d -> %d\n", R, nl, (int)rr);
    if (rr != CUDA_SUCCESS) return 1;
    cudaFuncSetAttribute(k<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, R * 4096 + 64);
    const int cases[][2] = { {64, 512}, {0, 0}, {128, 128}, {128, Ty - 8}, {64, 32}, {nrows - nl * R, Ty}, {nrows - nl * R, Ty + 64}, {8, 992} };
    int total = 0;
    for (auto& cs : cases) {
        cudaMemset(bad, 0, 4);
        k<R><<<1, 32, R * 4096 + 64>>>(map, d, Ty, nrows, nl, cs[0], cs[1], bad, cyc, 200);
        cudaError_t e = cudaDeviceSynchronize();
        int hb = -1; long long hc = 0;
        cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        printf("R=%d nl=%d row0=%d F0=%d: %d mismatches, %lld cycles per tile (%s)\n", R, nl, cs[0], cs[1], hb, hc, cudaGetErrorString(e));
        total += hb;
    }
    return total;
}
int main()
{
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    enc = (EncFn)fn;
    const int Ty = 1000, nrows = 1024;
    std::vector<float> h((size_t)nrows * Ty);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003) * 0.5f + 1.f;
    float* d; cudaMalloc(&d, h.size() * 4);     // exact size: an overrun past the last row shows up under compute-sanitizer
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    int bad = run<2>(d, Ty, nrows, 32) + run<4>(d, Ty, nrows, 32) + run<4>(d, Ty, nrows, 18) + run<5>(d, Ty, nrows, 28) + run<3>(d, Ty, nrows, 7);
    printf("%s\n", bad ? "FAILED" : "ALL OK");
    return 0;
}
