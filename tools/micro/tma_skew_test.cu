// Feasibility test: can ONE tensor map deliver a per-lane skewed tile?
// values viewed as [rows][Ty] fp32.  Pipeline position p = 4*m + c (m = 0..7, c = 0..3) owns rows p*R .. p*R+R-1 and wants frames
// F0 - p .. F0 - p + 31.  Map: dim0 = frames (4 B), dim1 = m with stride 4*R*Ty*4 - 16 B (four positions further and four frames
// earlier), dim2 = row with stride Ty*4.  One box (32 frames, 8 m, R rows) per class c at coordinates (F0 - c, 0, row0 + c*R),
// 128-byte swizzle so that the eight positions of a class read one 16-byte chunk each without bank conflicts.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_skew_test tma_skew_test.cu -lcuda
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

__device__ int g_mode;
__device__ int g_shift = 1;
template <int R>
__global__ void k(const __grid_constant__ CUtensorMap map, const float* g, int Ty, int nrows, int row0, int F0, int* bad, long long* cyc, int reps)
{
    extern __shared__ __align__(1024) unsigned char sm[];
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(sm);
    const uint32_t bar = s0 + 4 * R * 1024;
    const int lane = threadIdx.x;
    if (lane == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    long long t0 = clock64();
    for (int it = 0; it < reps; ++it) {
        if (lane == 0) {
            mbar_expect(bar, 4 * R * 1024);
            for (int c = 0; c < 4; ++c) { if (g_mode & 2) tma3(s0 + c * R * 1024, &map, F0 - g_shift * c, row0 + c * R, 0, bar); else tma3(s0 + c * R * 1024, &map, F0 - g_shift * c, 0, row0 + c * R, bar); }
        }
        mbar_wait(bar, it & 1);
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = (t1 - t0) / reps;
    // lane = c * 8 + m handles position p = 4 m + c
    const int c = lane >> 3, m = lane & 7, p = 4 * m + c;
    int nbad = 0;
    for (int r = 0; r < R; ++r)
        for (int kk = 0; kk < 32; ++kk) {
            const int rowi = (g_mode & 2) ? m * R + r : r * 8 + m;
            const uint32_t off = c * R * 1024 + rowi * 128 + ((((kk >> 2) ^ ((g_mode & 1) ? 0 : rowi)) & 7) << 4) + (kk & 3) * 4;
            const float v = *reinterpret_cast<const float*>(sm + off);
            const int row = row0 + p * R + r, f = F0 + kk - p, x = F0 - c + kk;     // x = the frame coordinate the hardware bounds-checks
            float want;
            if (x < 0 || x >= Ty + 28 || row0 + c * R + r >= nrows) want = 0.f;      // out of the map: zero fill
            else want = g[(long long)row * Ty + f];
            if (v != want) { ++nbad; if (nbad < 3) printf("lane %d r %d k %d: got %f want %f (row %d frame %d)\n", lane, r, kk, v, want, row, f); }
        }
    atomicAdd(bad, nbad);
}

template <int R> int run(CUtensorMap (*mk)(const float*, int, int, int), const float* d, const std::vector<float>& h, int Ty, int nrows)
{
    int* bad; long long* cyc;
    cudaMalloc(&bad, 4); cudaMalloc(&cyc, 8);
    CUtensorMap map = mk(d, Ty, nrows, R);
    cudaFuncSetAttribute(k<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * R * 1024 + 64);
    const int cases[][2] = { {64, 512}, {0, 0}, {64, 64}, {128, Ty - 8}, {64, 2}, {nrows - 32 * R, Ty}, {8, 992} };
    int total = 0;
    for (auto& cs : cases) {
        cudaMemset(bad, 0, 4);
        k<R><<<1, 32, 4 * R * 1024 + 64>>>(map, d, Ty, nrows, cs[0], cs[1], bad, cyc, 200);
        cudaError_t e = cudaDeviceSynchronize();
        int hb = -1; long long hc = 0;
        cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        printf("R=%d row0=%d F0=%d: %d mismatches, %lld cycles per 4-box tile (%s)\n", R, cs[0], cs[1], hb, hc, cudaGetErrorString(e));
        total += hb;
    }
    return total;
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncFn enc;
static int h_mode = 0;
CUtensorMap mk(const float* d, int Ty, int nrows, int R)
{
    CUtensorMap m;
    cuuint64_t dims[3] = { (cuuint64_t)Ty + 28, 8, (cuuint64_t)nrows };
    cuuint64_t strides[2] = { (cuuint64_t)4 * R * Ty * 4 - 16, (cuuint64_t)Ty * 4 };
    cuuint32_t box[3] = { 32, 8, (cuuint32_t)R }, es[3] = { 1, 1, 1 };
    if (h_mode == 4) { dims[0] = Ty; dims[1] = 8; dims[2] = nrows / 8; strides[0] = (cuuint64_t)Ty * 4; strides[1] = (cuuint64_t)Ty * 32; }
    else if (h_mode == 5) { dims[0] = Ty; }
    else if (h_mode & 2) { dims[1] = nrows; dims[2] = 8; strides[0] = (cuuint64_t)Ty * 4; strides[1] = (cuuint64_t)4 * R * Ty * 4 - 16; box[1] = R; box[2] = 8; }
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(d), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     (h_mode & 1) ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode R=%d -> %d\n", R, (int)r);
    if (r != CUDA_SUCCESS) exit(1);
    return m;
}
int main(int argc, char** argv)
{
    h_mode = argc > 1 ? atoi(argv[1]) : 0;
    const int onlyR = argc > 2 ? atoi(argv[2]) : 0;
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    enc = (EncFn)fn;
    const int Ty = 1000, nrows = 1024;
    std::vector<float> h((size_t)nrows * Ty + 4096);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003) * 0.5f + 1.f;
    float* d; cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(g_mode, &h_mode, 4);
    { int sh = argc > 3 ? atoi(argv[3]) : 1; cudaMemcpyToSymbol(g_shift, &sh, 4); printf("shift %d\n", sh); }
    printf("mode %d (bit0: no swizzle, bit1: dims ordered frames,row,m)\n", h_mode);
    int bad = 0;
    if (!onlyR || onlyR == 2) bad += run<2>(mk, d, h, Ty, nrows);
    if (!onlyR || onlyR == 3) bad += run<3>(mk, d, h, Ty, nrows);
    if (!onlyR || onlyR == 4) bad += run<4>(mk, d, h, Ty, nrows);
    printf("%s\n", bad ? "FAILED" : "ALL OK");
    return 0;
}
