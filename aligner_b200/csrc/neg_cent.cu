// neg_cent.cu -- the score matrices that feed monotonic alignment search, fp32 on sm_100a.
//
// The reference snapshot holds no code for these; the formulas are the published ones of
// Glow-TTS / VITS (`neg_cent1..4`) and of the OTA aligner (NeMo AlignmentEncoder), restated in fp64 in
// oracle/neg_cent.py.  Output layout is the reference API's: [b, t_text, t_mel], t_mel contiguous
// (monotonic_align/__init__.py:8), ready for alb200_mas_device.
//
//   gaussian  neg_cent[b,x,y] = sum_c log N(z[b,c,y]; m[b,c,x], exp(logs[b,c,x])^2)
//             = rowterm[x] + sum_k A[k][x] * Bm[k][y],  k = 2c   : A = s2        Bm = -0.5 z^2
//                                                       k = 2c+1 : A = m * s2    Bm = z
//             an [t_x, 2C] x [2C, t_y] contraction per utterance with a rank-1 epilogue.  Both operands are
//             read in their native layouts (t_x and t_y contiguous), transformed on the fly into shared
//             memory, accumulated in fp32 FFMA in a fixed order (deterministic, and within 1e-5 of fp64).
//   ota       logp[b,x,y] = log_softmax_x(-T * sum_c (q[b,c,y] - k[b,c,x])^2) + log(prior + 1e-8)
//             the squared distance is formed from differences (no |q|^2+|k|^2-2kq cancellation); a CTA owns
//             32 mel frames, keeps their whole text column in shared memory and normalises it there, so the
//             score matrix is written exactly once.
#include "../../include/aligner_b200.h"
#include "alb_opts.h"

#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

namespace alb {
extern thread_local char g_err[512];          // shared with mas_api.cu: alb200_last_error() reports both
extern thread_local uint64_t g_launches;
}

namespace albnc {

// ------------------------------------------------------------------ Gaussian prior (Glow-TTS / VITS)
constexpr int GBM = 64, GBN = 128, GKC = 8;     // tile: 64 text rows x 128 mel frames, 8 channels (16 k) per step

__global__ void __launch_bounds__(256) gaussian_kernel(const float* __restrict__ z, const float* __restrict__ m,
                                                       const float* __restrict__ logs, float* __restrict__ out, int C, int Tx, int Ty)
{
    __shared__ __align__(16) float As[2][2 * GKC][GBM];     // [buf][k][x]
    __shared__ __align__(16) float Bs[2][2 * GKC][GBN];     // [buf][k][y]
    __shared__ float rowpart[4][GBM];
    const int tid = threadIdx.x;
    const int b = blockIdx.z, x0 = blockIdx.y * GBM, y0 = blockIdx.x * GBN;
    const int tx8 = tid & 15, ty4 = tid >> 4;                // thread tile: rows ty4*4..+3, cols tx8*8..+7
    // loader roles
    const int ax = tid & 63, acg = tid >> 6;                 // A: x column ax, channels c0 + acg and c0 + acg + 4
    const int by = tid & 127, bcg = tid >> 7;                // B: y column by, channels c0 + bcg + 2j
    const float* mb = m + (size_t)b * C * Tx;
    const float* lb = logs + (size_t)b * C * Tx;
    const float* zb = z + (size_t)b * C * Ty;
    const bool a_ok = x0 + ax < Tx, b_ok = y0 + by < Ty;

    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float rsum = 0.f;                                        // this thread's share of rowterm[ax]
    float ra[2][2], rb[4][2];                                // prefetched, already transformed operands

    auto fetch = [&](int c0) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = c0 + acg + 4 * j;
            float s2 = 0.f, ms2 = 0.f;
            if (a_ok && c < C) {
                const float mm = mb[(size_t)c * Tx + x0 + ax], lg = lb[(size_t)c * Tx + x0 + ax];
                s2 = expf(-2.f * lg);
                ms2 = mm * s2;
                rsum += (-0.9189385332046727f - lg) - 0.5f * mm * ms2;      // -0.5 log(2 pi) - logs - 0.5 m^2 s2
            }
            ra[j][0] = s2; ra[j][1] = ms2;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + bcg + 2 * j;
            float zz = 0.f;
            if (b_ok && c < C) zz = zb[(size_t)c * Ty + y0 + by];
            rb[j][0] = -0.5f * zz * zz; rb[j][1] = zz;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int cc = acg + 4 * j;
            As[buf][2 * cc][ax] = ra[j][0];
            As[buf][2 * cc + 1][ax] = ra[j][1];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int cc = bcg + 2 * j;
            Bs[buf][2 * cc][by] = rb[j][0];
            Bs[buf][2 * cc + 1][by] = rb[j][1];
        }
    };

    fetch(0);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int c0 = 0; c0 < C; c0 += GKC) {
        const bool more = c0 + GKC < C;
        if (more) fetch(c0 + GKC);                           // global loads in flight while we multiply
#pragma unroll
        for (int kk = 0; kk < 2 * GKC; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][ty4 * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx8 * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx8 * 8 + 4]);
            const float av[4] = { a.x, a.y, a.z, a.w };
            const float bv[8] = { b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w };
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) stash(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
    rowpart[acg][ax] = rsum;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int xl = ty4 * 4 + i, x = x0 + xl;
        if (x >= Tx) continue;
        const float rt = ((rowpart[0][xl] + rowpart[1][xl]) + rowpart[2][xl]) + rowpart[3][xl];   // fixed order: deterministic
        float* o = out + ((size_t)b * Tx + x) * Ty + y0 + tx8 * 8;
        const int ybase = y0 + tx8 * 8;
        if (ybase + 8 <= Ty && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
            reinterpret_cast<float4*>(o)[0] = make_float4(acc[i][0] + rt, acc[i][1] + rt, acc[i][2] + rt, acc[i][3] + rt);
            reinterpret_cast<float4*>(o)[1] = make_float4(acc[i][4] + rt, acc[i][5] + rt, acc[i][6] + rt, acc[i][7] + rt);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (ybase + j < Ty) o[j] = acc[i][j] + rt;
        }
    }
}

// ------------------------------------------------------------------ OTA: L2 distance + log-softmax over text + prior
constexpr int OBN = 32, OBX = 64;     // a CTA owns 32 mel frames; text rows are taken 64 at a time

__global__ void __launch_bounds__(256) ota_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ prior,
                                                  const int32_t* __restrict__ x_lengths, float* __restrict__ out, float temperature,
                                                  int C, int Tx, int Ty, int use_global_scratch)
{
    extern __shared__ __align__(16) float sm[];
    float* qs = sm;                                  // [C][OBN]
    float* ks = qs + (size_t)C * OBN;                // [C][OBX]
    float* lse = ks + (size_t)C * OBX;               // [OBN]
    float* red = lse + OBN;                          // [8][OBN] x 2
    float* dbuf = red + 2 * 8 * OBN;                 // [Tx][OBN+1] when it fits
    const int tid = threadIdx.x;
    const int b = blockIdx.y, y0 = blockIdx.x * OBN;
    const int tlen = x_lengths ? min(max(x_lengths[b], 0), Tx) : Tx;
    const float* qb = q + (size_t)b * C * Ty;
    const float* kb = k + (size_t)b * C * Tx;
    float* ob = out + (size_t)b * Tx * Ty;
    const int DS = OBN + 1;

    for (int i = tid; i < C * OBN; i += 256) {
        const int c = i / OBN, yy = i - c * OBN;
        qs[i] = (y0 + yy < Ty) ? qb[(size_t)c * Ty + y0 + yy] : 0.f;
    }
    const int yp = tid & 15, xg = tid >> 4;          // thread tile: frames 2*yp, 2*yp+1; rows xg*4 .. +3 of the chunk
    for (int xc = 0; xc < tlen; xc += OBX) {
        __syncthreads();
        for (int i = tid; i < C * OBX; i += 256) {
            const int c = i / OBX, xx = i - c * OBX;
            ks[i] = (xc + xx < Tx) ? kb[(size_t)c * Tx + xc + xx] : 0.f;
        }
        __syncthreads();
        float d[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i][0] = d[i][1] = 0.f;
        for (int c = 0; c < C; ++c) {
            const float2 qq = *reinterpret_cast<const float2*>(&qs[c * OBN + 2 * yp]);
            const float4 kk = *reinterpret_cast<const float4*>(&ks[c * OBX + 4 * xg]);
            const float kv[4] = { kk.x, kk.y, kk.z, kk.w };
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float e0 = qq.x - kv[i], e1 = qq.y - kv[i];
                d[i][0] = fmaf(e0, e0, d[i][0]);
                d[i][1] = fmaf(e1, e1, d[i][1]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int x = xc + 4 * xg + i;
            if (x < tlen) {
                const float v0 = -temperature * d[i][0], v1 = -temperature * d[i][1];
                if (!use_global_scratch) {
                    dbuf[x * DS + 2 * yp] = v0; dbuf[x * DS + 2 * yp + 1] = v1;
                } else {
                    if (y0 + 2 * yp < Ty) ob[(size_t)x * Ty + y0 + 2 * yp] = v0;
                    if (y0 + 2 * yp + 1 < Ty) ob[(size_t)x * Ty + y0 + 2 * yp + 1] = v1;
                }
            }
        }
    }
    __syncthreads();
    // column-wise logsumexp over the text axis: 8 threads per frame, fixed reduction order
    {
        const int col = tid >> 3, part = tid & 7;
        const bool cok = y0 + col < Ty;
        float mx = -INFINITY;
        for (int x = part; x < tlen; x += 8) {
            const float v = !use_global_scratch ? dbuf[x * DS + col] : (cok ? ob[(size_t)x * Ty + y0 + col] : 0.f);
            mx = fmaxf(mx, v);
        }
        red[part * OBN + col] = mx;
        __syncthreads();
        float gm = red[col];
#pragma unroll
        for (int j = 1; j < 8; ++j) gm = fmaxf(gm, red[j * OBN + col]);
        float s = 0.f;
        for (int x = part; x < tlen; x += 8) {
            const float v = !use_global_scratch ? dbuf[x * DS + col] : (cok ? ob[(size_t)x * Ty + y0 + col] : 0.f);
            s += expf(v - gm);
        }
        red[8 * OBN + part * OBN + col] = s;
        __syncthreads();
        if (part == 0) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) t += red[8 * OBN + j * OBN + col];
            lse[col] = gm + logf(t);
        }
    }
    __syncthreads();
    {
        const int yy = tid & 31, xr = tid >> 5;
        const bool yok = y0 + yy < Ty;
        const float l = lse[yy];
        for (int x = xr; x < Tx; x += 8) {
            if (!yok) continue;
            const size_t idx = (size_t)x * Ty + y0 + yy;
            float v;
            if (x < tlen) {
                v = (!use_global_scratch ? dbuf[x * DS + yy] : ob[idx]) - l;
                if (prior) v += logf(prior[(size_t)b * Tx * Ty + idx] + 1e-8f);
            } else {
                v = -INFINITY;                        // text padding is excluded from the softmax
            }
            ob[idx] = v;
        }
    }
}

static int nc_fail(int code, const char* msg)
{
    snprintf(alb::g_err, sizeof(alb::g_err), "%s", msg);
    return code;
}

}  // namespace albnc

using namespace albnc;

extern "C" int alb200_neg_cent_gaussian_tc(const float*, const float*, const float*, float*, int, int, int, int, void*);
extern "C" int alb200_neg_cent_ota_tc(const float*, const float*, const float*, const int32_t*, float*, float, int, int, int, int, void*);
extern "C" int alb200_neg_cent_gaussian_v2(const float*, const float*, const float*, float*, int, int, int, int, void*, size_t, void*);
extern "C" int alb200_neg_cent_ota_v2(const float*, const float*, const float*, const int32_t*, float*, float, int, int, int, int, void*, size_t, void*);
extern "C" size_t alb200_neg_cent_workspace_bytes(int mode, int b, int c, int tx, int ty);

// Dispatch order (every path is CUDA; the later ones exist for shapes the earlier ones do not take and as cross-checks):
//   1. neg_cent_v2.cu   tcgen05 fp16x3, TMA-fed, warp-specialised, persistent   (t_x <= 512, t_y % 4 == 0, a workspace)
//   2. neg_cent_tc.cu   tcgen05 tf32x3, operands fetched with plain loads       (option "nc_v1"; OTA: t_x <= 512)
//   3. the fp32 FFMA kernels above                                              (option "nc_ffma"; OTA with t_x > 512)
extern "C" {

int alb200_neg_cent_gaussian_ws(const float* z, const float* m_p, const float* logs_p, float* out, int b, int c, int tx, int ty, void* workspace,
                                size_t workspace_bytes, void* stream)
{
    if (!z || !m_p || !logs_p || !out || b < 0 || c <= 0 || tx <= 0 || ty <= 0) return nc_fail(ALB200_E_INVALID, "neg_cent_gaussian: null pointer or bad shape");
    if (b == 0) return 0;
    if (b > 65535) return nc_fail(ALB200_E_UNSUPPORTED, "neg_cent_gaussian: batch > 65535");
    const alb::Opts& o = alb::opts();
    if (!o.nc_ffma) {
        if (!o.nc_v1 && workspace) {
            const int rc = alb200_neg_cent_gaussian_v2(z, m_p, logs_p, out, b, c, tx, ty, workspace, workspace_bytes, stream);
            if (rc != ALB200_E_UNSUPPORTED) return rc;
        }
        return alb200_neg_cent_gaussian_tc(z, m_p, logs_p, out, b, c, tx, ty, stream);
    }
    dim3 grid((ty + GBN - 1) / GBN, (tx + GBM - 1) / GBM, b);
    gaussian_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z, m_p, logs_p, out, c, tx, ty);
    ++alb::g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return nc_fail(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? ALB200_E_NO_DEVICE : ALB200_E_CUDA, cudaGetErrorString(e));
    return 0;
}

int alb200_neg_cent_ota_ws(const float* queries, const float* keys, const float* prior, const int32_t* x_lengths, float* out,
                           float temperature, int b, int c, int tx, int ty, void* workspace, size_t workspace_bytes, void* stream)
{
    if (!queries || !keys || !out || b < 0 || c <= 0 || tx <= 0 || ty <= 0) return nc_fail(ALB200_E_INVALID, "neg_cent_ota: null pointer or bad shape");
    if (b == 0) return 0;
    if (b > 65535) return nc_fail(ALB200_E_UNSUPPORTED, "neg_cent_ota: batch > 65535");
    const alb::Opts& o = alb::opts();
    if (tx <= 512 && !o.nc_ffma) {
        // the log-softmax over the text axis is done on the accumulator in tensor memory, which holds 512 tokens
        if (!o.nc_v1 && workspace) {
            const int rc = alb200_neg_cent_ota_v2(queries, keys, prior, x_lengths, out, temperature, b, c, tx, ty, workspace, workspace_bytes, stream);
            if (rc != ALB200_E_UNSUPPORTED) return rc;
        }
        return alb200_neg_cent_ota_tc(queries, keys, prior, x_lengths, out, temperature, b, c, tx, ty, stream);
    }
    size_t fixed = ((size_t)c * OBN + (size_t)c * OBX + OBN + 2 * 8 * OBN) * sizeof(float);
    size_t dbytes = (size_t)tx * (OBN + 1) * sizeof(float);
    int optin = 0, dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return nc_fail(ALB200_E_NO_DEVICE, cudaGetErrorString(e));
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (fixed > (size_t)optin) return nc_fail(ALB200_E_UNSUPPORTED, "neg_cent_ota: channel count too large for shared memory");
    // keep the column in shared memory when two CTAs still fit on an SM, else use the output as scratch (L2 resident)
    const int use_global = (fixed + dbytes > (size_t)optin / 2) ? 1 : 0;
    const size_t smem = fixed + (use_global ? 0 : dbytes);
    e = cudaFuncSetAttribute(ota_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
    if (e != cudaSuccess) return nc_fail(ALB200_E_CUDA, cudaGetErrorString(e));
    dim3 grid((ty + OBN - 1) / OBN, b);
    ota_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(queries, keys, prior, x_lengths, out, temperature, c, tx, ty, use_global);
    ++alb::g_launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return nc_fail(ALB200_E_CUDA, cudaGetErrorString(e));
    return 0;
}

// The entries without a workspace argument take the scratch of the TMA path from the stream-ordered allocator
// (cudaMallocAsync / cudaFreeAsync on `stream`: no synchronisation, capturable into a CUDA graph).
static int with_async_workspace(int mode, int b, int c, int tx, int ty, void* stream, void** ws, size_t* bytes)
{
    *ws = nullptr; *bytes = 0;
    const alb::Opts& o = alb::opts();
    if (o.nc_ffma || o.nc_v1 || b <= 0) return 0;
    const size_t need = alb200_neg_cent_workspace_bytes(mode, b, c, tx, ty);
    if (!need) return 0;
    if (cudaMallocAsync(ws, need, (cudaStream_t)stream) != cudaSuccess) { cudaGetLastError(); *ws = nullptr; return 0; }   // fall through to the next path
    *bytes = need;
    return 0;
}

int alb200_neg_cent_gaussian(const float* z, const float* m_p, const float* logs_p, float* out, int b, int c, int tx, int ty, void* stream)
{
    void* ws; size_t bytes;
    with_async_workspace(0, b, c > 0 ? c : 1, tx, ty, stream, &ws, &bytes);
    const int rc = alb200_neg_cent_gaussian_ws(z, m_p, logs_p, out, b, c, tx, ty, ws, bytes, stream);
    if (ws) cudaFreeAsync(ws, (cudaStream_t)stream);
    return rc;
}

int alb200_neg_cent_ota(const float* queries, const float* keys, const float* prior, const int32_t* x_lengths, float* out,
                        float temperature, int b, int c, int tx, int ty, void* stream)
{
    void* ws; size_t bytes;
    with_async_workspace(1, b, c > 0 ? c : 1, tx, ty, stream, &ws, &bytes);
    const int rc = alb200_neg_cent_ota_ws(queries, keys, prior, x_lengths, out, temperature, b, c, tx, ty, ws, bytes, stream);
    if (ws) cudaFreeAsync(ws, (cudaStream_t)stream);
    return rc;
}

}  // extern "C"
