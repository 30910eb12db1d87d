// mas_api.cu -- C ABI (include/aligner_b200.h) over the sm_100a MAS kernel.
//
// Host side of the boundary that replaces the reference's
//   maximum_path_c(paths, values, t_xs, t_ys, max_neg_val)   monotonic_align/core.pyx:40
// There is no CPU fallback in this file: without an sm_100 device every entry
// point returns ALB200_E_NO_DEVICE.
#include "mas_kernel.cuh"
#include "alb_opts.h"
#include "../../include/aligner_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <mutex>

namespace alb {

// ------------------------------------------------------------------ options (environment read once, alb200_set_option afterwards)
static Opts g_opts;
static std::once_flag g_opts_once;
static std::mutex g_opts_mu;
static int set_opt_locked(const char* name, const char* value)
{
    const bool unset = (value == nullptr || value[0] == '\0');
    if (!strcmp(name, "force")) { memset(g_opts.force, 0, sizeof(g_opts.force)); if (!unset) strncpy(g_opts.force, value, sizeof(g_opts.force) - 1); }
    else if (!strcmp(name, "latency_max_b")) g_opts.latency_max_b = unset ? -1 : atoi(value);
    else if (!strcmp(name, "tmap_promo")) g_opts.tmap_promo = unset ? -1 : atoi(value);
    else if (!strcmp(name, "no_tail_box")) g_opts.no_tail_box = unset ? 0 : atoi(value) != 0;
    else if (!strcmp(name, "no_shared_zero")) g_opts.no_shared_zero = unset ? 0 : atoi(value) != 0;
    else if (!strcmp(name, "force_unaligned")) g_opts.force_unaligned = unset ? 0 : atoi(value) != 0;
    else if (!strcmp(name, "dbg")) g_opts.dbg = unset ? 0 : atoi(value);
    else if (!strcmp(name, "nc_ffma")) g_opts.nc_ffma = unset ? 0 : atoi(value) != 0;
    else if (!strcmp(name, "nc_v1")) g_opts.nc_v1 = unset ? 0 : atoi(value) != 0;
    else if (!strcmp(name, "nc_no_pdl")) g_opts.nc_no_pdl = unset ? 0 : atoi(value) != 0;
    else if (!strcmp(name, "fused_seq")) g_opts.fused_seq = unset ? 0 : atoi(value);
    else return -1;
    ++g_opts.gen;
    return 0;
}
static void opts_from_env()
{
    memset(&g_opts, 0, sizeof(g_opts));
    g_opts.latency_max_b = -1; g_opts.tmap_promo = -1;
    static const char* const names[][2] = { {"ALB200_FORCE", "force"}, {"ALB200_LATENCY_MAX_B", "latency_max_b"}, {"ALB200_TMAP_PROMO", "tmap_promo"},
        {"ALB200_NO_TAIL_BOX", "no_tail_box"}, {"ALB200_NO_SHARED_ZERO", "no_shared_zero"}, {"ALB200_FORCE_UNALIGNED", "force_unaligned"}, {"ALB200_DBG", "dbg"}, {"ALB200_NC_FFMA", "nc_ffma"},
        {"ALB200_NC_V1", "nc_v1"}, {"ALB200_NC_NO_PDL", "nc_no_pdl"}, {"ALB200_FUSED_SEQ", "fused_seq"} };
    for (auto& n : names)
        if (const char* e = getenv(n[0])) set_opt_locked(n[1], e[0] ? e : "1");
}
const Opts& opts()
{
    std::call_once(g_opts_once, opts_from_env);
    return g_opts;
}

thread_local char g_err[512] = "";          // also written by neg_cent.cu
thread_local uint64_t g_launches = 0;
static thread_local uint64_t g_h2d = 0, g_d2h = 0;

static int fail(int code, const char* fmt, const char* a = "", long long b = 0, long long c = 0, long long d = 0)
{
    snprintf(g_err, sizeof(g_err), fmt, a, b, c, d);
    return code;
}
#define ALB_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            snprintf(g_err, sizeof(g_err), "%s failed: %s", #call, cudaGetErrorString(e_));  \
            return ALB200_E_CUDA;                                                            \
        }                                                                                    \
    } while (0)

struct DevInfo { int ok, dev, sms, smem_optin, cc_major; };

static int device_info(DevInfo* out)
{
    static thread_local DevInfo cache[16];
    static thread_local bool have[16] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(ALB200_E_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
    if (dev < 16 && have[dev]) { *out = cache[dev]; return 0; }
    DevInfo d; d.dev = dev; d.ok = 1;
    ALB_CUDA(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
    ALB_CUDA(cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    ALB_CUDA(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    if (d.cc_major != 10)
        return fail(ALB200_E_NO_DEVICE, "device is sm_%s%lld0, this library is sm_100a only (no fallback)", "", d.cc_major);
    if (dev < 16) { cache[dev] = d; have[dev] = true; }
    *out = d;
    return 0;
}

// ------------------------------------------------------------------ kernel table
typedef void (*KernelFn)(const MasParams, const CUtensorMap, const CUtensorMap);
struct KEntry { int R, TF, skew, nwmax, minb; KernelFn fn; int vt; };     // vt: score element type (0 fp32, 1 fp16, 2 bf16)
// lock-step form: a 255-register instance for the latency regime and a 128-register one for two CTAs per SM;
// skewed form: latency regime only.
#define ALB_K(R, TF) { R, TF, 0, 4, 1, mas_kernel<R, TF, false, 4, 1> }, { R, TF, 0, 4, 2, mas_kernel<R, TF, false, 4, 2> }
#define ALB_KS(R) { R, 32, 1, 4, 1, mas_kernel<R, 32, true, 4, 1> }
#define ALB_K8(R, TF) { R, TF, 0, 8, 1, mas_kernel<R, TF, false, 8, 1> }
#define ALB_KC(R) { R, 32, 2, 4, 1, mas_kernel<R, 32, true, 4, 1, true> }      // skew code 2 = skewed + cluster hand-off
#define ALB_KS8(R) { R, 32, 1, 8, 1, mas_kernel<R, 32, true, 8, 1> }
#define ALB_KL4(R) { R, 32, 4, 4, 1, mas_kernel<R, 32, true, 4, 1, false, 0, false, kLag4> }   // skew code 4 = skewed, 4-frame lag, pre-skewed boxes
static const KEntry g_kernels[] = {
#ifdef ALB200_FEW_KERNELS          // A/B builds while tuning (build_lib.py --variant): the instances of the latency regime + one throughput instance
    ALB_K(2, 32), ALB_K(4, 16), ALB_KS(1), ALB_KS(2), ALB_KS(3), ALB_KS8(1), ALB_KC(1), ALB_KC(2),
    ALB_KL4(2), ALB_KL4(3), ALB_KL4(4),
#else
    ALB_KL4(2), ALB_KL4(3), ALB_KL4(4),
    ALB_K(1, 32),
    ALB_K(2, 32), ALB_K(2, 16),
    ALB_K(3, 32), ALB_K(3, 16),
    ALB_K(4, 32), ALB_K(4, 16),
    ALB_K(6, 32), ALB_K(6, 16),
    ALB_K(8, 32), ALB_K(8, 16), ALB_K(8, 8),
    ALB_K(16, 16), ALB_K(16, 8),
    ALB_K8(8, 32), ALB_K8(8, 16), ALB_K8(8, 8),      // more than 4 compute warps: t_x > 1024
    ALB_K8(16, 16), ALB_K8(16, 8),
    ALB_KS(1), ALB_KS(2), ALB_KS(3), ALB_KS(4), ALB_KS(6), ALB_KS(8),     // skewed: 32-frame tiles only
    ALB_KS8(1), ALB_KS8(2), ALB_KS8(3), ALB_KS8(4), ALB_KS8(8),
    ALB_KC(1), ALB_KC(2), ALB_KC(3), ALB_KC(4),
    // fp16 / bf16 score input, promoted on load: the skewed/TMA form for the latency regime and the 128-register lock-step
    // form with 16-frame tiles for the throughput regime (the shapes the host picks); anything else reports
    // ALB200_E_UNSUPPORTED and the caller promotes to fp32 on the device first
#define ALB_KH(VT) \
    { 1, 32, 1, 4, 1, mas_kernel<1, 32, true, 4, 1, false, VT>, VT }, { 2, 32, 1, 4, 1, mas_kernel<2, 32, true, 4, 1, false, VT>, VT },   \
    { 3, 32, 1, 4, 1, mas_kernel<3, 32, true, 4, 1, false, VT>, VT }, { 4, 32, 1, 4, 1, mas_kernel<4, 32, true, 4, 1, false, VT>, VT },   \
    { 2, 16, 0, 4, 2, mas_kernel<2, 16, false, 4, 2, false, VT>, VT }, { 4, 16, 0, 4, 2, mas_kernel<4, 16, false, 4, 2, false, VT>, VT }, \
    { 8, 16, 0, 4, 2, mas_kernel<8, 16, false, 4, 2, false, VT>, VT }, { 16, 16, 0, 4, 2, mas_kernel<16, 16, false, 4, 2, false, VT>, VT }
    ALB_KH(1), ALB_KH(2),
    // VITS layout (scores and path stored [b, t_mel, t_text]): skewed/TMA form, fp32, skew code 3
    { 1, 32, 3, 4, 1, mas_kernel<1, 32, true, 4, 1, false, 0, true>, 0 }, { 2, 32, 3, 4, 1, mas_kernel<2, 32, true, 4, 1, false, 0, true>, 0 },
    { 3, 32, 3, 4, 1, mas_kernel<3, 32, true, 4, 1, false, 0, true>, 0 }, { 4, 32, 3, 4, 1, mas_kernel<4, 32, true, 4, 1, false, 0, true>, 0 },
#endif
};
// Latency regime = one utterance per SM at a time: 255-register instances, skewed/TMA form, deepest ring, bits in shared
// memory, and a persistent grid with the work cursor when the batch exceeds the SM count.  It obviously applies while
// b <= #SM; measured (profiles/r01_notes.md, item 21) it also beats the occupancy-driven throughput configuration for
// several SMs' worth of utterances as long as the skewed form applies (t_x <= 512, 16-byte aligned rows) -- the more
// compute warps one utterance keeps busy and the longer its mel axis, the longer.  ALB200_LATENCY_MAX_B overrides (tuning aid).
static bool is_latency(const DevInfo& di, int b, int tx, int ty, bool aligned, int vt)
{
    if (opts().latency_max_b >= 0) return b <= opts().latency_max_b;
    if (b <= di.sms) return true;
    // one CTA per SM cannot hide an utterance's pipeline fill (~63 frames per compute warp) and backtrack behind other
    // utterances, so the mel axis has to be long enough to amortise them
    // (half-precision scores: the lock-step form is the bandwidth-bound one and profits from the halved read -- 4096x200x1000
    //  bf16 1.04 ms lock-step vs 1.18 ms skewed -- so beyond #SM utterances they stay there)
    // (mel axis: re-measured at the end of round 2, tools/regime_check.sh -- with the cheaper unit and the in-place backtrack the
    //  persistent skewed form now also wins at 400-600 frames: 300x100x400 50 vs 60 us, 600x300x500 121 vs 138, 1500x160x500 179 vs
    //  189, 450x400x560 148 vs 156; it still loses at 800x120x360 (99 vs 77), 600x200x300 (109 vs 99) and beyond the batch limits
    //  below: 4096x160x500 451 vs 441, 3000x130x400 273 vs 257)
    if (!aligned || tx > 512 || ty < 400 || vt != 0) return false;
    if (tx <= 128) return b <= 3 * di.sms;                       // two compute warps per SM only
    if (tx <= 256) return ty >= 1000 || b <= 12 * di.sms;        // 4096x200x1000: 81.7 % of HBM peak vs 76.7 %
    // three rows per lane (end of round 2): 1000x300x1500 480 vs 549 us, 2000x300x1500 916 vs 996, 4096x300x1000 1789 vs 1983,
    // 1000x384x1200 455 vs 499, 2000x300x500 348 vs 369; 3000x320x800 838 vs 828.  Four rows per lane: 1200x512x1500 974 vs 924.
    if (tx <= 384) return ty >= 1000 || b <= 14 * di.sms;
    return b <= 5 * di.sms;
}
static KernelFn find_kernel(int R, int TF, int skew, int nw, int minb, int vt)
{
    for (const KEntry& k : g_kernels)
        if (k.R == R && k.TF == TF && k.skew == skew && nw <= k.nwmax && k.minb <= minb && k.vt == vt) return k.fn;   // minb 1 also serves minb 2 requests
    return nullptr;
}

struct Config {
    int R, TF, NW, NS, bits_smem, skew, grid, occ, nc, lag;
    uint32_t smem;
    int64_t bits_slot_words;
    KernelFn fn;
};

// 4-frame-lag form (forward_unit4): measured per 32-frame unit of a compute warp (tools/lag_sweep.py, slope between t_y = 1000
// and 2000, cycles): 1-frame lag R=1/2/3 1140/1120/1350; 4-frame lag R=2/3/4/5/6/8 1060/1210/1390/1690/1850/2360 -- 5-10 % less per
// unit at equal R, but 124 instead of 31 frames of fill per warp (three more units) and 128 more per hand-off.  That only pays for
// ONE compute warp and a long mel axis (64x96x2000: 60.3 -> 58.0 us, 64x64x2000: 53.8 -> 52.6 us; 64x200x1000 loses 10 %).
static bool choose_lag4(int tx, int ty, int* r4, int* nw4)
{
    if (tx > 96 || ty < 1536) return false;
    const int R = (tx + 31) / 32;
    if (R < 2 || tx % R != 0) return false;
    *r4 = R; *nw4 = 1;
    return true;
}

// Picks rows-per-lane, tile width, ring depth and where the direction bits live.
//   latency regime   (b <= #SM): one CTA per SM, <= 4 compute warps when possible (one per scheduler), 255-register
//                    instance, deepest ring that fits, bits in shared memory.
//   throughput regime (b > #SM): 4 rows per lane, and the smallest ring (>= 2 stages) that lets the 128-register
//                    instances reach their register-limited occupancy (8 / compute-warps CTAs per SM), so that about 8
//                    compute warps per SM hide each other's latency and one item's backtrack overlaps others' streaming.
static int select_config_uncached(const DevInfo& di, int b, int tx, int ty, bool want_dur, bool aligned, int vt, int vl, Config* c)
{
    const bool latency = is_latency(di, b, tx, ty, aligned, vt);
    int R, NW;
    if (tx <= 32) { R = 1; NW = 1; }
    else if (latency) {
        // one warp per scheduler: <= 4 compute warps, as few rows per lane as that allows
        if (tx <= 1024) {
            int raw = (tx + 127) / 128;
            if (raw < 2) raw = 2;
            R = raw <= 4 ? raw : (raw <= 6 ? 6 : 8);
        } else R = tx <= 2048 ? 8 : 16;
        NW = (tx + 32 * R - 1) / (32 * R);
    } else {
        // throughput: 4 rows per lane measured best (profiles/r01_shape_sweep.json); what matters is ~8 resident compute
        // warps per SM, i.e. as many CTAs as the 128-register instances allow
        R = tx <= 64 ? 2 : (tx <= 512 ? 4 : (tx <= 2048 ? 8 : 16));
        NW = (tx + 32 * R - 1) / (32 * R);
    }
    int f_tf = 0, f_ns = 0, f_bits = -1, f_skew = -1, f_nc = 0, f_lag = 0;
    int fr = 0;
    if (opts().force[0])                           // "R,TF,NS,bits_smem,skew,cluster,lag" -- tuning / tests only
        sscanf(opts().force, "%d,%d,%d,%d,%d,%d,%d", &fr, &f_tf, &f_ns, &f_bits, &f_skew, &f_nc, &f_lag);
    // Cluster mode: an utterance whose text axis needs more than 4 rows per lane in one CTA (t_x > 512) loses the fast
    // skewed form; split its rows over the CTAs of a thread-block cluster instead (2 rows per lane, 4 compute warps per
    // CTA, boundary rows handed over through distributed shared memory).  Only when every cluster gets its own SMs.
    int NC = 1;
    if (f_nc > 0) NC = f_nc;
    else if (latency && aligned && tx > 512 && fr == 0 && f_skew != 0 && vt == 0 && !vl) {
        const int want = (tx + 255) / 256;
        if (want <= 8 && (int64_t)b * want <= di.sms) NC = want;
    }
    if (NC > 1 && fr == 0) { R = 2; NW = 4; }
    if (fr > 0) { R = fr; NW = (tx + 32 * R * NC - 1) / (32 * R * NC); }
    // 4-frame lag on pre-skewed boxes (forward_unit4): fewer instructions per frame, but 124 frames of fill per warp instead of 31.
    // fp32 [t_text, t_mel] scores, one CTA per utterance, t_x a multiple of the rows per lane (a lane's rows must not run past
    // the tensor).  Chosen where measured faster (choose_lag4); any other shape only through the force option.
    int lag = 1;
    const bool lag4_ok = latency && aligned && vt == 0 && !vl && NC == 1 && tx > 32 && f_skew != 0;
    if (f_lag == kLag4) {
        if (!lag4_ok || fr == 0 || tx % R != 0 || NW > 4) return fail(ALB200_E_UNSUPPORTED, "the 4-frame-lag form cannot take t_x=%s%lld at %lld rows per lane", "", tx, R);
        lag = kLag4;
    } else if (lag4_ok && fr == 0 && f_lag == 0 && b <= di.sms) {
        int r4 = 0, nw4 = 0;
        if (choose_lag4(tx, ty, &r4, &nw4)) { lag = kLag4; R = r4; NW = nw4; }
    }
    if (NC > 1 && (!aligned || NC > 8 || 32 * R * NW * NC < tx))
        return fail(ALB200_E_UNSUPPORTED, "cluster of %s%lld CTAs cannot take t_x=%lld", "", NC, tx);
    if (NW > kMaxWarps)
        return fail(ALB200_E_UNSUPPORTED, "t_x=%s%lld needs too many compute warps at %lld rows per lane", "", tx, R);
    const int nblk = (ty + 31) / 32;
    const int per_sm = di.smem_optin + 1024;                   // 228 KB on sm_100
    const int tfs[3] = { R >= 16 ? 16 : 32, R == 1 ? 32 : 16, R >= 8 ? 8 : (R == 1 ? 32 : 16) };
    int best_tf = 0, best_ns = 0, best_bits = 0;
    // skewed (systolic) forward: lane l runs one frame behind lane l-1, so the neighbour exchange leaves the per-frame
    // dependency chain; costs 31 frames of fill per warp and 32 more per warp hand-off.  32-frame tiles only.  Measured
    // (profiles/r01_skew_sweep.json): faster or equal wherever an utterance owns its SM and has <= 4 rows per lane.
    int want_skew = f_skew >= 0 ? f_skew : ((latency && R <= 4) ? 1 : 0);
    if (NC > 1 || lag == kLag4) want_skew = 1;
    if (vl && (NC > 1 || vt != 0)) return fail(ALB200_E_UNSUPPORTED, "the [t_mel, t_text] layout has fp32 single-CTA kernels only%s", "");
    if (!aligned || R > 8) want_skew = 0;                      // its tiles come in by TMA: 16-byte aligned rows, <= 256 rows per box
    auto try_fit = [&](int tf, int ns, int bs, int budget) -> bool {
        if (want_skew && tf != 32) return false;
        if (vt && !want_skew && tf != 16) return false;          // half-precision lock-step instances exist for 16-frame tiles
        if (f_tf && tf != f_tf) return false;
        if (f_ns && ns != f_ns) return false;
        if (f_bits >= 0 && bs != f_bits) return false;
        if (NC > 1 && bs != 0) return false;                   // the walker (CTA 0) reads every CTA's bits: L2 slot
        const SmemLayout L = make_layout(NW, ns, R, tf, bs, nblk, want_dur, want_skew, NC, vt ? 2 : 4, lag);
        if ((int64_t)L.total > budget) return false;
        best_tf = tf; best_ns = ns; best_bits = bs;
        return true;
    };
    if (latency) {
        // deepest ring that fits one CTA per SM, bits in shared memory when possible
        for (int pass = 0; pass < 4 && !best_tf; ++pass) {
            if (pass == 2) { if (f_skew >= 0 || !want_skew || lag == kLag4) break; want_skew = 0; }    // does not fit skewed: lock-step
            const int p2 = pass & 1;
            for (int ti = 0; ti < 3 && !best_tf; ++ti)
                for (int bs = 1; bs >= 0 && !best_tf; --bs)
                    for (int ns = 8; ns >= (p2 == 0 ? 3 : 2) && !best_tf; --ns) try_fit(tfs[ti], ns, bs, di.smem_optin);
        }
    } else {
        const int reg_occ = NW <= 4 ? 8 / (NW == 3 ? 4 : NW) : 1;          // 128-register instances: 512 compute+loader threads... 8/NW CTAs
        for (int occ = reg_occ; occ >= 1 && !best_tf; --occ) {
            const int budget = per_sm / occ - 1024;
            for (int ti = 0; ti < 3 && !best_tf; ++ti)
                for (int ns = 3; ns >= 2 && !best_tf; --ns)
                    for (int bs = 1; bs >= 0 && !best_tf; --bs) try_fit(tfs[ti], ns, bs, budget);
        }
    }
    if (!best_tf)
        return fail(ALB200_E_UNSUPPORTED, "t_x=%s%lld does not fit the shared-memory ring (max about 3300)", "", tx);
    c->R = R; c->TF = best_tf; c->NW = NW; c->NS = best_ns; c->bits_smem = best_bits; c->nc = NC;
    c->skew = want_skew;
    c->lag = lag;
    c->fn = nullptr;
    if (!latency && !c->skew)
        for (const KEntry& k : g_kernels)
            if (k.R == R && k.TF == best_tf && k.skew == 0 && NW <= k.nwmax && k.minb == 2 && k.vt == vt) { c->fn = k.fn; break; }
    if (vl && !c->skew) return fail(ALB200_E_UNSUPPORTED, "the [t_mel, t_text] layout needs the skewed/TMA form (latency regime, <= 4 rows per lane, 16-byte aligned rows)%s", "");
    if (!c->fn) c->fn = find_kernel(R, best_tf, vl ? 3 : (NC > 1 ? 2 : (lag == kLag4 ? 4 : c->skew)), NW, latency ? 1 : 2, vt);
    if (!c->fn) return fail(ALB200_E_UNSUPPORTED, "no kernel instance for R=%s%lld TF=%lld (score type %lld)", "", R, best_tf, vt);
    SmemLayout L = make_layout(NW, best_ns, R, best_tf, best_bits, nblk, want_dur, want_skew, NC, vt ? 2 : 4, lag);
    c->smem = L.total;
    c->bits_slot_words = (int64_t)nblk * NC * NW * 32 * R;
    ALB_CUDA(cudaFuncSetAttribute(c->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, di.smem_optin));   // once per instance, never lowered
    int occ = 0;
    ALB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, c->fn, 2 * NW * 32, c->smem));
    if (occ < 1) return fail(ALB200_E_UNSUPPORTED, "kernel does not fit on an SM (smem %s%lld bytes)", "", c->smem);
    if (!latency && occ > 8) occ = 8;
    c->occ = occ;
    return 0;
}

// select_config_uncached costs a few driver calls; remember the answers, failures included (per thread; 64 entries hold the
// dur x aligned x dtype probes of alb200_mas_workspace_bytes for several shapes at once).
struct CfgKey { int dev, latency, tx, ty, dur, aligned, bclass, vt, vl; unsigned gen; };
constexpr int kCfgCache = 64;
static int select_config(const DevInfo& di, int b, int tx, int ty, bool want_dur, bool aligned, Config* c, int vt = 0, int vl = 0)
{
    static thread_local CfgKey keys[kCfgCache];
    static thread_local Config vals[kCfgCache];
    static thread_local int rcs[kCfgCache];
    static thread_local char errs[kCfgCache][160];
    static thread_local int used = 0, next = 0;
    CfgKey k;
    memset(&k, 0, sizeof(k));
    k.dev = di.dev; k.latency = is_latency(di, b, tx, ty, aligned, vt); k.tx = tx; k.ty = ty; k.dur = want_dur; k.aligned = aligned; k.vt = vt; k.vl = vl;
    k.bclass = (k.latency && tx > 512) ? b : 0;                 // the cluster decision depends on how many clusters fit
    k.gen = opts().gen;
    int hit = -1;
    for (int i = 0; i < used; ++i)
        if (memcmp(&keys[i], &k, sizeof(k)) == 0) { hit = i; break; }
    if (hit < 0) {
        Config fresh;
        memset(&fresh, 0, sizeof(fresh));
        const int rc = select_config_uncached(di, b, tx, ty, want_dur, aligned, vt, vl, &fresh);
        if (rc == ALB200_E_CUDA) return rc;                      // a failing driver call is not a property of the shape
        hit = next; next = (next + 1) % kCfgCache; if (used < kCfgCache) ++used;
        keys[hit] = k; vals[hit] = fresh; rcs[hit] = rc;
        strncpy(errs[hit], g_err, sizeof(errs[hit]) - 1); errs[hit][sizeof(errs[hit]) - 1] = 0;
    }
    if (rcs[hit]) { snprintf(g_err, sizeof(g_err), "%s", errs[hit]); return rcs[hit]; }
    *c = vals[hit];
    int64_t g = (int64_t)di.sms * c->occ;
    c->grid = c->nc > 1 ? b * c->nc : (int)(b < g ? b : g);
    return 0;
}

// Tensor map over values viewed as [b * t_x rows, t_y frames] fp32 with a (box_rows x box_frames) box; the skewed form's
// loader fetches one box per tile (cp.async.bulk.tensor).  Encoding is host-side arithmetic; the last few are remembered.
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmapEncodeFn tmap_encode_fn()          // also used by neg_cent_v2.cu; nullptr when the driver lacks the entry point
{
    static TmapEncodeFn enc = nullptr;
    if (!enc) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && fn && q == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<TmapEncodeFn>(fn);
    }
    return enc;
}
static int values_tensor_map(const void* values, int vt, int vl, int b, int tx, int ty, int box_rows, int box_frames, CUtensorMap* out)
{
    TmapEncodeFn enc = tmap_encode_fn();
    if (!enc) return fail(ALB200_E_CUDA, "cuTensorMapEncodeTiled is not available in this driver%s", "");
    struct Key { const void* v; int b, tx, ty, br, bf, vt, vl; unsigned gen; };
    static thread_local Key keys[16];
    static thread_local CUtensorMap maps[16];
    static thread_local int used = 0, next = 0;
    const Key k = { values, b, tx, ty, box_rows, box_frames, vt, vl, opts().gen };
    for (int i = 0; i < used; ++i)
        if (keys[i].v == k.v && keys[i].gen == k.gen && keys[i].b == b && keys[i].tx == tx && keys[i].ty == ty && keys[i].br == box_rows && keys[i].bf == box_frames && keys[i].vt == vt && keys[i].vl == vl) {
            *out = maps[i];
            return 0;
        }
    if ((int64_t)b * tx > 0x7fffffffLL) return fail(ALB200_E_UNSUPPORTED, "b * t_x = %s%lld rows exceed the tensor-map coordinate range", "", (long long)b * tx);
    if ((int64_t)b * ty > 0x7fffffffLL) return fail(ALB200_E_UNSUPPORTED, "b * t_y = %s%lld rows exceed the tensor-map coordinate range", "", (long long)b * ty);
    // [b*t_x, t_y] with a (tokens x frames) box -- or, VITS layout, [b*t_y, t_x] with the same box seen the other way round
    cuuint64_t dims[2] = { (cuuint64_t)(vl ? tx : ty), (cuuint64_t)b * (cuuint64_t)(vl ? ty : tx) };
    cuuint64_t strides[1] = { (cuuint64_t)(vl ? tx : ty) * (vt ? 2 : 4) };
    cuuint32_t box[2] = { (cuuint32_t)(vl ? box_rows : box_frames), (cuuint32_t)(vl ? box_frames : box_rows) }, es[2] = { 1, 1 };
    const CUtensorMapDataType dt = vt == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (vt == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    if (opts().tmap_promo >= 0) {                           // tuning aid: 0 none, 1 64 B, 2 128 B, 3 256 B
        const int v = opts().tmap_promo;
        promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : v == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    }
    CUresult r = enc(out, dt, 2, const_cast<void*>(values), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ALB200_E_CUDA, "cuTensorMapEncodeTiled failed with %s%lld", "", (long long)r);
    keys[next] = k; maps[next] = *out;
    next = (next + 1) % 16; if (used < 16) ++used;
    return 0;
}

// Pre-skewed boxes for the 4-frame-lag form: values viewed as (frame, lane, row) with lane stride R*t_y*4 - 16 bytes -- one lane
// further is R rows down and four frames back.  Box = (32 frames, 32 lanes, R rows), 128-byte swizzle; `lanes` bounds the lane
// coordinate (lanes past it are zero-filled without a fetch) and t_y + 4*(lanes-1) bounds the frame coordinate, so the last lane
// never reads past its row's end (tools/micro/tma_lag4_test.cu checks both under compute-sanitizer).
static int values_tensor_map_lag4(const void* values, int b, int tx, int ty, int R, int lanes, CUtensorMap* out)
{
    TmapEncodeFn enc = tmap_encode_fn();
    if (!enc) return fail(ALB200_E_CUDA, "cuTensorMapEncodeTiled is not available in this driver%s", "");
    struct Key { const void* v; int b, tx, ty, R, lanes; unsigned gen; };
    static thread_local Key keys[16];
    static thread_local CUtensorMap maps[16];
    static thread_local int used = 0, next = 0;
    for (int i = 0; i < used; ++i)
        if (keys[i].v == values && keys[i].gen == opts().gen && keys[i].b == b && keys[i].tx == tx && keys[i].ty == ty && keys[i].R == R && keys[i].lanes == lanes) {
            *out = maps[i];
            return 0;
        }
    if ((int64_t)b * tx > 0x7fffffffLL) return fail(ALB200_E_UNSUPPORTED, "b * t_x = %s%lld rows exceed the tensor-map coordinate range", "", (long long)b * tx);
    if (lanes < 1 || lanes > 32) return fail(ALB200_E_INVALID, "bad lane count %s%lld for the pre-skewed tensor map", "", lanes);
    cuuint64_t dims[3] = { (cuuint64_t)ty + (cuuint64_t)kLag4 * (lanes - 1), (cuuint64_t)lanes, (cuuint64_t)b * (cuuint64_t)tx };
    cuuint64_t strides[2] = { (cuuint64_t)R * ty * 4 - 16, (cuuint64_t)ty * 4 };
    cuuint32_t box[3] = { 32, 32, (cuuint32_t)R }, es[3] = { 1, 1, 1 };
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(values), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ALB200_E_CUDA, "cuTensorMapEncodeTiled (pre-skewed box) failed with %s%lld", "", (long long)r);
    keys[next] = Key{ values, b, tx, ty, R, lanes, opts().gen }; maps[next] = *out;
    next = (next + 1) % 16; if (used < 16) ++used;
    return 0;
}

static size_t ws_bytes_for(const Config& c)
{
    return sizeof(WsHeader) + (c.bits_smem ? 0 : (size_t)(c.grid / c.nc) * c.bits_slot_words * 4);
}

static int launch_mas(const void* values, const int32_t* t_xs, const int32_t* t_ys, const void* mask, int mask_dtype,
                      int64_t msb, int64_t msx, int64_t msy, void* paths, int esize, uint64_t one, int zero_fill,
                      int32_t* frame_tok, int32_t* durations, int32_t* lens_out, int b, int tx, int ty, float neg,
                      void* workspace, size_t workspace_bytes, cudaStream_t stream, int vt = 0, int vl = 0,
                      const int* tile_ready = nullptr, int ready_epoch = 0, int ready_tiles = 0, int pdl_wait = 0)
{
    if (!values || b < 0 || tx <= 0 || ty <= 0) return fail(ALB200_E_INVALID, "null values or non-positive shape%s", "");
    if (!mask && (!t_xs || !t_ys)) return fail(ALB200_E_INVALID, "need lengths or a mask%s", "");
    if (paths && esize != 1 && esize != 2 && esize != 4 && esize != 8)
        return fail(ALB200_E_INVALID, "path element size %s%lld not in {1,2,4,8}", "", esize);
    if (mask && (mask_dtype < 0 || mask_dtype > ALB200_I64)) return fail(ALB200_E_INVALID, "unknown mask dtype %s%lld", "", mask_dtype);
    if (!workspace) return fail(ALB200_E_INVALID, "null workspace%s", "");
    if (b == 0) return 0;
    DevInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    Config c;
    if (vt < 0 || vt > 2) return fail(ALB200_E_INVALID, "unknown score dtype %s%lld", "", vt);
    bool aligned = (reinterpret_cast<uintptr_t>(values) & 15) == 0 && ((int64_t)(vl ? tx : ty) * (vt ? 2 : 4)) % 16 == 0;
    if (opts().force_unaligned) aligned = false;
    rc = select_config(di, b, tx, ty, durations != nullptr, aligned, &c, vt, vl);
    if (rc) return rc;
    if (workspace_bytes < ws_bytes_for(c))
        return fail(ALB200_E_INVALID, "workspace too small: %s%lld < %lld bytes", "", (long long)workspace_bytes, (long long)ws_bytes_for(c));
    MasParams p;
    memset(&p, 0, sizeof(p));
    p.values = values; p.paths = paths; p.t_xs = t_xs; p.t_ys = t_ys;
    p.mask = mask; p.msb = msb; p.msx = msx; p.msy = msy; p.mask_dtype = mask_dtype;
    p.frame_tok = frame_tok; p.durations = durations; p.lens_out = lens_out;
    p.ws = reinterpret_cast<WsHeader*>(workspace);
    p.bits_ws = c.bits_smem ? nullptr : reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(workspace) + sizeof(WsHeader));
    p.one = one; p.bits_slot_words = c.bits_slot_words;
    p.B = b; p.Tx = tx; p.Ty = ty; p.esize = esize; p.zero_fill = zero_fill;
    p.nw = c.NW; p.ns = c.NS; p.nblk = (ty + 31) / 32; p.nc = c.nc;
    {
        const SmemLayout L = make_layout(c.NW, c.NS, c.R, c.TF, c.bits_smem, p.nblk, durations != nullptr, c.skew, c.nc, vt ? 2 : 4, c.lag);
        p.off_full = L.off_full; p.off_empty = L.off_empty; p.off_xbar = L.off_xbar; p.off_flags = L.off_flags; p.off_misc = L.off_misc; p.off_bnd = L.off_bnd;
        p.off_zero = L.off_zero; p.off_ring = L.off_ring; p.off_bits = L.off_bits; p.off_dur = L.off_dur; p.off_bt = L.off_bt; p.stage_bytes = L.stage_bytes;
    }
    p.aligned = aligned ? 1 : 0;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    CUtensorMap tmap_tail;
    memset(&tmap_tail, 0, sizeof(tmap_tail));
    p.tail_rows = 32 * c.R;
    if (c.lag == kLag4) {
        // full warps: 32 lanes, frames may run 124 past a row's end (into the next row, which exists: they are not the last warp);
        // last warp: only the lanes that have rows, and a frame extent that keeps its last lane inside the last row
        const int last_rows = tx - (c.NW - 1) * 32 * c.R;
        rc = values_tensor_map_lag4(values, b, tx, ty, c.R, 32, &tmap);
        if (!rc) rc = values_tensor_map_lag4(values, b, tx, ty, c.R, last_rows / c.R, &tmap_tail);
        if (rc) return rc;
    } else if (c.skew) {
        rc = values_tensor_map(values, vt, vl, b, tx, ty, 32 * c.R, c.TF, &tmap);
        if (rc) return rc;
        // rows the last compute warp of the padded text axis really has, rounded up to 8 (TMA box rows are not free)
        const int last = tx - (c.nc * c.NW - 1) * 32 * c.R;
        if (!vl && last > 0 && last < 32 * c.R && !opts().no_tail_box) {
            p.tail_rows = (last + 7) & ~7;
            if (p.tail_rows < 32 * c.R) {
                rc = values_tensor_map(values, vt, vl, b, tx, ty, p.tail_rows, c.TF, &tmap_tail);
                if (rc) return rc;
            }
        }
    }
    p.neg = neg;
    p.tile_ready = tile_ready; p.ready_epoch = ready_epoch; p.ready_tiles = ready_tiles; p.pdl_wait = pdl_wait;
    // Shared zero fill (mas_kernel.cuh): fewer utterances than SMs, one CTA each, a plain launch, 16-byte aligned output -- pad the
    // grid with filler CTAs for the idle SMs.
    int fillers = 0;
    {
        const int64_t total = (int64_t)b * tx * ty * esize;
        // (A/B on one box, us, own fill / shared: C3 32x300x1500 68.6 / 61.5; C2 64x200x1000 38.3 / 38.9; C1 16x100x800 24.0 / 25.1 -- it pays
        //  when one utterance's output takes an SM longer to zero than its forward pass takes: from about 1 KB per frame)
        if (paths && zero_fill && !opts().no_shared_zero && (int64_t)tx * esize >= 1000 && (int64_t)b * c.nc <= c.grid && c.grid < di.sms && tile_ready == nullptr && !pdl_wait &&
            (reinterpret_cast<uintptr_t>(paths) & 15) == 0 && total % 16 == 0 && total / kZeroChunk < 0x3fffffff) {
            int occ = 0;
            ALB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, c.fn, 2 * c.NW * 32, c.smem));
            if (occ == 1) {                              // (one CTA per SM: the fillers land on the idle SMs, not beside a search CTA)
                p.zero_chunks = (int)((total + kZeroChunk - 1) / kZeroChunk);
                p.search_ctas = c.grid;
                const int want = (p.zero_chunks + 63) / 64;
                fillers = di.sms - c.grid < want ? di.sms - c.grid : want;
                fillers -= fillers % c.nc;                   // cluster launches: whole filler clusters
                if (fillers == 0) { p.zero_chunks = 0; p.search_ctas = 0; }
            }
        }
    }
    static long long* d_dbg = nullptr;
    const bool dbg = kDbgBuild && opts().dbg == 1;
    const size_t dbg_n = (size_t)(c.grid + fillers) * ((2 * kMaxWarps + 2) * 2 + kMaxWarps * 8);
    if (dbg) {   // developer aid: per-warp clock64 stamps of the first item of every CTA, printed to stderr
        if (d_dbg) cudaFree(d_dbg);
        ALB_CUDA(cudaMalloc(&d_dbg, dbg_n * 8));
        ALB_CUDA(cudaMemset(d_dbg, 0, dbg_n * 8));
        p.dbg = d_dbg;
    }
    void* args[] = { &p, &tmap, &tmap_tail };
    if (c.nc > 1 || tile_ready != nullptr || pdl_wait) {
        cudaLaunchConfig_t lc;
        memset(&lc, 0, sizeof(lc));
        lc.gridDim = dim3(c.grid + fillers); lc.blockDim = dim3(2 * c.NW * 32); lc.dynamicSmemBytes = c.smem; lc.stream = stream;
        cudaLaunchAttribute at[2];
        int na = 0;
        if (c.nc > 1) {
            at[na].id = cudaLaunchAttributeClusterDimension;
            at[na].val.clusterDim.x = c.nc; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
            ++na;
        }
        if (tile_ready != nullptr || pdl_wait) {
            // launched programmatically dependent on the score kernel just before on this stream (it signals at its start).
            // Pipelined form: starts beside it on the SMs it left free, the loader warps poll tile_ready, nothing waits for that
            // kernel as a whole.  Back-to-back form (pdl_wait): CTAs start as SMs drain, set up, start the zero fill, and the
            // loaders execute griddepcontrol.wait before the first score load.
            at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[na].val.programmaticStreamSerializationAllowed = 1;
            ++na;
        }
        lc.attrs = at; lc.numAttrs = na;
        ALB_CUDA(cudaLaunchKernelExC(&lc, (const void*)c.fn, args));
    } else {
        ALB_CUDA(cudaLaunchKernel((const void*)c.fn, dim3(c.grid + fillers), dim3(2 * c.NW * 32), args, c.smem, stream));
    }
    ++g_launches;
    if (dbg) {
        long long* h = (long long*)malloc(dbg_n * 8);
        ALB_CUDA(cudaMemcpy(h, d_dbg, dbg_n * 8, cudaMemcpyDeviceToHost));
        for (int cta = 0; cta < c.grid && cta < (c.nc > 1 ? c.nc : 2); ++cta) {
            long long* d = h + (size_t)cta * (2 * kMaxWarps + 2) * 2;
            const long long t0 = d[2 * kMaxWarps * 2];
            fprintf(stderr, "[alb200 dbg] cta %d: lengths %lld |", cta, d[0] - t0);
            for (int w = 0; w < c.NW; ++w) fprintf(stderr, " w%d fwd %lld (end@%lld)", w, d[w * 2 + 1] - d[w * 2], d[w * 2 + 1] - t0);
            for (int w = c.NW; w < 2 * c.NW; ++w) fprintf(stderr, " ld%d end@%lld", w - c.NW, d[w * 2 + 1] - t0);
            fprintf(stderr, " | item done @%lld\n", d[2 * kMaxWarps * 2 + 1] - t0);
            {
                long long* e = h + (size_t)c.grid * (2 * kMaxWarps + 2) * 2 + ((size_t)cta * kMaxWarps + (kMaxWarps - 1)) * 4;
                if (e[3] < 0) fprintf(stderr, "[alb200 dbg]   backtrack: %lld blocks; per block: windows+walk %lld, publish %lld cycles\n", -e[3], e[1], e[2]);
            }
            for (int w = 0; w < c.NW && w < kMaxWarps - 1; ++w) {
                long long* e = h + (size_t)c.grid * (2 * kMaxWarps + 2) * 2 + ((size_t)cta * kMaxWarps + w) * 4;
                if (e[3] > 0)
                    fprintf(stderr, "[alb200 dbg]   w%d: %lld units; per unit: wait-full %lld, polls %lld, compute %lld, other %lld cycles\n", w, e[3],
                            e[0] / e[3], e[1] / e[3], e[2] / e[3], ((d[w * 2 + 1] - d[w * 2]) - e[0] - e[1] - e[2]) / e[3]);
            }
            for (int w = 0; w < c.NW; ++w) {
                long long* e = h + (size_t)c.grid * ((2 * kMaxWarps + 2) * 2 + kMaxWarps * 4) + ((size_t)cta * kMaxWarps + w) * 4;
                if (e[3] > 0)
                    fprintf(stderr, "[alb200 dbg]   ld%d: %lld tiles; per tile: wait-empty %lld, issue copies %lld, zero fill %lld cycles\n", w, e[3],
                            e[0] / e[3], e[1] / e[3], e[2] / e[3]);
            }
        }
        free(h);
    }
    return 0;
}

// ------------------------------------------------------------------ host-pointer path
struct HostCtx {
    bool init = false;
    int dev = -1;
    cudaStream_t st[2] = {nullptr, nullptr};
    cudaEvent_t ev[32];
    cudaEvent_t ev_lens = nullptr;
    float* d_values = nullptr;   size_t cap_values = 0;
    int32_t* d_ftok = nullptr;   size_t cap_ftok = 0;
    int32_t* d_lens = nullptr;   size_t cap_lens = 0;
    void* d_ws[2] = {nullptr, nullptr}; size_t cap_ws = 0;
    int32_t* h_ftok = nullptr;   size_t cap_hftok = 0;
    int32_t* h_lens = nullptr;   size_t cap_hlens = 0;
};
static thread_local HostCtx g_ctx;

template <typename T>
static int grow_dev(T** ptr, size_t* cap, size_t need)
{
    if (*cap >= need) return 0;
    if (*ptr) ALB_CUDA(cudaFree(*ptr));
    *ptr = nullptr; *cap = 0;
    ALB_CUDA(cudaMalloc(reinterpret_cast<void**>(ptr), need));
    *cap = need;
    return 0;
}
template <typename T>
static int grow_host(T** ptr, size_t* cap, size_t need)
{
    if (*cap >= need) return 0;
    if (*ptr) ALB_CUDA(cudaFreeHost(*ptr));
    *ptr = nullptr; *cap = 0;
    ALB_CUDA(cudaMallocHost(reinterpret_cast<void**>(ptr), need));
    *cap = need;
    return 0;
}

}  // namespace alb

using namespace alb;

extern "C" {

const char* alb200_last_error(void) { return g_err; }
const char* alb200_version(void) { return "aligner_b200 0.2.0 sm_100a"; }
int alb200_set_option(const char* name, const char* value)
{
    if (!name) return fail(ALB200_E_INVALID, "null option name%s", "");
    opts();                                                   // the environment is read first, once
    std::lock_guard<std::mutex> lk(g_opts_mu);
    if (set_opt_locked(name, value)) return fail(ALB200_E_INVALID, "unknown option '%s'", name);
    return 0;
}
uint64_t alb200_launch_count(void) { return g_launches; }
void alb200_last_transfer_bytes(uint64_t* h2d, uint64_t* d2h) { if (h2d) *h2d = g_h2d; if (d2h) *d2h = g_d2h; }

int alb200_mas_device(const float* values, const int32_t* t_xs, const int32_t* t_ys, void* paths, int path_elem_size,
                      uint64_t path_one, int zero_fill, int32_t* frame_tok, int32_t* durations, int b, int tx, int ty,
                      float max_neg_val, void* workspace, size_t workspace_bytes, void* stream)
{
    return launch_mas(values, t_xs, t_ys, nullptr, 0, 0, 0, 0, paths, path_elem_size, path_one, zero_fill, frame_tok,
                      durations, nullptr, b, tx, ty, max_neg_val, workspace, workspace_bytes, (cudaStream_t)stream);
}

int alb200_mas_device_ex(const void* values, int value_dtype, const int32_t* t_xs, const int32_t* t_ys, const void* mask, int mask_dtype,
                         int64_t msb, int64_t msx, int64_t msy, void* paths, int path_elem_size, uint64_t path_one,
                         int zero_fill, int32_t* frame_tok, int32_t* durations, int32_t* lens_out, int b, int tx, int ty, float max_neg_val,
                         void* workspace, size_t workspace_bytes, void* stream)
{
    const int vl = (value_dtype & ALB200_LAYOUT_VITS) ? 1 : 0;
    value_dtype &= ~ALB200_LAYOUT_VITS;
    int vt;
    switch (value_dtype) {
        case ALB200_F32: vt = 0; break;
        case ALB200_F16: vt = 1; break;
        case ALB200_BF16: vt = 2; break;
        default: return fail(ALB200_E_INVALID, "score dtype %s%lld is not fp32, fp16 or bf16", "", value_dtype);
    }
    return launch_mas(values, t_xs, t_ys, mask, mask_dtype, msb, msx, msy, paths, path_elem_size, path_one, zero_fill, frame_tok,
                      durations, lens_out, b, tx, ty, max_neg_val, workspace, workspace_bytes, (cudaStream_t)stream, vt, vl);
}

int alb200_mas_device_masked(const float* values, const void* mask, int mask_dtype, int64_t msb, int64_t msx, int64_t msy,
                             void* paths, int path_elem_size, uint64_t path_one, int zero_fill, int32_t* frame_tok,
                             int32_t* durations, int32_t* lens_out, int b, int tx, int ty, float max_neg_val,
                             void* workspace, size_t workspace_bytes, void* stream)
{
    if (!mask) return fail(ALB200_E_INVALID, "null mask%s", "");
    return launch_mas(values, nullptr, nullptr, mask, mask_dtype, msb, msx, msy, paths, path_elem_size, path_one, zero_fill,
                      frame_tok, durations, lens_out, b, tx, ty, max_neg_val, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t alb200_mas_workspace_bytes(int b, int tx, int ty)
{
    DevInfo di;
    if (b <= 0 || tx <= 0 || ty <= 0 || device_info(&di)) return sizeof(WsHeader);
    size_t need = sizeof(WsHeader);
    for (int v = 0; v < 4; ++v) {
        Config c;
        for (int vt = 0; vt < 3; ++vt)
            if (select_config(di, b, tx, ty, (v & 1) != 0, (v & 2) != 0, &c, vt) == 0) need = std::max(need, ws_bytes_for(c));
    }
    return need;
}

int alb200_mas_describe(int b, int tx, int ty, int want_durations, char* buf, size_t buf_bytes)
{
    DevInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    Config c;
    rc = select_config(di, b, tx, ty, want_durations != 0, (ty & 3) == 0, &c);
    if (rc) return rc;
    if (buf && buf_bytes)
        snprintf(buf, buf_bytes, "rows_per_lane=%d tile_frames=%d warps=%d stages=%d bits=%s form=%s smem=%u grid=%d ctas_per_sm=%d cluster=%d",
                 c.R, c.TF, c.NW, c.NS, c.bits_smem ? "smem" : "global", c.lag == kLag4 ? "skewed4" : (c.skew ? "skewed" : "lockstep"), c.smem, c.grid, c.occ, c.nc);
    return 0;
}

// ---- fused score + search, pipelined form
extern "C" size_t alb200_neg_cent_workspace_bytes(int mode, int b, int c, int tx, int ty);
extern "C" int alb200_neg_cent_gaussian_ws(const float*, const float*, const float*, float*, int, int, int, int, void*, size_t, void*);
extern "C" int alb200_neg_cent_gaussian_v2_pipelined(const float*, const float*, const float*, float*, int, int, int, int, void*, size_t, void*, int*, int, int);

size_t alb200_fused_workspace_bytes(int b, int c, int tx, int ty)
{
    const size_t nc = alb200_neg_cent_workspace_bytes(0, b, c, tx, ty);
    const size_t flags = ((size_t)b * ((ty + 127) / 128) * 4 + 1024 + 255) & ~(size_t)255;     // + slack for the developer time stamps
    return ((nc + 255) & ~(size_t)255) + flags + ((alb200_mas_workspace_bytes(b, tx, ty) + 255) & ~(size_t)255);
}

int alb200_gaussian_mas_fused(const float* z, const float* m_p, const float* logs_p, float* neg_cent, const int32_t* t_xs, const int32_t* t_ys,
                              const void* mask, int mask_dtype, int64_t msb, int64_t msx, int64_t msy, void* paths, int path_elem_size,
                              uint64_t path_one, int zero_fill, int32_t* frame_tok, int32_t* durations, int b, int c, int tx, int ty,
                              float max_neg_val, void* workspace, size_t workspace_bytes, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!z || !m_p || !logs_p || !neg_cent || !workspace || b < 0 || c <= 0 || tx <= 0 || ty <= 0)
        return fail(ALB200_E_INVALID, "null pointer or non-positive shape%s", "");
    if (b == 0) return 0;
    if (workspace_bytes < alb200_fused_workspace_bytes(b, c, tx, ty)) return fail(ALB200_E_INVALID, "workspace too small (alb200_fused_workspace_bytes)%s", "");
    const size_t nc_bytes = (alb200_neg_cent_workspace_bytes(0, b, c, tx, ty) + 255) & ~(size_t)255;
    const int n_mt = (ty + 127) / 128;
    const size_t flag_bytes = ((size_t)b * n_mt * 4 + 1024 + 255) & ~(size_t)255;
    // Layout: [search workspace][score scratch][tile flags].  The search's 64-byte header comes FIRST so that it stays where it is
    // from call to call whatever the shapes: its counters re-arm themselves and rely on the zero-initialised buffer, which a header
    // at a shape-dependent offset (inside what an earlier call used as score scratch) would not have.
    const size_t mas_ws_bytes = (alb200_mas_workspace_bytes(b, tx, ty) + 255) & ~(size_t)255;
    void* mas_ws = workspace;
    char* ws = reinterpret_cast<char*>(workspace) + mas_ws_bytes;       // score scratch
    int* flags = reinterpret_cast<int*>(ws + nc_bytes);
    DevInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    // Pipelined when the search (one CTA or cluster per utterance) needs at most a third of the SMs: the score kernel runs on the
    // others in tile-major order and publishes every 128-frame tile; the search is launched programmatically dependent, starts
    // beside it and its loader warps wait per tile, so a tile is read out of L2 while later ones are still being computed.
    // Measured (profiles/r02_notes.md): 32 x 300 x 1500 145 -> 117 us.  With more utterances the SMs taken away from the score
    // kernel cost as much as the overlap gains (64 x 200 x 1000: 94 vs 93 us), and a batch that fills the machine is
    // bandwidth-bound with nothing to overlap: those run the two kernels back to back.
    Config cfg;
    const bool aligned = (reinterpret_cast<uintptr_t>(neg_cent) & 15) == 0 && ((int64_t)ty * 4) % 16 == 0 && !opts().force_unaligned;
    rc = select_config(di, b, tx, ty, durations != nullptr, aligned, &cfg, 0, 0);
    if (rc) return rc;
    const int free_sms = di.sms - cfg.grid;
    const int fmode = opts().fused_seq;                       // tuning: 1 = always back to back, 2 = pipelined whenever possible
    const bool can_pipe = nc_bytes > 0 && is_latency(di, b, tx, ty, aligned, 0) && cfg.grid <= di.sms && free_sms >= 32 && !opts().nc_ffma && !opts().nc_v1;
    // ... and only when the score kernel has at least two rounds of tiles on the SMs left to it: with a single round every tile is
    // finished at about the same time and there is nothing to overlap (16 x 100 x 800: 47 us pipelined, 45 back to back)
    const bool pipelined = can_pipe && fmode != 1 && (fmode == 2 || (cfg.grid * 3 <= di.sms && (int64_t)b * n_mt >= 2 * (int64_t)free_sms));
    if (pipelined) {
        ALB_CUDA(cudaMemsetAsync(flags, 0, (size_t)b * n_mt * 4, stream));
        if (kDbgBuild && opts().dbg == 2) {      // developer aid: global-timer stamps of both kernels (see the kernels)
            unsigned long long init[4] = { ~0ull, 0ull, ~0ull, 0ull };
            ALB_CUDA(cudaMemcpyAsync(flags + (size_t)b * n_mt + 64, init, sizeof(init), cudaMemcpyHostToDevice, stream));
        }
        rc = alb200_neg_cent_gaussian_v2_pipelined(z, m_p, logs_p, neg_cent, b, c, tx, ty, ws, nc_bytes, stream, flags, 1, free_sms);
        if (rc == 0) {
            rc = launch_mas(neg_cent, t_xs, t_ys, mask, mask_dtype, msb, msx, msy, paths, path_elem_size, path_one, zero_fill, frame_tok, durations,
                            nullptr, b, tx, ty, max_neg_val, mas_ws, mas_ws_bytes, stream, 0, 0, flags, 1, n_mt);
            if (kDbgBuild && opts().dbg == 2 && rc == 0) {
                unsigned long long h[4];
                cudaStreamSynchronize(stream);
                cudaMemcpy(h, flags + (size_t)b * n_mt + 64, sizeof(h), cudaMemcpyDeviceToHost);
                fprintf(stderr, "[fused dbg] score kernel %.1f us (start 0), search starts at %.1f us, ends at %.1f us\n", (h[1] - h[0]) * 1e-3,
                        ((double)h[2] - (double)h[0]) * 1e-3, ((double)h[3] - (double)h[0]) * 1e-3);
            }
            return rc;
        }
        if (rc != ALB200_E_UNSUPPORTED) return rc;
    }
    rc = alb200_neg_cent_gaussian_ws(z, m_p, logs_p, neg_cent, b, c, tx, ty, nc_bytes ? ws : nullptr, nc_bytes, stream);
    if (rc) return rc;
    // back to back, but still programmatically dependent when the TMA / tcgen05 score kernel ran (it signals its dependents):
    // the search's launch, set-up and the start of its zero fill hide under the score kernel's tail
    const int pdl = (nc_bytes > 0 && !opts().nc_ffma && !opts().nc_v1 && !opts().nc_no_pdl) ? 1 : 0;
    return launch_mas(neg_cent, t_xs, t_ys, mask, mask_dtype, msb, msx, msy, paths, path_elem_size, path_one, zero_fill, frame_tok, durations,
                      nullptr, b, tx, ty, max_neg_val, mas_ws, mas_ws_bytes, stream, 0, 0, nullptr, 0, 0, pdl);
}

int alb200_mas_status(void* workspace, void* stream)
{
    if (!workspace) return fail(ALB200_E_INVALID, "null workspace%s", "");
    int st = 0;
    WsHeader* h = reinterpret_cast<WsHeader*>(workspace);
    ALB_CUDA(cudaMemcpyAsync(&st, &h->status, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    ALB_CUDA(cudaMemsetAsync(&h->status, 0, sizeof(int), (cudaStream_t)stream));
    ALB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return st;
}

int alb200_maximum_path_c(int32_t* paths, const float* values, const int32_t* t_xs, const int32_t* t_ys, int b, int tx,
                          int ty, float max_neg_val)
{
    if (!paths || !values || !t_xs || !t_ys || b < 0 || tx <= 0 || ty <= 0)
        return fail(ALB200_E_INVALID, "null pointer or non-positive shape%s", "");
    g_h2d = g_d2h = 0;
    if (b == 0) return 0;
    for (int i = 0; i < b; ++i) {
        const int a = t_xs[i], c = t_ys[i];
        if (a > 0 && c > 0 && (a > c || a > tx || c > ty))
            return fail(ALB200_E_LENGTHS, "item %s%lld has t_x=%lld, t_y=%lld (need t_x <= t_y, inside the tensor)", "", i, a, c);
    }
    DevInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    HostCtx& X = g_ctx;
    if (X.init && X.dev != di.dev) return fail(ALB200_E_INVALID, "host context was created on another device%s", "");
    if (!X.init) {
        for (int s = 0; s < 2; ++s) ALB_CUDA(cudaStreamCreateWithFlags(&X.st[s], cudaStreamNonBlocking));
        for (int i = 0; i < 32; ++i) ALB_CUDA(cudaEventCreateWithFlags(&X.ev[i], cudaEventDisableTiming));
        ALB_CUDA(cudaEventCreateWithFlags(&X.ev_lens, cudaEventDisableTiming));
        X.dev = di.dev; X.init = true;
    }
    const size_t item_bytes = (size_t)tx * ty * 4;
    // chunks of about 8 MB so copies, kernels and the host scatter overlap
    int nch = (int)std::min<size_t>((size_t)b, std::max<size_t>(1, (item_bytes * b) / (8u << 20)));
    nch = std::min(nch, 32);
    const int per = (b + nch - 1) / nch;
    nch = (b + per - 1) / per;

    if ((rc = grow_dev(&X.d_values, &X.cap_values, item_bytes * b))) return rc;
    if ((rc = grow_dev(&X.d_ftok, &X.cap_ftok, (size_t)b * ty * 4))) return rc;
    if ((rc = grow_dev(&X.d_lens, &X.cap_lens, (size_t)b * 8))) return rc;
    if ((rc = grow_host(&X.h_ftok, &X.cap_hftok, (size_t)b * ty * 4))) return rc;
    if ((rc = grow_host(&X.h_lens, &X.cap_hlens, (size_t)b * 8))) return rc;
    // the last chunk can be smaller and pick another configuration (e.g. cluster mode with its bits in L2): size for both
    const int last_nb = b - (nch - 1) * per;
    const size_t wsz = std::max(alb200_mas_workspace_bytes(per, tx, ty), alb200_mas_workspace_bytes(last_nb, tx, ty));
    if (X.cap_ws < wsz) {
        for (int s = 0; s < 2; ++s) {
            if (X.d_ws[s]) ALB_CUDA(cudaFree(X.d_ws[s]));
            X.d_ws[s] = nullptr;
            ALB_CUDA(cudaMalloc(&X.d_ws[s], wsz));
            ALB_CUDA(cudaMemset(X.d_ws[s], 0, wsz));
        }
        X.cap_ws = wsz;
    }
    memcpy(X.h_lens, t_xs, (size_t)b * 4);
    memcpy(X.h_lens + b, t_ys, (size_t)b * 4);
    ALB_CUDA(cudaMemcpyAsync(X.d_lens, X.h_lens, (size_t)b * 8, cudaMemcpyHostToDevice, X.st[0]));
    ALB_CUDA(cudaEventRecord(X.ev_lens, X.st[0]));
    ALB_CUDA(cudaStreamWaitEvent(X.st[1], X.ev_lens, 0));
    g_h2d += (uint64_t)b * 8;

    for (int c = 0; c < nch; ++c) {
        const int b0 = c * per, nb = std::min(per, b - b0);
        cudaStream_t s = X.st[c & 1];
        // Ship only what the search can read (core.pyx:18): row x of an item is live on frames [x, x + t_y - t_x], a parallelogram
        // whose rows start (t_y_pad + 1) elements apart -- exactly one pitched 2-D copy per item (pitch = t_y_pad + 1 elements,
        // width = t_y - t_x + 1, height = t_x).  Cells outside the band are never copied; the kernel may load them into shared
        // memory with whole tiles but no in-band cell depends on them ("lower band edge is optional").  When the band
        // holds most of the matrix one copy per chunk, trimmed to the longest item's rows, is faster: measured on the C2 shape (band =
        // 80 % of the cells) 64 pitched copies with 3.2 KB rows 1.137 ms against 1.076 ms for the plain chunk copies -- the DMA
        // engine pays per row.  The per-item form is taken when it saves at least 40 % of the bytes.
        int mx = 0;
        uint64_t band = 0;
        for (int i = b0; i < b0 + nb; ++i)
            if (t_xs[i] > 0 && t_ys[i] > 0) { mx = std::max(mx, t_xs[i]); band += (uint64_t)t_xs[i] * (uint64_t)(t_ys[i] - t_xs[i] + 1) * 4; }
        const uint64_t whole = (uint64_t)mx * ty * 4 * nb;
        if (mx > 0 && band * 10 <= whole * 6) {
            for (int i = b0; i < b0 + nb; ++i) {
                if (t_xs[i] <= 0 || t_ys[i] <= 0) continue;
                const size_t pitch = (size_t)(ty + 1) * 4, width = (size_t)(t_ys[i] - t_xs[i] + 1) * 4;
                ALB_CUDA(cudaMemcpy2DAsync(X.d_values + (size_t)i * tx * ty, pitch, values + (size_t)i * tx * ty, pitch, width, (size_t)t_xs[i],
                                           cudaMemcpyHostToDevice, s));
            }
            g_h2d += band;
        } else if (mx > 0) {   // rows past the longest item of the chunk are never read: do not ship them
            ALB_CUDA(cudaMemcpy2DAsync(X.d_values + (size_t)b0 * tx * ty, item_bytes, values + (size_t)b0 * tx * ty, item_bytes,
                                       (size_t)mx * ty * 4, nb, cudaMemcpyHostToDevice, s));
            g_h2d += whole;
        }
        rc = launch_mas(X.d_values + (size_t)b0 * tx * ty, X.d_lens + b0, X.d_lens + b + b0, nullptr, 0, 0, 0, 0, nullptr, 4, 1, 0,
                        X.d_ftok + (size_t)b0 * ty, nullptr, nullptr, nb, tx, ty, max_neg_val, X.d_ws[c & 1], X.cap_ws, s);
        if (rc) return rc;
        ALB_CUDA(cudaMemcpyAsync(X.h_ftok + (size_t)b0 * ty, X.d_ftok + (size_t)b0 * ty, (size_t)nb * ty * 4, cudaMemcpyDeviceToHost, s));
        g_d2h += (uint64_t)nb * ty * 4;
        ALB_CUDA(cudaEventRecord(X.ev[c], s));
    }
    for (int c = 0; c < nch; ++c) {
        const int b0 = c * per, nb = std::min(per, b - b0);
        ALB_CUDA(cudaEventSynchronize(X.ev[c]));
        for (int i = b0; i < b0 + nb; ++i) {                     // path[index, y] = 1   (core.pyx:33)
            const int n = (t_xs[i] > 0 && t_ys[i] > 0) ? t_ys[i] : 0;
            const int32_t* ft = X.h_ftok + (size_t)i * ty;
            int32_t* pi = paths + (size_t)i * tx * ty;
            for (int y = 0; y < n; ++y) pi[(size_t)ft[y] * ty + y] = 1;
        }
    }
    return 0;
}

}  // extern "C"
