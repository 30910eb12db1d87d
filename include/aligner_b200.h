/*
 * aligner_b200.h -- C ABI of libaligner_b200.so (sm_100a only, no CPU fallback).
 *
 * This is the drop-in boundary for the one hot path of xiaozhah/Aligner:
 * monotonic alignment search behind monotonic_align.maximum_path, plus the
 * score matrices that feed it.  Plain pointers and sizes; no torch, numpy or
 * Python types.  Every function returns 0 on success or a negative
 * ALB200_E_* code and never throws; alb200_last_error() describes the last
 * failure on the calling thread.
 *
 * "Reference" citations are relative to the root of xiaozhah/Aligner.
 */
#ifndef ALIGNER_B200_H_
#define ALIGNER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALB200_OK              0
#define ALB200_E_INVALID      -1   /* bad argument (null pointer, non-positive size, unknown dtype) */
#define ALB200_E_UNSUPPORTED  -2   /* shape outside what the kernels cover (t_x too long for the shared-memory ring), or no native
                                      kernel for the requested score dtype / layout at this shape (alb200_mas_device_ex) */
#define ALB200_E_CUDA         -3   /* a CUDA runtime call failed; see alb200_last_error() */
#define ALB200_E_LENGTHS      -4   /* an item has t_x > t_y, or lengths outside the tensor */
#define ALB200_E_NO_DEVICE    -5   /* no sm_100 device: this library has no fallback */

/* element types: of a mask tensor (alb200_mas_device_masked, all of them) and of the scores (alb200_mas_device_ex: F32, F16, BF16) */
#define ALB200_F32  0
#define ALB200_F16  1
#define ALB200_BF16 2
#define ALB200_F64  3
#define ALB200_U8   4   /* also torch.bool */
#define ALB200_I8   5
#define ALB200_I16  6
#define ALB200_I32  7
#define ALB200_I64  8
/* OR into value_dtype of alb200_mas_device_ex: scores AND path are stored [b, t_mel, t_text] (the VITS convention,
 * its core indexes value[y, x]) instead of the reference's [b, t_text, t_mel]; tx / ty keep their meaning. */
#define ALB200_LAYOUT_VITS 0x100

const char *alb200_last_error(void);
/* "aligner_b200 <version> sm_100a" */
const char *alb200_version(void);

/* Tuning / test options.  They are read from the environment once, when the library is first used (ALB200_FORCE,
 * ALB200_LATENCY_MAX_B, ALB200_TMAP_PROMO, ALB200_NO_TAIL_BOX, ALB200_FORCE_UNALIGNED, ALB200_DBG, ALB200_NC_FFMA,
 * ALB200_NC_V1), and can be changed afterwards with this call: name = the variable without the prefix, lower case
 * ("force", "latency_max_b", ...); value NULL or "" restores the default.  No launch path reads the environment.
 * Process-wide; not meant to be changed while another thread is launching. */
int alb200_set_option(const char *name, const char *value);

/* ------------------------------------------------------------------------
 * Monotonic alignment search, device pointers, asynchronous on `stream`.
 *
 * Replaces the body of  maximum_path_c(paths, values, t_xs, t_ys, max_neg_val)
 * (reference monotonic_align/core.pyx:38-45, per-item work core.pyx:7-35)
 * for buffers that already live in HBM.
 *
 *  values     const float [b, tx, ty], C-contiguous, NOT modified (the reference
 *             overwrites it with cumulative scores; nothing downstream reads them).
 *  t_xs,t_ys  const int32 [b] on the device (token / frame counts per item).
 *  paths      [b, tx, ty] of `path_elem_size`-byte elements, or NULL.  A cell on
 *             the path receives the bit pattern `path_one` (e.g. 0x3f800000 for
 *             fp32 1.0, 1 for int32).  If zero_fill != 0 the kernel also writes
 *             every other cell to zero (fused with the forward pass); with
 *             zero_fill == 0 the caller has pre-zeroed it, exactly like the
 *             reference contract (monotonic_align/__init__.py:15).
 *  frame_tok  optional int32 [b, ty]: token index of every frame, -1 past t_y.
 *  durations  optional int32 [b, tx]: frames per token (= path.sum(-1)).
 *  workspace  device scratch of at least alb200_mas_workspace_bytes() bytes that
 *             was ZERO when first handed to this library and is not touched by
 *             anyone else afterwards (it holds a self-resetting work counter,
 *             the status word and spilled direction bits).
 *  stream     cudaStream_t (NULL = legacy default stream).
 *
 * Items with t_x <= 0 or t_y <= 0 produce an all-zero path.  Items with
 * t_x > t_y (reference: out-of-bounds reads) produce an all-zero
 * path and set bit 0 of the status word, readable with alb200_mas_status().
 * ------------------------------------------------------------------------ */
int alb200_mas_device(const float *values, const int32_t *t_xs, const int32_t *t_ys,
                      void *paths, int path_elem_size, uint64_t path_one, int zero_fill,
                      int32_t *frame_tok, int32_t *durations,
                      int b, int tx, int ty, float max_neg_val,
                      void *workspace, size_t workspace_bytes, void *stream);

/* The general device entry: every option of the two entries around it, plus the score element type.
 *   value_dtype  ALB200_F32, ALB200_F16 or ALB200_BF16.  Half-precision scores are promoted to fp32 as they are loaded,
 *                which is exactly the reference's `.astype(np.float32)` (monotonic_align/__init__.py:14): the path is
 *                bit-identical to the reference run on the promoted values, and the kernel reads 2 instead of 4 bytes
 *                per cell.  Native half-precision kernels exist for the shapes the host heuristics
 *                pick (<= 4 rows per lane in the latency regime, the throughput regime); otherwise the call returns
 *                ALB200_E_UNSUPPORTED and the caller promotes to fp32 on the device first (the Python layer does).
 *   lengths      either (t_xs, t_ys) or mask (+ strides in elements), as in alb200_mas_device / _masked.
 *   layout       value_dtype | ALB200_LAYOUT_VITS: scores and path are [b, t_mel, t_text] (mask strides are still given
 *                as (b, text, mel)).  Native for fp32 in the latency regime with <= 4 rows per lane and t_x % 4 == 0;
 *                otherwise ALB200_E_UNSUPPORTED (the Python layer then transposes on the device). */
int alb200_mas_device_ex(const void *values, int value_dtype, const int32_t *t_xs, const int32_t *t_ys,
                         const void *mask, int mask_dtype, int64_t mask_stride_b, int64_t mask_stride_x, int64_t mask_stride_y,
                         void *paths, int path_elem_size, uint64_t path_one, int zero_fill,
                         int32_t *frame_tok, int32_t *durations, int32_t *lens_out,
                         int b, int tx, int ty, float max_neg_val,
                         void *workspace, size_t workspace_bytes, void *stream);

/* Same, with the lengths derived inside the kernel from a [b,tx,ty] mask the way
 * the reference's Python layer does (monotonic_align/__init__.py:18-19):
 *   t_x = sum_x mask[b, x, 0],  t_y = sum_y mask[b, 0, y], truncated to int32.
 * Strides are in ELEMENTS, so expanded / broadcast masks work.  Only those two
 * slices of the mask are read: the kernel relies on the mask being the outer
 * product of two prefix masks, for which `value * mask` (__init__.py:11) is
 * the identity on every cell the search can read.
 * lens_out: optional int32 [2, b] receiving (t_x, t_y). */
int alb200_mas_device_masked(const float *values, const void *mask, int mask_dtype,
                             int64_t mask_stride_b, int64_t mask_stride_x, int64_t mask_stride_y,
                             void *paths, int path_elem_size, uint64_t path_one, int zero_fill,
                             int32_t *frame_tok, int32_t *durations, int32_t *lens_out,
                             int b, int tx, int ty, float max_neg_val,
                             void *workspace, size_t workspace_bytes, void *stream);

/* Scratch size for a [b, tx, ty] problem on the current device. */
size_t alb200_mas_workspace_bytes(int b, int tx, int ty);

/* Writes a one-line description of the kernel configuration chosen for a
 * [b, tx, ty] problem (rows per lane, tile width, ring depth, grid) into buf. */
int alb200_mas_describe(int b, int tx, int ty, int want_durations, char *buf, size_t buf_bytes);

/* Synchronises `stream` and returns the status word accumulated in `workspace`
 * since the last call (then clears it): bit 0 = some item had t_x > t_y or
 * lengths outside [0,tx]x[0,ty].  Negative = ALB200_E_CUDA. */
int alb200_mas_status(void *workspace, void *stream);

/* ------------------------------------------------------------------------
 * Host-pointer entry with the reference's exact calling convention:
 *   maximum_path_c(paths, values, t_xs, t_ys, max_neg_val=-1e9)
 * (reference monotonic_align/core.pyx:40; called from __init__.py:20).
 *
 *  paths   int32 [b,tx,ty] in HOST memory, pre-zeroed by the caller; ones are
 *          written in place (core.pyx:33).
 *  values  const float [b,tx,ty] in HOST memory (pinned memory makes the copy
 *          asynchronous); NOT clobbered.
 * Blocking.  Values are staged to the device in chunks on internal streams, the
 * search runs on the GPU, only the per-frame token indices come back, and the
 * ones are scattered into `paths` on the host.  Returns ALB200_E_LENGTHS if an
 * item has t_x > t_y or lengths outside the tensor (checked on the host before
 * any work).
 * ------------------------------------------------------------------------ */
int alb200_maximum_path_c(int32_t *paths, const float *values, const int32_t *t_xs,
                          const int32_t *t_ys, int b, int tx, int ty, float max_neg_val);

/* Bytes moved by the last alb200_maximum_path_c call on this thread
 * (host->device, device->host); used by bench.py's e2e accounting. */
void alb200_last_transfer_bytes(uint64_t *h2d, uint64_t *d2h);

/* ------------------------------------------------------------------------
 * Score matrices that feed the search (device pointers, asynchronous on `stream`).
 * The reference snapshot has no code for these (its MoBo/RoMo/OTA branches are not in the
 * tree, README.md:9-25); they implement the published formulas of the projects it links to
 * and produce the layout the reference API documents: [b, t_text, t_mel], t_mel contiguous
 * (monotonic_align/__init__.py:8-9).  fp32 in, fp32 out.  Default path: 5th-generation tensor cores (tcgen05) with every
 * operand split into two fp16 parts after an exact power-of-two scaling (three products, fp32 accumulation in tensor
 * memory; a tile whose mel-side values leave the fp16 range is recomputed in plain fp32) -- within 1e-5 of the fp64
 * value relative to the largest |score| of the utterance; the Gaussian score is formed as the contraction plus a per-token
 * term, the OTA distance as |q|^2 + |k|^2 - 2 q.k.  Option "nc_ffma" selects fixed-order fp32 FFMA kernels instead.
 *
 * Gaussian prior (Glow-TTS / VITS `neg_cent1..4`):
 *   out[b,x,y] = sum_c log N(z[b,c,y]; m_p[b,c,x], exp(logs_p[b,c,x])^2)
 *   z [b,c,ty], m_p [b,c,tx], logs_p [b,c,tx], out [b,tx,ty].
 * ------------------------------------------------------------------------ */
int alb200_neg_cent_gaussian(const float *z, const float *m_p, const float *logs_p, float *out,
                             int b, int c, int tx, int ty, void *stream);

/* OTA aligner (arXiv 2108.10447, README.md:50; NeMo AlignmentEncoder):
 *   d[b,x,y]   = -temperature * sum_c (queries[b,c,y] - keys[b,c,x])^2
 *   out[b,x,y] = log_softmax(d, over x) + log(prior[b,x,y] + 1e-8)
 * queries [b,c,ty] (mel side), keys [b,c,tx] (text side), prior optional [b,tx,ty],
 * x_lengths optional int32 [b]: text positions past it are left out of the softmax (-inf). */
int alb200_neg_cent_ota(const float *queries, const float *keys, const float *prior,
                        const int32_t *x_lengths, float *out, float temperature,
                        int b, int c, int tx, int ty, void *stream);

/* The same two score matrices with caller-provided device scratch (what the Python layer calls: the scratch comes from the
 * framework's caching allocator).  The default tensor-core path stages the text-side operand of every utterance once, as
 * scaled fp16 hi / lo parts, in `workspace`; alb200_neg_cent_workspace_bytes(mode, ...) gives its size (mode 0 = Gaussian,
 * 1 = OTA; 0 = this shape takes a path that needs no scratch).  workspace must be 256-byte aligned; NULL selects the
 * paths without scratch.  The entries above allocate the same scratch with cudaMallocAsync / cudaFreeAsync on `stream`. */
size_t alb200_neg_cent_workspace_bytes(int mode, int b, int c, int tx, int ty);
int alb200_neg_cent_gaussian_ws(const float *z, const float *m_p, const float *logs_p, float *out,
                                int b, int c, int tx, int ty, void *workspace, size_t workspace_bytes, void *stream);
int alb200_neg_cent_ota_ws(const float *queries, const float *keys, const float *prior,
                           const int32_t *x_lengths, float *out, float temperature,
                           int b, int c, int tx, int ty, void *workspace, size_t workspace_bytes, void *stream);

/* OTA score with the beta-binomial alignment prior of the OTA paper GENERATED inside the kernel instead of
 * read from a [b, tx, ty] tensor:  prior[b, x, y] = BetaBinom(x; n = t_x[b] - 1, a = s (y + 1), b = s (t_y[b] - y)) for
 * y < t_y[b], 0 beyond (what a zero-padded prior tensor holds); out = log_softmax_x(d) + log(prior + 1e-8).
 * x_lengths / y_lengths: optional int32 [b] (NULL = tx / ty), s = prior_scaling > 0.  Same workspace as
 * alb200_neg_cent_ota_ws.  ALB200_E_UNSUPPORTED for shapes outside the TMA / tcgen05 path (tx > 512, ty % 4 != 0): the
 * caller then materialises the prior and uses alb200_neg_cent_ota_ws (the Python layer does). */
int alb200_neg_cent_ota_bb(const float *queries, const float *keys, const int32_t *x_lengths, const int32_t *y_lengths,
                           float prior_scaling, float *out, float temperature, int b, int c, int tx, int ty,
                           void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------
 * Fused score + search (the seam it removes is monotonic_align/__init__.py:11-14, where the reference
 * materialises the score matrix, stages it and only then consumes it):
 *   neg_cent = gaussian score of (z, m_p, logs_p)         [b, tx, ty], also returned (callers log it / reuse it)
 *   paths    = monotonic alignment search over neg_cent    same contract as alb200_mas_device_ex (lengths or mask)
 * in ONE call.  When the search leaves SMs free (batch <= #SM: one CTA or cluster per utterance) the two kernels run
 * CONCURRENTLY: the score kernel produces 128-frame tiles in the order the search consumes them, on the free SMs, and
 * publishes each tile; the search is launched programmatically dependent, starts beside it and its loader warps wait
 * per tile, so a tile is read out of L2 while the next ones are still being computed.  Results are bit-identical to
 * alb200_neg_cent_gaussian followed by alb200_mas_device_ex (same kernels, same values).  Larger batches run the two
 * kernels back to back.  workspace: alb200_fused_workspace_bytes() bytes, 256-byte aligned, ZERO when first handed
 * over and private to this (device, stream) afterwards.
 * ------------------------------------------------------------------------ */
size_t alb200_fused_workspace_bytes(int b, int c, int tx, int ty);
int alb200_gaussian_mas_fused(const float *z, const float *m_p, const float *logs_p, float *neg_cent,
                              const int32_t *t_xs, const int32_t *t_ys,
                              const void *mask, int mask_dtype, int64_t mask_stride_b, int64_t mask_stride_x, int64_t mask_stride_y,
                              void *paths, int path_elem_size, uint64_t path_one, int zero_fill,
                              int32_t *frame_tok, int32_t *durations,
                              int b, int c, int tx, int ty, float max_neg_val,
                              void *workspace, size_t workspace_bytes, void *stream);

/* Number of kernels this library has launched on this thread since load. */
uint64_t alb200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* ALIGNER_B200_H_ */
