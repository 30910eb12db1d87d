/*
 * oracle/mas_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's monotonic alignment search, used as the
 * checker for the CUDA path.  Nothing under aligner_b200/ may link, import or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * Follows (reference = xiaozhah/Aligner, paths relative to its root):
 *   monotonic_align/core.pyx:7-35   maximum_path_each  (forward DP + backtrack)
 *   monotonic_align/core.pyx:38-45  maximum_path_c     (batch loop)
 *   monotonic_align/core.c:19384-19396  max() lowered to (v_prev > v_cur) ? v_prev : v_cur, fp32 add
 *   monotonic_align/core.c:19444        strict fp32 '<' in the backtrack
 *
 * Parity status: PINNED.  tests/test_oracle.py checks these functions against
 * (a) the hand-checkable known answers recorded in SURVEY.md section 8c,
 * (b) tests/golden/*.npz, produced by running the UNMODIFIED reference
 *     (re-cythonized core.pyx + its own __init__.py) via oracle/make_golden.py,
 * (c) oracle/_ref/ (the compiled reference itself) on random inputs, when present.
 *
 * Two restatements live here:
 *   mas_oracle_full     - in-place table DP exactly as the reference does it
 *                         (values are overwritten with cumulative scores).
 *   mas_oracle_bits     - the formulation the GPU kernel uses: one running
 *                         fp32 column + one direction bit per cell + bit-driven
 *                         backtrack.  Proven equal to mas_oracle_full by test.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- single utterance, table form (core.pyx:9-35) ------------------------ */
static void each_full(int32_t *path, float *value, int64_t ld, int t_x, int t_y,
                      float neg)
{
    /* forward: frames outer, tokens inner, band limited (core.pyx:17-18) */
    for (int y = 0; y < t_y; ++y) {
        int lo = t_x + y - t_y;
        if (lo < 0) lo = 0;
        int hi = (y + 1 < t_x) ? y + 1 : t_x;
        for (int x = lo; x < hi; ++x) {
            float stay = (x == y) ? neg : value[(int64_t)x * ld + (y - 1)];      /* core.pyx:19-22 */
            float move;
            if (x == 0)                                                          /* core.pyx:23-27 */
                move = (y == 0) ? 0.0f : neg;
            else
                move = value[(int64_t)(x - 1) * ld + (y - 1)];                   /* core.pyx:29 */
            float best = (move > stay) ? move : stay;                            /* core.c:19384-19391 */
            value[(int64_t)x * ld + y] = best + value[(int64_t)x * ld + y];      /* core.pyx:30 */
        }
    }
    /* backtrack (core.pyx:32-35) */
    int tok = t_x - 1;
    for (int y = t_y - 1; y >= 0; --y) {
        path[(int64_t)tok * ld + y] = 1;
        if (tok != 0 &&
            (tok == y || value[(int64_t)tok * ld + (y - 1)] < value[(int64_t)(tok - 1) * ld + (y - 1)]))
            tok -= 1;
    }
}

/*
 * paths  int32 [b, tx, ty]  pre-zeroed by the caller, ones are written in place
 * values float [b, tx, ty]  CLOBBERED with cumulative scores, like the reference
 * Returns 0.  Lengths are trusted exactly as the reference trusts them, except
 * that items with t_x<=0, t_y<=0 or t_x>t_y are skipped (the reference's
 * behaviour there is out-of-bounds / meaningless).
 */
int mas_oracle_full(int32_t *paths, float *values, const int32_t *t_xs,
                    const int32_t *t_ys, int b, int tx, int ty, float neg)
{
    int64_t item = (int64_t)tx * ty;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int i = 0; i < b; ++i) {
        int t_x = t_xs[i], t_y = t_ys[i];
        if (t_x <= 0 || t_y <= 0 || t_x > t_y || t_x > tx || t_y > ty) continue;
        each_full(paths + i * item, values + i * item, ty, t_x, t_y, neg);
    }
    return 0;
}

/* ---- single utterance, running-column + direction-bit form ---------------- */
static void each_bits(int32_t *path, const float *value, int64_t ld, int t_x,
                      int t_y, float neg, int32_t *frame_tok)
{
    int words = (t_y + 31) / 32;
    uint32_t *bits = (uint32_t *)calloc((size_t)t_x * words, sizeof(uint32_t));
    float *col = (float *)malloc(sizeof(float) * (size_t)t_x);
    float *nxt = (float *)malloc(sizeof(float) * (size_t)t_x);
    for (int x = 0; x < t_x; ++x) col[x] = neg;   /* rows above the diagonal are held at neg */

    for (int y = 0; y < t_y; ++y) {
        int hi = (y + 1 < t_x) ? y + 1 : t_x;     /* upper band edge is semantic, lower is not */
        for (int x = 0; x < hi; ++x) {
            float stay = col[x];                  /* == neg when x == y because row x was held */
            float move = (x == 0) ? ((y == 0) ? 0.0f : neg) : col[x - 1];
            int take = move > stay;
            nxt[x] = (take ? move : stay) + value[(int64_t)x * ld + y];
            /* row 0 can never step down; the diagonal must (core.pyx:34) */
            if (x == 0) take = 0;
            else if (x == y) take = 1;
            if (take) bits[(size_t)x * words + (y >> 5)] |= 1u << (y & 31);
        }
        for (int x = 0; x < hi; ++x) col[x] = nxt[x];
    }
    int tok = t_x - 1;
    for (int y = t_y - 1; y >= 0; --y) {
        if (path) path[(int64_t)tok * ld + y] = 1;
        if (frame_tok) frame_tok[y] = tok;
        if ((bits[(size_t)tok * words + (y >> 5)] >> (y & 31)) & 1u) tok -= 1;
    }
    free(bits); free(col); free(nxt);
}

/* values are NOT modified.  frame_tok (optional) int32 [b, ty]: token per frame, -1 past t_y. */
int mas_oracle_bits(int32_t *paths, const float *values, const int32_t *t_xs,
                    const int32_t *t_ys, int b, int tx, int ty, float neg,
                    int32_t *frame_tok)
{
    int64_t item = (int64_t)tx * ty;
    if (frame_tok) for (int64_t k = 0; k < (int64_t)b * ty; ++k) frame_tok[k] = -1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int i = 0; i < b; ++i) {
        int t_x = t_xs[i], t_y = t_ys[i];
        if (t_x <= 0 || t_y <= 0 || t_x > t_y || t_x > tx || t_y > ty) continue;
        each_bits(paths ? paths + i * item : 0, values + i * item, ty, t_x, t_y, neg,
                  frame_tok ? frame_tok + (int64_t)i * ty : 0);
    }
    return 0;
}

int mas_oracle_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
