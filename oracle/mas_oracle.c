/*
 * mas_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the reference's Monotonic Alignment Search hot path,
 * written from the behaviour of the reference, not copied from it.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product path (art_tts_b200/) never does.
 *
 * Parity status: PINNED.  oracle/build_ref.py compiles the reference's own
 * Cython kernel from /root/reference into oracle/_ref/ and
 * tests/golden/make_golden.py records its outputs; tests/test_oracle.py checks
 * this file against those golden vectors (and against oracle/_ref directly
 * whenever it is present).
 *
 * Reference citations (relative to /root/reference/):
 *   src/model/monotonic_align/core.pyx:9-35    forward DP + backtrack (per utterance)
 *   src/model/monotonic_align/core.pyx:38-45   batch loop (prange)
 *   src/model/monotonic_align/__init__.py:8-23 mask multiply, fp32 cast, lengths from mask
 *   src/model/tts.py:483-495                   Gaussian log-prior, expanded form
 *   src/model/tts.py:503-505                   durations = sum_y path
 *   src/model/utils.py:26-43                   generate_path (durations -> path)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define MAS_NEG (-1e9f)

/* core.pyx:9-35.  value is [t_x_stride rows][T_y] fp32, y contiguous, updated in place
 * exactly like the reference; path is int32, same layout, pre-zeroed by the caller.
 * Arithmetic: one fp32 compare-select ((prev > cur) ? prev : cur) and one fp32 add. */
static void oracle_each(int32_t *path, float *value, int64_t T_y, int t_x, int t_y,
                        float neg)
{
    int x, y;
    int index = t_x - 1;
    for (y = 0; y < t_y; ++y) {
        int lo = t_x + y - t_y;
        int hi = (t_x < y + 1) ? t_x : (y + 1);
        if (lo < 0) lo = 0;
        for (x = lo; x < hi; ++x) {
            float v_cur, v_prev, m;
            v_cur = (x == y) ? neg : value[(int64_t)x * T_y + (y - 1)];
            if (x == 0)
                v_prev = (y == 0) ? 0.0f : neg;
            else
                v_prev = value[(int64_t)(x - 1) * T_y + (y - 1)];
            m = (v_prev > v_cur) ? v_prev : v_cur;
            value[(int64_t)x * T_y + y] = m + value[(int64_t)x * T_y + y];
        }
    }
    for (y = t_y - 1; y >= 0; --y) {
        path[(int64_t)index * T_y + y] = 1;
        /* short-circuit order as in core.pyx:34; the y==0 read of column -1 is only
         * reached when index != 0 and index != y, and its result is never used. */
        if (index != 0 &&
            (index == y ||
             (y > 0 && value[(int64_t)index * T_y + (y - 1)] <
                           value[(int64_t)(index - 1) * T_y + (y - 1)])))
            index -= 1;
    }
}

/* core.pyx:38-45.  values is clobbered (cumulative scores), paths must be zeroed.
 * n_threads <= 1 -> serial (what the reference's setup.py actually builds);
 * n_threads  > 1 -> OpenMP over utterances (what prange would do with -fopenmp). */
void mas_oracle_maximum_path_c(int32_t *paths, float *values, const int32_t *t_xs,
                               const int32_t *t_ys, int B, int T_x, int T_y,
                               float max_neg_val, int n_threads)
{
    int64_t per = (int64_t)T_x * T_y;
    int i;
    (void)n_threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads > 1 ? n_threads : 1)
#endif
    for (i = 0; i < B; ++i) {
        if (t_xs[i] < 1 || t_ys[i] < 0) continue; /* reference is UB here */
        oracle_each(paths + i * per, values + i * per, T_y, t_xs[i], t_ys[i], max_neg_val);
    }
}

/* __init__.py:13-21 for an fp32 value and fp32 0/1 mask: v = value*mask,
 * t_x = sum_x mask[b,x,0], t_y = sum_y mask[b,0,y] (truncated to int32). */
void mas_oracle_apply_mask(float *value, const float *mask, int32_t *t_xs, int32_t *t_ys,
                           int B, int T_x, int T_y)
{
    int64_t per = (int64_t)T_x * T_y;
    for (int b = 0; b < B; ++b) {
        float sx = 0.0f, sy = 0.0f;
        for (int x = 0; x < T_x; ++x) sx += mask[b * per + (int64_t)x * T_y];
        for (int y = 0; y < T_y; ++y) sy += mask[b * per + y];
        t_xs[b] = (int32_t)sx;
        t_ys[b] = (int32_t)sy;
        for (int64_t i = 0; i < per; ++i) value[b * per + i] *= mask[b * per + i];
    }
}

/* tts.py:483-495 restated cell by cell in fp32, expanded form
 *   lp = y_square - y_mu_double + mu_square + const
 * with y_square = sum_f(-0.5 * y^2), y_mu_double = sum_f((2*(-0.5*mu)) * y),
 * mu_square = sum_f(-0.5 * mu^2), const = -0.5*log(2*pi)*F.
 * Sums run over f in ascending order (BLAS order is unspecified; tests therefore
 * use a tolerance against the torch result and this function also offers fp64). */
void mas_oracle_log_prior_f32(float *lp, const float *mu_x, const float *y, int B, int F,
                              int T_x, int T_y)
{
    const float cst = (float)(-0.5 * log(2.0 * M_PI) * (double)F);
    for (int b = 0; b < B; ++b) {
        const float *mu = mu_x + (int64_t)b * F * T_x;
        const float *yy = y + (int64_t)b * F * T_y;
        for (int i = 0; i < T_x; ++i) {
            float musq = 0.0f;
            for (int f = 0; f < F; ++f) {
                float m = mu[(int64_t)f * T_x + i];
                musq += -0.5f * (m * m);
            }
            for (int j = 0; j < T_y; ++j) {
                float ysq = 0.0f, ymu = 0.0f;
                for (int f = 0; f < F; ++f) {
                    float m = mu[(int64_t)f * T_x + i];
                    float v = yy[(int64_t)f * T_y + j];
                    ysq += -0.5f * (v * v);
                    ymu += (2.0f * (-0.5f * m)) * v;
                }
                lp[((int64_t)b * T_x + i) * T_y + j] = ((ysq - ymu) + musq) + cst;
            }
        }
    }
}

void mas_oracle_log_prior_f64(double *lp, const float *mu_x, const float *y, int B, int F,
                              int T_x, int T_y)
{
    const double cst = -0.5 * log(2.0 * M_PI) * (double)F;
    for (int b = 0; b < B; ++b) {
        const float *mu = mu_x + (int64_t)b * F * T_x;
        const float *yy = y + (int64_t)b * F * T_y;
        for (int i = 0; i < T_x; ++i)
            for (int j = 0; j < T_y; ++j) {
                double s = 0.0;
                for (int f = 0; f < F; ++f) {
                    double d = (double)yy[(int64_t)f * T_y + j] - (double)mu[(int64_t)f * T_x + i];
                    s += d * d;
                }
                lp[((int64_t)b * T_x + i) * T_y + j] = -0.5 * s + cst;
            }
    }
}

/* tts.py:503-505: durations[b,x] = sum_y path[b,x,y]. */
void mas_oracle_durations(int32_t *dur, const int32_t *paths, int B, int T_x, int T_y)
{
    for (int64_t r = 0; r < (int64_t)B * T_x; ++r) {
        int32_t s = 0;
        for (int y = 0; y < T_y; ++y) s += paths[r * T_y + y];
        dur[r] = s;
    }
}

/* utils.py:26-43: cum = cumsum(dur); path[x,y] = (y < cum[x]) - (y < cum[x-1]); then
 * multiplied by the mask, which for a rectangular mask is (x < t_x && y < t_y). */
void mas_oracle_generate_path(int32_t *paths, const int32_t *dur, const int32_t *t_xs,
                              const int32_t *t_ys, int B, int T_x, int T_y)
{
    for (int b = 0; b < B; ++b) {
        int64_t cum = 0;
        for (int x = 0; x < T_x; ++x) {
            int64_t prev = cum;
            cum += dur[(int64_t)b * T_x + x];
            for (int y = 0; y < T_y; ++y) {
                int v = ((y < cum) ? 1 : 0) - ((y < prev) ? 1 : 0);
                if (!(x < t_xs[b] && y < t_ys[b])) v = 0;
                paths[((int64_t)b * T_x + x) * T_y + y] = v;
            }
        }
    }
}

int mas_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
