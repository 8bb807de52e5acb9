/*
 * ifk_oracle.c -- CPU restatement of Inverse-Flow's inverse-convolution hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker* for the CUDA path in
 * inverse_flow_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product never does.
 *
 * What it restates (paths relative to the reference checkout):
 *   inverse, groups == 1 : inf/utils/solve_mc.py:88-114 (raster `solve`), which is
 *                          inf/utils/solve_mc.py:8-50 / fastflow_inverse/
 *                          solve_parallel_mc.pyx:100-124 (wavefront) with another
 *                          visiting order, and cinc_kernel_level1.cu:57-69.
 *   inverse, groups  > 1 : inf/utils/inv_conv_cuda/cinc_kernel_level2.cu:59-72
 *                          (reads channel k_c + order*order_stride of the group).
 *   conv (sampling)      : inv_conv_with_bp_kernel_general.cu:182-198 with the
 *                          centre-tap triangular mask of solve_mc.py:104-108.
 *   dX, dW               : the reference has NO CPU backward.  They are the adjoint
 *                          of the solver above and
 *                          are pinned by the derivative of solve_mc.py's own solver
 *                          (linearity + finite differences), tests/golden/make_golden.py.
 *   literal_* functions  : bug-for-bug restatement of the shipped CUDA kernels
 *                          inv_conv_with_bp_kernel_general.cu:52-65 (inverse),
 *                          :307-327 + :371-383 (dy) -- compat mode only.
 *
 * Weight layout: (C, Cw, KH, KW) contiguous, Cw >= C/groups; the weight that
 * multiplies the neighbour at shift (k_h, k_w) is W[c][k_c][KH-1-k_h][KW-1-k_w]
 * (solve_mc.py:109-110).  The tap (0,0,k_c==c) is the implicit unit diagonal and
 * the centre tap with k_c > c is masked (solve_mc.py:104-108).
 *
 * Parity pinning: checked bit-for-bit (float64) against solve_mc.py and against
 * the compiled reference Cython solver oracle/_ref; see tests/test_oracle.py.
 *
 * Build: make -C oracle      (plain C11, -ffp-contract=off so that the float64
 * results are bit-identical with the reference's non-FMA arithmetic).
 */
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int ifk_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

#define IDX4(n1, n2, n3, i0, i1, i2, i3) \
    ((((size_t)(i0) * (n1) + (i1)) * (n2) + (i2)) * (n3) + (i3))

#define DEFINE_ORACLE(T, SUFFIX)                                                            \
                                                                                            \
/* y = L^-1 x.  solve_mc.py:88-114 visiting order (b, h, w, c); groups generalise       */ \
/* cinc_kernel_level2.cu:59-72.                                                         */ \
void ifk_oracle_inverse_##SUFFIX(const T *x, const T *wt, T *y, int B, int C, int H,        \
                                 int W, int KH, int KW, int Cw, int groups, int threads)    \
{                                                                                           \
    const int Cg = C / groups;                                                              \
    memcpy(y, x, sizeof(T) * (size_t)B * C * H * W);        /* y = x.clone() */             \
    _Pragma("omp parallel for schedule(static) num_threads(threads)")                       \
    for (int b = 0; b < B; b++)                                                             \
        for (int h = 0; h < H; h++)                                                         \
            for (int w = 0; w < W; w++)                                                     \
                for (int c = 0; c < C; c++) {                                               \
                    const int base = (c / Cg) * Cg, cl = c - base;                          \
                    T *out = &y[IDX4(C, H, W, b, c, h, w)];                                 \
                    for (int kh = 0; kh < KH; kh++) {                                       \
                        if (h - kh < 0) break;                                              \
                        for (int kw = 0; kw < KW; kw++) {                                   \
                            if (w - kw < 0) break;                                          \
                            for (int kc = 0; kc < Cg; kc++) {                               \
                                if (kh == 0 && kw == 0) {                                   \
                                    if (kc == cl) continue;                                 \
                                    if (cl - kc < 0) break;                                 \
                                }                                                           \
                                *out -= y[IDX4(C, H, W, b, base + kc, h - kh, w - kw)] *    \
                                        wt[IDX4(Cw, KH, KW, c, kc, KH - kh - 1,             \
                                                KW - kw - 1)];                              \
                            }                                                               \
                        }                                                                   \
                    }                                                                       \
                }                                                                           \
}                                                                                           \
                                                                                            \
/* Same solve in the anti-diagonal wavefront order of solve_mc.py:8-50 /                */ \
/* solve_parallel_mc.pyx:100-124 (step i, channel c, pixel (j, i-j)), with the step     */ \
/* count corrected to H+W-1 (the reference assumes H == W, .pyx:95-98).                 */ \
void ifk_oracle_inverse_wavefront_##SUFFIX(const T *x, const T *wt, T *y, int B, int C,     \
                                           int H, int W, int KH, int KW, int Cw,            \
                                           int groups, int threads)                         \
{                                                                                           \
    const int Cg = C / groups;                                                              \
    memcpy(y, x, sizeof(T) * (size_t)B * C * H * W);                                        \
    _Pragma("omp parallel for schedule(static) num_threads(threads)")                       \
    for (int b = 0; b < B; b++)                                                             \
        for (int i = 0; i < H + W - 1; i++)                                                 \
            for (int c = 0; c < C; c++) {                                                   \
                const int base = (c / Cg) * Cg, cl = c - base;                              \
                for (int j = 0; j < H; j++) {                                               \
                    if (j > i) break;                                                       \
                    const int h = j, w = i - j;                                             \
                    if (w >= W) continue;                                                   \
                    T *out = &y[IDX4(C, H, W, b, c, h, w)];                                 \
                    for (int kh = 0; kh < KH; kh++) {                                       \
                        if (h - kh < 0) break;                                              \
                        for (int kw = 0; kw < KW; kw++) {                                   \
                            if (w - kw < 0) break;                                          \
                            for (int kc = 0; kc < Cg; kc++) {                               \
                                if (kh == 0 && kw == 0) {                                   \
                                    if (kc == cl) continue;                                 \
                                    if (cl - kc < 0) break;                                 \
                                }                                                           \
                                *out -= y[IDX4(C, H, W, b, base + kc, h - kh, w - kw)] *    \
                                        wt[IDX4(Cw, KH, KW, c, kc, KH - kh - 1,             \
                                                KW - kw - 1)];                              \
                            }                                                               \
                        }                                                                   \
                    }                                                                       \
                }                                                                           \
            }                                                                               \
}                                                                                           \
                                                                                            \
/* x = L y : the masked (autoregressive) convolution, sampling direction.               */ \
/* inv_conv_with_bp_kernel_general.cu:182-198 + triangular centre mask.                 */ \
void ifk_oracle_conv_##SUFFIX(const T *y, const T *wt, T *x, int B, int C, int H, int W,    \
                              int KH, int KW, int Cw, int groups, int threads)              \
{                                                                                           \
    const int Cg = C / groups;                                                              \
    _Pragma("omp parallel for schedule(static) num_threads(threads)")                       \
    for (int b = 0; b < B; b++)                                                             \
        for (int c = 0; c < C; c++) {                                                       \
            const int base = (c / Cg) * Cg, cl = c - base;                                  \
            for (int h = 0; h < H; h++)                                                     \
                for (int w = 0; w < W; w++) {                                               \
                    T acc = y[IDX4(C, H, W, b, c, h, w)];                                   \
                    for (int kh = 0; kh < KH; kh++) {                                       \
                        if (h - kh < 0) break;                                              \
                        for (int kw = 0; kw < KW; kw++) {                                   \
                            if (w - kw < 0) break;                                          \
                            for (int kc = 0; kc < Cg; kc++) {                               \
                                if (kh == 0 && kw == 0) {                                   \
                                    if (kc == cl) continue;                                 \
                                    if (cl - kc < 0) break;                                 \
                                }                                                           \
                                acc += y[IDX4(C, H, W, b, base + kc, h - kh, w - kw)] *     \
                                       wt[IDX4(Cw, KH, KW, c, kc, KH - kh - 1,              \
                                               KW - kw - 1)];                               \
                            }                                                               \
                        }                                                                   \
                    }                                                                       \
                    x[IDX4(C, H, W, b, c, h, w)] = acc;                                     \
                }                                                                           \
        }                                                                                   \
}                                                                                           \
                                                                                            \
/* dX = L^-T g : adjoint solve, reverse raster order.      */ \
/* dX[b,kc,p] = g[b,kc,p] - sum_{(q,c)!=(0,kc), centre: c>kc} W[c,kc,q] dX[b,c,p+q]     */ \
void ifk_oracle_bwd_input_##SUFFIX(const T *g, const T *wt, T *dx, int B, int C, int H,     \
                                   int W, int KH, int KW, int Cw, int groups, int threads)  \
{                                                                                           \
    const int Cg = C / groups;                                                              \
    memcpy(dx, g, sizeof(T) * (size_t)B * C * H * W);                                       \
    _Pragma("omp parallel for schedule(static) num_threads(threads)")                       \
    for (int b = 0; b < B; b++)                                                             \
        for (int h = H - 1; h >= 0; h--)                                                    \
            for (int w = W - 1; w >= 0; w--)                                                \
                for (int k = C - 1; k >= 0; k--) {                                          \
                    const int base = (k / Cg) * Cg, kl = k - base;                          \
                    T *out = &dx[IDX4(C, H, W, b, k, h, w)];                                \
                    for (int kh = 0; kh < KH; kh++) {                                       \
                        if (h + kh >= H) break;                                             \
                        for (int kw = 0; kw < KW; kw++) {                                   \
                            if (w + kw >= W) break;                                         \
                            for (int cl = Cg - 1; cl >= 0; cl--) {                          \
                                if (kh == 0 && kw == 0 && cl <= kl) break;                  \
                                *out -= dx[IDX4(C, H, W, b, base + cl, h + kh, w + kw)] *   \
                                        wt[IDX4(Cw, KH, KW, base + cl, kl, KH - kh - 1,     \
                                                KW - kw - 1)];                              \
                            }                                                               \
                        }                                                                   \
                    }                                                                       \
                }                                                                           \
}                                                                                           \
                                                                                            \
/* dW[c,kc,KH-1-qh,KW-1-qw] = - sum_{b,p} dX[b,c,p] y[b,base+kc,p-q]; masked taps and   */ \
/* columns kc >= Cg get 0.  Accumulates in double, fixed   */ \
/* (b, h, w) order, so the result does not depend on `threads`.                         */ \
void ifk_oracle_bwd_weight_##SUFFIX(const T *dx, const T *y, T *dw, int B, int C, int H,    \
                                    int W, int KH, int KW, int Cw, int groups, int threads) \
{                                                                                           \
    const int Cg = C / groups;                                                              \
    memset(dw, 0, sizeof(T) * (size_t)C * Cw * KH * KW);                                    \
    _Pragma("omp parallel for collapse(2) schedule(static) num_threads(threads)")           \
    for (int c = 0; c < C; c++) {                                                           \
        for (int kc = 0; kc < Cg; kc++) {                                                   \
            const int base = (c / Cg) * Cg, cl = c - base;                                  \
            for (int qh = 0; qh < KH; qh++)                                                 \
                for (int qw = 0; qw < KW; qw++) {                                           \
                    if (qh == 0 && qw == 0 && kc >= cl) continue;                           \
                    double acc = 0.0;                                                       \
                    for (int b = 0; b < B; b++)                                             \
                        for (int h = qh; h < H; h++)                                        \
                            for (int w = qw; w < W; w++)                                    \
                                acc += (double)dx[IDX4(C, H, W, b, c, h, w)] *              \
                                       (double)y[IDX4(C, H, W, b, base + kc, h - qh,        \
                                                      w - qw)];                             \
                    dw[IDX4(Cw, KH, KW, c, kc, KH - 1 - qh, KW - 1 - qw)] = (T)(-acc);      \
                }                                                                           \
        }                                                                                   \
    }                                                                                       \
}                                                                                           \
                                                                                            \
/* ---- literal (bug-for-bug) restatement of the shipped kernels: compat oracle ------- */ \
/* inv_conv_with_bp_kernel_general.cu:52-65.  4 hard-coded channel groups, reads its    */ \
/* OWN channel (c + order*os) on the right-hand side, in-place read-modify-write, no    */ \
/* triangular mask; C < 4 leaves the pre-zeroed output untouched (.cu:97-98).           */ \
void ifk_oracle_literal_inverse_##SUFFIX(const T *x, const T *wt, T *y, int B, int C,       \
                                         int H, int W, int KH, int KW, int Cw)              \
{                                                                                           \
    const int os = C / 4;                                                                   \
    memset(y, 0, sizeof(T) * (size_t)B * C * H * W);                                        \
    for (int d = 1; d <= H + W - 1; d++)                                                    \
        for (int c = 0; c < os; c++) {                                                      \
            int rt = d;                                                                     \
            const int n = H < W ? H : W, m = H > W ? H : W;                                 \
            if (d > n) rt = (d <= m) ? n : m + n - d;                                       \
            for (int tid = 0; tid < rt; tid++)                                              \
                for (int order = 0; order < 4; order++)                                     \
                    for (int b = 0; b < B; b++) {                                           \
                        int h, w;                                                           \
                        if (d <= H) { h = d - 1 - tid; w = tid; }                           \
                        else        { w = (d - H) + tid; h = H - 1 - tid; }                 \
                        if (h < 0 || w < 0 || h >= H || w >= W) continue;                   \
                        const int cc = c + order * os;                                      \
                        T *out = &y[IDX4(C, H, W, b, cc, h, w)];                            \
                        *out = x[IDX4(C, H, W, b, cc, h, w)];                               \
                        for (int kh = 0; kh < KH; kh++) {                                   \
                            if (h - kh < 0) break;                                          \
                            for (int kw = 0; kw < KW; kw++) {                               \
                                if (w - kw < 0) break;                                      \
                                for (int kc = 0; kc < os; kc++) {                           \
                                    if (kh == 0 && kw == 0 && kc == c) continue;            \
                                    *out -= y[IDX4(C, H, W, b, cc, h - kh, w - kw)] *       \
                                            wt[IDX4(Cw, KH, KW, cc, kc, KH - kh - 1,        \
                                                    KW - kw - 1)];                          \
                                }                                                           \
                            }                                                               \
                        }                                                                   \
                    }                                                                       \
        }                                                                                   \
}                                                                                           \
                                                                                            \
/* inv_conv_with_bp_kernel_general.cu:307-327 (impulse response M) followed by          */ \
/* :371-383 (dense causal correlation): out = L^-1 g, NOT the true input gradient.      */ \
/* Only meaningful for C == 4 (os == 1); O((HW)^2) per channel-image.                   */ \
void ifk_oracle_literal_dy_##SUFFIX(const T *g, const T *wt, T *out, int B, int C, int H,   \
                                    int W, int KH, int KW, int Cw)                          \
{                                                                                           \
    const int os = C / 4;                                                                   \
    T *M = (T *)calloc((size_t)B * C * H * W, sizeof(T));                                   \
    memset(out, 0, sizeof(T) * (size_t)B * C * H * W);                                      \
    for (int pass = 0; pass < 2; pass++)                                                    \
        for (int s = 0; s <= H + W - 2; s++)                                                \
            for (int c = 0; c < os; c++)                                                    \
                for (int h = 0; h < H; h++) {                                               \
                    const int w = s - h;                                                    \
                    if (w < 0 || w >= W) continue;                                          \
                    for (int order = 0; order < 4; order++)                                 \
                        for (int b = 0; b < B; b++) {                                       \
                            const int cc = c + order * os;                                  \
                            if (pass == 0) {                                                \
                                T *m = &M[IDX4(C, H, W, b, cc, h, w)];                      \
                                if (h == 0 && w == 0) { *m = (T)1.0; continue; }            \
                                for (int kh = 0; kh < KH && h - kh >= 0; kh++)              \
                                    for (int kw = 0; kw < KW && w - kw >= 0; kw++)          \
                                        for (int kc = 0; kc < os; kc++) {                   \
                                            if (kh == 0 && kw == 0 && kc == c) continue;    \
                                            *m -= wt[IDX4(Cw, KH, KW, cc, kc, KH - 1 - kh,  \
                                                          KW - 1 - kw)] *                   \
                                                  M[IDX4(C, H, W, b, cc, h - kh, w - kw)];  \
                                        }                                                   \
                            } else {                                                        \
                                T *o = &out[IDX4(C, H, W, b, cc, h, w)];                    \
                                for (int kh = 0; kh < H && h - kh >= 0; kh++)               \
                                    for (int kw = 0; kw < W && w - kw >= 0; kw++)           \
                                        *o += g[IDX4(C, H, W, b, cc, kh, kw)] *             \
                                              M[IDX4(C, H, W, b, cc, h - kh, w - kw)];      \
                            }                                                               \
                        }                                                                   \
                }                                                                           \
    free(M);                                                                                \
}

DEFINE_ORACLE(float, f32)
DEFINE_ORACLE(double, f64)
