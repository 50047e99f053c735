/*
 * msda_oracle.c — CPU restatement of TCE-RVOS's multi-scale deformable attention.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may include, link, import or execute
 * this file; it is the checker used by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * leg.  The product path has no CPU fallback.
 *
 * Parity status: PINNED.  oracle/gen_golden.py imports the reference's own
 * ms_deform_attn_core_pytorch (models/ops/functions/ms_deform_attn_func.py:67-87) in the build
 * container, runs it (fp64 forward + autograd backward) on seeded inputs and commits inputs and
 * outputs under tests/golden/; tests/test_oracle.py checks this file against every one of them.
 *
 * What is restated (file:line under the reference tree, models/ops/src/cuda/):
 *   forward, per output element     ms_deform_im2col_cuda.cuh:421-452
 *     h_im = loc_h*H - 0.5, w_im = loc_w*W - 0.5; the point counts only if
 *     h_im > -1 && w_im > -1 && h_im < H && w_im < W                      (:441)
 *   bilinear with per-corner zero padding, weights hh*hw, hh*lw, lh*hw, lh*lw   (:34-85)
 *   backward formulas               ms_deform_im2col_cuda.cuh:88-160
 *     grad_value[corner] += w_corner * top_grad * attn                    (:126,:135,:144,:153)
 *     grad_attn  = top_grad * val                                        (:157)
 *     grad_loc.x = W * grad_w_weight * top_grad * attn                   (:158)
 *     grad_loc.y = H * grad_h_weight * top_grad * attn                   (:159)
 *     summed over the D channels of the head                             (:537-553)
 *   the fork's 3-D variant (third coordinate = fractional frame inside a group of G frames):
 *     forward :366-420, backward :162-242 and :565-712.  Its frame gradient is scaled by (G-1)
 *     although the forward scales by G (:241) — restated as is.
 *
 * The same body is compiled twice, for double (`_f64`) and float (`_f32`) arithmetic.
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared).
 */
#ifndef MSDA_ORACLE_BODY

#include <math.h>
#include <stdint.h>
#include <string.h>

#define MSDA_ORACLE_BODY
#define REAL double
#define FN(name) name##_f64
#define FLOOR floor
#include "msda_oracle.c"
#undef REAL
#undef FN
#undef FLOOR
#define REAL float
#define FN(name) name##_f32
#define FLOOR floorf
#include "msda_oracle.c"

int msda_oracle_version(void) { return 1; }

#else /* ---------------------------------- body ---------------------------------- */

/* one sampling point of one (frame, query, head): geometry shared by forward and backward */
typedef struct {
    int inside;
    int h_low, w_low;
    REAL lh, lw;
} FN(point_t);

static FN(point_t) FN(resolve)(REAL loc_w, REAL loc_h, int64_t H, int64_t W)
{
    FN(point_t) p;
    const REAL h_im = loc_h * (REAL)H - (REAL)0.5;
    const REAL w_im = loc_w * (REAL)W - (REAL)0.5;
    p.inside = (h_im > (REAL)-1 && w_im > (REAL)-1 && h_im < (REAL)H && w_im < (REAL)W);
    p.h_low = 0; p.w_low = 0; p.lh = 0; p.lw = 0;
    if (p.inside) {
        const REAL hf = FLOOR(h_im), wf = FLOOR(w_im);
        p.h_low = (int)hf; p.w_low = (int)wf;
        p.lh = h_im - hf; p.lw = w_im - wf;
    }
    return p;
}

/* bilinear read of channel c of head m from one level of one frame */
static REAL FN(bilinear)(const REAL *level, int64_t H, int64_t W, int M, int D, const FN(point_t) *p, int m, int c)
{
    const REAL hh = 1 - p->lh, hw = 1 - p->lw;
    const int64_t ws = (int64_t)M * D, hs = W * ws;
    const int hl = p->h_low, wl = p->w_low, hu = hl + 1, wu = wl + 1;
    const int64_t b = (int64_t)m * D + c;
    REAL v1 = 0, v2 = 0, v3 = 0, v4 = 0;
    if (hl >= 0 && wl >= 0) v1 = level[hl * hs + wl * ws + b];
    if (hl >= 0 && wu <= W - 1) v2 = level[hl * hs + wu * ws + b];
    if (hu <= H - 1 && wl >= 0) v3 = level[hu * hs + wl * ws + b];
    if (hu <= H - 1 && wu <= W - 1) v4 = level[hu * hs + wu * ws + b];
    return hh * hw * v1 + hh * p->lw * v2 + p->lh * hw * v3 + p->lh * p->lw * v4;
}

/* forward.  is_3d: loc has 3 coordinates, frames are blended inside groups of `group` frames. */
void FN(msda_oracle_forward)(const REAL *value, const int64_t *shapes, const int64_t *lsi,
                             const REAL *loc, const REAL *aw,
                             int64_t N, int64_t S, int64_t Lq, int M, int D, int L, int P,
                             int is_3d, int64_t group, REAL *out)
{
    const int LD = is_3d ? 3 : 2;
    const int64_t pix = (int64_t)M * D;
    if (!is_3d || group <= 0 || group > N) group = N;
    #pragma omp parallel for schedule(static)
    for (int64_t nq = 0; nq < N * Lq; ++nq) {
        const int64_t n = nq / Lq;
        const int64_t g0 = is_3d ? (n / group) * group : n;          /* first frame of the group */
        const int64_t G = is_3d ? ((g0 + group <= N) ? group : N - g0) : 1;
        for (int m = 0; m < M; ++m) {
            const REAL *lp = loc + (nq * M + m) * (int64_t)L * P * LD;
            const REAL *ap = aw + (nq * M + m) * (int64_t)L * P;
            for (int c = 0; c < D; ++c) {
                REAL col = 0;
                for (int l = 0; l < L; ++l) {
                    const int64_t H = shapes[2 * l], W = shapes[2 * l + 1];
                    for (int p = 0; p < P; ++p) {
                        const int k = l * P + p;
                        const FN(point_t) pt = FN(resolve)(lp[k * LD], lp[k * LD + 1], H, W);
                        if (!pt.inside) continue;
                        if (!is_3d) {
                            col += FN(bilinear)(value + (n * S + lsi[l]) * pix, H, W, M, D, &pt, m, c) * ap[k];
                        } else {
                            const REAL f_im = lp[k * LD + 2] * (REAL)G - (REAL)0.5;
                            const int64_t f_low = (int64_t)FLOOR(f_im), f_high = f_low + 1;
                            REAL lo = 0, hi = 0;
                            if (f_low >= 0 && f_low < G)
                                lo = FN(bilinear)(value + ((g0 + f_low) * S + lsi[l]) * pix, H, W, M, D, &pt, m, c);
                            if (f_high >= 0 && f_high < G)
                                hi = FN(bilinear)(value + ((g0 + f_high) * S + lsi[l]) * pix, H, W, M, D, &pt, m, c);
                            col += (lo * ((REAL)f_high - f_im) + hi * (f_im - (REAL)f_low)) * ap[k];
                        }
                    }
                }
                out[(nq * M + m) * D + c] = col;
            }
        }
    }
}

/* backward of one corner set of one frame; accumulates the per-channel partial gradients */
static void FN(col2im)(const REAL *level, REAL *glevel, int64_t H, int64_t W, int M, int D,
                       const FN(point_t) *p, int m, int c, REAL fw, REAL fsign,
                       REAL top_grad, REAL attn, REAL *g_w, REAL *g_h, REAL *g_f, REAL *g_a)
{
    const REAL lh = p->lh, lw = p->lw, hh = 1 - lh, hw = 1 - lw;
    const int64_t ws = (int64_t)M * D, hs = W * ws;
    const int hl = p->h_low, wl = p->w_low, hu = hl + 1, wu = wl + 1;
    const int64_t b = (int64_t)m * D + c;
    const REAL tgv = top_grad * attn;
    REAL gh = 0, gw = 0, gf = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0;
    if (hl >= 0 && wl >= 0) {
        const int64_t o = hl * hs + wl * ws + b;
        v1 = level[o]; gh -= hw * fw * v1; gw -= hh * fw * v1; gf += fsign * hw * hh * v1;
        glevel[o] += hh * hw * fw * tgv;
    }
    if (hl >= 0 && wu <= W - 1) {
        const int64_t o = hl * hs + wu * ws + b;
        v2 = level[o]; gh -= lw * fw * v2; gw += hh * fw * v2; gf += fsign * lw * hh * v2;
        glevel[o] += hh * lw * fw * tgv;
    }
    if (hu <= H - 1 && wl >= 0) {
        const int64_t o = hu * hs + wl * ws + b;
        v3 = level[o]; gh += hw * fw * v3; gw -= lh * fw * v3; gf += fsign * hw * lh * v3;
        glevel[o] += lh * hw * fw * tgv;
    }
    if (hu <= H - 1 && wu <= W - 1) {
        const int64_t o = hu * hs + wu * ws + b;
        v4 = level[o]; gh += lw * fw * v4; gw += lh * fw * v4; gf += fsign * lw * lh * v4;
        glevel[o] += lh * lw * fw * tgv;
    }
    const REAL val = (hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4) * fw;
    *g_a += top_grad * val;
    *g_w += (REAL)W * gw * tgv;
    *g_h += (REAL)H * gh * tgv;
    *g_f += gf * tgv;
}

/* backward.  Outputs are fully overwritten (grad_value is zeroed here).  Parallel over frame
 * groups only, so grad_value accumulation is race free and deterministic. */
void FN(msda_oracle_backward)(const REAL *value, const int64_t *shapes, const int64_t *lsi,
                              const REAL *loc, const REAL *aw, const REAL *grad_out,
                              int64_t N, int64_t S, int64_t Lq, int M, int D, int L, int P,
                              int is_3d, int64_t group,
                              REAL *grad_value, REAL *grad_loc, REAL *grad_aw)
{
    const int LD = is_3d ? 3 : 2;
    const int64_t pix = (int64_t)M * D;
    if (!is_3d) group = 1;
    else if (group <= 0 || group > N) group = N;
    memset(grad_value, 0, sizeof(REAL) * (size_t)(N * S * pix));
    const int64_t n_groups = (N + group - 1) / (group > 0 ? group : 1);
    #pragma omp parallel for schedule(dynamic, 1)
    for (int64_t gi = 0; gi < n_groups; ++gi) {
        const int64_t g0 = gi * group;
        const int64_t G = (g0 + group <= N) ? group : N - g0;
        for (int64_t n = g0; n < g0 + G; ++n)
        for (int64_t q = 0; q < Lq; ++q)
        for (int m = 0; m < M; ++m) {
            const int64_t qh = (n * Lq + q) * M + m;
            const REAL *lp = loc + qh * (int64_t)L * P * LD;
            const REAL *ap = aw + qh * (int64_t)L * P;
            for (int l = 0; l < L; ++l) {
                const int64_t H = shapes[2 * l], W = shapes[2 * l + 1];
                for (int p = 0; p < P; ++p) {
                    const int k = l * P + p;
                    REAL g_w = 0, g_h = 0, g_f = 0, g_a = 0;
                    const FN(point_t) pt = FN(resolve)(lp[k * LD], lp[k * LD + 1], H, W);
                    if (pt.inside) {
                        for (int c = 0; c < D; ++c) {
                            const REAL tg = grad_out[qh * D + c];
                            if (!is_3d) {
                                const int64_t o = (n * S + lsi[l]) * pix;
                                FN(col2im)(value + o, grad_value + o, H, W, M, D, &pt, m, c, (REAL)1, (REAL)0,
                                           tg, ap[k], &g_w, &g_h, &g_f, &g_a);
                            } else {
                                const REAL f_im = lp[k * LD + 2] * (REAL)G - (REAL)0.5;
                                const int64_t f_low = (int64_t)FLOOR(f_im), f_high = f_low + 1;
                                const REAL lf = f_im - (REAL)f_low, hf = 1 - lf;
                                if (f_low >= 0 && f_low < G) {
                                    const int64_t o = ((g0 + f_low) * S + lsi[l]) * pix;
                                    FN(col2im)(value + o, grad_value + o, H, W, M, D, &pt, m, c, hf, (REAL)-1,
                                               tg, ap[k], &g_w, &g_h, &g_f, &g_a);
                                }
                                if (f_high >= 0 && f_high < G) {
                                    const int64_t o = ((g0 + f_high) * S + lsi[l]) * pix;
                                    FN(col2im)(value + o, grad_value + o, H, W, M, D, &pt, m, c, lf, (REAL)1,
                                               tg, ap[k], &g_w, &g_h, &g_f, &g_a);
                                }
                            }
                        }
                    }
                    grad_aw[qh * (int64_t)L * P + k] = g_a;
                    grad_loc[(qh * (int64_t)L * P + k) * LD] = g_w;
                    grad_loc[(qh * (int64_t)L * P + k) * LD + 1] = g_h;
                    if (is_3d) grad_loc[(qh * (int64_t)L * P + k) * LD + 2] = (REAL)(G - 1) * g_f;
                }
            }
        }
    }
}

#endif
