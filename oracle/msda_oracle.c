/*
 * msda_oracle.c -- CPU restatement of the reference's multi-scale deformable
 * attention (MSDA) arithmetic.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product path (the CUDA library
 * behind include/msda_b200.h) never links, loads or calls anything in oracle/.
 *
 * What it restates (all file:line relative to
 * /root/reference/model/modeling/pixel_decoder/ops/src/cuda/):
 *   - pixel coordinate from a normalised location, `loc*size - 0.5`
 *     (ms_deform_im2col_cuda.cuh:290-291), evaluated as ONE fused multiply-add:
 *     that is what nvcc makes of the expression with the reference's build flags
 *     (`FFMA R, R, R, -0.5` in the SASS of the reference op compiled for sm_100,
 *     see baseline/build_reference_cuda.py), so it is the reference kernel's actual
 *     integer-deciding arithmetic.  fmaf()/fma() are exact here whatever
 *     -ffp-contract says;
 *   - the point validity test (cuh:293);
 *   - floor / fractional split and the per-corner bounds test (cuh:43-50, 60-83);
 *   - flat element offsets: level base `level_start*M*D` (cuh:279-283), row
 *     stride `W*M*D`, pixel stride `M*D`, channel base `m*D+c` (cuh:52-58);
 *   - forward accumulation `sum_l sum_p weight * bilinear` (cuh:285-302);
 *   - backward: grad_value scatter, grad_attn_weight and grad_sampling_loc
 *     (scaled by W and H) (cuh:119-163), zero for invalid points (cuh:370-372).
 *
 * Parity pin: this file is checked (tests/test_oracle.py) against golden
 * vectors produced by running the reference's own
 * ms_deform_attn_core_pytorch (ops/functions/ms_deform_attn_func.py:55-75) in
 * fp64 inside the build container (tests/golden/make_golden.py).
 *
 * Layouts (all contiguous, row-major):
 *   value  [N, S, M, D]      loc [N, Lq, M, L, P, 2] (x then y)
 *   weight [N, Lq, M, L, P]  out [N, Lq, M, D]
 *   shapes [L, 2] int64 (H, W)      level_start [L] int64
 *
 * The real type is chosen with -DREAL=float|double through the two wrapper
 * translation units at the bottom (the file includes itself).
 */
#ifndef MSDA_ORACLE_BODY

#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#define MSDA_ORACLE_BODY
/* REAL: type of value / weights / accumulation.  GEO: type the sampling-point
 * geometry (pixel coordinate, floor, fractions) is computed in.
 *   f32     = (float,  float)   the reference CUDA kernel's fp32 arithmetic
 *   f64     = (double, double)  the float golden
 *   f64g32  = (double, float)   fp32 geometry exactly as the reference kernel computes
 *             it (its `loc*W - 0.5` is one fp32 FMA, cuh:290-291), everything after
 *             that in fp64: "the reference fp32 kernel with exact accumulation".  This
 *             is what an fp32 implementation can be held to 1e-5 against on level
 *             sizes that are not powers of two, where the fp32 product loc*W is
 *             inexact and the reference's own fp32 paths sit ~1.7e-5 from fp64. */
#define REAL float
#define GEO float
#define SUFFIX(name) name##_f32
#define FLOOR floorf
#define FMA fmaf
#include "msda_oracle.c"
#undef REAL
#undef GEO
#undef SUFFIX
#undef FLOOR
#undef FMA

#define REAL double
#define GEO double
#define SUFFIX(name) name##_f64
#define FLOOR floor
#define FMA fma
#include "msda_oracle.c"
#undef REAL
#undef GEO
#undef SUFFIX
#undef FLOOR
#undef FMA

#define REAL double
#define GEO float
#define SUFFIX(name) name##_f64g32
#define FLOOR floorf
#define FMA fmaf
#include "msda_oracle.c"
#undef REAL
#undef GEO
#undef SUFFIX
#undef FLOOR
#undef FMA

int msda_oracle_abi_version(void) { return 1; }

#else /* MSDA_ORACLE_BODY: one instantiation for REAL */

/* Decomposition of one sampling point.  Everything integer here is what the
 * KATs pin bit-exactly against the CUDA path's debug entry. */
typedef struct {
    int valid;           /* point passes the range test (cuh:293)            */
    int h_low, w_low;    /* floor of the pixel coordinate (cuh:43-44)        */
    int cmask;           /* bit k set <=> corner k is inside the level:
                            k=0 (h_low,w_low) 1 (h_low,w_high)
                            k=2 (h_high,w_low) 3 (h_high,w_high)             */
    REAL lh, lw;         /* fractional parts (cuh:48-49)                     */
} SUFFIX(point_t);

static inline SUFFIX(point_t)
SUFFIX(decompose)(REAL loc_x_, REAL loc_y_, int H, int W)
{
    SUFFIX(point_t) pt;
    const GEO loc_x = (GEO)loc_x_, loc_y = (GEO)loc_y_;
    /* one rounding: fused multiply-add, as the compiled reference kernel (cuh:290-291) */
    const GEO h_im = FMA(loc_y, (GEO)H, (GEO)-0.5);
    const GEO w_im = FMA(loc_x, (GEO)W, (GEO)-0.5);
    pt.valid = (h_im > -1 && w_im > -1 && h_im < H && w_im < W);
    pt.h_low = (int)FLOOR(h_im);
    pt.w_low = (int)FLOOR(w_im);
    pt.lh = (REAL)(h_im - (GEO)pt.h_low);
    pt.lw = (REAL)(w_im - (GEO)pt.w_low);
    const int h_high = pt.h_low + 1, w_high = pt.w_low + 1;
    pt.cmask = 0;
    if (pt.valid) {
        if (pt.h_low >= 0 && pt.w_low >= 0) pt.cmask |= 1;
        if (pt.h_low >= 0 && w_high <= W - 1) pt.cmask |= 2;
        if (h_high <= H - 1 && pt.w_low >= 0) pt.cmask |= 4;
        if (h_high <= H - 1 && w_high <= W - 1) pt.cmask |= 8;
    }
    return pt;
}

/* element offset of (pixel h,w ; head m ; channel 0) inside one batch image */
static inline int64_t
SUFFIX(pix_off)(int64_t level_start, int W, int M, int D, int h, int w, int m)
{
    return ((level_start + (int64_t)h * W + w) * M + m) * (int64_t)D;
}

/* Integer known-answer output: for every (n,q,m,l,p)
 *   idx[.,0]=valid idx[.,1]=h_low idx[.,2]=w_low idx[.,3]=cmask
 *   off[.,k] = flat element offset of corner k (channel 0) from the start of
 *              the whole value tensor, or -1 if that corner does not contribute. */
void SUFFIX(msda_oracle_indices)(
    const int64_t *shapes, const int64_t *level_start, const REAL *loc,
    int N, int S, int M, int D, int L, int Lq, int P,
    int32_t *idx, int64_t *off)
{
    const int64_t total = (int64_t)N * Lq * M;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < total; ++r) {
        const int64_t n = r / ((int64_t)Lq * M);
        for (int l = 0; l < L; ++l) {
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
            for (int p = 0; p < P; ++p) {
                const int64_t s = (r * L + l) * P + p;
                const int m = (int)(r % M);
                SUFFIX(point_t) pt = SUFFIX(decompose)(loc[2 * s], loc[2 * s + 1], H, W);
                idx[4 * s + 0] = pt.valid;
                idx[4 * s + 1] = pt.h_low;
                idx[4 * s + 2] = pt.w_low;
                idx[4 * s + 3] = pt.cmask;
                for (int k = 0; k < 4; ++k) {
                    const int h = pt.h_low + (k >> 1), w = pt.w_low + (k & 1);
                    off[4 * s + k] = (pt.cmask >> k & 1)
                        ? n * (int64_t)S * M * D + SUFFIX(pix_off)(level_start[l], W, M, D, h, w, m)
                        : -1;
                }
            }
        }
    }
}

void SUFFIX(msda_oracle_forward)(
    const REAL *value, const int64_t *shapes, const int64_t *level_start,
    const REAL *loc, const REAL *weight,
    int N, int S, int M, int D, int L, int Lq, int P, REAL *out)
{
    const int64_t total = (int64_t)N * Lq * M;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < total; ++r) {
        const int64_t n = r / ((int64_t)Lq * M);
        const int m = (int)(r % M);
        const REAL *vb = value + n * (int64_t)S * M * D;
        REAL *o = out + r * D;
        for (int c = 0; c < D; ++c) o[c] = 0;
        for (int l = 0; l < L; ++l) {
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
            for (int p = 0; p < P; ++p) {
                const int64_t s = (r * L + l) * P + p;
                SUFFIX(point_t) pt = SUFFIX(decompose)(loc[2 * s], loc[2 * s + 1], H, W);
                if (!pt.valid) continue;
                const REAL a = weight[s];
                const REAL hh = 1 - pt.lh, hw = 1 - pt.lw;
                const REAL cw[4] = { hh * hw, hh * pt.lw, pt.lh * hw, pt.lh * pt.lw };
                const REAL *cp[4];
                for (int k = 0; k < 4; ++k)
                    cp[k] = (pt.cmask >> k & 1)
                        ? vb + SUFFIX(pix_off)(level_start[l], W, M, D,
                                               pt.h_low + (k >> 1), pt.w_low + (k & 1), m)
                        : NULL;
                for (int c = 0; c < D; ++c) {
                    REAL v[4];
                    for (int k = 0; k < 4; ++k) v[k] = cp[k] ? cp[k][c] : 0;
                    /* same association as cuh:85-87, then `* weight` (cuh:295) */
                    const REAL val = cw[0] * v[0] + cw[1] * v[1] + cw[2] * v[2] + cw[3] * v[3];
                    o[c] += val * a;
                }
            }
        }
    }
}

/* Backward.  grad_value must be zero on entry (the reference allocates it with
 * zeros, ms_deform_attn_cuda.cu:126).  The scatter is sequential per image so
 * the result is deterministic; images run in parallel. */
void SUFFIX(msda_oracle_backward)(
    const REAL *grad_out, const REAL *value, const int64_t *shapes,
    const int64_t *level_start, const REAL *loc, const REAL *weight,
    int N, int S, int M, int D, int L, int Lq, int P,
    REAL *grad_value, REAL *grad_loc, REAL *grad_weight)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t n = 0; n < N; ++n) {
        const REAL *vb = value + n * (int64_t)S * M * D;
        REAL *gvb = grad_value + n * (int64_t)S * M * D;
        for (int64_t qm = 0; qm < (int64_t)Lq * M; ++qm) {
            const int64_t r = n * (int64_t)Lq * M + qm;
            const int m = (int)(qm % M);
            const REAL *go = grad_out + r * D;
            for (int l = 0; l < L; ++l) {
                const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
                for (int p = 0; p < P; ++p) {
                    const int64_t s = (r * L + l) * P + p;
                    SUFFIX(point_t) pt = SUFFIX(decompose)(loc[2 * s], loc[2 * s + 1], H, W);
                    REAL g_a = 0, g_x = 0, g_y = 0;
                    if (pt.valid) {
                        const REAL a = weight[s];
                        const REAL hh = 1 - pt.lh, hw = 1 - pt.lw;
                        const REAL cw[4] = { hh * hw, hh * pt.lw, pt.lh * hw, pt.lh * pt.lw };
                        /* d val / d h and d val / d w per corner (cuh:128-156) */
                        const REAL dh[4] = { -hw, -pt.lw, hw, pt.lw };
                        const REAL dw[4] = { -hh, hh, -pt.lh, pt.lh };
                        for (int c = 0; c < D; ++c) {
                            const REAL tg = go[c];
                            const REAL tgv = tg * a;           /* cuh:120 */
                            REAL val = 0, gh = 0, gw = 0;
                            for (int k = 0; k < 4; ++k) {
                                if (!(pt.cmask >> k & 1)) continue;
                                const int64_t o = SUFFIX(pix_off)(level_start[l], W, M, D,
                                    pt.h_low + (k >> 1), pt.w_low + (k & 1), m) + c;
                                const REAL v = vb[o];
                                gh += dh[k] * v;
                                gw += dw[k] * v;
                                val += cw[k] * v;
                                gvb[o] += cw[k] * tgv;         /* cuh:130,139,148,157 */
                            }
                            g_a += tg * val;                   /* cuh:161 */
                            g_x += (REAL)W * gw * tgv;         /* cuh:162 */
                            g_y += (REAL)H * gh * tgv;         /* cuh:163 */
                        }
                    }
                    grad_weight[s] = g_a;
                    grad_loc[2 * s] = g_x;
                    grad_loc[2 * s + 1] = g_y;
                }
            }
        }
    }
}

#endif /* MSDA_ORACLE_BODY */
