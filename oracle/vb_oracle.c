/* vb_oracle.c — see vb_oracle.h. TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Build with -ffp-contract=off and without -ffast-math: every '*' and '+' below is meant to be one
 * IEEE-754 rounding, because the reference's OpenCV calls round after each element-wise op
 * (checked against cv2 4.13.0, tests/golden/gen_golden.py).
 */
#include "vb_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ============================ score reduction order ========================================== */

double vbo_score_sum(const float *e, int n) {
    const int group_elems = VBO_SUM_CHUNK * VBO_SUM_GROUP;
    double total = 0.0;
    for (int g0 = 0; g0 < n; g0 += group_elems) {
        int g1 = g0 + group_elems < n ? g0 + group_elems : n;
        double gs = 0.0;
        for (int c0 = g0; c0 < g1; c0 += VBO_SUM_CHUNK) {
            int c1 = c0 + VBO_SUM_CHUNK < g1 ? c0 + VBO_SUM_CHUNK : g1;
            double cs = 0.0;
            for (int i = c0; i < c1; i++) cs = cs + (double)e[i];
            gs = gs + cs;
        }
        total = total + gs;
    }
    return total;
}

/* ============================ mt19937 / uniform_int_distribution ============================== */

void vbo_mt_seed(vbo_mt19937 *g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; i++) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

static void mt_twist(vbo_mt19937 *g) {
    uint32_t *mt = g->mt;
    for (int k = 0; k < 624; k++) {
        uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
}

uint32_t vbo_mt_next(vbo_mt19937 *g) {
    if (g->idx >= 624) mt_twist(g);
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* libstdc++ 13 bits/uniform_int_dist.h:_S_nd (Lemire multiply-shift with rejection), reached because
 * mt19937's range is exactly 2^32-1. range = hi + 1. */
int vbo_uniform_int(vbo_mt19937 *g, int hi) {
    uint32_t range = (uint32_t)hi + 1u;
    uint64_t product = (uint64_t)vbo_mt_next(g) * (uint64_t)range;
    uint32_t low = (uint32_t)product;
    if (low < range) {
        uint32_t threshold = (0u - range) % range;
        while (low < threshold) {
            product = (uint64_t)vbo_mt_next(g) * (uint64_t)range;
            low = (uint32_t)product;
        }
    }
    return (int)(product >> 32);
}

/* src/RansacFilter.cpp:6-34. The reference copies the whole 0..n-1 pool per hypothesis (:20) and
 * removes a drawn slot by overwriting it with the last (:28-31); only <= 8 slots ever differ from
 * the identity, so a small displacement list reproduces it. */
void vbo_initialize_sets(int n_matches, int min_items, int max_iterations, uint32_t seed, int32_t *sets) {
    vbo_mt19937 g;
    vbo_mt_seed(&g, seed);
    for (int i = 0; i < max_iterations; i++) {
        int32_t pos[8], val[8];
        int nd = 0;
        int size = n_matches;
        for (int j = 0; j < 8; j++) sets[i * 8 + j] = 0;
        for (int j = 0; j < min_items; j++) {
            int r = vbo_uniform_int(&g, size - 1);
            int vr = r, vlast = size - 1;
            for (int t = 0; t < nd; t++) {
                if (pos[t] == r) vr = val[t];
                if (pos[t] == size - 1) vlast = val[t];
            }
            sets[i * 8 + j] = vr;
            /* pool[r] = pool.back(); pool.pop_back(); */
            int found = 0;
            for (int t = 0; t < nd; t++)
                if (pos[t] == r) { val[t] = vlast; found = 1; }
            if (!found) { pos[nd] = r; val[nd] = vlast; nd++; }
            size--;
        }
    }
}

/* ============================ 8-point solve =================================================== */

void vbo_null_vector_8x9(const float *A, float *f9) {
    /* B = A^T (9x8); Householder QR; null vector = Q e_8. */
    double B[9][8], V[8][9], beta[8];
    for (int i = 0; i < 9; i++)
        for (int k = 0; k < 8; k++) B[i][k] = (double)A[k * 9 + i];
    for (int k = 0; k < 8; k++) {
        double sigma = 0.0;
        for (int i = k; i < 9; i++) sigma = sigma + B[i][k] * B[i][k];
        double norm = sqrt(sigma);
        double alpha = (B[k][k] >= 0.0) ? -norm : norm;
        for (int i = 0; i < 9; i++) V[k][i] = 0.0;
        V[k][k] = B[k][k] - alpha;
        for (int i = k + 1; i < 9; i++) V[k][i] = B[i][k];
        double vtv = 0.0;
        for (int i = k; i < 9; i++) vtv = vtv + V[k][i] * V[k][i];
        beta[k] = (vtv > 0.0) ? 2.0 / vtv : 0.0;
        for (int j = k + 1; j < 8; j++) {
            double dot = 0.0;
            for (int i = k; i < 9; i++) dot = dot + V[k][i] * B[i][j];
            double t = beta[k] * dot;
            for (int i = k; i < 9; i++) B[i][j] = B[i][j] - t * V[k][i];
        }
    }
    double z[9] = {0, 0, 0, 0, 0, 0, 0, 0, 1.0};
    for (int k = 7; k >= 0; k--) {
        double dot = 0.0;
        for (int i = k; i < 9; i++) dot = dot + V[k][i] * z[i];
        double t = beta[k] * dot;
        for (int i = k; i < 9; i++) z[i] = z[i] - t * V[k][i];
    }
    for (int i = 0; i < 9; i++) f9[i] = (float)z[i];
}

#define VBO_SVD3_MAX_SWEEPS 30
#define VBO_SVD3_EPS 2.220446049250313e-16 /* 2^-52 */

void vbo_svd3x3(const float *F, float *U, float *D, float *Vt) {
    double G[3][3], V[3][3];
    static const int PQ[3][2] = {{0, 1}, {0, 2}, {1, 2}};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            G[i][j] = (double)F[i * 3 + j];
            V[i][j] = (i == j) ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < VBO_SVD3_MAX_SWEEPS; sweep++) {
        int rotated = 0;
        for (int r = 0; r < 3; r++) {
            int p = PQ[r][0], q = PQ[r][1];
            double alpha = (G[0][p] * G[0][p] + G[1][p] * G[1][p]) + G[2][p] * G[2][p];
            double bet = (G[0][q] * G[0][q] + G[1][q] * G[1][q]) + G[2][q] * G[2][q];
            double gamma = (G[0][p] * G[0][q] + G[1][p] * G[1][q]) + G[2][p] * G[2][q];
            if (gamma == 0.0 || fabs(gamma) <= VBO_SVD3_EPS * sqrt(alpha * bet)) continue;
            rotated = 1;
            double zeta = (bet - alpha) / (2.0 * gamma);
            double t = 1.0 / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            if (zeta < 0.0) t = -t;
            double c = 1.0 / sqrt(1.0 + t * t);
            double s = c * t;
            for (int k = 0; k < 3; k++) {
                double gp = G[k][p], gq = G[k][q];
                G[k][p] = c * gp - s * gq;
                G[k][q] = s * gp + c * gq;
                double vp = V[k][p], vq = V[k][q];
                V[k][p] = c * vp - s * vq;
                V[k][q] = s * vp + c * vq;
            }
        }
        if (!rotated) break;
    }
    double sv[3];
    int ord[3] = {0, 1, 2};
    for (int j = 0; j < 3; j++) sv[j] = sqrt((G[0][j] * G[0][j] + G[1][j] * G[1][j]) + G[2][j] * G[2][j]);
    /* stable insertion sort, descending */
    for (int a = 1; a < 3; a++)
        for (int b = a; b > 0 && sv[ord[b]] > sv[ord[b - 1]]; b--) {
            int tmp = ord[b]; ord[b] = ord[b - 1]; ord[b - 1] = tmp;
        }
    for (int j = 0; j < 3; j++) {
        int c = ord[j];
        double s = sv[c];
        D[j] = (float)s;
        for (int k = 0; k < 3; k++) {
            U[k * 3 + j] = (s > 0.0) ? (float)(G[k][c] / s) : 0.0f;
            Vt[j * 3 + k] = (float)V[k][c];
        }
    }
}

/* fp32 3x3 product in OpenCV's small-matrix gemm order: (a0*b0 + a1*b1) + a2*b2. */
static void mat3_mul_f32(const float *A, const float *B, float *C) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            float p0 = A[i * 3 + 0] * B[0 * 3 + j];
            float p1 = A[i * 3 + 1] * B[1 * 3 + j];
            float p2 = A[i * 3 + 2] * B[2 * 3 + j];
            float s = p0 + p1;
            C[i * 3 + j] = s + p2;
        }
}

void vbo_compute_fundamental(const float *p1set, const float *p2set, float *F) {
    float A[72];
    for (int i = 0; i < 8; i++) {
        const float u1 = p1set[2 * i], v1 = p1set[2 * i + 1];
        const float u2 = p2set[2 * i], v2 = p2set[2 * i + 1];
        float *r = A + 9 * i; /* src/RansacFilter.cpp:81-89 */
        r[0] = u2 * u1; r[1] = u2 * v1; r[2] = u2;
        r[3] = v2 * u1; r[4] = v2 * v1; r[5] = v2;
        r[6] = u1;      r[7] = v1;      r[8] = 1.0f;
    }
    float f9[9], U[9], D[3], Vt[9], Dg[9], T[9];
    vbo_null_vector_8x9(A, f9);      /* :94-95  F = Vt.row(8).reshape(0,3) */
    vbo_svd3x3(f9, U, D, Vt);        /* :98 */
    D[2] = 0.0f;                     /* :99 */
    for (int i = 0; i < 9; i++) Dg[i] = 0.0f;
    Dg[0] = D[0]; Dg[4] = D[1]; Dg[8] = D[2];
    mat3_mul_f32(U, Dg, T);          /* :101 U * diag(D) * V_t, left to right */
    mat3_mul_f32(T, Vt, F);
}

/* ============================ residual ======================================================== */

float vbo_residual_one(const float *F, float x1, float y1, float x2, float y2) {
    /* F * x1 (gemm, fp32) */
    float a0 = (F[0] * x1 + F[1] * y1) + F[2];
    float a1 = (F[3] * x1 + F[4] * y1) + F[5];
    float a2 = (F[6] * x1 + F[7] * y1) + F[8];
    /* F.t() * x2 (gemm GEMM_1_T: fp64 accumulate, round once) */
    float b0 = (float)(((double)F[0] * (double)x2 + (double)F[3] * (double)y2) + (double)F[6]);
    float b1 = (float)(((double)F[1] * (double)x2 + (double)F[4] * (double)y2) + (double)F[7]);
    /* reduce(x2.mul(F_x1), 0, SUM): (x2*a0 + y2*a1) + 1*a2 in fp32 */
    float s = (x2 * a0 + y2 * a1) + a2;
    /* :126 as parsed: ((s*s)/(a0*a0)) + a1*a1 + b0*b0 + b1*b1 */
    float num = s * s, den = a0 * a0;
    float e = num / den;
    e = e + a1 * a1;
    e = e + b0 * b0;
    e = e + b1 * b1;
    return e;
}

void vbo_compute_fundamental_residual(const float *p1, const float *p2, const int32_t *matches, int m,
                                      const float *F, float threshold, uint8_t *mask, float *e_out,
                                      int *n_inliers, float *score) {
    float *e = e_out ? e_out : (float *)malloc(sizeof(float) * (size_t)(m > 0 ? m : 1));
    int n = 0;
    for (int i = 0; i < m; i++) {
        const float *a = p1 + 2 * (size_t)matches[2 * i];
        const float *b = p2 + 2 * (size_t)matches[2 * i + 1];
        e[i] = vbo_residual_one(F, a[0], a[1], b[0], b[1]);
        int in = e[i] <= threshold; /* :130; NaN -> outlier */
        if (mask) mask[i] = (uint8_t)in;
        n += in;
    }
    *n_inliers = n;
    *score = (float)vbo_score_sum(e, m); /* :138 (float) cv::sum(e_sq)[0] */
    if (!e_out) free(e);
}

/* ============================ find_fundamental ================================================ */

int vbo_find_fundamental(const float *p1, const float *p2, const int32_t *matches, int m, int min_items,
                         int max_iterations, float threshold, uint32_t seed, float *F, uint8_t *mask,
                         int *n_inliers, float *score, int *best_hyp, int32_t *sets_out, float *F_all,
                         int32_t *cnt_all, float *score_all) {
    if (min_items < 1 || min_items > 8 || m < min_items || max_iterations < 0) return -1;
    int32_t *sets = (int32_t *)malloc(sizeof(int32_t) * 8 * (size_t)(max_iterations > 0 ? max_iterations : 1));
    uint8_t *cur = (uint8_t *)malloc((size_t)(m > 0 ? m : 1));
    vbo_initialize_sets(m, min_items, max_iterations, seed, sets); /* :38 */
    float best_score = 0.0f; /* :44 */
    int best_n = 0;          /* :45 */
    int best = -1;
    for (int i = 0; i < max_iterations; i++) { /* :49 */
        float s1[16], s2[16], Fi[9];
        for (int j = 0; j < 8; j++) { /* :50-54 — all 8 slots, whatever min_items is */
            int idx = sets[i * 8 + j];
            s1[2 * j] = p1[2 * (size_t)matches[2 * idx]];
            s1[2 * j + 1] = p1[2 * (size_t)matches[2 * idx] + 1];
            s2[2 * j] = p2[2 * (size_t)matches[2 * idx + 1]];
            s2[2 * j + 1] = p2[2 * (size_t)matches[2 * idx + 1] + 1];
        }
        vbo_compute_fundamental(s1, s2, Fi);
        int n;
        float sc;
        vbo_compute_fundamental_residual(p1, p2, matches, m, Fi, threshold, cur, NULL, &n, &sc);
        if (F_all) memcpy(F_all + 9 * (size_t)i, Fi, sizeof(Fi));
        if (cnt_all) cnt_all[i] = n;
        if (score_all) score_all[i] = sc;
        if (n > best_n || (n == best_n && sc > best_score)) { /* :59 */
            best_n = n;
            best_score = sc;
            best = i;
            memcpy(F, Fi, sizeof(Fi));
            if (mask) memcpy(mask, cur, (size_t)m);
        }
    }
    if (sets_out) memcpy(sets_out, sets, sizeof(int32_t) * 8 * (size_t)max_iterations);
    if (n_inliers) *n_inliers = best_n;
    if (score) *score = best_score;
    if (best_hyp) *best_hyp = best;
    free(sets);
    free(cur);
    return 0;
}

/* ============================ opt-in mode: Hartley normalisation + true Sampson distance ====== */
/* NOT reference behaviour: the two repairs the reference itself flags and never made — `//TODO: normalize`
 * (src/RansacFilter.cpp:40) and the mis-parenthesised residual (:125-126) — as an explicit opt-in (flags != 0). The default
 * path above is untouched. Every operation below is one IEEE rounding in the order written; the GPU repeats it exactly.
 *
 * VBO_RANSAC_HARTLEY  each 8-point sample is translated to its centroid and scaled to mean distance sqrt(2) per image
 *                     (Hartley 1997) before the same 8-point solve, and F is mapped back: F = T2^T F^ T1, unit Frobenius norm.
 * VBO_RANSAC_SAMPSON  e = (x2^T F x1)^2 / (a0^2 + a1^2 + b0^2 + b1^2), a = F x1, b = F^T x2 — the first-order geometric
 *                     error the reference's expression was meant to be — evaluated in double, narrowed to f32 once; the
 *                     inlier test (e <= threshold), the score sum and the selection rule stay the reference's. */
static void hartley_norm(const float *pts /* [8][2] */, float *out /* [8][2] */, double *s_out, double *tx, double *ty) {
    double cx = 0.0, cy = 0.0;
    for (int i = 0; i < 8; i++) { cx = cx + (double)pts[2 * i]; cy = cy + (double)pts[2 * i + 1]; }
    cx = cx / 8.0; cy = cy / 8.0;
    double md = 0.0;
    for (int i = 0; i < 8; i++) {
        const double dx = (double)pts[2 * i] - cx, dy = (double)pts[2 * i + 1] - cy;
        md = md + sqrt(dx * dx + dy * dy);
    }
    md = md / 8.0;
    const double s = (md > 0.0) ? 1.4142135623730951 / md : 1.0;
    for (int i = 0; i < 8; i++) {
        out[2 * i] = (float)(s * ((double)pts[2 * i] - cx));
        out[2 * i + 1] = (float)(s * ((double)pts[2 * i + 1] - cy));
    }
    *s_out = s; *tx = s * cx; *ty = s * cy;
}

void vbo_compute_fundamental_hartley(const float *p1set, const float *p2set, float *F) {
    float n1[16], n2[16], Fh[9];
    double s1, tx1, ty1, s2, tx2, ty2;
    hartley_norm(p1set, n1, &s1, &tx1, &ty1);
    hartley_norm(p2set, n2, &s2, &tx2, &ty2);
    vbo_compute_fundamental(n1, n2, Fh);
    /* G = F^ T1, T = [[s, 0, -tx], [0, s, -ty], [0, 0, 1]] */
    double G[9], R[9];
    for (int i = 0; i < 3; i++) {
        const double f0 = (double)Fh[3 * i], f1 = (double)Fh[3 * i + 1], f2 = (double)Fh[3 * i + 2];
        G[3 * i] = f0 * s1;
        G[3 * i + 1] = f1 * s1;
        G[3 * i + 2] = (f2 - f0 * tx1) - f1 * ty1;
    }
    /* R = T2^T G */
    for (int j = 0; j < 3; j++) {
        R[j] = s2 * G[j];
        R[3 + j] = s2 * G[3 + j];
        R[6 + j] = (G[6 + j] - tx2 * G[j]) - ty2 * G[3 + j];
    }
    double nn = 0.0;
    for (int i = 0; i < 9; i++) nn = nn + R[i] * R[i];
    const double nrm = sqrt(nn);
    for (int i = 0; i < 9; i++) F[i] = (float)((nrm > 0.0) ? R[i] / nrm : R[i]);
}

float vbo_sampson_one(const float *F, float x1f, float y1f, float x2f, float y2f) {
    const double x1 = x1f, y1 = y1f, x2 = x2f, y2 = y2f;
    double f[9];
    for (int i = 0; i < 9; i++) f[i] = (double)F[i];
    const double a0 = (f[0] * x1 + f[1] * y1) + f[2];
    const double a1 = (f[3] * x1 + f[4] * y1) + f[5];
    const double a2 = (f[6] * x1 + f[7] * y1) + f[8];
    const double b0 = (f[0] * x2 + f[3] * y2) + f[6];
    const double b1 = (f[1] * x2 + f[4] * y2) + f[7];
    const double sv = (x2 * a0 + y2 * a1) + a2;
    const double den = ((a0 * a0 + a1 * a1) + b0 * b0) + b1 * b1;
    return (float)((sv * sv) / den);
}

int vbo_find_fundamental_ex(const float *p1, const float *p2, const int32_t *matches, int m, int min_items, int max_iterations,
                            float threshold, uint32_t seed, unsigned flags, float *F, uint8_t *mask, int *n_inliers, float *score,
                            int *best_hyp, float *F_all, int32_t *cnt_all, float *score_all) {
    if (flags == 0)
        return vbo_find_fundamental(p1, p2, matches, m, min_items, max_iterations, threshold, seed, F, mask, n_inliers, score,
                                    best_hyp, NULL, F_all, cnt_all, score_all);
    if (min_items < 1 || min_items > 8 || m < min_items || max_iterations < 0) return -1;
    int32_t *sets = (int32_t *)malloc(sizeof(int32_t) * 8 * (size_t)(max_iterations > 0 ? max_iterations : 1));
    uint8_t *cur = (uint8_t *)malloc((size_t)(m > 0 ? m : 1));
    float *e = (float *)malloc(sizeof(float) * (size_t)(m > 0 ? m : 1));
    vbo_initialize_sets(m, min_items, max_iterations, seed, sets);
    float best_score = 0.0f;
    int best_n = 0, best = -1;
    for (int i = 0; i < max_iterations; i++) {
        float s1[16], s2[16], Fi[9];
        for (int j = 0; j < 8; j++) {
            int idx = sets[i * 8 + j];
            s1[2 * j] = p1[2 * (size_t)matches[2 * idx]];
            s1[2 * j + 1] = p1[2 * (size_t)matches[2 * idx] + 1];
            s2[2 * j] = p2[2 * (size_t)matches[2 * idx + 1]];
            s2[2 * j + 1] = p2[2 * (size_t)matches[2 * idx + 1] + 1];
        }
        if (flags & VBO_RANSAC_HARTLEY) vbo_compute_fundamental_hartley(s1, s2, Fi);
        else vbo_compute_fundamental(s1, s2, Fi);
        int n = 0;
        for (int k = 0; k < m; k++) {
            const float *a = p1 + 2 * (size_t)matches[2 * k];
            const float *b = p2 + 2 * (size_t)matches[2 * k + 1];
            e[k] = (flags & VBO_RANSAC_SAMPSON) ? vbo_sampson_one(Fi, a[0], a[1], b[0], b[1])
                                                : vbo_residual_one(Fi, a[0], a[1], b[0], b[1]);
            cur[k] = (uint8_t)(e[k] <= threshold);
            n += cur[k];
        }
        const float sc = (float)vbo_score_sum(e, m);
        if (F_all) memcpy(F_all + 9 * (size_t)i, Fi, sizeof(Fi));
        if (cnt_all) cnt_all[i] = n;
        if (score_all) score_all[i] = sc;
        if (n > best_n || (n == best_n && sc > best_score)) {
            best_n = n; best_score = sc; best = i;
            memcpy(F, Fi, sizeof(Fi));
            if (mask) memcpy(mask, cur, (size_t)m);
        }
    }
    if (n_inliers) *n_inliers = best_n;
    if (score) *score = best_score;
    if (best_hyp) *best_hyp = best;
    free(sets); free(cur); free(e);
    return 0;
}

/* ============================ matcher ========================================================= */

static inline int hamming_bytes(const uint8_t *a, const uint8_t *b, int bytes) {
    int d = 0, k = 0;
    for (; k + 8 <= bytes; k += 8) {
        uint64_t x, y;
        memcpy(&x, a + k, 8);
        memcpy(&y, b + k, 8);
        d += __builtin_popcountll(x ^ y);
    }
    for (; k < bytes; k++) d += __builtin_popcount((unsigned)(a[k] ^ b[k]));
    return d;
}

void vbo_knn2_hamming(const uint8_t *d1, int n1, const uint8_t *d2, int n2, int bytes, int32_t *idx,
                      int32_t *dist) {
    for (int q = 0; q < n1; q++) {
        int b0 = 0x7fffffff, b1 = 0x7fffffff, i0 = -1, i1 = -1;
        const uint8_t *a = d1 + (size_t)q * bytes;
        for (int t = 0; t < n2; t++) {
            int d = hamming_bytes(a, d2 + (size_t)t * bytes, bytes);
            if (d < b0) { b1 = b0; i1 = i0; b0 = d; i0 = t; }
            else if (d < b1) { b1 = d; i1 = t; }
        }
        idx[2 * q] = i0; idx[2 * q + 1] = i1;
        dist[2 * q] = b0; dist[2 * q + 1] = b1;
    }
}

int vbo_ratio_keep(int d0, int d1, double ratio) {
    float f0 = (float)d0, f1 = (float)d1; /* DMatch::distance is float */
    return (double)f0 < (double)f1 * ratio; /* src/Frame.cpp:91 */
}

int vbo_match_hamming(const uint8_t *d1, int n1, const uint8_t *d2, int n2, int bytes, double ratio,
                      int32_t *out_pairs) {
    if (n2 < 2) return 0; /* reference: m[1] out of bounds (UB) */
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(n1 > 0 ? n1 : 1));
    int32_t *dist = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(n1 > 0 ? n1 : 1));
    vbo_knn2_hamming(d1, n1, d2, n2, bytes, idx, dist);
    int m = 0;
    for (int q = 0; q < n1; q++)
        if (vbo_ratio_keep(dist[2 * q], dist[2 * q + 1], ratio)) {
            out_pairs[2 * m] = q;
            out_pairs[2 * m + 1] = idx[2 * q];
            m++;
        }
    free(idx);
    free(dist);
    return m;
}

void vbo_knn2_l2f(const float *d1, int n1, const float *d2, int n2, int dim, int32_t *idx, float *dist) {
    for (int q = 0; q < n1; q++) {
        float b0 = INFINITY, b1 = INFINITY;
        int i0 = -1, i1 = -1;
        const float *a = d1 + (size_t)q * dim;
        for (int t = 0; t < n2; t++) {
            const float *b = d2 + (size_t)t * dim;
            float acc = 0.0f;
            for (int k = 0; k < dim; k++) {
                float df = a[k] - b[k];
                acc = acc + df * df;
            }
            if (acc < b0) { b1 = b0; i1 = i0; b0 = acc; i0 = t; }
            else if (acc < b1) { b1 = acc; i1 = t; }
        }
        idx[2 * q] = i0; idx[2 * q + 1] = i1;
        dist[2 * q] = sqrtf(b0); dist[2 * q + 1] = sqrtf(b1);
    }
}

int vbo_match_l2f(const float *d1, int n1, const float *d2, int n2, int dim, double ratio, int32_t *out_pairs) {
    if (n2 < 2) return 0;
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(n1 > 0 ? n1 : 1));
    float *dist = (float *)malloc(sizeof(float) * 2 * (size_t)(n1 > 0 ? n1 : 1));
    vbo_knn2_l2f(d1, n1, d2, n2, dim, idx, dist);
    int m = 0;
    for (int q = 0; q < n1; q++)
        if ((double)dist[2 * q] < (double)dist[2 * q + 1] * ratio) {
            out_pairs[2 * m] = q;
            out_pairs[2 * m + 1] = idx[2 * q];
            m++;
        }
    free(idx);
    free(dist);
    return m;
}

int vbo_match_features(const float *p1, const uint8_t *d1, int n1, const float *p2, const uint8_t *d2,
                       int n2, int bytes, double ratio, int min_items, int max_iterations,
                       float threshold, uint32_t seed, int32_t *out_matches, float *F, int *n_tentative,
                       int *best_hyp) {
    int32_t *tent = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(n1 > 0 ? n1 : 1));
    int m = vbo_match_hamming(d1, n1, d2, n2, bytes, ratio, tent);
    if (n_tentative) *n_tentative = m;
    uint8_t *mask = (uint8_t *)calloc((size_t)(m > 0 ? m : 1), 1);
    int n_inl = 0, bh = -1;
    float sc;
    int rc = vbo_find_fundamental(p1, p2, tent, m, min_items, max_iterations, threshold, seed, F, mask,
                                  &n_inl, &sc, &bh, NULL, NULL, NULL, NULL);
    if (best_hyp) *best_hyp = bh;
    int out = -1;
    if (rc == 0) {
        out = 0;
        if (bh >= 0) /* inliers stays empty in the reference when no hypothesis was ever accepted */
            for (int i = 0; i < m; i++)
                if (mask[i]) { /* src/Frame.cpp:98-102 */
                    out_matches[2 * out] = tent[2 * i];
                    out_matches[2 * out + 1] = tent[2 * i + 1];
                    out++;
                }
    }
    free(tent);
    free(mask);
    return out;
}

/* ============================ KD-tree ========================================================= */

typedef struct { const float *pts; int axis; } kd_cmp_ctx;
static __thread kd_cmp_ctx g_kd_ctx;

static int kd_cmp(const void *a, const void *b) {
    int ia = *(const int32_t *)a, ib = *(const int32_t *)b;
    float ca = g_kd_ctx.pts[2 * (size_t)ia + g_kd_ctx.axis], cb = g_kd_ctx.pts[2 * (size_t)ib + g_kd_ctx.axis];
    if (ca < cb) return -1;
    if (cb < ca) return 1;
    return (ia > ib) - (ia < ib);
}

static void kd_build_rec(const float *pts, int32_t *work, int l, int r, int axis, int32_t *pre, int *slot) {
    if (l >= r) return; /* src/KDTree.cpp:6 */
    int len = r - l, m = l + len / 2; /* :7-8 */
    g_kd_ctx.pts = pts;
    g_kd_ctx.axis = axis;
    qsort(work + l, (size_t)len, sizeof(int32_t), kd_cmp); /* stands for nth_element, :10-12 */
    pre[(*slot)++] = work[m];                               /* :16-17 */
    kd_build_rec(pts, work, l, m, 1 - axis, pre, slot);     /* :18 */
    kd_build_rec(pts, work, m + 1, r, 1 - axis, pre, slot); /* :19 */
}

void vbo_kdtree_build(const float *pts, int n, int32_t *pre_idx) {
    if (n <= 0) return;
    int32_t *work = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    for (int i = 0; i < n; i++) work[i] = i;
    int slot = 0;
    kd_build_rec(pts, work, 0, n, 0, pre_idx, &slot);
    free(work);
}

int vbo_kdtree_height(int n) { return n > 0 ? (int)floor(log2((double)n)) + 1 : 0; } /* :33 */

typedef struct {
    const float *pts;
    const int32_t *pre;
    float qx, qy;
    float best;
    int best_slot;
} kd_nn;

static void kd_nearest_rec(kd_nn *s, int slot, int len, int axis) {
    if (len <= 0) return; /* :47 */
    const float *pt = s->pts + 2 * (size_t)s->pre[slot];
    const float q_ax = axis ? s->qy : s->qx;
    const float split = q_ax - pt[axis]; /* :51 */
    const int llen = len / 2, rlen = len - llen - 1;
    const int lslot = slot + 1, rslot = slot + 1 + llen;
    int far_slot, far_len;
    if (split < 0) { /* :54-60 */
        kd_nearest_rec(s, lslot, llen, 1 - axis);
        far_slot = rslot; far_len = rlen;
    } else {
        kd_nearest_rec(s, rslot, rlen, 1 - axis);
        far_slot = lslot; far_len = llen;
    }
    float dx = pt[0] - s->qx, dy = pt[1] - s->qy; /* :62-63 */
    float d2 = dx * dx + dy * dy;
    if (d2 < s->best) { s->best = d2; s->best_slot = slot; } /* :64-67 */
    if (split * split < s->best) kd_nearest_rec(s, far_slot, far_len, 1 - axis); /* :68-70 */
}

int vbo_kdtree_nearest(const float *pts, const int32_t *pre_idx, int n, float qx, float qy, float max_d2,
                       float *out_d2) {
    kd_nn s = {pts, pre_idx, qx, qy, max_d2, -1};
    kd_nearest_rec(&s, 0, n, 0);
    if (out_d2) *out_d2 = s.best;
    return s.best_slot;
}

/* k nearest neighbours (k > 1). The reference only has commented-out declarations for this (include/KDTree.h:39-42,
 * 74-77), so the behaviour is DEFINED here as the direct generalisation of `nearest` (src/KDTree.cpp:45-71): the same
 * visiting order (near child, node, far child), the bound of the strict `<` tests (:64, :68) is the k-th best squared
 * distance so far (max_distance_sq until k candidates exist), and a candidate goes into the ascending list BEHIND entries
 * of equal distance — so among equidistant points the first visited wins, exactly as for k = 1. k = 1 is vbo_kdtree_nearest. */
typedef struct {
    const float *pts;
    const int32_t *pre;
    float qx, qy, max_d2;
    int k, cnt;
    int32_t *slot;
    float *d2;
} kd_knn;

static void kd_knn_rec(kd_knn *s, int slot, int len, int axis) {
    if (len <= 0) return;
    const float *pt = s->pts + 2 * (size_t)s->pre[slot];
    const float q_ax = axis ? s->qy : s->qx;
    const float split = q_ax - pt[axis];
    const int llen = len / 2, rlen = len - llen - 1;
    const int lslot = slot + 1, rslot = slot + 1 + llen;
    int far_slot, far_len;
    if (split < 0) {
        kd_knn_rec(s, lslot, llen, 1 - axis);
        far_slot = rslot; far_len = rlen;
    } else {
        kd_knn_rec(s, rslot, rlen, 1 - axis);
        far_slot = lslot; far_len = llen;
    }
    float dx = pt[0] - s->qx, dy = pt[1] - s->qy;
    float d2 = dx * dx + dy * dy;
    float bound = (s->cnt < s->k) ? s->max_d2 : s->d2[s->k - 1];
    if (d2 < bound) {
        int j = (s->cnt < s->k) ? s->cnt : s->k - 1;   /* position being filled: a new tail, or the evicted k-th */
        while (j > 0 && s->d2[j - 1] > d2) { s->d2[j] = s->d2[j - 1]; s->slot[j] = s->slot[j - 1]; j--; }
        s->d2[j] = d2; s->slot[j] = slot;
        if (s->cnt < s->k) s->cnt++;
    }
    bound = (s->cnt < s->k) ? s->max_d2 : s->d2[s->k - 1];
    if (split * split < bound) kd_knn_rec(s, far_slot, far_len, 1 - axis);
}

int vbo_kdtree_knn(const float *pts, const int32_t *pre_idx, int n, float qx, float qy, int k, float max_d2,
                   int32_t *out_slot, float *out_d2) {
    kd_knn s = {pts, pre_idx, qx, qy, max_d2, k, 0, out_slot, out_d2};
    if (k <= 0) return 0;
    kd_knn_rec(&s, 0, n, 0);
    return s.cnt;
}

typedef struct {
    const float *pts;
    const int32_t *pre;
    float qx, qy, r, r2;
    int32_t *out;
    int cap, cnt;
} kd_rs;

static void kd_radius_rec(kd_rs *s, int slot, int len, int axis) {
    if (len <= 0) return;
    const int pi = s->pre[slot];
    const float *pt = s->pts + 2 * (size_t)pi;
    const float q_ax = axis ? s->qy : s->qx;
    const float split = q_ax - pt[axis];
    const int llen = len / 2, rlen = len - llen - 1;
    const float as = (split > 0) ? split : -split; /* ABS macro, include/KDTree.h:10 */
    if (as <= s->r) { /* :88 */
        float dx = s->qx - pt[0], dy = s->qy - pt[1];
        float d2 = dx * dx + dy * dy;
        if (d2 < s->r2) { /* :91 */
            if (s->cnt < s->cap) s->out[s->cnt] = pi;
            s->cnt++;
        }
        kd_radius_rec(s, slot + 1, llen, 1 - axis);
        kd_radius_rec(s, slot + 1 + llen, rlen, 1 - axis);
    } else if (split < 0) {
        kd_radius_rec(s, slot + 1, llen, 1 - axis);
    } else {
        kd_radius_rec(s, slot + 1 + llen, rlen, 1 - axis);
    }
}

int vbo_kdtree_radius(const float *pts, const int32_t *pre_idx, int n, float qx, float qy, float radius,
                      int32_t *out_idx, int cap) {
    kd_rs s = {pts, pre_idx, qx, qy, radius, radius * radius, out_idx, cap, 0};
    kd_radius_rec(&s, 0, n, 0);
    return s.cnt;
}

/* ============================ next rows ======================================================= */

uint32_t vbo_orb_distance(const uint8_t *desc, const uint8_t *obs, int k, int bytes) {
    uint32_t mn = 0xffffffffu; /* u32_max, src/PointMap.cpp:37 */
    for (int i = 0; i < k; i++) {
        uint32_t d = (uint32_t)hamming_bytes(desc, obs + (size_t)i * bytes, bytes);
        if (d < mn) mn = d;
    }
    return mn;
}

/* Projection of homogeneous map points by a 3x4 camera, as cv::Mat `points * c2.t()` evaluates it
 * (src/vslam.cpp:131; MatExpr -> cv::gemm(points, c2, 1, noArray(), 0, GEMM_2_T)). Observed on cv2 4.13.0
 * (tests/golden, `projection_*`): fewer than 100 rows take OpenCV's own kernel — products and sums in
 * double, one rounding to float; 100 rows or more are handed to the BLAS sgemm, which here evaluates
 * ((x0*c0 + x1*c1) + x2*c2) + x3*c3 in fp32, one rounding per operation. */
void vbo_project_points(const float *X, int n, const float *c2, float *out3) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 3; j++) {
            const float *x = X + (size_t)i * 4, *c = c2 + (size_t)j * 4;
            float r;
            if (n < 100) {
                double acc = (double)x[0] * (double)c[0];
                acc += (double)x[1] * (double)c[1];
                acc += (double)x[2] * (double)c[2];
                acc += (double)x[3] * (double)c[3];
                r = (float)acc;
            } else {
                float acc = x[0] * c[0];
                acc = acc + x[1] * c[1];
                acc = acc + x[2] * c[2];
                acc = acc + x[3] * c[3];
                r = acc;
            }
            out3[(size_t)i * 3 + j] = r;
        }
}

/* src/vslam.cpp:129-161. Returns the number of map points that claimed a frame keypoint. */
int vbo_search_by_projection(const float *X, int n, const float *c2, int W, int H, const float *pts,
                             const int32_t *pre_idx, int k, const uint8_t *desc, int bytes,
                             int32_t *map_point_ids, const int32_t *obs_off, const uint8_t *obs_desc,
                             float radius, uint32_t dist_thr, int32_t *assign, float *proj_xy,
                             uint8_t *in_view) {
    float *pr = (float *)malloc(sizeof(float) * 3 * (size_t)(n > 0 ? n : 1));
    int32_t *hits = (int32_t *)malloc(sizeof(int32_t) * (size_t)(k > 0 ? k : 1));
    vbo_project_points(X, n, c2, pr); /* :131 */
    int claimed = 0;
    for (int i = 0; i < n; i++) { /* :136-145: make non homogeneous, in-image test */
        const float h = pr[3 * i + 2];
        pr[3 * i + 0] = pr[3 * i + 0] / h;
        pr[3 * i + 1] = pr[3 * i + 1] / h;
        const float x = pr[3 * i + 0], y = pr[3 * i + 1];
        const int in = (x >= 0 && x < (float)W && y >= 0 && y < (float)H) ? 1 : 0;
        if (in_view) in_view[i] = (uint8_t)in;
        if (proj_xy) { proj_xy[2 * i] = x; proj_xy[2 * i + 1] = y; }
        assign[i] = -1;
    }
    for (int i = 0; i < n; i++) { /* :147-160 */
        const float x = pr[3 * i + 0], y = pr[3 * i + 1];
        if (!(x >= 0 && x < (float)W && y >= 0 && y < (float)H)) continue;
        const int nh = vbo_kdtree_radius(pts, pre_idx, k, x, y, radius, hits, k); /* :149, pre-order */
        for (int t = 0; t < nh; t++) {
            const int idx = hits[t];
            if (map_point_ids[idx] >= 0) continue; /* :151 */
            const uint32_t d = vbo_orb_distance(desc + (size_t)idx * bytes, obs_desc + (size_t)obs_off[i] * bytes,
                                                obs_off[i + 1] - obs_off[i], bytes); /* :152 */
            if (d < dist_thr) { /* :153 */
                map_point_ids[idx] = i;
                assign[i] = idx;
                claimed++;
                break;
            }
        }
    }
    free(pr);
    free(hits);
    return claimed;
}

/* ---- pose recovery and triangulation (src/helpers.cpp) ------------------------------------- */

/* E = K.t() * F * K (src/helpers.cpp:4): cv::gemm(K, F, GEMM_1_T) — products and sums in double, one rounding — then
 * a flag-free 3x3 product in the fp32 small-matrix order. Both pinned against cv2 (tests/golden). */
void vbo_essential(const float *F, const float *K, float *E) {
    float T[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double acc = (double)K[0 * 3 + i] * (double)F[0 * 3 + j];
            acc += (double)K[1 * 3 + i] * (double)F[1 * 3 + j];
            acc += (double)K[2 * 3 + i] * (double)F[2 * 3 + j];
            T[i * 3 + j] = (float)acc;
        }
    mat3_mul_f32(T, K, E);
}

static float det3_f32(const float *m) { /* only its sign is used (:20,25) */
    double d = (double)m[0] * ((double)m[4] * m[8] - (double)m[5] * m[7]) -
               (double)m[1] * ((double)m[3] * m[8] - (double)m[5] * m[6]) +
               (double)m[2] * ((double)m[3] * m[7] - (double)m[4] * m[6]);
    return (float)d;
}

/* src/helpers.cpp:3-35. cv::SVD::compute is replaced by vbo_svd3x3 (defined sequence; PARITY UNPINNED there). */
void vbo_extract_rt(const float *F, const float *K, float *R, float *t) {
    float E[9], U[9], D[3], Vt[9];
    vbo_essential(F, K, E);
    vbo_svd3x3(E, U, D, Vt);                                   /* :7 */
    /* An exactly rank-2 E (e.g. F = [t]x with K = I) has a singular value of exactly 0, for which vbo_svd3x3 returns a zero
     * column of U (it cannot divide by 0). cv::SVD::compute with FULL_UV completes the basis instead, so U.col(2) is the
     * unit null vector of E^T (up to sign, which :31 fixes). Same here: the normalised cross product of the first two
     * columns, in double, when the third column is not a unit vector or its singular value is below 2^-40 of the largest
     * (the Jacobi sweep stops at 2^-52 relative, so such a column is rounding noise, not a direction). */
    {
        const double n2 = ((double)U[2] * U[2] + (double)U[5] * U[5]) + (double)U[8] * U[8];
        if (!(n2 >= 0.5) || !(D[2] > D[0] * 9.094947017729282e-13f)) {
            const double c0 = (double)U[3] * U[7] - (double)U[6] * U[4];
            const double c1 = (double)U[6] * U[1] - (double)U[0] * U[7];
            const double c2 = (double)U[0] * U[4] - (double)U[3] * U[1];
            const double ci = 1.0 / sqrt((c0 * c0 + c1 * c1) + c2 * c2);
            U[2] = (float)(c0 * ci); U[5] = (float)(c1 * ci); U[8] = (float)(c2 * ci);
        }
    }
    for (int i = 0; i < 3; i++) t[i] = U[i * 3 + 2];           /* :9  U.col(2) */
    const double nrm = sqrt(((double)t[0] * t[0] + (double)t[1] * t[1]) + (double)t[2] * t[2]); /* cv::norm: double */
    const double inv = 1.0 / nrm;                              /* :11 Mat /= s scales by 1/s in double */
    for (int i = 0; i < 3; i++) t[i] = (float)((double)t[i] * inv);
    static const float Wm[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};   /* :13-16 */
    static const float Wt[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1};
    float T[9], R1[9], R2[9];
    mat3_mul_f32(U, Wm, T); mat3_mul_f32(T, Vt, R1);           /* :18 */
    if (det3_f32(R1) < 0) for (int i = 0; i < 9; i++) R1[i] = -R1[i];
    mat3_mul_f32(U, Wt, T); mat3_mul_f32(T, Vt, R2);           /* :23 */
    if (det3_f32(R2) < 0) for (int i = 0; i < 9; i++) R2[i] = -R2[i];
    const float tr = (R1[0] + R1[4]) + R1[8];                  /* :29 */
    for (int i = 0; i < 9; i++) R[i] = (tr < 0) ? R2[i] : R1[i];
    if (t[2] < 0) for (int i = 0; i < 3; i++) t[i] = -t[i];    /* :31-33 */
}

/* Right singular vector of the smallest singular value of a 4x4 (fp32 in, fp32 out): fp64 one-sided Jacobi over the
 * column pairs (0,1)(0,2)(0,3)(1,2)(1,3)(2,3), same rotation and stopping rule as vbo_svd3x3; ties keep the later
 * column. Replaces cv::SVD::compute(A, ..., MODIFY_A | FULL_UV) + V_t.row(3) (src/helpers.cpp:57,67). */
void vbo_null_vector_4x4(const float *A, float *v4) {
    double G[4][4], V[4][4];
    static const int PQ[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            G[i][j] = (double)A[i * 4 + j];
            V[i][j] = (i == j) ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < VBO_SVD3_MAX_SWEEPS; sweep++) {
        int rotated = 0;
        for (int r = 0; r < 6; r++) {
            int p = PQ[r][0], q = PQ[r][1];
            double alpha = 0.0, bet = 0.0, gamma = 0.0;
            for (int k = 0; k < 4; k++) {
                alpha += G[k][p] * G[k][p];
                bet += G[k][q] * G[k][q];
                gamma += G[k][p] * G[k][q];
            }
            if (gamma == 0.0 || fabs(gamma) <= VBO_SVD3_EPS * sqrt(alpha * bet)) continue;
            rotated = 1;
            double zeta = (bet - alpha) / (2.0 * gamma);
            double t = 1.0 / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            if (zeta < 0.0) t = -t;
            double c = 1.0 / sqrt(1.0 + t * t);
            double s = c * t;
            for (int k = 0; k < 4; k++) {
                double gp = G[k][p], gq = G[k][q];
                G[k][p] = c * gp - s * gq;
                G[k][q] = s * gp + c * gq;
                double vp = V[k][p], vq = V[k][q];
                V[k][p] = c * vp - s * vq;
                V[k][q] = s * vp + c * vq;
            }
        }
        if (!rotated) break;
    }
    int best = 0;
    double bestn = 0.0;
    for (int j = 0; j < 4; j++) {
        double n2 = 0.0;
        for (int k = 0; k < 4; k++) n2 += G[k][j] * G[k][j];
        if (j == 0 || n2 <= bestn) { best = j; bestn = n2; }
    }
    for (int k = 0; k < 4; k++) v4[k] = (float)V[k][best];
}

/* src/helpers.cpp:37-80. p1, p2: [n][2]; c1, c2: 3x4 cameras; out: [n][4] = (X/w, Y/w, Z/w, 1). */
void vbo_triangulate(const float *p1, const float *p2, int n, const float *c1, const float *c2, float *out) {
    for (int i = 0; i < n; i++) {
        float A[16], v[4];
        for (int j = 0; j < 4; j++) { /* :49-52, one fp32 rounding per multiply and per subtract */
            A[0 * 4 + j] = p1[2 * i] * c1[2 * 4 + j] - c1[0 * 4 + j];
            A[1 * 4 + j] = p1[2 * i + 1] * c1[2 * 4 + j] - c1[1 * 4 + j];
            A[2 * 4 + j] = p2[2 * i] * c2[2 * 4 + j] - c2[0 * 4 + j];
            A[3 * 4 + j] = p2[2 * i + 1] * c2[2 * 4 + j] - c2[1 * 4 + j];
        }
        vbo_null_vector_4x4(A, v); /* :57,67 */
        out[4 * i + 0] = v[0] / v[3]; /* :71-74 */
        out[4 * i + 1] = v[1] / v[3];
        out[4 * i + 2] = v[2] / v[3];
        out[4 * i + 3] = 1.0f;
    }
}

/* Reprojection gate after triangulation, src/vslam.cpp:186-251 as written:
 *   reproj_k = points_4d * c_k.t()            :192-193  (cv::gemm GEMM_2_T on an n x 4 by 3 x 4: vbo_project_points' rule)
 *   for (i = 0; i < reproj.rows; i += 3)      :201-211  the loop counts ROWS but indexes the FLAT data, so it divides
 *       data[i] /= h, data[i+1] /= h, h = 1             x, y by h only for the first ceil(n / 3) points; the rest keep
 *                                                       their raw homogeneous x, y. Reproduced, not repaired.
 *   d_k = reproj_k.colRange(0, 2) - initial_points_k    :231-232  fp32 subtraction
 *   per row i:                                          :237-251
 *       if (map_point_ids[i] > 0) continue              (strictly positive: id 0 is NOT skipped; indexed by ROW)
 *       re1 = d1.row(i).dot(d1.row(i))                  cv::Mat::dot on two floats: products and the sum in double
 *       if (re1 > thresholdSq) continue                 (dotProd_32f's scalar tail), result narrowed to f32;
 *       re2 likewise; if (re2 > thresholdSq) continue   NaN compares false, so a NaN error PASSES the gate
 *       reprojection_inliers.push_back(i); reproj_error += re1 + re2   (f32 add, accumulated in f64, row order)
 * The Mat::dot rule is restated from OpenCV's source (there is no Python binding to pin it with); everything else is
 * pinned by cv2 goldens (tests/golden/gen_golden_gate.py). re1 / re2 are reported for every row. Returns the inlier count. */
int vbo_reprojection_gate(const float *points4, int n, const float *c1, const float *c2, const float *ip1, const float *ip2,
                          const int32_t *map_point_ids, float threshold_sq, float *re1_out, float *re2_out,
                          int32_t *inlier_idx, double *reproj_error) {
    float *r1 = (float *)malloc(sizeof(float) * 3 * (size_t)(n > 0 ? n : 1));
    float *r2 = (float *)malloc(sizeof(float) * 3 * (size_t)(n > 0 ? n : 1));
    vbo_project_points(points4, n, c1, r1);
    vbo_project_points(points4, n, c2, r2);
    for (size_t i = 0; i < (size_t)n; i += 3) {
        const float h1 = r1[i + 2];
        r1[i] = r1[i] / h1; r1[i + 1] = r1[i + 1] / h1; r1[i + 2] = 1.0f;
        const float h2 = r2[i + 2];
        r2[i] = r2[i] / h2; r2[i + 1] = r2[i + 1] / h2; r2[i + 2] = 1.0f;
    }
    int cnt = 0;
    double err = 0.0;
    for (int i = 0; i < n; i++) {
        const float d1x = r1[3 * i] - ip1[2 * i], d1y = r1[3 * i + 1] - ip1[2 * i + 1];
        const float d2x = r2[3 * i] - ip2[2 * i], d2y = r2[3 * i + 1] - ip2[2 * i + 1];
        const float re1 = (float)((double)d1x * (double)d1x + (double)d1y * (double)d1y);
        const float re2 = (float)((double)d2x * (double)d2x + (double)d2y * (double)d2y);
        if (re1_out) re1_out[i] = re1;
        if (re2_out) re2_out[i] = re2;
        if (map_point_ids && map_point_ids[i] > 0) continue;
        if (re1 > threshold_sq) continue;
        if (re2 > threshold_sq) continue;
        if (inlier_idx) inlier_idx[cnt] = i;
        cnt++;
        err += (double)(re1 + re2);
    }
    if (reproj_error) *reproj_error = err;
    free(r1);
    free(r2);
    return cnt;
}

/* ============================ seed hook ======================================================= */

static unsigned g_ref_seed = 0;
void vbo_ref_seed_set(unsigned seed) { g_ref_seed = seed; }
unsigned vbo_ref_seed_next(void) { return g_ref_seed; }

/* ============================ bench helper ==================================================== */

long vbo_pairs_run(const float *pts, const uint8_t *desc, int nframes, int k, int bytes, double ratio,
                   int max_iterations, float threshold, uint32_t seed0, int n_threads, int *threads_used) {
    long total = 0;
    int used = 1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel
    {
#pragma omp single
        used = omp_get_num_threads();
    }
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
#endif
    for (int i = 0; i < nframes - 1; i++) {
        int32_t *out = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)k);
        float F[9];
        int nt, bh;
        int r = vbo_match_features(pts + (size_t)i * k * 2, desc + (size_t)i * k * bytes, k,
                                   pts + (size_t)(i + 1) * k * 2, desc + (size_t)(i + 1) * k * bytes, k, bytes,
                                   ratio, 8, max_iterations, threshold, seed0 + (uint32_t)i, out, F, &nt, &bh);
        total += (r > 0 ? r : 0);
        free(out);
    }
    if (threads_used) *threads_used = used;
    return total;
}
