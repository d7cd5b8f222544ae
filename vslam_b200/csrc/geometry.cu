// geometry.cu — the two consumers of F downstream of the correspondence path:
//   extract_Rt   reference src/helpers.cpp:3-35    E = K^T F K, 3x3 SVD, rotation pick, unit translation
//   triangulate  reference src/helpers.cpp:37-80   per match a 4x4 DLT system, its null vector, dehomogenise
// One thread per problem / per match; every operation is an explicit round-to-nearest intrinsic in the order the
// CPU checker defines. E is bit-identical to what cv::gemm produces (fp64-accumulated K^T F, then the fp32
// small-matrix product); the SVDs replace cv::SVD::compute by the same fp64 one-sided Jacobi the 8-point solve uses
// (solve8.cuh), so R, t and the points are defined by this build and agree with an OpenCV build to its SVD's accuracy.
#include "common.cuh"
#include "solve8.cuh"

namespace vb {

struct Mat9 { float m[9]; };
struct Cam12 { float c[12]; };

__device__ __forceinline__ double det3_d(const float (&m)[9]) {
    const double a = __dsub_rn(__dmul_rn((double)m[4], (double)m[8]), __dmul_rn((double)m[5], (double)m[7]));
    const double b = __dsub_rn(__dmul_rn((double)m[3], (double)m[8]), __dmul_rn((double)m[5], (double)m[6]));
    const double c = __dsub_rn(__dmul_rn((double)m[3], (double)m[7]), __dmul_rn((double)m[4], (double)m[6]));
    return __dadd_rn(__dsub_rn(__dmul_rn((double)m[0], a), __dmul_rn((double)m[1], b)), __dmul_rn((double)m[2], c));
}

__global__ void __launch_bounds__(64) k_extract_rt(const float *__restrict__ F_all, uint32_t P, Mat9 Km, float *__restrict__ R_all,
                                                   float *__restrict__ t_all, float *__restrict__ E_all) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float F[9], T[9], E[9], U[9], D[3], Vt[9];
#pragma unroll
    for (int i = 0; i < 9; i++) F[i] = F_all[(size_t)p * 9 + i];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {   // K^T * F, cv::gemm GEMM_1_T: products (exact in fp64) and sums in double, one rounding
            double acc = __dmul_rn((double)Km.m[0 * 3 + i], (double)F[0 * 3 + j]);
            acc = __dadd_rn(acc, __dmul_rn((double)Km.m[1 * 3 + i], (double)F[1 * 3 + j]));
            acc = __dadd_rn(acc, __dmul_rn((double)Km.m[2 * 3 + i], (double)F[2 * 3 + j]));
            T[i * 3 + j] = __double2float_rn(acc);
        }
    mat3_mul_f32(T, Km.m, E);
    if (E_all)
#pragma unroll
        for (int i = 0; i < 9; i++) E_all[(size_t)p * 9 + i] = E[i];
    svd3x3(E, U, D, Vt);                                            // :7
    {   // exactly singular E: svd3x3 leaves a zero (sv == 0) or noise (sv < 2^-40 sv0) third column of U; complete the basis
        // as cv::SVD's FULL_UV does
        const double n2c = __dadd_rn(__dadd_rn(__dmul_rn((double)U[2], (double)U[2]), __dmul_rn((double)U[5], (double)U[5])),
                                     __dmul_rn((double)U[8], (double)U[8]));
        if (!(n2c >= 0.5) || !(D[2] > __fmul_rn(D[0], 9.094947017729282e-13f))) {
            const double c0 = __dsub_rn(__dmul_rn((double)U[3], (double)U[7]), __dmul_rn((double)U[6], (double)U[4]));
            const double c1 = __dsub_rn(__dmul_rn((double)U[6], (double)U[1]), __dmul_rn((double)U[0], (double)U[7]));
            const double c2 = __dsub_rn(__dmul_rn((double)U[0], (double)U[4]), __dmul_rn((double)U[3], (double)U[1]));
            const double ci = __ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(c0, c0), __dmul_rn(c1, c1)), __dmul_rn(c2, c2))));
            U[2] = __double2float_rn(__dmul_rn(c0, ci));
            U[5] = __double2float_rn(__dmul_rn(c1, ci));
            U[8] = __double2float_rn(__dmul_rn(c2, ci));
        }
    }
    float t[3] = {U[2], U[5], U[8]};                                // :9
    const double n2 = __dadd_rn(__dadd_rn(__dmul_rn((double)t[0], (double)t[0]), __dmul_rn((double)t[1], (double)t[1])),
                                __dmul_rn((double)t[2], (double)t[2]));
    const double inv = __ddiv_rn(1.0, __dsqrt_rn(n2));              // :11
#pragma unroll
    for (int i = 0; i < 3; i++) t[i] = __double2float_rn(__dmul_rn((double)t[i], inv));
    const float Wm[9] = {0.f, -1.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 1.f};   // :13-16
    const float Wt[9] = {0.f, 1.f, 0.f, -1.f, 0.f, 0.f, 0.f, 0.f, 1.f};
    float X[9], R1[9], R2[9];
    mat3_mul_f32(U, Wm, X);
    mat3_mul_f32(X, Vt, R1);                                        // :18
    if (__double2float_rn(det3_d(R1)) < 0.f)
#pragma unroll
        for (int i = 0; i < 9; i++) R1[i] = -R1[i];
    mat3_mul_f32(U, Wt, X);
    mat3_mul_f32(X, Vt, R2);                                        // :23
    if (__double2float_rn(det3_d(R2)) < 0.f)
#pragma unroll
        for (int i = 0; i < 9; i++) R2[i] = -R2[i];
    const float tr = __fadd_rn(__fadd_rn(R1[0], R1[4]), R1[8]);     // :29
#pragma unroll
    for (int i = 0; i < 9; i++) R_all[(size_t)p * 9 + i] = (tr < 0.f) ? R2[i] : R1[i];
    const bool flip = t[2] < 0.f;                                   // :31
#pragma unroll
    for (int i = 0; i < 3; i++) t_all[(size_t)p * 3 + i] = flip ? -t[i] : t[i];
}

// Right singular vector of the smallest singular value of a 4x4: fp64 one-sided Jacobi, pairs (0,1)(0,2)(0,3)(1,2)(1,3)(2,3).
__device__ __forceinline__ void null_vector_4x4(const float (&A)[16], float (&v)[4]) {
    double G[4][4], V[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            G[i][j] = (double)A[i * 4 + j];
            V[i][j] = (i == j) ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < VB_SVD3_MAX_SWEEPS; sweep++) {
        bool rotated = false;
#pragma unroll
        for (int r = 0; r < 6; r++) {
            const int p = (r < 3) ? 0 : (r < 5) ? 1 : 2;
            const int q = (r == 0) ? 1 : (r == 1 || r == 3) ? 2 : 3;
            double alpha = __dmul_rn(G[0][p], G[0][p]), bet = __dmul_rn(G[0][q], G[0][q]), gamma = __dmul_rn(G[0][p], G[0][q]);
#pragma unroll
            for (int k = 1; k < 4; k++) {
                alpha = __dadd_rn(alpha, __dmul_rn(G[k][p], G[k][p]));
                bet = __dadd_rn(bet, __dmul_rn(G[k][q], G[k][q]));
                gamma = __dadd_rn(gamma, __dmul_rn(G[k][p], G[k][q]));
            }
            if (gamma == 0.0 || fabs(gamma) <= __dmul_rn(VB_SVD3_EPS, __dsqrt_rn(__dmul_rn(alpha, bet)))) continue;
            rotated = true;
            const double zeta = __ddiv_rn(__dsub_rn(bet, alpha), __dmul_rn(2.0, gamma));
            double t = __ddiv_rn(1.0, __dadd_rn(fabs(zeta), __dsqrt_rn(__dadd_rn(1.0, __dmul_rn(zeta, zeta)))));
            if (zeta < 0.0) t = -t;
            const double c = __ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(1.0, __dmul_rn(t, t))));
            const double s = __dmul_rn(c, t);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const double gp = G[k][p], gq = G[k][q];
                G[k][p] = __dsub_rn(__dmul_rn(c, gp), __dmul_rn(s, gq));
                G[k][q] = __dadd_rn(__dmul_rn(s, gp), __dmul_rn(c, gq));
                const double vp = V[k][p], vq = V[k][q];
                V[k][p] = __dsub_rn(__dmul_rn(c, vp), __dmul_rn(s, vq));
                V[k][q] = __dadd_rn(__dmul_rn(s, vp), __dmul_rn(c, vq));
            }
        }
        if (!rotated) break;
    }
    double bestn = 0.0;
    double bv[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        double n2 = __dmul_rn(G[0][j], G[0][j]);
#pragma unroll
        for (int k = 1; k < 4; k++) n2 = __dadd_rn(n2, __dmul_rn(G[k][j], G[k][j]));
        if (j == 0 || n2 <= bestn) {
            bestn = n2;
#pragma unroll
            for (int k = 0; k < 4; k++) bv[k] = V[k][j];
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = __double2float_rn(bv[k]);
}

__global__ void __launch_bounds__(128) k_triangulate(const float2 *__restrict__ p1, const float2 *__restrict__ p2, uint32_t n,
                                                     Cam12 c1, Cam12 c2, float4 *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 a = p1[i], b = p2[i];
    float A[16], v[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {   // :49-52
        A[0 * 4 + j] = __fsub_rn(__fmul_rn(a.x, c1.c[8 + j]), c1.c[j]);
        A[1 * 4 + j] = __fsub_rn(__fmul_rn(a.y, c1.c[8 + j]), c1.c[4 + j]);
        A[2 * 4 + j] = __fsub_rn(__fmul_rn(b.x, c2.c[8 + j]), c2.c[j]);
        A[3 * 4 + j] = __fsub_rn(__fmul_rn(b.y, c2.c[8 + j]), c2.c[4 + j]);
    }
    null_vector_4x4(A, v);            // :57, :67
    out[i] = make_float4(__fdiv_rn(v[0], v[3]), __fdiv_rn(v[1], v[3]), __fdiv_rn(v[2], v[3]), 1.0f);   // :71-74
}

// One row of `points_4d * c.t()` (src/vslam.cpp:192-193) in cv::gemm's arithmetic for that shape: fewer than 100 rows ->
// products and sums in double, one rounding; otherwise the fp32 chain ((x0 c0 + x1 c1) + x2 c2) + x3 c3 (see k_sbp_project).
__device__ __forceinline__ void reproject_row(const float4 x, const Cam12 &cam, int small, float (&r)[3]) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const float *c = cam.c + 4 * j;
        if (small) {
            double acc = __dmul_rn((double)x.x, (double)c[0]);
            acc = __dadd_rn(acc, __dmul_rn((double)x.y, (double)c[1]));
            acc = __dadd_rn(acc, __dmul_rn((double)x.z, (double)c[2]));
            acc = __dadd_rn(acc, __dmul_rn((double)x.w, (double)c[3]));
            r[j] = __double2float_rn(acc);
        } else {
            float acc = __fmul_rn(x.x, c[0]);
            acc = __fadd_rn(acc, __fmul_rn(x.y, c[1]));
            acc = __fadd_rn(acc, __fmul_rn(x.z, c[2]));
            acc = __fadd_rn(acc, __fmul_rn(x.w, c[3]));
            r[j] = acc;
        }
    }
}

// triangulate (src/helpers.cpp:37-80) fused with the reprojection gate that consumes it (src/vslam.cpp:186-251): the point
// never leaves the thread's registers between the two. The reference's dehomogenisation loop (:201-211) counts rows but
// indexes the flat data, so only rows r with 3 r < n are divided by h — reproduced as written. flag[i] = passes the gate.
__global__ void __launch_bounds__(128) k_triangulate_gate(const float2 *__restrict__ p1, const float2 *__restrict__ p2, uint32_t n,
                                                          Cam12 c1, Cam12 c2, const int32_t *__restrict__ ids, float thr_sq,
                                                          int small, float4 *__restrict__ out, float *__restrict__ re1_out,
                                                          float *__restrict__ re2_out, uint8_t *__restrict__ flag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 a = p1[i], b = p2[i];
    float A[16], v[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        A[0 * 4 + j] = __fsub_rn(__fmul_rn(a.x, c1.c[8 + j]), c1.c[j]);
        A[1 * 4 + j] = __fsub_rn(__fmul_rn(a.y, c1.c[8 + j]), c1.c[4 + j]);
        A[2 * 4 + j] = __fsub_rn(__fmul_rn(b.x, c2.c[8 + j]), c2.c[j]);
        A[3 * 4 + j] = __fsub_rn(__fmul_rn(b.y, c2.c[8 + j]), c2.c[4 + j]);
    }
    null_vector_4x4(A, v);
    const float4 X = make_float4(__fdiv_rn(v[0], v[3]), __fdiv_rn(v[1], v[3]), __fdiv_rn(v[2], v[3]), 1.0f);
    out[i] = X;
    float r1[3], r2[3];
    reproject_row(X, c1, small, r1);
    reproject_row(X, c2, small, r2);
    if (3ull * i < (unsigned long long)n) {   // :201-211 as written
        r1[0] = __fdiv_rn(r1[0], r1[2]); r1[1] = __fdiv_rn(r1[1], r1[2]);
        r2[0] = __fdiv_rn(r2[0], r2[2]); r2[1] = __fdiv_rn(r2[1], r2[2]);
    }
    const float d1x = __fsub_rn(r1[0], a.x), d1y = __fsub_rn(r1[1], a.y);   // :231
    const float d2x = __fsub_rn(r2[0], b.x), d2y = __fsub_rn(r2[1], b.y);   // :232
    // cv::Mat::dot of two floats: exact products and their sum in double, narrowed once (:240, :242)
    const float re1 = __double2float_rn(__dadd_rn(__dmul_rn((double)d1x, (double)d1x), __dmul_rn((double)d1y, (double)d1y)));
    const float re2 = __double2float_rn(__dadd_rn(__dmul_rn((double)d2x, (double)d2x), __dmul_rn((double)d2y, (double)d2y)));
    if (re1_out) re1_out[i] = re1;
    if (re2_out) re2_out[i] = re2;
    bool pass = !(ids && ids[i] > 0);      // :239 (strictly positive, indexed by row)
    pass = pass && !(re1 > thr_sq) && !(re2 > thr_sq);   // :241, :243 — a NaN error passes, as in the reference
    flag[i] = pass ? 1 : 0;
}

// reprojection_inliers.push_back(i) in row order (:245) and reproj_error += re1 + re2 (:249, f32 add accumulated in f64
// in row order — one thread walks the list at the end so the sum has the reference's order).
__global__ void __launch_bounds__(256) k_gate_compact(const uint8_t *__restrict__ flag, const float *__restrict__ re1,
                                                      const float *__restrict__ re2, uint32_t n, uint32_t *__restrict__ idx,
                                                      uint32_t *__restrict__ count, double *__restrict__ err) {
    __shared__ int s_scan[8];
    __shared__ int s_base;
    const uint32_t tid = threadIdx.x;
    const int lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < n; i0 += blockDim.x) {
        const uint32_t i = i0 + tid;
        const int in = (i < n && flag[i]) ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        const int wpre = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) s_scan[w] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int j = 0; j < 8; j++) {
            if (j < w) woff += s_scan[j];
            tot += s_scan[j];
        }
        const int base = s_base;
        if (in) idx[base + woff + wpre] = i;
        __syncthreads();
        if (tid == 0) s_base = base + tot;
        __syncthreads();
    }
    if (tid == 0) {
        const uint32_t c = (uint32_t)s_base;
        *count = c;
        double e = 0.0;
        for (uint32_t j = 0; j < c; j++) {
            const uint32_t i = idx[j];
            e = __dadd_rn(e, (double)__fadd_rn(re1[i], re2[i]));
        }
        *err = e;
    }
}

}  // namespace vb

using namespace vb;

extern "C" {

int vb_triangulate_gated(vb_ctx *ctx, const float *p1, const float *p2, uint32_t n, const float *c1, const float *c2,
                         const int32_t *map_point_ids, float threshold_sq, float *points4, float *re1, float *re2,
                         uint32_t *inlier_idx, uint32_t *n_inliers, double *reproj_error) {
    VB_REQUIRE(ctx && c1 && c2 && n_inliers && (n == 0 || (p1 && p2)), VB_ERR_INVALID, "NULL argument");
    *n_inliers = 0;
    if (reproj_error) *reproj_error = 0.0;
    if (n == 0) return VB_OK;
    VB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    // layout: p1 | p2 | points4 | re1 | re2 | idx | ids | err (8 B aligned first) | count | flag
    const size_t bytes = (size_t)n * (8 + 8 + 16 + 4 + 4 + 4 + 4) + 16 + (size_t)n;
    if ((rc = ctx->ws_ensure(WS_SBP_X, bytes))) return rc;
    uint8_t *base = ctx->ws[WS_SBP_X].as<uint8_t>();
    double *err_d = reinterpret_cast<double *>(base);
    uint32_t *cnt_d = reinterpret_cast<uint32_t *>(base + 8);
    float4 *out_d = reinterpret_cast<float4 *>(base + 16);
    float2 *p1_d = reinterpret_cast<float2 *>(out_d + n), *p2_d = p1_d + n;
    float *re1_d = reinterpret_cast<float *>(p2_d + n), *re2_d = re1_d + n;
    uint32_t *idx_d = reinterpret_cast<uint32_t *>(re2_d + n);
    int32_t *ids_d = reinterpret_cast<int32_t *>(idx_d + n);
    uint8_t *flag_d = reinterpret_cast<uint8_t *>(ids_d + n);
    VB_CUDA(cudaMemcpyAsync(p1_d, p1, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    VB_CUDA(cudaMemcpyAsync(p2_d, p2, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (map_point_ids) VB_CUDA(cudaMemcpyAsync(ids_d, map_point_ids, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    Cam12 a, b;
    memcpy(a.c, c1, sizeof(a.c));
    memcpy(b.c, c2, sizeof(b.c));
    ctx->prof_begin("triangulate");
    k_triangulate_gate<<<div_up(n, 128), 128, 0, ctx->stream>>>(p1_d, p2_d, n, a, b, map_point_ids ? ids_d : nullptr, threshold_sq,
                                                               n < 100 ? 1 : 0, out_d, re1_d, re2_d, flag_d);
    k_gate_compact<<<1, 256, 0, ctx->stream>>>(flag_d, re1_d, re2_d, n, idx_d, cnt_d, err_d);
    ctx->prof_end("triangulate");
    ctx->launches += 2;
    VB_CUDA(cudaGetLastError());
    struct { double err; uint32_t cnt; uint32_t pad; } head;
    VB_CUDA(cudaMemcpyAsync(&head, base, 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (points4) VB_CUDA(cudaMemcpyAsync(points4, out_d, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (re1) VB_CUDA(cudaMemcpyAsync(re1, re1_d, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (re2) VB_CUDA(cudaMemcpyAsync(re2, re2_d, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    VB_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_inliers = head.cnt;
    if (reproj_error) *reproj_error = head.err;
    if (inlier_idx && head.cnt) VB_CUDA(cudaMemcpy(inlier_idx, idx_d, (size_t)head.cnt * 4, cudaMemcpyDeviceToHost));
    return VB_OK;
}

int vb_extract_rt(vb_ctx *ctx, const float *F, uint32_t P, const float *K, float *R, float *t, float *E_out) {
    VB_REQUIRE(ctx && K && (P == 0 || (F && R && t)), VB_ERR_INVALID, "NULL argument");
    if (P == 0) return VB_OK;
    VB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ctx->ws_ensure(WS_SBP_X, (size_t)P * (9 + 9 + 3 + 9) * 4))) return rc;
    float *F_d = ctx->ws[WS_SBP_X].as<float>(), *R_d = F_d + (size_t)P * 9, *t_d = R_d + (size_t)P * 9, *E_d = t_d + (size_t)P * 3;
    VB_CUDA(cudaMemcpyAsync(F_d, F, (size_t)P * 36, cudaMemcpyHostToDevice, ctx->stream));
    Mat9 Km;
    memcpy(Km.m, K, sizeof(Km.m));
    ctx->prof_begin("extract_rt");
    k_extract_rt<<<div_up(P, 64), 64, 0, ctx->stream>>>(F_d, P, Km, R_d, t_d, E_d);
    ctx->prof_end("extract_rt");
    ctx->launches++;
    VB_CUDA(cudaGetLastError());
    VB_CUDA(cudaMemcpyAsync(R, R_d, (size_t)P * 36, cudaMemcpyDeviceToHost, ctx->stream));
    VB_CUDA(cudaMemcpyAsync(t, t_d, (size_t)P * 12, cudaMemcpyDeviceToHost, ctx->stream));
    if (E_out) VB_CUDA(cudaMemcpyAsync(E_out, E_d, (size_t)P * 36, cudaMemcpyDeviceToHost, ctx->stream));
    VB_CUDA(cudaStreamSynchronize(ctx->stream));
    return VB_OK;
}

int vb_triangulate(vb_ctx *ctx, const float *p1, const float *p2, uint32_t n, const float *c1, const float *c2, float *points4) {
    VB_REQUIRE(ctx && c1 && c2 && (n == 0 || (p1 && p2 && points4)), VB_ERR_INVALID, "NULL argument");
    if (n == 0) return VB_OK;
    VB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ctx->ws_ensure(WS_SBP_X, (size_t)n * (8 + 8 + 16)))) return rc;
    float2 *p1_d = ctx->ws[WS_SBP_X].as<float2>(), *p2_d = p1_d + n;
    float4 *out_d = reinterpret_cast<float4 *>(p2_d + n);
    VB_CUDA(cudaMemcpyAsync(p1_d, p1, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    VB_CUDA(cudaMemcpyAsync(p2_d, p2, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    Cam12 a, b;
    memcpy(a.c, c1, sizeof(a.c));
    memcpy(b.c, c2, sizeof(b.c));
    ctx->prof_begin("triangulate");
    k_triangulate<<<div_up(n, 128), 128, 0, ctx->stream>>>(p1_d, p2_d, n, a, b, out_d);
    ctx->prof_end("triangulate");
    ctx->launches++;
    VB_CUDA(cudaGetLastError());
    VB_CUDA(cudaMemcpyAsync(points4, out_d, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    VB_CUDA(cudaStreamSynchronize(ctx->stream));
    return VB_OK;
}

}  // extern "C"
