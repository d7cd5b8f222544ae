// l2f.cu — float-descriptor matching (BASELINE config 3: 128-d fp32 descriptors). The reference has no
// float path (src/Frame.cpp:83 is NORM_HAMMING only); the semantics are the matcher's, carried over:
// two nearest train descriptors per query by L2 distance, ties to the lower train index, Lowe ratio in
// double on the float distances.
//
// Distance definition (build-defined, restated by the CPU checker): d2 = sum_k (a_k - b_k)^2 accumulated sequentially in fp32,
// one rounding per subtract / multiply / add; distance = sqrtf(d2).
//
// k_l2_partial<DIM>  exact kernel: one query per thread (descriptor in registers), train tiles staged
//                    in shared memory and read as broadcast LDS.128; contiguous train splits merge by
//                    (distance bits, index) like the Hamming path.
#include "common.cuh"
#include "ransac_dev.cuh"

namespace vb {

// tensor-core path (l2f_tc.cu): same results, large problems
bool l2_tc_eligible(const vb_ctx *ctx, uint32_t n1, uint32_t n2, uint32_t dim);
int l2_tc_launch(vb_ctx *ctx, const float *d1, uint32_t n1, const float *d2, uint32_t n2, uint32_t dim, ulonglong2 *out);

constexpr int L2_THREADS = 128;
constexpr int L2_TILE = 32;

template <int DIM>
__global__ void __launch_bounds__(L2_THREADS) k_l2_partial(const float *__restrict__ d1, const float *__restrict__ d2,
                                                           uint32_t n1, uint32_t n2, uint32_t split_len, uint32_t nsplits,
                                                           ulonglong2 *__restrict__ part) {
    __shared__ __align__(16) float tile[L2_TILE * DIM];
    const uint32_t s = blockIdx.y, tid = threadIdx.x;
    const uint32_t q = blockIdx.x * L2_THREADS + tid;
    const uint32_t t0 = s * split_len, t1 = min(t0 + split_len, n2);
    float a[DIM];
    {
        const uint32_t qs = q < n1 ? q : 0;
        const float4 *src = reinterpret_cast<const float4 *>(d1 + (size_t)qs * DIM);
#pragma unroll
        for (int i = 0; i < DIM / 4; i++) {
            const float4 v = __ldg(src + i);
            a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
        }
    }
    float bd1 = INFINITY, bd2 = INFINITY;
    uint32_t bj1 = 0xffffffffu, bj2 = 0xffffffffu;
    for (uint32_t ts = t0; ts < t1; ts += L2_TILE) {
        const uint32_t n_here = min((uint32_t)L2_TILE, t1 - ts);
        __syncthreads();
        {
            const float4 *src = reinterpret_cast<const float4 *>(d2 + (size_t)ts * DIM);
            float4 *dst = reinterpret_cast<float4 *>(tile);
            for (uint32_t i = tid; i < n_here * (DIM / 4); i += L2_THREADS) dst[i] = __ldg(src + i);
        }
        __syncthreads();
        for (uint32_t j = 0; j < n_here; j++) {
            const float4 *b = reinterpret_cast<const float4 *>(tile + j * DIM);
            float acc = 0.0f;
#pragma unroll
            for (int i = 0; i < DIM / 4; i++) {
                const float4 v = b[i];
                float df;
                df = __fsub_rn(a[4 * i], v.x);     acc = __fadd_rn(acc, __fmul_rn(df, df));
                df = __fsub_rn(a[4 * i + 1], v.y); acc = __fadd_rn(acc, __fmul_rn(df, df));
                df = __fsub_rn(a[4 * i + 2], v.z); acc = __fadd_rn(acc, __fmul_rn(df, df));
                df = __fsub_rn(a[4 * i + 3], v.w); acc = __fadd_rn(acc, __fmul_rn(df, df));
            }
            if (acc < bd2) {
                const uint32_t jg = ts + j;
                if (acc < bd1) { bd2 = bd1; bj2 = bj1; bd1 = acc; bj1 = jg; }
                else { bd2 = acc; bj2 = jg; }
            }
        }
    }
    if (q < n1) {
        ulonglong2 o;
        o.x = ((unsigned long long)__float_as_uint(bd1) << 32) | bj1;   // d2 >= 0: float bits are monotone
        o.y = ((unsigned long long)__float_as_uint(bd2) << 32) | bj2;
        part[(size_t)s * n1 + q] = o;
    }
}

__global__ void __launch_bounds__(1024) k_l2_finish(const ulonglong2 *__restrict__ part, uint32_t nsplits, uint32_t n1,
                                                   double ratio, int32_t *__restrict__ knn_idx, float *__restrict__ knn_dist,
                                                   int2 *__restrict__ tent, uint32_t *__restrict__ m_out,
                                                   const float2 *__restrict__ p1, const float2 *__restrict__ p2,
                                                   float4 *__restrict__ corr) {
    __shared__ int s_scan[32];
    __shared__ int s_base;
    const uint32_t tid = threadIdx.x;
    const int lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (uint32_t q0 = 0; q0 < n1; q0 += blockDim.x) {
        const uint32_t q = q0 + tid;
        int keep = 0;
        unsigned long long k1 = ~0ull, k2 = ~0ull;
        if (q < n1) {
            for (uint32_t s = 0; s < nsplits; s++) {
                const ulonglong2 v = part[(size_t)s * n1 + q];
                const unsigned long long ks[2] = {v.x, v.y};
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    const unsigned long long k = ks[t];
                    if (k < k2) {
                        if (k < k1) { k2 = k1; k1 = k; } else { k2 = k; }
                    }
                }
            }
            const float e0 = __fsqrt_rn(__uint_as_float((uint32_t)(k1 >> 32)));
            const float e1 = __fsqrt_rn(__uint_as_float((uint32_t)(k2 >> 32)));
            if (knn_idx) {
                knn_idx[2 * (size_t)q] = (int32_t)(uint32_t)k1;
                knn_idx[2 * (size_t)q + 1] = (int32_t)(uint32_t)k2;
                knn_dist[2 * (size_t)q] = e0;
                knn_dist[2 * (size_t)q + 1] = e1;
            }
            keep = ((double)e0 < (double)e1 * ratio) ? 1 : 0;
        }
        if (tent) {
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            const int wpre = __popc(bal & ((1u << lane) - 1u));
            if (lane == 0) s_scan[w] = __popc(bal);
            __syncthreads();
            int woff = 0, tot = 0;
            for (int j = 0; j < (int)(blockDim.x >> 5); j++) {
                if (j < w) woff += s_scan[j];
                tot += s_scan[j];
            }
            const int base = s_base;
            if (keep) {
                tent[base + woff + wpre] = make_int2((int)q, (int)(uint32_t)k1);
                if (corr) {   // whole-pair path: the correspondence the RANSAC stage reads
                    const float2 a = p1[q], b = p2[(uint32_t)k1];
                    corr[base + woff + wpre] = make_float4(a.x, a.y, b.x, b.y);
                }
            }
            __syncthreads();
            if (tid == 0) s_base = base + tot;
            __syncthreads();
        }
    }
    if (tid == 0 && m_out) *m_out = (uint32_t)s_base;
}

// Matcher on device-resident descriptors: leaves kNN-2 in WS_KNN, the ratio-test survivors in WS_TENT, their count in
// WS_M and — when p1_d / p2_d are given — the float4 correspondences in WS_CORR. Enqueues only.
static int l2_device(vb_ctx *ctx, const float *d1_d, uint32_t n1, const float *d2_d, uint32_t n2, uint32_t dim, double ratio,
                     const float2 *p1_d, const float2 *p2_d) {
    int rc;
    const bool use_tc = l2_tc_eligible(ctx, n1, n2, dim);
    const uint32_t qtiles = div_up(n1, L2_THREADS);
    uint32_t ns = use_tc ? 1 : div_up(4u * ctx->sm_count, qtiles);
    const uint32_t max_splits = div_up(n2, L2_TILE);
    if (ns > max_splits) ns = max_splits;
    if (ns > 64) ns = 64;
    if (ns < 1) ns = 1;
    const uint32_t split_len = div_up(div_up(n2, ns), L2_TILE) * L2_TILE;
    const uint32_t nsplits = div_up(n2, split_len);
    if ((rc = ctx->ws_ensure(WS_L2C, (size_t)nsplits * n1 * sizeof(ulonglong2)))) return rc;
    if ((rc = ctx->ws_ensure(WS_KNN, (size_t)n1 * 16))) return rc;
    if ((rc = ctx->ws_ensure(WS_TENT, (size_t)n1 * 8))) return rc;
    if ((rc = ctx->ws_ensure(WS_M, 16))) return rc;
    if (p1_d && (rc = ctx->ws_ensure(WS_CORR, (size_t)n1 * sizeof(float4)))) return rc;
    dim3 grid(qtiles, nsplits);
    if (use_tc) {
        if ((rc = l2_tc_launch(ctx, d1_d, n1, d2_d, n2, dim, ctx->ws[WS_L2C].as<ulonglong2>()))) return rc;
    } else {
        ctx->prof_begin("l2f");
        if (dim == 128)
            k_l2_partial<128><<<grid, L2_THREADS, 0, ctx->stream>>>(d1_d, d2_d, n1, n2, split_len, nsplits,
                                                                   ctx->ws[WS_L2C].as<ulonglong2>());
        else
            k_l2_partial<64><<<grid, L2_THREADS, 0, ctx->stream>>>(d1_d, d2_d, n1, n2, split_len, nsplits,
                                                                  ctx->ws[WS_L2C].as<ulonglong2>());
        ctx->prof_end("l2f");
        ctx->launches++;
    }
    int32_t *kidx = ctx->ws[WS_KNN].as<int32_t>();
    float *kdist = reinterpret_cast<float *>(kidx + (size_t)n1 * 2);
    ctx->prof_begin("finish");
    // one CTA (the compaction is ordered); 1 024 threads for config-3 sized inputs: 20 rounds instead of 79
    k_l2_finish<<<1, n1 > 2048 ? 1024 : 256, 0, ctx->stream>>>(ctx->ws[WS_L2C].as<ulonglong2>(), nsplits, n1, ratio, kidx, kdist,
                                            ctx->ws[WS_TENT].as<int2>(), ctx->ws[WS_M].as<uint32_t>(), p1_d, p2_d,
                                            p1_d ? ctx->ws[WS_CORR].as<float4>() : nullptr);
    ctx->prof_end("finish");
    ctx->launches++;
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

static int l2_host(vb_ctx *ctx, const float *d1, uint32_t n1, const float *d2, uint32_t n2, uint32_t dim, double ratio,
                   int32_t *idx, float *dist, int32_t *out_pairs, uint32_t *out_m) {
    VB_REQUIRE(ctx && d1 && d2, VB_ERR_INVALID, "NULL argument");
    VB_REQUIRE(dim == 64 || dim == 128, VB_ERR_INVALID, "descriptor dim must be 64 or 128");
    VB_REQUIRE(n2 >= 2, VB_ERR_TOO_FEW, "knnMatch(k=2) needs at least 2 train descriptors");
    if (n1 == 0) { if (out_m) *out_m = 0; return VB_OK; }
    VB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ctx->ws_ensure(WS_L2A, (size_t)n1 * dim * 4))) return rc;
    if ((rc = ctx->ws_ensure(WS_L2B, (size_t)n2 * dim * 4))) return rc;
    VB_CUDA(cudaMemcpyAsync(ctx->ws[WS_L2A].p, d1, (size_t)n1 * dim * 4, cudaMemcpyHostToDevice, ctx->stream));
    VB_CUDA(cudaMemcpyAsync(ctx->ws[WS_L2B].p, d2, (size_t)n2 * dim * 4, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = l2_device(ctx, ctx->ws[WS_L2A].as<float>(), n1, ctx->ws[WS_L2B].as<float>(), n2, dim, ratio, nullptr, nullptr)))
        return rc;
    int32_t *kidx = ctx->ws[WS_KNN].as<int32_t>();
    float *kdist = reinterpret_cast<float *>(kidx + (size_t)n1 * 2);
    uint32_t m = 0;
    VB_CUDA(cudaMemcpyAsync(&m, ctx->ws[WS_M].p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (idx) VB_CUDA(cudaMemcpyAsync(idx, kidx, (size_t)n1 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (dist) VB_CUDA(cudaMemcpyAsync(dist, kdist, (size_t)n1 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    VB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (out_pairs && m) VB_CUDA(cudaMemcpy(out_pairs, ctx->ws[WS_TENT].p, (size_t)m * 8, cudaMemcpyDeviceToHost));
    if (out_m) *out_m = m;
    return VB_OK;
}

// match_features with float descriptors (BASELINE config 3), device-resident inputs: matcher -> ratio test -> correspondences
// -> RansacFilter -> inlier copy-out, nothing returning to the host in between. results_d[0], out_matches_d[n1].
static int l2_pair_device(vb_ctx *ctx, const float2 *p1_d, const float *d1_d, uint32_t n1, const float2 *p2_d, const float *d2_d,
                          uint32_t n2, uint32_t dim, const vb_pair_params &prm, vb_pair_result *result_d, int2 *out_matches_d) {
    int rc;
    if ((rc = l2_device(ctx, d1_d, n1, d2_d, n2, dim, prm.ratio, p1_d, p2_d))) return rc;
    RansacPlan rp;
    if ((rc = ransac_plan(ctx, 1, n1, n1, prm.max_iterations, prm.min_items, &rp))) return rc;
    ProblemDims dims{ctx->ws[WS_M].as<uint32_t>(), 0, prm.seed0};
    return ransac_run(ctx, rp, ctx->ws[WS_CORR].as<float4>(), dims, prm.threshold, result_d, nullptr, ctx->ws[WS_TENT].as<int2>(),
                      out_matches_d, true);
}

}  // namespace vb

extern "C" {

int vb_knn2_l2f(vb_ctx *ctx, const float *d1, uint32_t n1, const float *d2, uint32_t n2, uint32_t dim, int32_t *idx,
                float *dist) {
    return vb::l2_host(ctx, d1, n1, d2, n2, dim, 0.7, idx, dist, nullptr, nullptr);
}

int vb_match_l2f(vb_ctx *ctx, const float *d1, uint32_t n1, const float *d2, uint32_t n2, uint32_t dim, double ratio,
                 int32_t *out_pairs, uint32_t *out_m) {
    VB_REQUIRE(out_pairs && out_m, VB_ERR_INVALID, "NULL output");
    return vb::l2_host(ctx, d1, n1, d2, n2, dim, ratio, nullptr, nullptr, out_pairs, out_m);
}

static int l2_pair_check(const vb_pair_params *p, uint32_t n1, uint32_t n2, uint32_t dim) {
    VB_REQUIRE(p != nullptr, VB_ERR_INVALID, "params is NULL");
    VB_REQUIRE(p->min_items >= 1 && p->min_items <= 8, VB_ERR_INVALID, "min_items must be in 1..8");
    VB_REQUIRE(p->max_iterations > 0, VB_ERR_INVALID, "max_iterations is 0");
    VB_REQUIRE(dim == 64 || dim == 128, VB_ERR_INVALID, "descriptor dim must be 64 or 128");
    VB_REQUIRE(n1 > 0, VB_ERR_TOO_FEW, "no query keypoints");
    VB_REQUIRE(n2 >= 2, VB_ERR_TOO_FEW, "knnMatch(k=2) needs at least 2 train descriptors");
    return VB_OK;
}

int vb_match_features_l2f_d(vb_ctx *ctx, const float *p1_d, const float *d1_d, uint32_t n1, const float *p2_d, const float *d2_d,
                            uint32_t n2, uint32_t dim, const vb_pair_params *params, int32_t *out_matches_d,
                            vb_pair_result *result_d) {
    VB_REQUIRE(ctx && p1_d && d1_d && p2_d && d2_d && result_d, VB_ERR_INVALID, "NULL argument");
    int rc;
    if ((rc = l2_pair_check(params, n1, n2, dim))) return rc;
    VB_CUDA(cudaSetDevice(ctx->device));
    return vb::l2_pair_device(ctx, reinterpret_cast<const float2 *>(p1_d), d1_d, n1, reinterpret_cast<const float2 *>(p2_d), d2_d,
                              n2, dim, *params, result_d, reinterpret_cast<int2 *>(out_matches_d));
}

int vb_match_features_l2f(vb_ctx *ctx, const float *p1, const float *d1, uint32_t n1, const float *p2, const float *d2,
                          uint32_t n2, uint32_t dim, const vb_pair_params *params, int32_t *out_matches, vb_pair_result *result) {
    VB_REQUIRE(ctx && p1 && d1 && p2 && d2 && result, VB_ERR_INVALID, "NULL argument");
    int rc;
    if ((rc = l2_pair_check(params, n1, n2, dim))) return rc;
    VB_CUDA(cudaSetDevice(ctx->device));
    if ((rc = ctx->ws_ensure(vb::WS_P1, (size_t)n1 * 8))) return rc;
    if ((rc = ctx->ws_ensure(vb::WS_P2, (size_t)n2 * 8))) return rc;
    if ((rc = ctx->ws_ensure(vb::WS_L2A, (size_t)n1 * dim * 4))) return rc;
    if ((rc = ctx->ws_ensure(vb::WS_L2B, (size_t)n2 * dim * 4))) return rc;
    if ((rc = ctx->ws_ensure(vb::WS_RESULT, sizeof(vb_pair_result)))) return rc;
    if ((rc = ctx->ws_ensure(vb::WS_OUTMATCH, (size_t)n1 * 8))) return rc;
    VB_CUDA(cudaMemcpyAsync(ctx->ws[vb::WS_P1].p, p1, (size_t)n1 * 8, cudaMemcpyHostToDevice, ctx->stream));
    VB_CUDA(cudaMemcpyAsync(ctx->ws[vb::WS_P2].p, p2, (size_t)n2 * 8, cudaMemcpyHostToDevice, ctx->stream));
    VB_CUDA(cudaMemcpyAsync(ctx->ws[vb::WS_L2A].p, d1, (size_t)n1 * dim * 4, cudaMemcpyHostToDevice, ctx->stream));
    VB_CUDA(cudaMemcpyAsync(ctx->ws[vb::WS_L2B].p, d2, (size_t)n2 * dim * 4, cudaMemcpyHostToDevice, ctx->stream));
    rc = vb::l2_pair_device(ctx, ctx->ws[vb::WS_P1].as<float2>(), ctx->ws[vb::WS_L2A].as<float>(), n1, ctx->ws[vb::WS_P2].as<float2>(),
                            ctx->ws[vb::WS_L2B].as<float>(), n2, dim, *params, ctx->ws[vb::WS_RESULT].as<vb_pair_result>(),
                            ctx->ws[vb::WS_OUTMATCH].as<int2>());
    if (rc) {   // the uploads above may still be reading the caller's buffers
        cudaStreamSynchronize(ctx->stream);
        return rc;
    }
    VB_CUDA(cudaMemcpyAsync(result, ctx->ws[vb::WS_RESULT].p, sizeof(vb_pair_result), cudaMemcpyDeviceToHost, ctx->stream));
    VB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (out_matches && result->n_matches > 0)
        VB_CUDA(cudaMemcpy(out_matches, ctx->ws[vb::WS_OUTMATCH].p, (size_t)result->n_matches * 8, cudaMemcpyDeviceToHost));
    return VB_OK;
}

}  // extern "C"
