// hamming.cu — brute-force binary-descriptor matching: two nearest train descriptors per query
// (XOR + population count), Lowe ratio test, ordered compaction of the survivors.
//
// Replaces cv::BFMatcher(NORM_HAMMING)::knnMatch(d1, d2, k=2) and the ratio loop of match_features
// (reference src/Frame.cpp:83-95). knnMatch resolves equal distances to the lower train index
// (checked against cv2 4.13, tests/golden); scanning train descriptors in increasing index with strict
// `<` updates reproduces that without carrying indices through the comparison.
//
// k_knn2_partial<W,QPT>  each thread keeps QPT query descriptors (W 32-bit words each) in registers;
//                        the CTA streams 128-descriptor train tiles through shared memory with
//                        cp.async double buffering and every thread reads them as broadcast
//                        LDS.128. The 8-word distance uses a carry-save adder tree so that 256 bits
//                        cost 4 POPC instead of 8 (POPC is a quarter-rate instruction).
//                        grid = (query tiles, train splits, problems); a split scans a contiguous
//                        train range so per-split results merge by (distance, index).
//   k_knn2_finish        one CTA per problem: merge splits, ratio test in double exactly as :91,
//                        ballot/scan compaction in query order; for the pair pipeline it also writes
//                        the float4 correspondences and the match count RANSAC consumes.
#include <cuda_pipeline.h>

#include "common.cuh"
#include "hamming_dev.cuh"

namespace vb {

// Carry-save adder on three bit-planes, pinned to exactly two LOP3 (left to itself the compiler folds the
// preceding XORs into the majority and ends up with ~27 LOP3 per 256-bit distance instead of 16).
__device__ __forceinline__ void csa(uint32_t a, uint32_t b, uint32_t c, uint32_t &sum, uint32_t &carry) {
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(sum) : "r"(a), "r"(b), "r"(c));     // a ^ b ^ c
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(carry) : "r"(a), "r"(b), "r"(c));   // majority
}

// Integer multiply-add pinned to the FMA pipe: the multiplier is a run-time register (the kernel is handed
// the constants 1, 2, 4 as launch values), so ptxas cannot turn it into an ALU-pipe add/shift. The
// ALU pipe is the saturated one in this kernel (ncu: 91 % busy, FMA pipe 5 %).
__device__ __forceinline__ int imad(uint32_t a, int m, int c) {
    int r;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"((int)a), "r"(m), "r"(c));
    return r;
}
struct Mul124 { int m1, m2, m4; };

// (Hamming distance of W words) + acc, with 4 POPC per 8 words: ones + 2*twos + 4*fours.
template <int W>
__device__ __forceinline__ int hamming_acc(const uint32_t (&a)[W], const uint32_t *b, int acc, const Mul124 &m) {
    if constexpr (W % 8 == 0) {
#pragma unroll
        for (int g = 0; g < W / 8; g++) {
            uint32_t x[8];
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = a[g * 8 + i] ^ b[g * 8 + i];
            uint32_t s1, c1, s2, c2, s3, c3, s4, c4;
            csa(x[0], x[1], x[2], s1, c1);
            csa(x[3], x[4], x[5], s2, c2);
            csa(s1, s2, x[6], s3, c3);
            csa(c1, c2, c3, s4, c4);
            acc = imad(__popc(s3), m.m1, acc);
            acc = imad(__popc(x[7]), m.m1, acc);
            acc = imad(__popc(s4), m.m2, acc);
            acc = imad(__popc(c4), m.m4, acc);
        }
    } else {
#pragma unroll
        for (int i = 0; i < W; i++) acc = imad(__popc(a[i] ^ b[i]), m.m1, acc);
    }
    return acc;
}

constexpr int KNN_THREADS = 256;
constexpr int KNN_TILE = 128;   // train descriptors per shared-memory stage

template <int W, int QPT>
__global__ void __launch_bounds__(KNN_THREADS) k_knn2_partial(const uint32_t *__restrict__ d1_base,
                                                              const uint32_t *__restrict__ d2_base,
                                                              size_t problem_stride_words, uint32_t n1, uint32_t n2,
                                                              uint32_t split_len, uint32_t nsplits, int one,
                                                              uint2 *__restrict__ part) {
    __shared__ __align__(16) uint32_t tile[2][KNN_TILE * W];
    const Mul124 mul{one, one + one, one * 4};
    const uint32_t p = blockIdx.z, s = blockIdx.y, tid = threadIdx.x;
    const uint32_t *d1 = d1_base + (size_t)p * problem_stride_words;
    const uint32_t *d2 = d2_base + (size_t)p * problem_stride_words;
    const uint32_t t0 = s * split_len;
    const uint32_t t1 = min(t0 + split_len, n2);

    uint32_t a[QPT][W];
    uint32_t q[QPT];
#pragma unroll
    for (int k = 0; k < QPT; k++) {
        q[k] = (blockIdx.x * QPT + k) * KNN_THREADS + tid;
        const uint32_t qs = q[k] < n1 ? q[k] : 0;
        const uint4 *src = reinterpret_cast<const uint4 *>(d1 + (size_t)qs * W);
#pragma unroll
        for (int i = 0; i < W / 4; i++) {
            const uint4 v = __ldg(src + i);
            a[k][4 * i] = v.x; a[k][4 * i + 1] = v.y; a[k][4 * i + 2] = v.z; a[k][4 * i + 3] = v.w;
        }
    }
    uint32_t bd1[QPT], bj1[QPT], bd2[QPT], bj2[QPT];
#pragma unroll
    for (int k = 0; k < QPT; k++) { bd1[k] = bd2[k] = 0x3ffu; bj1[k] = bj2[k] = 0x3fffffu; }
    int nbd2[QPT];   // -bd2: the distance is accumulated on top of it, so "d < bd2" is just a sign bit
#pragma unroll
    for (int k = 0; k < QPT; k++) nbd2[k] = -(int)bd2[k];

    constexpr int PIECES = KNN_TILE * W / 4;   // 16-byte pieces per tile
    auto issue_tile = [&](uint32_t tile_start, int buf) {
        const uint32_t n_here = min((uint32_t)KNN_TILE, t1 - tile_start);
        const uint4 *src = reinterpret_cast<const uint4 *>(d2 + (size_t)tile_start * W);
        uint4 *dst = reinterpret_cast<uint4 *>(tile[buf]);
        for (uint32_t i = tid; i < (uint32_t)PIECES; i += KNN_THREADS)
            if (i < n_here * (W / 4)) __pipeline_memcpy_async(dst + i, src + i, 16);
        __pipeline_commit();
    };

    if (t0 < t1) issue_tile(t0, 0);
    int buf = 0;
    for (uint32_t ts = t0; ts < t1; ts += KNN_TILE, buf ^= 1) {
        if (ts + KNN_TILE < t1) {
            issue_tile(ts + KNN_TILE, buf ^ 1);
            __pipeline_wait_prior(1);
        } else {
            __pipeline_wait_prior(0);
        }
        __syncthreads();
        const uint32_t n_here = min((uint32_t)KNN_TILE, t1 - ts);
        const uint32_t *tb = tile[buf];
#pragma unroll 2
        for (uint32_t j = 0; j < n_here; j++) {
            uint32_t b[W];
#pragma unroll
            for (int i = 0; i < W / 4; i++) {
                const uint4 v = *reinterpret_cast<const uint4 *>(tb + j * W + 4 * i);
                b[4 * i] = v.x; b[4 * i + 1] = v.y; b[4 * i + 2] = v.z; b[4 * i + 3] = v.w;
            }
            int t[QPT];
            int sign = 0;
#pragma unroll
            for (int k = 0; k < QPT; k++) {
                t[k] = hamming_acc<W>(a[k], b, nbd2[k], mul);   // d - bd2
                sign |= t[k];
            }
            // The top-2 update is rare after the first few hundred candidates (probability ~2/j at
            // candidate j): one sign test and one warp vote per train descriptor keep it off the common
            // path instead of ~10 predicated instructions per distance.
            if (__any_sync(0xffffffffu, sign < 0)) {
                const uint32_t jg = ts + j;
#pragma unroll
                for (int k = 0; k < QPT; k++)
                    if (t[k] < 0) {   // d < bd2; equal distance never displaces an earlier (lower) index
                        const uint32_t d = (uint32_t)(t[k] + (int)bd2[k]);
                        if (d < bd1[k]) { bd2[k] = bd1[k]; bj2[k] = bj1[k]; bd1[k] = d; bj1[k] = jg; }
                        else { bd2[k] = d; bj2[k] = jg; }
                        nbd2[k] = -(int)bd2[k];
                    }
            }
        }
        __syncthreads();   // everyone is done with tile[buf] before it is refilled two iterations later
    }
#pragma unroll
    for (int k = 0; k < QPT; k++)
        if (q[k] < n1)
            part[((size_t)p * nsplits + s) * n1 + q[k]] =
                make_uint2((bd1[k] << KNN_IDX_BITS) | bj1[k], (bd2[k] << KNN_IDX_BITS) | bj2[k]);
}

__global__ void __launch_bounds__(256) k_knn2_finish(KnnFinishArgs a) {
    __shared__ int s_scan[8];
    __shared__ int s_base;
    const uint32_t p = blockIdx.x, tid = threadIdx.x;
    const int lane = tid & 31, w = tid >> 5;
    const uint2 *part = a.part + (size_t)p * a.nsplits * a.n1;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (uint32_t q0 = 0; q0 < a.n1; q0 += blockDim.x) {
        const uint32_t q = q0 + tid;
        int keep = 0;
        uint32_t k1 = 0xffffffffu, k2 = 0xffffffffu;
        if (q < a.n1) {
            for (uint32_t s = 0; s < a.nsplits; s++) {
                const uint2 v = part[(size_t)s * a.n1 + q];
                const uint32_t ks[2] = {v.x, v.y};
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    const uint32_t k = ks[t];
                    if (k < k2) {
                        if (k < k1) { k2 = k1; k1 = k; } else { k2 = k; }
                    }
                }
            }
            const int d0 = (int)(k1 >> KNN_IDX_BITS), d1 = (int)(k2 >> KNN_IDX_BITS);
            const int i0 = (int)(k1 & KNN_IDX_MASK), i1 = (int)(k2 & KNN_IDX_MASK);
            if (a.knn_idx) {
                int32_t *o = a.knn_idx + ((size_t)p * a.n1 + q) * 2;
                o[0] = i0; o[1] = i1;
                int32_t *od = a.knn_dist + ((size_t)p * a.n1 + q) * 2;
                od[0] = d0; od[1] = d1;
            }
            // src/Frame.cpp:91: m[0].distance < m[1].distance * 0.7 — float distances, double product
            keep = ((double)(float)d0 < (double)(float)d1 * a.ratio) ? 1 : 0;
        }
        if (a.tent) {
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
                const int o = base + woff + wpre;
                const int i0 = (int)(k1 & KNN_IDX_MASK);
                a.tent[(size_t)p * a.mcap + o] = make_int2((int)q, i0);
                if (a.corr) {
                    const float2 x1 = a.p1_base[(size_t)p * a.pts_stride + q];
                    const float2 x2 = a.p2_base[(size_t)p * a.pts_stride + i0];
                    a.corr[(size_t)p * a.mcap + o] = make_float4(x1.x, x1.y, x2.x, x2.y);
                }
            }
            __syncthreads();
            if (tid == 0) s_base = base + tot;
            __syncthreads();
        }
    }
    if (tid == 0 && a.m_out) a.m_out[p] = (uint32_t)s_base;
}

int hamming_plan(vb_ctx *ctx, uint32_t P, uint32_t n1, uint32_t n2, uint32_t bytes, HammingPlan *pl) {
    pl->P = P; pl->n1 = n1; pl->n2 = n2; pl->W = bytes / 4;
    // two queries per thread once the grid is large anyway (halves the shared-memory reads per pair)
    const uint32_t qt1 = div_up(n1, KNN_THREADS);
    pl->qpt = ((uint64_t)P * qt1 >= 16ull * ctx->sm_count && pl->W <= 8) ? 4
              : ((uint64_t)P * qt1 >= 8ull * ctx->sm_count) ? 2 : 1;
    if (const long long e = ctx->opt("hamming_qpt", 0)) pl->qpt = e == 4 ? 4 : e == 2 ? 2 : 1;
    pl->qtiles = div_up(n1, KNN_THREADS * pl->qpt);
    // split the train set so that a small batch still fills the machine (~4 CTAs per SM)
    uint64_t want = 4ull * ctx->sm_count;
    uint64_t have = (uint64_t)P * pl->qtiles;
    uint32_t ns = (uint32_t)((want + have - 1) / have);
    const uint32_t max_splits = div_up(n2, KNN_TILE);
    if (ns > max_splits) ns = max_splits;
    if (ns > 64) ns = 64;
    if (ns < 1) ns = 1;
    pl->split_len = div_up(div_up(n2, ns), KNN_TILE) * KNN_TILE;
    pl->nsplits = div_up(n2, pl->split_len);
    return ctx->ws_ensure(WS_KNN_PART, (size_t)P * pl->nsplits * n1 * sizeof(uint2));
}

template <int W> static void launch_partial(vb_ctx *ctx, const HammingPlan &pl, const uint32_t *d1, const uint32_t *d2,
                                            size_t stride_words) {
    dim3 grid(pl.qtiles, pl.nsplits, pl.P);
    uint2 *part = ctx->ws[WS_KNN_PART].as<uint2>();
    if (pl.qpt == 4)
        k_knn2_partial<W, 4><<<grid, KNN_THREADS, 0, ctx->stream>>>(d1, d2, stride_words, pl.n1, pl.n2, pl.split_len,
                                                                    pl.nsplits, 1, part);
    else if (pl.qpt == 2)
        k_knn2_partial<W, 2><<<grid, KNN_THREADS, 0, ctx->stream>>>(d1, d2, stride_words, pl.n1, pl.n2, pl.split_len,
                                                                    pl.nsplits, 1, part);
    else
        k_knn2_partial<W, 1><<<grid, KNN_THREADS, 0, ctx->stream>>>(d1, d2, stride_words, pl.n1, pl.n2, pl.split_len,
                                                                    pl.nsplits, 1, part);
}

int hamming_launch(vb_ctx *ctx, const HammingPlan &pl, const uint32_t *d1, const uint32_t *d2, size_t stride_words,
                   KnnFinishArgs fin) {
    uint32_t nsplits = pl.nsplits;
    const uint2 *part = nullptr;
    if (hamming_tc_eligible(ctx, pl)) {
        int rc = hamming_tc_launch(ctx, pl, d1, d2, stride_words, &part, fin.knn_idx != nullptr, fin.ratio);
        if (rc) return rc;
        nsplits = 1;
    } else {
        ctx->prof_begin("hamming");
        switch (pl.W) {
            case 4: launch_partial<4>(ctx, pl, d1, d2, stride_words); break;
            case 8: launch_partial<8>(ctx, pl, d1, d2, stride_words); break;
            case 16: launch_partial<16>(ctx, pl, d1, d2, stride_words); break;
            default: set_error("descriptor bytes must be 16, 32 or 64"); return VB_ERR_INVALID;
        }
        ctx->prof_end("hamming");
        ctx->launches++;
        VB_CUDA(cudaGetLastError());
        part = ctx->ws[WS_KNN_PART].as<uint2>();
    }
    fin.part = part;
    fin.nsplits = nsplits;
    fin.n1 = pl.n1;
    ctx->prof_begin("finish");
    k_knn2_finish<<<pl.P, 256, 0, ctx->stream>>>(fin);
    ctx->prof_end("finish");
    ctx->launches++;
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

static int hamming_host(vb_ctx *ctx, const uint8_t *d1, uint32_t n1, const uint8_t *d2, uint32_t n2, uint32_t bytes,
                        double ratio, int32_t *idx, int32_t *dist, int32_t *out_pairs, uint32_t *out_m) {
    VB_REQUIRE(ctx && d1 && d2, VB_ERR_INVALID, "NULL argument");
    VB_REQUIRE(bytes == 16 || bytes == 32 || bytes == 64, VB_ERR_INVALID, "descriptor bytes must be 16, 32 or 64");
    VB_REQUIRE(n2 >= 2, VB_ERR_TOO_FEW, "knnMatch(k=2) needs at least 2 train descriptors");
    VB_REQUIRE(n2 <= KNN_IDX_MASK, VB_ERR_INVALID, "too many train descriptors");
    if (n1 == 0) { if (out_m) *out_m = 0; return VB_OK; }
    VB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    HammingPlan pl;
    if ((rc = hamming_plan(ctx, 1, n1, n2, bytes, &pl))) return rc;
    if ((rc = ctx->ws_ensure(WS_D1, (size_t)n1 * bytes))) return rc;
    if ((rc = ctx->ws_ensure(WS_D2, (size_t)n2 * bytes))) return rc;
    if ((rc = ctx->ws_ensure(WS_KNN, (size_t)n1 * 16))) return rc;
    if ((rc = ctx->ws_ensure(WS_TENT, (size_t)n1 * 8))) return rc;
    if ((rc = ctx->ws_ensure(WS_M, 16))) return rc;
    VB_CUDA(cudaMemcpyAsync(ctx->ws[WS_D1].p, d1, (size_t)n1 * bytes, cudaMemcpyHostToDevice, ctx->stream));
    VB_CUDA(cudaMemcpyAsync(ctx->ws[WS_D2].p, d2, (size_t)n2 * bytes, cudaMemcpyHostToDevice, ctx->stream));
    KnnFinishArgs fin;
    memset(&fin, 0, sizeof(fin));
    fin.ratio = ratio;
    fin.knn_idx = ctx->ws[WS_KNN].as<int32_t>();
    fin.knn_dist = ctx->ws[WS_KNN].as<int32_t>() + (size_t)n1 * 2;
    fin.tent = ctx->ws[WS_TENT].as<int2>();
    fin.mcap = n1;
    fin.m_out = ctx->ws[WS_M].as<uint32_t>();
    if ((rc = hamming_launch(ctx, pl, ctx->ws[WS_D1].as<uint32_t>(), ctx->ws[WS_D2].as<uint32_t>(), 0, fin))) return rc;
    uint32_t m = 0;
    VB_CUDA(cudaMemcpyAsync(&m, ctx->ws[WS_M].p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (idx) VB_CUDA(cudaMemcpyAsync(idx, fin.knn_idx, (size_t)n1 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (dist) VB_CUDA(cudaMemcpyAsync(dist, fin.knn_dist, (size_t)n1 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    VB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (out_pairs && m) VB_CUDA(cudaMemcpy(out_pairs, ctx->ws[WS_TENT].p, (size_t)m * 8, cudaMemcpyDeviceToHost));
    if (out_m) *out_m = m;
    return VB_OK;
}

}  // namespace vb

extern "C" {

int vb_knn2_hamming(vb_ctx *ctx, const uint8_t *d1, uint32_t n1, const uint8_t *d2, uint32_t n2, uint32_t bytes,
                    int32_t *idx, int32_t *dist) {
    return vb::hamming_host(ctx, d1, n1, d2, n2, bytes, 0.7, idx, dist, nullptr, nullptr);
}

int vb_match_hamming(vb_ctx *ctx, const uint8_t *d1, uint32_t n1, const uint8_t *d2, uint32_t n2, uint32_t bytes,
                     double ratio, int32_t *out_pairs, uint32_t *out_m) {
    VB_REQUIRE(out_pairs && out_m, VB_ERR_INVALID, "NULL output");
    return vb::hamming_host(ctx, d1, n1, d2, n2, bytes, ratio, nullptr, nullptr, out_pairs, out_m);
}

}  // extern "C"
