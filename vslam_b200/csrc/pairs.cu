// pairs.cu — whole-pair pipeline: match_features (reference src/Frame.cpp:82-105) for one pair or for
// every consecutive pair of a frame sequence, with all intermediates kept on the device.
//
// Per batch of P pairs the launch sequence is fixed (independent of P). Large 256-bit problems:
//   k_expand_e2m1 -> k_knn2_tc4 (tcgen05 kNN candidates) -> k_knn2_tc_fix (exact XOR+POPC on the candidate group)
//   -> k_knn2_finish (ratio test, ordered compaction, float4 correspondences, match count)
//   -> k_sample_sets -> k_solve8 -> k_bq_init + k_count_queue (bounded inlier counting; k_corr_bounds + k_count2 for small
//   batches) -> k_select (tie scores, best model, mask, inlier matches in order).
// Small or non-256-bit problems replace the first three by k_knn2_partial (XOR + POPC tiles).
// Pair i samples with std::mt19937(seed0 + i), i.e. what the reference would draw had
// std::random_device returned seed0 + i for that frame.
#include "common.cuh"
#include "hamming_dev.cuh"
#include "ransac_dev.cuh"
#include "pairs_dev.cuh"

namespace vb {

int pairs_core(vb_ctx *ctx, uint32_t P, const float2 *p1_base, const float2 *p2_base, size_t pts_stride,
                      const uint32_t *d1_base, const uint32_t *d2_base, size_t desc_stride_words, uint32_t n1, uint32_t n2,
                      uint32_t bytes, const vb_pair_params &prm, uint32_t seed0, vb_pair_result *results_d,
                      int2 *out_matches_d) {
    int rc;
    HammingPlan hp;
    if ((rc = hamming_plan(ctx, P, n1, n2, bytes, &hp))) return rc;
    if ((rc = ctx->ws_ensure(WS_TENT, (size_t)P * n1 * sizeof(int2)))) return rc;
    if ((rc = ctx->ws_ensure(WS_CORR, (size_t)P * n1 * sizeof(float4)))) return rc;
    if ((rc = ctx->ws_ensure(WS_M, (size_t)P * sizeof(uint32_t)))) return rc;
    KnnFinishArgs fin;
    memset(&fin, 0, sizeof(fin));
    fin.ratio = prm.ratio;
    fin.tent = ctx->ws[WS_TENT].as<int2>();
    fin.mcap = n1;
    fin.m_out = ctx->ws[WS_M].as<uint32_t>();
    fin.corr = ctx->ws[WS_CORR].as<float4>();
    fin.p1_base = p1_base;
    fin.p2_base = p2_base;
    fin.pts_stride = pts_stride;
    if ((rc = hamming_launch(ctx, hp, d1_base, d2_base, desc_stride_words, fin))) return rc;
    RansacPlan rp;
    if ((rc = ransac_plan(ctx, P, n1, n1, prm.max_iterations, prm.min_items, &rp))) return rc;
    ProblemDims dims{ctx->ws[WS_M].as<uint32_t>(), 0, seed0};
    return ransac_run(ctx, rp, ctx->ws[WS_CORR].as<float4>(), dims, prm.threshold, results_d, nullptr,
                      ctx->ws[WS_TENT].as<int2>(), out_matches_d, true);
}

int pairs_check_params(const vb_pair_params *p, uint32_t bytes, uint32_t n2) {
    VB_REQUIRE(p != nullptr, VB_ERR_INVALID, "params is NULL");
    VB_REQUIRE(p->min_items >= 1 && p->min_items <= 8, VB_ERR_INVALID, "min_items must be in 1..8");
    VB_REQUIRE(p->max_iterations > 0, VB_ERR_INVALID, "max_iterations is 0");
    VB_REQUIRE(bytes == 16 || bytes == 32 || bytes == 64, VB_ERR_INVALID, "descriptor bytes must be 16, 32 or 64");
    VB_REQUIRE(n2 >= 2, VB_ERR_TOO_FEW, "knnMatch(k=2) needs at least 2 train descriptors");
    return VB_OK;
}

}  // namespace vb

using namespace vb;

extern "C" {

int vb_pairs_run_d(vb_ctx *ctx, const float *pts_d, const uint8_t *desc_d, uint32_t nframes, uint32_t k, uint32_t bytes,
                   const vb_pair_params *params, vb_pair_result *results_d, int32_t *out_matches_d) {
    VB_REQUIRE(ctx && pts_d && desc_d && results_d, VB_ERR_INVALID, "NULL argument");
    int rc;
    if ((rc = pairs_check_params(params, bytes, k))) return rc;
    if (nframes < 2) return VB_OK;
    VB_CUDA(cudaSetDevice(ctx->device));
    const uint32_t P = nframes - 1, W = bytes / 4;
    const float2 *pts = reinterpret_cast<const float2 *>(pts_d);
    const uint32_t *desc = reinterpret_cast<const uint32_t *>(desc_d);
    // Batches alternate between the context and its twin (own stream + workspaces) so that the tail of one batch's kernels
    // overlaps the head of the next one's; the twin starts after everything already queued on the caller's stream and
    // the caller's stream ends by waiting for the twin.
    // Off by default since the counting stage became one persistent kernel per batch: its fixed cost (the rounds of the last
    // problems, ~0.09 ms) is paid per launch and two halves no longer overlap anything else — 5.13 ms whole against 5.20 ms
    // split per 1 024 pairs. The split survives as option pairs_split of a -DVB_TUNING build.
#ifdef VB_TUNING
    const bool use_twin = ctx->opt("pairs_split", 0) != 0;
#else
    const bool use_twin = false;
#endif
    // (not while the profiling brackets are on: they time the context's own stream, one whole batch at a time)
    const bool split = use_twin && !ctx->profile && P >= 512;
    const uint32_t batch = split ? (P < 2 * PAIRS_MAX_BATCH ? (P + 1) / 2 : PAIRS_MAX_BATCH) : PAIRS_MAX_BATCH;
    const bool twin = split && P > batch;
    if (twin) {
        if (!ctx->twin) {
            vb_ctx *t = new vb_ctx();
            t->device = ctx->device;
            t->sm_count = ctx->sm_count;
            if (cudaStreamCreateWithFlags(&t->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
                delete t;
                set_error("cudaStreamCreate failed for the twin context");
                return VB_ERR_CUDA;
            }
            t->stream = t->own_stream;
            t->opt_parent = ctx;
            ctx->twin = t;
        }
        while (ctx->events.size() < 2) {
            cudaEvent_t e;
            VB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->events.push_back(e);
        }
        VB_CUDA(cudaEventRecord(ctx->events[0], ctx->stream));
        VB_CUDA(cudaStreamWaitEvent(ctx->twin->stream, ctx->events[0], 0));
    }
    uint32_t nb = 0;
    for (uint32_t b0 = 0; b0 < P; b0 += batch, nb++) {
        const uint32_t pb = (P - b0 < batch) ? P - b0 : batch;
        vb_ctx *cx = (twin && (nb & 1)) ? ctx->twin : ctx;
        rc = pairs_core(cx, pb, pts + (size_t)b0 * k, pts + (size_t)(b0 + 1) * k, k, desc + (size_t)b0 * k * W,
                        desc + (size_t)(b0 + 1) * k * W, (size_t)k * W, k, k, bytes, *params, params->seed0 + b0,
                        results_d + b0, out_matches_d ? reinterpret_cast<int2 *>(out_matches_d) + (size_t)b0 * k : nullptr);
        if (rc) return rc;
    }
    if (twin) {
        VB_CUDA(cudaEventRecord(ctx->events[1], ctx->twin->stream));
        VB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->events[1], 0));
    }
    return VB_OK;
}

int vb_pairs_run(vb_ctx *ctx, const float *pts, const uint8_t *desc, uint32_t nframes, uint32_t k, uint32_t bytes,
                 const vb_pair_params *params, vb_pair_result *results, int32_t *out_matches) {
    VB_REQUIRE(ctx && pts && desc && results, VB_ERR_INVALID, "NULL argument");
    int rc;
    if ((rc = pairs_check_params(params, bytes, k))) return rc;
    if (nframes < 2) return VB_OK;
    VB_CUDA(cudaSetDevice(ctx->device));
    const uint32_t P = nframes - 1;
    const size_t pts_bytes = (size_t)nframes * k * 8, desc_bytes = (size_t)nframes * k * bytes;
    if ((rc = ctx->ws_ensure(WS_PTS, pts_bytes))) return rc;
    if ((rc = ctx->ws_ensure(WS_DESC, desc_bytes))) return rc;
    if ((rc = ctx->ws_ensure(WS_RESULT, (size_t)P * sizeof(vb_pair_result)))) return rc;
    if (out_matches && (rc = ctx->ws_ensure(WS_OUTMATCH, (size_t)P * k * 8))) return rc;
    // Software pipeline over sub-batches: the upload of batch b+1 and the download of batch b-1 run on two
    // copy streams while batch b computes (they only overlap when the caller's buffers are pinned).
    std::vector<uint32_t> cut;   // cut[b] .. cut[b+1] = pairs of sub-batch b
    cut.push_back(0);
    if (P >= 512) {
        // doubling sizes: each upload is hidden by the (about equally long) compute of the batch before it; the first
        // batch is tiny because nothing hides its upload, the last one moderate because nothing hides its download
        uint32_t sz = 32;
        while (P - cut.back() > sz + 192) {
            cut.push_back(cut.back() + sz);
            if (sz < 384) sz *= 2;
        }
        if (P - cut.back() > 256) cut.push_back(P - 192);
    }
    cut.push_back(P);
#ifdef VB_TUNING
    if (const char *e = getenv("VB_PAIRS_SCHEDULE")) {   // measurement override: comma-separated sub-batch sizes
        cut.assign(1, 0u);
        for (const char *q = e; *q && cut.back() < P;) {
            const uint32_t sz = (uint32_t)strtoul(q, const_cast<char **>(&q), 10);
            if (sz == 0) break;
            cut.push_back(cut.back() + sz < P ? cut.back() + sz : P);
            if (*q == ',') q++;
        }
        if (cut.back() < P) cut.push_back(P);
    }
#endif
    const uint32_t nb = (uint32_t)cut.size() - 1;
    if (!ctx->copy_in) {
        VB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
        VB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    }
    // Odd sub-batches run on a twin context (its own stream and workspaces) so that the tail of one sub-batch's kernels
    // overlaps the head of the next one's: sub-batches are too small to fill the machine through every kernel.
#ifdef VB_TUNING
    const bool use_twin = ctx->opt("pairs_twin", 1) != 0;
#else
    const bool use_twin = true;
#endif
    if (use_twin && nb > 1 && !ctx->twin) {
        vb_ctx *t = new vb_ctx();
        t->device = ctx->device;
        t->sm_count = ctx->sm_count;
        if (cudaStreamCreateWithFlags(&t->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete t;
            set_error("cudaStreamCreate failed for the twin context");
            return VB_ERR_CUDA;
        }
        t->stream = t->own_stream;
        t->opt_parent = ctx;
        ctx->twin = t;
    }
    while (ctx->events.size() < 2 * (size_t)nb + 1) {
        cudaEvent_t e;
        VB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->events.push_back(e);
    }
    const uint32_t W = bytes / 4;
    float *pts_d = ctx->ws[WS_PTS].as<float>();
    uint8_t *desc_d = ctx->ws[WS_DESC].as<uint8_t>();
    vb_pair_result *res_d = ctx->ws[WS_RESULT].as<vb_pair_result>();
    int2 *outm_d = out_matches ? ctx->ws[WS_OUTMATCH].as<int2>() : nullptr;
    // Everything that enqueues work runs inside `enqueue`: if any step fails after the first asynchronous copy, the DMA
    // engines may still be reading the caller's pts / desc or writing its results / out_matches, so every stream is
    // drained before the error is returned (the caller is free to release those buffers as soon as this function returns).
    auto enqueue = [&]() -> int {
        // earlier work on the compute stream may still use these buffers
        cudaEvent_t ev_prev = ctx->events[2 * nb];
        VB_CUDA(cudaEventRecord(ev_prev, ctx->stream));
        VB_CUDA(cudaStreamWaitEvent(ctx->copy_in, ev_prev, 0));
        for (uint32_t b = 0; b < nb; b++) {
            const uint32_t f0 = b == 0 ? 0 : cut[b] + 1;   // first frame not yet uploaded
            const uint32_t f1 = cut[b + 1] + 1;            // one past the halo frame
            VB_CUDA(cudaMemcpyAsync(pts_d + (size_t)f0 * k * 2, pts + (size_t)f0 * k * 2, (size_t)(f1 - f0) * k * 8,
                                    cudaMemcpyHostToDevice, ctx->copy_in));
            VB_CUDA(cudaMemcpyAsync(desc_d + (size_t)f0 * k * bytes, desc + (size_t)f0 * k * bytes, (size_t)(f1 - f0) * k * bytes,
                                    cudaMemcpyHostToDevice, ctx->copy_in));
            VB_CUDA(cudaEventRecord(ctx->events[b], ctx->copy_in));
        }
        const float2 *pts2 = reinterpret_cast<const float2 *>(pts_d);
        const uint32_t *desc32 = reinterpret_cast<const uint32_t *>(desc_d);
        for (uint32_t b = 0; b < nb; b++) {
            const uint32_t p0 = cut[b], pb = cut[b + 1] - cut[b];
            vb_ctx *cx = (use_twin && ctx->twin && (b & 1)) ? ctx->twin : ctx;
            VB_CUDA(cudaStreamWaitEvent(cx->stream, ctx->events[b], 0));
            rc = pairs_core(cx, pb, pts2 + (size_t)p0 * k, pts2 + (size_t)(p0 + 1) * k, k, desc32 + (size_t)p0 * k * W,
                            desc32 + (size_t)(p0 + 1) * k * W, (size_t)k * W, k, k, bytes, *params, params->seed0 + p0,
                            res_d + p0, outm_d ? outm_d + (size_t)p0 * k : nullptr);
            if (rc) return rc;
            VB_CUDA(cudaEventRecord(ctx->events[nb + b], cx->stream));
            VB_CUDA(cudaStreamWaitEvent(ctx->copy_out, ctx->events[nb + b], 0));
            VB_CUDA(cudaMemcpyAsync(results + p0, res_d + p0, (size_t)pb * sizeof(vb_pair_result), cudaMemcpyDeviceToHost,
                                    ctx->copy_out));
            if (out_matches)
                VB_CUDA(cudaMemcpyAsync(out_matches + (size_t)p0 * k * 2, outm_d + (size_t)p0 * k, (size_t)pb * k * 8,
                                        cudaMemcpyDeviceToHost, ctx->copy_out));
        }
    return VB_OK;
    };
    rc = enqueue();
    if (rc != VB_OK) {
        cudaStreamSynchronize(ctx->copy_in);
        cudaStreamSynchronize(ctx->copy_out);
        cudaStreamSynchronize(ctx->stream);
        if (ctx->twin) cudaStreamSynchronize(ctx->twin->stream);
        return rc;
    }
    VB_CUDA(cudaStreamSynchronize(ctx->copy_out));
    VB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->twin) {
        VB_CUDA(cudaStreamSynchronize(ctx->twin->stream));
    }
    return VB_OK;
}

int vb_match_features(vb_ctx *ctx, const float *p1, const uint8_t *d1, uint32_t n1, const float *p2, const uint8_t *d2,
                      uint32_t n2, uint32_t bytes, const vb_pair_params *params, int32_t *out_matches, vb_pair_result *result) {
    VB_REQUIRE(ctx && p1 && d1 && p2 && d2 && result, VB_ERR_INVALID, "NULL argument");
    int rc;
    if ((rc = pairs_check_params(params, bytes, n2))) return rc;
    VB_REQUIRE(n1 > 0, VB_ERR_TOO_FEW, "no query keypoints");
    VB_CUDA(cudaSetDevice(ctx->device));
    if ((rc = ctx->ws_ensure(WS_P1, (size_t)n1 * 8))) return rc;
    if ((rc = ctx->ws_ensure(WS_P2, (size_t)n2 * 8))) return rc;
    if ((rc = ctx->ws_ensure(WS_D1, (size_t)n1 * bytes))) return rc;
    if ((rc = ctx->ws_ensure(WS_D2, (size_t)n2 * bytes))) return rc;
    if ((rc = ctx->ws_ensure(WS_RESULT, sizeof(vb_pair_result)))) return rc;
    if ((rc = ctx->ws_ensure(WS_OUTMATCH, (size_t)n1 * 8))) return rc;
    VB_CUDA(cudaMemcpyAsync(ctx->ws[WS_P1].p, p1, (size_t)n1 * 8, cudaMemcpyHostToDevice, ctx->stream));
    VB_CUDA(cudaMemcpyAsync(ctx->ws[WS_P2].p, p2, (size_t)n2 * 8, cudaMemcpyHostToDevice, ctx->stream));
    VB_CUDA(cudaMemcpyAsync(ctx->ws[WS_D1].p, d1, (size_t)n1 * bytes, cudaMemcpyHostToDevice, ctx->stream));
    VB_CUDA(cudaMemcpyAsync(ctx->ws[WS_D2].p, d2, (size_t)n2 * bytes, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = pairs_core(ctx, 1, ctx->ws[WS_P1].as<float2>(), ctx->ws[WS_P2].as<float2>(), 0, ctx->ws[WS_D1].as<uint32_t>(),
                         ctx->ws[WS_D2].as<uint32_t>(), 0, n1, n2, bytes, *params, params->seed0,
                         ctx->ws[WS_RESULT].as<vb_pair_result>(), ctx->ws[WS_OUTMATCH].as<int2>())))
        return rc;
    VB_CUDA(cudaMemcpyAsync(result, ctx->ws[WS_RESULT].p, sizeof(vb_pair_result), cudaMemcpyDeviceToHost, ctx->stream));
    VB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (out_matches && result->n_matches > 0)
        VB_CUDA(cudaMemcpy(out_matches, ctx->ws[WS_OUTMATCH].p, (size_t)result->n_matches * 8, cudaMemcpyDeviceToHost));
    return VB_OK;
}

}  // extern "C"
