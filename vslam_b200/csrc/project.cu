// project.cu — search by projection: the loop at reference src/vslam.cpp:129-161 (with orb_distance,
// src/PointMap.cpp:36-46) as a batch over all map points.
//
//   for each map point i, in index order:                                  (reference, sequential)
//       (x, y) = project(point i); skip unless inside the W x H image      :131-145
//       for idx in radius_search(frame.kdtree, frame.points, (x, y), 2):   :149   (DFS pre-order)
//           if frame.map_point_ids[idx] >= 0: continue                      :151   (already claimed)
//           if orb_distance(pm, i, frame, idx) < 64: claim idx; break      :152-158
//
// The claim makes it sequential: a keypoint taken by map point j is gone for every i > j. That is a
// matching with one common priority order (lower map index first) in which every map point walks down
// its own candidate list, so its unique stable outcome is the sequential one, and deferred acceptance
// finds it in parallel:
//   k_sbp_project   one thread per map point: `points * c2.t()` in the arithmetic cv::gemm uses for that
//                   shape (fp64 accumulate below 100 rows, fp32 sequential from 100 rows on; pinned against
//                   cv2 4.13 by the oracle's golden test), x/h, y/h, the in-image test. Points outside get a
//                   far-away query, so the radius search finds nothing for them.
//   k_kd_radius     (kdtree.cu) the reference's radius search, CSR result in pre-order.
//   k_sbp_filter    one thread per map point: for each candidate keypoint, free-at-entry test and
//                   orb_distance (min XOR+POPC distance over the point's observations) < threshold.
//   k_sbp_round     one thread per map point: skip candidates that failed the filter or are held by a
//                   lower map index, then atomicMin the own index into the candidate's owner slot.
//                   Repeated until a round changes no owner (owners only ever decrease): then every
//                   proposer holds its candidate, which is exactly the reference's outcome.
//   k_sbp_finish    assign[i] / map_point_ids[idx] and the claim count.
#include <algorithm>
#include <climits>

#include "common.cuh"

namespace vb {

int kd_radius_ws(vb_tree *t, const float2 *q_d, uint32_t nq, float radius, uint64_t *total_out);
int kd_radius_ws_async(vb_tree *t, const float2 *q_d, uint32_t nq, float radius, uint32_t cap);

struct Cam34 {
    float c[12];
};

__global__ void __launch_bounds__(256) k_sbp_project(const float4 *__restrict__ X, uint32_t n, Cam34 cam, float Wf, float Hf,
                                                     int small, float2 *__restrict__ q, float2 *__restrict__ proj,
                                                     uint8_t *__restrict__ in_view) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 x = X[i];
    float r[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const float *c = cam.c + 4 * j;
        if (small) {   // OpenCV's own kernel: products and sums in double, one rounding (products of floats are exact)
            double acc = __dmul_rn((double)x.x, (double)c[0]);
            acc = __dadd_rn(acc, __dmul_rn((double)x.y, (double)c[1]));
            acc = __dadd_rn(acc, __dmul_rn((double)x.z, (double)c[2]));
            acc = __dadd_rn(acc, __dmul_rn((double)x.w, (double)c[3]));
            r[j] = __double2float_rn(acc);
        } else {       // BLAS sgemm path: ((x0*c0 + x1*c1) + x2*c2) + x3*c3 in fp32, one rounding per operation
            float acc = __fmul_rn(x.x, c[0]);
            acc = __fadd_rn(acc, __fmul_rn(x.y, c[1]));
            acc = __fadd_rn(acc, __fmul_rn(x.z, c[2]));
            acc = __fadd_rn(acc, __fmul_rn(x.w, c[3]));
            r[j] = acc;
        }
    }
    const float px = __fdiv_rn(r[0], r[2]), py = __fdiv_rn(r[1], r[2]);   // :139-140
    const bool in = px >= 0.f && px < Wf && py >= 0.f && py < Hf;         // :141 (false for NaN)
    if (proj) proj[i] = make_float2(px, py);
    if (in_view) in_view[i] = in ? 1 : 0;
    q[i] = in ? make_float2(px, py) : make_float2(-3.0e38f, -3.0e38f);
}

template <int W>
__global__ void __launch_bounds__(128) k_sbp_filter(const uint32_t *__restrict__ offs, const uint32_t *__restrict__ cand,
                                                    uint32_t cap, uint32_t n, const uint32_t *__restrict__ desc,
                                                    const int32_t *__restrict__ ids, const uint32_t *__restrict__ obs_off,
                                                    const uint32_t *__restrict__ obs, uint32_t thr,
                                                    uint8_t *__restrict__ accept) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t o0 = obs_off[i], o1 = obs_off[i + 1];
    for (uint32_t c = offs[i]; c < offs[i + 1] && c < cap; c++) {
        const uint32_t idx = cand[c];
        uint8_t ok = 0;
        if (ids[idx] < 0) {   // :151 — free when the search starts; claims made during it are the rounds' business
            uint32_t a[W];
#pragma unroll
            for (int w = 0; w < W; w++) a[w] = __ldg(desc + (size_t)idx * W + w);
            uint32_t mn = 0xffffffffu;   // u32_max, src/PointMap.cpp:37
            for (uint32_t o = o0; o < o1; o++) {
                uint32_t d = 0;
#pragma unroll
                for (int w = 0; w < W; w++) d += __popc(a[w] ^ __ldg(obs + (size_t)o * W + w));
                mn = min(mn, d);
            }
            ok = mn < thr ? 1 : 0;   // :153
        }
        accept[c] = ok;
    }
}

__global__ void __launch_bounds__(256) k_sbp_round(const uint32_t *__restrict__ offs, const uint32_t *__restrict__ cand,
                                                   const uint8_t *__restrict__ accept, uint32_t n, uint32_t *__restrict__ cur,
                                                   int32_t *__restrict__ owner, uint32_t *__restrict__ changed) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c = cur[i];
    const uint32_t end = offs[i + 1];
    while (c < end && (!accept[c] || owner[cand[c]] < (int32_t)i)) c++;
    cur[i] = c;
    if (c < end) {
        const int32_t old = atomicMin(owner + cand[c], (int32_t)i);
        if (old > (int32_t)i) *changed = 1;
    }
}

// Every round of the deferred acceptance in ONE launch: a grid of co-resident CTAs (grid-stride over the map points) with
// a sense-reversing barrier in global memory between rounds, until a round lowers no owner. The host is not involved:
// vb_search_by_projection used to synchronise once per round (0.44 ms around 0.14 ms of device work). Owners are read with
// ld.cg (they change under atomicMin in L2); the three `changed` flags rotate so that the flag a slow CTA still reads is
// never the one a fast CTA clears for the round after next.
__device__ __forceinline__ void sbp_grid_barrier(unsigned int *bar, unsigned int nblocks, unsigned int &sense) {
    __syncthreads();
    if (threadIdx.x == 0) {
        sense ^= 1u;
        __threadfence();
        if (atomicAdd(&bar[0], 1u) == nblocks - 1u) {
            bar[0] = 0u;
            __threadfence();
            *reinterpret_cast<volatile unsigned int *>(&bar[1]) = sense;
        } else {
            while (*reinterpret_cast<volatile unsigned int *>(&bar[1]) != sense) __nanosleep(64);
        }
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) k_sbp_converge(const uint32_t *__restrict__ offs, const uint32_t *__restrict__ cand,
                                                      const uint8_t *__restrict__ accept, uint32_t cap, uint32_t n,
                                                      uint32_t *__restrict__ cur, int32_t *owner, unsigned int *flags /* [3] changed, [3..4] barrier, [5] rounds */) {
    unsigned int sense = 0;
    for (uint32_t round = 0;; round++) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<volatile unsigned int *>(&flags[(round + 1) % 3]) = 0u;
        bool any = false;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            uint32_t c = cur[i];
            uint32_t end = offs[i + 1];
            if (end > cap) end = cap;
            while (c < end && (!accept[c] || __ldcg(owner + cand[c]) < (int32_t)i)) c++;
            cur[i] = c;
            if (c < end) {
                const int32_t old = atomicMin(owner + cand[c], (int32_t)i);
                if (old > (int32_t)i) any = true;
            }
        }
        if (any) *reinterpret_cast<volatile unsigned int *>(&flags[round % 3]) = 1u;
        sbp_grid_barrier(flags + 3, gridDim.x, sense);
        const unsigned int changed = *reinterpret_cast<volatile unsigned int *>(&flags[round % 3]);
        if (!changed || round > n) {
            if (blockIdx.x == 0 && threadIdx.x == 0) flags[5] = round + 1u;
            return;
        }
    }
}

__global__ void __launch_bounds__(256) k_sbp_finish(const uint32_t *__restrict__ offs, const uint32_t *__restrict__ cand,
                                                    uint32_t cap, uint32_t n, const uint32_t *__restrict__ cur, int32_t *__restrict__ assign,
                                                    int32_t *__restrict__ ids, uint32_t *__restrict__ count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = cur[i];
    if (c < min(offs[i + 1], cap)) {
        const uint32_t idx = cand[c];
        assign[i] = (int32_t)idx;
        ids[idx] = (int32_t)i;   // :154 — one owner per keypoint by construction
        atomicAdd(count, 1u);
    } else {
        assign[i] = -1;
    }
}

__global__ void __launch_bounds__(256) k_fill_i32(int32_t *p, uint32_t n, int32_t v) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void __launch_bounds__(256) k_sbp_init_cur(const uint32_t *__restrict__ offs, uint32_t n, uint32_t *__restrict__ cur) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cur[i] = offs[i];
}

}  // namespace vb

using namespace vb;

extern "C" {

int vb_search_by_projection(vb_ctx *ctx, vb_tree *tree, const float *map_points, uint32_t n, const float *camera, int width,
                            int height, const uint8_t *frame_desc, uint32_t bytes, int32_t *map_point_ids,
                            const uint32_t *obs_offsets, const uint8_t *obs_desc, float radius, uint32_t dist_threshold,
                            int32_t *assign, float *proj_xy, uint8_t *in_view, uint32_t *n_claimed) {
    VB_REQUIRE(ctx && tree && camera && map_point_ids && assign, VB_ERR_INVALID, "NULL argument");
    VB_REQUIRE(tree->ctx == ctx, VB_ERR_INVALID, "tree belongs to another context");
    VB_REQUIRE(bytes == 16 || bytes == 32 || bytes == 64, VB_ERR_INVALID, "descriptor bytes must be 16, 32 or 64");
    VB_REQUIRE(n == 0 || (map_points && obs_offsets), VB_ERR_INVALID, "NULL argument");
    if (n_claimed) *n_claimed = 0;
    if (n == 0) return VB_OK;
    const uint32_t k = tree->n;
    VB_REQUIRE(k == 0 || frame_desc, VB_ERR_INVALID, "frame_desc is NULL");
    VB_CUDA(cudaSetDevice(ctx->device));
    const uint32_t nobs = obs_offsets[n];
    VB_REQUIRE(nobs == 0 || obs_desc, VB_ERR_INVALID, "obs_desc is NULL");
    int rc;
    if ((rc = ctx->ws_ensure(WS_SBP_X, (size_t)n * 16))) return rc;
    if ((rc = ctx->ws_ensure(WS_SBP_Q, (size_t)n * 8 * 2 + n))) return rc;
    if ((rc = ctx->ws_ensure(WS_SBP_DESC, (size_t)(k ? k : 1) * bytes))) return rc;
    if ((rc = ctx->ws_ensure(WS_SBP_IDS, (size_t)(k ? k : 1) * 4))) return rc;
    if ((rc = ctx->ws_ensure(WS_SBP_OBSOFF, (size_t)(n + 1) * 4))) return rc;
    if ((rc = ctx->ws_ensure(WS_SBP_OBS, (size_t)(nobs ? nobs : 1) * bytes))) return rc;
    if ((rc = ctx->ws_ensure(WS_SBP_CUR, (size_t)n * 4 + 32))) return rc;
    if ((rc = ctx->ws_ensure(WS_SBP_OWNER, (size_t)(k ? k : 1) * 4))) return rc;
    if ((rc = ctx->ws_ensure(WS_SBP_ASSIGN, (size_t)n * 4))) return rc;
    cudaStream_t st = ctx->stream;
    float4 *X_d = ctx->ws[WS_SBP_X].as<float4>();
    float2 *q_d = ctx->ws[WS_SBP_Q].as<float2>();
    float2 *proj_d = q_d + n;
    uint8_t *inview_d = reinterpret_cast<uint8_t *>(proj_d + n);
    uint32_t *desc_d = ctx->ws[WS_SBP_DESC].as<uint32_t>();
    int32_t *ids_d = ctx->ws[WS_SBP_IDS].as<int32_t>();
    uint32_t *obsoff_d = ctx->ws[WS_SBP_OBSOFF].as<uint32_t>();
    uint32_t *obs_d = ctx->ws[WS_SBP_OBS].as<uint32_t>();
    uint32_t *cur_d = ctx->ws[WS_SBP_CUR].as<uint32_t>();
    uint32_t *flags_d = cur_d + n;   // [0..2] changed (rotating), [3..4] grid barrier, [5] rounds run, [6] claim count
    int32_t *owner_d = ctx->ws[WS_SBP_OWNER].as<int32_t>();
    int32_t *assign_d = ctx->ws[WS_SBP_ASSIGN].as<int32_t>();
    VB_CUDA(cudaMemcpyAsync(X_d, map_points, (size_t)n * 16, cudaMemcpyHostToDevice, st));
    if (k) VB_CUDA(cudaMemcpyAsync(desc_d, frame_desc, (size_t)k * bytes, cudaMemcpyHostToDevice, st));
    if (k) VB_CUDA(cudaMemcpyAsync(ids_d, map_point_ids, (size_t)k * 4, cudaMemcpyHostToDevice, st));
    VB_CUDA(cudaMemcpyAsync(obsoff_d, obs_offsets, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, st));
    if (nobs) VB_CUDA(cudaMemcpyAsync(obs_d, obs_desc, (size_t)nobs * bytes, cudaMemcpyHostToDevice, st));
    Cam34 cam;
    memcpy(cam.c, camera, sizeof(cam.c));
    // No host synchronisation between the upload and the download: the candidate array is sized for 16 hits per map point
    // (or whatever an earlier call needed) and the whole pass is repeated with the exact size in the rare case that was
    // not enough — the total is only known once the results are back.
    static_assert(sizeof(unsigned long long) == 8, "");
    uint32_t cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>((uint64_t)n * 16, ctx->sbp_cap_hint), 0xfffffff0ull);
    uint64_t total = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        ctx->prof_begin("sbp");
        k_sbp_project<<<div_up(n, 256), 256, 0, st>>>(X_d, n, cam, (float)width, (float)height, n < 100 ? 1 : 0, q_d, proj_d, inview_d);
        ctx->launches++;
        if ((rc = kd_radius_ws_async(tree, q_d, n, radius, cap))) return rc;
        const uint32_t *offs_d = ctx->ws[WS_OUT0].as<uint32_t>();
        const uint32_t *cand_d = ctx->ws[WS_OUT1].as<uint32_t>();
        if ((rc = ctx->ws_ensure(WS_SBP_ACC, (size_t)cap))) return rc;
        uint8_t *acc_d = ctx->ws[WS_SBP_ACC].as<uint8_t>();
        switch (bytes / 4) {
            case 4: k_sbp_filter<4><<<div_up(n, 128), 128, 0, st>>>(offs_d, cand_d, cap, n, desc_d, ids_d, obsoff_d, obs_d, dist_threshold, acc_d); break;
            case 8: k_sbp_filter<8><<<div_up(n, 128), 128, 0, st>>>(offs_d, cand_d, cap, n, desc_d, ids_d, obsoff_d, obs_d, dist_threshold, acc_d); break;
            default: k_sbp_filter<16><<<div_up(n, 128), 128, 0, st>>>(offs_d, cand_d, cap, n, desc_d, ids_d, obsoff_d, obs_d, dist_threshold, acc_d); break;
        }
        if (k) k_fill_i32<<<div_up(k, 256), 256, 0, st>>>(owner_d, k, INT_MAX);
        k_sbp_init_cur<<<div_up(n, 256), 256, 0, st>>>(offs_d, n, cur_d);
        VB_CUDA(cudaMemsetAsync(flags_d, 0, 32, st));
        int per_sm = 0;
        VB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sbp_converge, 256, 0));
        const uint32_t resident = (uint32_t)(per_sm < 1 ? 1 : per_sm) * (uint32_t)ctx->sm_count;   // the barrier needs co-residency
        const uint32_t grid = std::min(div_up(n, 256), resident);
        k_sbp_converge<<<grid, 256, 0, st>>>(offs_d, cand_d, acc_d, cap, n, cur_d, owner_d, flags_d);
        k_sbp_finish<<<div_up(n, 256), 256, 0, st>>>(offs_d, cand_d, cap, n, cur_d, assign_d, ids_d, flags_d + 6);
        ctx->launches += 5;
        ctx->prof_end("sbp");
        VB_CUDA(cudaGetLastError());
        VB_CUDA(cudaMemcpyAsync(&total, ctx->ws[WS_MISC].p, 8, cudaMemcpyDeviceToHost, st));
        if (attempt == 0 && k) {   // keep the caller's ids intact until the pass is known to be complete
            VB_CUDA(cudaStreamSynchronize(st));
            if (total > cap) {
                VB_REQUIRE(total <= 0xfffffff0ull, VB_ERR_CAPACITY, "radius result too large for 32-bit CSR offsets");
                cap = (uint32_t)total;
                ctx->sbp_cap_hint = total;
                VB_CUDA(cudaMemcpyAsync(ids_d, map_point_ids, (size_t)k * 4, cudaMemcpyHostToDevice, st));   // finish wrote claims into it
                continue;
            }
        }
        break;
    }
    uint32_t claimed = 0;
    VB_CUDA(cudaMemcpyAsync(assign, assign_d, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (k) VB_CUDA(cudaMemcpyAsync(map_point_ids, ids_d, (size_t)k * 4, cudaMemcpyDeviceToHost, st));
    if (proj_xy) VB_CUDA(cudaMemcpyAsync(proj_xy, proj_d, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    if (in_view) VB_CUDA(cudaMemcpyAsync(in_view, inview_d, (size_t)n, cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaMemcpyAsync(&claimed, flags_d + 6, 4, cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    if (n_claimed) *n_claimed = claimed;
    return VB_OK;
}

}  // extern "C"
