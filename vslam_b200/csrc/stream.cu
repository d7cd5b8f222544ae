// stream.cu — asynchronous, double-buffered submission of frame sequences with a compact result download.
//
// vb_pairs_run (pairs.cu) is a blocking call: upload, compute and download of ONE sequence overlap only inside that
// call, so every call pays for a short first sub-batch and an exposed last download. A caller that streams sequences
// (a SLAM front end consuming a camera, the benchmark loop, one host thread per GPU in vb_multi) instead keeps two
// calls in flight:
//
//     vb_pairs_submit(n)      H2D(n)  on the upload stream      --+
//     vb_pairs_submit(n+1)    H2D(n+1) right behind it            |  compute(n) on the context's stream
//     vb_pairs_wait(n)        results(n) D2H on slot n's stream --+  compute(n+1) follows without a gap
//
// Each of the two slots owns its device-side input, result and match buffers AND its compute stream + workspaces (even
// tickets run on the context, odd ones on its twin context): consecutive submissions are independent, so the small
// kernels and the inlier counting of submission n run beside the matcher of submission n+1 where registers and shared
// memory allow (measured: 5.14 -> 4.91 ms per 1 024-pair step; option pairs_overlap = 0 serialises them on one stream). Results come back compact: per pair its vb_pair_result, and the inlier
// matches of all pairs packed back to back as (query, train) uint16 pairs (k <= 65 535) — what match_features appends to
// frame1.matches at reference src/Frame.cpp:98-102 — instead of a [pairs][k] int32 slab of which two fifths are unused:
// 12 KB instead of 40 KB per pair at k = 5 000. The exact byte count is only known once the results have landed, so the
// download is two-phase: results + offsets first, then exactly `total` matches.
#include <new>

#include "common.cuh"
#include "pairs_dev.cuh"

namespace vb {

struct PairSlot {
    DevBuf pts, desc, res, outm, pack, offs;
    cudaStream_t s_out = nullptr;
    cudaEvent_t ev_in = nullptr, ev_in0 = nullptr, ev_done = nullptr, ev_res = nullptr, ev_m = nullptr;
    bool busy = false, device_io = false;
    vb_ctx *cx = nullptr;   // the context (stream + workspaces) this slot computes on
    uint32_t P = 0;
    uint64_t base = 0, cap = 0;
    vb_pair_result *results = nullptr;
    uint32_t *offsets = nullptr;
    uint16_t *matches16 = nullptr;
    int ticket = -1;
};

struct PairsStream {
    PairSlot slot[PAIRS_DEPTH];
    int next_ticket = 0;
};

// offs[p] = base + sum of n_matches of the pairs before p; the pair's matches narrowed to uint16 pairs behind it.
__global__ void __launch_bounds__(256) k_pack_matches(const vb_pair_result *__restrict__ res, const int2 *__restrict__ outm,
                                                      uint32_t k, uint32_t base, uint32_t *__restrict__ offs,
                                                      ushort2 *__restrict__ pack) {
    __shared__ uint32_t red[8];
    const uint32_t p = blockIdx.x, tid = threadIdx.x;
    uint32_t s = 0;
    for (uint32_t q = tid; q < p; q += blockDim.x) s += (uint32_t)res[q].n_matches;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    uint32_t off = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) off += red[w];
    if (tid == 0) offs[p] = base + off;
    const uint32_t n = (uint32_t)res[p].n_matches;
    const int2 *src = outm + (size_t)p * k;
    for (uint32_t i = tid; i < n; i += blockDim.x) {
        const int2 v = src[i];
        pack[(size_t)off + i] = make_ushort2((unsigned short)v.x, (unsigned short)v.y);
    }
}

static PairsStream *stream_state(vb_ctx *ctx) {
    if (!ctx->pairs_stream) ctx->pairs_stream = new (std::nothrow) PairsStream();
    return static_cast<PairsStream *>(ctx->pairs_stream);
}

void pairs_stream_release(vb_ctx *ctx) {
    PairsStream *ps = static_cast<PairsStream *>(ctx->pairs_stream);
    if (!ps) return;
    for (PairSlot &s : ps->slot) {
        if (s.s_out) cudaStreamSynchronize(s.s_out);
        for (DevBuf *b : {&s.pts, &s.desc, &s.res, &s.outm, &s.pack, &s.offs}) b->release();
        for (cudaEvent_t e : {s.ev_in, s.ev_in0, s.ev_done, s.ev_res, s.ev_m})
            if (e) cudaEventDestroy(e);
        if (s.s_out) cudaStreamDestroy(s.s_out);
    }
    delete ps;
    ctx->pairs_stream = nullptr;
}

static void drain(vb_ctx *ctx, PairSlot &s) {
    if (ctx->copy_in) cudaStreamSynchronize(ctx->copy_in);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->twin) cudaStreamSynchronize(ctx->twin->stream);
    if (s.s_out) cudaStreamSynchronize(s.s_out);
}

// The context a ticket computes on: the caller's context for even tickets, its twin (own stream, own workspaces, options
// looked up in the parent) for odd ones. One stream when profiling brackets are on (they time one stream) or when asked.
static int compute_ctx(vb_ctx *ctx, int ticket, vb_ctx **out) {
    *out = ctx;
    if (!(ticket & 1) || ctx->profile || ctx->opt("pairs_overlap", 1) == 0) return VB_OK;
    if (!ctx->twin) {
        vb_ctx *t = new (std::nothrow) vb_ctx();
        VB_REQUIRE(t != nullptr, VB_ERR_CUDA, "out of host memory");
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
    *out = ctx->twin;
    return VB_OK;
}

int pairs_submit(vb_ctx *ctx, const float *pts, const uint8_t *desc, uint32_t nframes, uint32_t k, uint32_t bytes,
                 const vb_pair_params *params, vb_pair_result *results, uint32_t *match_offsets, uint16_t *matches16,
                 uint64_t cap, uint64_t base, int *ticket) {
    VB_REQUIRE(ctx && pts && desc && results && ticket, VB_ERR_INVALID, "NULL argument");
    VB_REQUIRE((match_offsets == nullptr) == (matches16 == nullptr), VB_ERR_INVALID,
               "match_offsets and matches16 go together (both NULL skips the match download)");
    VB_REQUIRE(matches16 == nullptr || k <= 65535u, VB_ERR_INVALID, "compact matches hold uint16 indices: k must be <= 65535");
    int rc;
    if ((rc = pairs_check_params(params, bytes, k))) return rc;
    VB_CUDA(cudaSetDevice(ctx->device));
    PairsStream *ps = stream_state(ctx);
    VB_REQUIRE(ps != nullptr, VB_ERR_CUDA, "out of host memory");
    PairSlot &s = ps->slot[ps->next_ticket % PAIRS_DEPTH];
    VB_REQUIRE(!s.busy, VB_ERR_CAPACITY, "three submissions are already in flight: vb_pairs_wait the oldest ticket first");
    const uint32_t P = nframes < 2 ? 0 : nframes - 1;
    VB_REQUIRE(base + (uint64_t)P * k <= 0xffffffffull, VB_ERR_INVALID, "match offsets are 32-bit: too many pairs x keypoints");
    s.P = P; s.base = base; s.cap = cap; s.results = results; s.offsets = match_offsets; s.matches16 = matches16;
    s.ticket = ps->next_ticket;
    s.device_io = false;
    if ((rc = compute_ctx(ctx, s.ticket, &s.cx))) return rc;
    vb_ctx *cx = s.cx;
    if (P == 0) {
        s.busy = true;
        *ticket = ps->next_ticket++;
        return VB_OK;
    }
    if (!s.s_out) {
        VB_CUDA(cudaStreamCreateWithFlags(&s.s_out, cudaStreamNonBlocking));
        for (cudaEvent_t *e : {&s.ev_in, &s.ev_in0, &s.ev_done, &s.ev_res, &s.ev_m}) VB_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    if (!ctx->copy_in) {
        VB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
        VB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    }
    const size_t pts_bytes = (size_t)nframes * k * 8, desc_bytes = (size_t)nframes * k * bytes;
    // (growing a slot buffer frees the old one: cudaFree waits for the device, so nothing in flight still uses it)
    if ((rc = s.pts.ensure(pts_bytes))) return rc;
    if ((rc = s.desc.ensure(desc_bytes))) return rc;
    if ((rc = s.res.ensure((size_t)P * sizeof(vb_pair_result)))) return rc;
    if (matches16) {
        if ((rc = s.outm.ensure((size_t)P * k * sizeof(int2)))) return rc;
        if ((rc = s.pack.ensure((size_t)P * k * sizeof(ushort2)))) return rc;
        if ((rc = s.offs.ensure((size_t)P * sizeof(uint32_t)))) return rc;
    }
    const uint32_t W = bytes / 4;
    // Nothing else in flight (the first submission of a run, or a caller that submits and waits): the upload is exposed, so it
    // goes in two pieces and the first half of the pairs starts as soon as its frames are there — half the pipeline fill for
    // one extra launch sequence. With other submissions in flight the upload hides behind their kernels and the batch stays whole.
    bool alone = true;
    for (const PairSlot &o : ps->slot) alone = alone && (&o == &s || !o.busy);
    const uint32_t split = (alone && P >= 512) ? P / 2 : 0;   // pairs in the first piece
    auto enqueue = [&]() -> int {
        if (split) {
            const size_t f0 = (size_t)split + 1;   // frames the first piece needs
            VB_CUDA(cudaMemcpyAsync(s.pts.p, pts, f0 * k * 8, cudaMemcpyHostToDevice, ctx->copy_in));
            VB_CUDA(cudaMemcpyAsync(s.desc.p, desc, f0 * k * bytes, cudaMemcpyHostToDevice, ctx->copy_in));
            VB_CUDA(cudaEventRecord(s.ev_in0, ctx->copy_in));
            VB_CUDA(cudaMemcpyAsync(s.pts.as<uint8_t>() + f0 * k * 8, reinterpret_cast<const uint8_t *>(pts) + f0 * k * 8,
                                    pts_bytes - f0 * k * 8, cudaMemcpyHostToDevice, ctx->copy_in));
            VB_CUDA(cudaMemcpyAsync(s.desc.as<uint8_t>() + f0 * k * bytes, desc + f0 * k * bytes, desc_bytes - f0 * k * bytes,
                                    cudaMemcpyHostToDevice, ctx->copy_in));
        } else {
            VB_CUDA(cudaMemcpyAsync(s.pts.p, pts, pts_bytes, cudaMemcpyHostToDevice, ctx->copy_in));
            VB_CUDA(cudaMemcpyAsync(s.desc.p, desc, desc_bytes, cudaMemcpyHostToDevice, ctx->copy_in));
        }
        VB_CUDA(cudaEventRecord(s.ev_in, ctx->copy_in));
        VB_CUDA(cudaStreamWaitEvent(cx->stream, split ? s.ev_in0 : s.ev_in, 0));
        const float2 *p2 = s.pts.as<float2>();
        const uint32_t *d32 = s.desc.as<uint32_t>();
        vb_pair_result *res_d = s.res.as<vb_pair_result>();
        int2 *outm_d = matches16 ? s.outm.as<int2>() : nullptr;
        for (uint32_t b0 = 0, pb = 0; b0 < P; b0 += pb) {
            pb = (P - b0 < PAIRS_MAX_BATCH) ? P - b0 : PAIRS_MAX_BATCH;
            if (split && b0 < split) pb = (split - b0 < pb) ? split - b0 : pb;
            if (split && b0 == split) VB_CUDA(cudaStreamWaitEvent(cx->stream, s.ev_in, 0));
            int r = pairs_core(cx, pb, p2 + (size_t)b0 * k, p2 + (size_t)(b0 + 1) * k, k, d32 + (size_t)b0 * k * W,
                               d32 + (size_t)(b0 + 1) * k * W, (size_t)k * W, k, k, bytes, *params, params->seed0 + b0,
                               res_d + b0, outm_d ? outm_d + (size_t)b0 * k : nullptr);
            if (r) return r;
        }
        if (matches16) {
            k_pack_matches<<<P, 256, 0, cx->stream>>>(res_d, outm_d, k, (uint32_t)base, s.offs.as<uint32_t>(),
                                                      s.pack.as<ushort2>());
            cx->launches++;
            VB_CUDA(cudaGetLastError());
        }
        VB_CUDA(cudaEventRecord(s.ev_done, cx->stream));
        VB_CUDA(cudaStreamWaitEvent(s.s_out, s.ev_done, 0));
        VB_CUDA(cudaMemcpyAsync(results, res_d, (size_t)P * sizeof(vb_pair_result), cudaMemcpyDeviceToHost, s.s_out));
        if (matches16)
            VB_CUDA(cudaMemcpyAsync(match_offsets, s.offs.p, (size_t)P * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.s_out));
        VB_CUDA(cudaEventRecord(s.ev_res, s.s_out));
        return VB_OK;
    };
    rc = enqueue();
    if (rc != VB_OK) {   // the DMA engines may still be reading the caller's buffers: drain before reporting
        drain(ctx, s);
        return rc;
    }
    s.busy = true;
    *ticket = ps->next_ticket++;
    return VB_OK;
}

int pairs_wait(vb_ctx *ctx, int ticket, uint64_t *total_matches) {
    VB_REQUIRE(ctx != nullptr, VB_ERR_INVALID, "ctx is NULL");
    PairsStream *ps = static_cast<PairsStream *>(ctx->pairs_stream);
    VB_REQUIRE(ps != nullptr && ticket >= 0, VB_ERR_INVALID, "unknown ticket");
    PairSlot &s = ps->slot[ticket % PAIRS_DEPTH];
    VB_REQUIRE(s.busy && s.ticket == ticket, VB_ERR_INVALID, "unknown or already completed ticket");
    VB_CUDA(cudaSetDevice(ctx->device));
    s.busy = false;
    uint64_t total = 0;
    if (s.P && s.device_io) {
        cudaError_t e = cudaEventSynchronize(s.ev_done);
        if (e != cudaSuccess) {
            drain(ctx, s);
            set_error("vb_pairs_wait: %s", cudaGetErrorString(e));
            return VB_ERR_CUDA;
        }
    } else if (s.P) {
        cudaError_t e = cudaEventSynchronize(s.ev_res);
        if (e != cudaSuccess) {
            drain(ctx, s);
            set_error("vb_pairs_wait: %s", cudaGetErrorString(e));
            return VB_ERR_CUDA;
        }
        if (s.matches16) {
            total = (uint64_t)s.offsets[s.P - 1] - s.base + (uint64_t)(uint32_t)s.results[s.P - 1].n_matches;
            if (total_matches) *total_matches = total;
            if (total > s.cap) {
                set_error("vb_pairs_wait: %llu matches, room for %llu", (unsigned long long)total, (unsigned long long)s.cap);
                return VB_ERR_CAPACITY;
            }
            if (total) {
                VB_CUDA(cudaMemcpyAsync(s.matches16 + 2 * s.base, s.pack.p, total * sizeof(ushort2), cudaMemcpyDeviceToHost, s.s_out));
                VB_CUDA(cudaEventRecord(s.ev_m, s.s_out));
                VB_CUDA(cudaEventSynchronize(s.ev_m));
            }
        }
    }
    if (total_matches) *total_matches = total;
    return VB_OK;
}

// Device-resident twin of pairs_submit: inputs and outputs are device pointers, nothing is copied; the ticket completes when
// the kernels have. Work of an odd ticket runs on the twin's stream after everything queued on the context's stream at
// submission time (the inputs may have been produced there).
int pairs_submit_d(vb_ctx *ctx, const float *pts_d, const uint8_t *desc_d, uint32_t nframes, uint32_t k, uint32_t bytes,
                   const vb_pair_params *params, vb_pair_result *results_d, int32_t *out_matches_d, int *ticket) {
    VB_REQUIRE(ctx && pts_d && desc_d && results_d && ticket, VB_ERR_INVALID, "NULL argument");
    int rc;
    if ((rc = pairs_check_params(params, bytes, k))) return rc;
    VB_CUDA(cudaSetDevice(ctx->device));
    PairsStream *ps = stream_state(ctx);
    VB_REQUIRE(ps != nullptr, VB_ERR_CUDA, "out of host memory");
    PairSlot &s = ps->slot[ps->next_ticket % PAIRS_DEPTH];
    VB_REQUIRE(!s.busy, VB_ERR_CAPACITY, "three submissions are already in flight: vb_pairs_wait the oldest ticket first");
    const uint32_t P = nframes < 2 ? 0 : nframes - 1;
    s.P = P; s.base = 0; s.cap = 0; s.results = nullptr; s.offsets = nullptr; s.matches16 = nullptr;
    s.ticket = ps->next_ticket;
    s.device_io = true;
    if ((rc = compute_ctx(ctx, s.ticket, &s.cx))) return rc;
    vb_ctx *cx = s.cx;
    if (P) {
        if (!s.s_out) {
            VB_CUDA(cudaStreamCreateWithFlags(&s.s_out, cudaStreamNonBlocking));
            for (cudaEvent_t *e : {&s.ev_in, &s.ev_in0, &s.ev_done, &s.ev_res, &s.ev_m}) VB_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        }
        if (cx != ctx) {
            VB_CUDA(cudaEventRecord(s.ev_in, ctx->stream));
            VB_CUDA(cudaStreamWaitEvent(cx->stream, s.ev_in, 0));
        }
        const uint32_t W = bytes / 4;
        const float2 *p2 = reinterpret_cast<const float2 *>(pts_d);
        const uint32_t *d32 = reinterpret_cast<const uint32_t *>(desc_d);
        int2 *outm = reinterpret_cast<int2 *>(out_matches_d);
        for (uint32_t b0 = 0; b0 < P; b0 += PAIRS_MAX_BATCH) {
            const uint32_t pb = (P - b0 < PAIRS_MAX_BATCH) ? P - b0 : PAIRS_MAX_BATCH;
            rc = pairs_core(cx, pb, p2 + (size_t)b0 * k, p2 + (size_t)(b0 + 1) * k, k, d32 + (size_t)b0 * k * W,
                            d32 + (size_t)(b0 + 1) * k * W, (size_t)k * W, k, k, bytes, *params, params->seed0 + b0, results_d + b0,
                            outm ? outm + (size_t)b0 * k : nullptr);
            if (rc) {
                cudaStreamSynchronize(cx->stream);
                return rc;
            }
        }
        VB_CUDA(cudaEventRecord(s.ev_done, cx->stream));
    }
    s.busy = true;
    *ticket = ps->next_ticket++;
    return VB_OK;
}

}  // namespace vb

using namespace vb;

extern "C" {

int vb_pairs_submit_d(vb_ctx *ctx, const float *pts_d, const uint8_t *desc_d, uint32_t nframes, uint32_t k, uint32_t bytes,
                      const vb_pair_params *params, vb_pair_result *results_d, int32_t *out_matches_d, int *ticket) {
    return pairs_submit_d(ctx, pts_d, desc_d, nframes, k, bytes, params, results_d, out_matches_d, ticket);
}

int vb_pairs_submit(vb_ctx *ctx, const float *pts, const uint8_t *desc, uint32_t nframes, uint32_t k, uint32_t bytes,
                    const vb_pair_params *params, vb_pair_result *results, uint32_t *match_offsets, uint16_t *matches16,
                    uint64_t cap_matches, int *ticket) {
    return pairs_submit(ctx, pts, desc, nframes, k, bytes, params, results, match_offsets, matches16, cap_matches, 0, ticket);
}

int vb_pairs_wait(vb_ctx *ctx, int ticket, uint64_t *total_matches) { return pairs_wait(ctx, ticket, total_matches); }

int vb_pairs_run_compact(vb_ctx *ctx, const float *pts, const uint8_t *desc, uint32_t nframes, uint32_t k, uint32_t bytes,
                         const vb_pair_params *params, vb_pair_result *results, uint32_t *match_offsets, uint16_t *matches16,
                         uint64_t cap_matches, uint64_t *total_matches) {
    int ticket = -1;
    int rc = vb_pairs_submit(ctx, pts, desc, nframes, k, bytes, params, results, match_offsets, matches16, cap_matches, &ticket);
    if (rc) return rc;
    return vb_pairs_wait(ctx, ticket, total_matches);
}

/* Pinned host memory for the caller's frame and result buffers (asynchronous copies only overlap with pinned memory). */
int vb_host_alloc(size_t bytes, void **out) {
    VB_REQUIRE(out != nullptr, VB_ERR_INVALID, "out is NULL");
    VB_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return VB_OK;
}
int vb_host_free(void *p) {
    if (p) VB_CUDA(cudaFreeHost(p));
    return VB_OK;
}
int vb_host_register(void *p, size_t bytes) {
    VB_REQUIRE(p != nullptr && bytes > 0, VB_ERR_INVALID, "NULL or empty range");
    VB_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return VB_OK;
}
int vb_host_unregister(void *p) {
    VB_REQUIRE(p != nullptr, VB_ERR_INVALID, "NULL pointer");
    VB_CUDA(cudaHostUnregister(p));
    return VB_OK;
}

}  // extern "C"
