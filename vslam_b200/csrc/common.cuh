// common.cuh — context, error plumbing and small device helpers shared by the kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/vslam_b200.h"

namespace vb {

void set_error(const char *fmt, ...);

#define VB_CUDA(call)                                                                             \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            vb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return VB_ERR_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

#define VB_REQUIRE(cond, code, msg)                        \
    do {                                                   \
        if (!(cond)) {                                     \
            vb::set_error("%s: %s", __func__, msg);        \
            return code;                                   \
        }                                                  \
    } while (0)

// Grow-only device buffer.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return VB_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e));
            return VB_ERR_CUDA;
        }
        cap = want;
        return VB_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Grow-only pinned host buffer (staging for the host-pointer entry points).
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return VB_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) {
            set_error("cudaMallocHost(%zu) -> %s", want, cudaGetErrorString(e));
            return VB_ERR_CUDA;
        }
        cap = want;
        return VB_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

enum WsSlot {
    WS_P1 = 0, WS_P2, WS_MATCHES, WS_CORR, WS_M, WS_SETS, WS_RAW, WS_FALL, WS_PART_CNT, WS_PART_SUM, WS_CNT, WS_SCORE,
    WS_RESULT, WS_MASK, WS_OUTMATCH, WS_D1, WS_D2, WS_KNN_PART, WS_KNN, WS_TENT, WS_FLAGS, WS_Q, WS_OUT0, WS_OUT1,
    WS_OUT2, WS_OFFS, WS_SCAN, WS_PTS, WS_DESC, WS_SEEDS, WS_MISC, WS_L2A, WS_L2B, WS_L2C, WS_L2N, WS_L2M, WS_L2H, WS_EXP, WS_SBP_X, WS_SBP_Q, WS_SBP_DESC, WS_SBP_IDS, WS_SBP_OBSOFF,
    WS_SBP_OBS, WS_SBP_ACC, WS_SBP_CUR, WS_SBP_OWNER, WS_SBP_ASSIGN, WS_BOUNDS, WS_TIED, WS_PRUNE, WS_PRUNE_STATS, WS_ALIVE, WS_BQ_ITEMS, WS_BQ_VALID, WS_BQ_CTL, WS_BQ_ORDER, WS_FOLD_CNT, WS_FOLD_SUM, WS_SCAN2, WS_KD_TOP, WS_COUNT
};

struct ProfEntry {
    cudaEvent_t a = nullptr, b = nullptr;
    bool used = false;
};

}  // namespace vb

struct vb_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;   // upload / download streams of vb_pairs_run
    vb_ctx *twin = nullptr;   // second stream + second set of workspaces: vb_pairs_run alternates sub-batches between the two
    std::vector<cudaEvent_t> events;                      // untimed events for the copy/compute pipeline
    void *pairs_stream = nullptr;                         // vb_pairs_submit / vb_pairs_wait slots (stream.cu)
    // Path selectors for tests and measurements (vb_set_option). Nothing in the library reads the environment; a twin
    // context looks its parent's options up.
    std::map<std::string, long long> opts;
    const vb_ctx *opt_parent = nullptr;
    long long opt(const char *name, long long dflt) const {
        const vb_ctx *c = opt_parent ? opt_parent : this;
        auto it = c->opts.find(name);
        return it == c->opts.end() ? dflt : it->second;
    }
    vb::DevBuf ws[vb::WS_COUNT];
    vb::PinBuf pin[4];
    uint64_t launches = 0;
    std::vector<uint4> kd_segs_host;   // segment table of the last top-down k-d tree build (source of an asynchronous upload)
    uint64_t sbp_cap_hint = 0;     // largest candidate count a search-by-projection call has needed so far
    uint32_t func_attr_done = 0;   // per-context (hence per-device) bits: large-smem attribute set for kernel i
    bool profile = false;
    std::map<std::string, vb::ProfEntry> prof;

    int ws_ensure(int slot, size_t bytes) { return ws[slot].ensure(bytes); }
    // profiling brackets (only when enabled; events on the launching stream)
    void prof_begin(const char *name);
    void prof_end(const char *name);
};

namespace vb { void pairs_stream_release(vb_ctx *ctx); }

struct vb_tree {
    vb_ctx *ctx = nullptr;
    uint32_t n = 0;
    uint32_t height = 0;
    float *x = nullptr;      // [n] pre-order
    float *y = nullptr;      // [n]
    uint32_t *idx = nullptr; // [n] original index
    void *block = nullptr;   // one allocation backing the three arrays
};

namespace vb {

// ---- device helpers --------------------------------------------------------------------------
__device__ __forceinline__ int warp_reduce_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

inline unsigned div_up(unsigned a, unsigned b) { return (a + b - 1) / b; }
inline size_t div_up64(size_t a, size_t b) { return (a + b - 1) / b; }

}  // namespace vb
