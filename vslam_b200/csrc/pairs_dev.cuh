// pairs_dev.cuh — the batched pair pipeline shared by pairs.cu (blocking entry points) and stream.cu (asynchronous,
// compact-output entry points).
#pragma once
#include "common.cuh"

namespace vb {

constexpr uint32_t PAIRS_MAX_BATCH = 1024;   // pairs per launch sequence (workspaces are sized for this)
constexpr int PAIRS_DEPTH = 3;               // submissions in flight per context: one uploading, two computing (ctx + twin)

// match_features (reference src/Frame.cpp:82-105) for P pairs whose inputs are already on the device. Pair i reads
// p1_base + i * pts_stride / d1_base + i * desc_stride_words as frame 1 and the *2* bases as frame 2, samples with
// std::mt19937(seed0 + i), and writes results_d[i] and (optionally) its inlier matches to out_matches_d[i * n1 ...].
int pairs_core(vb_ctx *ctx, uint32_t P, const float2 *p1_base, const float2 *p2_base, size_t pts_stride,
               const uint32_t *d1_base, const uint32_t *d2_base, size_t desc_stride_words, uint32_t n1, uint32_t n2,
               uint32_t bytes, const vb_pair_params &prm, uint32_t seed0, vb_pair_result *results_d, int2 *out_matches_d);
int pairs_check_params(const vb_pair_params *p, uint32_t bytes, uint32_t n2);

}  // namespace vb
