// hamming_dev.cuh — plan / argument structs shared by hamming.cu and pairs.cu.
#pragma once
#include "common.cuh"

namespace vb {

// (distance, train index) packed as distance << 22 | index: unsigned order == (distance, index) order.
constexpr uint32_t KNN_IDX_BITS = 22;
constexpr uint32_t KNN_IDX_MASK = (1u << KNN_IDX_BITS) - 1u;

struct HammingPlan {
    uint32_t P, n1, n2, W;
    int qpt;            // query descriptors per thread
    uint32_t qtiles;    // grid.x
    uint32_t nsplits;   // grid.y: contiguous train ranges
    uint32_t split_len;
};

struct KnnFinishArgs {
    const uint2 *part;   // [P][nsplits][n1] (filled in by hamming_launch)
    uint32_t nsplits, n1;
    double ratio;
    int32_t *knn_idx;    // [P][n1][2] or nullptr
    int32_t *knn_dist;   // [P][n1][2]
    int2 *tent;          // [P][mcap] ratio-test survivors in query order, or nullptr
    uint32_t mcap;
    uint32_t *m_out;     // [P] survivor count
    // pair pipeline: also emit float4 correspondences (p1[q], p2[train])
    float4 *corr;        // [P][mcap] or nullptr
    const float2 *p1_base, *p2_base;
    size_t pts_stride;   // points between consecutive problems
};

int hamming_plan(vb_ctx *ctx, uint32_t P, uint32_t n1, uint32_t n2, uint32_t bytes, HammingPlan *pl);
// tensor-core path (hamming_tc.cu): 256-bit descriptors; *final_part = [P][n1] keys (one split) inside WS_KNN_PART
bool hamming_tc_eligible(const vb_ctx *ctx, const HammingPlan &pl);
// need_second_index = false: the second neighbour's key carries its exact distance but only its group's first column, and
// queries that are certain to fail Lowe's test with `ratio` (>= 0) keep their group keys unevaluated
int hamming_tc_launch(vb_ctx *ctx, const HammingPlan &pl, const uint32_t *d1, const uint32_t *d2, size_t stride_words,
                      const uint2 **final_part, bool need_second_index, double ratio);
int hamming_launch(vb_ctx *ctx, const HammingPlan &pl, const uint32_t *d1, const uint32_t *d2, size_t stride_words,
                   KnnFinishArgs fin);

}  // namespace vb
