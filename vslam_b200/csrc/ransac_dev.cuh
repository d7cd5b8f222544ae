// ransac_dev.cuh — constants and host-side plan shared by ransac.cu and pairs.cu.
#pragma once
#include "common.cuh"

namespace vb {

// Score reduction order (must equal the oracle's VBO_SUM_CHUNK / VBO_SUM_GROUP definition):
// level 1 = 128 consecutive matches summed sequentially in fp64, level 2 = 64 consecutive level-1
// sums added in order, level 3 = level-2 sums added in order, then one rounding to fp32.
constexpr int SUM_CHUNK = 128;
constexpr int SUM_GROUP = 64;
constexpr int SCORE_THREADS = 128;   // == SUM_CHUNK: one tile entry staged per thread
constexpr int SELECT_THREADS = 256;

// Number of matches / seed of problem p: from a device array (pair pipeline) or a launch constant.
struct ProblemDims {
    const uint32_t *m_arr;  // [P] or nullptr
    uint32_t m_fixed;
    uint32_t seed0;         // problem p samples with std::mt19937(seed0 + p)
    __device__ __forceinline__ uint32_t m(uint32_t p) const { return m_arr ? m_arr[p] : m_fixed; }
};

struct RansacPlan {
    uint32_t P, mcap, H;
    int min_items;
    int hpt;              // hypotheses per thread in k_score
    uint32_t htiles;      // grid.x
    uint32_t grid_y;      // chunk ranges
    uint32_t chunks_per_cta;
    int unit_is_group;    // partials are per level-2 group (1) or per level-1 chunk (0)
    uint32_t nunits;      // partial slots per (problem, hypothesis)
    uint32_t nraw;        // mt19937 words generated per problem
};

struct RansacSelectArgs {
    const float4 *corr;
    ProblemDims dims;
    uint32_t mcap;
    const float *F_all;
    uint32_t H;
    float thr;
    const int32_t *part_cnt;
    const double *part_sum;
    uint32_t nunits;
    int unit_is_group;
    int32_t *cnt;     // [P][H]
    float *score;     // [P][H]
    const int32_t *status;  // [P] or nullptr (all OK)
    vb_pair_result *results;
    uint8_t *mask;          // [P][mcap] or nullptr
    const int2 *tent;       // [P][mcap] tentative matches (pair pipeline) or nullptr
    int2 *out_matches;      // [P][mcap] or nullptr
    int score_only;
    int prefolded;          // cnt/score already hold the folded totals (k_fold ran first)
    int sampson;            // opt-in mode: the winner's mask uses the true Sampson distance (VB_RANSAC_SAMPSON)
    int lazy;               // counts came from k_count: no residual sums; k_select scores the tied hypotheses itself
    uint32_t *tied;         // [P][H] scratch list of tied hypotheses (lazy mode)
    const unsigned int *queue_timeouts;   // bounded counting: non-zero if a work-queue wait gave up (counts incomplete), or nullptr
};

int ransac_plan(vb_ctx *ctx, uint32_t P, uint32_t mcap, uint32_t m_upper, uint32_t H, int min_items, RansacPlan *pl);
int ransac_launch_score(vb_ctx *ctx, const RansacPlan &pl, const float4 *corr, ProblemDims dims, const float *F_all,
                        float thr);
// lazy = true: count with k_count (approximate residual + exact fallback) and compute residual sums only for the
// hypotheses that tie at the largest count; per-hypothesis scores of the others are then not available (0).
int ransac_run(vb_ctx *ctx, const RansacPlan &pl, const float4 *corr, ProblemDims dims, float thr,
               vb_pair_result *results_d, uint8_t *mask_d, const int2 *tent_d, int2 *out_matches_d, bool lazy,
               uint32_t flags = 0);

}  // namespace vb
