// RansacFilter.cpp — implementation of include/RansacFilter.h over the C ABI. Compiled by the integrator
// in place of the reference's src/RansacFilter.cpp and linked with libvslam_b200.so.
#include "RansacFilter.h"

#include <atomic>
#include <stdexcept>

#include "adapter_common.h"

using namespace vslam_b200_adapter;

namespace {
std::atomic<bool> g_seed_fixed(false);
std::atomic<unsigned long long> g_seed_base(0), g_seed_calls(0);

std::vector<float> flatten(const std::vector<cv::Point2f> &pts) {
    std::vector<float> f(pts.size() * 2);
    for (size_t i = 0; i < pts.size(); i++) { f[2 * i] = pts[i].x; f[2 * i + 1] = pts[i].y; }
    return f;
}
std::vector<int32_t> flatten(const std::vector<std::pair<int, int> > &m) {
    std::vector<int32_t> f(m.size() * 2);
    for (size_t i = 0; i < m.size(); i++) { f[2 * i] = m[i].first; f[2 * i + 1] = m[i].second; }
    return f;
}
}  // namespace

void vslam_b200_set_ransac_seed(unsigned long long seed) {
    g_seed_base = seed;
    g_seed_calls = 0;
    g_seed_fixed = true;
}
void vslam_b200_clear_ransac_seed() { g_seed_fixed = false; }

// Opt-in repairs of the reference's own TODOs (VB_RANSAC_HARTLEY | VB_RANSAC_SAMPSON, include/vslam_b200.h); 0 = the
// reference's behaviour, which is the default.
static unsigned g_ransac_flags = 0;
void vslam_b200_set_ransac_flags(unsigned flags) { g_ransac_flags = flags; }

RansacFilter::RansacFilter(const int min_items, const int max_iterations, const float threshold)
    : min_items(min_items), max_iterations(max_iterations), threshold(threshold) {}

// What `std::random_device rd; std::mt19937 gen(rd());` (reference src/RansacFilter.cpp:15-16) is seeded
// with: a fresh OS value by default, or the out-of-band deterministic sequence.
unsigned RansacFilter::next_seed() {
    if (!g_seed_fixed) {
        if (const char *e = std::getenv("VSLAM_RANSAC_SEED")) vslam_b200_set_ransac_seed(std::strtoull(e, NULL, 10));
    }
    if (g_seed_fixed) return (unsigned)(g_seed_base + g_seed_calls++);
    std::random_device rd;
    return rd();
}

void RansacFilter::initialize_sets(const int n_matches) {
    ransac_sets = std::vector<std::vector<int> >(max_iterations, std::vector<int>(8, 0));
    if (max_iterations <= 0) return;
    std::vector<int32_t> flat((size_t)max_iterations * 8);
    check(vb_ransac_sample_sets(context(), (uint32_t)n_matches, min_items, (uint32_t)max_iterations, next_seed(), flat.data()),
          "vb_ransac_sample_sets");
    for (int i = 0; i < max_iterations; i++)
        for (int j = 0; j < 8; j++) ransac_sets[i][j] = flat[(size_t)i * 8 + j];
}

void RansacFilter::find_fundamental(const std::vector<cv::Point2f> &p1, const std::vector<cv::Point2f> &p2,
                                    const std::vector<std::pair<int, int> > &matches, std::vector<bool> &inliers,
                                    cv::Mat &fundamental) {
    const std::vector<float> a = flatten(p1), b = flatten(p2);
    const std::vector<int32_t> mm = flatten(matches);
    float F[9];
    std::vector<uint8_t> mask(matches.size() ? matches.size() : 1);
    int32_t n_in = 0, best = -1;
    float score = 0.f;
    const int rc = vb_ransac_fundamental_ex(context(), a.data(), (uint32_t)p1.size(), b.data(), (uint32_t)p2.size(), mm.data(),
                                            (uint32_t)matches.size(), min_items, (uint32_t)(max_iterations > 0 ? max_iterations : 0),
                                            threshold, next_seed(), g_ransac_flags, F, mask.data(), &n_in, &score, &best);
    // No accepted hypothesis: the reference leaves `fundamental` empty and `inliers` as passed (:59-65).
    // Fewer matches than min_items is undefined behaviour there; here it is the same quiet no-op.
    if (rc == VB_ERR_NO_MODEL || rc == VB_ERR_TOO_FEW) return;
    check(rc, "vb_ransac_fundamental");
    fundamental = cv::Mat(3, 3, CV_32FC1);
    for (int i = 0; i < 9; i++) fundamental.at<float>(i / 3, i % 3) = F[i];
    std::vector<bool> cur(matches.size());
    for (size_t i = 0; i < matches.size(); i++) cur[i] = mask[i] != 0;
    inliers.swap(cur);
}

void RansacFilter::compute_fundamental(const std::vector<cv::Point2f> &p1_set, const std::vector<cv::Point2f> &p2_set,
                                       cv::Mat &temp_F) {
    if (p1_set.size() != 8 || p2_set.size() != 8)
        throw std::invalid_argument("vslam_b200: compute_fundamental takes exactly 8 correspondences");
    const std::vector<float> a = flatten(p1_set), b = flatten(p2_set);
    float F[9];
    check(vb_ransac_solve8(context(), a.data(), b.data(), 1, F), "vb_ransac_solve8");
    temp_F = cv::Mat(3, 3, CV_32FC1);
    for (int i = 0; i < 9; i++) temp_F.at<float>(i / 3, i % 3) = F[i];
}

std::pair<int, float> RansacFilter::compute_fundamental_residual(const std::vector<cv::Point2f> &p1,
                                                                 const std::vector<cv::Point2f> &p2,
                                                                 const std::vector<std::pair<int, int> > &matches,
                                                                 const cv::Mat &F, std::vector<bool> &inliers) {
    const std::vector<float> a = flatten(p1), b = flatten(p2);
    const std::vector<int32_t> mm = flatten(matches);
    float Ff[9];
    for (int i = 0; i < 9; i++) Ff[i] = F.at<float>(i / 3, i % 3);
    std::vector<uint8_t> mask(matches.size() ? matches.size() : 1);
    int32_t n_in = 0;
    float score = 0.f;
    check(vb_ransac_residual(context(), a.data(), (uint32_t)p1.size(), b.data(), (uint32_t)p2.size(), mm.data(),
                             (uint32_t)matches.size(), Ff, threshold, mask.data(), &n_in, &score),
          "vb_ransac_residual");
    inliers.resize(matches.size());
    for (size_t i = 0; i < matches.size(); i++) inliers[i] = mask[i] != 0;
    return std::make_pair((int)n_in, score);
}
