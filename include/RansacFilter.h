// RansacFilter.h — drop-in replacement for the reference's include/RansacFilter.h, backed by
// libvslam_b200.so.
//
// Same class, same public const members, same method signatures (reference include/RansacFilter.h:9-25),
// so `RansacFilter rf(8, 100, 10)` (src/vslam.cpp:19) and `rf.find_fundamental(frame1.points,
// frame2.points, i_matches, inliers, F)` (src/Frame.cpp:97) compile and behave as before:
//
//   * find_fundamental writes a 3x3 CV_32F `fundamental` and an `inliers` vector of matches.size() flags
//     when a hypothesis is accepted; when none is (or there are fewer than min_items matches, which is
//     undefined behaviour in the reference) both outputs are left exactly as they were passed in.
//   * Sampling reproduces std::mt19937 + std::uniform_int_distribution draw for draw. The reference
//     seeds from std::random_device on every call (src/RansacFilter.cpp:15-16); that stays the default.
//     For reproducible runs set the seed out of band: vslam_b200_set_ransac_seed(seed) or the
//     environment variable VSLAM_RANSAC_SEED (call n uses seed + n).
//   * The 8-point solve is a fully specified fp64 sequence instead of cv::SVDecomp (whose result depends
//     on the OpenCV build); see DESIGN.md "8-point solve".
#ifndef VSLAM_B200_RANSAC_FILTER_H
#define VSLAM_B200_RANSAC_FILTER_H

#include <cstdlib>
#include <opencv2/core.hpp>
#include <random>
#include <utility>
#include <vector>

class RansacFilter {
   public:
    const int min_items;
    const int max_iterations;
    const float threshold;

    RansacFilter(const int min_items = 8, const int max_iterations = 100, const float threshold = 0.2);

    void initialize_sets(const int n_matches);
    void find_fundamental(const std::vector<cv::Point2f> &p1, const std::vector<cv::Point2f> &p2,
                          const std::vector<std::pair<int, int> > &matches, std::vector<bool> &inliers,
                          cv::Mat &fundamental);
    // p1_set / p2_set must hold exactly 8 points (the only size find_fundamental ever passes).
    void compute_fundamental(const std::vector<cv::Point2f> &p1_set, const std::vector<cv::Point2f> &p2_set,
                             cv::Mat &temp_F);
    std::pair<int, float> compute_fundamental_residual(const std::vector<cv::Point2f> &p1,
                                                       const std::vector<cv::Point2f> &p2,
                                                       const std::vector<std::pair<int, int> > &matches,
                                                       const cv::Mat &F, std::vector<bool> &inliers);

    // Read-only view of the last sample sets (the reference keeps them private; exposed for tests).
    const std::vector<std::vector<int> > &sample_sets() const { return ransac_sets; }

   private:
    std::vector<std::vector<int> > ransac_sets;
    unsigned next_seed();
};

// Deterministic sampling: call n (0-based, process-wide) uses seed + n. vslam_b200_clear_ransac_seed()
// returns to one fresh std::random_device value per call.
void vslam_b200_set_ransac_seed(unsigned long long seed);
void vslam_b200_clear_ransac_seed();
// Opt-in, NOT reference behaviour: VB_RANSAC_HARTLEY (1) normalises every 8-point sample (the `//TODO: normalize` at reference
// src/RansacFilter.cpp:40), VB_RANSAC_SAMPSON (2) replaces the mis-parenthesised residual of :125-126 by the true Sampson
// distance (threshold then in squared pixels). 0 (default) is the reference's behaviour.
void vslam_b200_set_ransac_flags(unsigned flags);

#endif
