// KDTree.h — drop-in replacement for the reference's include/KDTree.h, backed by libvslam_b200.so.
//
// The two tree structs and the free functions Frame.cpp / vslam.cpp / tests/test_kdtree.cpp use keep
// their names, signatures and observable behaviour (reference include/KDTree.h:13-23,25,30,44,47-57,
// 60,79; src/KDTree.cpp:25-35,37-43,73-78,107-121,145-150):
//
//   * construct_kdtree still hands back `root` = ONE malloc() block of N pointer nodes in DFS pre-order
//     with valid left/right links, `size` advanced by N and `height` = floor(log2 N)+1, because callers
//     walk the nodes (tests/test_kdtree.cpp:16-37) and release them with free() (src/vslam.cpp:295-297).
//     The node contents come from the GPU build (vb_kdtree_build + vb_kdtree_export).
//   * nearest / radius_search run on the GPU copy of the tree. The structs are PODs that get copied,
//     moved inside std::vector<Frame> and freed behind our back, so the device handle cannot live in
//     them: it lives in a process-wide side table keyed by `root` (LRU-bounded; a tree that is not in
//     the table — evicted, or built by other code — is re-imported from its host node array on demand).
//
// Additions (not in the reference): *_batch overloads that answer many queries in one launch — that is
// where the GPU pays off; a single query per call is dominated by launch latency.
//
// The reference also declares nearest_approx / nearest(frame_kdtree) / k_nearest (include/KDTree.h:34-42,
// 65-77) but never defines them, so no caller can link against them; they are not declared here.
#ifndef VSLAM_B200_KDTREE_H
#define VSLAM_B200_KDTREE_H

#include <algorithm>
#include <cmath>
#include <opencv2/core.hpp>
#include <vector>

#include <vslam_internal.h>

#ifndef SQ
#define SQ(x) ((x) * (x))
#endif
#ifndef ABS
#define ABS(x) (((x) > 0) ? x : -x)
#endif
#ifndef P
#define P(pt, i) ((float *)&(pt))[i]
#endif

struct KDTree {
    struct KDTreeNode {
        cv::Point2f pt;
        KDTreeNode *left;
        KDTreeNode *right;
    };

    KDTreeNode *root;
    u32 size = 0;
    u8 height = 0;
};

struct frame_kdtree {
    struct KDTreeNode {
        usize pt_index;
        KDTreeNode *left;
        KDTreeNode *right;
    };

    KDTreeNode *root;
    u32 size = 0;
    u8 height = 0;
};

// ---- value tree ------------------------------------------------------------------------------
void construct_kdtree(KDTree &kdtree, const std::vector<cv::Point2f> &points);
cv::Point2f nearest(const KDTree &kdtree, const cv::Point2f &query_pt, float max_distance_sq = INFINITY);
std::vector<cv::Point2f> radius_search(const KDTree &kdtree, const cv::Point2f &query_pt, float radius);

// ---- index tree (the one Frame carries, include/Frame.h:24) -------------------------------------
void construct_kdtree(frame_kdtree &kdtree, const std::vector<cv::Point2f> &points);
std::vector<usize> radius_search(const frame_kdtree kdtree, const std::vector<cv::Point2f> &points,
                                 const cv::Point2f &query_pt, float radius);

// ---- batched additions --------------------------------------------------------------------------
// One launch for all queries; results[i] equals what the single-query call returns for queries[i].
std::vector<cv::Point2f> nearest_batch(const KDTree &kdtree, const std::vector<cv::Point2f> &queries,
                                       float max_distance_sq = INFINITY);
std::vector<std::vector<cv::Point2f> > radius_search_batch(const KDTree &kdtree, const std::vector<cv::Point2f> &queries,
                                                           float radius);
std::vector<std::vector<usize> > radius_search_batch(const frame_kdtree &kdtree, const std::vector<cv::Point2f> &points,
                                                     const std::vector<cv::Point2f> &queries, float radius);

// Drop the device copy that belongs to `root` now (optional: free(root) alone is still correct, the
// side table is LRU-bounded). Call before free(root) if device memory matters.
void vslam_b200_kdtree_release(const void *root);
// GPU used by the adapters (default: $VSLAM_B200_DEVICE or 0). Must be called before the first use.
void vslam_b200_set_device(int device);

#endif
