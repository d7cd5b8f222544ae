// ref_wrap.cpp — C entry points around the REFERENCE'S OWN sources, for tests only.
//
// TEST INFRASTRUCTURE, NOT PRODUCT. Built by oracle/Makefile into oracle/_ref/libvbref.so together
// with /root/reference/src/KDTree.cpp and /root/reference/src/RansacFilter.cpp, which are compiled
// unmodified from where they lie (never copied into this repo). OpenCV is replaced by the test-only
// stand-in tests/cvlite/opencv2/core.hpp; see that file for what it defines and how each rule was
// checked against cv2.
#define private public  // reach RansacFilter::ransac_sets (include/RansacFilter.h:24) to compare sample sets
#include "RansacFilter.h"
#undef private
#include "KDTree.h"

#include <chrono>
#include <cstdint>

extern "C" {
void vbo_ref_seed_set(unsigned seed);

// ---- value tree (include/KDTree.h:13-45) ------------------------------------------------------
// pre_pts: n x 2 floats, the node array root[0..n) which the reference fills in DFS pre-order.
int vbref_kdtree_build(const float *pts, int n, float *pre_pts, int *height, int *links_ok) {
    std::vector<cv::Point2f> v(n);
    for (int i = 0; i < n; i++) v[i] = cv::Point2f(pts[2 * i], pts[2 * i + 1]);
    KDTree t;
    construct_kdtree(t, v);
    if (n == 0) { *height = t.height; *links_ok = (t.root == NULL); return 0; }
    int ok = 1;
    // check the layout claim the GPU build relies on: left == this+1, right == this+1+len/2
    struct Item { KDTree::KDTreeNode *nd; int len; };
    std::vector<Item> st; st.push_back({t.root, n});
    while (!st.empty()) {
        Item it = st.back(); st.pop_back();
        int ll = it.len / 2, rl = it.len - ll - 1;
        if (ll > 0) { if (it.nd->left != it.nd + 1) ok = 0; else st.push_back({it.nd->left, ll}); } else if (it.nd->left) ok = 0;
        if (rl > 0) { if (it.nd->right != it.nd + 1 + ll) ok = 0; else st.push_back({it.nd->right, rl}); } else if (it.nd->right) ok = 0;
    }
    for (int i = 0; i < n; i++) { pre_pts[2 * i] = t.root[i].pt.x; pre_pts[2 * i + 1] = t.root[i].pt.y; }
    *height = t.height;
    *links_ok = ok && (int)t.size == n;
    free(t.root);
    return 0;
}

int vbref_kdtree_nearest(const float *pts, int n, const float *q, int nq, float max_d2, float *out_pts) {
    std::vector<cv::Point2f> v(n);
    for (int i = 0; i < n; i++) v[i] = cv::Point2f(pts[2 * i], pts[2 * i + 1]);
    KDTree t;
    construct_kdtree(t, v);
    for (int i = 0; i < nq; i++) {
        cv::Point2f r = nearest(t, cv::Point2f(q[2 * i], q[2 * i + 1]), max_d2);
        out_pts[2 * i] = r.x; out_pts[2 * i + 1] = r.y;
    }
    free(t.root);
    return 0;
}

// CSR of points; returns total (entries beyond cap are counted, not written)
long vbref_kdtree_radius(const float *pts, int n, const float *q, int nq, float radius, int *offsets,
                         float *out_pts, long cap) {
    std::vector<cv::Point2f> v(n);
    for (int i = 0; i < n; i++) v[i] = cv::Point2f(pts[2 * i], pts[2 * i + 1]);
    KDTree t;
    construct_kdtree(t, v);
    long tot = 0;
    for (int i = 0; i < nq; i++) {
        offsets[i] = (int)tot;
        std::vector<cv::Point2f> r = radius_search(t, cv::Point2f(q[2 * i], q[2 * i + 1]), radius);
        for (size_t k = 0; k < r.size(); k++, tot++)
            if (tot < cap) { out_pts[2 * tot] = r[k].x; out_pts[2 * tot + 1] = r[k].y; }
    }
    offsets[nq] = (int)tot;
    free(t.root);
    return tot;
}

// ---- index tree (include/KDTree.h:47-80) ------------------------------------------------------
int vbref_frame_kdtree_build(const float *pts, int n, int64_t *pre_idx, int *height) {
    std::vector<cv::Point2f> v(n);
    for (int i = 0; i < n; i++) v[i] = cv::Point2f(pts[2 * i], pts[2 * i + 1]);
    frame_kdtree t;
    construct_kdtree(t, v);
    for (int i = 0; i < n; i++) pre_idx[i] = (int64_t)t.root[i].pt_index;
    *height = t.height;
    if (n) free(t.root);
    return 0;
}

long vbref_frame_kdtree_radius(const float *pts, int n, const float *q, int nq, float radius, int *offsets,
                               int64_t *out_idx, long cap) {
    std::vector<cv::Point2f> v(n);
    for (int i = 0; i < n; i++) v[i] = cv::Point2f(pts[2 * i], pts[2 * i + 1]);
    frame_kdtree t;
    construct_kdtree(t, v);
    long tot = 0;
    for (int i = 0; i < nq; i++) {
        offsets[i] = (int)tot;
        std::vector<usize> r = radius_search(t, v, cv::Point2f(q[2 * i], q[2 * i + 1]), radius);
        for (size_t k = 0; k < r.size(); k++, tot++)
            if (tot < cap) out_idx[tot] = (int64_t)r[k];
    }
    offsets[nq] = (int)tot;
    if (n) free(t.root);
    return tot;
}

// Timings of the reference's own code (ms), for bench.py's kd rows. which: 0 value-tree build,
// 1 nearest x nq, 2 radius x nq (value tree), 3 frame_kdtree build, 4 frame_kdtree radius x nq.
double vbref_kdtree_time_ms(int which, const float *pts, int n, const float *q, int nq, float radius, int reps) {
    std::vector<cv::Point2f> v(n);
    for (int i = 0; i < n; i++) v[i] = cv::Point2f(pts[2 * i], pts[2 * i + 1]);
    volatile float sink = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (int rep = 0; rep < reps; rep++) {
        if (which == 0) { KDTree t; construct_kdtree(t, v); sink += t.root[0].pt.x; free(t.root); }
        if (which == 3) { frame_kdtree t; construct_kdtree(t, v); sink += (float)t.root[0].pt_index; free(t.root); }
    }
    if (which == 1 || which == 2) {
        KDTree t; construct_kdtree(t, v);
        t0 = std::chrono::steady_clock::now();
        for (int rep = 0; rep < reps; rep++)
            for (int i = 0; i < nq; i++) {
                if (which == 1) sink += nearest(t, cv::Point2f(q[2 * i], q[2 * i + 1])).x;
                else sink += (float)radius_search(t, cv::Point2f(q[2 * i], q[2 * i + 1]), radius).size();
            }
        auto t1 = std::chrono::steady_clock::now();
        free(t.root);
        return std::chrono::duration<double, std::milli>(t1 - t0).count() / reps;
    }
    if (which == 4) {
        frame_kdtree t; construct_kdtree(t, v);
        t0 = std::chrono::steady_clock::now();
        for (int rep = 0; rep < reps; rep++)
            for (int i = 0; i < nq; i++) sink += (float)radius_search(t, v, cv::Point2f(q[2 * i], q[2 * i + 1]), radius).size();
        auto t1 = std::chrono::steady_clock::now();
        free(t.root);
        return std::chrono::duration<double, std::milli>(t1 - t0).count() / reps;
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::milli>(t1 - t0).count() / reps;
}

// ---- RansacFilter (include/RansacFilter.h:9-25) -----------------------------------------------
int vbref_initialize_sets(int n_matches, int min_items, int max_iterations, unsigned seed, int32_t *sets) {
    RansacFilter rf(min_items, max_iterations, 1.0f);
    vbo_ref_seed_set(seed);
    rf.initialize_sets(n_matches);
    for (int i = 0; i < max_iterations; i++)
        for (int j = 0; j < 8; j++) sets[i * 8 + j] = rf.ransac_sets[i][j];
    return 0;
}

// returns 1 if a model was accepted (fundamental non-empty), else 0
int vbref_find_fundamental(const float *p1, int n1, const float *p2, int n2, const int32_t *matches, int m,
                           int min_items, int max_iterations, float threshold, unsigned seed, float *F,
                           uint8_t *mask, int *mask_len) {
    std::vector<cv::Point2f> a(n1), b(n2);
    for (int i = 0; i < n1; i++) a[i] = cv::Point2f(p1[2 * i], p1[2 * i + 1]);
    for (int i = 0; i < n2; i++) b[i] = cv::Point2f(p2[2 * i], p2[2 * i + 1]);
    std::vector<std::pair<int, int> > mm(m);
    for (int i = 0; i < m; i++) mm[i] = std::make_pair(matches[2 * i], matches[2 * i + 1]);
    RansacFilter rf(min_items, max_iterations, threshold);
    vbo_ref_seed_set(seed);
    std::vector<bool> inl;
    cv::Mat Fm;
    rf.find_fundamental(a, b, mm, inl, Fm);
    *mask_len = (int)inl.size();
    for (size_t i = 0; i < inl.size(); i++) mask[i] = inl[i] ? 1 : 0;
    if (Fm.empty()) return 0;
    for (int i = 0; i < 9; i++) F[i] = Fm.at<float>(i / 3, i % 3);
    return 1;
}

int vbref_compute_fundamental(const float *p1set, const float *p2set, float *F) {
    std::vector<cv::Point2f> a(8), b(8);
    for (int i = 0; i < 8; i++) { a[i] = cv::Point2f(p1set[2 * i], p1set[2 * i + 1]); b[i] = cv::Point2f(p2set[2 * i], p2set[2 * i + 1]); }
    RansacFilter rf;
    cv::Mat Fm;
    rf.compute_fundamental(a, b, Fm);
    for (int i = 0; i < 9; i++) F[i] = Fm.at<float>(i / 3, i % 3);
    return 0;
}

int vbref_residual(const float *p1, int n1, const float *p2, int n2, const int32_t *matches, int m,
                   const float *F, float threshold, uint8_t *mask, int *n_inl, float *score) {
    std::vector<cv::Point2f> a(n1), b(n2);
    for (int i = 0; i < n1; i++) a[i] = cv::Point2f(p1[2 * i], p1[2 * i + 1]);
    for (int i = 0; i < n2; i++) b[i] = cv::Point2f(p2[2 * i], p2[2 * i + 1]);
    std::vector<std::pair<int, int> > mm(m);
    for (int i = 0; i < m; i++) mm[i] = std::make_pair(matches[2 * i], matches[2 * i + 1]);
    cv::Mat Fm(3, 3, CV_32FC1);
    for (int i = 0; i < 9; i++) Fm.at<float>(i / 3, i % 3) = F[i];
    RansacFilter rf(8, 1, threshold);
    std::vector<bool> inl;
    std::pair<int, float> r = rf.compute_fundamental_residual(a, b, mm, Fm, inl);
    for (int i = 0; i < m; i++) mask[i] = inl[i] ? 1 : 0;
    *n_inl = r.first;
    *score = r.second;
    return 0;
}

// Wall time (ms) of the reference's find_fundamental on one problem — bench.py reference arm.
double vbref_find_fundamental_time_ms(const float *p1, int n1, const float *p2, int n2, const int32_t *matches,
                                      int m, int max_iterations, float threshold, unsigned seed, int reps) {
    std::vector<cv::Point2f> a(n1), b(n2);
    for (int i = 0; i < n1; i++) a[i] = cv::Point2f(p1[2 * i], p1[2 * i + 1]);
    for (int i = 0; i < n2; i++) b[i] = cv::Point2f(p2[2 * i], p2[2 * i + 1]);
    std::vector<std::pair<int, int> > mm(m);
    for (int i = 0; i < m; i++) mm[i] = std::make_pair(matches[2 * i], matches[2 * i + 1]);
    RansacFilter rf(8, max_iterations, threshold);
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; r++) {
        vbo_ref_seed_set(seed + r);
        std::vector<bool> inl;
        cv::Mat Fm;
        rf.find_fundamental(a, b, mm, inl, Fm);
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::milli>(t1 - t0).count() / reps;
}
}
