// adapter_harness.cpp — drives the drop-in C++ adapters (include/KDTree.h, include/RansacFilter.h) exactly
// the way the reference's callers do (src/Frame.cpp:76,97; src/vslam.cpp:19,149,295-297;
// tests/test_kdtree.cpp:16-37,54-57,103-107) and dumps what they return, so a pytest can compare it with
// the oracle. Built against the test-only OpenCV stand-in (tests/cvlite) — a real integrator uses OpenCV.
//
//   adapter_harness <in.bin> <out.bin>      run the calls described by in.bin
//   adapter_harness --protocol <trials>     replay the reference's own randomized kd-tree test protocol
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "KDTree.h"
#include "RansacFilter.h"

template <typename T> static void rd(FILE *f, T *p, size_t n) {
    if (n && fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}
template <typename T> static void wr(FILE *f, const T *p, size_t n) {
    if (n && fwrite(p, sizeof(T), n, f) != n) { fprintf(stderr, "short write\n"); exit(2); }
}

// walk the pointer tree the way tests/test_kdtree.cpp:15-37 does (via left/right), pre-order
static void walk(const KDTree::KDTreeNode *nd, std::vector<float> &out) {
    if (!nd) return;
    out.push_back(nd->pt.x); out.push_back(nd->pt.y);
    walk(nd->left, out);
    walk(nd->right, out);
}
static void walk(const frame_kdtree::KDTreeNode *nd, std::vector<int64_t> &out) {
    if (!nd) return;
    out.push_back((int64_t)nd->pt_index);
    walk(nd->left, out);
    walk(nd->right, out);
}

static int protocol(int trials) {
    // tests/test_kdtree.cpp:47-146: integer points in [0,100)^2, 2500..2999 of them; nearest passes on
    // equal distance, radius passes on set equality.
    srand(12345);
    int ok_nn = 0, ok_rs = 0;
    for (int t = 0; t < trials; t++) {
        const int size = rand() % 500 + 2500;
        std::vector<cv::Point2f> arr;
        for (int j = 0; j < size; j++) arr.emplace_back((float)(rand() % 100), (float)(rand() % 100));
        KDTree kd;
        construct_kdtree(kd, arr);
        cv::Point2f qp((float)(rand() % 100), (float)(rand() % 100));
        cv::Point2f nn = nearest(kd, qp);
        float best = INFINITY;
        for (auto &p : arr) { cv::Point2f d = qp - p; float c = d.dot(d); if (c < best) best = c; }
        cv::Point2f dn = qp - nn;
        if (dn.dot(dn) == best) ok_nn++;
        const float radius = 10.f + 90.f * (float)rand() / (float)RAND_MAX;
        std::vector<cv::Point2f> found = radius_search(kd, qp, radius), want;
        for (auto &p : arr) { cv::Point2f d = qp - p; if (d.dot(d) < radius * radius) want.push_back(p); }
        auto cmp = [](const cv::Point2f &a, const cv::Point2f &b) { return a.x == b.x ? a.y < b.y : a.x < b.x; };
        std::sort(found.begin(), found.end(), cmp);
        std::sort(want.begin(), want.end(), cmp);
        bool same = found.size() == want.size();
        for (size_t i = 0; same && i < found.size(); i++) same = found[i] == want[i];
        if (same) ok_rs++;
        free(kd.root);   // the caller's side of the ownership contract (tests/test_kdtree.cpp:88)
    }
    printf("%d successes out of %d trials\n%d successes out of %d trials\n", ok_nn, trials, ok_rs, trials);
    return (ok_nn == trials && ok_rs == trials) ? 0 : 1;
}

int main(int argc, char **argv) {
    if (argc == 3 && std::string(argv[1]) == "--protocol") return protocol(atoi(argv[2]));
    if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin | --protocol N\n", argv[0]); return 2; }
    FILE *fi = fopen(argv[1], "rb"), *fo = fopen(argv[2], "wb");
    if (!fi || !fo) { perror("open"); return 2; }
    int32_t hdr[7];
    float fh[2];
    rd(fi, hdr, 7);
    rd(fi, fh, 2);
    const int n1 = hdr[0], n2 = hdr[1], m = hdr[2], nq = hdr[3], iters = hdr[4], min_items = hdr[5];
    const uint32_t seed = (uint32_t)hdr[6];
    const float thr = fh[0], radius = fh[1];
    std::vector<float> p1f(2 * n1), p2f(2 * n2), qf(2 * nq);
    std::vector<int32_t> mmf(2 * m);
    rd(fi, p1f.data(), p1f.size()); rd(fi, p2f.data(), p2f.size()); rd(fi, mmf.data(), mmf.size()); rd(fi, qf.data(), qf.size());
    fclose(fi);
    std::vector<cv::Point2f> p1(n1), p2(n2), q(nq);
    for (int i = 0; i < n1; i++) p1[i] = cv::Point2f(p1f[2 * i], p1f[2 * i + 1]);
    for (int i = 0; i < n2; i++) p2[i] = cv::Point2f(p2f[2 * i], p2f[2 * i + 1]);
    for (int i = 0; i < nq; i++) q[i] = cv::Point2f(qf[2 * i], qf[2 * i + 1]);
    std::vector<std::pair<int, int> > matches(m);
    for (int i = 0; i < m; i++) matches[i] = std::make_pair(mmf[2 * i], mmf[2 * i + 1]);

    // ---- value tree over frame-2 points ----
    KDTree kd;
    construct_kdtree(kd, p2);
    std::vector<float> pre;
    walk(kd.root, pre);
    int32_t meta[2] = {(int32_t)kd.size, (int32_t)kd.height};
    wr(fo, meta, 2);
    wr(fo, pre.data(), pre.size());
    std::vector<float> nn(2 * nq);
    for (int i = 0; i < nq; i++) {   // single-query calls for the first few, batch for the rest
        if (i < 8) { cv::Point2f r = nearest(kd, q[i]); nn[2 * i] = r.x; nn[2 * i + 1] = r.y; }
    }
    std::vector<cv::Point2f> nb = nearest_batch(kd, q);
    for (int i = 8; i < nq; i++) { nn[2 * i] = nb[i].x; nn[2 * i + 1] = nb[i].y; }
    for (int i = 0; i < nq && i < 8; i++)
        if (nb[i] != cv::Point2f(nn[2 * i], nn[2 * i + 1])) { fprintf(stderr, "single/batch nearest differ\n"); return 3; }
    wr(fo, nn.data(), nn.size());
    std::vector<std::vector<cv::Point2f> > rv = radius_search_batch(kd, q, radius);
    for (int i = 0; i < nq; i++) {
        int32_t c = (int32_t)rv[i].size();
        wr(fo, &c, 1);
        for (auto &p : rv[i]) { float xy[2] = {p.x, p.y}; wr(fo, xy, 2); }
    }
    // forget the device copy: the next query must transparently re-import the tree from the host nodes
    vslam_b200_kdtree_release(kd.root);
    if (nq) {
        std::vector<cv::Point2f> again = radius_search(kd, q[0], radius);
        if (again.size() != rv[0].size()) { fprintf(stderr, "re-import changed the result\n"); return 3; }
        for (size_t k = 0; k < again.size(); k++) if (again[k] != rv[0][k]) { fprintf(stderr, "re-import order\n"); return 3; }
    }
    free(kd.root);
    // the same struct constructed a second time: `size` is now n2 + n1 (the reference never resets it, src/KDTree.cpp:16)
    // while root holds n1 nodes — a re-import after eviction must count the nodes, not trust `size`
    construct_kdtree(kd, p1);
    if ((int)kd.size != n1 + n2) { fprintf(stderr, "size must accumulate like the reference's\n"); return 3; }
    if (nq) {
        std::vector<cv::Point2f> before = radius_search(kd, q[0], radius * 4);
        vslam_b200_kdtree_release(kd.root);
        std::vector<cv::Point2f> after = radius_search(kd, q[0], radius * 4);
        if (before.size() != after.size()) { fprintf(stderr, "re-import of a twice-constructed struct changed the result\n"); return 3; }
        for (size_t k = 0; k < after.size(); k++) if (before[k] != after[k]) { fprintf(stderr, "re-import (2) order\n"); return 3; }
    }
    free(kd.root);

    // ---- index tree (what Frame carries) ----
    frame_kdtree fk;
    construct_kdtree(fk, p2);
    std::vector<int64_t> fpre;
    walk(fk.root, fpre);
    meta[0] = (int32_t)fk.size; meta[1] = (int32_t)fk.height;
    wr(fo, meta, 2);
    wr(fo, fpre.data(), fpre.size());
    for (int i = 0; i < nq; i++) {
        std::vector<usize> r = (i < 8) ? radius_search(fk, p2, q[i], radius)
                                       : radius_search_batch(fk, p2, std::vector<cv::Point2f>(1, q[i]), radius)[0];
        int32_t c = (int32_t)r.size();
        wr(fo, &c, 1);
        for (usize v : r) { int64_t w = (int64_t)v; wr(fo, &w, 1); }
    }
    free(fk.root);   // src/vslam.cpp:295-297

    // ---- RansacFilter ----
    vslam_b200_set_ransac_seed(seed);
    RansacFilter rf(min_items, iters, thr);
    std::vector<bool> inliers;
    cv::Mat F;
    rf.find_fundamental(p1, p2, matches, inliers, F);   // uses seed
    int32_t accepted = F.empty() ? 0 : 1, ilen = (int32_t)inliers.size();
    wr(fo, &accepted, 1);
    wr(fo, &ilen, 1);
    float Ff[9] = {0};
    if (accepted) for (int i = 0; i < 9; i++) Ff[i] = F.at<float>(i / 3, i % 3);
    wr(fo, Ff, 9);
    for (int i = 0; i < ilen; i++) { uint8_t b = inliers[i] ? 1 : 0; wr(fo, &b, 1); }
    int32_t cnt = -1;
    float score = 0.f;
    if (accepted) {
        std::vector<bool> in2;
        std::pair<int, float> r = rf.compute_fundamental_residual(p1, p2, matches, F, in2);
        cnt = r.first; score = r.second;
        if (in2 != inliers) { fprintf(stderr, "residual mask differs from find_fundamental mask\n"); return 3; }
    }
    wr(fo, &cnt, 1);
    wr(fo, &score, 1);
    if (m >= min_items) {
        rf.initialize_sets(m);   // uses seed + 1
        for (int i = 0; i < iters; i++)
            for (int j = 0; j < 8; j++) { int32_t v = rf.sample_sets()[i][j]; wr(fo, &v, 1); }
    }
    if (m >= 8) {
        std::vector<cv::Point2f> s1(8), s2(8);
        for (int j = 0; j < 8; j++) { s1[j] = p1[matches[j].first]; s2[j] = p2[matches[j].second]; }
        cv::Mat F8;
        rf.compute_fundamental(s1, s2, F8);
        for (int i = 0; i < 9; i++) Ff[i] = F8.at<float>(i / 3, i % 3);
        wr(fo, Ff, 9);
    }
    fclose(fo);
    return 0;
}
