/* vb_oracle — CPU restatement of the reference's frame-to-frame correspondence path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, load or call anything in oracle/. The product
 * library (vslam_b200/csrc) never includes, links or falls back to this code.
 *
 * Everything here restates rahulaggarwal965/vslam (paths relative to /root/reference):
 *   KD-tree        src/KDTree.cpp:3-35 (build), :37-71 (nearest), :73-101 and :145-171 (radius)
 *   matcher        src/Frame.cpp:82-105 (BFMatcher NORM_HAMMING knnMatch k=2, ratio 0.7, inlier copy-out)
 *   RansacFilter   src/RansacFilter.cpp:6-34 (sample sets), :36-67 (loop + best model),
 *                  :69-103 (8-point), :105-140 (residual / inliers / score)
 *   search by projection   src/vslam.cpp:129-161, src/PointMap.cpp:36-46 (orb_distance)
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - KD-tree: pinned against the reference's own src/KDTree.cpp compiled unmodified
 *     (oracle/_ref/libvbref.so) and its tests/test_kdtree.cpp (1000/1000 twice).
 *   - sampling: pinned against libstdc++'s std::mt19937 + uniform_int_distribution through the
 *     reference's own initialize_sets (oracle/_ref, seed hook in tests/cvlite).
 *   - residual, mask, count, knnMatch order, ratio test: pinned bit-for-bit against Python cv2 4.13.0
 *     calling the same OpenCV entry points the reference calls (tests/golden/gen_golden.py).
 *   - 8-point solve (cv::SVDecomp) and score reduction order (cv::sum): the reference leaves these to
 *     whatever OpenCV build is installed (LAPACK vs Jacobi, SIMD width). They are DEFINED here
 *     (fp64 Householder null vector, fp64 one-sided Jacobi 3x3 SVD, three-level blocked fp64 sum)
 *     and only checked against cv2 to a tolerance. PARITY UNPINNED for those two steps.
 */
#ifndef VB_ORACLE_H
#define VB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- score reduction order (defines cv::sum at src/RansacFilter.cpp:138) ------------------- */
#define VBO_SUM_CHUNK 128 /* matches per level-1 block, summed sequentially in fp64            */
#define VBO_SUM_GROUP 64  /* level-1 blocks per level-2 block; level-3 adds level-2 sums in order */
double vbo_score_sum(const float *e, int n);

/* ---- std::mt19937 + libstdc++ uniform_int_distribution (src/RansacFilter.cpp:15-16,24-26) -- */
typedef struct {
    uint32_t mt[624];
    int idx;
} vbo_mt19937;
void vbo_mt_seed(vbo_mt19937 *g, uint32_t seed);
uint32_t vbo_mt_next(vbo_mt19937 *g);
/* uniform_int_distribution<int>(0, hi)(g) as implemented by libstdc++ 13 for a 32-bit URBG. */
int vbo_uniform_int(vbo_mt19937 *g, int hi);

/* src/RansacFilter.cpp:6-34. sets is [max_iterations][8] (inner size is hard-coded 8 at :17;
 * entries j >= min_items stay 0). Requires 1 <= min_items <= 8 and n_matches >= min_items. */
void vbo_initialize_sets(int n_matches, int min_items, int max_iterations, uint32_t seed, int32_t *sets);

/* ---- 8-point solve (src/RansacFilter.cpp:69-103) ------------------------------------------- */
/* A (8x9, fp32, rows built exactly as :81-89) -> unit null vector, fp64 Householder QR of A^T. */
void vbo_null_vector_8x9(const float *A, float *f9);
/* F (3x3 fp32) -> U, D (descending), Vt in fp32 via fp64 one-sided Jacobi (Hestenes). */
void vbo_svd3x3(const float *F, float *U, float *D, float *Vt);
/* p1set/p2set: 8 (x,y) pairs. F row-major with x2^T F x1 = 0, rank 2 enforced as at :98-101. */
void vbo_compute_fundamental(const float *p1set, const float *p2set, float *F);

/* ---- residual (src/RansacFilter.cpp:105-140) ------------------------------------------------ */
/* e = s*s/(a0*a0) + a1*a1 + b0*b0 + b1*b1 as parsed at :126, for one correspondence. */
float vbo_residual_one(const float *F, float x1, float y1, float x2, float y2);
/* p1: n1x2, p2: n2x2, matches: mx2 (first indexes p1, second indexes p2). mask/e_out may be NULL. */
void vbo_compute_fundamental_residual(const float *p1, const float *p2, const int32_t *matches, int m,
                                      const float *F, float threshold, uint8_t *mask, float *e_out,
                                      int *n_inliers, float *score);

/* ---- find_fundamental (src/RansacFilter.cpp:36-67) ------------------------------------------ */
/* Returns 0, or -1 if m < min_items / bad min_items (the reference is UB there; nothing is written).
 * best_hyp = -1 when no hypothesis ever replaced the initial (0 inliers, score 0) best: then F and
 * mask are left untouched, as the reference leaves `fundamental` empty and `inliers` as passed.
 * Optional per-hypothesis outputs (may be NULL): sets_out [iters][8], F_all [iters][9],
 * cnt_all [iters], score_all [iters]. */
int vbo_find_fundamental(const float *p1, const float *p2, const int32_t *matches, int m, int min_items,
                         int max_iterations, float threshold, uint32_t seed, float *F, uint8_t *mask,
                         int *n_inliers, float *score, int *best_hyp, int32_t *sets_out, float *F_all,
                         int32_t *cnt_all, float *score_all);

/* ---- matcher (src/Frame.cpp:82-105) --------------------------------------------------------- */
/* BFMatcher(NORM_HAMMING).knnMatch(d1, d2, k=2): per query the two smallest distances, ties to the
 * lower train index (cv2 4.13 behaviour). idx/dist are [n1][2]. Requires n2 >= 2. */
void vbo_knn2_hamming(const uint8_t *d1, int n1, const uint8_t *d2, int n2, int bytes, int32_t *idx,
                      int32_t *dist);
/* Lowe ratio test exactly as :91 — float distances, double product and compare. */
int vbo_ratio_keep(int d0, int d1, double ratio);
/* knn2 + ratio; out_pairs [<=n1][2] in query order; returns the number kept. */
int vbo_match_hamming(const uint8_t *d1, int n1, const uint8_t *d2, int n2, int bytes, double ratio,
                      int32_t *out_pairs);
/* Float descriptors (BASELINE config 3; no reference counterpart — build-defined): squared L2 as a
 * sequential fp32 sum of (a_k-b_k)^2, distance = sqrtf, same tie and ratio rules. */
void vbo_knn2_l2f(const float *d1, int n1, const float *d2, int n2, int dim, int32_t *idx, float *dist);
int vbo_match_l2f(const float *d1, int n1, const float *d2, int n2, int dim, double ratio, int32_t *out_pairs);

/* Whole match_features (:82-105): matcher, RANSAC, inlier copy-out. out_matches [<=n1][2].
 * Returns number of final matches, or -1 when RANSAC could not run (fewer than min_items tentative
 * matches). n_tentative receives the ratio-test survivor count. */
int vbo_match_features(const float *p1, const uint8_t *d1, int n1, const float *p2, const uint8_t *d2,
                       int n2, int bytes, double ratio, int min_items, int max_iterations,
                       float threshold, uint32_t seed, int32_t *out_matches, float *F, int *n_tentative,
                       int *best_hyp);

/* ---- KD-tree (src/KDTree.cpp) --------------------------------------------------------------- */
/* The reference stores nodes in DFS pre-order in one array (root[size++], :16/:136) with
 * left = len/2 points and right = len - len/2 - 1 (:8/:127). This restatement returns that array:
 * pre_idx[k] = index into pts of the point stored at pre-order slot k. Ties on the split coordinate
 * are ordered by original index (std::nth_element leaves them unspecified), so the result equals
 * the reference's whenever split coordinates are distinct. */
void vbo_kdtree_build(const float *pts, int n, int32_t *pre_idx);
int vbo_kdtree_height(int n);
/* nearest (:37-71): returns the pre-order slot of the winner or -1 if nothing is closer than
 * max_d2 (the reference then returns a default Point2f{0,0}, :38-42). */
int vbo_kdtree_nearest(const float *pts, const int32_t *pre_idx, int n, float qx, float qy, float max_d2,
                       float *out_d2);
/* radius_search (:73-101, :145-171): writes matching point indices in DFS pre-order; returns count
 * (may exceed cap; only cap entries are written). */
/* k > 1 nearest neighbours: build-defined generalisation of `nearest` (no reference behaviour exists: the declarations at
 * include/KDTree.h:39-42,74-77 are commented out). Ascending by squared distance, first visited first among equals.
 * out_slot / out_d2 have room for k entries; returns how many were found (< k when fewer points lie within max_d2). */
int vbo_kdtree_knn(const float *pts, const int32_t *pre_idx, int n, float qx, float qy, int k, float max_d2,
                   int32_t *out_slot, float *out_d2);
int vbo_kdtree_radius(const float *pts, const int32_t *pre_idx, int n, float qx, float qy, float radius,
                      int32_t *out_idx, int cap);

/* ---- "next" rows ---------------------------------------------------------------- */
/* src/PointMap.cpp:36-46 orb_distance: min Hamming distance between desc and each of k observations. */
uint32_t vbo_orb_distance(const uint8_t *desc, const uint8_t *obs, int k, int bytes);
/* src/vslam.cpp:131: X [n][4] homogeneous map points times c2^T (c2 = 3x4 camera, row-major) -> out3 [n][3],
 * in the arithmetic cv::gemm uses for that shape (pinned against cv2 4.13 goldens; see the .c file). */
void vbo_project_points(const float *X, int n, const float *c2, float *out3);
/* src/vslam.cpp:129-161 search by projection: project every map point, keep those that land inside the W x H
 * image, radius_search(r) the frame's kd-tree (pts [k][2], pre_idx from vbo_kdtree_build) around each, and let the
 * map point claim the first hit (pre-order) that is still free (map_point_ids[idx] < 0) and whose orb_distance
 * to the point's observations (obs_desc rows obs_off[i] .. obs_off[i+1]) is < dist_thr. Map points are processed
 * in index order; a claimed keypoint is no longer free for later ones. assign[i] = claimed keypoint or -1;
 * map_point_ids is updated in place; proj_xy [n][2] / in_view [n] are optional outputs. Returns the claim count. */
int vbo_search_by_projection(const float *X, int n, const float *c2, int W, int H, const float *pts,
                             const int32_t *pre_idx, int k, const uint8_t *desc, int bytes,
                             int32_t *map_point_ids, const int32_t *obs_off, const uint8_t *obs_desc,
                             float radius, uint32_t dist_thr, int32_t *assign, float *proj_xy,
                             uint8_t *in_view);

/* src/helpers.cpp:3-35 extract_Rt and :37-80 triangulate. E = K^T F K is pinned bit for bit
 * against cv2; the SVDs (cv::SVD::compute, OpenCV-build dependent) are DEFINED here as fp64 one-sided Jacobi —
 * PARITY UNPINNED for R, t and the triangulated points, which are checked against cv2 to a tolerance only. */
void vbo_essential(const float *F, const float *K, float *E);
void vbo_extract_rt(const float *F, const float *K, float *R, float *t);
void vbo_null_vector_4x4(const float *A, float *v4);
void vbo_triangulate(const float *p1, const float *p2, int n, const float *c1, const float *c2, float *out);
/* src/vslam.cpp:186-251: reprojection into both cameras, the (partial, as written) dehomogenisation, squared errors, gate. */
int vbo_reprojection_gate(const float *points4, int n, const float *c1, const float *c2, const float *ip1, const float *ip2,
                          const int32_t *map_point_ids, float threshold_sq, float *re1_out, float *re2_out,
                          int32_t *inlier_idx, double *reproj_error);

/* ---- opt-in mode (NOT reference behaviour; the default path is untouched): Hartley-normalised solve and / or the true
 * Sampson distance, the two defects the reference flags itself (src/RansacFilter.cpp:40, :125-126). ---- */
#define VBO_RANSAC_HARTLEY 1u
#define VBO_RANSAC_SAMPSON 2u
void vbo_compute_fundamental_hartley(const float *p1set, const float *p2set, float *F);
float vbo_sampson_one(const float *F, float x1, float y1, float x2, float y2);
int vbo_find_fundamental_ex(const float *p1, const float *p2, const int32_t *matches, int m, int min_items, int max_iterations,
                            float threshold, uint32_t seed, unsigned flags, float *F, uint8_t *mask, int *n_inliers, float *score,
                            int *best_hyp, float *F_all, int32_t *cnt_all, float *score_all);

/* ---- seed hook consumed by the cvlite random_device stand-in (oracle/_ref builds only) -------- */
void vbo_ref_seed_set(unsigned seed);
unsigned vbo_ref_seed_next(void);

/* ---- bench helper: many pairs over all host threads (OpenMP), for bench.py's CPU legs --------- */
/* frames: pts [nframes][k][2], desc [nframes][k][bytes]; pair i = (frame i, frame i+1).
 * Returns total final matches (checksum). n_threads <= 0 -> omp default. */
long vbo_pairs_run(const float *pts, const uint8_t *desc, int nframes, int k, int bytes, double ratio,
                   int max_iterations, float threshold, uint32_t seed0, int n_threads, int *threads_used);

#ifdef __cplusplus
}
#endif
#endif
