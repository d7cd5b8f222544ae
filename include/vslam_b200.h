/* vslam_b200.h — C ABI of the B200-native correspondence path (libvslam_b200.so).
 *
 * This is the drop-in boundary for the reference's frame-to-frame correspondence hot path
 * (rahulaggarwal965/vslam; file:line below are relative to the reference tree):
 *
 *   KD-tree build      construct_kdtree        src/KDTree.cpp:25-35, :107-121   include/KDTree.h:25,60
 *   KD-tree 1-NN       nearest                 src/KDTree.cpp:37-71             include/KDTree.h:30
 *   KD-tree radius     radius_search           src/KDTree.cpp:73-101, :145-171  include/KDTree.h:44,79
 *   descriptor match   match_features (front)  src/Frame.cpp:82-95              include/Frame.h:34
 *   RANSAC             RansacFilter::*         src/RansacFilter.cpp:6-140       include/RansacFilter.h:9-25
 *
 * The reference has no FFI: callers include KDTree.h / RansacFilter.h and link the objects. The
 * adapters in include/KDTree.h and include/RansacFilter.h of THIS repo keep those C++ interfaces and
 * forward to the functions below (INTEGRATION.md shows the wiring).
 *
 * Conventions
 *   - plain C types only; every function returns VB_OK (0) or a VB_ERR_* code and writes no output on
 *     failure. vb_last_error() returns a thread-local message for the last failure.
 *   - pointers are HOST pointers unless the function name ends in _d, in which case every array
 *     argument is a DEVICE pointer on the context's GPU and the call only enqueues work on the
 *     context's stream (call vb_synchronize before reading results).
 *   - points are (x, y) float pairs; matches / pairs are (first, second) int32 pairs, first indexing
 *     frame 1 (query) and second frame 2 (train), as std::pair<int,int> in the reference.
 *   - there is NO CPU fallback: without a CUDA device vb_create fails with VB_ERR_CUDA.
 */
#ifndef VSLAM_B200_H
#define VSLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB_OK 0
#define VB_ERR_INVALID 1      /* bad argument (NULL, zero size where forbidden, min_items outside 1..8 ...) */
#define VB_ERR_CUDA 2         /* a CUDA runtime call failed; see vb_last_error() */
#define VB_ERR_TOO_FEW 3      /* fewer matches than min_items / fewer than 2 train descriptors (UB in the reference) */
#define VB_ERR_CAPACITY 4     /* caller-provided output capacity too small; required size is reported */
#define VB_ERR_NO_MODEL 5     /* RANSAC accepted no hypothesis (reference leaves `fundamental` empty) */

typedef struct vb_ctx vb_ctx;   /* one per GPU per host thread; owns a stream and grow-only workspaces */
typedef struct vb_tree vb_tree; /* device-resident SoA k-d tree */

int vb_version(void);
const char *vb_last_error(void);

int vb_create(int device, vb_ctx **out);
int vb_destroy(vb_ctx *ctx);
/* Run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL restores the context's own. */
int vb_set_stream(vb_ctx *ctx, void *cuda_stream);
int vb_synchronize(vb_ctx *ctx);
/* Path selectors for tests and measurements — the library never reads the environment. Every option chooses between code
 * paths that return the same bits (the GPU tests force each of them against the oracle); unset = the built-in rule.
 *   hamming_tc 0/1        popcount matcher / tcgen05 matcher regardless of problem size
 *   hamming_fp4 0         tcgen05 matcher on kind::f8f6f4 (e4m3) instead of kind::mxf4
 *   tc_fix8 0             the matcher's fix pass always reads both candidate groups
 *   hamming_qpt 1/2/4     queries per thread of the popcount matcher
 *   l2_tc 0/1             float descriptors: exact SIMT kernel / bf16 tcgen05 prefilter + exact re-evaluation
 *   ransac_lazy 0         the pair pipeline scores every hypothesis in full instead of counting
 *   ransac_prune 0/1/2    bounded counting never / for batches that fill the machine (default) / always
 *   prune_first_chunks, prune_first16, prune_growth16, prune_rounds, prune_item_chunks   its checkpoint schedule
 *   count_packed 0, score_packed 0   scalar-instruction versions of the counting / scoring kernels
 *   kd_lanes_per_query 1/8/32        k-d tree 1-NN: lanes that share one query (A/B of the thread-per-query mapping)
 *   tc_drain 0..7 (default 6), tc_svc_hi 0/1     organisation of the tensor matcher's drain / placement of its service warps
 *   pairs_overlap 0                  vb_pairs_submit[_d]: every submission on one compute stream instead of alternating two
 * A build with -DVB_TUNING adds timing-only options (tc_dbg, prune_ctas_per_sm, pairs_twin, pairs_split). Unknown names
 * return VB_ERR_INVALID. vb_reset_options restores every built-in rule. */
int vb_set_option(vb_ctx *ctx, const char *name, long long value);
int vb_reset_options(vb_ctx *ctx);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
uint64_t vb_launch_count(const vb_ctx *ctx);

/* ------------------------------------------------------------------------------------------------
 * KD-tree over 2-D keypoints. Replaces construct_kdtree / nearest / radius_search.
 * The tree is an implicit balanced tree stored in DFS pre-order (the reference's node-array order,
 * src/KDTree.cpp:16): slot s with subtree length len has its left child at s+1 (len/2 points) and its
 * right child at s+1+len/2 (len-len/2-1 points). Node arrays are SoA in HBM: x[], y[], idx[].
 * Ties on the split coordinate are ordered by original index (std::nth_element leaves them
 * unspecified), so the layout equals the reference's whenever split coordinates are distinct.
 * ---------------------------------------------------------------------------------------------- */
int vb_kdtree_build(vb_ctx *ctx, const float *pts_xy, uint32_t n, vb_tree **out);
int vb_kdtree_build_d(vb_ctx *ctx, const float *pts_xy_d, uint32_t n, vb_tree **out);
/* One tree per frame of a sequence in a single launch (one CTA per tree): pts_xy_d [ntrees][n][2] on the device, out[ntrees].
 * The trees share one allocation; release them together with vb_kdtree_free_batch. (construct_kdtree once per frame,
 * src/Frame.cpp:76, for every frame of a batch.) */
int vb_kdtree_build_batch_d(vb_ctx *ctx, const float *pts_xy_d, uint32_t ntrees, uint32_t n, vb_tree **out);
int vb_kdtree_free_batch(vb_tree **trees, uint32_t ntrees);
/* Re-create the device tree from a pre-order node array that already exists on the host (what the
 * reference's KDTree::root / frame_kdtree::root hold): pts_preorder[n*2], idx_preorder[n] (NULL = slot
 * numbers). Used by the C++ adapter when a caller hands it a tree it has no device copy of. */
int vb_kdtree_import(vb_ctx *ctx, const float *pts_preorder, const uint32_t *idx_preorder, uint32_t n, vb_tree **out);
int vb_kdtree_free(vb_tree *tree);
uint32_t vb_kdtree_size(const vb_tree *tree);
uint32_t vb_kdtree_height(const vb_tree *tree); /* floor(log2 n)+1, src/KDTree.cpp:33 */
/* Pre-order export (either output may be NULL): idx_preorder[n] = original index of the point at each
 * slot (frame_kdtree::KDTreeNode::pt_index), pts_preorder[n*2] = its coordinates (KDTree::KDTreeNode::pt). */
int vb_kdtree_export(vb_tree *tree, uint32_t *idx_preorder, float *pts_preorder);
/* Exact 1-NN for nq queries, same visiting order as the reference (near child, node, far child iff
 * split^2 < best), so equidistant ties resolve identically. Outputs (any may be NULL):
 * out_pt[nq*2] point value ({0,0} if nothing is closer than max_d2, as src/KDTree.cpp:38-42),
 * out_idx[nq] original point index or -1, out_d2[nq] squared distance (max_d2 if none). */
int vb_kdtree_nearest(vb_tree *tree, const float *q_xy, uint32_t nq, float max_d2, float *out_pt,
                      int32_t *out_idx, float *out_d2);
int vb_kdtree_nearest_d(vb_tree *tree, const float *q_xy_d, uint32_t nq, float max_d2, float *out_pt_d,
                        int32_t *out_idx_d, float *out_d2_d);
/* k nearest neighbours per query, 1 <= k <= 32 — the north star's "kNN". The reference has no behaviour to match here: its
 * k_nearest declarations are commented out (include/KDTree.h:39-42, 74-77). Defined as the direct generalisation of `nearest`
 * (src/KDTree.cpp:45-71): same visiting order, the strict `<` tests (:64, :68) compare against the k-th best squared distance
 * so far (max_d2 until k candidates exist), equidistant points rank in visiting order. out_idx / out_d2 are [nq][k], ascending;
 * unused slots hold -1 / max_d2; out_count[nq] (optional) = how many were found. k = 1 equals vb_kdtree_nearest. */
int vb_kdtree_knn(vb_tree *tree, const float *q_xy, uint32_t nq, uint32_t k, float max_d2, int32_t *out_idx, float *out_d2,
                  uint32_t *out_count);
int vb_kdtree_knn_d(vb_tree *tree, const float *q_xy_d, uint32_t nq, uint32_t k, float max_d2, int32_t *out_idx_d,
                    float *out_d2_d, uint32_t *out_count_d);
/* Radius search for nq queries: all points with dist^2 < r^2 (strict, :91) in DFS pre-order, as CSR.
 * out_offsets[nq+1]; out_idx[cap] original indices. *out_total receives the total hit count; if it
 * exceeds cap the call returns VB_ERR_CAPACITY after filling out_offsets (retry with a larger buffer). */
int vb_kdtree_radius(vb_tree *tree, const float *q_xy, uint32_t nq, float radius, uint32_t *out_offsets,
                     uint32_t *out_idx, uint64_t cap, uint64_t *out_total);
int vb_kdtree_radius_d(vb_tree *tree, const float *q_xy_d, uint32_t nq, float radius, uint32_t *out_offsets_d,
                       uint32_t *out_idx_d, uint64_t cap, uint64_t *out_total /* host */);

/* ------------------------------------------------------------------------------------------------
 * Descriptor matching. Replaces BFMatcher(NORM_HAMMING).knnMatch(k=2) + Lowe ratio (src/Frame.cpp:83-95).
 * ---------------------------------------------------------------------------------------------- */
/* Two nearest train descriptors per query; ties go to the lower train index. idx/dist are [n1][2].
 * bytes must be a multiple of 4 and <= 64 (ORB: 32). Requires n2 >= 2. */
int vb_knn2_hamming(vb_ctx *ctx, const uint8_t *d1, uint32_t n1, const uint8_t *d2, uint32_t n2, uint32_t bytes,
                    int32_t *idx, int32_t *dist);
/* knn2 + ratio test `(double)d0 < (double)d1 * ratio` (:91); survivors in query order.
 * out_pairs has room for n1 pairs; *out_m receives the count. */
int vb_match_hamming(vb_ctx *ctx, const uint8_t *d1, uint32_t n1, const uint8_t *d2, uint32_t n2, uint32_t bytes,
                     double ratio, int32_t *out_pairs, uint32_t *out_m);
/* Float descriptors (BASELINE config 3; not in the reference, which is Hamming-only at :83). */
int vb_knn2_l2f(vb_ctx *ctx, const float *d1, uint32_t n1, const float *d2, uint32_t n2, uint32_t dim, int32_t *idx,
                float *dist);
int vb_match_l2f(vb_ctx *ctx, const float *d1, uint32_t n1, const float *d2, uint32_t n2, uint32_t dim, double ratio,
                 int32_t *out_pairs, uint32_t *out_m);

/* ------------------------------------------------------------------------------------------------
 * RansacFilter. Replaces find_fundamental and its helpers (src/RansacFilter.cpp:36-67).
 * seed is the value the reference would have obtained from std::random_device at :15 — the same seed
 * gives the same std::mt19937 stream, hence the same sample sets.
 * ---------------------------------------------------------------------------------------------- */
/* Outputs: F[9] row-major (x2^T F x1 = 0), inlier_mask[m] (0/1), *n_inliers, *score (the (float)cv::sum
 * of the winner's residuals), *best_hyp (index of the winning hypothesis). Optional outputs may be NULL.
 * Returns VB_ERR_TOO_FEW if m < min_items, VB_ERR_NO_MODEL if no hypothesis beat the initial
 * (0 inliers, score 0) best — in both cases nothing is written except *best_hyp = -1. */
int vb_ransac_fundamental(vb_ctx *ctx, const float *p1_xy, uint32_t n1, const float *p2_xy, uint32_t n2,
                          const int32_t *matches, uint32_t m, int min_items, uint32_t max_iterations, float threshold,
                          uint32_t seed, float *F, uint8_t *inlier_mask, int32_t *n_inliers, float *score,
                          int32_t *best_hyp);
/* Opt-in mode — NOT reference behaviour; flags = 0 is exactly vb_ransac_fundamental. The two defects the reference flags
 * itself and never repaired, as explicit switches:
 *   VB_RANSAC_HARTLEY  `//TODO: normalize` (src/RansacFilter.cpp:40): every 8-point sample is translated to its centroid and
 *                      scaled to mean distance sqrt(2) per image before the same solve; F = T2^T F^ T1, unit Frobenius norm.
 *   VB_RANSAC_SAMPSON  the mis-parenthesised residual (:125-126) becomes the true Sampson distance
 *                      (x2^T F x1)^2 / (a0^2 + a1^2 + b0^2 + b1^2), a = F x1, b = F^T x2 (double, narrowed to f32 once);
 *                      `threshold` is then in squared pixels.
 * Sample sets, the inlier test (e <= threshold), the score and the selection rule (:59) stay the reference's. Defined
 * operation by operation by the CPU checker's mode of the same name; bit-exact against it. */
#define VB_RANSAC_HARTLEY 1u
#define VB_RANSAC_SAMPSON 2u
int vb_ransac_fundamental_ex(vb_ctx *ctx, const float *p1_xy, uint32_t n1, const float *p2_xy, uint32_t n2,
                             const int32_t *matches, uint32_t m, int min_items, uint32_t max_iterations, float threshold,
                             uint32_t seed, uint32_t flags, float *F, uint8_t *inlier_mask, int32_t *n_inliers, float *score,
                             int32_t *best_hyp);
/* Per-hypothesis view of the same run, for parity tests: sets[iters][8], F_all[iters][9],
 * n_inliers[iters], score[iters]. Any output may be NULL. */
int vb_ransac_hypotheses(vb_ctx *ctx, const float *p1_xy, uint32_t n1, const float *p2_xy, uint32_t n2,
                         const int32_t *matches, uint32_t m, int min_items, uint32_t max_iterations, float threshold,
                         uint32_t seed, int32_t *sets, float *F_all, int32_t *n_inliers, float *score);
/* initialize_sets alone (:6-34): sets[iters][8] for n_matches candidates. */
int vb_ransac_sample_sets(vb_ctx *ctx, uint32_t n_matches, int min_items, uint32_t max_iterations, uint32_t seed,
                          int32_t *sets);
/* compute_fundamental_residual for ONE model with its inlier mask (:105-140). */
int vb_ransac_residual(vb_ctx *ctx, const float *p1_xy, uint32_t n1, const float *p2_xy, uint32_t n2,
                       const int32_t *matches, uint32_t m, const float *F, float threshold, uint8_t *inlier_mask,
                       int32_t *n_inliers, float *score);
/* Scoring only (compute_fundamental_residual, :105-140, for h models at once): corr[m][4] rows
 * (x1,y1,x2,y2); F[h][9]; outputs n_inliers[h], score[h]. The BASELINE "hypotheses scored/s" entry. */
int vb_ransac_score(vb_ctx *ctx, const float *corr, uint32_t m, const float *F, uint32_t h, float threshold,
                    int32_t *n_inliers, float *score);
int vb_ransac_score_d(vb_ctx *ctx, const float *corr_d, uint32_t m, const float *F_d, uint32_t h, float threshold,
                      int32_t *n_inliers_d, float *score_d);
/* Inlier counts only (the first member of compute_fundamental_residual's return pair, :138) — the kernel the pair pipeline
 * uses: a and s = x2 . a follow the reference's rounding sequence, the rest of the residual is evaluated with FMAs and an
 * approximate reciprocal under a rigorous error bound, and any evaluation within that bound of the threshold is redone
 * with the reference's sequence, so the counts equal vb_ransac_score's exactly. */
int vb_ransac_counts(vb_ctx *ctx, const float *corr, uint32_t m, const float *F, uint32_t h, float threshold,
                     int32_t *n_inliers);
int vb_ransac_counts_d(vb_ctx *ctx, const float *corr_d, uint32_t m, const float *F_d, uint32_t h, float threshold,
                       int32_t *n_inliers_d);
/* Work done by the pair pipeline's bounded counting since the context was created (or last reset): hypothesis x match
 * evaluations actually performed, and the number a full evaluation of every hypothesis on every match would have taken
 * (find_fundamental's loop, :52-64). A hypothesis is abandoned only when its count so far plus every match it has not seen
 * is below a count another hypothesis is known to reach, so the winner, its count, score and mask are unaffected. */
int vb_ransac_prune_stats(vb_ctx *ctx, uint64_t *evaluated, uint64_t *total, int reset);
/* The 8-point solve alone (compute_fundamental, :69-103) for h minimal samples: p1set/p2set [h][8][2]. */
int vb_ransac_solve8(vb_ctx *ctx, const float *p1set, const float *p2set, uint32_t h, float *F);

/* ------------------------------------------------------------------------------------------------
 * Whole pair(s): match_features (src/Frame.cpp:82-105) = knn2 + ratio + find_fundamental + inlier
 * copy-out, with everything between the input upload and the result download kept on the device.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    double ratio;            /* 0.7 at src/Frame.cpp:91 */
    int32_t min_items;       /* 8   at src/vslam.cpp:19 */
    uint32_t max_iterations; /* 100 at src/vslam.cpp:19; 1024 in BASELINE config 2 */
    float threshold;         /* 10  at src/vslam.cpp:19 */
    uint32_t seed0;          /* pair i uses seed0 + i */
} vb_pair_params;

typedef struct {
    int32_t status;      /* VB_OK, VB_ERR_TOO_FEW or VB_ERR_NO_MODEL for this pair (VB_ERR_CUDA: internal failure, no result) */
    int32_t n_tentative; /* matches surviving the ratio test */
    int32_t n_matches;   /* final (RANSAC inlier) matches written for this pair */
    int32_t best_hyp;
    int32_t n_inliers;
    float score;
    float F[9];
} vb_pair_result;

/* One pair. out_matches has room for n1 pairs. */
int vb_match_features(vb_ctx *ctx, const float *p1_xy, const uint8_t *d1, uint32_t n1, const float *p2_xy,
                      const uint8_t *d2, uint32_t n2, uint32_t bytes, const vb_pair_params *params,
                      int32_t *out_matches, vb_pair_result *result);
/* A sequence: frames [nframes][k] keypoints, pair i = (frame i, frame i+1), i in [0, nframes-1).
 * pts [nframes][k][2], desc [nframes][k][bytes]; results [nframes-1]; out_matches [nframes-1][k][2]
 * (may be NULL to skip the match download). All pairs run in a handful of batched launches. */
int vb_pairs_run(vb_ctx *ctx, const float *pts, const uint8_t *desc, uint32_t nframes, uint32_t k, uint32_t bytes,
                 const vb_pair_params *params, vb_pair_result *results, int32_t *out_matches);
int vb_pairs_run_d(vb_ctx *ctx, const float *pts_d, const uint8_t *desc_d, uint32_t nframes, uint32_t k,
                   uint32_t bytes, const vb_pair_params *params, vb_pair_result *results_d, int32_t *out_matches_d);
/* match_features with FLOAT descriptors (BASELINE config 3; the reference matcher is Hamming-only, src/Frame.cpp:83, so the
 * matcher's contract is vb_knn2_l2f's and everything downstream — ratio test :91, find_fundamental :97, inlier copy-out
 * :98-102 — is the reference's): d1 [n1][dim], d2 [n2][dim] fp32, dim 64 or 128. One call, nothing returns to the host
 * between the matcher and the inlier list. The _d variant takes device pointers, writes result_d[0] / out_matches_d[n1][2]
 * on the device and only enqueues. */
int vb_match_features_l2f(vb_ctx *ctx, const float *p1_xy, const float *d1, uint32_t n1, const float *p2_xy, const float *d2,
                          uint32_t n2, uint32_t dim, const vb_pair_params *params, int32_t *out_matches,
                          vb_pair_result *result);
int vb_match_features_l2f_d(vb_ctx *ctx, const float *p1_xy_d, const float *d1_d, uint32_t n1, const float *p2_xy_d,
                            const float *d2_d, uint32_t n2, uint32_t dim, const vb_pair_params *params,
                            int32_t *out_matches_d, vb_pair_result *result_d);

/* ------------------------------------------------------------------------------------------------
 * Streaming submission with a compact result download — for callers that process one sequence after another (the
 * reference's main loop calls match_features once per frame, src/vslam.cpp:92; a batch of frames is one submission).
 * Up to three submissions per context are in flight (a fourth returns VB_ERR_CAPACITY until the oldest ticket has been
 * waited for): one uploading while two compute — consecutive submissions run on two internal compute streams, the counting
 * and the small kernels of one beside the matcher of the other (option pairs_overlap = 0: one stream) — and the download
 * of the oldest one overlaps both. All pointers are HOST pointers and must stay valid until the ticket has been
 * waited for; copies overlap with compute only when they are pinned (vb_host_alloc / vb_host_register).
 *   results        [nframes-1]
 *   match_offsets  [nframes-1]  index (in matches, not bytes) of pair i's first match inside matches16
 *   matches16      [cap_matches][2]  (query, train) keypoint indices as uint16 — what match_features appends to
 *                  frame1.matches (src/Frame.cpp:98-102); pair i owns results[i].n_matches entries from match_offsets[i].
 *                  Requires k <= 65535. match_offsets == matches16 == NULL skips the match download.
 * vb_pairs_wait returns VB_ERR_CAPACITY (and the required size in *total_matches) when cap_matches is too small; the
 * vb_pair_result entries are valid in that case. Same results, bit for bit, as vb_pairs_run.
 * ---------------------------------------------------------------------------------------------- */
int vb_pairs_submit(vb_ctx *ctx, const float *pts, const uint8_t *desc, uint32_t nframes, uint32_t k, uint32_t bytes,
                    const vb_pair_params *params, vb_pair_result *results, uint32_t *match_offsets, uint16_t *matches16,
                    uint64_t cap_matches, int *ticket);
int vb_pairs_wait(vb_ctx *ctx, int ticket, uint64_t *total_matches);
/* The same with device-resident inputs and outputs (arguments as vb_pairs_run_d; nothing is copied). An odd ticket's work
 * runs on an internal stream that first waits for everything queued on the context's stream at submission time;
 * vb_pairs_wait(ticket) returns once its kernels have finished (*total_matches is not written). Two consecutive submissions
 * overlap on the GPU — the counting and the small kernels of one beside the matcher of the other — which a sequence of
 * vb_pairs_run_d calls on one stream cannot. */
int vb_pairs_submit_d(vb_ctx *ctx, const float *pts_d, const uint8_t *desc_d, uint32_t nframes, uint32_t k, uint32_t bytes,
                      const vb_pair_params *params, vb_pair_result *results_d, int32_t *out_matches_d, int *ticket);
/* submit + wait */
int vb_pairs_run_compact(vb_ctx *ctx, const float *pts, const uint8_t *desc, uint32_t nframes, uint32_t k, uint32_t bytes,
                         const vb_pair_params *params, vb_pair_result *results, uint32_t *match_offsets, uint16_t *matches16,
                         uint64_t cap_matches, uint64_t *total_matches);
/* Pinned (page-locked, portable across contexts) host memory, or pinning of memory the caller already owns. */
int vb_host_alloc(size_t bytes, void **out);
int vb_host_free(void *p);
int vb_host_register(void *p, size_t bytes);
int vb_host_unregister(void *p);

/* ------------------------------------------------------------------------------------------------
 * Several GPUs of one box from ONE process: a vb_multi owns one host thread + one context + its own streams
 * per listed device. The nframes-1 pairs of a sequence are cut into contiguous ranges, one per GPU (one-frame halo);
 * every GPU downloads straight into the caller's arrays at its range's position, which is the host gather — there is no
 * exchange between GPUs, hence no collective. Pair i samples with std::mt19937(seed0 + i) on whichever GPU it runs, so
 * the output does not depend on the device list. Arguments as vb_pairs_submit, except that matches16 must have room for
 * (nframes-1) * k matches: the range of GPU g starts at match index first_pair(g) * k and is packed from there
 * (match_offsets[] holds the absolute start of every pair).
 * ---------------------------------------------------------------------------------------------- */
typedef struct vb_multi vb_multi;
int vb_multi_create(const int *devices, uint32_t ndev, vb_multi **out);
int vb_multi_destroy(vb_multi *m);
uint32_t vb_multi_device_count(const vb_multi *m);
int vb_multi_pairs_submit(vb_multi *m, const float *pts, const uint8_t *desc, uint32_t nframes, uint32_t k, uint32_t bytes,
                          const vb_pair_params *params, vb_pair_result *results, uint32_t *match_offsets,
                          uint16_t *matches16, uint64_t cap_matches, int *ticket);
int vb_multi_pairs_wait(vb_multi *m, int ticket, uint64_t *total_matches);
int vb_multi_pairs_run(vb_multi *m, const float *pts, const uint8_t *desc, uint32_t nframes, uint32_t k, uint32_t bytes,
                       const vb_pair_params *params, vb_pair_result *results, uint32_t *match_offsets, uint16_t *matches16,
                       uint64_t cap_matches, uint64_t *total_matches);

/* ------------------------------------------------------------------------------------------------
 * Search by projection — replaces the loop at reference src/vslam.cpp:129-161 together with orb_distance
 * (src/PointMap.cpp:36-46); the sole production caller of radius_search.
 *   tree            kd-tree of the current frame's keypoints (vb_kdtree_build on frame.points, src/Frame.cpp:76)
 *   map_points      [n][4] rows of pm.points (homogeneous), camera = c2 = K * R_t.rowRange(0,3), row-major 3x4
 *   width, height   image size of the in-image test (:141)
 *   frame_desc      [tree size][bytes] descriptors of the frame's keypoints
 *   map_point_ids   [tree size] frame.map_point_ids, in/out: >= 0 means already claimed (:151); claims are written back
 *   obs_offsets     [n+1] CSR over obs_desc: rows obs_offsets[i] .. obs_offsets[i+1] are the descriptors of map
 *                   point i's observations (pm.frames[frame_ids[i][j]].descriptors.row(frame_point_ids[i][j]))
 *   radius, dist_threshold   2 and DISTANCE_THRESHOLD = 64 in the reference (:149, :39)
 *   assign          [n] out: keypoint index claimed by map point i, or -1 (the caller appends frame.id / idx to
 *                   pm.frame_ids[i] / pm.frame_point_ids[i] for those, :155-156)
 *   proj_xy [n][2], in_view [n]   optional outputs of the projection step (may be NULL)
 * Result equals the reference's sequential first-unclaimed-wins loop exactly (map points in index order).
 * ---------------------------------------------------------------------------------------------- */
int vb_search_by_projection(vb_ctx *ctx, vb_tree *tree, const float *map_points, uint32_t n, const float *camera,
                            int width, int height, const uint8_t *frame_desc, uint32_t bytes, int32_t *map_point_ids,
                            const uint32_t *obs_offsets, const uint8_t *obs_desc, float radius, uint32_t dist_threshold,
                            int32_t *assign, float *proj_xy, uint8_t *in_view, uint32_t *n_claimed);

/* ------------------------------------------------------------------------------------------------
 * Downstream of F.
 * vb_extract_rt  — reference extract_Rt, src/helpers.cpp:3-35, for P fundamental matrices at once (F [P][9] row-major,
 *   K [9]): R [P][9], t [P][3]; E_out [P][9] (optional) receives E = K^T F K, which is bit-identical to the two
 *   cv::gemm calls of :4. The SVD of :7 is this build's fp64 Jacobi (as in the 8-point solve), not an OpenCV build's.
 * vb_triangulate — reference triangulate, src/helpers.cpp:37-80: p1, p2 [n][2] matched keypoints, c1, c2 3x4 cameras
 *   (row-major), points4 [n][4] = (X/w, Y/w, Z/w, 1) as at :71-74.
 * ---------------------------------------------------------------------------------------------- */
int vb_extract_rt(vb_ctx *ctx, const float *F, uint32_t P, const float *K, float *R, float *t, float *E_out);
int vb_triangulate(vb_ctx *ctx, const float *p1, const float *p2, uint32_t n, const float *c1, const float *c2,
                   float *points4);
/* triangulate + the reprojection gate that consumes it (reference src/vslam.cpp:186-251), fused: points4 [n][4] as above;
 * re1 / re2 [n] = squared reprojection error in camera 1 / 2 (`d.row(i).dot(d.row(i))`, :240, :242) of every row; inlier_idx
 * = the rows pushed to reprojection_inliers, in order: map_point_ids[i] <= 0 (or map_point_ids == NULL), re1 <= threshold_sq
 * and re2 <= threshold_sq (thresholdSq = 4, :53); *reproj_error = the f64 running sum of (re1 + re2) over them (:249).
 * Kept as written: the loop that makes the reprojections non-homogeneous (:201-211) only reaches the first ceil(n/3) rows,
 * and a NaN error passes the gate. Any output except n_inliers may be NULL. */
int vb_triangulate_gated(vb_ctx *ctx, const float *p1, const float *p2, uint32_t n, const float *c1, const float *c2,
                         const int32_t *map_point_ids, float threshold_sq, float *points4, float *re1, float *re2,
                         uint32_t *inlier_idx, uint32_t *n_inliers, double *reproj_error);

/* ------------------------------------------------------------------------------------------------
 * Timing hook for bench.py: CUDA-event time (ms) of the kernels of one class recorded on the context's
 * stream during the last call, keyed by name ("score", "hamming", "solve", ...). Returns <0 if unknown.
 * ---------------------------------------------------------------------------------------------- */
int vb_profile_enable(vb_ctx *ctx, int on);
/* Measurement hook (not on the product path): the rate at which this GPU retires back-to-back tcgen05.mma instructions of
 * one kind from one issuing thread per SM — operands resident in shared memory, accumulator in TMEM, no loads, no epilogue.
 * kind 0 = kind::mxf4.block_scale (e2m1, K = 64: the Hamming matcher's instruction), 1 = kind::f8f6f4 (e4m3, K = 32),
 * 2 = kind::f16 (bf16, K = 16); M = 128, N = n_cols. Launches reps + 1 times, reports the fastest timed launch (CUDA events
 * on the context's stream) and the flop count of one launch; bench.py prints flop / ms as roofline.peak_measured_*. */
int vb_probe_tensor_peak(vb_ctx *ctx, int kind, uint32_t n_cols, uint32_t iters, uint32_t reps, float *best_ms,
                         double *flop_per_launch);
float vb_profile_last_ms(vb_ctx *ctx, const char *kernel_class);

#ifdef __cplusplus
}
#endif
#endif /* VSLAM_B200_H */
