// kdtree.cu — 2-D k-d tree over keypoints: device build, batched exact 1-NN, batched radius search.
//
// Replaces construct_kdtree / nearest / radius_search (reference src/KDTree.cpp:3-35, :37-71, :73-101,
// :107-171; include/KDTree.h:13-80).
//
// Layout. The reference mallocs N pointer nodes and fills them in DFS pre-order (root[size++], :16), so
// a node's left child is always the next slot and its right child follows the len/2 nodes of the left
// subtree. That makes the pointers redundant: the tree here is three SoA arrays in HBM — x[], y[],
// idx[] in pre-order — and a traversal carries (slot, len) instead of a pointer.
//
// Build (k_kd_build, one CTA per tree). The reference recursively nth_element's each segment around
// position l + len/2 on alternating axes. Here both axis orders are sorted once (bitonic sort of
// (orderable(coord) << 32 | index) keys), and the two index lists are kept segment-aligned: at a level
// that splits on x, every segment's median is simply Lx[l + len/2], Lx needs no change, and Ly is
// stably partitioned per segment into [left | median | right] with one block-wide scan. All segments
// of a level are processed in the same passes, so the build is height x O(N) work with ~6 block
// barriers per level and no recursion. Working arrays live in shared memory when 33*N bytes fit
// (N <= ~6500, which covers the 5 000-keypoint configs) and in an HBM workspace otherwise.
//
// Queries (k_kd_nearest, k_kd_radius): one thread per query, explicit stack in shared memory laid out
// [depth][thread] (bank-conflict free). The visiting order is exactly the reference's recursion, so
// equidistant nearest-neighbour ties and the pre-order of radius results come out identical.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace vb {

constexpr int KD_BUILD_THREADS = 1024;
constexpr int KD_Q_THREADS = 128;
constexpr int KD_MAX_DEPTH = 32;

__device__ __forceinline__ uint32_t float_orderable(float f) {
    f = __fadd_rn(f, 0.0f);   // -0 -> +0 so that the two zeros tie, as they do under operator<
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// in-place bitonic sort of npad (power of two) 64-bit keys by the whole CTA
__device__ void bitonic_sort_u64(uint64_t *keys, uint32_t npad) {
    for (uint32_t k = 2; k <= npad; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < npad; i += blockDim.x) {
                const uint32_t ixj = i ^ j;
                if (ixj > i) {
                    const uint64_t a = keys[i], b = keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// sorts two key arrays with shared barriers (same stage structure)
__device__ void bitonic_sort2_u64(uint64_t *ka, uint64_t *kb, uint32_t npad) {
    for (uint32_t k = 2; k <= npad; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < npad; i += blockDim.x) {
                const uint32_t ixj = i ^ j;
                if (ixj > i) {
                    const bool up = (i & k) == 0;
                    uint64_t a = ka[i], b = ka[ixj];
                    if ((a > b) == up) { ka[i] = b; ka[ixj] = a; }
                    a = kb[i]; b = kb[ixj];
                    if ((a > b) == up) { kb[i] = b; kb[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// exclusive block scan of packed (left, median) flag counts over positions [0, n); out[n] = total
__device__ void block_scan_flags(const uint32_t *list, const uint8_t *side, const uint32_t *segL, const uint32_t *segR,
                                 uint32_t n, uint64_t *out, uint64_t *warp_tot) {
    const uint32_t per = (n + blockDim.x - 1) / blockDim.x;
    const uint32_t b = min(threadIdx.x * per, n), e = min(b + per, n);
    uint64_t local = 0;
    for (uint32_t p = b; p < e; p++) {
        if (segR[p] > segL[p]) {
            const uint8_t s = side[list[p]];
            local += (s == 0 ? 1ull : 0ull) + (s == 1 ? (1ull << 32) : 0ull);
        }
    }
    uint64_t incl = local;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    uint64_t woff = 0;
    for (int i = 0; i < w; i++) woff += warp_tot[i];
    uint64_t run = woff + incl - local;
    for (uint32_t p = b; p < e; p++) {
        out[p] = run;
        if (segR[p] > segL[p]) {
            const uint8_t s = side[list[p]];
            run += (s == 0 ? 1ull : 0ull) + (s == 1 ? (1ull << 32) : 0ull);
        }
    }
    __syncthreads();
}

struct KdBuildArgs {
    const float2 *pts;       // [ntrees][pts_stride]
    size_t pts_stride;
    uint32_t n, npad;
    float *out_x, *out_y;    // [ntrees][n] pre-order
    uint32_t *out_idx;
    size_t out_stride;
    uint8_t *ws;             // global workspace, ws_stride bytes per tree (lists; and keys when keys_global)
    size_t ws_stride;
    int lists_in_smem;       // working arrays in dynamic shared memory
    int keys_mode;           // 0: both key arrays in smem, 1: one at a time in smem, 2: global
    // sub-tree mode (top-down build of a large tree): CTA b builds the sub-tree of segment segs[b] = (start, len, slot):
    // its points are pts[start .. start + len), its nodes go to out[slot ...], its first level splits on axis0, and the
    // original point index of local point i is ord[start + i]
    const uint4 *segs;
    const uint32_t *ord;
    uint32_t axis0;
};

__global__ void __launch_bounds__(KD_BUILD_THREADS) k_kd_build(KdBuildArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint64_t warp_tot[KD_BUILD_THREADS / 32];
    const uint32_t tid = threadIdx.x;
    uint32_t n = a.n, npad = a.npad, seg_start = 0, seg_slot = 0;
    if (a.segs) {
        const uint4 sg = a.segs[blockIdx.x];
        seg_start = sg.x; n = sg.y; seg_slot = sg.z;
        npad = 1;
        while (npad < n) npad <<= 1;
    }
    const float2 *pts = a.segs ? a.pts + seg_start : a.pts + (size_t)blockIdx.x * a.pts_stride;
    uint8_t *gws = a.ws ? a.ws + (size_t)blockIdx.x * a.ws_stride : nullptr;

    // ---- carve the working arrays -----------------------------------------------------------
    uint8_t *base = a.lists_in_smem ? smem : gws;
    uint64_t *scan = reinterpret_cast<uint64_t *>(base);                  // [n+1]
    uint32_t *L0 = reinterpret_cast<uint32_t *>(scan + (n + 1));          // three list buffers
    uint32_t *L1 = L0 + n;
    uint32_t *L2 = L1 + n;
    uint32_t *segL = L2 + n;
    uint32_t *segR = segL + n;
    uint32_t *segB = segR + n;                                            // pre-order base of the segment
    uint8_t *side = reinterpret_cast<uint8_t *>(segB + n);
    const size_t lists_bytes = (size_t)(n + 1) * 8 + (size_t)n * 24 + ((n + 7) / 8) * 8;

    // ---- sort both axis orders ---------------------------------------------------------------
    uint32_t *Lx = L0, *Ly = L1, *spare = L2;
    if (a.keys_mode == 0) {
        // keys alias the (not yet used) working arrays in shared memory
        uint64_t *kx = reinterpret_cast<uint64_t *>(smem), *ky = kx + npad;
        for (uint32_t i = tid; i < npad; i += blockDim.x) {
            uint64_t vx = ~0ull, vy = ~0ull;
            if (i < n) {
                const float2 p = pts[i];
                vx = ((uint64_t)float_orderable(p.x) << 32) | i;
                vy = ((uint64_t)float_orderable(p.y) << 32) | i;
            }
            kx[i] = vx; ky[i] = vy;
        }
        __syncthreads();
        bitonic_sort2_u64(kx, ky, npad);
        // move out through registers (n <= 8 * blockDim in this mode) because the lists alias the keys
        uint32_t rx[8], ry[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint32_t i = tid + r * blockDim.x;
            rx[r] = (i < n) ? (uint32_t)kx[i] : 0;
            ry[r] = (i < n) ? (uint32_t)ky[i] : 0;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint32_t i = tid + r * blockDim.x;
            if (i < n) { Lx[i] = rx[r]; Ly[i] = ry[r]; }
        }
    } else {
        uint64_t *keys = (a.keys_mode == 1) ? reinterpret_cast<uint64_t *>(smem)
                                            : reinterpret_cast<uint64_t *>(gws + ((lists_bytes + 15) / 16) * 16);
        for (int axis = 0; axis < 2; axis++) {
            for (uint32_t i = tid; i < npad; i += blockDim.x) {
                uint64_t v = ~0ull;
                if (i < n) {
                    const float2 p = pts[i];
                    v = ((uint64_t)float_orderable(axis ? p.y : p.x) << 32) | i;
                }
                keys[i] = v;
            }
            __syncthreads();
            bitonic_sort_u64(keys, npad);
            uint32_t *dst = axis ? Ly : Lx;   // lists are global in this mode: no aliasing with smem keys
            for (uint32_t i = tid; i < n; i += blockDim.x) dst[i] = (uint32_t)keys[i];
            __syncthreads();
        }
    }
    for (uint32_t p = tid; p < n; p += blockDim.x) { segL[p] = 0; segR[p] = n; segB[p] = 0; }
    __syncthreads();

    float *ox = a.segs ? a.out_x + seg_slot : a.out_x + (size_t)blockIdx.x * a.out_stride;
    float *oy = a.segs ? a.out_y + seg_slot : a.out_y + (size_t)blockIdx.x * a.out_stride;
    uint32_t *oi = a.segs ? a.out_idx + seg_slot : a.out_idx + (size_t)blockIdx.x * a.out_stride;
    const uint32_t *omap = a.segs ? a.ord + seg_start : nullptr;

    // ---- one pass per tree level ---------------------------------------------------------------
    uint32_t height = 0;
    for (uint32_t t = n; t > 0; t >>= 1) height++;
    for (uint32_t level = 0; level < height; level++) {
        const uint32_t ax = (level + a.axis0) & 1u;
        uint32_t *A = ax ? Ly : Lx;    // list sorted along this level's axis
        uint32_t *Bl = ax ? Lx : Ly;   // the other list, to be partitioned
        // pass 1: classify every point of every live segment; medians become tree nodes
        for (uint32_t p = tid; p < n; p += blockDim.x) {
            const uint32_t l = segL[p], r = segR[p];
            if (r > l) {
                const uint32_t m = l + (r - l) / 2;   // src/KDTree.cpp:8
                const uint32_t pt = A[p];
                side[pt] = (p < m) ? 0 : (p == m) ? 1 : 2;
                if (p == m) {
                    const float2 v = pts[pt];
                    const uint32_t slot = segB[p];
                    ox[slot] = v.x; oy[slot] = v.y; oi[slot] = omap ? omap[pt] : pt;
                }
            }
        }
        __syncthreads();
        // pass 2: stable three-way partition of the other list, all segments at once
        block_scan_flags(Bl, side, segL, segR, n, scan, warp_tot);
        for (uint32_t p = tid; p < n; p += blockDim.x) {
            const uint32_t l = segL[p], r = segR[p];
            const uint32_t pt = Bl[p];
            uint32_t np = p;
            if (r > l) {
                const uint32_t m = l + (r - l) / 2;
                const uint64_t d = scan[p] - scan[l];
                const uint32_t cl = (uint32_t)d, cm = (uint32_t)(d >> 32);
                const uint8_t s = side[pt];
                np = (s == 0) ? l + cl : (s == 1) ? m : m + 1 + (p - l - cl - cm);
            }
            spare[np] = pt;
        }
        __syncthreads();
        // pass 3: shrink the segments
        for (uint32_t p = tid; p < n; p += blockDim.x) {
            const uint32_t l = segL[p], r = segR[p];
            if (r > l) {
                const uint32_t m = l + (r - l) / 2;
                if (p < m) { segR[p] = m; segB[p] += 1; }
                else if (p == m) { segL[p] = 0; segR[p] = 0; }
                else { segL[p] = m + 1; segB[p] += (m - l) + 1; }
            }
        }
        // rotate buffers: the partitioned copy becomes the current "other" list
        if (ax) { uint32_t *t = Lx; Lx = spare; spare = t; }
        else { uint32_t *t = Ly; Ly = spare; spare = t; }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// LPQ = lanes that share one query. LPQ = 1 is the product mapping (the batch dimension on the lanes). LPQ = 8 / 32 are
// the north star's "warp-per-query traversal with shared-memory stack staging" taken literally: the exact 1-NN descent of
// a 2-D tree is one dependent chain (load node -> compare -> push / pop), so the group's leader lane walks it with the
// stack in shared memory and the other lanes have nothing to do. They exist so that the choice is a measurement
// (option kd_lanes_per_query; bench.py kdtree.lanes_per_query_ab), not an argument.
template <int LPQ>
__global__ void __launch_bounds__(KD_Q_THREADS) k_kd_nearest(const float *__restrict__ tx, const float *__restrict__ ty,
                                                             const uint32_t *__restrict__ tidx, uint32_t n,
                                                             const float2 *__restrict__ q, uint32_t nq, float max_d2,
                                                             float2 *__restrict__ out_pt, int32_t *__restrict__ out_idx,
                                                             float *__restrict__ out_d2) {
    __shared__ uint2 stack[KD_MAX_DEPTH][KD_Q_THREADS / LPQ];
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) / LPQ, t = threadIdx.x / LPQ;
    if (i >= nq || (threadIdx.x % LPQ) != 0) return;
    const float2 qp = q[i];
    float best = max_d2;
    int best_slot = -1;
    int sp = 0;
    // frame = (slot, len | axis << 30 | state << 31)
    if (n > 0) stack[sp++][t] = make_uint2(0u, n);
    while (sp > 0) {
        const uint2 f = stack[sp - 1][t];
        const uint32_t slot = f.x, len = f.y & 0x3fffffffu, axis = (f.y >> 30) & 1u, state = f.y >> 31;
        const float x = __ldg(tx + slot), y = __ldg(ty + slot);
        const float split = axis ? __fsub_rn(qp.y, y) : __fsub_rn(qp.x, x);   // :51
        const uint32_t llen = len >> 1, rlen = len - llen - 1;
        const bool go_left = split < 0.0f;                                    // :54
        if (state == 0) {
            stack[sp - 1][t].y = f.y | 0x80000000u;
            const uint32_t cs = go_left ? slot + 1 : slot + 1 + llen, cl = go_left ? llen : rlen;
            if (cl > 0) { stack[sp][t] = make_uint2(cs, cl | ((axis ^ 1u) << 30)); sp++; }
        } else {
            sp--;
            const float dx = __fsub_rn(x, qp.x), dy = __fsub_rn(y, qp.y);     // :62-63
            const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            if (d2 < best) { best = d2; best_slot = (int)slot; }              // :64-67
            const uint32_t cs = go_left ? slot + 1 + llen : slot + 1, cl = go_left ? rlen : llen;
            if (cl > 0 && __fmul_rn(split, split) < best) {                   // :68-70
                stack[sp][t] = make_uint2(cs, cl | ((axis ^ 1u) << 30));
                sp++;
            }
        }
    }
    if (out_pt) out_pt[i] = best_slot >= 0 ? make_float2(tx[best_slot], ty[best_slot]) : make_float2(0.f, 0.f);
    if (out_idx) out_idx[i] = best_slot >= 0 ? (int32_t)tidx[best_slot] : -1;
    if (out_d2) out_d2[i] = best;
}

// k nearest neighbours, k <= KD_KNN_MAX: the traversal of k_kd_nearest with the k-th best distance as the bound and an
// ascending candidate list per query in shared memory ([k][thread], conflict-free). Build-defined (the reference only has
// commented-out declarations, include/KDTree.h:39-42): first visited first among equal distances, like `nearest`.
constexpr int KD_KNN_MAX = 32;
constexpr int KD_KNN_THREADS = 64;
__global__ void __launch_bounds__(KD_KNN_THREADS) k_kd_knn(const float *__restrict__ tx, const float *__restrict__ ty,
                                                           const uint32_t *__restrict__ tidx, uint32_t n,
                                                           const float2 *__restrict__ q, uint32_t nq, uint32_t k, float max_d2,
                                                           int32_t *__restrict__ out_idx, float *__restrict__ out_d2,
                                                           uint32_t *__restrict__ out_cnt) {
    __shared__ uint2 stack[KD_MAX_DEPTH][KD_KNN_THREADS];
    __shared__ float ld2[KD_KNN_MAX][KD_KNN_THREADS];
    __shared__ uint32_t lslot[KD_KNN_MAX][KD_KNN_THREADS];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, t = threadIdx.x;
    if (i >= nq) return;
    const float2 qp = q[i];
    uint32_t cnt = 0;
    float bound = max_d2;
    int sp = 0;
    if (n > 0) stack[sp++][t] = make_uint2(0u, n);
    while (sp > 0) {
        const uint2 f = stack[sp - 1][t];
        const uint32_t slot = f.x, len = f.y & 0x3fffffffu, axis = (f.y >> 30) & 1u, state = f.y >> 31;
        const float x = __ldg(tx + slot), y = __ldg(ty + slot);
        const float split = axis ? __fsub_rn(qp.y, y) : __fsub_rn(qp.x, x);
        const uint32_t llen = len >> 1, rlen = len - llen - 1;
        const bool go_left = split < 0.0f;
        if (state == 0) {
            stack[sp - 1][t].y = f.y | 0x80000000u;
            const uint32_t cs = go_left ? slot + 1 : slot + 1 + llen, cl = go_left ? llen : rlen;
            if (cl > 0) { stack[sp][t] = make_uint2(cs, cl | ((axis ^ 1u) << 30)); sp++; }
        } else {
            sp--;
            const float dx = __fsub_rn(x, qp.x), dy = __fsub_rn(y, qp.y);
            const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            if (d2 < bound) {
                uint32_t j = cnt < k ? cnt : k - 1;
                while (j > 0 && ld2[j - 1][t] > d2) { ld2[j][t] = ld2[j - 1][t]; lslot[j][t] = lslot[j - 1][t]; j--; }
                ld2[j][t] = d2; lslot[j][t] = slot;
                if (cnt < k) cnt++;
                bound = cnt < k ? max_d2 : ld2[k - 1][t];
            }
            const uint32_t cs = go_left ? slot + 1 + llen : slot + 1, cl = go_left ? rlen : llen;
            if (cl > 0 && __fmul_rn(split, split) < bound) {
                stack[sp][t] = make_uint2(cs, cl | ((axis ^ 1u) << 30));
                sp++;
            }
        }
    }
    for (uint32_t j = 0; j < k; j++) {
        out_idx[(size_t)i * k + j] = j < cnt ? (int32_t)tidx[lslot[j][t]] : -1;
        out_d2[(size_t)i * k + j] = j < cnt ? ld2[j][t] : max_d2;
    }
    if (out_cnt) out_cnt[i] = cnt;
}

template <bool FILL>
__global__ void __launch_bounds__(KD_Q_THREADS) k_kd_radius(const float *__restrict__ tx, const float *__restrict__ ty,
                                                            const uint32_t *__restrict__ tidx, uint32_t n,
                                                            const float2 *__restrict__ q, uint32_t nq, float radius,
                                                            uint32_t *__restrict__ counts,
                                                            const uint32_t *__restrict__ offsets,
                                                            uint32_t *__restrict__ out_idx, uint32_t out_cap = 0xffffffffu) {
    __shared__ uint2 stack[KD_MAX_DEPTH][KD_Q_THREADS];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, t = threadIdx.x;
    if (i >= nq) return;
    const float2 qp = q[i];
    const float r2 = __fmul_rn(radius, radius);   // SQ(radius), :74
    uint32_t cnt = 0;
    const uint32_t off = FILL ? offsets[i] : 0;
    int sp = 0;
    if (n > 0) stack[sp++][t] = make_uint2(0u, n);
    while (sp > 0) {
        const uint2 f = stack[--sp][t];
        uint32_t slot = f.x, len = f.y & 0x3fffffffu, axis = (f.y >> 30) & 1u;
        while (len > 0) {
            const float x = __ldg(tx + slot), y = __ldg(ty + slot);
            const float split = axis ? __fsub_rn(qp.y, y) : __fsub_rn(qp.x, x);
            const uint32_t llen = len >> 1, rlen = len - llen - 1;
            const float as = (split > 0.0f) ? split : -split;   // ABS macro, include/KDTree.h:10
            if (as <= radius) {                                 // :88
                const float dx = __fsub_rn(qp.x, x), dy = __fsub_rn(qp.y, y);
                const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                if (d2 < r2) {                                  // :91
                    if (FILL && off + cnt < out_cap) out_idx[off + cnt] = tidx[slot];
                    cnt++;
                }
                if (rlen > 0) { stack[sp][t] = make_uint2(slot + 1 + llen, rlen | ((axis ^ 1u) << 30)); sp++; }
                slot = slot + 1; len = llen;
            } else if (split < 0.0f) {
                slot = slot + 1; len = llen;
            } else {
                slot = slot + 1 + llen; len = rlen;
            }
            axis ^= 1u;
        }
    }
    if (!FILL) counts[i] = cnt;
}

// single-CTA exclusive scan: out[0..n] (out[n] = total, also stored to *total as 64-bit)
__global__ void __launch_bounds__(1024) k_scan_u32(const uint32_t *__restrict__ in, uint32_t n, uint32_t *__restrict__ out,
                                                   unsigned long long *__restrict__ total) {
    __shared__ unsigned long long wt[32];
    const uint32_t per = (n + blockDim.x - 1) / blockDim.x;
    const uint32_t b = min(threadIdx.x * per, n), e = min(b + per, n);
    unsigned long long local = 0;
    for (uint32_t p = b; p < e; p++) local += in[p];
    unsigned long long incl = local;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wt[w] = incl;
    __syncthreads();
    unsigned long long woff = 0, all = 0;
    for (int i = 0; i < 32; i++) {
        if (i < w) woff += wt[i];
        all += wt[i];
    }
    unsigned long long run = woff + incl - local;
    for (uint32_t p = b; p < e; p++) {
        out[p] = (uint32_t)run;
        run += in[p];
    }
    if (threadIdx.x == 0) {
        out[n] = (uint32_t)all;
        *total = all;
    }
}

// Large query batches: the same exclusive scan in three coalesced passes (per-block scan of 4 096 counts, scan of the block
// sums, add). The single-CTA version above walks n / 1 024 consecutive elements per thread — fine for a frame's worth of
// queries, 1.3 ms of uncoalesced loads at 2^20.
constexpr uint32_t SCAN_BLOCK = 4096;
__global__ void __launch_bounds__(1024) k_scan_blocks(const uint32_t *__restrict__ in, uint32_t n, uint32_t *__restrict__ out,
                                                      uint32_t *__restrict__ block_sums) {
    __shared__ uint32_t wt[32];
    const uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = (base + k < n) ? in[base + k] : 0u;
    const uint32_t local = v[0] + v[1] + v[2] + v[3];
    uint32_t incl = local;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wt[w] = incl;
    __syncthreads();
    uint32_t woff = 0, all = 0;
    for (int i = 0; i < 32; i++) {
        if (i < w) woff += wt[i];
        all += wt[i];
    }
    uint32_t run = woff + incl - local;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = all;
}
__global__ void __launch_bounds__(1024) k_scan_add(uint32_t *__restrict__ out, uint32_t n, const uint32_t *__restrict__ block_offs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += block_offs[i / SCAN_BLOCK];
}

// out[0..n] = exclusive scan of in[0..n) (out[n] = total, also as 64 bit in *total). scratch: 2 * (n / SCAN_BLOCK + 2) words.
static void scan_counts(cudaStream_t st, const uint32_t *in, uint32_t n, uint32_t *out, unsigned long long *total, uint32_t *scratch,
                        uint64_t *launches) {
    if (n <= 4 * SCAN_BLOCK || scratch == nullptr) {
        k_scan_u32<<<1, 1024, 0, st>>>(in, n, out, total);
        *launches += 1;
        return;
    }
    const uint32_t nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;   // <= 2^20 blocks for n < 2^32; the sums' own scan stays single-CTA
    uint32_t *sums = scratch, *offs = scratch + nb + 1;
    k_scan_blocks<<<nb, 1024, 0, st>>>(in, n, out, sums);
    k_scan_u32<<<1, 1024, 0, st>>>(sums, nb, offs, total);
    k_scan_add<<<(n + 1023) / 1024, 1024, 0, st>>>(out, n, offs);
    cudaMemcpyAsync(out + n, offs + nb, 4, cudaMemcpyDeviceToDevice, st);   // out[n] = total (fits 32 bits or the caller rejects it)
    *launches += 3;
}

// ---- top-down levels of a LARGE tree --------------------------------------------------------------------------------
// k_kd_build sorts a whole tree inside one CTA; past ~6 000 points its working set leaves shared memory and a single CTA
// grinds through an HBM-resident bitonic sort (3.3 ms for 20 000 points — slower than one host core). The shape of the
// tree does not depend on the data (the median sits at position l + len/2, src/KDTree.cpp:8), so the segments of every
// level are known up front: the top levels are peeled off with one CTA per segment — radix-select the median of the level's
// coordinate, emit it as the node, stably partition the segment's index list around it — until the segments fit shared
// memory, and k_kd_build then builds all the sub-trees in one launch (sub-tree mode). Ties on the coordinate fall to the
// lower original index, as everywhere: the index list starts ascending and every partition is stable.
constexpr uint32_t KD_SPLIT_THREADS = 1024;
constexpr uint32_t KD_SMEM_MAX_POINTS = 6144;

__global__ void __launch_bounds__(KD_SPLIT_THREADS) k_kd_split(const float2 *__restrict__ pts, const uint32_t *__restrict__ ord_in,
                                                               uint32_t *__restrict__ ord_out, const uint4 *__restrict__ segs,
                                                               uint32_t axis, float *__restrict__ ox, float *__restrict__ oy,
                                                               uint32_t *__restrict__ oi) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_rank, s_less, s_pick;
    __shared__ uint32_t wsum[32][3];
    __shared__ uint32_t s_run[3];
    const uint4 sg = segs[blockIdx.x];
    const uint32_t l = sg.x, len = sg.y, slot = sg.z, tid = threadIdx.x;
    if (len == 0) return;
    const uint32_t k = len / 2;   // rank of the median in (coordinate, position) order; position order = original index order
    auto coord = [&](uint32_t i) {
        const float2 v = pts[ord_in[l + i]];
        return float_orderable(axis ? v.y : v.x);
    };
    // -- radix select of the k-th smallest coordinate (4 passes of 8 bits, most significant first) --
    if (tid == 0) { s_prefix = 0; s_rank = k; }
    for (int pass = 0; pass < 4; pass++) {
        const int shift = 24 - 8 * pass;
        for (uint32_t b = tid; b < 256; b += blockDim.x) hist[b] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix, himask = pass ? (0xffffffffu << (shift + 8)) : 0u;
        for (uint32_t i = tid; i < len; i += blockDim.x) {
            const uint32_t c = coord(i);
            if ((c & himask) == prefix) atomicAdd(&hist[(c >> shift) & 0xffu], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t r = s_rank, b = 0;
            for (; b < 256; b++) {
                if (r < hist[b]) break;
                r -= hist[b];
            }
            s_prefix = prefix | (b << shift);
            s_rank = r;
        }
        __syncthreads();
    }
    const uint32_t cstar = s_prefix;        // the median's coordinate
    const uint32_t eq_rank = s_rank;        // ... it is the eq_rank-th (0-based, in position order) of the points with that coordinate
    // -- stable three-way partition: [l, l + k) <- smaller, l + k <- the median, (l + k, l + len) <- larger --
    if (tid < 3) s_run[tid] = 0;            // running counts: left, equal-to-cstar seen, right
    __syncthreads();
    const int lane = tid & 31, w = tid >> 5;
    for (uint32_t i0 = 0; i0 < len; i0 += blockDim.x) {
        const uint32_t i = i0 + tid;
        uint32_t c = 0, id = 0;
        int isl = 0, ise = 0, isr = 0;
        if (i < len) {
            id = ord_in[l + i];
            const float2 v = pts[id];
            c = float_orderable(axis ? v.y : v.x);
            isl = c < cstar; ise = c == cstar; isr = c > cstar;
        }
        const unsigned bl = __ballot_sync(0xffffffffu, isl), be = __ballot_sync(0xffffffffu, ise), br = __ballot_sync(0xffffffffu, isr);
        if (lane == 0) { wsum[w][0] = __popc(bl); wsum[w][1] = __popc(be); wsum[w][2] = __popc(br); }
        __syncthreads();
        uint32_t pl = s_run[0], pe = s_run[1], pr = s_run[2], tl = 0, te = 0, tr = 0;
        for (int q = 0; q < 32; q++) {
            if (q < w) { pl += wsum[q][0]; pe += wsum[q][1]; pr += wsum[q][2]; }
            tl += wsum[q][0]; te += wsum[q][1]; tr += wsum[q][2];
        }
        const unsigned below = (1u << lane) - 1u;
        pl += __popc(bl & below); pe += __popc(be & below); pr += __popc(br & below);
        if (i < len) {
            // equal coordinates: the first eq_rank of them (position order) are "smaller", the next one is the median
            // number of equal-coordinate points that end up left of a given equal point = min(its equal-rank, eq_rank)
            uint32_t dst;
            if (isl) dst = l + pl + min(pe, eq_rank);
            else if (ise && pe < eq_rank) dst = l + pl + pe;
            else if (ise && pe == eq_rank) dst = l + k;
            else if (ise) dst = l + k + 1 + pr + (pe - eq_rank - 1);
            else dst = l + k + 1 + pr + (pe > eq_rank ? pe - eq_rank - 1 : 0);
            ord_out[dst] = id;
            if (ise && pe == eq_rank) {
                const float2 v = pts[id];
                ox[slot] = v.x; oy[slot] = v.y; oi[slot] = id;
            }
        }
        __syncthreads();
        if (tid == 0) { s_run[0] += tl; s_run[1] += te; s_run[2] += tr; }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_kd_iota(uint32_t *p, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}
__global__ void __launch_bounds__(256) k_kd_gather(const float2 *__restrict__ pts, const uint32_t *__restrict__ ord, uint32_t n,
                                                   float2 *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = pts[ord[i]];
}

static uint32_t next_pow2(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

// Builds ntrees trees of n points each. out arrays are [ntrees][out_stride].
// One large tree, top-down (see k_kd_split). Workspace (WS_KD_TOP): two index lists, the gathered points, the segment tables.
static int kd_build_topdown(vb_ctx *ctx, const float2 *pts_d, uint32_t n, float *ox, float *oy, uint32_t *oi) {
    uint32_t levels = 0;
    while (((n >> levels) + 1) > KD_SMEM_MAX_POINTS) levels++;   // segment lengths at level L are <= ceil(n / 2^L)
    // the segments of every level (data-independent): (start, len, pre-order slot)
    std::vector<std::vector<uint4>> segs(levels + 1);
    segs[0].push_back(make_uint4(0u, n, 0u, 0u));
    for (uint32_t L = 0; L < levels; L++)
        for (const uint4 &sg : segs[L]) {
            const uint32_t l = sg.x, len = sg.y, slot = sg.z, k = len / 2;
            segs[L + 1].push_back(make_uint4(l, k, slot + 1, 0u));
            segs[L + 1].push_back(make_uint4(l + k + 1, len - k - 1, slot + 1 + k, 0u));
        }
    size_t nseg_total = 0;
    for (auto &v : segs) nseg_total += v.size();
    const size_t off_ord = 0, off_pts = (size_t)n * 8, off_segs = off_pts + (size_t)n * 8;
    int rc;
    if ((rc = ctx->ws_ensure(WS_KD_TOP, off_segs + nseg_total * sizeof(uint4)))) return rc;
    uint8_t *base = ctx->ws[WS_KD_TOP].as<uint8_t>();
    uint32_t *ordA = reinterpret_cast<uint32_t *>(base + off_ord), *ordB = ordA + n;
    float2 *gpts = reinterpret_cast<float2 *>(base + off_pts);
    uint4 *segs_d = reinterpret_cast<uint4 *>(base + off_segs);
    std::vector<uint4> &flat = ctx->kd_segs_host;   // kept alive by the context: the upload below is asynchronous
    flat.clear();
    std::vector<size_t> level_off;
    for (auto &v : segs) { level_off.push_back(flat.size()); flat.insert(flat.end(), v.begin(), v.end()); }
    VB_CUDA(cudaStreamSynchronize(ctx->stream));   // an earlier build may still be reading the table (rare path: large trees)
    VB_CUDA(cudaMemcpyAsync(segs_d, flat.data(), flat.size() * sizeof(uint4), cudaMemcpyHostToDevice, ctx->stream));
    ctx->prof_begin("kd_build");
    k_kd_iota<<<div_up(n, 256), 256, 0, ctx->stream>>>(ordA, n);
    uint32_t *cur = ordA, *nxt = ordB;
    for (uint32_t L = 0; L < levels; L++) {
        // a level's partitions only fill the children's ranges; the medians' own slots in the next list are never read
        k_kd_split<<<(unsigned)segs[L].size(), KD_SPLIT_THREADS, 0, ctx->stream>>>(pts_d, cur, nxt, segs_d + level_off[L], L & 1u, ox, oy, oi);
        std::swap(cur, nxt);
    }
    k_kd_gather<<<div_up(n, 256), 256, 0, ctx->stream>>>(pts_d, cur, n, gpts);
    KdBuildArgs a;
    memset(&a, 0, sizeof(a));
    a.pts = gpts; a.n = (n >> levels) + 1; a.npad = 1;
    while (a.npad < a.n) a.npad <<= 1;
    a.out_x = ox; a.out_y = oy; a.out_idx = oi;
    a.lists_in_smem = 1; a.keys_mode = 0;
    a.segs = segs_d + level_off[levels]; a.ord = cur; a.axis0 = levels & 1u;
    const size_t lists_bytes = (size_t)(a.n + 1) * 8 + (size_t)a.n * 24 + ((a.n + 7) / 8) * 8;
    const size_t keys2 = (size_t)a.npad * 16;
    const size_t smem = lists_bytes > keys2 ? lists_bytes : keys2;
    VB_CUDA(cudaFuncSetAttribute(k_kd_build, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k_kd_build<<<(unsigned)segs[levels].size(), KD_BUILD_THREADS, smem, ctx->stream>>>(a);
    ctx->prof_end("kd_build");
    ctx->launches += 3 + levels;
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

int kd_build_launch(vb_ctx *ctx, const float2 *pts_d, size_t pts_stride, uint32_t ntrees, uint32_t n, float *ox, float *oy,
                    uint32_t *oi, size_t out_stride) {
    if (n == 0 || ntrees == 0) return VB_OK;
    if (n > KD_SMEM_MAX_POINTS) {   // large trees: top-down, many CTAs per tree
        for (uint32_t t = 0; t < ntrees; t++) {
            int rc = kd_build_topdown(ctx, pts_d + (size_t)t * pts_stride, n, ox + (size_t)t * out_stride, oy + (size_t)t * out_stride,
                                      oi + (size_t)t * out_stride);
            if (rc) return rc;
        }
        return VB_OK;
    }
    KdBuildArgs a;
    memset(&a, 0, sizeof(a));
    a.pts = pts_d; a.pts_stride = pts_stride; a.n = n; a.npad = next_pow2(n);
    a.out_x = ox; a.out_y = oy; a.out_idx = oi; a.out_stride = out_stride;
    const size_t lists_bytes = (size_t)(n + 1) * 8 + (size_t)n * 24 + ((n + 7) / 8) * 8;
    const size_t keys2 = (size_t)a.npad * 16, keys1 = (size_t)a.npad * 8;
    const size_t smem_max = 200 * 1024;
    size_t smem = 0, ws_per_tree = 0;
    if (lists_bytes <= smem_max && keys2 <= smem_max && n <= 8u * KD_BUILD_THREADS) {
        a.lists_in_smem = 1; a.keys_mode = 0;
        smem = lists_bytes > keys2 ? lists_bytes : keys2;
    } else {
        a.lists_in_smem = 0;
        ws_per_tree = ((lists_bytes + 15) / 16) * 16;
        if (keys1 <= smem_max) { a.keys_mode = 1; smem = keys1; }
        else { a.keys_mode = 2; ws_per_tree += keys1; }
    }
    if (ws_per_tree) {
        int rc = ctx->ws_ensure(WS_SCAN, ws_per_tree * ntrees);
        if (rc) return rc;
        a.ws = ctx->ws[WS_SCAN].as<uint8_t>();
        a.ws_stride = ws_per_tree;
    }
    VB_CUDA(cudaFuncSetAttribute(k_kd_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    ctx->prof_begin("kd_build");
    k_kd_build<<<ntrees, KD_BUILD_THREADS, smem, ctx->stream>>>(a);
    ctx->prof_end("kd_build");
    ctx->launches++;
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

static int tree_alloc(vb_ctx *ctx, uint32_t n, vb_tree **out) {
    vb_tree *t = new vb_tree();
    t->ctx = ctx; t->n = n;
    uint32_t h = 0;
    for (uint32_t v = n; v > 0; v >>= 1) h++;
    t->height = h;   // floor(log2 n) + 1
    if (n) {
        cudaError_t e = cudaMallocAsync(&t->block, (size_t)n * 12, ctx->stream);
        if (e != cudaSuccess) {
            set_error("cudaMallocAsync(%zu) -> %s", (size_t)n * 12, cudaGetErrorString(e));
            delete t;
            return VB_ERR_CUDA;
        }
        t->x = reinterpret_cast<float *>(t->block);
        t->y = t->x + n;
        t->idx = reinterpret_cast<uint32_t *>(t->y + n);
    }
    *out = t;
    return VB_OK;
}

}  // namespace vb

using namespace vb;

namespace vb {
// The same without any host synchronisation: the hit array has room for `cap` entries (hits beyond it are dropped, the
// offsets stay exact) and the total is left in WS_MISC[0] (64-bit) for the caller to read when it synchronises anyway.
int kd_radius_ws_async(vb_tree *t, const float2 *q_d, uint32_t nq, float radius, uint32_t cap) {
    vb_ctx *ctx = t->ctx;
    int rc;
    if ((rc = ctx->ws_ensure(WS_OFFS, (size_t)(nq + 1) * 4))) return rc;
    if ((rc = ctx->ws_ensure(WS_OUT0, (size_t)(nq + 1) * 4))) return rc;
    if ((rc = ctx->ws_ensure(WS_OUT1, (size_t)(cap ? cap : 1) * 4))) return rc;
    if ((rc = ctx->ws_ensure(WS_MISC, 64))) return rc;
    uint32_t *counts = ctx->ws[WS_OFFS].as<uint32_t>();
    uint32_t *offs = ctx->ws[WS_OUT0].as<uint32_t>();
    unsigned long long *total_d = ctx->ws[WS_MISC].as<unsigned long long>();
    ctx->prof_begin("kd_radius");
    if (nq)
        k_kd_radius<false><<<div_up(nq, KD_Q_THREADS), KD_Q_THREADS, 0, ctx->stream>>>(t->x, t->y, t->idx, t->n, q_d, nq, radius,
                                                                                      counts, nullptr, nullptr);
    if ((rc = ctx->ws_ensure(WS_SCAN2, (size_t)2 * (nq / SCAN_BLOCK + 3) * 4))) return rc;
    scan_counts(ctx->stream, counts, nq, offs, total_d, ctx->ws[WS_SCAN2].as<uint32_t>(), &ctx->launches);
    if (nq)
        k_kd_radius<true><<<div_up(nq, KD_Q_THREADS), KD_Q_THREADS, 0, ctx->stream>>>(t->x, t->y, t->idx, t->n, q_d, nq, radius,
                                                                                     nullptr, offs, ctx->ws[WS_OUT1].as<uint32_t>(), cap);
    ctx->prof_end("kd_radius");
    ctx->launches += nq ? 2 : 0;
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

// Internal CSR radius search for callers inside the library (search by projection): offsets land in WS_OUT0
// ([nq+1]), hit indices (original point indices, DFS pre-order per query) in WS_OUT1.
int kd_radius_ws(vb_tree *t, const float2 *q_d, uint32_t nq, float radius, uint64_t *total_out) {
    vb_ctx *ctx = t->ctx;
    int rc;
    if ((rc = ctx->ws_ensure(WS_OFFS, (size_t)(nq + 1) * 4))) return rc;
    if ((rc = ctx->ws_ensure(WS_OUT0, (size_t)(nq + 1) * 4))) return rc;
    if ((rc = ctx->ws_ensure(WS_MISC, 64))) return rc;
    uint32_t *counts = ctx->ws[WS_OFFS].as<uint32_t>();
    uint32_t *offs = ctx->ws[WS_OUT0].as<uint32_t>();
    unsigned long long *total_d = ctx->ws[WS_MISC].as<unsigned long long>();
    ctx->prof_begin("kd_radius");
    if (nq)
        k_kd_radius<false><<<div_up(nq, KD_Q_THREADS), KD_Q_THREADS, 0, ctx->stream>>>(t->x, t->y, t->idx, t->n, q_d, nq, radius,
                                                                                      counts, nullptr, nullptr);
    if ((rc = ctx->ws_ensure(WS_SCAN2, (size_t)2 * (nq / SCAN_BLOCK + 3) * 4))) return rc;
    scan_counts(ctx->stream, counts, nq, offs, total_d, ctx->ws[WS_SCAN2].as<uint32_t>(), &ctx->launches);
    ctx->launches += nq ? 1 : 0;
    VB_CUDA(cudaGetLastError());
    unsigned long long total = 0;
    VB_CUDA(cudaMemcpyAsync(&total, total_d, 8, cudaMemcpyDeviceToHost, ctx->stream));
    VB_CUDA(cudaStreamSynchronize(ctx->stream));
    *total_out = total;
    if (total > 0xffffffffull) { set_error("radius result too large for 32-bit CSR offsets"); return VB_ERR_CAPACITY; }
    if ((rc = ctx->ws_ensure(WS_OUT1, (size_t)(total ? total : 1) * 4))) return rc;
    if (total && nq) {
        k_kd_radius<true><<<div_up(nq, KD_Q_THREADS), KD_Q_THREADS, 0, ctx->stream>>>(t->x, t->y, t->idx, t->n, q_d, nq, radius,
                                                                                     nullptr, offs, ctx->ws[WS_OUT1].as<uint32_t>());
        ctx->launches++;
        VB_CUDA(cudaGetLastError());
    }
    ctx->prof_end("kd_radius");
    return VB_OK;
}
}  // namespace vb

extern "C" {

int vb_kdtree_build_d(vb_ctx *ctx, const float *pts_d, uint32_t n, vb_tree **out) {
    VB_REQUIRE(ctx && out && (pts_d || n == 0), VB_ERR_INVALID, "NULL argument");
    VB_REQUIRE(n < (1u << 30), VB_ERR_INVALID, "too many points");
    VB_CUDA(cudaSetDevice(ctx->device));
    vb_tree *t = nullptr;
    int rc = tree_alloc(ctx, n, &t);
    if (rc) return rc;
    rc = kd_build_launch(ctx, reinterpret_cast<const float2 *>(pts_d), 0, 1, n, t->x, t->y, t->idx, 0);
    if (rc) { vb_kdtree_free(t); return rc; }
    *out = t;
    return VB_OK;
}

int vb_kdtree_build_batch_d(vb_ctx *ctx, const float *pts_d, uint32_t ntrees, uint32_t n, vb_tree **out) {
    VB_REQUIRE(ctx && out && (pts_d || n == 0 || ntrees == 0), VB_ERR_INVALID, "NULL argument");
    VB_REQUIRE(n < (1u << 30), VB_ERR_INVALID, "too many points");
    if (ntrees == 0) return VB_OK;
    VB_CUDA(cudaSetDevice(ctx->device));
    // one allocation backs every tree ([tree][x | y | idx]); tree 0 owns it, the others are views
    vb_tree *first = nullptr;
    int rc = tree_alloc(ctx, n * ntrees, &first);
    if (rc) return rc;
    float *base = reinterpret_cast<float *>(first->block);
    rc = kd_build_launch(ctx, reinterpret_cast<const float2 *>(pts_d), n, ntrees, n, base, base + n,
                         reinterpret_cast<uint32_t *>(base + 2 * (size_t)n), 3 * (size_t)n);
    if (rc) { vb_kdtree_free(first); return rc; }
    uint32_t height = 0;
    for (uint32_t v = n; v; v >>= 1) height++;
    for (uint32_t i = 0; i < ntrees; i++) {
        vb_tree *t = i ? new vb_tree() : first;
        t->ctx = ctx;
        t->n = n;
        t->height = height;
        t->x = base + (size_t)i * 3 * n;
        t->y = t->x + n;
        t->idx = reinterpret_cast<uint32_t *>(t->y + n);
        if (i) t->block = nullptr;
        out[i] = t;
    }
    return VB_OK;
}

int vb_kdtree_free_batch(vb_tree **trees, uint32_t ntrees) {
    if (!trees) return VB_OK;
    for (uint32_t i = ntrees; i-- > 0;) vb_kdtree_free(trees[i]);   // views first, the owner (tree 0) last
    return VB_OK;
}

int vb_kdtree_build(vb_ctx *ctx, const float *pts, uint32_t n, vb_tree **out) {
    VB_REQUIRE(ctx && out && (pts || n == 0), VB_ERR_INVALID, "NULL argument");
    VB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ctx->ws_ensure(WS_PTS, (size_t)(n ? n : 1) * 8))) return rc;
    if (n) VB_CUDA(cudaMemcpyAsync(ctx->ws[WS_PTS].p, pts, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = vb_kdtree_build_d(ctx, ctx->ws[WS_PTS].as<float>(), n, out))) return rc;
    VB_CUDA(cudaStreamSynchronize(ctx->stream));   // the caller may free or reuse pts immediately
    return VB_OK;
}

int vb_kdtree_import(vb_ctx *ctx, const float *pts_preorder, const uint32_t *idx_preorder, uint32_t n, vb_tree **out) {
    VB_REQUIRE(ctx && out && (pts_preorder || n == 0), VB_ERR_INVALID, "NULL argument");
    VB_REQUIRE(n < (1u << 30), VB_ERR_INVALID, "too many points");
    VB_CUDA(cudaSetDevice(ctx->device));
    vb_tree *t = nullptr;
    int rc = tree_alloc(ctx, n, &t);
    if (rc) return rc;
    if (n) {
        std::vector<float> soa((size_t)n * 3);
        uint32_t *ii = reinterpret_cast<uint32_t *>(soa.data() + (size_t)n * 2);
        for (uint32_t i = 0; i < n; i++) {
            soa[i] = pts_preorder[2 * i];
            soa[n + i] = pts_preorder[2 * i + 1];
            ii[i] = idx_preorder ? idx_preorder[i] : i;
        }
        cudaError_t e = cudaMemcpyAsync(t->block, soa.data(), (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            set_error("vb_kdtree_import: %s", cudaGetErrorString(e));
            vb_kdtree_free(t);
            return VB_ERR_CUDA;
        }
    }
    *out = t;
    return VB_OK;
}

int vb_kdtree_free(vb_tree *t) {
    if (!t) return VB_OK;
    if (t->block) {
        cudaSetDevice(t->ctx->device);
        cudaFreeAsync(t->block, t->ctx->stream);
    }
    delete t;
    return VB_OK;
}

uint32_t vb_kdtree_size(const vb_tree *t) { return t ? t->n : 0; }
uint32_t vb_kdtree_height(const vb_tree *t) { return t ? t->height : 0; }

int vb_kdtree_export(vb_tree *t, uint32_t *idx_preorder, float *pts_preorder) {
    VB_REQUIRE(t != nullptr, VB_ERR_INVALID, "tree is NULL");
    if (t->n == 0) return VB_OK;
    vb_ctx *ctx = t->ctx;
    VB_CUDA(cudaSetDevice(ctx->device));
    std::vector<float> xs, ys;
    if (pts_preorder) {
        xs.resize(t->n); ys.resize(t->n);
        VB_CUDA(cudaMemcpyAsync(xs.data(), t->x, (size_t)t->n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        VB_CUDA(cudaMemcpyAsync(ys.data(), t->y, (size_t)t->n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (idx_preorder) VB_CUDA(cudaMemcpyAsync(idx_preorder, t->idx, (size_t)t->n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    VB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (pts_preorder)
        for (uint32_t i = 0; i < t->n; i++) { pts_preorder[2 * i] = xs[i]; pts_preorder[2 * i + 1] = ys[i]; }
    return VB_OK;
}

int vb_kdtree_nearest_d(vb_tree *t, const float *q_d, uint32_t nq, float max_d2, float *out_pt_d, int32_t *out_idx_d,
                        float *out_d2_d) {
    VB_REQUIRE(t && (q_d || nq == 0), VB_ERR_INVALID, "NULL argument");
    if (nq == 0) return VB_OK;
    vb_ctx *ctx = t->ctx;
    VB_CUDA(cudaSetDevice(ctx->device));
    ctx->prof_begin("kd_nearest");
    const float2 *q2 = reinterpret_cast<const float2 *>(q_d);
    float2 *op = reinterpret_cast<float2 *>(out_pt_d);
    const long long lpq = ctx->opt("kd_lanes_per_query", 1);
    if (lpq == 32)
        k_kd_nearest<32><<<(unsigned)div_up64((size_t)nq * 32, KD_Q_THREADS), KD_Q_THREADS, 0, ctx->stream>>>(
            t->x, t->y, t->idx, t->n, q2, nq, max_d2, op, out_idx_d, out_d2_d);
    else if (lpq == 8)
        k_kd_nearest<8><<<(unsigned)div_up64((size_t)nq * 8, KD_Q_THREADS), KD_Q_THREADS, 0, ctx->stream>>>(
            t->x, t->y, t->idx, t->n, q2, nq, max_d2, op, out_idx_d, out_d2_d);
    else
        k_kd_nearest<1><<<div_up(nq, KD_Q_THREADS), KD_Q_THREADS, 0, ctx->stream>>>(t->x, t->y, t->idx, t->n, q2, nq, max_d2, op,
                                                                                   out_idx_d, out_d2_d);
    ctx->prof_end("kd_nearest");
    ctx->launches++;
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

int vb_kdtree_nearest(vb_tree *t, const float *q, uint32_t nq, float max_d2, float *out_pt, int32_t *out_idx, float *out_d2) {
    VB_REQUIRE(t && (q || nq == 0), VB_ERR_INVALID, "NULL argument");
    if (nq == 0) return VB_OK;
    vb_ctx *ctx = t->ctx;
    VB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ctx->ws_ensure(WS_Q, (size_t)nq * 8))) return rc;
    if ((rc = ctx->ws_ensure(WS_OUT0, (size_t)nq * 8))) return rc;
    if ((rc = ctx->ws_ensure(WS_OUT1, (size_t)nq * 4))) return rc;
    if ((rc = ctx->ws_ensure(WS_OUT2, (size_t)nq * 4))) return rc;
    VB_CUDA(cudaMemcpyAsync(ctx->ws[WS_Q].p, q, (size_t)nq * 8, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = vb_kdtree_nearest_d(t, ctx->ws[WS_Q].as<float>(), nq, max_d2, ctx->ws[WS_OUT0].as<float>(),
                                  ctx->ws[WS_OUT1].as<int32_t>(), ctx->ws[WS_OUT2].as<float>())))
        return rc;
    if (out_pt) VB_CUDA(cudaMemcpyAsync(out_pt, ctx->ws[WS_OUT0].p, (size_t)nq * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_idx) VB_CUDA(cudaMemcpyAsync(out_idx, ctx->ws[WS_OUT1].p, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_d2) VB_CUDA(cudaMemcpyAsync(out_d2, ctx->ws[WS_OUT2].p, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    VB_CUDA(cudaStreamSynchronize(ctx->stream));
    return VB_OK;
}

int vb_kdtree_knn_d(vb_tree *t, const float *q_d, uint32_t nq, uint32_t k, float max_d2, int32_t *out_idx_d, float *out_d2_d,
                    uint32_t *out_count_d) {
    VB_REQUIRE(t && out_idx_d && out_d2_d && (q_d || nq == 0), VB_ERR_INVALID, "NULL argument");
    VB_REQUIRE(k >= 1 && k <= (uint32_t)KD_KNN_MAX, VB_ERR_INVALID, "k must be in 1..32");
    if (nq == 0) return VB_OK;
    vb_ctx *ctx = t->ctx;
    VB_CUDA(cudaSetDevice(ctx->device));
    ctx->prof_begin("kd_knn");
    k_kd_knn<<<div_up(nq, KD_KNN_THREADS), KD_KNN_THREADS, 0, ctx->stream>>>(t->x, t->y, t->idx, t->n,
                                                                            reinterpret_cast<const float2 *>(q_d), nq, k, max_d2,
                                                                            out_idx_d, out_d2_d, out_count_d);
    ctx->prof_end("kd_knn");
    ctx->launches++;
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

int vb_kdtree_knn(vb_tree *t, const float *q, uint32_t nq, uint32_t k, float max_d2, int32_t *out_idx, float *out_d2,
                  uint32_t *out_count) {
    VB_REQUIRE(t && out_idx && out_d2 && (q || nq == 0), VB_ERR_INVALID, "NULL argument");
    VB_REQUIRE(k >= 1 && k <= (uint32_t)KD_KNN_MAX, VB_ERR_INVALID, "k must be in 1..32");
    if (nq == 0) return VB_OK;
    vb_ctx *ctx = t->ctx;
    VB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ctx->ws_ensure(WS_Q, (size_t)nq * 8))) return rc;
    if ((rc = ctx->ws_ensure(WS_OUT0, (size_t)nq * k * 4))) return rc;
    if ((rc = ctx->ws_ensure(WS_OUT1, (size_t)nq * k * 4))) return rc;
    if ((rc = ctx->ws_ensure(WS_OUT2, (size_t)nq * 4))) return rc;
    VB_CUDA(cudaMemcpyAsync(ctx->ws[WS_Q].p, q, (size_t)nq * 8, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = vb_kdtree_knn_d(t, ctx->ws[WS_Q].as<float>(), nq, k, max_d2, ctx->ws[WS_OUT0].as<int32_t>(), ctx->ws[WS_OUT1].as<float>(),
                              ctx->ws[WS_OUT2].as<uint32_t>())))
        return rc;
    VB_CUDA(cudaMemcpyAsync(out_idx, ctx->ws[WS_OUT0].p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    VB_CUDA(cudaMemcpyAsync(out_d2, ctx->ws[WS_OUT1].p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_count) VB_CUDA(cudaMemcpyAsync(out_count, ctx->ws[WS_OUT2].p, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    VB_CUDA(cudaStreamSynchronize(ctx->stream));
    return VB_OK;
}

int vb_kdtree_radius_d(vb_tree *t, const float *q_d, uint32_t nq, float radius, uint32_t *out_offsets_d, uint32_t *out_idx_d,
                       uint64_t cap, uint64_t *out_total) {
    VB_REQUIRE(t && out_offsets_d && (q_d || nq == 0), VB_ERR_INVALID, "NULL argument");
    vb_ctx *ctx = t->ctx;
    VB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ctx->ws_ensure(WS_OFFS, (size_t)(nq + 1) * 4))) return rc;
    if ((rc = ctx->ws_ensure(WS_MISC, 64))) return rc;
    uint32_t *counts = ctx->ws[WS_OFFS].as<uint32_t>();
    unsigned long long *total_d = ctx->ws[WS_MISC].as<unsigned long long>();
    const float2 *q2 = reinterpret_cast<const float2 *>(q_d);
    ctx->prof_begin("kd_radius");
    if (nq)
        k_kd_radius<false><<<div_up(nq, KD_Q_THREADS), KD_Q_THREADS, 0, ctx->stream>>>(t->x, t->y, t->idx, t->n, q2, nq, radius,
                                                                                      counts, nullptr, nullptr);
    if ((rc = ctx->ws_ensure(WS_SCAN2, (size_t)2 * (nq / SCAN_BLOCK + 3) * 4))) return rc;
    scan_counts(ctx->stream, counts, nq, out_offsets_d, total_d, ctx->ws[WS_SCAN2].as<uint32_t>(), &ctx->launches);
    ctx->launches += nq ? 1 : 0;
    VB_CUDA(cudaGetLastError());
    unsigned long long total = 0;
    VB_CUDA(cudaMemcpyAsync(&total, total_d, 8, cudaMemcpyDeviceToHost, ctx->stream));
    VB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (out_total) *out_total = total;
    if (total > 0xffffffffull) { set_error("radius result too large for 32-bit CSR offsets"); return VB_ERR_CAPACITY; }
    if (total > cap) {
        ctx->prof_end("kd_radius");
        set_error("radius search produced %llu hits, capacity %llu", total, (unsigned long long)cap);
        return VB_ERR_CAPACITY;
    }
    if (total && nq) {
        VB_REQUIRE(out_idx_d != nullptr, VB_ERR_INVALID, "out_idx is NULL");
        k_kd_radius<true><<<div_up(nq, KD_Q_THREADS), KD_Q_THREADS, 0, ctx->stream>>>(t->x, t->y, t->idx, t->n, q2, nq, radius,
                                                                                     nullptr, out_offsets_d, out_idx_d);
        ctx->launches++;
        VB_CUDA(cudaGetLastError());
    }
    ctx->prof_end("kd_radius");
    return VB_OK;
}

int vb_kdtree_radius(vb_tree *t, const float *q, uint32_t nq, float radius, uint32_t *out_offsets, uint32_t *out_idx,
                     uint64_t cap, uint64_t *out_total) {
    VB_REQUIRE(t && out_offsets && (q || nq == 0), VB_ERR_INVALID, "NULL argument");
    vb_ctx *ctx = t->ctx;
    VB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ctx->ws_ensure(WS_Q, (size_t)(nq ? nq : 1) * 8))) return rc;
    if ((rc = ctx->ws_ensure(WS_OUT0, (size_t)(nq + 1) * 4))) return rc;
    if ((rc = ctx->ws_ensure(WS_OUT1, (size_t)(cap ? cap : 1) * 4))) return rc;
    if (nq) VB_CUDA(cudaMemcpyAsync(ctx->ws[WS_Q].p, q, (size_t)nq * 8, cudaMemcpyHostToDevice, ctx->stream));
    uint64_t total = 0;
    rc = vb_kdtree_radius_d(t, ctx->ws[WS_Q].as<float>(), nq, radius, ctx->ws[WS_OUT0].as<uint32_t>(),
                            ctx->ws[WS_OUT1].as<uint32_t>(), cap, &total);
    if (out_total) *out_total = total;
    if (rc != VB_OK && rc != VB_ERR_CAPACITY) return rc;
    VB_CUDA(cudaMemcpyAsync(out_offsets, ctx->ws[WS_OUT0].p, (size_t)(nq + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (rc == VB_OK && total && out_idx)
        VB_CUDA(cudaMemcpyAsync(out_idx, ctx->ws[WS_OUT1].p, (size_t)total * 4, cudaMemcpyDeviceToHost, ctx->stream));
    VB_CUDA(cudaStreamSynchronize(ctx->stream));
    return rc;
}

}  // extern "C"
