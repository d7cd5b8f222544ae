// hamming_tc.cu — 256-bit Hamming kNN-2 on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Same contract as k_knn2_partial in hamming.cu (reference src/Frame.cpp:83-85: BFMatcher(NORM_HAMMING)
// knnMatch k=2, equal distances to the lower train index), different machine mapping. Brute-force matching
// is k^2 distance evaluations over 2k descriptors — compute-bound by a factor of ~10^4 over its compulsory
// bytes — and the XOR+POPC formulation saturates the integer pipe at ~0.87 T distances/s. With every
// descriptor bit b mapped to the 8-bit float s = 1 - 2b (e4m3 +1.0 / -1.0) the distance becomes one exact
// dot product:   hamming(a, b) = (256 - <s_a, s_b>) / 2     (products are +-1, sums are integers <= 256,
// exact in the f32 accumulator), so a tile of 128 x 256 distances is 8 tcgen05.mma (K = 32 bytes each).
//
//   k_expand_pm1   descriptor bits -> e4m3 +-1 rows of 256 bytes (one pass, 16-byte stores).
//   k_knn2_tc      persistent CTAs (one per SM) over work units; a unit = 256 queries of one pair (two M = 128
//                  accumulators of 256 columns = all 512 TMEM columns) against that pair's whole train set in
//                  256-wide tiles. Barriers, TMEM and the B ring carry over from unit to unit.
//                    warp 0     TMA producer: A once (64 KB), B tiles (64 KB) through a 2-stage ring
//                    warp 1     one thread issues the UMMAs: acc0 <- A0 * B^T, acc1 <- A1 * B^T per tile
//                    warp 2     TMEM allocation / release
//                    warps 4-7  drain acc0, warps 8-11 drain acc1 (one query row per thread): while one
//                               accumulator is being drained the other one is being computed.
//                  The drain keeps (best, second) as (dot, index) per row. Candidates arrive in increasing
//                  train index and only a strictly larger dot replaces, which is knnMatch's tie order. The
//                  common case is "no candidate in this group of 4 beats the current second best": one
//                  3-input max, two compares and a warp vote per 4 distances.
#include "common.cuh"
#include "hamming_dev.cuh"
#include "tc_common.cuh"
#include "tc_host.cuh"

namespace vb {

using namespace tc;

constexpr int TC_COLSPLIT = 2;         // threads draining one accumulator row (column parts)
constexpr int TC_THREADS = 128 + 256 * TC_COLSPLIT;   // 4 service warps + draining warps
constexpr int TC_QROWS = 256;          // queries per CTA
constexpr int TC_NCOLS = 256;          // train descriptors per tile (UMMA N)
constexpr int TC_KBYTES = 256;         // expanded descriptor: 256 e4m3 values
constexpr int TC_BOX_ROWS = 128;       // TMA box = 128 rows x 128 bytes (one swizzle atom wide) = 16 KB
constexpr uint32_t TC_BOX_BYTES = 128 * TC_BOX_ROWS;
constexpr uint32_t TC_A_BYTES = 4 * TC_BOX_BYTES;    // [half][khalf][128 rows][128 B]
constexpr uint32_t TC_B_BYTES = 4 * TC_BOX_BYTES;    // [khalf][256 rows][128 B]
constexpr uint32_t TC_SMEM_BYTES = TC_A_BYTES + 2 * TC_B_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;

__global__ void __launch_bounds__(256) k_expand_pm1(const uint32_t *__restrict__ src, size_t stride_words, uint32_t rows,
                                                    uint32_t P, uint8_t *__restrict__ dst) {
    // one thread = 16 descriptor bits -> 16 output bytes
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)P * rows * 16;
    if (t >= total) return;
    const uint32_t piece = (uint32_t)(t & 15);
    const size_t grow = t >> 4;
    const uint32_t p = (uint32_t)(grow / rows), r = (uint32_t)(grow % rows);
    const uint32_t w = __ldg(src + (size_t)p * stride_words + (size_t)r * 8 + (piece >> 1));
    const uint32_t bits = (piece & 1) ? (w >> 16) : (w & 0xffffu);
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t nib = (bits >> (4 * i)) & 0xfu;
        const uint32_t spread = (nib * 0x00204081u) & 0x01010101u;   // bit i -> LSB of byte i
        o[i] = 0x70707070u | (spread << 7);                          // e4m3: 0x70 = +128, 0xF0 = -128
    }
    *reinterpret_cast<uint4 *>(dst + grow * TC_KBYTES + piece * 16) = make_uint4(o[0], o[1], o[2], o[3]);
}

// ---- the drain ---------------------------------------------------------------------------------------
// An accumulator value is v = 2^14 * dot (see k_expand_pm1). A candidate's key is  col - v  =
// 2^14 * (2 * distance - 256) + col: an integer of magnitude below 2^23, exact in fp32, whose order is
// (distance, train index) — knnMatch's order including its lower-index-first ties. Keys are never equal.
//
// The drain works on 8-column groups. Per group a thread takes the largest of the 8 accumulator values with a
// 3-input max tree (4 FMNMX3/FMNMX) and turns it into ONE key, (first column of the group) - 2^14 * (best dot of
// the group), whose order is (best distance in the group, group position); the four group keys of a 32-column
// chunk are offered, two at a time, to the row's running (r0 <= r1). No per-value work beyond the max tree, no
// branch, no vote, nothing data-dependent. What this yields per row is the group holding the nearest neighbour
// (equal distances resolve to the lower group, hence to the lower index) and the best OTHER group. The two
// nearest neighbours lie in those two groups: the second one is either another member of the best's group or the
// best member of the best other group. k_knn2_tc_fix evaluates those 16 distances directly (XOR + POPC on the
// original descriptors) and takes their top-2 by (distance, index).
constexpr int TC_KEY_SHIFT = 14;
constexpr uint32_t TC_MAX_TRAIN = 1u << TC_KEY_SHIFT;
constexpr float TC_KEY_BIAS = 4194304.f;   // 2^22 = 2^14 * 256: makes keys non-negative
constexpr float TC_KEY_NONE = 3.0e7f;      // above every real key (< 2^24)

__device__ __forceinline__ float min3(float a, float b, float c) { return fminf(fminf(a, b), c); }
__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// One 32-column chunk of one row: the four minimum keys of its 8-column groups go into the running (r0 <= r1).
// Columns are counted from the start of the warp's column part (C0 = 32 * chunk); `base` makes them global.
constexpr int TC_GROUP = 8;   // columns represented by one offered key (what k_knn2_tc_fix re-examines)
__device__ __forceinline__ void offer2(float a, float b, float &r0, float &r1) {
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    r1 = min3(r1, fmaxf(r0, lo), hi);
    r0 = fminf(r0, lo);
}
__device__ __forceinline__ void offer1(float a, float &r0, float &r1) {
    r1 = fminf(r1, fmaxf(r0, a));
    r0 = fminf(r0, a);
}
template <int C0, bool MASKED, int NCOLS = 32>
__device__ __forceinline__ void drain_chunk(const uint32_t (&raw)[NCOLS], uint32_t nvalid, float base, float &r0, float &r1) {
    static_assert(NCOLS == 32 || NCOLS == 16 || NCOLS == 8, "whole 8-column groups");
    float g[NCOLS / 8];
#pragma unroll
    for (int j = 0; j < NCOLS / 8; j++) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            v[i] = __uint_as_float(raw[8 * j + i]);
            if (MASKED) v[i] = (uint32_t)(C0 + 8 * j + i) < nvalid ? v[i] : -TC_KEY_NONE;
        }
        const float vmax = fmaxf(max3(max3(v[0], v[1], v[2]), max3(v[3], v[4], v[5]), v[6]), v[7]);
        g[j] = __fadd_rn(__fsub_rn((float)(C0 + 8 * j), vmax), base);   // group key: first column of the group - 2^14 * best dot
    }
    if (NCOLS == 8) {
        offer1(g[0], r0, r1);
    } else {
#pragma unroll
        for (int j = 0; j + 1 < NCOLS / 8; j += 2) offer2(g[j], g[j + 1], r0, r1);
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
k_knn2_tc(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_t, uint32_t n1, uint32_t n2,
          uint32_t rowstride_q, uint32_t rowstride_t, uint32_t nunits, uint2 *__restrict__ part, int dbg) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzle-128B tiles need 1024-byte alignment
    const uint32_t sA = smem0;
    const uint32_t sB = smem0 + TC_A_BYTES;
    const uint32_t sBar = sB + 2 * TC_B_BYTES;
    const uint32_t bar_a = sBar, bar_full = sBar + 8, bar_empty = sBar + 24, bar_tfull = sBar + 40, bar_tempty = sBar + 56;
    const uint32_t s_tmem = sBar + 72, bar_afree = sBar + 80;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Persistent CTA: work units u = blockIdx.x, blockIdx.x + gridDim.x, ... with unit = (pair p, 256-query block).
    // Barriers, TMEM and the B ring live across units; only A is reloaded (once the previous unit's last MMA has
    // retired), so the per-unit cost is one exposed A load instead of a CTA launch + TMEM allocation + pipeline fill.
    const uint32_t qblocks = (n1 + TC_QROWS - 1) / TC_QROWS;
    const uint32_t ntiles = (n2 + TC_NCOLS - 1) / TC_NCOLS;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_t);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(bar_a, 1);
        mbar_init(bar_afree, 1);
        for (int s = 0; s < 2; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
            mbar_init(bar_tfull + 8 * s, 1);
            mbar_init(bar_tempty + 8 * s, 4 * TC_COLSPLIT);   // one arrival per draining warp
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(s_tmem));

    // Service warps walk their loops whole (warp-uniform control flow, every lane waits on the barriers) and one elected lane
    // issues: under elect.sync ptxas emits the uniform-datapath instructions (UTMALDG, UTCxMMA, UTCBAR) straight-line.
    if (warp == 0) {
        {
            const bool leader = elect_one();
            uint32_t g = 0, ul = 0;   // tiles / units issued so far by this CTA
            for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x, ul++) {
                const uint32_t p = u / qblocks, q0 = (u % qblocks) * TC_QROWS;
                const int32_t qrow = (int32_t)(p * rowstride_q + q0);
                mbar_wait(bar_afree, (ul & 1) ^ 1);   // the previous unit's MMAs no longer read A
                if (leader) {
                    mbar_expect_tx(bar_a, TC_A_BYTES);
#pragma unroll
                    for (int h = 0; h < 2; h++)
#pragma unroll
                        for (int kk = 0; kk < 2; kk++)
                            tma_load_2d(sA + (h * 2 + kk) * TC_BOX_BYTES, &map_q, kk * 128, qrow + h * 128, bar_a);
                }
                const int32_t trow = (int32_t)(p * rowstride_t);
                for (uint32_t j = 0; j < ntiles; j++, g++) {
                    const uint32_t s = g & 1, ph = (g >> 1) & 1;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    if ((dbg & 4) && g >= 2) {   // measurement only: reuse the resident stage, no L2 -> SM traffic
                        if (leader) mbar_arrive(bar_full + 8 * s);
                        continue;
                    }
                    if (leader) {
                        mbar_expect_tx(bar_full + 8 * s, TC_B_BYTES);
                        const uint32_t dst = sB + s * TC_B_BYTES;
#pragma unroll
                        for (int kk = 0; kk < 2; kk++)
#pragma unroll
                            for (int r = 0; r < 2; r++)
                                tma_load_2d(dst + (kk * 2 + r) * TC_BOX_BYTES, &map_t, kk * 128,
                                            trow + (int32_t)(j * TC_NCOLS) + r * 128, bar_full + 8 * s);
                    }
                }
            }
        }
    } else if (warp == 1) {
        {
            const bool leader = elect_one();
            constexpr uint32_t idesc = umma_idesc(UMMA_FMT_E4M3, 128, TC_NCOLS);
            uint32_t g = 0, ul = 0;
            for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x, ul++) {
                mbar_wait(bar_a, ul & 1);
                tc_fence_after();
                for (uint32_t j = 0; j < ntiles; j++, g++) {
                    const uint32_t s = g & 1, ph = (g >> 1) & 1;
                    mbar_wait(bar_full + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t bbase = sB + s * TC_B_BYTES;
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        mbar_wait(bar_tempty + 8 * h, (g & 1) ^ 1);   // accumulator h drained (previous tile)
                        tc_fence_after();
                        if (leader) {
#pragma unroll
                            for (int kk = 0; kk < 2; kk++)
#pragma unroll
                                for (int k = 0; k < 4; k++) {
                                    const uint64_t ad = smem_desc_sw128(sA + (h * 2 + kk) * TC_BOX_BYTES + k * 32);
                                    const uint64_t bd = smem_desc_sw128(bbase + kk * 2 * TC_BOX_BYTES + k * 32);
                                    umma_f8f6f4(tmem_base + h * TC_NCOLS, ad, bd, idesc, (kk | k) != 0 ? 1u : 0u);
                                }
                            umma_commit(bar_tfull + 8 * h);
                        }
                    }
                    if (leader) umma_commit(bar_empty + 8 * s);   // both halves have consumed this B stage
                }
                if (leader) umma_commit(bar_afree);   // ... and every MMA of this unit has consumed A
            }
        }
    } else if (warp >= 4) {
        const uint32_t ew = warp - 4;
        const uint32_t quad = warp & 3, h = (ew >> 2) & 1, ch = ew >> 3;   // TMEM lane quadrant, accumulator, column part
        const uint32_t row = h * 128 + quad * 32 + lane;
        constexpr uint32_t CW = TC_NCOLS / TC_COLSPLIT;
        constexpr int NCH = CW / 32;
        static_assert(NCH == 4, "the drain below is written out for four 32-column chunks per warp");
        const uint32_t taddr = tmem_base + ((quad * 32u) << 16) + h * TC_NCOLS + ch * CW;
        uint32_t raw0[32], raw1[32];
        uint32_t g = 0;   // tiles drained so far by this CTA (accumulator barrier phase)
        for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
            const uint32_t p = u / qblocks, q = (u % qblocks) * TC_QROWS + row;
            // running (best, best outside the best's group) as biased global keys
            float r0 = TC_KEY_NONE, r1 = TC_KEY_NONE;
            float tbase = (float)(ch * CW) + TC_KEY_BIAS;   // key bias + first column of this warp's part of tile j
            for (uint32_t j = 0; j < ntiles; j++, g++, tbase += (float)TC_NCOLS) {
                mbar_wait(bar_tfull + 8 * h, g & 1);
                tc_fence_after();
                const uint32_t tile0 = j * TC_NCOLS + ch * CW;
                if ((dbg & 2) || tile0 >= n2) {   // (dbg 2: MMA floor measurement) / nothing valid in this part of the last tile
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8 * h);
                    continue;
                }
                const bool masked = tile0 + CW > n2;
                const uint32_t nvalid = n2 - tile0;
                // chunk c+1 is in flight while chunk c is reduced; the accumulator goes back to the MMA warp as soon
                // as its last chunk has landed in registers
                tmem_ld32(taddr, raw0);
                tmem_wait_ld_regs(raw0);
                tmem_ld32(taddr + 32, raw1);
                if (masked) drain_chunk<0, true>(raw0, nvalid, tbase, r0, r1); else drain_chunk<0, false>(raw0, nvalid, tbase, r0, r1);
                tmem_wait_ld_regs(raw1);
                tmem_ld32(taddr + 64, raw0);
                if (masked) drain_chunk<32, true>(raw1, nvalid, tbase, r0, r1); else drain_chunk<32, false>(raw1, nvalid, tbase, r0, r1);
                tmem_wait_ld_regs(raw0);
                tmem_ld32(taddr + 96, raw1);
                if (masked) drain_chunk<64, true>(raw0, nvalid, tbase, r0, r1); else drain_chunk<64, false>(raw0, nvalid, tbase, r0, r1);
                tmem_wait_ld_regs(raw1);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * h);
                if (masked) drain_chunk<96, true>(raw1, nvalid, tbase, r0, r1); else drain_chunk<96, false>(raw1, nvalid, tbase, r0, r1);
            }
            if (q < n1) {
                // biased key -> (distance << KNN_IDX_BITS | index); a part that saw fewer than two groups reports the
                // largest key, which the merge ignores
                uint32_t out[2];
                const float ks[2] = {r0, r1};
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const uint32_t ki = __float2uint_rz(ks[i]);
                    out[i] = ks[i] < 16777216.f ? ((ki >> (TC_KEY_SHIFT + 1)) << KNN_IDX_BITS) | (ki & (TC_MAX_TRAIN - 1u))
                                                : 0xffffffffu;
                }
                part[((size_t)p * TC_COLSPLIT + ch) * n1 + q] = make_uint2(out[0], out[1]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---- FP4 variant ---------------------------------------------------------------------------------------
// Same algorithm with the descriptors expanded to packed e2m1 (+1 = 0x2, -1 = 0xA; 128 bytes per descriptor) and
// block-scaled UMMAs (kind::mxf4, K = 64 per instruction: twice the fp8 rate, half the shared-memory and L2 bytes).
// Every scale factor is the same ue8m0 value 2^7, for A and for B, so an accumulator still holds 2^14 * dot exactly
// and the drain is unchanged; because the factors are all equal, their TMEM layout does not matter — a 32-column
// strip of TMEM is simply filled with 0x86 bytes once per CTA. That strip costs accumulator width: two accumulators
// of 240 columns (512 = 32 + 2 * 240), so a train tile is 240 descriptors.
constexpr int T4_NCOLS = 240;                          // train descriptors per tile (UMMA N)
constexpr int T4_STAGES = 4;                           // B ring depth
constexpr int T4_PARTS = 4;                            // column parts of an accumulator (one draining warp each per quadrant)
constexpr int T4_ROWBYTES = 128;                       // expanded descriptor: 256 e2m1 values
constexpr uint32_t T4_A_BYTES = 2 * 128 * T4_ROWBYTES;        // [half][128 rows][128 B]
constexpr uint32_t T4_B_BYTES = T4_NCOLS * T4_ROWBYTES;       // [240 rows][128 B] = 30 atoms of 1 KB
constexpr uint32_t T4_SF_COLS = 32;                    // TMEM columns [0, 32): scale factors
constexpr uint32_t T4_SMEM_BYTES = T4_A_BYTES + T4_STAGES * T4_B_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
static_assert(T4_B_BYTES % 1024 == 0, "stages must stay 1024-byte aligned for the 128-byte swizzle");

__global__ void __launch_bounds__(256) k_expand_e2m1(const uint32_t *__restrict__ src, size_t stride_words, uint32_t rows,
                                                     uint32_t P, uint8_t *__restrict__ dst) {
    // one thread = one 32-bit word of a descriptor -> 16 output bytes (two bits per byte, low nibble first)
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;   // word index within frame p = blockIdx.y
    if (t >= rows * 8) return;
    const uint32_t p = blockIdx.y;
    const uint32_t w = __ldg(src + (size_t)p * stride_words + t);
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t x = (w >> (8 * i)) & 0xffu;
        const uint32_t sp = (x | (x << 6) | (x << 12) | (x << 18)) & 0x03030303u;   // bits (2j, 2j+1) -> byte j, positions 0, 1
        o[i] = 0x22222222u | ((sp & 0x01010101u) << 3) | ((sp & 0x02020202u) << 6);  // nibble 0x2 = +1.0, 0xA = -1.0
    }
    *reinterpret_cast<uint4 *>(dst + ((size_t)p * rows * 8 + t) * 16) = make_uint4(o[0], o[1], o[2], o[3]);
}

// ---- packed-integer drain (DRAIN 6) ------------------------------------------------------------------------------------
// The ALU pipe issues one warp instruction per two cycles, and the float drain needs 13 of them per 16 columns: 447 cycles
// per accumulator for the max trees and offers alone, against 494 cycles of UMMA time (tools/probe/drain_probe.cu measures
// the slice in isolation at 28-30 cycles). Packed 16-bit integer maxima (VIMNMX3.U16x2, same issue rate) halve that, if an
// accumulator word carries an INTEGER in its low half. It does when the accumulator starts from a magic constant instead of
// zero: the draining warps write T6_MAGIC = 1.5 * 2^23 + 16512 into their columns right after reading them
// (tcgen05.st), every UMMA accumulates, and with scale factors 2^3 * 2^3 a word ends as MAGIC + 64 * dot — exact (the
// products are +-64, the sum stays below 2^24; tools/probe/tmem_probe.cu checks the tensor pipe on it), upper half
// constant 0x4B40, low half = 128 * (257 - distance). tcgen05.ld.pack::16b then delivers two columns per register.
// A group is the 8 EVEN or the 8 ODD columns of a 16-column span (the two halves of the packed registers); its key is
// 128 * (257 - best distance) + (127 - position), position = 4 * tile + span within the warp's 64-column part, so a
// larger key is a smaller (distance, position) and 7 bits of position cover 32 train tiles (n2 <= 7 680; larger train
// sets take variant 1). Keys below 128 mean "no candidate" (a real key has 257 - distance >= 1).
constexpr uint32_t T6_MAGIC = 0x4B404080u;
// Timeline of one CTA (builds with -DVB_TC_TRACE on top of TUNING=1, tc_dbg & 32; the trace points cost the UMMA warp time even
// when they are switched off, so the plain TUNING build leaves them out): SM clock at the hand-shake points of the UMMA thread (role 0), two draining
// warps (1, 2) and a re-arming warp (3), per step = 2 * tile + accumulator; read back with vb_debug_tc_trace (tools/tc_trace.py).
#if defined(VB_TUNING) && defined(VB_TC_TRACE)
constexpr uint32_t TC_TRACE_STEPS = 128, TC_TRACE_EVENTS = 6;
constexpr uint32_t TC_TRACE_ROLES = 18;   // 0 = UMMA thread, 1 = re-arming warp 20, 2 + ew = draining warp ew
__device__ long long g_tc_trace[TC_TRACE_ROLES * TC_TRACE_STEPS * TC_TRACE_EVENTS];
__device__ __forceinline__ uint32_t mbar_wait_count(uint32_t bar, uint32_t parity) {   // mbar_wait, returning the failed polls
    uint32_t n = 0;
    while (!vb::tc::mbar_try_wait(bar, parity)) n++;
    return n;
}
#define TC_TRACE(role, step, ev)                                                                                       \
    do {                                                                                                               \
        if ((role) >= 0 && (step) < TC_TRACE_STEPS) g_tc_trace[((role) * TC_TRACE_STEPS + (step)) * TC_TRACE_EVENTS + (ev)] = clock64(); \
    } while (0)
#else
#define TC_TRACE(role, step, ev) do { } while (0)
#endif
constexpr uint32_t T6_MAX_TILES = 32;   // 4 spans per part (variant 6)
constexpr uint32_t T7_MAX_TILES = 25;   // 5 spans per part (variant 7)
__device__ __forceinline__ void tmem_st32_unpack16(uint32_t taddr, const uint32_t (&v)[16]) {   // 32 columns from 16 registers
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.unpack::16b.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
template <int NREG>
__device__ __forceinline__ void tmem_st16_unpack16(uint32_t taddr, const uint32_t (&v)[NREG]) {   // 16 columns from 8 registers
    static_assert(NREG >= 8, "eight packed registers");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.unpack::16b.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ uint32_t vmax3u2(uint32_t a, uint32_t b, uint32_t c) { return __vmaxu2(__vmaxu2(a, b), c); }
template <bool MASKED>
__device__ __forceinline__ void drain_span16(const uint32_t *raw, uint32_t c_first, uint32_t nvalid, uint32_t posc, uint32_t &r0,
                                             uint32_t &r1) {
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        v[i] = raw[i];
        if (MASKED)
            v[i] &= ((c_first + 2u * i < nvalid) ? 0x0000ffffu : 0u) | ((c_first + 2u * i + 1u < nvalid) ? 0xffff0000u : 0u);
    }
    const uint32_t key = vmax3u2(vmax3u2(v[0], v[1], v[2]), vmax3u2(v[3], v[4], v[5]), __vmaxu2(v[6], v[7])) + posc;
    const uint32_t t = __vminu2(r0, key);
    r0 = __vmaxu2(r0, key);
    r1 = __vmaxu2(r1, t);
}
// 16-bit key -> the fix kernel's group key (distance << 22 | first column of the group); part_c0 = first column of the part
template <uint32_t SPANS = 4>   // spans per part: position = SPANS * tile + span
__device__ __forceinline__ uint32_t t6_group_key(uint32_t key16, uint32_t part_c0, uint32_t parity) {
    if (key16 < 128u) return 0xffffffffu;
    const uint32_t pos = 127u - (key16 & 127u);
    const uint32_t col = (pos / SPANS) * (uint32_t)T4_NCOLS + part_c0 + (pos % SPANS) * 16u + parity;
    return ((257u - (key16 >> 7)) << KNN_IDX_BITS) | col;
}

// One step (one accumulator of one tile) of the packed drain for a warp's column part. FULL = every column of the part is a
// real train descriptor (all tiles but the last): no bounds, no branches — loads, the constant back, hand-back, max trees.
// UNPACK (variant 10): the constant is the denormal of variant 9 and goes back with tcgen05.st.unpack::16b, two columns per
// register — half the store traffic of the hand-back.
template <bool WIDE, bool FULL, bool UNPACK = false>
__device__ __forceinline__ void t6_step(uint32_t taddr, uint32_t bar_full_h, uint32_t bar_empty_h, uint32_t parity, uint32_t lane,
                                        uint32_t (&ra)[16], uint32_t (&rb)[16], const uint32_t (&cst)[8], bool skip, bool masked,
                                        uint32_t nvalid, uint32_t posc, uint32_t &r0, uint32_t &r1, int dbg, int trole = -1,
                                        uint32_t tstep = 0) {
    TC_TRACE(trole, tstep, 0);
    mbar_wait(bar_full_h, parity);
    tc_fence_after();
    TC_TRACE(trole, tstep, 1);
    if (FULL || !skip) {
        tmem_ld32_pack16(taddr, ra);
        tmem_ld32_pack16(taddr + (WIDE ? 32u : 16u), rb);   // narrow part: columns [16, 48), the upper half is its third span
        tmem_wait_ld_regs16(ra);
    }
    if (UNPACK) {
        tmem_st16_unpack16(taddr, cst);
        tmem_st16_unpack16(taddr + 16, cst);
    } else if (FULL || !(dbg & 16)) {   // (dbg 16: timing without the stores, results invalid)
        tmem_st8(taddr, cst);    // the next tile accumulates onto the constant again
        tmem_st8(taddr + 8, cst);
        tmem_st8(taddr + 16, cst);
        tmem_st8(taddr + 24, cst);
    }
    if (FULL || !skip) tmem_wait_ld_regs16(rb);
    TC_TRACE(trole, tstep, 2);
    if (UNPACK) {
        tmem_st16_unpack16(taddr + 32, cst);
        if (WIDE) tmem_st16_unpack16(taddr + 48, cst);
        tmem_wait_st();
    } else if (FULL || !(dbg & 16)) {
        tmem_st8(taddr + 32, cst);
        tmem_st8(taddr + 40, cst);
        if (WIDE) {
            tmem_st8(taddr + 48, cst);
            tmem_st8(taddr + 56, cst);
        }
        tmem_wait_st();
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty_h);   // handed back before anything is reduced
    TC_TRACE(trole, tstep, 3);
    if (!FULL && skip) return;
    if (!FULL && masked) {
        drain_span16<true>(&ra[0], 0u, nvalid, posc, r0, r1);
        drain_span16<true>(&ra[8], 16u, nvalid, posc - 0x00010001u, r0, r1);
        if (WIDE) {
            drain_span16<true>(&rb[0], 32u, nvalid, posc - 2u * 0x00010001u, r0, r1);
            drain_span16<true>(&rb[8], 48u, nvalid, posc - 3u * 0x00010001u, r0, r1);
        } else {
            drain_span16<true>(&rb[8], 32u, nvalid, posc - 2u * 0x00010001u, r0, r1);
        }
    } else {
        drain_span16<false>(&ra[0], 0u, nvalid, posc, r0, r1);
        drain_span16<false>(&ra[8], 16u, nvalid, posc - 0x00010001u, r0, r1);
        if (WIDE) {
            drain_span16<false>(&rb[0], 32u, nvalid, posc - 2u * 0x00010001u, r0, r1);
            drain_span16<false>(&rb[8], 48u, nvalid, posc - 3u * 0x00010001u, r0, r1);
        } else {
            drain_span16<false>(&rb[8], 32u, nvalid, posc - 2u * 0x00010001u, r0, r1);
        }
    }
#if defined(VB_TUNING) && defined(VB_TC_TRACE)
    if (trole >= 0) { asm volatile("" ::"r"(r0), "r"(r1) : "memory"); TC_TRACE(trole, tstep, 4); }
#endif
}
// All tiles of one unit for a warp's part: the leading full tiles through the branch-free step, the rest through the general one.
template <bool WIDE, bool UNPACK = false>
__device__ __forceinline__ void t6_unit(uint32_t lane_base, uint32_t bar_tfull, uint32_t bar_tempty, uint32_t &g, uint32_t lane,
                                        uint32_t c0, uint32_t cw, uint32_t n2, uint32_t ntiles, uint32_t nfull,
                                        uint32_t (&ra)[16], uint32_t (&rb)[16], const uint32_t (&cst)[8], uint32_t (&r0)[2],
                                        uint32_t (&r1)[2], int dbg, int trole = -1) {
    uint32_t posc = 127u * 0x00010001u;
    uint32_t j = 0;
    for (; j < nfull; j++, g++, posc -= 4u * 0x00010001u) {
        t6_step<WIDE, true, UNPACK>(lane_base, bar_tfull, bar_tempty, g & 1, lane, ra, rb, cst, false, false, cw, posc, r0[0], r1[0], 0, trole, 2 * g);
        t6_step<WIDE, true, UNPACK>(lane_base + T4_NCOLS, bar_tfull + 8, bar_tempty + 8, g & 1, lane, ra, rb, cst, false, false, cw, posc,
                                    r0[1], r1[1], 0, trole, 2 * g + 1);
    }
    for (; j < ntiles; j++, g++, posc -= 4u * 0x00010001u) {
        const uint32_t tile0 = j * T4_NCOLS + c0;
        const bool skip = (dbg & 2) || tile0 >= n2;
        const bool masked = tile0 + cw > n2;
        const uint32_t nvalid = skip ? 0 : n2 - tile0;
        t6_step<WIDE, false, UNPACK>(lane_base, bar_tfull, bar_tempty, g & 1, lane, ra, rb, cst, skip, masked, nvalid, posc, r0[0], r1[0], dbg);
        t6_step<WIDE, false, UNPACK>(lane_base + T4_NCOLS, bar_tfull + 8, bar_tempty + 8, g & 1, lane, ra, rb, cst, skip, masked, nvalid, posc,
                                     r0[1], r1[1], dbg);
    }
}

// ---- packed drain with re-arming warps (DRAIN 8, 9) -------------------------------------------------------------------------
// In variant 6 a draining warp's step is load -> store of the constant -> wait for the store -> hand-back -> max trees, one
// after the other, and that serial chain (not a pipe) is what a step costs. Here four more warps (one per TMEM lane quadrant)
// do nothing but re-arm: a draining warp loads its part, says so on a "read" barrier and goes straight to its max trees;
// the re-arming warp of the quadrant waits for that barrier, writes the constant over the quadrant's 240 columns and hands
// the accumulator back to the UMMA warp. The store round trip runs beside the max trees instead of in front of them.
//   8  constant 1.5 * 2^23 + 16512 as in variant 6 (full-width stores)
//   9  constant 16512 * 2^-149 — a DENORMAL whose upper half-word is zero — with scale factors 2^-71 * 2^-72, so a product
//      of +-1 is +-64 units of 2^-149 and the low half-word is the same integer as in variant 6; the constant then goes
//      back with tcgen05.st.unpack::16b (two columns per register: half the store traffic). Needs the tensor pipe to
//      accumulate denormals exactly, which tools/probe/tmem_probe.cu checks.
constexpr uint32_t T9_MAGIC = 0x00004080u;
constexpr int T8_THREADS = 128 + 32 * 16 + 32 * 4;
template <bool WIDE, bool FULL>
__device__ __forceinline__ void t8_step(uint32_t taddr, uint32_t bar_full_h, uint32_t bar_read_h, uint32_t parity, uint32_t lane,
                                        uint32_t (&ra)[16], uint32_t (&rb)[16], bool skip, bool masked, uint32_t nvalid,
                                        uint32_t posc, uint32_t &r0, uint32_t &r1, int trole = -1, uint32_t tstep = 0) {
    TC_TRACE(trole, tstep, 0);
    mbar_wait(bar_full_h, parity);
    tc_fence_after();
    TC_TRACE(trole, tstep, 1);
    if (FULL || !skip) {
        tmem_ld32_pack16(taddr, ra);
        tmem_ld32_pack16(taddr + (WIDE ? 32u : 16u), rb);
        tmem_wait_ld_regs16(ra);
        tmem_wait_ld_regs16(rb);
    }
    TC_TRACE(trole, tstep, 2);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_read_h);   // the part is in registers: the quadrant's re-arming warp may overwrite it
    TC_TRACE(trole, tstep, 3);
    if (!FULL && skip) return;
    if (!FULL && masked) {
        drain_span16<true>(&ra[0], 0u, nvalid, posc, r0, r1);
        drain_span16<true>(&ra[8], 16u, nvalid, posc - 0x00010001u, r0, r1);
        if (WIDE) {
            drain_span16<true>(&rb[0], 32u, nvalid, posc - 2u * 0x00010001u, r0, r1);
            drain_span16<true>(&rb[8], 48u, nvalid, posc - 3u * 0x00010001u, r0, r1);
        } else {
            drain_span16<true>(&rb[8], 32u, nvalid, posc - 2u * 0x00010001u, r0, r1);
        }
    } else {
        drain_span16<false>(&ra[0], 0u, nvalid, posc, r0, r1);
        drain_span16<false>(&ra[8], 16u, nvalid, posc - 0x00010001u, r0, r1);
        if (WIDE) {
            drain_span16<false>(&rb[0], 32u, nvalid, posc - 2u * 0x00010001u, r0, r1);
            drain_span16<false>(&rb[8], 48u, nvalid, posc - 3u * 0x00010001u, r0, r1);
        } else {
            drain_span16<false>(&rb[8], 32u, nvalid, posc - 2u * 0x00010001u, r0, r1);
        }
    }
#if defined(VB_TUNING) && defined(VB_TC_TRACE)
    if (trole >= 0) { asm volatile("" ::"r"(r0), "r"(r1) : "memory"); TC_TRACE(trole, tstep, 4); }
#endif
}
template <bool WIDE>
__device__ __forceinline__ void t8_unit(uint32_t lane_base, uint32_t bar_tfull, uint32_t bar_tread, uint32_t &g, uint32_t lane,
                                        uint32_t c0, uint32_t cw, uint32_t n2, uint32_t ntiles, uint32_t nfull,
                                        uint32_t (&ra)[16], uint32_t (&rb)[16], uint32_t (&r0)[2], uint32_t (&r1)[2], int trole = -1) {
    uint32_t posc = 127u * 0x00010001u;
    uint32_t j = 0;
    for (; j < nfull; j++, g++, posc -= 4u * 0x00010001u) {
        t8_step<WIDE, true>(lane_base, bar_tfull, bar_tread, g & 1, lane, ra, rb, false, false, cw, posc, r0[0], r1[0], trole, 2 * g);
        t8_step<WIDE, true>(lane_base + T4_NCOLS, bar_tfull + 8, bar_tread + 8, g & 1, lane, ra, rb, false, false, cw, posc, r0[1], r1[1], trole, 2 * g + 1);
    }
    for (; j < ntiles; j++, g++, posc -= 4u * 0x00010001u) {
        const uint32_t tile0 = j * T4_NCOLS + c0;
        const bool skip = tile0 >= n2;
        const bool masked = tile0 + cw > n2;
        const uint32_t nvalid = skip ? 0 : n2 - tile0;
        t8_step<WIDE, false>(lane_base, bar_tfull, bar_tread, g & 1, lane, ra, rb, skip, masked, nvalid, posc, r0[0], r1[0]);
        t8_step<WIDE, false>(lane_base + T4_NCOLS, bar_tfull + 8, bar_tread + 8, g & 1, lane, ra, rb, skip, masked, nvalid, posc, r0[1], r1[1]);
    }
}

// DRAIN selects how a draining warp moves its 64-column part of an accumulator out of TMEM:
//   0  the whole part in one go (two x32 loads), accumulator handed back, then the max trees — TMEM reads (480 clk per
//      accumulator at 64 B/clk per scheduler) and ALU work (420 clk) of a step run one after the other.
//   1  16-column slices through two rotating 16-register buffers: slice s+1 is in flight while slice s is reduced, so the
//      TMEM port and the ALU pipe of a scheduler overlap inside a step, and half as many raw registers are live.
//   4  24 draining warps (6 per scheduler, 40-column parts) instead of 16: the profile of variants 0/1 shows the drain
//      latency-bound per warp (one instruction per ~6.4 cycles: dependency waits, branch resolution, instruction fetch)
//      with only four warps per scheduler to hide it behind; three buffers (16 + 16 + 8 registers) so that all of a part's
//      loads are in flight at once, and the masked path (last tile only) kept out of the hot loop.
constexpr int T4_THREADS_WIDE = 128 + 32 * 24;
// SVC_HI puts the four service warps (TMA producer, UMMA issuer, TMEM allocator, spare) at the HIGHEST warp ids of the CTA
// instead of the lowest: the warp arbiter of a scheduler prefers the highest warp id among eligible warps, and the one thread
// that issues the UMMAs shares its scheduler with four draining warps that almost always have an instruction ready.
// ISSUERS = 2: the spare service warp issues the UMMAs of accumulator 1 and warp 1 those of accumulator 0. One thread's chain per
// tile — two commits, the B-stage wait, two accumulator waits, eight UMMA issues, each a dependent long-latency operation — is
// what a tile costs (tools/tc_trace.py: the thread never finds a barrier incomplete, and the draining warps idle 40 % of the
// time, yet the tensor pipe is 68 % active); two threads halve it.
template <int DRAIN, bool SVC_HI = false, int ISSUERS = 1>
__global__ void __launch_bounds__((DRAIN == 4 || DRAIN == 5 || DRAIN == 7) ? T4_THREADS_WIDE : (DRAIN == 8 || DRAIN == 9) ? T8_THREADS : TC_THREADS, 1) __maxnreg__((DRAIN >= 6) ? 72 : 128)
k_knn2_tc4(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_t, uint32_t n1, uint32_t n2,
           uint32_t rowstride_q, uint32_t rowstride_t, uint32_t nunits, uint2 *__restrict__ part, int dbg) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = smem0;
    const uint32_t sB = smem0 + T4_A_BYTES;
    const uint32_t sBar = sB + T4_STAGES * T4_B_BYTES;
    const uint32_t bar_a = sBar, bar_afree = sBar + 8, bar_tfull = sBar + 16, bar_tempty = sBar + 32, s_tmem = sBar + 48,
                   bar_full = sBar + 64, bar_empty = sBar + 64 + 8 * T4_STAGES, bar_tread = sBar + 64 + 16 * T4_STAGES;
    static_assert(64 + 16 * T4_STAGES + 16 <= 256, "barrier block");
    constexpr bool REARM = DRAIN == 8 || DRAIN == 9;
    constexpr bool DENORM = DRAIN == 9 || DRAIN == 10;   // accumulators in the denormal range, unpack::16b stores
    constexpr uint32_t MAGIC = DENORM ? T9_MAGIC : T6_MAGIC;

    const uint32_t wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ uint32_t s_cst[16];
    __shared__ uint2 s_keys[DRAIN == 10 ? 2 : 1][DRAIN == 10 ? 4 : 1][DRAIN == 10 ? 3 : 1][DRAIN == 10 ? 2 : 1][DRAIN == 10 ? 32 : 1];   // variant 10: unit parity x quadrant x part 1..3 x row half x lane
    // read back after the first __syncthreads (variant 9 stores two columns per register)
    if (DRAIN >= 6 && threadIdx.x < 16) s_cst[threadIdx.x] = DENORM ? (T9_MAGIC | (T9_MAGIC << 16)) : T6_MAGIC;
    constexpr uint32_t NDW = (DRAIN == 4 || DRAIN == 5 || DRAIN == 7) ? 24u : 16u;   // draining warps
    // role index: 0..3 = service warps, 4.. = draining warps ((wid + 4) & 3 == wid & 3, so the TMEM lane quadrant is unchanged)
    const uint32_t warp = SVC_HI ? (wid >= NDW ? wid - NDW : wid + 4u) : wid;
    const uint32_t qblocks = (n1 + TC_QROWS - 1) / TC_QROWS;
    const uint32_t ntiles = (n2 + T4_NCOLS - 1) / T4_NCOLS;
#if defined(VB_TUNING) && defined(VB_TC_TRACE)
    const int trole = ((dbg & 32) && blockIdx.x == 0 && lane == 0) ? (warp == 1 ? 0 : warp == 20 ? 1 : (warp >= 4 && warp < 20) ? (int)warp - 2 : -1) : -1;
#else
    constexpr int trole = -1;
#endif

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_t);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(bar_a, 1);
        mbar_init(bar_afree, ISSUERS);
        for (int s = 0; s < 2; s++) {
            mbar_init(bar_tfull + 8 * s, 1);
            mbar_init(bar_tempty + 8 * s, REARM ? 4 : DRAIN == 3 ? 8 : (DRAIN == 5 || DRAIN == 7) ? 12 : DRAIN == 4 ? 24 : 4 * T4_PARTS);   // draining (re-arming) warps per accumulator
            if (REARM) mbar_init(bar_tread + 8 * s, 4 * T4_PARTS);
        }
        for (int s = 0; s < T4_STAGES; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, ISSUERS);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(s_tmem));
    if (warp >= 4 && warp < 8) {   // ue8m0 2^7 everywhere (packed drain: 2^3, products are +-64)
        if (DENORM) {              // A's factors (columns [0, 16)) 2^-71, B's ([16, 32)) 2^-72: products are +-64 * 2^-149
            tmem_st16_const(tmem_base + (((warp & 3) * 32u) << 16), 0x38383838u);
            tmem_st16_const(tmem_base + (((warp & 3) * 32u) << 16) + 16u, 0x37373737u);
            tmem_wait_st();
        } else {
            tmem_st32_const(tmem_base + (((warp & 3) * 32u) << 16), DRAIN >= 6 ? 0x82828282u : 0x86868686u);
        }
    }
    if (DRAIN == 7 && warp >= 4) {   // each warp its 80 columns of its accumulator
        const uint32_t ew7 = warp - 4, t0 = tmem_base + T4_SF_COLS + (((warp & 3) * 32u) << 16) + ((ew7 >> 2) & 1u) * T4_NCOLS + (ew7 >> 3) * 80u;
        tmem_st32_const_async(t0, T6_MAGIC);
        tmem_st32_const_async(t0 + 32, T6_MAGIC);
        tmem_st16_const(t0 + 64, T6_MAGIC);
        tmem_wait_st();
    }
    if ((DRAIN == 6 || DRAIN == 10 || REARM) && warp >= 4 && warp < 20) {   // both accumulators start from the magic constant
        const uint32_t cp = (warp - 4) >> 2, t0 = tmem_base + T4_SF_COLS + (((warp & 3) * 32u) << 16) + cp * 64u;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            tmem_st32_const_async(t0 + h * T4_NCOLS, MAGIC);
            if (cp != 3) tmem_st32_const_async(t0 + h * T4_NCOLS + 32, MAGIC);
            else tmem_st16_const(t0 + h * T4_NCOLS + 32, MAGIC);
        }
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t acc0 = tmem_base + T4_SF_COLS;

    if (warp == 0) {
        {
            const bool leader = elect_one();
            uint32_t g = 0, ul = 0;
            for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x, ul++) {
                const uint32_t p = u / qblocks, q0 = (u % qblocks) * TC_QROWS;
                const int32_t qrow = (int32_t)(p * rowstride_q + q0);
                mbar_wait(bar_afree, (ul & 1) ^ 1);
                if (leader) {
                    mbar_expect_tx(bar_a, T4_A_BYTES);
                    tma_load_2d(sA, &map_q, 0, qrow, bar_a);
                    tma_load_2d(sA + 128 * T4_ROWBYTES, &map_q, 0, qrow + 128, bar_a);
                }
                const int32_t trow = (int32_t)(p * rowstride_t);
                for (uint32_t j = 0; j < ntiles; j++, g++) {
                    const uint32_t s = g % T4_STAGES, ph = (g / T4_STAGES) & 1;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    if ((dbg & 4) && g >= T4_STAGES) {   // timing floor without the B traffic (results invalid): stages keep their first tile
                        if (leader) mbar_arrive(bar_full + 8 * s);
                        continue;
                    }
                    if (leader) {
                        mbar_expect_tx(bar_full + 8 * s, T4_B_BYTES);
                        tma_load_2d(sB + s * T4_B_BYTES, &map_t, 0, trow + (int32_t)(j * T4_NCOLS), bar_full + 8 * s);
                    }
                }
            }
        }
    } else if (warp == 1 || (ISSUERS == 2 && warp == 3)) {
        // The whole warp walks the loop and waits on the barriers (warp-uniform control flow); one elected lane issues.
        {
            const bool leader = elect_one();
            constexpr uint32_t idesc = umma_idesc_mxf4(128, T4_NCOLS);
            const uint32_t sfa = tmem_base, sfb = tmem_base + 16;
            const int hmine = warp == 3 ? 1 : 0;   // ISSUERS == 2: this warp's accumulator
#if defined(VB_TUNING) && defined(VB_TC_TRACE)
            const int mrole = (ISSUERS == 2 && warp == 3) ? -1 : trole;
#else
            constexpr int mrole = -1;
#endif
            uint32_t g = 0, ul = 0;
            for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x, ul++) {
                mbar_wait(bar_a, ul & 1);
                tc_fence_after();
                for (uint32_t j = 0; j < ntiles; j++, g++) {
                    const uint32_t s = g % T4_STAGES, ph = (g / T4_STAGES) & 1;
                    TC_TRACE(mrole, 2 * g, 3);
                    mbar_wait(bar_full + 8 * s, ph);
                    tc_fence_after();
                    TC_TRACE(mrole, 2 * g, 4);
                    const uint32_t bbase = sB + s * T4_B_BYTES;
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        if (ISSUERS == 2 && h != hmine) continue;
                        TC_TRACE(mrole, 2 * g + h, 0);
                        mbar_wait(bar_tempty + 8 * h, (g & 1) ^ 1);
                        tc_fence_after();
                        TC_TRACE(mrole, 2 * g + h, 1);
                        if (leader) {
#pragma unroll
                            for (int k = 0; k < 4; k++) {   // 4 K-steps of 64 e2m1 (32 B) in the 128-byte row
                                const uint64_t ad = smem_desc_sw128(sA + h * 128 * T4_ROWBYTES + k * 32);
                                const uint64_t bd = smem_desc_sw128(bbase + k * 32);
                                umma_mxf4(acc0 + h * T4_NCOLS, ad, bd, idesc, sfa, sfb, (DRAIN >= 6 || k != 0 || (dbg & 8)) ? 1u : 0u);
                            }
                            umma_commit(bar_tfull + 8 * h);
                        }
                        TC_TRACE(mrole, 2 * g + h, 2);
                    }
                    if (leader) umma_commit(bar_empty + 8 * s);
                }
                if (leader) umma_commit(bar_afree);
            }
        }
    } else if (REARM && warp >= 20) {
        // re-arming warps, one per TMEM lane quadrant: when the quadrant's four draining warps have an accumulator in registers
        // the constant goes back over its 240 columns and the accumulator returns to the UMMA warp
        const uint32_t quad = warp & 3;
        uint32_t cst[16];
#pragma unroll
        for (int i = 0; i < 16; i++) cst[i] = *reinterpret_cast<volatile uint32_t *>(&s_cst[i]);
        const uint32_t lane_base = __shfl_sync(0xffffffffu, acc0 + ((quad * 32u) << 16), 0);
        uint32_t g = 0;
        for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
            for (uint32_t j = 0; j < ntiles; j++, g++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t t0 = lane_base + h * T4_NCOLS;
                    TC_TRACE(trole, 2 * g + h, 0);
                    mbar_wait(bar_tread + 8 * h, g & 1);
                    tc_fence_after();
                    TC_TRACE(trole, 2 * g + h, 1);
                    if (DRAIN == 9) {
#pragma unroll
                        for (int c = 0; c < 224; c += 32) tmem_st32_unpack16(t0 + c, cst);
                        tmem_st16_unpack16(t0 + 224, cst);
                    } else {
#pragma unroll
                        for (int c = 0; c < 240; c += 16) tmem_st16(t0 + c, cst);
                    }
                    tmem_wait_st();
                    TC_TRACE(trole, 2 * g + h, 2);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8 * h);
                }
            }
        }
    } else if (warp >= 4) {
        // 16 draining warps = 4 TMEM lane quadrants x 4 column parts; every warp takes its part of BOTH accumulators
        // (a warp may read any column of its own quadrant), so a part is at most 64 columns: it is loaded into
        // registers in one go and the accumulator is handed back to the MMA warp before any of it is reduced.
        const uint32_t ew = warp - 4;
        const uint32_t quad = warp & 3, cp = ew >> 2;
        // parts of the 240 columns: [0,64) [64,128) [128,192) [192,240) — the last one is 32 + 16
        const uint32_t c0 = cp * 64;
        const bool wide = cp != 3;
        const uint32_t cw = wide ? 64 : T4_NCOLS - 192;
        uint32_t raw0[32], raw1[32];
        uint32_t g = 0;
        if (DRAIN == 4) {
            const uint32_t part6 = ew >> 2;                 // 0..5: columns [40 part6, 40 part6 + 40) of both accumulators
            const uint32_t pc0 = part6 * 40u;
            uint32_t(&ra)[16] = *reinterpret_cast<uint32_t(*)[16]>(raw0);
            uint32_t(&rb)[16] = *reinterpret_cast<uint32_t(*)[16]>(raw1);
            uint32_t(&rc)[8] = *reinterpret_cast<uint32_t(*)[8]>(&raw0[16]);
            const uint32_t lane_base = acc0 + ((quad * 32u) << 16) + pc0;
            for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
                const uint32_t p = u / qblocks, qb = (u % qblocks) * TC_QROWS + quad * 32 + lane;
                float r0[2] = {TC_KEY_NONE, TC_KEY_NONE}, r1[2] = {TC_KEY_NONE, TC_KEY_NONE};
                float tbase = (float)pc0 + TC_KEY_BIAS;
                for (uint32_t j = 0; j < ntiles; j++, g++, tbase += (float)T4_NCOLS) {
                    const uint32_t tile0 = j * T4_NCOLS + pc0;
                    const bool dead = (dbg & 2) || tile0 >= n2;
                    const bool masked = tile0 + 40u > n2;
                    const uint32_t nvalid = dead ? 0 : n2 - tile0;
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const uint32_t taddr = lane_base + h * T4_NCOLS;
                        mbar_wait(bar_tfull + 8 * h, g & 1);
                        tc_fence_after();
                        if (!dead) {
                            tmem_ld16(taddr, ra);
                            tmem_ld16(taddr + 16, rb);
                            tmem_ld8(taddr + 32, rc);
                            tmem_wait_ld_regs16(ra);
                            tmem_wait_ld_regs16(rb);
                            tmem_wait_ld_regs8(rc);
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tempty + 8 * h);
                        if (dead) continue;
                        if (!masked) {
                            drain_chunk<0, false, 16>(ra, nvalid, tbase, r0[h], r1[h]);
                            drain_chunk<16, false, 16>(rb, nvalid, tbase, r0[h], r1[h]);
                            drain_chunk<32, false, 8>(rc, nvalid, tbase, r0[h], r1[h]);
                        } else {
                            drain_chunk<0, true, 16>(ra, nvalid, tbase, r0[h], r1[h]);
                            drain_chunk<16, true, 16>(rb, nvalid, tbase, r0[h], r1[h]);
                            drain_chunk<32, true, 8>(rc, nvalid, tbase, r0[h], r1[h]);
                        }
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t q = qb + h * 128;
                    if (q < n1) {
                        uint32_t out[2];
                        const float ks[2] = {r0[h], r1[h]};
#pragma unroll
                        for (int i = 0; i < 2; i++) {
                            const uint32_t ki = __float2uint_rz(ks[i]);
                            out[i] = ks[i] < 16777216.f ? ((ki >> (TC_KEY_SHIFT + 1)) << KNN_IDX_BITS) | (ki & (TC_MAX_TRAIN - 1u))
                                                        : 0xffffffffu;
                        }
                        part[((size_t)p * 6 + part6) * n1 + q] = make_uint2(out[0], out[1]);
                    }
                }
            }
        } else if (DRAIN == 3) {
            // Dedicated warps: of the four draining warps on a scheduler (= TMEM lane quadrant) two serve accumulator 0
            // and two accumulator 1, each taking 120 of its accumulator's 240 columns in 16-column slices (the register
            // scoreboard lets slice s+1 load while slice s is reduced). The two accumulators fill half a step apart, so the
            // scheduler's TMEM port and ALU pipe are reading / reducing one accumulator while the tensor pipe writes the
            // other, instead of every warp loading both accumulators and then reducing both in lock-step.
            const uint32_t acc = (ew >> 2) & 1u, halfc = ew >> 3;          // which accumulator, which 120-column half
            const uint32_t hc0 = halfc * 120u;
            uint32_t(&ra)[16] = *reinterpret_cast<uint32_t(*)[16]>(raw0);
            uint32_t(&rb)[16] = *reinterpret_cast<uint32_t(*)[16]>(raw1);
            uint32_t(&rl)[8] = *reinterpret_cast<uint32_t(*)[8]>(raw1);
            const uint32_t taddr = acc0 + ((quad * 32u) << 16) + acc * T4_NCOLS + hc0;
            for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
                const uint32_t p = u / qblocks, q = (u % qblocks) * TC_QROWS + acc * 128 + quad * 32 + lane;
                float r0 = TC_KEY_NONE, r1 = TC_KEY_NONE;
                float tbase = (float)hc0 + TC_KEY_BIAS;
                for (uint32_t j = 0; j < ntiles; j++, g++, tbase += (float)T4_NCOLS) {
                    const uint32_t tile0 = j * T4_NCOLS + hc0;
                    const bool dead = (dbg & 2) || tile0 >= n2;
                    const bool masked = tile0 + 120u > n2;
                    const uint32_t nvalid = dead ? 0 : n2 - tile0;
                    mbar_wait(bar_tfull + 8 * acc, g & 1);
                    tc_fence_after();
                    if (dead) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                        continue;
                    }
#define VB_SLICE(C, BUF, NEXT_LD)                                                                   \
                    tmem_wait_ld_regs16(BUF);                                                       \
                    NEXT_LD;                                                                        \
                    if (masked) drain_chunk<C, true, 16>(BUF, nvalid, tbase, r0, r1);               \
                    else drain_chunk<C, false, 16>(BUF, nvalid, tbase, r0, r1);
                    tmem_ld16(taddr, ra);
                    VB_SLICE(0, ra, tmem_ld16(taddr + 16, rb))
                    VB_SLICE(16, rb, tmem_ld16(taddr + 32, ra))
                    VB_SLICE(32, ra, tmem_ld16(taddr + 48, rb))
                    VB_SLICE(48, rb, tmem_ld16(taddr + 64, ra))
                    VB_SLICE(64, ra, tmem_ld16(taddr + 80, rb))
                    VB_SLICE(80, rb, tmem_ld16(taddr + 96, ra))
                    // last two slices: 16 + 8 columns; the accumulator goes back once both have landed
                    tmem_wait_ld_regs16(ra);
                    tmem_ld8(taddr + 112, rl);
                    if (masked) drain_chunk<96, true, 16>(ra, nvalid, tbase, r0, r1);
                    else drain_chunk<96, false, 16>(ra, nvalid, tbase, r0, r1);
                    tmem_wait_ld_regs8(rl);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                    if (masked) drain_chunk<112, true, 8>(rl, nvalid, tbase, r0, r1);
                    else drain_chunk<112, false, 8>(rl, nvalid, tbase, r0, r1);
#undef VB_SLICE
                }
                if (q < n1) {
                    uint32_t out[2];
                    const float ks[2] = {r0, r1};
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        const uint32_t ki = __float2uint_rz(ks[i]);
                        out[i] = ks[i] < 16777216.f ? ((ki >> (TC_KEY_SHIFT + 1)) << KNN_IDX_BITS) | (ki & (TC_MAX_TRAIN - 1u))
                                                    : 0xffffffffu;
                    }
                    part[((size_t)p * 2 + halfc) * n1 + q] = make_uint2(out[0], out[1]);
                }
            }
        } else if (DRAIN == 5) {
            // Variant 3's dedicated warps with 24 draining warps: on every scheduler three warps serve accumulator 0 and three
            // accumulator 1, each taking 80 of its accumulator's 240 columns as five 16-column slices through two rotating
            // buffers. The two groups run half a step apart (their accumulators fill half a step apart), so one group's
            // barrier wait and first TMEM load fall into the other group's reduction; with two warps per group (variant 3)
            // the scheduler had too little independent work while one group waited.
            const uint32_t acc = (ew >> 2) & 1u, third = ew >> 3;          // which accumulator, which 80-column third
            const uint32_t hc0 = third * 80u;
            uint32_t(&ra)[16] = *reinterpret_cast<uint32_t(*)[16]>(raw0);
            uint32_t(&rb)[16] = *reinterpret_cast<uint32_t(*)[16]>(raw1);
            const uint32_t taddr = acc0 + ((quad * 32u) << 16) + acc * T4_NCOLS + hc0;
            for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
                const uint32_t p = u / qblocks, q = (u % qblocks) * TC_QROWS + acc * 128 + quad * 32 + lane;
                float r0 = TC_KEY_NONE, r1 = TC_KEY_NONE;
                float tbase = (float)hc0 + TC_KEY_BIAS;
                for (uint32_t j = 0; j < ntiles; j++, g++, tbase += (float)T4_NCOLS) {
                    const uint32_t tile0 = j * T4_NCOLS + hc0;
                    const bool dead = (dbg & 2) || tile0 >= n2;
                    const bool masked = tile0 + 80u > n2;
                    const uint32_t nvalid = dead ? 0 : n2 - tile0;
                    mbar_wait(bar_tfull + 8 * acc, g & 1);
                    tc_fence_after();
                    if (dead) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                        continue;
                    }
#define VB_SLICE(C, BUF, NEXT_LD)                                                                   \
                    tmem_wait_ld_regs16(BUF);                                                       \
                    NEXT_LD;                                                                        \
                    if (masked) drain_chunk<C, true, 16>(BUF, nvalid, tbase, r0, r1);               \
                    else drain_chunk<C, false, 16>(BUF, nvalid, tbase, r0, r1);
                    tmem_ld16(taddr, ra);
                    VB_SLICE(0, ra, tmem_ld16(taddr + 16, rb))
                    VB_SLICE(16, rb, tmem_ld16(taddr + 32, ra))
                    VB_SLICE(32, ra, tmem_ld16(taddr + 48, rb))
                    VB_SLICE(48, rb, tmem_ld16(taddr + 64, ra))
                    // the accumulator goes back once the last slice has landed
                    tmem_wait_ld_regs16(ra);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                    if (masked) drain_chunk<64, true, 16>(ra, nvalid, tbase, r0, r1);
                    else drain_chunk<64, false, 16>(ra, nvalid, tbase, r0, r1);
#undef VB_SLICE
                }
                if (q < n1) {
                    uint32_t out[2];
                    const float ks[2] = {r0, r1};
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        const uint32_t ki = __float2uint_rz(ks[i]);
                        out[i] = ks[i] < 16777216.f ? ((ki >> (TC_KEY_SHIFT + 1)) << KNN_IDX_BITS) | (ki & (TC_MAX_TRAIN - 1u))
                                                    : 0xffffffffu;
                    }
                    part[((size_t)p * 3 + third) * n1 + q] = make_uint2(out[0], out[1]);
                }
            }
        } else if (DRAIN == 6 || DRAIN == 10) {
            uint32_t(&ra)[16] = *reinterpret_cast<uint32_t(*)[16]>(raw0);   // columns [0, 32) of the part, two per register
            uint32_t(&rb)[16] = *reinterpret_cast<uint32_t(*)[16]>(raw1);   // columns [32, 64) (narrow part: [16, 48))
            // warp-uniform by construction; the shuffle says so to the compiler (TMEM addresses live in uniform registers)
            const uint32_t lane_base = __shfl_sync(0xffffffffu, acc0 + ((quad * 32u) << 16) + c0, 0);
            // the constant, 8 times, read through volatile shared-memory loads: values ptxas cannot re-create, so the vector
            // stays in 8 registers (a known constant is rebuilt with moves in front of every store)
            uint32_t cst[8];
#pragma unroll
            for (int i = 0; i < 8; i++) cst[i] = *reinterpret_cast<volatile uint32_t *>(&s_cst[i]);
            // tiles whose columns [c0, c0 + cw) are all real train descriptors
            uint32_t nfull = (n2 >= c0 + cw) ? (n2 - c0 - cw) / (uint32_t)T4_NCOLS + 1u : 0u;
            if (nfull > ntiles) nfull = ntiles;
            if (dbg & ~(32 | 4)) nfull = 0;   // the timing switches live in the general step
            // Variant 10 merges the four column parts of a row before they leave the SM: parts 1..3 of a quadrant put their two
            // keys in shared memory, the four warps meet at a named barrier (once per unit, 42 steps), and part 0's warp writes ONE
            // key pair per row — a quarter of the matcher's output and of what the fix pass reads, whose merge loop disappears.
            constexpr bool MERGE = DRAIN == 10;
            uint32_t ulocal = 0;
            for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x, ulocal++) {
                const uint32_t p = u / qblocks, qb = (u % qblocks) * TC_QROWS + quad * 32 + lane;
                uint32_t r0[2] = {0u, 0u}, r1[2] = {0u, 0u};   // per row half: packed (odd-column group, even-column group) keys
                if (wide) t6_unit<true, DRAIN == 10>(lane_base, bar_tfull, bar_tempty, g, lane, c0, cw, n2, ntiles, nfull, ra, rb, cst, r0, r1, dbg, trole);
                else t6_unit<false, DRAIN == 10>(lane_base, bar_tfull, bar_tempty, g, lane, c0, cw, n2, ntiles, nfull, ra, rb, cst, r0, r1, dbg, trole);
                uint2 mine[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    // four group keys (even / odd columns x best / second): the two smallest go to the fix kernel
                    const uint32_t ka = t6_group_key(r0[h] & 0xffffu, c0, 0u), kb = t6_group_key(r0[h] >> 16, c0, 1u);
                    const uint32_t kc = t6_group_key(r1[h] & 0xffffu, c0, 0u), kd = t6_group_key(r1[h] >> 16, c0, 1u);
                    const uint32_t lo1 = min(ka, kb), hi1 = max(ka, kb), lo2 = min(kc, kd);
                    mine[h] = make_uint2(lo1, min(hi1, lo2));   // ka < kc and kb < kd
                    const uint32_t q = qb + h * 128;
                    if (!MERGE && q < n1) part[((size_t)p * T4_PARTS + cp) * n1 + q] = mine[h];
                }
                if (MERGE) {
                    if (cp != 0) {
                        s_keys[ulocal & 1][quad][cp - 1][0][lane] = mine[0];
                        s_keys[ulocal & 1][quad][cp - 1][1][lane] = mine[1];
                    }
                    // the quadrant's four draining warps (immediate barrier ids: with a register id ptxas reserves all sixteen
                    // hardware barriers of the SM and nothing that calls __syncthreads could run beside this kernel)
                    if (quad == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
                    else if (quad == 1) asm volatile("bar.sync 2, 128;" ::: "memory");
                    else if (quad == 2) asm volatile("bar.sync 3, 128;" ::: "memory");
                    else asm volatile("bar.sync 4, 128;" ::: "memory");
                    if (cp == 0) {
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            uint32_t k1 = mine[h].x, k2 = mine[h].y;
#pragma unroll
                            for (int sp = 0; sp < 3; sp++) {
                                const uint2 v = s_keys[ulocal & 1][quad][sp][h][lane];
                                const uint32_t lo = min(k1, v.x), hi = max(k1, v.x);   // v.x < v.y and k1 < k2
                                k2 = min(min(k2, v.y), hi);
                                k1 = lo;
                            }
                            const uint32_t q = qb + h * 128;
                            if (q < n1) part[(size_t)p * n1 + q] = make_uint2(k1, k2);
                        }
                    }
                }
            }
        } else if (REARM) {
            uint32_t(&ra)[16] = *reinterpret_cast<uint32_t(*)[16]>(raw0);
            uint32_t(&rb)[16] = *reinterpret_cast<uint32_t(*)[16]>(raw1);
            const uint32_t lane_base = __shfl_sync(0xffffffffu, acc0 + ((quad * 32u) << 16) + c0, 0);
            uint32_t nfull = (n2 >= c0 + cw) ? (n2 - c0 - cw) / (uint32_t)T4_NCOLS + 1u : 0u;
            if (nfull > ntiles) nfull = ntiles;
            for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
                const uint32_t p = u / qblocks, qb = (u % qblocks) * TC_QROWS + quad * 32 + lane;
                uint32_t r0[2] = {0u, 0u}, r1[2] = {0u, 0u};
                if (wide) t8_unit<true>(lane_base, bar_tfull, bar_tread, g, lane, c0, cw, n2, ntiles, nfull, ra, rb, r0, r1, trole);
                else t8_unit<false>(lane_base, bar_tfull, bar_tread, g, lane, c0, cw, n2, ntiles, nfull, ra, rb, r0, r1, trole);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t q = qb + h * 128;
                    if (q < n1) {
                        const uint32_t ka = t6_group_key(r0[h] & 0xffffu, c0, 0u), kb = t6_group_key(r0[h] >> 16, c0, 1u);
                        const uint32_t kc = t6_group_key(r1[h] & 0xffffu, c0, 0u), kd = t6_group_key(r1[h] >> 16, c0, 1u);
                        const uint32_t lo1 = min(ka, kb), hi1 = max(ka, kb), lo2 = min(kc, kd);
                        part[((size_t)p * T4_PARTS + cp) * n1 + q] = make_uint2(lo1, min(hi1, lo2));
                    }
                }
            }
        } else if (DRAIN == 7) {
            // The packed drain with variant 5's geometry: 24 draining warps, on every scheduler three serve accumulator 0 and
            // three accumulator 1, 80 columns (five spans) each. The float drain gained nothing from the split because its ALU
            // work alone filled the pipe; the packed drain needs half of it, so one group's TMEM round trip (load, store of the
            // constant, hand-back) can hide behind the other group's max trees.
            const uint32_t acc = (ew >> 2) & 1u, third = ew >> 3;
            const uint32_t hc0 = third * 80u;
            uint32_t(&ra)[16] = *reinterpret_cast<uint32_t(*)[16]>(raw0);            // columns [0, 32)
            uint32_t(&rb)[16] = *reinterpret_cast<uint32_t(*)[16]>(raw1);            // columns [32, 64)
            uint32_t(&rc)[8] = *reinterpret_cast<uint32_t(*)[8]>(&raw0[16]);         // columns [64, 80)
            uint32_t cst[8];
#pragma unroll
            for (int i = 0; i < 8; i++) cst[i] = *reinterpret_cast<volatile uint32_t *>(&s_cst[i]);
            const uint32_t taddr = acc0 + ((quad * 32u) << 16) + acc * T4_NCOLS + hc0;
            for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
                const uint32_t p = u / qblocks, q = (u % qblocks) * TC_QROWS + acc * 128 + quad * 32 + lane;
                uint32_t r0 = 0u, r1 = 0u;
                uint32_t posc = 127u * 0x00010001u;
                for (uint32_t j = 0; j < ntiles; j++, g++, posc -= 5u * 0x00010001u) {
                    const uint32_t tile0 = j * T4_NCOLS + hc0;
                    const bool skip = (dbg & 2) || tile0 >= n2;
                    const bool masked = tile0 + 80u > n2;
                    const uint32_t nvalid = skip ? 0 : n2 - tile0;
                    mbar_wait(bar_tfull + 8 * acc, g & 1);
                    tc_fence_after();
                    if (!skip) {
                        tmem_ld32_pack16(taddr, ra);
                        tmem_ld32_pack16(taddr + 32, rb);
                        tmem_ld16_pack16(taddr + 64, rc);
                        tmem_wait_ld_regs16(ra);
                    }
#pragma unroll
                    for (int c = 0; c < 32; c += 8) tmem_st8(taddr + c, cst);
                    if (!skip) tmem_wait_ld_regs16(rb);
#pragma unroll
                    for (int c = 32; c < 64; c += 8) tmem_st8(taddr + c, cst);
                    if (!skip) tmem_wait_ld_regs8(rc);
                    tmem_st8(taddr + 64, cst);
                    tmem_st8(taddr + 72, cst);
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                    if (skip) continue;
                    if (masked) {
                        drain_span16<true>(&ra[0], 0u, nvalid, posc, r0, r1);
                        drain_span16<true>(&ra[8], 16u, nvalid, posc - 0x00010001u, r0, r1);
                        drain_span16<true>(&rb[0], 32u, nvalid, posc - 2u * 0x00010001u, r0, r1);
                        drain_span16<true>(&rb[8], 48u, nvalid, posc - 3u * 0x00010001u, r0, r1);
                        drain_span16<true>(&rc[0], 64u, nvalid, posc - 4u * 0x00010001u, r0, r1);
                    } else {
                        drain_span16<false>(&ra[0], 0u, nvalid, posc, r0, r1);
                        drain_span16<false>(&ra[8], 16u, nvalid, posc - 0x00010001u, r0, r1);
                        drain_span16<false>(&rb[0], 32u, nvalid, posc - 2u * 0x00010001u, r0, r1);
                        drain_span16<false>(&rb[8], 48u, nvalid, posc - 3u * 0x00010001u, r0, r1);
                        drain_span16<false>(&rc[0], 64u, nvalid, posc - 4u * 0x00010001u, r0, r1);
                    }
                }
                if (q < n1) {
                    const uint32_t ka = t6_group_key<5>(r0 & 0xffffu, hc0, 0u), kb = t6_group_key<5>(r0 >> 16, hc0, 1u);
                    const uint32_t kc = t6_group_key<5>(r1 & 0xffffu, hc0, 0u), kd = t6_group_key<5>(r1 >> 16, hc0, 1u);
                    const uint32_t lo1 = min(ka, kb), hi1 = max(ka, kb), lo2 = min(kc, kd);
                    part[((size_t)p * 3 + third) * n1 + q] = make_uint2(lo1, min(hi1, lo2));
                }
            }
        } else if (DRAIN == 2) {
            // Slices as in DRAIN 1, and the pipeline also runs ACROSS steps: as soon as a step's last slice has landed the
            // accumulator goes back to the MMA warp, the first slice of the next step (the other accumulator) is requested,
            // and only then is the last slice reduced — the TMEM port never waits for the ALU pipe inside a unit.
            // Every part issues four slice loads per step so that the two buffers keep their roles; the fourth slice of the
            // narrow part (columns 192..239 = three slices) re-reads its third and is not reduced.
            uint32_t(&ra)[16] = *reinterpret_cast<uint32_t(*)[16]>(raw0);
            uint32_t(&rb)[16] = *reinterpret_cast<uint32_t(*)[16]>(raw1);
            const uint32_t lane_base = acc0 + ((quad * 32u) << 16) + c0;
            const uint32_t last_off = wide ? 48u : 32u;
            for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
                const uint32_t p = u / qblocks, qb = (u % qblocks) * TC_QROWS + quad * 32 + lane;
                float r0[2] = {TC_KEY_NONE, TC_KEY_NONE}, r1[2] = {TC_KEY_NONE, TC_KEY_NONE};
                float tbase = (float)c0 + TC_KEY_BIAS;
                mbar_wait(bar_tfull, g & 1);   // first step of the unit: accumulator 0 of tile 0
                tc_fence_after();
                tmem_ld16(lane_base, ra);
                for (uint32_t j = 0; j < ntiles; j++, tbase += (float)T4_NCOLS) {
                    const uint32_t tile0 = j * T4_NCOLS + c0;
                    const bool dead = (dbg & 2) || tile0 >= n2;
                    const bool masked = tile0 + cw > n2;
                    const uint32_t nvalid = dead ? 0 : n2 - tile0;
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const uint32_t taddr = lane_base + h * T4_NCOLS;
                        tmem_wait_ld_regs16(ra);
                        tmem_ld16(taddr + 16, rb);
                        if (!dead) {
                            if (masked) drain_chunk<0, true, 16>(ra, nvalid, tbase, r0[h], r1[h]);
                            else drain_chunk<0, false, 16>(ra, nvalid, tbase, r0[h], r1[h]);
                        }
                        tmem_wait_ld_regs16(rb);
                        tmem_ld16(taddr + 32, ra);
                        if (!dead) {
                            if (masked) drain_chunk<16, true, 16>(rb, nvalid, tbase, r0[h], r1[h]);
                            else drain_chunk<16, false, 16>(rb, nvalid, tbase, r0[h], r1[h]);
                        }
                        tmem_wait_ld_regs16(ra);
                        tmem_ld16(taddr + last_off, rb);
                        if (!dead) {
                            if (masked) drain_chunk<32, true, 16>(ra, nvalid, tbase, r0[h], r1[h]);
                            else drain_chunk<32, false, 16>(ra, nvalid, tbase, r0[h], r1[h]);
                        }
                        tmem_wait_ld_regs16(rb);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tempty + 8 * h);
                        if (h == 1) g++;
                        if (h == 0 || j + 1 < ntiles) {   // next step: the other accumulator (this tile's, or the next tile's)
                            mbar_wait(bar_tfull + 8 * (h ^ 1), g & 1);
                            tc_fence_after();
                            tmem_ld16(lane_base + (h ^ 1) * T4_NCOLS, ra);
                        }
                        if (wide && !dead) {
                            if (masked) drain_chunk<48, true, 16>(rb, nvalid, tbase, r0[h], r1[h]);
                            else drain_chunk<48, false, 16>(rb, nvalid, tbase, r0[h], r1[h]);
                        }
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t q = qb + h * 128;
                    if (q < n1) {
                        uint32_t out[2];
                        const float ks[2] = {r0[h], r1[h]};
#pragma unroll
                        for (int i = 0; i < 2; i++) {
                            const uint32_t ki = __float2uint_rz(ks[i]);
                            out[i] = ks[i] < 16777216.f ? ((ki >> (TC_KEY_SHIFT + 1)) << KNN_IDX_BITS) | (ki & (TC_MAX_TRAIN - 1u))
                                                        : 0xffffffffu;
                        }
                        part[((size_t)p * T4_PARTS + cp) * n1 + q] = make_uint2(out[0], out[1]);
                    }
                }
            }
        } else
        for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
            const uint32_t p = u / qblocks, qb = (u % qblocks) * TC_QROWS + quad * 32 + lane;
            float r0[2] = {TC_KEY_NONE, TC_KEY_NONE}, r1[2] = {TC_KEY_NONE, TC_KEY_NONE};   // per row half
            float tbase = (float)c0 + TC_KEY_BIAS;
            for (uint32_t j = 0; j < ntiles; j++, g++, tbase += (float)T4_NCOLS) {
                const uint32_t tile0 = j * T4_NCOLS + c0;
                const bool skip = (dbg & 2) || tile0 >= n2;   // (dbg 2: MMA floor measurement) / nothing valid in this part
                const bool masked = tile0 + cw > n2;
                const uint32_t nvalid = skip ? 0 : n2 - tile0;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t taddr = acc0 + ((quad * 32u) << 16) + h * T4_NCOLS + c0;
                    mbar_wait(bar_tfull + 8 * h, g & 1);
                    tc_fence_after();
                    if (DRAIN == 1) {
                        if (skip) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(bar_tempty + 8 * h);
                            continue;
                        }
                        uint32_t(&ra)[16] = *reinterpret_cast<uint32_t(*)[16]>(raw0);
                        uint32_t(&rb)[16] = *reinterpret_cast<uint32_t(*)[16]>(raw1);
                        // slices: [0,16) [16,32) [32,48) and, for the three wide parts, [48,64)
                        tmem_ld16(taddr, ra);
                        tmem_wait_ld_regs16(ra);
                        tmem_ld16(taddr + 16, rb);
                        if (masked) drain_chunk<0, true, 16>(ra, nvalid, tbase, r0[h], r1[h]);
                        else drain_chunk<0, false, 16>(ra, nvalid, tbase, r0[h], r1[h]);
                        tmem_wait_ld_regs16(rb);
                        tmem_ld16(taddr + 32, ra);
                        if (masked) drain_chunk<16, true, 16>(rb, nvalid, tbase, r0[h], r1[h]);
                        else drain_chunk<16, false, 16>(rb, nvalid, tbase, r0[h], r1[h]);
                        tmem_wait_ld_regs16(ra);
                        if (wide) {
                            tmem_ld16(taddr + 48, rb);
                            if (masked) drain_chunk<32, true, 16>(ra, nvalid, tbase, r0[h], r1[h]);
                            else drain_chunk<32, false, 16>(ra, nvalid, tbase, r0[h], r1[h]);
                            tmem_wait_ld_regs16(rb);
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tempty + 8 * h);
                        if (wide) {
                            if (masked) drain_chunk<48, true, 16>(rb, nvalid, tbase, r0[h], r1[h]);
                            else drain_chunk<48, false, 16>(rb, nvalid, tbase, r0[h], r1[h]);
                        } else {
                            if (masked) drain_chunk<32, true, 16>(ra, nvalid, tbase, r0[h], r1[h]);
                            else drain_chunk<32, false, 16>(ra, nvalid, tbase, r0[h], r1[h]);
                        }
                        continue;
                    }
                    if (!skip) {
                        tmem_ld32(taddr, raw0);
                        if (wide) {
                            tmem_ld32(taddr + 32, raw1);
                            tmem_wait_ld_regs(raw0);
                            tmem_wait_ld_regs(raw1);
                        } else {
                            tmem_ld16(taddr + 32, *reinterpret_cast<uint32_t(*)[16]>(raw1));
                            tmem_wait_ld_regs(raw0);
                            tmem_wait_ld_regs16(*reinterpret_cast<uint32_t(*)[16]>(raw1));
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8 * h);
                    if (skip) continue;
                    if (masked) {
                        drain_chunk<0, true>(raw0, nvalid, tbase, r0[h], r1[h]);
                        if (wide) drain_chunk<32, true>(raw1, nvalid, tbase, r0[h], r1[h]);
                        else drain_chunk<32, true, 16>(*reinterpret_cast<uint32_t(*)[16]>(raw1), nvalid, tbase, r0[h], r1[h]);
                    } else {
                        drain_chunk<0, false>(raw0, nvalid, tbase, r0[h], r1[h]);
                        if (wide) drain_chunk<32, false>(raw1, nvalid, tbase, r0[h], r1[h]);
                        else drain_chunk<32, false, 16>(*reinterpret_cast<uint32_t(*)[16]>(raw1), nvalid, tbase, r0[h], r1[h]);
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t q = qb + h * 128;
                if (q < n1) {
                    uint32_t out[2];
                    const float ks[2] = {r0[h], r1[h]};
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        const uint32_t ki = __float2uint_rz(ks[i]);
                        out[i] = ks[i] < 16777216.f ? ((ki >> (TC_KEY_SHIFT + 1)) << KNN_IDX_BITS) | (ki & (TC_MAX_TRAIN - 1u))
                                                    : 0xffffffffu;
                    }
                    part[((size_t)p * T4_PARTS + cp) * n1 + q] = make_uint2(out[0], out[1]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// LANES lanes per query: merge the column parts into (best group, best other group) and evaluate candidates with XOR + POPC
// on the original descriptors. out[p][q] = final (best, second) keys.
//   LANES = 16: the 16 descriptors of both 8-column groups, their two smallest (distance, index) keys — knnMatch's pair,
//               second index included (the vb_knn2_hamming entry points).
//   LANES = 8:  the best group only. A group key carries the group's best DISTANCE exactly (it is the accumulator's
//               maximum), so the second neighbour's distance is min(second smallest in the best group, the other group's
//               key distance) without reading that group; only the second neighbour's index stays unknown (the key keeps
//               the group's first column), and match_features (src/Frame.cpp:91-95) never looks at it. Half the bytes.
//   STRIDE = 2 (packed drain): a group is 8 columns of one parity, first column + 2 i. Two groups with the same best distance
//               are ordered by first column, which orders their members too — except for the even and the odd group of ONE
//               span, whose columns interleave: when those two tie, LANES = 8 evaluates both (the nearest neighbour is the
//               lower index of the two groups' best members, and the second distance is then that same distance).
template <int LANES, int STRIDE = 1>
__global__ void __launch_bounds__(256) k_knn2_tc_fix(const uint32_t *__restrict__ d1_base, const uint32_t *__restrict__ d2_base,
                                                     size_t stride_words, uint32_t n1, uint32_t n2, uint32_t nparts,
                                                     const uint2 *__restrict__ part, uint2 *__restrict__ out, double ratio) {
    static_assert(TC_GROUP == 8 && (LANES == 8 || LANES == 16), "groups of eight lanes per query");
    const uint32_t sub = threadIdx.x & (LANES - 1);
    const uint32_t q = blockIdx.x * (blockDim.x / LANES) + (threadIdx.x / LANES), p = blockIdx.y;
    const bool live = q < n1;
    uint32_t key = 0xffffffffu, other = 0xffffffffu;
    if (live) {
        uint32_t k1 = 0xffffffffu, k2 = 0xffffffffu;   // group keys: (best distance in the group) << 22 | first column
        for (uint32_t s = 0; s < nparts; s++) {
            const uint2 v = part[((size_t)p * nparts + s) * n1 + q];
            const uint32_t lo = min(k1, v.x), hi = max(k1, v.x);   // v.x < v.y and k1 < k2
            k2 = min(min(k2, v.y), hi);
            k1 = lo;
        }
        // ratio >= 0 (LANES = 8, the caller applies Lowe's test with this ratio and never looks at a failing query's indices):
        // the second neighbour is at most as far as the other group's best member, so a query that fails the test against
        // THAT distance — the very comparison k_knn2_finish makes, monotone in the second distance — fails it for certain;
        // its descriptors are not read and the group keys go out as they are (exact nearest distance, a bound for the second).
        const bool skip = LANES == 8 && ratio >= 0.0 && k2 != 0xffffffffu &&
                          !((double)(float)(k1 >> KNN_IDX_BITS) < (double)(float)(k2 >> KNN_IDX_BITS) * ratio);
        const uint32_t gk = skip ? 0xffffffffu : (sub < 8) ? k1 : k2;
        const uint32_t col = (gk & KNN_IDX_MASK) + (sub & 7) * STRIDE;
        if (LANES == 8) other = k2;
        const uint4 *a = reinterpret_cast<const uint4 *>(d1_base + (size_t)p * stride_words + (size_t)q * 8);
        if (gk != 0xffffffffu && col < n2) {
            const uint4 *b = reinterpret_cast<const uint4 *>(d2_base + (size_t)p * stride_words + (size_t)col * 8);
            const uint4 a0 = __ldg(a), a1 = __ldg(a + 1), b0 = __ldg(b), b1 = __ldg(b + 1);
            const uint32_t d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                               __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
            key = (d << KNN_IDX_BITS) | col;
        }
        if (skip) {
            key = k1;
            other = k2;
        }
        if (STRIDE == 2 && LANES == 8 && !skip && k2 != 0xffffffffu && (k2 >> KNN_IDX_BITS) == (k1 >> KNN_IDX_BITS) &&
            (k2 & KNN_IDX_MASK) == (k1 & KNN_IDX_MASK) + 1u) {
            // the odd group of the best group's span ties with it: its members interleave with the best group's
            const uint32_t col2 = (k2 & KNN_IDX_MASK) + (sub & 7) * 2u;
            if (col2 < n2) {
                const uint4 *b = reinterpret_cast<const uint4 *>(d2_base + (size_t)p * stride_words + (size_t)col2 * 8);
                const uint4 a0 = __ldg(a), a1 = __ldg(a + 1), b0 = __ldg(b), b1 = __ldg(b + 1);
                const uint32_t d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                                   __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
                const uint32_t key2 = (d << KNN_IDX_BITS) | col2;
                other = max(key, key2);   // this lane's other candidate competes for second place with its exact key
                key = min(key, key2);
            }
        }
    }
    uint32_t best = key;
#pragma unroll
    for (int o = 1; o < LANES; o <<= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    uint32_t second = (key == best) ? other : key;   // keys of distinct columns are distinct
#pragma unroll
    for (int o = 1; o < LANES; o <<= 1) second = min(second, __shfl_xor_sync(0xffffffffu, second, o));
    if (live && sub == 0) out[(size_t)p * n1 + q] = make_uint2(best, second);
}

static int make_map(CUtensorMap *m, const void *base, uint64_t rows) {
    return tc::make_map_2d(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, base, TC_KBYTES, rows, 128, TC_BOX_ROWS);
}

bool hamming_tc_eligible(const vb_ctx *ctx, const HammingPlan &pl) {
    if (pl.W != 8 || pl.n2 > TC_MAX_TRAIN) return false;   // the packed float key holds 14 index bits
    const long long force = ctx->opt("hamming_tc", -1);
    if (force >= 0) return force != 0;
    // below a few thousand distance tiles the popcount kernel's finer CTA granularity wins
    return (uint64_t)pl.P * pl.n1 * pl.n2 >= (1ull << 22);
}

// WS_KNN_PART = [P][TC_COLSPLIT][n1] column-part results followed by [P][n1] final keys; *final_part points at
// the latter, which k_knn2_finish reads as a single split.
static bool hamming_tc_use_fp4(const vb_ctx *ctx) { return ctx->opt("hamming_fp4", 1) != 0; }

int hamming_tc_launch(vb_ctx *ctx, const HammingPlan &pl, const uint32_t *d1, const uint32_t *d2, size_t stride_words,
                      const uint2 **final_part, bool need_second_index, double ratio) {
    const bool fp4 = hamming_tc_use_fp4(ctx);
    if (!(ctx->func_attr_done & 1u)) {   // a function attribute is per device: remembered per context, not per process
        VB_CUDA(cudaFuncSetAttribute(k_knn2_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
        VB_CUDA(cudaFuncSetAttribute(k_knn2_tc4<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T4_SMEM_BYTES));
        VB_CUDA(cudaFuncSetAttribute(k_knn2_tc4<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T4_SMEM_BYTES));
        VB_CUDA(cudaFuncSetAttribute(k_knn2_tc4<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T4_SMEM_BYTES));
        VB_CUDA(cudaFuncSetAttribute(k_knn2_tc4<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T4_SMEM_BYTES));
        VB_CUDA(cudaFuncSetAttribute(k_knn2_tc4<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T4_SMEM_BYTES));
        VB_CUDA(cudaFuncSetAttribute(k_knn2_tc4<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T4_SMEM_BYTES));
        VB_CUDA(cudaFuncSetAttribute(k_knn2_tc4<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T4_SMEM_BYTES));
        VB_CUDA(cudaFuncSetAttribute(k_knn2_tc4<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T4_SMEM_BYTES));
        VB_CUDA(cudaFuncSetAttribute(k_knn2_tc4<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T4_SMEM_BYTES));
        VB_CUDA(cudaFuncSetAttribute(k_knn2_tc4<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T4_SMEM_BYTES));
        VB_CUDA(cudaFuncSetAttribute(k_knn2_tc4<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T4_SMEM_BYTES));
        VB_CUDA((cudaFuncSetAttribute(k_knn2_tc4<6, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T4_SMEM_BYTES)));
        VB_CUDA((cudaFuncSetAttribute(k_knn2_tc4<9, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T4_SMEM_BYTES)));
        VB_CUDA((cudaFuncSetAttribute(k_knn2_tc4<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T4_SMEM_BYTES)));
        VB_CUDA((cudaFuncSetAttribute(k_knn2_tc4<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T4_SMEM_BYTES)));
        VB_CUDA((cudaFuncSetAttribute(k_knn2_tc4<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T4_SMEM_BYTES)));
        ctx->func_attr_done |= 1u;
    }
    const uint32_t P = pl.P, n1 = pl.n1, n2 = pl.n2;
    const bool seq = P > 1 && n1 == n2 && stride_words == (size_t)n1 * 8 && d2 == d1 + stride_words;
    const uint32_t rowbytes = fp4 ? T4_ROWBYTES : TC_KBYTES;
    int rc;
    const size_t rows_total = seq ? (size_t)(P + 1) * n1 : (size_t)P * ((size_t)n1 + n2);
    if ((rc = ctx->ws_ensure(WS_EXP, rows_total * rowbytes))) return rc;
    // 10 = the packed-integer drain on denormal accumulators, constant re-armed with unpack::16b stores (1.88 us per 5k x 5k
    // pair; 6, the same drain on a normal-range constant with full-width stores: 1.96-2.07; the float organisations 0-5:
    // 2.47-2.52). 6-10 and 1 are held to 72 registers, which leaves room for another submission's
    // kernels beside the matcher (stream.cu).
    int drain = (int)ctx->opt("tc_drain", 10);
    // the packed drain: 7 bits of position in a 16-bit key; and when the even and the odd group of one span tie for SECOND place
    // the group pair it hands over has the right distances but may miss the lower index — callers that want the second index
    // (vb_knn2_hamming; match_features never looks at it) take variant 1
    if (drain == 7 && div_up(n2, (uint32_t)T4_NCOLS) > T7_MAX_TILES) drain = 6;
    if (drain >= 6 && (div_up(n2, (uint32_t)T4_NCOLS) > T6_MAX_TILES || need_second_index)) drain = 1;
    const uint32_t nparts = fp4 ? (drain == 10 ? 1u : drain == 3 ? 2u : (drain == 5 || drain == 7) ? 3u : drain == 4 ? 6u : (uint32_t)T4_PARTS) : (uint32_t)TC_COLSPLIT;
    if ((rc = ctx->ws_ensure(WS_KNN_PART, (size_t)P * (nparts + 1) * n1 * sizeof(uint2)))) return rc;
    uint8_t *E = ctx->ws[WS_EXP].as<uint8_t>();
    uint8_t *Eq = E, *Et;
    auto expand = [&](const uint32_t *src, uint32_t rows, uint32_t frames, uint8_t *dst) {
        if (fp4) {
            k_expand_e2m1<<<dim3(div_up(rows * 8, 256), frames), 256, 0, ctx->stream>>>(src, stride_words, rows, frames, dst);
        } else {
            k_expand_pm1<<<(unsigned)div_up64((size_t)frames * rows * 16, 256), 256, 0, ctx->stream>>>(src, stride_words, rows,
                                                                                                        frames, dst);
        }
        ctx->launches++;
    };
    ctx->prof_begin("expand");
    if (seq) {
        Et = E + (size_t)n1 * rowbytes;
        expand(d1, n1, P + 1, E);
    } else {
        Et = E + (size_t)P * n1 * rowbytes;
        expand(d1, n1, P, Eq);
        expand(d2, n2, P, Et);
    }
    ctx->prof_end("expand");
    VB_CUDA(cudaGetLastError());
    CUtensorMap mq, mt;
    if (fp4) {
        if ((rc = make_map_2d(&mq, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, Eq, T4_ROWBYTES, (uint64_t)P * n1, 128, 128))) return rc;
        if ((rc = make_map_2d(&mt, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, Et, T4_ROWBYTES, (uint64_t)P * n2, 128, T4_NCOLS))) return rc;
    } else {
        if ((rc = make_map(&mq, Eq, (uint64_t)P * n1))) return rc;
        if ((rc = make_map(&mt, Et, (uint64_t)P * n2))) return rc;
    }
    const uint32_t nunits = div_up(n1, TC_QROWS) * P;
    const uint32_t grid = nunits < (uint32_t)ctx->sm_count ? nunits : (uint32_t)ctx->sm_count;   // one persistent CTA per SM
    uint2 *part = ctx->ws[WS_KNN_PART].as<uint2>();
    uint2 *fixed = part + (size_t)P * nparts * n1;
#ifdef VB_TUNING
    const int dbg = (int)ctx->opt("tc_dbg", 0);   // timing floors (results invalid): 2 = no drain, 4 = fp8 kernel without B loads
#else
    const int dbg = 0;
#endif
    ctx->prof_begin("hamming");
    const bool svc_hi = ctx->opt("tc_svc_hi", 0) != 0;
    const int issuers = (int)ctx->opt("tc_issuers", 1);
    if (fp4 && svc_hi && drain == 4)
        k_knn2_tc4<4, true><<<grid, T4_THREADS_WIDE, T4_SMEM_BYTES, ctx->stream>>>(mq, mt, n1, n2, n1, n2, nunits, part, dbg);
    else if (fp4 && svc_hi && drain == 1)
        k_knn2_tc4<1, true><<<grid, TC_THREADS, T4_SMEM_BYTES, ctx->stream>>>(mq, mt, n1, n2, n1, n2, nunits, part, dbg);
    else if (fp4 && svc_hi && drain == 0)
        k_knn2_tc4<0, true><<<grid, TC_THREADS, T4_SMEM_BYTES, ctx->stream>>>(mq, mt, n1, n2, n1, n2, nunits, part, dbg);
    else if (fp4 && drain == 4)
        k_knn2_tc4<4><<<grid, T4_THREADS_WIDE, T4_SMEM_BYTES, ctx->stream>>>(mq, mt, n1, n2, n1, n2, nunits, part, dbg);
    else if (fp4 && drain == 9 && issuers == 2)
        k_knn2_tc4<9, false, 2><<<grid, T8_THREADS, T4_SMEM_BYTES, ctx->stream>>>(mq, mt, n1, n2, n1, n2, nunits, part, dbg);
    else if (fp4 && drain == 6 && issuers == 2)
        k_knn2_tc4<6, false, 2><<<grid, TC_THREADS, T4_SMEM_BYTES, ctx->stream>>>(mq, mt, n1, n2, n1, n2, nunits, part, dbg);
    else if (fp4 && drain == 10)
        k_knn2_tc4<10><<<grid, TC_THREADS, T4_SMEM_BYTES, ctx->stream>>>(mq, mt, n1, n2, n1, n2, nunits, part, dbg);
    else if (fp4 && drain == 9)
        k_knn2_tc4<9><<<grid, T8_THREADS, T4_SMEM_BYTES, ctx->stream>>>(mq, mt, n1, n2, n1, n2, nunits, part, dbg);
    else if (fp4 && drain == 8)
        k_knn2_tc4<8><<<grid, T8_THREADS, T4_SMEM_BYTES, ctx->stream>>>(mq, mt, n1, n2, n1, n2, nunits, part, dbg);
    else if (fp4 && drain == 7)
        k_knn2_tc4<7><<<grid, T4_THREADS_WIDE, T4_SMEM_BYTES, ctx->stream>>>(mq, mt, n1, n2, n1, n2, nunits, part, dbg);
    else if (fp4 && drain == 6)
        k_knn2_tc4<6><<<grid, TC_THREADS, T4_SMEM_BYTES, ctx->stream>>>(mq, mt, n1, n2, n1, n2, nunits, part, dbg);
    else if (fp4 && drain == 5)
        k_knn2_tc4<5><<<grid, T4_THREADS_WIDE, T4_SMEM_BYTES, ctx->stream>>>(mq, mt, n1, n2, n1, n2, nunits, part, dbg);
    else if (fp4 && drain == 3)
        k_knn2_tc4<3><<<grid, TC_THREADS, T4_SMEM_BYTES, ctx->stream>>>(mq, mt, n1, n2, n1, n2, nunits, part, dbg);
    else if (fp4 && drain == 2)
        k_knn2_tc4<2><<<grid, TC_THREADS, T4_SMEM_BYTES, ctx->stream>>>(mq, mt, n1, n2, n1, n2, nunits, part, dbg);
    else if (fp4 && drain == 1)
        k_knn2_tc4<1><<<grid, TC_THREADS, T4_SMEM_BYTES, ctx->stream>>>(mq, mt, n1, n2, n1, n2, nunits, part, dbg);
    else if (fp4)
        k_knn2_tc4<0><<<grid, TC_THREADS, T4_SMEM_BYTES, ctx->stream>>>(mq, mt, n1, n2, n1, n2, nunits, part, dbg);
    else
        k_knn2_tc<<<grid, TC_THREADS, TC_SMEM_BYTES, ctx->stream>>>(mq, mt, n1, n2, n1, n2, nunits, part, dbg);
    ctx->prof_end("hamming");
    ctx->prof_begin("knnfix");
    const bool fix8 = ctx->opt("tc_fix8", 1) != 0;
    const double fix_ratio = (need_second_index || ctx->opt("tc_fix_skip", 1) == 0) ? -1.0 : ratio;   // < 0: every query is evaluated
    const bool stride2 = fp4 && drain >= 6;
    if ((need_second_index || !fix8) && stride2)
        k_knn2_tc_fix<16, 2><<<dim3(div_up(n1, 16), P), 256, 0, ctx->stream>>>(d1, d2, stride_words, n1, n2, nparts, part, fixed, fix_ratio);
    else if (need_second_index || !fix8)
        k_knn2_tc_fix<16><<<dim3(div_up(n1, 16), P), 256, 0, ctx->stream>>>(d1, d2, stride_words, n1, n2, nparts, part, fixed, fix_ratio);
    else if (stride2)
        k_knn2_tc_fix<8, 2><<<dim3(div_up(n1, 32), P), 256, 0, ctx->stream>>>(d1, d2, stride_words, n1, n2, nparts, part, fixed, fix_ratio);
    else
        k_knn2_tc_fix<8><<<dim3(div_up(n1, 32), P), 256, 0, ctx->stream>>>(d1, d2, stride_words, n1, n2, nparts, part, fixed, fix_ratio);
    ctx->prof_end("knnfix");
    ctx->launches += 2;
    VB_CUDA(cudaGetLastError());
    *final_part = fixed;
    return VB_OK;
}

}  // namespace vb

#if defined(VB_TUNING) && defined(VB_TC_TRACE)
// TUNING + TRACE builds only (not in include/vslam_b200.h): the timeline recorded by the last k_knn2_tc4 launch with tc_dbg & 32.
extern "C" int vb_debug_tc_trace(long long *out, int n) {
    const size_t total = sizeof(vb::g_tc_trace) / sizeof(long long);
    if (n < 0 || (size_t)n > total) return -1;
    return cudaMemcpyFromSymbol(out, vb::g_tc_trace, (size_t)n * sizeof(long long)) == cudaSuccess ? 0 : -2;
}
#endif
