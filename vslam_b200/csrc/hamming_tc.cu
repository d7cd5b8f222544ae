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
//   k_knn2_tc      CTA = 256 queries (two M = 128 accumulators of 256 columns = all 512 TMEM columns)
//                  against the whole train set in 256-wide tiles.
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

namespace vb {

using namespace tc;

constexpr int TC_THREADS = 384;
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
        o[i] = 0x38383838u | (spread << 7);                          // e4m3: 0x38 = +1.0, 0xB8 = -1.0
    }
    *reinterpret_cast<uint4 *>(dst + grow * TC_KBYTES + piece * 16) = make_uint4(o[0], o[1], o[2], o[3]);
}

struct Top2 {
    float d1, d2;      // largest and second largest dot (= smallest and second smallest distance)
    uint32_t i1, i2;
    __device__ __forceinline__ void offer(float v, uint32_t col) {
        if (v > d2) {   // strictly better than the current second; an equal dot never displaces an earlier index
            if (v > d1) { d2 = d1; i2 = i1; d1 = v; i1 = col; }
            else { d2 = v; i2 = col; }
        }
    }
};

template <bool MASKED>
__device__ __forceinline__ void drain_chunk(const uint32_t (&raw)[32], uint32_t col0, uint32_t n2, Top2 &t) {
#pragma unroll
    for (int g = 0; g < 32; g += 4) {
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            v[i] = __uint_as_float(raw[g + i]);
            if (MASKED) v[i] = (col0 + g + i < n2) ? v[i] : -1.0e30f;
        }
        const float m = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
        if (__any_sync(0xffffffffu, m > t.d2)) {
#pragma unroll
            for (int i = 0; i < 4; i++) t.offer(v[i], col0 + g + i);
        }
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
k_knn2_tc(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_t, uint32_t n1, uint32_t n2,
          uint32_t rowstride_q, uint32_t rowstride_t, uint2 *__restrict__ part) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzle-128B tiles need 1024-byte alignment
    const uint32_t sA = smem0;
    const uint32_t sB = smem0 + TC_A_BYTES;
    const uint32_t sBar = sB + 2 * TC_B_BYTES;
    const uint32_t bar_a = sBar, bar_full = sBar + 8, bar_empty = sBar + 24, bar_tfull = sBar + 40, bar_tempty = sBar + 56;
    const uint32_t s_tmem = sBar + 72;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t p = blockIdx.y;
    const uint32_t q0 = blockIdx.x * TC_QROWS;
    const uint32_t ntiles = (n2 + TC_NCOLS - 1) / TC_NCOLS;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_t);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(bar_a, 1);
        for (int s = 0; s < 2; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
            mbar_init(bar_tfull + 8 * s, 1);
            mbar_init(bar_tempty + 8 * s, 4);   // one arrival per draining warp
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(s_tmem));

    if (warp == 0) {
        if (lane == 0) {
            const int32_t qrow = (int32_t)(p * rowstride_q + q0);
            mbar_expect_tx(bar_a, TC_A_BYTES);
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int kk = 0; kk < 2; kk++)
                    tma_load_2d(sA + (h * 2 + kk) * TC_BOX_BYTES, &map_q, kk * 128, qrow + h * 128, bar_a);
            const int32_t trow = (int32_t)(p * rowstride_t);
            for (uint32_t j = 0; j < ntiles; j++) {
                const uint32_t s = j & 1, ph = (j >> 1) & 1;
                mbar_wait(bar_empty + 8 * s, ph ^ 1);
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
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(UMMA_FMT_E4M3, 128, TC_NCOLS);
            mbar_wait(bar_a, 0);
            for (uint32_t j = 0; j < ntiles; j++) {
                const uint32_t s = j & 1, ph = (j >> 1) & 1;
                mbar_wait(bar_full + 8 * s, ph);
                tc_fence_after();
                const uint32_t bbase = sB + s * TC_B_BYTES;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    mbar_wait(bar_tempty + 8 * h, (j & 1) ^ 1);   // accumulator h drained (tile j-1)
                    tc_fence_after();
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
                umma_commit(bar_empty + 8 * s);   // both halves have consumed this B stage
            }
        }
    } else if (warp >= 4) {
        const uint32_t h = (warp - 4) >> 2, quad = warp & 3;
        const uint32_t row = h * 128 + quad * 32 + lane;
        const uint32_t q = q0 + row;
        const uint32_t taddr = tmem_base + ((quad * 32u) << 16) + h * TC_NCOLS;
        Top2 t;
        t.d1 = t.d2 = -1.0e30f;
        t.i1 = t.i2 = KNN_IDX_MASK;
        for (uint32_t j = 0; j < ntiles; j++) {
            mbar_wait(bar_tfull + 8 * h, j & 1);
            tc_fence_after();
            const uint32_t tile0 = j * TC_NCOLS;
            const bool masked = tile0 + TC_NCOLS > n2;
#pragma unroll 1
            for (int c = 0; c < TC_NCOLS / 32; c++) {
                uint32_t raw[32];
                tmem_ld32(taddr + c * 32, raw);
                tmem_wait_ld();
                if (c == TC_NCOLS / 32 - 1) {   // the accumulator is in registers: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8 * h);
                }
                if (masked) drain_chunk<true>(raw, tile0 + c * 32, n2, t);
                else drain_chunk<false>(raw, tile0 + c * 32, n2, t);
            }
        }
        if (q < n1) {
            const uint32_t h1 = (uint32_t)((256 - __float2int_rn(t.d1)) >> 1);
            const uint32_t h2 = (uint32_t)((256 - __float2int_rn(t.d2)) >> 1);
            part[(size_t)p * n1 + q] = make_uint2((h1 << KNN_IDX_BITS) | t.i1, (h2 << KNN_IDX_BITS) | t.i2);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

static int make_map(CUtensorMap *m, const void *base, uint64_t rows) {
    const cuuint64_t gdim[2] = {(cuuint64_t)TC_KBYTES, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)TC_KBYTES};
    const cuuint32_t box[2] = {128, (cuuint32_t)TC_BOX_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = cuTensorMapEncodeTiled(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), gdim, gstride,
                                              box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        const char *s = nullptr;
        cuGetErrorString(r, &s);
        set_error("cuTensorMapEncodeTiled -> %s", s ? s : "?");
        return VB_ERR_CUDA;
    }
    return VB_OK;
}

bool hamming_tc_eligible(const HammingPlan &pl) {
    if (pl.W != 8 || pl.n2 > KNN_IDX_MASK) return false;
    if (const char *e = getenv("VB_HAMMING_TC")) return atoi(e) != 0;
    // below a few thousand distance tiles the popcount kernel's finer CTA granularity wins
    return (uint64_t)pl.P * pl.n1 * pl.n2 >= (1ull << 22);
}

// Fills WS_KNN_PART as [P][1 split][n1], the layout k_knn2_finish reads with nsplits = 1.
int hamming_tc_launch(vb_ctx *ctx, const HammingPlan &pl, const uint32_t *d1, const uint32_t *d2, size_t stride_words) {
    static bool attr_set = false;
    if (!attr_set) {
        VB_CUDA(cudaFuncSetAttribute(k_knn2_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
        attr_set = true;
    }
    const uint32_t P = pl.P, n1 = pl.n1, n2 = pl.n2;
    const bool seq = P > 1 && n1 == n2 && stride_words == (size_t)n1 * 8 && d2 == d1 + stride_words;
    int rc;
    const size_t rows_total = seq ? (size_t)(P + 1) * n1 : (size_t)P * ((size_t)n1 + n2);
    if ((rc = ctx->ws_ensure(WS_EXP, rows_total * TC_KBYTES))) return rc;
    if ((rc = ctx->ws_ensure(WS_KNN_PART, (size_t)P * n1 * sizeof(uint2)))) return rc;
    uint8_t *E = ctx->ws[WS_EXP].as<uint8_t>();
    uint8_t *Eq = E, *Et;
    ctx->prof_begin("expand");
    if (seq) {
        Et = E + (size_t)n1 * TC_KBYTES;
        const size_t thr = (size_t)(P + 1) * n1 * 16;
        k_expand_pm1<<<(unsigned)div_up64(thr, 256), 256, 0, ctx->stream>>>(d1, stride_words, n1, P + 1, E);
        ctx->launches++;
    } else {
        Et = E + (size_t)P * n1 * TC_KBYTES;
        k_expand_pm1<<<(unsigned)div_up64((size_t)P * n1 * 16, 256), 256, 0, ctx->stream>>>(d1, stride_words, n1, P, Eq);
        k_expand_pm1<<<(unsigned)div_up64((size_t)P * n2 * 16, 256), 256, 0, ctx->stream>>>(d2, stride_words, n2, P, Et);
        ctx->launches += 2;
    }
    ctx->prof_end("expand");
    VB_CUDA(cudaGetLastError());
    CUtensorMap mq, mt;
    if ((rc = make_map(&mq, Eq, (uint64_t)P * n1))) return rc;
    if ((rc = make_map(&mt, Et, (uint64_t)P * n2))) return rc;
    dim3 grid(div_up(n1, TC_QROWS), P);
    ctx->prof_begin("hamming");
    k_knn2_tc<<<grid, TC_THREADS, TC_SMEM_BYTES, ctx->stream>>>(mq, mt, n1, n2, n1, n2, ctx->ws[WS_KNN_PART].as<uint2>());
    ctx->prof_end("hamming");
    ctx->launches++;
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

}  // namespace vb
