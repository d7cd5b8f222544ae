// l2f_tc.cu — float-descriptor kNN-2 (BASELINE config 3) with the ab term of ||a||^2 + ||b||^2 - 2ab on the
// 5th-generation tensor cores (tcgen05 kind::f16 on bf16 operands, TMA-fed, TMEM accumulators), and exact results.
//
// The contract is l2f.cu's: d2 = sum_k (a_k - b_k)^2 accumulated sequentially in fp32, two nearest train rows per
// query, equal d2 to the lower train index. A reduced-precision GEMM cannot deliver that bit pattern, so it is used
// for what it is good at — discarding almost every candidate — and the survivors are re-evaluated exactly.
// Why bf16 and not tf32: with fp32 operands a 128 x N x 8 UMMA reads 4 + N/16 KB of shared memory per N/2 cycles,
// which exceeds the 128 B/clk shared-memory port for N < 256, and the B stream is 4 bytes per element from L2
// (measured: 0.26 ms, smem/L2-bound). bf16 halves both and doubles the MMA rate; what it costs is a wider
// candidate window, which only the cheap exact pass sees.
//
//   k_l2_prep      one warp per row: fp32 -> bf16 (round to nearest even); for train rows also ||b||^2 (fp32, padded
//                  with +inf to whole tiles) and the largest ||b||^2.
//   k_l2_tc<DIM>   CTA = 256 queries (A resident in shared memory, two M = 128 row halves) against a range of
//                  256-row train tiles — the same pipeline geometry as k_knn2_tc (hamming_tc.cu):
//                    warp 0     TMA producer (A once, B tiles through a 2-stage ring, 128B swizzle)
//                    warp 1     one thread issues the UMMAs: DIM/16 per row half and tile, N = 256
//                    warp 2     TMEM allocation (two 256-column accumulators)
//                    warps 4-19 drain, one row and 128 columns per thread and tile: key = ||b||^2 - 2 ab
//                               (= d2 - ||a||^2, one FFMA per value), minimum per 32-column chunk (3-input min
//                               tree), one float per (query, chunk) to HBM.
//                  No top-2 bookkeeping, no branches: the only product is the chunk-minimum matrix.
//   k_l2_rerank    one warp per query: m2 = second smallest chunk minimum (so at least two candidates have an
//                  approximate key <= m2); every chunk whose minimum is <= m2 + 2E can hold a true top-2 member,
//                  where E bounds |approximate key - exact key| (bf16 operand rounding 2^-9 each, fp32
//                  accumulation, the fp32 rounding of the exact sequential sum itself). Those chunks (2-3 per
//                  query in practice) are scanned one train row per lane: first the bf16 dot product again
//                  (256 B per row), and only columns whose own approximate key is <= m2 + 2E (about one per
//                  chunk) have their fp32 row read — coalesced, by the whole warp — and summed in l2f.cu's exact
//                  order; top-2 by (d2 bits, index). Degenerate inputs (all rows equal) degrade to
//                  an exact brute-force scan, never to a wrong answer.
#include <cuda_bf16.h>

#include <cmath>

#include "common.cuh"
#include "tc_common.cuh"
#include "tc_host.cuh"

namespace vb {

using namespace tc;

constexpr int LT_COLSPLIT = 2;                          // threads draining one accumulator row (column parts)
constexpr int LT_THREADS = 128 + 256 * LT_COLSPLIT;     // 4 service warps + 16 draining warps
constexpr int LT_QROWS = 256;                           // queries per CTA
constexpr int LT_NCOLS = 256;                           // train rows per tile (UMMA N)
constexpr int LT_CHUNKS = LT_NCOLS / 32;                // chunk minima per query and tile
constexpr uint32_t LT_BOX = 128 * 128;                  // TMA box: 128 rows x 128 B (64 bf16)

template <int DIM> struct LtCfg {
    static constexpr int KATOMS = DIM * 2 / 128;        // 128-byte swizzle atoms per row
    static constexpr uint32_t A_BYTES = 2 * KATOMS * LT_BOX;    // [half][katom][128 rows][128 B]
    static constexpr uint32_t B_BYTES = 2 * KATOMS * LT_BOX;    // [katom][256 rows][128 B]
    static constexpr uint32_t SMEM = A_BYTES + 2 * B_BYTES + 1024 /*alignment*/ + 256 /*barriers*/ + 16 * 128 * 4 /*norm slots*/;
};

// One warp per row: dst[r] = bf16(src[r]) (RN); optionally norms[r] = sum_k src[r][k]^2 (fp32, lane-strided partials
// + butterfly), rows [n, n_pad) of norms = +inf, *max_bits = bits of the largest norm.
__global__ void __launch_bounds__(256) k_l2_prep(const float *__restrict__ src, uint32_t n, uint32_t n_pad, uint32_t dim,
                                                 __nv_bfloat16 *__restrict__ dst, float *__restrict__ norms,
                                                 uint32_t *__restrict__ max_bits) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_pad) return;
    if (r >= n) {
        if (norms && lane == 0) norms[r] = INFINITY;
        return;
    }
    float s = 0.f;
    for (uint32_t k = 2 * lane; k < dim; k += 64) {
        const float2 v = __ldg(reinterpret_cast<const float2 *>(src + (size_t)r * dim + k));
        s = fmaf(v.x, v.x, s);
        s = fmaf(v.y, v.y, s);
        *reinterpret_cast<__nv_bfloat162 *>(dst + (size_t)r * dim + k) = __floats2bfloat162_rn(v.x, v.y);
    }
    if (!norms) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        norms[r] = s;
        atomicMax(max_bits, __float_as_uint(s));   // s >= 0: float order == unsigned order of the bits
    }
}

// min over 32 keys  nb[i] - 2 * acc[i]; nb read from shared memory as broadcast LDS.128
__device__ __forceinline__ float chunk_min_l2(const uint32_t (&raw)[32], const float *nbs) {
    float k[32];
#pragma unroll
    for (int g = 0; g < 8; g++) {
        const float4 nb = *reinterpret_cast<const float4 *>(nbs + 4 * g);
        k[4 * g + 0] = fmaf(__uint_as_float(raw[4 * g + 0]), -2.f, nb.x);
        k[4 * g + 1] = fmaf(__uint_as_float(raw[4 * g + 1]), -2.f, nb.y);
        k[4 * g + 2] = fmaf(__uint_as_float(raw[4 * g + 2]), -2.f, nb.z);
        k[4 * g + 3] = fmaf(__uint_as_float(raw[4 * g + 3]), -2.f, nb.w);
    }
    float a[12];
#pragma unroll
    for (int i = 0; i < 10; i++) a[i] = fminf(fminf(k[3 * i], k[3 * i + 1]), k[3 * i + 2]);
    a[10] = k[30];
    a[11] = k[31];
    const float b0 = fminf(fminf(a[0], a[1]), a[2]), b1 = fminf(fminf(a[3], a[4]), a[5]), b2 = fminf(fminf(a[6], a[7]), a[8]),
                b3 = fminf(fminf(a[9], a[10]), a[11]);
    return fminf(fminf(fminf(b0, b1), b2), b3);
}

template <int DIM>
__global__ void __launch_bounds__(LT_THREADS, 1)
k_l2_tc(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_t, uint32_t n1, uint32_t n2,
        uint32_t tiles_per_split, const float *__restrict__ nb, float *__restrict__ chunkmin, uint32_t nct) {
    using Cfg = LtCfg<DIM>;
    constexpr int KATOMS = Cfg::KATOMS;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = smem0;
    const uint32_t sB = smem0 + Cfg::A_BYTES;
    const uint32_t sBar = sB + 2 * Cfg::B_BYTES;
    const uint32_t bar_a = sBar, bar_full = sBar + 8, bar_empty = sBar + 24, bar_tfull = sBar + 40, bar_tempty = sBar + 56;
    const uint32_t s_tmem = sBar + 72;
    float *s_nb = reinterpret_cast<float *>(smem_raw + (sBar + 256 - smem_u32(smem_raw)));   // [16 warps][128]

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q0 = blockIdx.x * LT_QROWS;
    const uint32_t ntiles_all = (n2 + LT_NCOLS - 1) / LT_NCOLS;
    const uint32_t t0 = blockIdx.y * tiles_per_split;
    const uint32_t ntiles = min(tiles_per_split, ntiles_all - t0);   // the host never launches an empty split

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
            mbar_init(bar_tempty + 8 * s, 4 * LT_COLSPLIT);   // one arrival per draining warp of that row half
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(s_tmem));

    // Service warps walk their loops whole (every lane waits on the barriers); one elected lane issues, so that ptxas emits
    // UTMALDG / UTCHMMA / UTCBAR straight-line instead of an ELECT / BRA.U.ANY loop around each (tc_common.cuh, elect_one).
    if (warp == 0) {
        const bool leader = elect_one();
        if (leader) {
            mbar_expect_tx(bar_a, Cfg::A_BYTES);
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int ka = 0; ka < KATOMS; ka++)
                    tma_load_2d(sA + (h * KATOMS + ka) * LT_BOX, &map_q, ka * 64, (int32_t)(q0 + h * 128), bar_a);
        }
        for (uint32_t j = 0; j < ntiles; j++) {
            const uint32_t s = j & 1, ph = (j >> 1) & 1;
            mbar_wait(bar_empty + 8 * s, ph ^ 1);
            if (leader) {
                mbar_expect_tx(bar_full + 8 * s, Cfg::B_BYTES);
                const uint32_t dst = sB + s * Cfg::B_BYTES;
#pragma unroll
                for (int ka = 0; ka < KATOMS; ka++)
#pragma unroll
                    for (int r = 0; r < 2; r++)
                        tma_load_2d(dst + (ka * 2 + r) * LT_BOX, &map_t, ka * 64, (int32_t)((t0 + j) * LT_NCOLS) + r * 128,
                                    bar_full + 8 * s);
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one();
        constexpr uint32_t idesc = umma_idesc(UMMA_FMT_BF16, 128, LT_NCOLS);
        mbar_wait(bar_a, 0);
        for (uint32_t j = 0; j < ntiles; j++) {
            const uint32_t s = j & 1, ph = (j >> 1) & 1;
            mbar_wait(bar_full + 8 * s, ph);
            tc_fence_after();
            const uint32_t bbase = sB + s * Cfg::B_BYTES;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                mbar_wait(bar_tempty + 8 * h, (j & 1) ^ 1);   // accumulator h drained (tile j-1)
                tc_fence_after();
                if (leader) {
#pragma unroll
                    for (int ka = 0; ka < KATOMS; ka++)
#pragma unroll
                        for (int k = 0; k < 4; k++) {   // 4 K-steps of 16 bf16 (32 B) per 128-byte atom
                            const uint64_t ad = smem_desc_sw128(sA + (h * KATOMS + ka) * LT_BOX + k * 32);
                            const uint64_t bd = smem_desc_sw128(bbase + ka * 2 * LT_BOX + k * 32);
                            umma_f16(tmem_base + h * LT_NCOLS, ad, bd, idesc, (ka | k) != 0 ? 1u : 0u);
                        }
                    umma_commit(bar_tfull + 8 * h);
                }
            }
            if (leader) umma_commit(bar_empty + 8 * s);   // both halves have consumed this B stage
        }
    } else if (warp >= 4) {
        const uint32_t ew = warp - 4;
        const uint32_t quad = warp & 3, h = (ew >> 2) & 1, ch = ew >> 3;   // TMEM lane quadrant, accumulator, column part
        const uint32_t q = q0 + h * 128 + quad * 32 + lane;
        constexpr uint32_t CW = LT_NCOLS / LT_COLSPLIT;
        static_assert(CW == 128, "the drain below is written out for four 32-column chunks per warp");
        const uint32_t taddr = tmem_base + ((quad * 32u) << 16) + h * LT_NCOLS + ch * CW;
        float *nbs = s_nb + ew * CW;
        float *out = chunkmin + (size_t)q * nct + ch * (CW / 32);
        const float *nbp = nb + (size_t)t0 * LT_NCOLS + ch * CW + lane;
        uint32_t raw0[32], raw1[32];
        // the 128 train norms of this warp's columns, fetched one tile ahead (nb is padded with +inf past n2, which
        // also masks the columns of the last tile)
        float pre[4];
#pragma unroll
        for (int i = 0; i < 4; i++) pre[i] = __ldg(nbp + 32 * i);
        for (uint32_t j = 0; j < ntiles; j++) {
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; i++) nbs[32 * i + lane] = pre[i];
            __syncwarp();
            if (j + 1 < ntiles) {
#pragma unroll
                for (int i = 0; i < 4; i++) pre[i] = __ldg(nbp + (size_t)(j + 1) * LT_NCOLS + 32 * i);
            }
            mbar_wait(bar_tfull + 8 * h, j & 1);
            tc_fence_after();
            float m[4];
            tmem_ld32(taddr, raw0);
            tmem_wait_ld_regs(raw0);
            tmem_ld32(taddr + 32, raw1);
            m[0] = chunk_min_l2(raw0, nbs);
            tmem_wait_ld_regs(raw1);
            tmem_ld32(taddr + 64, raw0);
            m[1] = chunk_min_l2(raw1, nbs + 32);
            tmem_wait_ld_regs(raw0);
            tmem_ld32(taddr + 96, raw1);
            m[2] = chunk_min_l2(raw0, nbs + 64);
            tmem_wait_ld_regs(raw1);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * h);
            m[3] = chunk_min_l2(raw1, nbs + 96);
            if (q < n1) *reinterpret_cast<float4 *>(out + (size_t)(t0 + j) * LT_CHUNKS) = make_float4(m[0], m[1], m[2], m[3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---- exact re-evaluation -----------------------------------------------------------------------------
constexpr int RR_WARPS = 8;

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w < v ? w : v;
    }
    return v;
}
__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

constexpr int RR_BATCH = 8;   // surviving columns evaluated exactly per round (shared-memory rows per warp)

template <int DIM>
__global__ void __launch_bounds__(RR_WARPS * 32) k_l2_rerank(const float *__restrict__ d1, const float *__restrict__ d2,
                                                             const __nv_bfloat16 *__restrict__ h2, const float *__restrict__ nb,
                                                             uint32_t n1, uint32_t n2, const float *__restrict__ chunkmin,
                                                             uint32_t nct, const uint32_t *__restrict__ bmax_bits,
                                                             ulonglong2 *__restrict__ out) {
    constexpr int TROW = DIM + 4;   // padded term row: lanes summing different rows hit different banks
    constexpr int WARP_FLOATS = 2 * DIM + RR_BATCH * TROW;
    __shared__ __align__(16) float rr_smem[RR_WARPS * WARP_FLOATS];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x * RR_WARPS + warp;
    if (q >= n1) return;
    float *as = rr_smem + warp * WARP_FLOATS;   // the query descriptor ...
    float *ah = as + DIM;                       // ... its bf16 rounding (what the GEMM saw), as floats ...
    float *terms = ah + DIM;                    // ... and (a_k - b_k)^2 of up to RR_BATCH surviving columns
    float na2 = 0.f;
    for (uint32_t k = lane; k < DIM; k += 32) {
        const float v = __ldg(d1 + (size_t)q * DIM + k);
        as[k] = v;
        ah[k] = __bfloat162float(__float2bfloat16_rn(v));
        na2 = fmaf(v, v, na2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) na2 += __shfl_xor_sync(0xffffffffu, na2, o);
    __syncwarp();
    // m2: second smallest chunk minimum of this query's row
    const float *row = chunkmin + (size_t)q * nct;
    float r0 = INFINITY, r1 = INFINITY;
#pragma unroll 8
    for (uint32_t i = lane; i < nct; i += 32) {
        const float v = __ldg(row + i);
        r1 = fminf(r1, fmaxf(r0, v));
        r0 = fminf(r0, v);
    }
    const float m1 = warp_min_f(r0);
    const uint32_t owner = __ffs(__ballot_sync(0xffffffffu, r0 == m1)) - 1;
    const float m2 = warp_min_f(lane == owner ? r1 : r0);
    // E >= |approximate key - exact key|:  2 * (2^-8 + 2^-16) * |a||b|  (bf16 operands rounded to nearest: 2^-9 each;
    //                                                                    fp32 accumulation of the products, any order)
    //                                      + 2^-16 (|a| + |b|)^2        (norm sums, the exact sum's own rounding)
    // taken with ~1.4x slack and the largest train norm. A member c of the exact top-2 has exact key <= m2 + E,
    // hence any approximate key of it (the GEMM's, or the bf16 dot product recomputed below) is <= T = m2 + 2E.
    const float an = sqrtf(na2) * 1.001f, bn = sqrtf(__uint_as_float(__ldg(bmax_bits))) * 1.001f;
    const float E = 0.011f * an * bn + 6.1035e-5f * (an + bn) * (an + bn);
    const float T = m2 + 2.f * E;
    unsigned long long k0 = ~0ull, k1 = ~0ull;
    for (uint32_t base = 0; base < nct; base += 32) {
        const uint32_t i = base + lane;
        uint32_t bal = __ballot_sync(0xffffffffu, i < nct && __ldg(row + i) <= T);
        while (bal) {
            const uint32_t c = base + __ffs(bal) - 1;
            bal &= bal - 1;
            // Stage 1, one train row per lane: the approximation the GEMM made (bf16 row, 256 B, all loads in flight
            // at once). Most of the chunk's 32 columns fail key <= T and never touch their fp32 row.
            const uint32_t col = c * 32 + lane;
            bool pass = false;
            if (col < n2) {
                const uint4 *hb = reinterpret_cast<const uint4 *>(h2 + (size_t)col * DIM);
                uint4 w[DIM / 8];
#pragma unroll
                for (int k8 = 0; k8 < DIM / 8; k8++) w[k8] = __ldg(hb + k8);
                const float nbc = __ldg(nb + col);
                float dot0 = 0.f, dot1 = 0.f;
#pragma unroll
                for (int k8 = 0; k8 < DIM / 8; k8++) {
                    const float4 a0 = *reinterpret_cast<const float4 *>(ah + 8 * k8);
                    const float4 a1 = *reinterpret_cast<const float4 *>(ah + 8 * k8 + 4);
                    dot0 = fmaf(a0.x, __uint_as_float(w[k8].x << 16), dot0);
                    dot1 = fmaf(a0.y, __uint_as_float(w[k8].x & 0xffff0000u), dot1);
                    dot0 = fmaf(a0.z, __uint_as_float(w[k8].y << 16), dot0);
                    dot1 = fmaf(a0.w, __uint_as_float(w[k8].y & 0xffff0000u), dot1);
                    dot0 = fmaf(a1.x, __uint_as_float(w[k8].z << 16), dot0);
                    dot1 = fmaf(a1.y, __uint_as_float(w[k8].z & 0xffff0000u), dot1);
                    dot0 = fmaf(a1.z, __uint_as_float(w[k8].w << 16), dot0);
                    dot1 = fmaf(a1.w, __uint_as_float(w[k8].w & 0xffff0000u), dot1);
                }
                pass = fmaf(dot0 + dot1, -2.f, nbc) <= T;
            }
            // Stage 2, the survivors, RR_BATCH at a time: the warp reads a survivor's fp32 row as one coalesced
            // request and forms the 128 squared differences in parallel; the sum itself has to run in index order
            // (that is the contract), so lane s then adds up row s — RR_BATCH sequential sums side by side.
            uint32_t sb = __ballot_sync(0xffffffffu, pass);
            while (sb) {
                uint32_t mycol = 0xffffffffu;
                int cnt = 0;
#pragma unroll 1
                for (; cnt < RR_BATCH && sb; cnt++) {
                    const uint32_t sc = c * 32 + __ffs(sb) - 1;
                    sb &= sb - 1;
                    if (lane == (uint32_t)cnt) mycol = sc;
                    if (lane < DIM / 4) {
                        const float4 bv = __ldg(reinterpret_cast<const float4 *>(d2 + (size_t)sc * DIM) + lane);
                        const float4 av = *reinterpret_cast<const float4 *>(as + 4 * lane);
                        float4 t;
                        float df;
                        df = __fsub_rn(av.x, bv.x); t.x = __fmul_rn(df, df);
                        df = __fsub_rn(av.y, bv.y); t.y = __fmul_rn(df, df);
                        df = __fsub_rn(av.z, bv.z); t.z = __fmul_rn(df, df);
                        df = __fsub_rn(av.w, bv.w); t.w = __fmul_rn(df, df);
                        *reinterpret_cast<float4 *>(terms + cnt * TROW + 4 * lane) = t;
                    }
                }
                __syncwarp();
                if (lane < (uint32_t)cnt) {
                    const float *tr = terms + lane * TROW;
                    float acc = 0.0f;
#pragma unroll 8
                    for (int k4 = 0; k4 < DIM / 4; k4++) {
                        const float4 t = *reinterpret_cast<const float4 *>(tr + 4 * k4);
                        acc = __fadd_rn(acc, t.x);
                        acc = __fadd_rn(acc, t.y);
                        acc = __fadd_rn(acc, t.z);
                        acc = __fadd_rn(acc, t.w);
                    }
                    const unsigned long long key = ((unsigned long long)__float_as_uint(acc) << 32) | mycol;
                    if (key < k1) {
                        if (key < k0) { k1 = k0; k0 = key; } else { k1 = key; }
                    }
                }
                __syncwarp();
            }
        }
    }
    const unsigned long long g0 = warp_min_u64(k0);
    const uint32_t own = __ffs(__ballot_sync(0xffffffffu, k0 == g0)) - 1;
    const unsigned long long g1 = warp_min_u64(lane == own ? k1 : k0);
    if (lane == 0) out[q] = make_ulonglong2(g0, g1);
}

// ---- host --------------------------------------------------------------------------------------------
bool l2_tc_eligible(const vb_ctx *ctx, uint32_t n1, uint32_t n2, uint32_t dim) {
    if (dim != 64 && dim != 128) return false;
    const long long force = ctx->opt("l2_tc", -1);
    if (force >= 0) return force != 0;
    return (uint64_t)n1 * n2 >= (1ull << 22);
}

template <int DIM>
static int l2_tc_run(vb_ctx *ctx, const float *d1, uint32_t n1, const float *d2, uint32_t n2, ulonglong2 *out) {
    using Cfg = LtCfg<DIM>;
    constexpr uint32_t bit = DIM == 128 ? 2u : 4u;
    if (!(ctx->func_attr_done & bit)) {   // a function attribute is per device: remembered per context, not per process
        VB_CUDA(cudaFuncSetAttribute(k_l2_tc<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        ctx->func_attr_done |= bit;
    }
    int rc;
    const uint32_t ntiles = div_up(n2, LT_NCOLS), n2_pad = ntiles * LT_NCOLS, nct = ntiles * LT_CHUNKS;
    if ((rc = ctx->ws_ensure(WS_L2N, (size_t)n2_pad * 4 + 256))) return rc;
    if ((rc = ctx->ws_ensure(WS_L2M, (size_t)n1 * nct * 4))) return rc;
    if ((rc = ctx->ws_ensure(WS_L2H, ((size_t)n1 + n2) * DIM * 2))) return rc;
    uint32_t *bmax = ctx->ws[WS_L2N].as<uint32_t>();          // [0] = bits of max ||b||^2, norms start 256 B in
    float *nb = reinterpret_cast<float *>(ctx->ws[WS_L2N].as<uint8_t>() + 256);
    float *chunkmin = ctx->ws[WS_L2M].as<float>();
    __nv_bfloat16 *hq = ctx->ws[WS_L2H].as<__nv_bfloat16>(), *ht = hq + (size_t)n1 * DIM;
    VB_CUDA(cudaMemsetAsync(bmax, 0, 4, ctx->stream));
    ctx->prof_begin("l2f");
    k_l2_prep<<<div_up(n1, 8), 256, 0, ctx->stream>>>(d1, n1, n1, DIM, hq, nullptr, nullptr);
    k_l2_prep<<<div_up(n2_pad, 8), 256, 0, ctx->stream>>>(d2, n2, n2_pad, DIM, ht, nb, bmax);
    CUtensorMap mq, mt;
    if ((rc = make_map_2d(&mq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, hq, DIM, n1, 64, 128))) return rc;
    if ((rc = make_map_2d(&mt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ht, DIM, n2, 64, 128))) return rc;
    // split the train tiles so that the CTA count fills whole waves of the SMs (>= 4 tiles per CTA keeps the
    // resident-A load amortised); the smallest split count within 4 % of the best wave efficiency wins
    const uint32_t rb = div_up(n1, LT_QROWS);
    uint32_t best_s = 1;
    double best_eff = 0.0;
    for (uint32_t s = 1; s <= 64 && (s == 1 || ntiles / s >= 4); s++) {
        const uint32_t tps = div_up(ntiles, s), used = div_up(ntiles, tps);
        const double waves = (double)rb * used / ctx->sm_count;
        const double eff = waves / ceil(waves);
        if (eff > best_eff + 0.04) { best_eff = eff; best_s = s; }
    }
    const uint32_t tiles_per_split = div_up(ntiles, best_s), nsplit = div_up(ntiles, tiles_per_split);
    ctx->prof_begin("l2f_gemm");
    k_l2_tc<DIM><<<dim3(rb, nsplit), LT_THREADS, Cfg::SMEM, ctx->stream>>>(mq, mt, n1, n2, tiles_per_split, nb, chunkmin, nct);
    ctx->prof_end("l2f_gemm");
    ctx->prof_begin("l2f_rerank");
    k_l2_rerank<DIM><<<div_up(n1, RR_WARPS), RR_WARPS * 32, 0, ctx->stream>>>(d1, d2, ht, nb, n1, n2, chunkmin, nct, bmax, out);
    ctx->prof_end("l2f_rerank");
    ctx->prof_end("l2f");
    ctx->launches += 4;
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

// out = [n1] (best, second) keys (d2 bits << 32 | train index), the layout k_l2_finish reads with nsplits = 1.
int l2_tc_launch(vb_ctx *ctx, const float *d1, uint32_t n1, const float *d2, uint32_t n2, uint32_t dim, ulonglong2 *out) {
    return dim == 128 ? l2_tc_run<128>(ctx, d1, n1, d2, n2, out) : l2_tc_run<64>(ctx, d1, n1, d2, n2, out);
}

}  // namespace vb
