// l2f_tc.cu — float-descriptor kNN-2 (BASELINE config 3) with the ab term of ||a||^2 + ||b||^2 - 2ab on the
// 5th-generation tensor cores (tcgen05 kind::tf32, TMA-fed, TMEM accumulators), and exact results.
//
// The contract is l2f.cu's: d2 = sum_k (a_k - b_k)^2 accumulated sequentially in fp32, two nearest train rows per
// query, equal d2 to the lower train index. A TF32 GEMM cannot deliver that bit pattern, so it is used for what it
// is good at — discarding almost every candidate — and the survivors are re-evaluated exactly:
//
//   k_rownorm2     ||b_j||^2 per train row (fp32), padded with +inf to a whole number of tiles; max ||b||^2.
//   k_l2_tc<DIM>   CTA = 256 queries (A resident in shared memory, two M = 128 row halves) against the whole train
//                  set in 64-row tiles:
//                    warp 0     TMA producer (A once, B tiles through a 2-stage ring, fp32 rows, 128B swizzle)
//                    warp 1     one thread issues the UMMAs (DIM/8 per row half and tile) into a ring of
//                               4 TMEM accumulator slots per half
//                    warp 2     TMEM allocation
//                    warps 4-11 drain: key = ||b||^2 - 2 ab (= d2 - ||a||^2, one FFMA per value), minimum per
//                               32-column chunk (3-input min tree), one float per (query, chunk) to HBM.
//                  No top-2 bookkeeping, no branches: the only product is the chunk-minimum matrix.
//   k_l2_rerank    one warp per query: m2 = second smallest chunk minimum (so at least two candidates have an
//                  approximate key <= m2); every chunk whose minimum is <= m2 + 2E can hold a true top-2 member,
//                  where E bounds |approximate key - exact key| (TF32 operand rounding 2^-10 each, fp32
//                  accumulation, the fp32 rounding of the exact sequential sum itself). Those chunks (2-3 per
//                  query in practice) are staged in shared memory and their 32 distances evaluated with exactly
//                  l2f.cu's arithmetic; top-2 by (d2 bits, index). Degenerate inputs (all rows equal) degrade to
//                  an exact brute-force scan, never to a wrong answer.
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_host.cuh"

namespace vb {

using namespace tc;

constexpr int LT_THREADS = 128 + 256;   // 4 service warps + 8 draining warps
constexpr int LT_QROWS = 256;           // queries per CTA
constexpr int LT_NCOLS = 64;            // train rows per tile (UMMA N)
constexpr int LT_SLOTS = 4;             // accumulator slots per row half: 2 halves x 4 x 64 = 512 TMEM columns
constexpr uint32_t LT_A_BOX = 128 * 128;        // 128 rows x 128 B (32 floats)
constexpr uint32_t LT_B_BOX = LT_NCOLS * 128;   // 64 rows x 128 B

template <int DIM> struct LtCfg {
    static constexpr int KATOMS = DIM / 32;
    static constexpr uint32_t A_BYTES = 2 * KATOMS * LT_A_BOX;
    static constexpr uint32_t B_BYTES = KATOMS * LT_B_BOX;
    static constexpr uint32_t SMEM = A_BYTES + 2 * B_BYTES + 1024 /*alignment*/ + 256 /*barriers*/ + 8 * 64 * 4 /*norm slots*/;
};

// One warp per row: out[r] = sum_k d[r][k]^2 (fp32, lane-strided partials + butterfly), rows [n, n_pad) = +inf.
__global__ void __launch_bounds__(256) k_rownorm2(const float *__restrict__ d, uint32_t n, uint32_t n_pad, uint32_t dim,
                                                  float *__restrict__ out, uint32_t *__restrict__ max_bits) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_pad) return;
    if (r >= n) {
        if (lane == 0) out[r] = INFINITY;
        return;
    }
    float s = 0.f;
    for (uint32_t k = lane; k < dim; k += 32) {
        const float v = __ldg(d + (size_t)r * dim + k);
        s = fmaf(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        out[r] = s;
        if (max_bits) atomicMax(max_bits, __float_as_uint(s));   // s >= 0: float order == unsigned order of the bits
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
        const float *__restrict__ nb, float *__restrict__ chunkmin, uint32_t nct) {
    using Cfg = LtCfg<DIM>;
    constexpr int KATOMS = Cfg::KATOMS;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = smem0;
    const uint32_t sB = smem0 + Cfg::A_BYTES;
    const uint32_t sBar = sB + 2 * Cfg::B_BYTES;
    const uint32_t bar_a = sBar, bar_full = sBar + 8, bar_empty = sBar + 24, bar_tfull = sBar + 40,
                   bar_tempty = sBar + 40 + 8 * 2 * LT_SLOTS;
    const uint32_t s_tmem = bar_tempty + 8 * 2 * LT_SLOTS;
    float *s_nb = reinterpret_cast<float *>(smem_raw + (sBar + 256 - smem_u32(smem_raw)));   // [8 warps][64]

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q0 = blockIdx.x * LT_QROWS;
    const uint32_t ntiles = (n2 + LT_NCOLS - 1) / LT_NCOLS;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_t);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(bar_a, 1);
        for (int s = 0; s < 2; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int i = 0; i < 2 * LT_SLOTS; i++) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 4);   // one arrival per draining warp of that row half
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
            mbar_expect_tx(bar_a, Cfg::A_BYTES);
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int ka = 0; ka < KATOMS; ka++)
                    tma_load_2d(sA + (h * KATOMS + ka) * LT_A_BOX, &map_q, ka * 32, (int32_t)(q0 + h * 128), bar_a);
            for (uint32_t j = 0; j < ntiles; j++) {
                const uint32_t s = j & 1, ph = (j >> 1) & 1;
                mbar_wait(bar_empty + 8 * s, ph ^ 1);
                mbar_expect_tx(bar_full + 8 * s, Cfg::B_BYTES);
#pragma unroll
                for (int ka = 0; ka < KATOMS; ka++)
                    tma_load_2d(sB + s * Cfg::B_BYTES + ka * LT_B_BOX, &map_t, ka * 32, (int32_t)(j * LT_NCOLS), bar_full + 8 * s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(UMMA_FMT_TF32, 128, LT_NCOLS);
            mbar_wait(bar_a, 0);
            for (uint32_t j = 0; j < ntiles; j++) {
                const uint32_t s = j & 1, ph = (j >> 1) & 1;
                const uint32_t slot = j % LT_SLOTS, sph = (j / LT_SLOTS) & 1;
                mbar_wait(bar_full + 8 * s, ph);
                tc_fence_after();
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    mbar_wait(bar_tempty + 8 * (h * LT_SLOTS + slot), sph ^ 1);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < DIM / 8; k++) {
                        const uint32_t atom = k >> 2, koff = (k & 3) * 32;
                        const uint64_t ad = smem_desc_sw128(sA + (h * KATOMS + atom) * LT_A_BOX + koff);
                        const uint64_t bd = smem_desc_sw128(sB + s * Cfg::B_BYTES + atom * LT_B_BOX + koff);
                        umma_tf32(tmem_base + h * (LT_SLOTS * LT_NCOLS) + slot * LT_NCOLS, ad, bd, idesc, k != 0 ? 1u : 0u);
                    }
                    umma_commit(bar_tfull + 8 * (h * LT_SLOTS + slot));
                }
                umma_commit(bar_empty + 8 * s);
            }
        }
    } else if (warp >= 4) {
        const uint32_t ew = warp - 4;
        const uint32_t quad = warp & 3, h = ew >> 2;   // TMEM lane quadrant (= warp % 4), row half
        const uint32_t q = q0 + h * 128 + quad * 32 + lane;
        const uint32_t taddr = tmem_base + ((quad * 32u) << 16) + h * (LT_SLOTS * LT_NCOLS);
        float *nbs = s_nb + ew * 64;
        float *out = chunkmin + (size_t)q * nct;
        uint32_t raw0[32], raw1[32];
        for (uint32_t j = 0; j < ntiles; j++) {
            const uint32_t slot = j % LT_SLOTS, sph = (j / LT_SLOTS) & 1;
            // this tile's 64 train norms -> the warp's own shared-memory slot (nb is padded with +inf past n2)
            const float nb0 = __ldg(nb + j * LT_NCOLS + lane), nb1 = __ldg(nb + j * LT_NCOLS + 32 + lane);
            __syncwarp();
            nbs[lane] = nb0;
            nbs[32 + lane] = nb1;
            __syncwarp();
            mbar_wait(bar_tfull + 8 * (h * LT_SLOTS + slot), sph);
            tc_fence_after();
            tmem_ld32(taddr + slot * LT_NCOLS, raw0);
            tmem_ld32(taddr + slot * LT_NCOLS + 32, raw1);
            tmem_wait_ld_regs(raw0);
            tmem_wait_ld_regs(raw1);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * (h * LT_SLOTS + slot));
            const float m0 = chunk_min_l2(raw0, nbs);
            const float m1 = chunk_min_l2(raw1, nbs + 32);
            if (q < n1) *reinterpret_cast<float2 *>(out + 2 * j) = make_float2(m0, m1);
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
constexpr int RR_WARPS = 4;
template <int DIM> struct RrCfg {
    static constexpr int ROWF = DIM + 4;   // padded row (floats): conflict-free LDS.128 with one row per lane
    static constexpr uint32_t WARP_FLOATS = DIM + 32 * ROWF;
    static constexpr uint32_t SMEM = RR_WARPS * WARP_FLOATS * 4;
};

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

template <int DIM>
__global__ void __launch_bounds__(RR_WARPS * 32) k_l2_rerank(const float *__restrict__ d1, const float *__restrict__ d2,
                                                             uint32_t n1, uint32_t n2, const float *__restrict__ chunkmin,
                                                             uint32_t nct, const uint32_t *__restrict__ bmax_bits,
                                                             ulonglong2 *__restrict__ out) {
    using Cfg = RrCfg<DIM>;
    extern __shared__ float rr_smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x * RR_WARPS + warp;
    if (q >= n1) return;
    float *as = rr_smem + warp * Cfg::WARP_FLOATS;
    float *tile = as + DIM;
    // the query descriptor, and ||a||^2 for the error bound
    float na2 = 0.f;
    for (uint32_t k = lane; k < DIM; k += 32) {
        const float v = __ldg(d1 + (size_t)q * DIM + k);
        as[k] = v;
        na2 = fmaf(v, v, na2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) na2 += __shfl_xor_sync(0xffffffffu, na2, o);
    __syncwarp();
    // m2: second smallest chunk minimum of this query's row
    const float *row = chunkmin + (size_t)q * nct;
    float r0 = INFINITY, r1 = INFINITY;
    for (uint32_t i = lane; i < nct; i += 32) {
        const float v = __ldg(row + i);
        r1 = fminf(r1, fmaxf(r0, v));
        r0 = fminf(r0, v);
    }
    const float m1 = warp_min_f(r0);
    const uint32_t owner = __ffs(__ballot_sync(0xffffffffu, r0 == m1)) - 1;
    const float m2 = warp_min_f(lane == owner ? r1 : r0);
    // E >= |approximate key - exact key|:  2 * (2^-9 + 2^-16) * |a||b|  (TF32 operands, fp32 accumulation)
    //                                      + 2^-16 (|a| + |b|)^2        (norm sums, the exact sum's own rounding)
    // taken with ~1.4x slack and the largest train norm
    const float an = sqrtf(na2) * 1.001f, bn = sqrtf(__uint_as_float(__ldg(bmax_bits))) * 1.001f;
    const float E = 0.0055243f * an * bn + 6.1035e-5f * (an + bn) * (an + bn);
    const float T = m2 + 2.f * E;
    unsigned long long k0 = ~0ull, k1 = ~0ull;
    for (uint32_t base = 0; base < nct; base += 32) {
        const uint32_t i = base + lane;
        uint32_t bal = __ballot_sync(0xffffffffu, i < nct && __ldg(row + i) <= T);
        while (bal) {
            const uint32_t c = base + __ffs(bal) - 1;
            bal &= bal - 1;
            const uint32_t col0 = c * 32;
            __syncwarp();
            // 32 consecutive train rows = one contiguous block: coalesced float4 loads into the padded tile
            for (uint32_t t = lane; t < 32 * (DIM / 4); t += 32) {
                const uint32_t r = t / (DIM / 4), pos = t % (DIM / 4);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (col0 + r < n2) v = __ldg(reinterpret_cast<const float4 *>(d2 + (size_t)(col0 + r) * DIM) + pos);
                *reinterpret_cast<float4 *>(tile + r * Cfg::ROWF + pos * 4) = v;
            }
            __syncwarp();
            const uint32_t col = col0 + lane;
            if (col < n2) {
                const float *b = tile + lane * Cfg::ROWF;
                float acc = 0.0f;
#pragma unroll 8
                for (int k4 = 0; k4 < DIM / 4; k4++) {
                    const float4 bv = *reinterpret_cast<const float4 *>(b + 4 * k4);
                    const float4 av = *reinterpret_cast<const float4 *>(as + 4 * k4);
                    float df;
                    df = __fsub_rn(av.x, bv.x); acc = __fadd_rn(acc, __fmul_rn(df, df));
                    df = __fsub_rn(av.y, bv.y); acc = __fadd_rn(acc, __fmul_rn(df, df));
                    df = __fsub_rn(av.z, bv.z); acc = __fadd_rn(acc, __fmul_rn(df, df));
                    df = __fsub_rn(av.w, bv.w); acc = __fadd_rn(acc, __fmul_rn(df, df));
                }
                const unsigned long long key = ((unsigned long long)__float_as_uint(acc) << 32) | col;
                if (key < k1) {
                    if (key < k0) { k1 = k0; k0 = key; } else { k1 = key; }
                }
            }
        }
    }
    const unsigned long long g0 = warp_min_u64(k0);
    const uint32_t own = __ffs(__ballot_sync(0xffffffffu, k0 == g0)) - 1;
    const unsigned long long g1 = warp_min_u64(lane == own ? k1 : k0);
    if (lane == 0) out[q] = make_ulonglong2(g0, g1);
}

// ---- host --------------------------------------------------------------------------------------------
bool l2_tc_eligible(uint32_t n1, uint32_t n2, uint32_t dim) {
    if (dim != 64 && dim != 128) return false;
    if (const char *e = getenv("VB_L2_TC")) return atoi(e) != 0;
    return (uint64_t)n1 * n2 >= (1ull << 22);
}

template <int DIM>
static int l2_tc_run(vb_ctx *ctx, const float *d1, uint32_t n1, const float *d2, uint32_t n2, ulonglong2 *out) {
    using Cfg = LtCfg<DIM>;
    static bool attr_set = false;
    if (!attr_set) {
        VB_CUDA(cudaFuncSetAttribute(k_l2_tc<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        VB_CUDA(cudaFuncSetAttribute(k_l2_rerank<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RrCfg<DIM>::SMEM));
        attr_set = true;
    }
    int rc;
    const uint32_t ntiles = div_up(n2, LT_NCOLS), n2_pad = ntiles * LT_NCOLS, nct = ntiles * 2;
    if ((rc = ctx->ws_ensure(WS_L2N, (size_t)n2_pad * 4 + 256))) return rc;
    if ((rc = ctx->ws_ensure(WS_L2M, (size_t)n1 * nct * 4))) return rc;
    uint32_t *bmax = ctx->ws[WS_L2N].as<uint32_t>();          // [0] = bits of max ||b||^2, norms start 256 B in
    float *nb = reinterpret_cast<float *>(ctx->ws[WS_L2N].as<uint8_t>() + 256);
    float *chunkmin = ctx->ws[WS_L2M].as<float>();
    VB_CUDA(cudaMemsetAsync(bmax, 0, 4, ctx->stream));
    ctx->prof_begin("l2f");
    k_rownorm2<<<div_up(n2_pad, 8), 256, 0, ctx->stream>>>(d2, n2, n2_pad, DIM, nb, bmax);
    CUtensorMap mq, mt;
    if ((rc = make_map_2d(&mq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d1, DIM, n1, 32, 128))) return rc;
    if ((rc = make_map_2d(&mt, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d2, DIM, n2, 32, LT_NCOLS))) return rc;
    ctx->prof_begin("l2f_gemm");
    k_l2_tc<DIM><<<div_up(n1, LT_QROWS), LT_THREADS, Cfg::SMEM, ctx->stream>>>(mq, mt, n1, n2, nb, chunkmin, nct);
    ctx->prof_end("l2f_gemm");
    ctx->prof_begin("l2f_rerank");
    k_l2_rerank<DIM><<<div_up(n1, RR_WARPS), RR_WARPS * 32, RrCfg<DIM>::SMEM, ctx->stream>>>(d1, d2, n1, n2, chunkmin, nct, bmax, out);
    ctx->prof_end("l2f_rerank");
    ctx->prof_end("l2f");
    ctx->launches += 3;
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

// out = [n1] (best, second) keys (d2 bits << 32 | train index), the layout k_l2_finish reads with nsplits = 1.
int l2_tc_launch(vb_ctx *ctx, const float *d1, uint32_t n1, const float *d2, uint32_t n2, uint32_t dim, ulonglong2 *out) {
    return dim == 128 ? l2_tc_run<128>(ctx, d1, n1, d2, n2, out) : l2_tc_run<64>(ctx, d1, n1, d2, n2, out);
}

}  // namespace vb
