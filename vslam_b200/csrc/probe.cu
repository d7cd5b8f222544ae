// probe.cu — measurement hooks for bench.py: what the tensor pipe of THIS GPU delivers when nothing but UMMAs runs.
//
// MEASURED_PEAKS.json carries a dense bf16 figure (cuBLAS, power-throttled run) but nothing for the block-scaled fp4
// instructions the Hamming matcher issues, so the roofline denominator of k_knn2_tc4 used to be derived (4 x bf16).
// k_probe_umma issues the matcher's own instruction — tcgen05.mma kind::mxf4.block_scale, M = 128, K = 64 e2m1 — back
// to back from one thread per SM, operands resident in shared memory (no TMA traffic), accumulating into TMEM, with no
// epilogue: the rate of that loop is the ceiling any kernel built on this instruction shape can approach on this part.
// The same loop with kind::f8f6f4 (e4m3) and kind::f16 (bf16) cross-checks the method against the cuBLAS bf16 figure.
// Nothing here is on the product path.
#include "common.cuh"
#include "tc_common.cuh"

namespace vb {

using namespace tc;

constexpr uint32_t PROBE_SMEM = 16384 + 32768 + 1024 + 64;   // A [128][128 B], B [256][128 B], alignment slack, barrier

// kind: 0 = mxf4 (K = 64), 1 = f8f6f4 e4m3 (K = 32), 2 = f16 bf16 (K = 16). N = UMMA N (multiple of 16, <= 256).
__global__ void __launch_bounds__(128, 1) k_probe_umma(int kind, uint32_t N, uint32_t iters, uint32_t *sink) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = smem0, sB = smem0 + 16384, bar = sB + 32768, s_tmem = bar + 8;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // operand bytes: 0x22 / 0x2A / 0xA2 / 0xAA by a hash of the position — valid and finite in all three formats
    // (e2m1 pairs of +-1, small e4m3 values, small bf16 values), and not constant, so the datapath toggles
    for (uint32_t i = threadIdx.x; i < (16384u + 32768u) / 4u; i += blockDim.x) {
        uint32_t h = (i + blockIdx.x * 7919u) * 2654435761u;
        uint32_t w = 0x22222222u;
        w |= (h & 0x00000100u) ? 0x00000080u : 0u;
        w |= (h & 0x00000200u) ? 0x00000008u : 0u;
        w |= (h & 0x00000400u) ? 0x00008000u : 0u;
        w |= (h & 0x00000800u) ? 0x00000800u : 0u;
        w |= (h & 0x00001000u) ? 0x00800000u : 0u;
        w |= (h & 0x00002000u) ? 0x00080000u : 0u;
        w |= (h & 0x00004000u) ? 0x80000000u : 0u;
        w |= (h & 0x00008000u) ? 0x08000000u : 0u;
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(sA + 4u * i), "r"(w) : "memory");
    }
    if (warp == 1 && lane == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(s_tmem, 512);
    fence_proxy_async();   // generic-proxy writes above must be visible to the UMMA unit's async-proxy reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(s_tmem));
    tmem_st32_const(tmem_base + ((warp * 32u) << 16), 0x7f7f7f7fu);   // ue8m0 2^0 scale factors (columns [0, 32))
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1 && elect_one()) {   // warp-uniform condition, then one elected lane: straight-line UTCxMMA (tc_common.cuh)
        const uint32_t acc = tmem_base + 32, sfa = tmem_base, sfb = tmem_base + 16;
        const uint32_t idesc4 = umma_idesc_mxf4(128, N);
        const uint32_t idesc8 = umma_idesc(UMMA_FMT_E4M3, 128, N);
        const uint32_t idesc16 = umma_idesc(UMMA_FMT_BF16, 128, N);
        for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
            for (int k = 0; k < 4; k++) {   // four K-steps of 32 bytes inside the 128-byte swizzled row
                const uint64_t ad = smem_desc_sw128(sA + k * 32), bd = smem_desc_sw128(sB + k * 32);
                const uint32_t accum = (it | (uint32_t)k) ? 1u : 0u;
                if (kind == 0) umma_mxf4(acc, ad, bd, idesc4, sfa, sfb, accum);
                else if (kind == 1) umma_f8f6f4(acc, ad, bd, idesc8, accum);
                else umma_f16(acc, ad, bd, idesc16, accum);
            }
        }
        umma_commit(bar);
        mbar_wait(bar, 0);
        tc_fence_after();
    }
    __syncthreads();
    if (warp == 0) {   // keep the accumulator observable
        uint32_t v[4];
        tmem_ld4(tmem_base + 32, v);
        tmem_wait_ld_regs4(v);
        if (lane == 0 && sink) sink[blockIdx.x] = v[0] ^ v[1] ^ v[2] ^ v[3];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace vb

using namespace vb;

extern "C" int vb_probe_tensor_peak(vb_ctx *ctx, int kind, uint32_t n_cols, uint32_t iters, uint32_t reps, float *best_ms,
                                    double *flop_per_launch) {
    VB_REQUIRE(ctx && best_ms && flop_per_launch, VB_ERR_INVALID, "NULL argument");
    VB_REQUIRE(kind >= 0 && kind <= 2, VB_ERR_INVALID, "kind must be 0 (mxf4), 1 (f8f6f4) or 2 (f16)");
    VB_REQUIRE(n_cols >= 16 && n_cols <= 256 && n_cols % 16 == 0, VB_ERR_INVALID, "n_cols must be a multiple of 16 in [16, 256]");
    VB_REQUIRE(iters > 0 && reps > 0, VB_ERR_INVALID, "iters and reps must be positive");
    VB_CUDA(cudaSetDevice(ctx->device));
    VB_CUDA(cudaFuncSetAttribute(k_probe_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PROBE_SMEM));
    int rc;
    if ((rc = ctx->ws_ensure(WS_MISC, 4096 * sizeof(uint32_t)))) return rc;
    cudaEvent_t a, b;
    VB_CUDA(cudaEventCreate(&a));
    VB_CUDA(cudaEventCreate(&b));
    const uint32_t grid = (uint32_t)ctx->sm_count;
    const uint32_t kelems = kind == 0 ? 64u : kind == 1 ? 32u : 16u;
    float best = 1e30f;
    for (uint32_t r = 0; r < reps + 1; r++) {   // first launch is warm-up
        cudaEventRecord(a, ctx->stream);
        k_probe_umma<<<grid, 128, PROBE_SMEM, ctx->stream>>>(kind, n_cols, iters, ctx->ws[WS_MISC].as<uint32_t>());
        cudaEventRecord(b, ctx->stream);
        cudaError_t e = cudaEventSynchronize(b);
        if (e != cudaSuccess) {
            cudaEventDestroy(a);
            cudaEventDestroy(b);
            set_error("k_probe_umma: %s", cudaGetErrorString(e));
            return VB_ERR_CUDA;
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms < best) best = ms;
        ctx->launches++;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *best_ms = best;
    *flop_per_launch = 2.0 * 128.0 * (double)n_cols * (double)kelems * 4.0 * (double)iters * (double)grid;
    return VB_OK;
}
