// tmem_probe.cu — development probe for a packed-integer drain of the fp4 matcher:
//   A  what tcgen05.ld .pack::16b returns and what tcgen05.st .unpack::16b writes (which halves, what happens to the others)
//   B  whether tcgen05.mma kind::mxf4 accumulates +-1 products exactly onto an accumulator pre-filled with 1.5 * 2^23 + 512
//   C  the rate of LDTM / STTM in the shapes such a drain would use, 16 warps at a time
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I vslam_b200/csrc -o tools/probe/tmem_probe tools/probe/tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include "tc_common.cuh"
using namespace vb::tc;

#define LD_X16(pack, addr, v)                                                                                                   \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16" pack ".b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),   \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                      \
                 : "r"(addr) : "memory")
#define ST_X16(unpack, addr, v)                                                                                                 \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16" unpack ".b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" \
                 ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), \
                 "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory")
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait16(uint32_t (&v)[16]) { tmem_wait_ld_regs16(v); }

constexpr uint32_t SMEM = 16384 + 32768 + 1024 + 64;
constexpr uint32_t MAGIC = 0x4B400200u;   // 1.5 * 2^23 + 512

// out: [0,16) plain read-back of what A wrote; [16,32) pack::16b read of the same columns (base 64); [32,48) pack read at
// base 64 + 16; [48,64) plain read after an unpack::16b store over pre-filled cells; [64, 64+32) accumulators after the MMAs
__global__ void __launch_bounds__(128, 1) k_semantics(uint32_t *out, int bmode, uint32_t magic, uint32_t sfa_word, uint32_t sfb_word) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = smem0, sB = smem0 + 16384, bar = sB + 32768, s_tmem = bar + 8;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // A: every e2m1 element +1. B: every 32-byte block (one K = 64 step of one row, wherever the swizzle puts it) holds
    // 31 bytes (+1, -1) and one byte that is (+1, +1) [bmode 0: block sum +2], (-1, -1) [1: -2] or (+1, -1) [2: 0]
    for (uint32_t i = threadIdx.x; i < 16384u / 4u; i += blockDim.x) asm volatile("st.shared.u32 [%0], %1;" ::"r"(sA + 4u * i), "r"(0x22222222u) : "memory");
    for (uint32_t i = threadIdx.x; i < 32768u / 4u; i += blockDim.x) {
        uint32_t w = 0x2A2A2A2Au;
        if ((i & 7u) == 0u) w = (w & 0xffffff00u) | (bmode == 0 ? 0x22u : bmode == 1 ? 0xAAu : 0x2Au);
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(sB + 4u * i), "r"(w) : "memory");
    }
    if (warp == 1 && lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 2) tmem_alloc(s_tmem, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(s_tmem));
    const uint32_t lane_base = tmem_base + ((warp * 32u) << 16);
    tmem_st16_const(lane_base, sfa_word);        // scale factors of A in columns [0, 16) (default 2^0)
    tmem_st16_const(lane_base + 16, sfb_word);   // scale factors of B in columns [16, 32)
    wait_st();
    // ---- A: pack / unpack semantics on columns [320, 352)
    uint32_t v[16], r[16];
#pragma unroll
    for (int j = 0; j < 16; j++) v[j] = ((0x1100u + j) << 16) | (0x2200u + j);
    ST_X16("", lane_base + 320, v);
#pragma unroll
    for (int j = 0; j < 16; j++) v[j] = ((0x3300u + j) << 16) | (0x4400u + j);
    ST_X16("", lane_base + 336, v);
    wait_st();
    LD_X16("", lane_base + 320, r); wait16(r);
    if (threadIdx.x == 0) for (int j = 0; j < 16; j++) out[j] = r[j];
    LD_X16(".pack::16b", lane_base + 320, r); wait16(r);
    if (threadIdx.x == 0) for (int j = 0; j < 16; j++) out[16 + j] = r[j];
    LD_X16(".pack::16b", lane_base + 336, r); wait16(r);
    if (threadIdx.x == 0) for (int j = 0; j < 16; j++) out[32 + j] = r[j];
#pragma unroll
    for (int j = 0; j < 16; j++) v[j] = ((0x5500u + j) << 16) | (0x6600u + j);
    ST_X16(".unpack::16b", lane_base + 320, v);
    wait_st();
    LD_X16("", lane_base + 320, r); wait16(r);
    if (threadIdx.x == 0) for (int j = 0; j < 16; j++) out[48 + j] = r[j];
    LD_X16("", lane_base + 336, r); wait16(r);
    if (threadIdx.x == 0) for (int j = 0; j < 16; j++) out[96 + j] = r[j];
    // ---- B: exact accumulation onto the magic constant, accumulator columns [32, 32 + 240)
#pragma unroll
    for (int j = 0; j < 16; j++) v[j] = magic;
    for (uint32_t c = 32; c < 32 + 240; c += 16) ST_X16("", lane_base + c, v);
    wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1 && lane == 0) {
        const uint32_t idesc4 = umma_idesc_mxf4(128, 240);
#pragma unroll
        for (int k = 0; k < 4; k++) umma_mxf4(tmem_base + 32, smem_desc_sw128(sA + k * 32), smem_desc_sw128(sB + k * 32), idesc4, tmem_base, tmem_base + 16, 1u);
        umma_commit(bar);
        mbar_wait(bar, 0);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    LD_X16("", lane_base + 32, r); wait16(r);
    if (threadIdx.x == 37) for (int j = 0; j < 16; j++) out[64 + j] = r[j];
    LD_X16("", lane_base + 32 + 224, r); wait16(r);
    if (threadIdx.x == 101) for (int j = 0; j < 16; j++) out[80 + j] = r[j];
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// C: 16 warps (4 per sub-partition), each over its lane quadrant: REPS x (64 columns) in the given mode
__global__ void __launch_bounds__(512, 1) k_rate(int mode, uint32_t reps, long long *clk, uint32_t *sink) {
    __shared__ uint32_t s_tmem;
    const uint32_t warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = s_tmem + (((warp & 3u) * 32u) << 16) + (warp >> 2) * 64u;   // 4 quadrants x 4 column parts
    uint32_t a[16], b[16], acc = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) a[j] = b[j] = MAGIC + j;
    ST_X16("", base, a); ST_X16("", base + 16, a); ST_X16("", base + 32, a); ST_X16("", base + 48, a);
    wait_st();
    __syncthreads();
    const long long t0 = clock64();
    for (uint32_t it = 0; it < reps; it++) {
        if (mode == 0) {          // plain loads: 64 columns = 4 x LDTM.x16
            LD_X16("", base, a); LD_X16("", base + 16, b); wait16(a); wait16(b);
            acc += a[0] ^ b[15];
            LD_X16("", base + 32, a); LD_X16("", base + 48, b); wait16(a); wait16(b);
            acc += a[3] ^ b[7];
        } else if (mode == 1) {   // packed loads: 64 columns = 2 x LDTM.x16.pack::16b
            LD_X16(".pack::16b", base, a); LD_X16(".pack::16b", base + 32, b); wait16(a); wait16(b);
            acc += a[0] ^ b[15];
        } else if (mode == 2) {   // plain stores
            ST_X16("", base, a); ST_X16("", base + 16, a); ST_X16("", base + 32, a); ST_X16("", base + 48, a);
            wait_st();
        } else if (mode == 3) {   // unpack stores: 64 columns = 2 x STTM.x16.unpack::16b
            ST_X16(".unpack::16b", base, a); ST_X16(".unpack::16b", base + 32, a);
            wait_st();
        } else if (mode == 4) {   // packed load + unpack store of the same 64 columns (the drain's hand-back)
            LD_X16(".pack::16b", base, a); LD_X16(".pack::16b", base + 32, b); wait16(a); wait16(b);
            acc += a[0] ^ b[15];
            ST_X16(".unpack::16b", base, a); ST_X16(".unpack::16b", base + 32, b); wait_st();
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * 512 + threadIdx.x] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(s_tmem, 512); }
}

int main() {
    uint32_t *out; long long *clk; uint32_t *sink;
    cudaMalloc(&out, 4096); cudaMalloc(&clk, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
    cudaFuncSetAttribute(k_semantics, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    for (int bmode = 0; bmode < 3; bmode++) {
        cudaMemset(out, 0, 4096);
        k_semantics<<<1, 128, SMEM>>>(out, bmode, MAGIC, 0x7f7f7f7fu, 0x7f7f7f7fu);
        cudaError_t e = cudaDeviceSynchronize();
        uint32_t h[128];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("== semantics, B block sum %s: %s\n", bmode == 0 ? "+2" : bmode == 1 ? "-2" : "0", cudaGetErrorString(e));
        if (bmode == 0) {
            const char *names[] = {"plain read [320,336)", "pack::16b read @320", "pack::16b read @336", "plain read [320,336) after unpack::16b store @320"};
            for (int s = 0; s < 4; s++) { printf("%-50s", names[s]); for (int j = 0; j < 16; j++) printf(" %08x", h[16 * s + j]); printf("\n"); }
            printf("%-50s", "plain read [336,352) after that store"); for (int j = 0; j < 16; j++) printf(" %08x", h[96 + j]); printf("\n");
        }
        printf("accumulators (magic %08x, expected magic %+d): ", MAGIC, bmode == 0 ? 8 : bmode == 1 ? -8 : 0);
        for (int j = 0; j < 32; j++) printf(" %08x", h[64 + j]);
        printf("\n");
    }
    // D: the same accumulation in the DENORMAL range: magic 0x00004080 (upper half-word zero, so that unpack::16b stores can
    // re-arm an accumulator), scale factors 2^-71 (A) * 2^-72 (B): a product of +-1 is +-64 units of 2^-149
    for (int bmode = 0; bmode < 3; bmode++) {
        cudaMemset(out, 0, 4096);
        k_semantics<<<1, 128, SMEM>>>(out, bmode, 0x00004080u, 0x38383838u, 0x37373737u);
        cudaError_t e = cudaDeviceSynchronize();
        uint32_t h[128];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("== denormal accumulators, B block sum %s: %s\n", bmode == 0 ? "+2" : bmode == 1 ? "-2" : "0", cudaGetErrorString(e));
        printf("accumulators (magic 00004080, expected %08x): ", 0x4080 + (bmode == 0 ? 512 : bmode == 1 ? -512 : 0));
        for (int j = 0; j < 32; j++) printf(" %08x", h[64 + j]);
        printf("\n");
    }
    // E: equal scale factors 2^-72 * 2^-72 (no assumption on which columns hold A's and which B's): +-32 units
    for (int bmode = 0; bmode < 2; bmode++) {
        cudaMemset(out, 0, 4096);
        k_semantics<<<1, 128, SMEM>>>(out, bmode, 0x00004080u, 0x37373737u, 0x37373737u);
        cudaError_t e = cudaDeviceSynchronize();
        uint32_t h[128];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("== denormal accumulators, equal scales, B block sum %s: %s\n", bmode == 0 ? "+2" : "-2", cudaGetErrorString(e));
        printf("accumulators (magic 00004080, expected %08x): ", 0x4080 + (bmode == 0 ? 256 : -256));
        for (int j = 0; j < 32; j++) printf(" %08x", h[64 + j]);
        printf("\n");
    }
    const char *mn[] = {"LDTM.x16 x4 (64 cols)", "LDTM.x16.pack::16b x2 (64 cols)", "STTM.x16 x4 (64 cols)", "STTM.x16.unpack::16b x2 (64 cols)", "packed load + unpack store"};
    for (int mode = 0; mode < 5; mode++) {
        const uint32_t reps = 4096;
        k_rate<<<148, 512>>>(mode, reps, clk, sink);
        k_rate<<<148, 512>>>(mode, reps, clk, sink);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; i++) avg += (double)h[i]; avg /= 148;
        printf("%-36s %.1f clk per 128 lanes x 256 columns (16 warps x 64 columns), %s\n", mn[mode], avg / reps, cudaGetErrorString(e));
    }
    return 0;
}
