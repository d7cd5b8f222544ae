// pipe_probe.cu — reciprocal throughput (clk per warp instruction per SM sub-partition) of the instructions the matcher's
// drain is made of, and of candidate replacements. Development aid; build: nvcc -gencode arch=compute_100a,code=sm_100a.
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>

constexpr int CH = 8;        // independent chains per thread
constexpr int ITERS = 2048;

template <int OP> __device__ __forceinline__ void step(uint32_t (&x)[CH], uint32_t a, uint32_t b) {
#pragma unroll
    for (int i = 0; i < CH; i++) {
        float f = __uint_as_float(x[i]), fa = __uint_as_float(a), fb = __uint_as_float(b);
        if (OP == 0) x[i] = __float_as_uint(fmaxf(f, fa));                                  // FMNMX
        if (OP == 1) x[i] = __float_as_uint(fmaxf(fmaxf(f, fa), fb));                       // FMNMX3
        if (OP == 2) x[i] = __float_as_uint(__fadd_rn(f, fa));                              // FADD
        if (OP == 3) x[i] = __float_as_uint(__fmaf_rn(f, fa, fb));                          // FFMA 3-reg
        if (OP == 4) x[i] = x[i] * a + b;                                                   // IMAD 3-reg
        if (OP == 5) x[i] = __vmaxu2(x[i], a);                                              // VIMNMX.U16x2
        if (OP == 6) { __half2 h = __hmax2(*reinterpret_cast<__half2 *>(&x[i]), *reinterpret_cast<__half2 *>(&a)); x[i] = *reinterpret_cast<uint32_t *>(&h); }
        if (OP == 7) { __half2 h = __floats2half2_rn(f, fa); x[i] = *reinterpret_cast<uint32_t *>(&h) ^ b; }   // F2FP + LOP3
        if (OP == 8) x[i] = __byte_perm(x[i], a, 0x5410);                                   // PRMT
        if (OP == 9) x[i] = max(max((int)x[i], (int)a), (int)b);                            // VIMNMX3 ?
        if (OP == 10) x[i] = __vmaxu2(__vmaxu2(x[i], a), b);                                // VIMNMX3.U16x2 ?
        if (OP == 11) x[i] = x[i] * 65536u + a;                                             // IMAD imm
        if (OP == 12) x[i] = __float_as_uint(__fmaf_rn(f, 1.5f, fa));                       // FFMA imm
        if (OP == 13) x[i] = max((int)x[i], (int)a);                                        // VIMNMX 2-input
        if (OP == 14) x[i] = __vmaxs2(x[i], a);                                             // VIMNMX.S16x2
        if (OP == 15) x[i] = __float_as_uint(__fadd_rn(f, fabsf(fa)));                      // FADD |b|
    }
}
// two different ops interleaved 1:1 (dual-pipe issue)
template <int OPA, int OPB> __device__ __forceinline__ void step2(uint32_t (&x)[CH], uint32_t (&y)[CH], uint32_t a, uint32_t b) {
    step<OPA>(x, a, b);
    step<OPB>(y, a, b);
}

template <int OPA, int OPB>
__global__ void __launch_bounds__(512) k_probe(const uint32_t *in, uint32_t *out, long long *clk) {
    uint32_t x[CH], y[CH];
    const uint32_t a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
#pragma unroll
    for (int i = 0; i < CH; i++) { x[i] = in[64 + i] + threadIdx.x; y[i] = in[80 + i] ^ threadIdx.x; }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (OPB < 0) { step<OPA>(x, a, b); step<OPA>(y, a, b); }
        else step2<OPA, OPB>(x, y, a, b);
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += x[i] ^ y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OPA, int OPB> void run(const char *name, const uint32_t *in, uint32_t *out, long long *clk, int threads) {
    k_probe<OPA, OPB><<<148, threads>>>(in, out, clk);
    k_probe<OPA, OPB><<<148, threads>>>(in, out, clk);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; i++) avg += (double)h[i];
    avg /= 148;
    const double instr_per_warp = (double)ITERS * 2 * CH;          // warp instructions of the probed kind(s) per warp
    const double warps_per_smsp = threads / 32 / 4.0;
    printf("%-34s threads=%4d  clk per warp-instr per SMSP = %.3f   (err %s)\n", name, threads, avg / (instr_per_warp * warps_per_smsp),
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    uint32_t *in, *out; long long *clk;
    cudaMalloc(&in, 4096); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
    uint32_t h[128];
    for (int i = 0; i < 128; i++) h[i] = 0x3f800000u + i * 7919u;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int threads : {128, 512}) {
        run<0, -1>("FMNMX", in, out, clk, threads);
        run<1, -1>("FMNMX3", in, out, clk, threads);
        run<2, -1>("FADD", in, out, clk, threads);
        run<15, -1>("FADD |b|", in, out, clk, threads);
        run<3, -1>("FFMA 3-reg", in, out, clk, threads);
        run<12, -1>("FFMA imm", in, out, clk, threads);
        run<4, -1>("IMAD 3-reg", in, out, clk, threads);
        run<11, -1>("IMAD imm", in, out, clk, threads);
        run<13, -1>("VIMNMX s32", in, out, clk, threads);
        run<9, -1>("VIMNMX3 s32 (?)", in, out, clk, threads);
        run<5, -1>("VIMNMX.U16x2", in, out, clk, threads);
        run<14, -1>("VIMNMX.S16x2", in, out, clk, threads);
        run<10, -1>("VIMNMX3.U16x2 (?)", in, out, clk, threads);
        run<6, -1>("HMNMX2", in, out, clk, threads);
        run<7, -1>("F2FP.PACK_AB + LOP3", in, out, clk, threads);
        run<8, -1>("PRMT", in, out, clk, threads);
        run<0, 2>("FMNMX + FADD 1:1", in, out, clk, threads);
        run<1, 2>("FMNMX3 + FADD 1:1", in, out, clk, threads);
        run<1, 11>("FMNMX3 + IMAD imm 1:1", in, out, clk, threads);
        run<5, 11>("VIMNMX.U16x2 + IMAD imm 1:1", in, out, clk, threads);
        run<5, 2>("VIMNMX.U16x2 + FADD 1:1", in, out, clk, threads);
        run<6, 2>("HMNMX2 + FADD 1:1", in, out, clk, threads);
        run<1, 1>("FMNMX3 + FMNMX3", in, out, clk, threads);
    }
    return 0;
}
