// drain_probe.cu — the matcher's 16-column drain slice (two 8-column max trees, two keys, one offer2) on values that come
// from shared memory instead of TMEM, nothing else running: what the ALU pipe sustains for exactly this instruction mix.
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ float min3(float a, float b, float c) { return fminf(fminf(a, b), c); }
__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
__device__ __forceinline__ void offer2(float a, float b, float &r0, float &r1) {
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    r1 = min3(r1, fmaxf(r0, lo), hi);
    r0 = fminf(r0, lo);
}
__device__ __forceinline__ void slice(const float (&v)[16], float base, float &r0, float &r1) {
    float g[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const float vmax = fmaxf(max3(max3(v[8 * j], v[8 * j + 1], v[8 * j + 2]), max3(v[8 * j + 3], v[8 * j + 4], v[8 * j + 5]), v[8 * j + 6]), v[8 * j + 7]);
        g[j] = __fadd_rn(__fsub_rn((float)(8 * j), vmax), base);
    }
    offer2(g[0], g[1], r0, r1);
}
constexpr int ITERS = 4096;
__global__ void __launch_bounds__(1024) k_drain(const float *in, float *out, long long *clk, int mode) {
    __shared__ float4 s[32 * 4 * 8];   // 8 variants x 32 lanes x 16 floats
    for (int i = threadIdx.x; i < 32 * 4 * 8; i += blockDim.x) s[i] = reinterpret_cast<const float4 *>(in)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float r0 = 3e7f, r1 = 3e7f, base = 4194304.f;
    float v[16];
#pragma unroll
    for (int q = 0; q < 16; q++) v[q] = reinterpret_cast<const float *>(s)[64 + q * 33 + lane];
    const long long t0 = clock64();
#pragma unroll 2
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int q = 0; q < 16; q++) asm volatile("" : "+f"(v[q]));   // opaque: the values "arrive" anew every iteration at no cost
        if (mode == 0) slice(v, base, r0, r1);
        else { float a = v[0]; 
#pragma unroll
            for (int q = 1; q < 16; q++) a += v[q];   // FADD-only control: 15 fma-pipe instructions on the same loads
            r0 += a; }
        base += 16.f;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = r0 + r1;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
int main() {
    float *in, *out; long long *clk;
    cudaMalloc(&in, 32 * 4 * 8 * 16); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
    float h[32 * 4 * 8 * 4];
    for (int i = 0; i < 32 * 4 * 8 * 4; i++) h[i] = (float)(((i * 2654435761u) >> 20) & 511) * 32.f - 8192.f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 2; mode++)
        for (int threads : {128, 256, 512, 768, 1024}) {
            k_drain<<<148, threads>>>(in, out, clk, mode);
            k_drain<<<148, threads>>>(in, out, clk, mode);
            cudaDeviceSynchronize();
            long long c[148]; cudaMemcpy(c, clk, sizeof(c), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; i++) avg += (double)c[i]; avg /= 148;
            const double wps = threads / 32 / 4.0;
            printf("%s warps/SMSP=%.0f: %.1f clk per slice per warp, %.1f clk per slice per SMSP (%s)\n", mode == 0 ? "drain slice (13 ALU + 4 FADD, values in registers)" : "FADD control", wps,
                   avg / ITERS, avg / ITERS / wps, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
