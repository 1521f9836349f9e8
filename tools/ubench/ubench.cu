// Latency micro-benchmarks for the instructions the Riccati loop depends on (one warp, one block).
#include <cstdio>
#include <cuda_runtime.h>
#define REP 512
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
#define KEEP(x) asm volatile("" : "+d"(x) :: "memory")
__global__ void k_lat(double* out, long long* cyc, double seed, double b, double c) {
    __shared__ double sm[256];
    for (int i = threadIdx.x; i < 256; i += 32) sm[i] = (double)((i * 7 + 1) % 256);
    __syncwarp();
    double a = seed + threadIdx.x;
    long long t0, t1;
    // dependent DFMA
    t0 = clk();
#pragma unroll
    for (int i = 0; i < REP; i++) a = fma(a, b, c);
    KEEP(a); t1 = clk(); cyc[0] = t1 - t0;
    // dependent DADD
    t0 = clk();
#pragma unroll
    for (int i = 0; i < REP; i++) a = a + c;
    KEEP(a); t1 = clk(); cyc[1] = t1 - t0;
    // dependent DMUL
    t0 = clk();
#pragma unroll
    for (int i = 0; i < REP; i++) a = a * b;
    KEEP(a); t1 = clk(); cyc[2] = t1 - t0;
    // independent DFMA x4 (throughput, one warp)
    double a1 = a + 1, a2 = a + 2, a3 = a + 3;
    t0 = clk();
#pragma unroll
    for (int i = 0; i < REP; i++) { a = fma(a, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); }
    KEEP(a); t1 = clk(); cyc[3] = t1 - t0;
    a += a1 + a2 + a3;
    // dependent LDS.64 (pointer chase through shared)
    int idx = threadIdx.x;
    t0 = clk();
#pragma unroll
    for (int i = 0; i < REP; i++) idx = (int)sm[idx & 255];
    KEEP(a); t1 = clk(); cyc[4] = t1 - t0;
    a += idx;
    // dependent double shuffle
    t0 = clk();
#pragma unroll
    for (int i = 0; i < REP; i++) a = __shfl_xor_sync(0xffffffffu, a, 1);
    KEEP(a); t1 = clk(); cyc[5] = t1 - t0;
    // STS -> bar.warp.sync -> LDS round trip
    t0 = clk();
#pragma unroll
    for (int i = 0; i < REP; i++) { sm[threadIdx.x] = a; __syncwarp(); a = sm[(threadIdx.x + 1) & 31]; __syncwarp(); }
    KEEP(a); t1 = clk(); cyc[6] = t1 - t0;
    // dependent division
    t0 = clk();
#pragma unroll 16
    for (int i = 0; i < REP; i++) a = 1.0 / (a + 1.5);
    KEEP(a); t1 = clk(); cyc[7] = t1 - t0;
    // dependent FFMA for comparison
    float f = (float)a;
    t0 = clk();
#pragma unroll
    for (int i = 0; i < REP; i++) f = fmaf(f, 1.0000001f, 1e-9f);
    KEEP(a); t1 = clk(); cyc[8] = t1 - t0;
    // sincos
    t0 = clk();
#pragma unroll 8
    for (int i = 0; i < 64; i++) { double s, cc; sincos(a, &s, &cc); a = s + cc; }
    KEEP(a); t1 = clk(); cyc[9] = t1 - t0;
    out[threadIdx.x] = a + f;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 32 * 8); cudaMalloc(&c, 16 * 8);
    for (int r = 0; r < 2; r++) k_lat<<<1, 32>>>(d, c, 0.5, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, c, 16 * 8, cudaMemcpyDeviceToHost);
    const char* nm[] = {"dep DFMA", "dep DADD", "dep DMUL", "4x indep DFMA (per group)", "dep LDS.64", "dep SHFL f64", "STS+sync+LDS+sync", "dep 1/x f64", "dep FFMA", "sincos f64 (per call, /64)"};
    for (int i = 0; i < 10; i++) printf("%-28s %8.2f cycles\n", nm[i], (double)h[i] / (i == 9 ? 64 : REP));
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
