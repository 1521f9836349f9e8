// Shared-memory pipe cost of partially active warps: many warps hammer LDS; report cycles per LDS instruction per SM.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
template <int MODE>
__global__ void k(double* out, long long* cyc, int active, int stride) {
    __shared__ double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x >> 5) * 1024;
    const unsigned addr = base + (lane * stride) * 8;
    double acc = 0;
    __syncthreads();
    long long t0 = clk();
    if (lane < active) {
#pragma unroll 8
        for (int i = 0; i < 1024; i++) {
            if (MODE == 0) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr + (i & 7) * 8)); acc += v; }
            if (MODE == 1) { double v, w; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v), "=d"(w) : "r"(addr + (i & 3) * 16)); acc += v + w; }
            if (MODE == 2) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr + (i & 7) * 8), "d"(acc)); }
        }
    }
    __syncthreads();
    long long t1 = clk();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 1 << 20); cudaMalloc(&c, 1024);
    const int warps = 16;
    const char* nm[] = {"LDS.64", "LDS.128", "STS.64"};
    for (int mode = 0; mode < 3; mode++)
        for (int stride : {0, 1, 2, 6})
            for (int active : {32, 16, 8, 6, 1}) {
                for (int r = 0; r < 2; r++) {
                    if (mode == 0) k<0><<<1, warps * 32>>>(d, c, active, stride);
                    if (mode == 1) k<1><<<1, warps * 32>>>(d, c, active, stride);
                    if (mode == 2) k<2><<<1, warps * 32>>>(d, c, active, stride);
                }
                cudaDeviceSynchronize();
                long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
                printf("%-8s lane-stride %d doubles, %2d active lanes: %.2f cycles per warp-instruction (16 warps on one SM)\n", nm[mode], stride, active, (double)h / (1024.0 * warps));
            }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
