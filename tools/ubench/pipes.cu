// Throughput of DMMA (mma.sync m8n8k4 f64), SHFL, LDS on one SM, alone and combined: do they share a pipe?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
// what: bit0 = DMMA, bit1 = SHFL(f64 = 2 x 32-bit), bit2 = LDS.64 distinct addresses, bit3 = LDS.64 uniform address, bit4: dependent DMMA chain
__global__ void k(double* out, long long* cyc, int what, int active) {
    __shared__ double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-3;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x >> 5) * 1024;
    const unsigned addr = base + lane * 8, uaddr = base;
    double a = 1.0 + lane * 1e-3, b = 1.0 - lane * 1e-3, c0 = 0, c1 = 0, d0 = 0, d1 = 0, e0 = 0, e1 = 0, f0 = 0, f1 = 0, s = lane, acc = 0;
    __syncthreads();
    long long t0 = clk();
    if (lane < active) {
#pragma unroll 4
    for (int i = 0; i < 512; i++) {
        if (what & 1) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(e0), "+d"(e1) : "d"(a), "d"(b));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(f0), "+d"(f1) : "d"(a), "d"(b));
        }
        if (what & 16) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(c0), "d"(b));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(c1), "d"(b));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(c0), "d"(b));
        }
        if (what & 2) {
            s = __shfl_xor_sync(0xffffffffu, s, 1); acc += __shfl_xor_sync(0xffffffffu, a, 2);
            acc += __shfl_xor_sync(0xffffffffu, b, 4); acc += __shfl_xor_sync(0xffffffffu, a, 8);
        }
        if (what & 4) {
            double v0, v1, v2, v3;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v0) : "r"(addr + (i & 3) * 256));
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v1) : "r"(addr + 1024 + (i & 3) * 256));
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v2) : "r"(addr + 2048 + (i & 3) * 256));
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v3) : "r"(addr + 3072 + (i & 3) * 256));
            acc += (v0 + v1) + (v2 + v3);
        }
        if (what & 8) {
            double v0, v1, v2, v3;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v0) : "r"(uaddr + (i & 3) * 256));
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v1) : "r"(uaddr + 1024 + (i & 3) * 256));
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v2) : "r"(uaddr + 2048 + (i & 3) * 256));
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v3) : "r"(uaddr + 3072 + (i & 3) * 256));
            acc += (v0 + v1) + (v2 + v3);
        }
    }
    }
    __syncthreads();
    long long t1 = clk();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + c0 + c1 + d0 + d1 + e0 + e1 + f0 + f1 + s;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 1 << 20); cudaMalloc(&c, 1024);
    struct { int what; const char* nm; } T[] = {{1, "DMMA x4 indep"}, {16, "DMMA x4 dependent"}, {2, "SHFL f64 x4"}, {4, "LDS.64 distinct x4"}, {8, "LDS.64 uniform x4"},
                                                {1 | 2, "DMMA + SHFL"}, {2 | 4, "SHFL + LDS distinct"}, {1 | 4, "DMMA + LDS distinct"}};
    for (auto& t : T)
        for (int warps : {1, 4, 8, 16})
            for (int active : {32, 8}) {
                if ((t.what & 19) && active != 32) continue;
                for (int r = 0; r < 2; r++) k<<<1, warps * 32>>>(d, c, t.what, active);
                cudaDeviceSynchronize();
                long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
                printf("%-22s warps %2d active lanes %2d: %8.2f cycles per loop trip (4 ops each kind) -> %.2f cycles/op/SM\n", t.nm, warps, active, (double)h / 512.0, (double)h / 512.0 / (4.0 * warps));
            }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
