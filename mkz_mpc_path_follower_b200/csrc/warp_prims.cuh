// warp_prims.cuh -- the few warp-collective primitives the solver uses.
//
// On the device these are the sm_100a intrinsics.  With MPC_HOST_EMU defined (tests only,
// tests/emu/) the same names resolve to a single-threaded 32-lane coroutine emulator so the
// exact kernel source can be stepped on a CPU; the shipped library never defines it.
#pragma once

#ifdef MPC_HOST_EMU
#include "warp_emu.h"
#else
#include <cuda_runtime.h>
#include <math.h>

#define MPC_DEV __device__ __forceinline__
#define MPC_DEV_NOINLINE __device__ __noinline__
#define MPC_FULL 0xffffffffu

namespace mpcb200 {
MPC_DEV int lane_id() { return threadIdx.x & 31; }
// block-level primitives: used only when several warps share one long-horizon problem (one team = one block)
MPC_DEV int thread_in_block() { return threadIdx.x; }
MPC_DEV void block_sync() { asm volatile("bar.sync 0;" ::: "memory"); }
MPC_DEV bool block_all(bool p) { return __syncthreads_and(p ? 1 : 0) != 0; }
MPC_DEV double shfl(double v, int src) {   // src is taken modulo 32 by the hardware
    const int lo = __shfl_sync(MPC_FULL, __double2loint(v), src), hi = __shfl_sync(MPC_FULL, __double2hiint(v), src);
    return __hiloint2double(hi, lo);
}
MPC_DEV int shfl(int v, int src) { return __shfl_sync(MPC_FULL, v, src); }
MPC_DEV double shfl_down(double v, int d) { return __shfl_down_sync(MPC_FULL, v, d); }
MPC_DEV double shfl_up(double v, int d) { return __shfl_up_sync(MPC_FULL, v, d); }
MPC_DEV double shfl_xor(double v, int m) { return __shfl_xor_sync(MPC_FULL, v, m); }
MPC_DEV void syncwarp() { asm volatile("bar.warp.sync 0xffffffff;" ::: "memory"); }
MPC_DEV bool warp_all(bool p) { return __all_sync(MPC_FULL, p); }
MPC_DEV bool warp_any(bool p) { return __any_sync(MPC_FULL, p); }
MPC_DEV void mpc_sincos(double x, double* s, double* c) { sincos(x, s, c); }
// 1 / x for a positive, normal x (checked by the caller): hardware seed + two Newton steps, no special
// cases -- about half the instructions of the IEEE division sequence; accurate to ~1 ulp
MPC_DEV double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
MPC_DEV float fast_log2(float x) { return __log2f(x); }
MPC_DEV float fast_exp2(float x) { return exp2f(x); }
// a * b and a + b rounded separately (never contracted into an FMA): the reference generator is numpy, which
// rounds every operation, and its nearest-sample search and np.interp are reproduced bit for bit
MPC_DEV double mul_rn(double a, double b) { return __dmul_rn(a, b); }
MPC_DEV double add_rn(double a, double b) { return __dadd_rn(a, b); }
MPC_DEV void st_global_v2(double* p, double x, double y) { *reinterpret_cast<double2*>(p) = make_double2(x, y); }
MPC_DEV double int2_as_double(int lo, int hi) { return __hiloint2double(hi, lo); }

// Shared memory through 32-bit shared-window addresses and explicit ld/st.shared: the base is an
// opaque register value, so the compiler cannot rematerialise generic->shared conversions or lane
// role arithmetic inside the Riccati loops.  Offsets are in BYTES on the device.
typedef unsigned smem_t;
#define SO(x) ((x) * 8)
MPC_DEV smem_t smem_base(double* p) { return (smem_t)__cvta_generic_to_shared(p); }
MPC_DEV double lds(smem_t b, int off) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(b + off) : "memory");
    return v;
}
struct d2 { double x, y; };
MPC_DEV d2 lds2(smem_t b, int off) {
    d2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(b + off) : "memory");
    return v;
}
MPC_DEV void sts(smem_t b, int off, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(b + off), "d"(v) : "memory"); }
MPC_DEV void sts2(smem_t b, int off, double x, double y) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(b + off), "d"(x), "d"(y) : "memory"); }
MPC_DEV int launder(int v) { int r; asm volatile("mov.b32 %0, %1;" : "=r"(r) : "r"(v)); return r; }
// one lane's row of the role table (20 ints, 16-byte aligned) through the read-only path
MPC_DEV void ld_roles(const int* p, int* out, int n = 20) {
    const int4* q = reinterpret_cast<const int4*>(p);
    for (int i = 0; i < n / 4; i++) { const int4 v = __ldg(q + i); out[4 * i] = v.x; out[4 * i + 1] = v.y; out[4 * i + 2] = v.z; out[4 * i + 3] = v.w; }
}
}  // namespace mpcb200
#endif
