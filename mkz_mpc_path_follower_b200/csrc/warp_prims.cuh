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
MPC_DEV double shfl(double v, int src) { return __shfl_sync(MPC_FULL, v, src); }
MPC_DEV int shfl(int v, int src) { return __shfl_sync(MPC_FULL, v, src); }
MPC_DEV double shfl_down(double v, int d) { return __shfl_down_sync(MPC_FULL, v, d); }
MPC_DEV double shfl_up(double v, int d) { return __shfl_up_sync(MPC_FULL, v, d); }
MPC_DEV double shfl_xor(double v, int m) { return __shfl_xor_sync(MPC_FULL, v, m); }
MPC_DEV void syncwarp() { __syncwarp(); }
MPC_DEV bool warp_all(bool p) { return __all_sync(MPC_FULL, p); }
MPC_DEV bool warp_any(bool p) { return __any_sync(MPC_FULL, p); }
MPC_DEV void mpc_sincos(double x, double* s, double* c) { sincos(x, s, c); }
}  // namespace mpcb200
#endif
