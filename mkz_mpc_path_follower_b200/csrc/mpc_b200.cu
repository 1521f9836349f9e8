// mpc_b200.cu -- C ABI (include/mpc_b200.h) and kernel launches for sm_100a.
//
// One handle = one GPU + one stream.  Problems are independent: a persistent grid of warps
// pulls problem indices from a device counter (iteration counts differ per problem, so a
// dynamic queue keeps the SMs level), solves each with mpc_kernel.cuh::solve_problem and
// writes the per-problem record.  No CPU path exists in this library.
#include "../../include/mpc_b200.h"
#include "mpc_kernel.cuh"
#include "tpp_solver.cuh"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <utility>

using namespace mpcb200;

namespace {

constexpr int WARPS_PER_BLOCK = 4;

// 168 registers per thread: three blocks (12 warps) per SM at the N = 20 shared-memory footprint.
// The iterate's bound multipliers, the step and the model evaluation live in per-thread shared-memory
// fields (mpc_kernel.cuh, LF_*), which is what lets the solver fit.
#ifndef MPC_MIN_BLOCKS
#define MPC_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, MPC_MIN_BLOCKS)
mpc_solve_kernel(const KCfg cfg, const BatchPtrs io, const RefGen rg, const long long B, unsigned long long* counter) {
    extern __shared__ double smem_all[];
    const int warp = threadIdx.x >> 5;
    const smem_t smem = smem_base(smem_all + (size_t)warp * smem_doubles_per_team(cfg.N));
    TeamSolver<1>::init_work(smem, cfg.N);
    for (;;) {
        unsigned long long b = 0;
        if ((threadIdx.x & 31) == 0) b = atomicAdd(counter, 1ULL);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= (unsigned long long)B) break;
        solve_problem<1>(cfg, io, rg, (long)b, smem);
        __syncwarp();
    }
}

// Frenet-frame variant (MKZMPCPathFollowerFrenet.jl): same driver, dense s / e_y columns (mpc_kernel.cuh, MODEL 1);
// 62-double records, 21.3 KB of shared memory per problem at N = 20.  Two builds: blocks of four warps at 255 registers
// (two per SM) and blocks of three at 168 registers with spill code (up to four per SM: 12 warps at the N = 8 footprint);
// mpcb200_create_frenet picks per horizon (see there)
template <int WPB>
__global__ void __launch_bounds__(WPB * 32, WPB == 3 ? 3 : 2)
mpc_solve_frenet_kernel(const KCfg cfg, const BatchPtrs io, const RefGen rg, const long long B, unsigned long long* counter) {
    extern __shared__ double smem_all[];
    const int warp = threadIdx.x >> 5;
    const smem_t smem = smem_base(smem_all + (size_t)warp * smem_doubles_per_team(cfg.N, 1));
    TeamSolver<1, 1>::init_work(smem, cfg.N);
    for (;;) {
        unsigned long long b = 0;
        if ((threadIdx.x & 31) == 0) b = atomicAdd(counter, 1ULL);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= (unsigned long long)B) break;
        solve_problem<1, 1>(cfg, io, rg, (long)b, smem);
        __syncwarp();
    }
}

// Long horizons (32 <= N <= 95): one problem per block of W = 2 or 3 warps, thread k = stage k.
// register budget (measured at N = 40 / 80): W = 2 -> 6 blocks (12 warps) per SM at 168 registers; W = 3 -> 2 blocks at 255
#ifndef MPC_LONG_MIN_BLOCKS2
#define MPC_LONG_MIN_BLOCKS2 6
#define MPC_LONG_MIN_BLOCKS3 2
#endif
template <int W, int MODEL = 0>
__global__ void __launch_bounds__(W * 32, MODEL ? 2 : (W == 2 ? MPC_LONG_MIN_BLOCKS2 : MPC_LONG_MIN_BLOCKS3))
mpc_solve_long_kernel(const KCfg cfg, const BatchPtrs io, const RefGen rg, const long long B, unsigned long long* counter) {
    extern __shared__ double smem_all[];
    __shared__ unsigned long long next_problem;
    const smem_t smem = smem_base(smem_all);
    TeamSolver<W, MODEL>::init_work(smem, cfg.N);
    for (;;) {
        if (threadIdx.x == 0) next_problem = atomicAdd(counter, 1ULL);
        __syncthreads();
        const unsigned long long b = next_problem;
        __syncthreads();
        if (b >= (unsigned long long)B) break;
        solve_problem<W, MODEL>(cfg, io, rg, (long)b, smem);
    }
}


// Large batches of the XY model: one THREAD per problem (tpp_solver.cuh).  A lane owns a slot of the state arrays
// ([field][stage][slot]: the 32 lanes of a warp touch 256 contiguous bytes per access) for the whole launch; when its
// problem ends it writes the result and takes the next problem index from the device counter.  The warp leaves when the
// counter has run past the batch and every lane is idle.
#ifndef MPC_TPP_BLOCK
#define MPC_TPP_BLOCK 256
#define MPC_TPP_MIN_BLOCKS 1
#endif
template <int MODEL>
__global__ void __launch_bounds__(MPC_TPP_BLOCK, MPC_TPP_MIN_BLOCKS)
mpc_solve_tpp_kernel(const KCfg cfg, const BatchPtrs io, const long long B, double* st, double* filt, unsigned long long* counter,
                     const double v_des0) {
    const long slot = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const TppMemT<MODEL> mem(st, filt, cfg.N, slot);
    TppSolverT<MODEL> sv(cfg, mem);
    bool alive = true;   // the counter has not run past the batch for this lane yet
    long b = -1;
    while (__syncthreads_or(alive)) {
        if (alive && b < 0) {
            const unsigned long long nb = atomicAdd(counter, 1ULL);
            if (nb >= (unsigned long long)B) alive = false;
            else { b = (long)nb; sv.begin(io, b, v_des0); }
        }
        const bool done = sv.tick(b >= 0);
        if (b >= 0 && done) { sv.finish(io, b); b = -1; }
    }
}

// ---- mpcb200_solve_batch_on_path with the thread-per-problem layout: the waypoints of the whole batch first (warp = problem,
// get_waypoints as in solve_problem), into `ref` = [B][3][N+1] (the caller's ref_out if it asked for one), which the solve kernel
// then reads like references that came from the host; a problem whose path id was never set gets NaN references (its solve ends
// at once) and on_path_invalid_kernel gives it the answer solve_problem gives: MPCB200_ERROR, zero commands.
__global__ void __launch_bounds__(128)
on_path_ref_kernel(const KCfg cfg, const long long B, const double* state, const RefGen rg, double* ref) {
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31, N = cfg.N;
    TeamSolver<1> S(cfg, (smem_t)0);   // (one-warp teams use no shared memory in get_waypoints)
    for (long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < B; b += warps) {
        const int pid = rg.path_of[b];
        double xr = __longlong_as_double(0x7ff8000000000000LL), yr = xr, pr = xr;
        bool sc = false;
        if (!(pid < 0 || pid > 2 || rg.paths[pid].n < 2))
            sc = S.get_waypoints(rg.paths[pid], cfg.dt, state[4 * b], state[4 * b + 1], state[4 * b + 2], !rg.track_using_time, rg.target_vel, xr, yr, pr);
        if (lane <= N) { double* ro = ref + 3LL * (N + 1) * b; ro[lane] = xr; ro[(N + 1) + lane] = yr; ro[2 * (N + 1) + lane] = pr; }
        if (rg.stop && lane == 0) rg.stop[b] = sc ? 1 : 0;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(128)
on_path_invalid_kernel(const long long B, const RefGen rg, const BatchPtrs io) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int pid = rg.path_of[b];
    if (!(pid < 0 || pid > 2 || rg.paths[pid].n < 2)) return;
    if (io.u0) { io.u0[2 * b] = 0.0; io.u0[2 * b + 1] = 0.0; }
    if (io.cost) io.cost[b] = 0.0;
    if (io.status) io.status[b] = 4;
    if (io.iters) io.iters[b] = 0;
    if (io.rec) { st_global_v2(io.rec + 4 * b, 0.0, 0.0); st_global_v2(io.rec + 4 * b + 2, 0.0, int2_as_double(4, 0)); }
    if (io.resto) io.resto[b] = 0;
}

// ---- Closed-loop fleets with the thread-per-problem layout: one control period = three launches over the whole fleet
// instead of one persistent kernel that carries four vehicles per block through all their periods.
//   plant      thread = vehicle: the ten 100 Hz publishes of ten Euler sub-steps (vehicle_simulator.py:58-112): every lane of a
//              warp integrates its own vehicle (in mpc_rollout_kernel one warp per block does it for four while three wait)
//   waypoints  warp = vehicle: get_waypoints (ref_gps_traj.py:131-218) written straight into the vehicle's slot of the solver state
//   solve      thread = vehicle: the slot still holds the previous solution = the warm start (mpc_cmd_pub.jl:115-141), all
//              vehicles start together and take about the same handful of iterations, which is where this layout is at its best
// Vehicle state is structure-of-arrays [12][Bp]: X Y psi vx vy wz acc df | acc_des df_des | d_f_current acc_current.
#define RV_NF 12
__global__ void __launch_bounds__(128)
rollout_plant_kernel(const long long B, const long long Bp, double* veh, const double* pose0, const int first) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= B) return;
    double st[8], ad, dd;
    if (first) {
        st[0] = pose0[3 * v]; st[1] = pose0[3 * v + 1]; st[2] = pose0[3 * v + 2];
        for (int i = 3; i < 8; i++) st[i] = 0.0;
        ad = 0.0; dd = 0.0;
        veh[8 * Bp + v] = 0.0; veh[9 * Bp + v] = 0.0; veh[10 * Bp + v] = 0.0; veh[11 * Bp + v] = 0.0;
    } else {
        for (int i = 0; i < 8; i++) st[i] = veh[i * Bp + v];
        ad = veh[8 * Bp + v]; dd = veh[9 * Bp + v];
    }
    MPC_NOUNROLL for (int i = 0; i < 10; i++) plant_step(st, ad, dd);
    for (int i = 0; i < 8; i++) veh[i * Bp + v] = st[i];
}

__global__ void __launch_bounds__(128)
rollout_waypoints_kernel(const KCfg cfg, const long long B, const long long Bp, const double* veh, const int* path_of, const RefGen rg,
                         double* tpp_state, double* tpp_filt, int* stop, const int first) {
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    TeamSolver<1> S(cfg, (smem_t)0);   // (one-warp teams use no shared memory in get_waypoints)
    for (long long v = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < B; v += warps) {
        if (first && lane == 0) stop[v] = 0;
        if (!first && stop[v]) continue;   // the stop latch (mpc_cmd_pub.jl:102-111): no more solves, no more references
        double xr, yr, pr;
        const bool sc = S.get_waypoints(rg.paths[path_of[v]], cfg.dt, veh[v], veh[Bp + v], veh[2 * Bp + v], !rg.track_using_time, rg.target_vel, xr, yr, pr);
        if (lane <= cfg.N) {
            const TppMem mem(tpp_state, tpp_filt, cfg.N, (long)v);
            mem.sto(TF_XR, lane, xr); mem.sto(TF_YR, lane, yr); mem.sto(TF_PR, lane, pr);
        }
        if (sc && lane == 0) stop[v] = 1;
        __syncwarp();
    }
}

// MODEL 0: reference samples already in the slots, stop latch, the first solve from the module-load solution `warm0`.
// MODEL 1 (closed loop of the Frenet node, gazebo_sim_mpc_cmd_pub_frenet.jl:112-153): `fref` = [6][Bp] curvature polynomial (4),
// e_y, psi_start per vehicle from rollout_frenet_ref_kernel; state (0, e_y, -psi_start, v); the first solve from all zeros.
template <int MODEL>
__global__ void __launch_bounds__(MPC_TPP_BLOCK, MPC_TPP_MIN_BLOCKS)
rollout_solve_tpp_kernel(const KCfg cfg, const long long B, const long long Bp, double* veh, const int* stop, double* tpp_state, double* tpp_filt,
                         const double* warm0, const double des_speed, BatchPtrs out, double* log_row, const double* fref, const int first) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = v < B;
    const TppMemT<MODEL> mem(tpp_state, tpp_filt, cfg.N, (long)(valid ? v : 0));
    TppSolverT<MODEL> sv(cfg, mem);
    const bool stopped = valid && !MODEL && stop[v] != 0;
    bool live = valid && !stopped;
    double st4[4] = {0.0, 0.0, 0.0, 0.0};
    if (valid) for (int i = 0; i < 4; i++) st4[i] = veh[i * Bp + v];
    if (live) {
        if (MODEL) {
            for (int i = 0; i < 4; i++) sv.kp[i] = fref[i * Bp + v];
            const double c7[7] = {0.0, fref[4 * Bp + v], -fref[5 * Bp + v], st4[3], veh[10 * Bp + v], veh[11 * Bp + v], des_speed};
            if (first) {   // start = 0.0, then the previous solution
                for (int k = 0; k <= cfg.N; k++) for (int f = TF_SX; f <= TF_UD; f++) mem.sto(f, k, 0.0);
                for (int k = 0; k <= cfg.N; k++) { mem.sto(TF_XR, k, 0.0); mem.sto(TF_YR, k, 0.0); mem.sto(TF_PR, k, 0.0); }
            }
            sv.begin_in_place(c7, nullptr);
        } else {
            const double c7[7] = {st4[0], st4[1], st4[2], st4[3], veh[10 * Bp + v], veh[11 * Bp + v], des_speed};
            sv.begin_in_place(c7, warm0);
        }
    }
    bool solved = false;
    while (__syncthreads_or(live)) {
        const bool done = sv.tick(live);
        if (live && done) { sv.finish(out, (long)v); live = false; solved = true; }
    }
    if (!valid) return;
    double acc_des = -1.0, df_des = 0.0, status = -1.0, iters = 0.0;   // stopped: mpc_cmd_pub.jl:148-153
    if (solved) {
        acc_des = out.u0[2 * v]; df_des = out.u0[2 * v + 1]; status = (double)out.status[v]; iters = (double)out.iters[v];
        veh[10 * Bp + v] = df_des; veh[11 * Bp + v] = acc_des;   // update_current_input(df_opt, a_opt) (:140)
    }
    veh[8 * Bp + v] = acc_des; veh[9 * Bp + v] = df_des;         // published whatever the status (:129-132)
    if (log_row) {
        double* r = log_row + 8 * v;
        r[0] = st4[0]; r[1] = st4[1]; r[2] = st4[2]; r[3] = st4[3]; r[4] = acc_des; r[5] = df_des; r[6] = status; r[7] = iters;
    }
}

// Frenet node, per control period and vehicle (warp = vehicle): frenet_reference (mpc_kernel.cuh; nav_msgs_path_frenet.py:44-86) ->
// curvature polynomial, e_y, psi_start, kept per vehicle for the solve kernel
__global__ void __launch_bounds__(128)
rollout_frenet_ref_kernel(const KCfg cfg, const FrenetRolloutArgs a, const long long Bp, const double* veh, double* fref) {
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int k = threadIdx.x & 31;
    TeamSolver<1, 1> S(cfg, (smem_t)0);   // (one-warp teams: collectives are shuffles only)
    for (long long v = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < a.B; v += warps) {
        double kc[4], ey, psi0;
        frenet_reference(S, a.paths[a.path_of[v]], a, veh[v], veh[Bp + v], veh[2 * Bp + v], kc, ey, psi0);
        if (k < 6) fref[k * Bp + v] = (k < 4) ? sel8(kc, k) : (k == 4 ? ey : psi0);
        __syncwarp();
    }
}

__global__ void rollout_final_kernel(const long long B, const long long Bp, const double* veh, double* final_state) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= B) return;
    for (int i = 0; i < 8; i++) final_state[8 * v + i] = veh[i * Bp + v];
}

#ifndef MPC_ROLLOUT_MIN_BLOCKS
#define MPC_ROLLOUT_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, MPC_ROLLOUT_MIN_BLOCKS)
mpc_rollout_kernel(const KCfg cfg, const RolloutArgs args, unsigned long long* counter) {
    extern __shared__ double smem_all[];
    const int warp = threadIdx.x >> 5;
    const smem_t smem = smem_base(smem_all + (size_t)warp * smem_doubles_per_team(cfg.N));
    const smem_t px = smem_base(smem_all + (size_t)WARPS_PER_BLOCK * smem_doubles_per_team(cfg.N));   // plant exchange area
    __shared__ unsigned long long next_group;
    TeamSolver<1>::init_work(smem, cfg.N);
    for (;;) {   // the block takes WARPS_PER_BLOCK vehicles at a time and steps them together
        if (threadIdx.x == 0) next_group = atomicAdd(counter, (unsigned long long)WARPS_PER_BLOCK);
        __syncthreads();
        const unsigned long long b0 = next_group;
        __syncthreads();
        if (b0 >= (unsigned long long)args.B) break;
        rollout_group<1>(cfg, args, (long)b0, smem, px, WARPS_PER_BLOCK);
    }
}

// Closed-loop rollouts at long horizons (32 <= N <= 95): one vehicle per block of W = 2 or 3 warps
template <int W>
__global__ void __launch_bounds__(W * 32, W == 2 ? 4 : 2)
mpc_rollout_long_kernel(const KCfg cfg, const RolloutArgs args, unsigned long long* counter) {
    extern __shared__ double smem_all[];
    __shared__ unsigned long long next_vehicle;
    const smem_t smem = smem_base(smem_all);
    const smem_t px = smem_base(smem_all + (size_t)smem_doubles_per_team(cfg.N));
    TeamSolver<W>::init_work(smem, cfg.N);
    for (;;) {
        if (threadIdx.x == 0) next_vehicle = atomicAdd(counter, 1ULL);
        __syncthreads();
        const unsigned long long b0 = next_vehicle;
        __syncthreads();
        if (b0 >= (unsigned long long)args.B) break;
        rollout_group<W>(cfg, args, (long)b0, smem, px, 1);
    }
}

// Closed loop on the Frenet-frame module (gazebo_sim_mpc_cmd_pub_frenet.jl:112-153): four vehicles per block
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, 2)
mpc_rollout_frenet_kernel(const KCfg cfg, const FrenetRolloutArgs args, unsigned long long* counter) {
    extern __shared__ double smem_all[];
    const int warp = threadIdx.x >> 5;
    const smem_t smem = smem_base(smem_all + (size_t)warp * smem_doubles_per_team(cfg.N, 1));
    const smem_t px = smem_base(smem_all + (size_t)WARPS_PER_BLOCK * smem_doubles_per_team(cfg.N, 1));
    __shared__ unsigned long long next_group;
    TeamSolver<1, 1>::init_work(smem, cfg.N);
    for (;;) {
        if (threadIdx.x == 0) next_group = atomicAdd(counter, (unsigned long long)WARPS_PER_BLOCK);
        __syncthreads();
        const unsigned long long b0 = next_group;
        __syncthreads();
        if (b0 >= (unsigned long long)args.B) break;
        rollout_group_frenet(cfg, args, (long)b0, smem, px, WARPS_PER_BLOCK);
    }
}

// FP64 FMA throughput probe: 8 independent chains per thread
__global__ void fp64_peak_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct mpcb200_handle {
    mpcb200_config cfg;
    double w[8];
    int device = 0;
    int num_sms = 0;
    int blocks_per_sm = 0;
    int team_warps = 1;     /* warps per problem: 1 (N <= 31), 2 (N <= 63), 3 (N <= 95) */
    int model = 0;          /* 0: XY model; 1: Frenet-frame variant (mpcb200_create_frenet) */
    int frenet_wpb = WARPS_PER_BLOCK;   /* warps (= problems) per block of mpc_solve_frenet_kernel */
    size_t smem_bytes = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    unsigned long long* d_counter = nullptr;
    int* d_roles = nullptr;   /* lane roles of the Riccati recursion for this horizon (riccati_roles) */
    DevBuf d_state, d_ref, d_vdes, d_uprev, d_warm, d_u0, d_cost, d_status, d_iters, d_traj;
    DevBuf d_path[3], d_pose, d_pathof, d_log, d_final, d_stop;
    DevBuf d_tpp_state, d_tpp_filt;   /* thread-per-problem path: slot state [warp][stage][field][lane], filters */
    DevBuf d_genref;                  /* ... on a path: the generated waypoints [B][3][N+1] when the caller did not ask for them */
    DevBuf d_veh, d_vstop;            /* ... closed-loop fleets: vehicle state [12][Bp], stop latches */
    int tpp_blocks_per_sm = 0;        /* resident blocks of mpc_solve_tpp_kernel per SM */
    int tpp_block = MPC_TPP_BLOCK;    /* threads per block of mpc_solve_tpp_kernel (a multiple of 32, <= MPC_TPP_BLOCK) */
    bool tpp_block_forced = false;    /* ... given by MPCB200_TPP_BLOCK: no per-batch choice */
    int64_t tpp_min_batch = 0;        /* batches of at least this many problems take the thread-per-problem path (0: never) */
    bool tpp_default_rule = true;     /* ... as set by the default rule (then warm / rollout starts switch at half of it) */
    DevBuf d_rec, d_resto;        /* packed 32-byte records; restoration count per problem of the last solve */
    DevBuf d_seed;                /* the module-load solution (start point of a rollout's first solve), [6N+4] */
    DevBuf d_fit;                 /* Frenet rollouts: the two least-squares fit matrices [4][n1], [4][n2] */
    double fit_window = -1.0;     /* ... built for this window length */
    int fit_n1 = 0, fit_n2 = 0;
    int frenet_rollout_blocks_per_sm = 0;
    bool seed_valid = false;
    int64_t last_B = 0;           /* batch of the last solve (mpcb200_get_restorations) */
    int32_t small_resto[64];      /* ... its restoration counts when it went through the small-batch path */
    bool resto_on_host = false;
    /* multi-GPU: the handle itself works on devices[0]; sub[i] on devices[i + 1] */
    int n_sub = 0;
    mpcb200_handle* sub[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int64_t slice_lo[8] = {0, 0, 0, 0, 0, 0, 0, 0}, slice_hi[8] = {0, 0, 0, 0, 0, 0, 0, 0};   /* slices of the last sharded call */
    bool last_sharded = false;
    DevBuf d_stage;               /* small batches: one contiguous device buffer, one copy each way */
    void* h_stage = nullptr;      /* its pinned host mirror */
    size_t h_stage_cap = 0;
    /* small batches on the handle's own stream: copy in -> kernel -> copy out captured once per shape as a CUDA graph
     * (one cudaGraphLaunch instead of five driver calls per solve); dropped when a buffer, the weights or the stream change */
    struct SmallGraph { int64_t B; int flags; cudaGraphExec_t exec; };
    SmallGraph graphs[8];
    int n_graphs = 0;
    bool graphs_ok = true;
    int path_n[3] = {0, 0, 0};
    int rollout_blocks_per_sm = 0;
    mpcb200_stats stats;
    char err[512];
};

static thread_local char g_err[512] = "";

/* The dynamic shared-memory limit of a kernel is an attribute of the FUNCTION (per device), not of a launch, and the team
 * footprint depends on the horizon: handles with different horizons in one process share it.  It is therefore only ever
 * raised (a larger limit is fine for a smaller launch); lowering it made the next launch of a longer-horizon handle fail
 * with "invalid argument". */
static cudaError_t raise_dyn_smem(const void* fn, int device, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> current;
    std::lock_guard<std::mutex> g(mu);
    size_t& cur = current[std::make_pair(fn, device)];
    if (cur >= bytes) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}

static int fail(mpcb200_handle* h, int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    char* dst = h ? h->err : g_err;
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(h, expr)                                                                          \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) return fail(h, MPCB200_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

static void drop_graphs(mpcb200_handle* h) {
    for (int i = 0; i < h->n_graphs; i++) if (h->graphs[i].exec) cudaGraphExecDestroy(h->graphs[i].exec);
    h->n_graphs = 0;
}

static int ensure(mpcb200_handle* h, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    if (b.p) CUDA_TRY(h, cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    size_t cap = bytes + bytes / 4;
    CUDA_TRY(h, cudaMalloc(&b.p, cap));
    b.cap = cap;
    return 0;
}

static KCfg make_kcfg(const mpcb200_handle* h) {
    KCfg k;
    memset(&k, 0, sizeof(k));
    const mpcb200_config& c = h->cfg;
    k.N = c.N; k.max_iter = c.max_iter; k.start_mode = c.start_mode;
    k.dt = c.dt; k.dtc = c.dt_control; k.La = c.L_a; k.Lb = c.L_b;
    k.vmin = c.v_min; k.vmax = c.v_max; k.amax = c.a_max; k.smax = c.steer_max;
    k.admax = c.a_dmax; k.sdmax = c.steer_dmax; k.tol = c.tol;
    for (int i = 0; i < 8; i++) k.w[i] = h->w[i];
    kcfg_finalize(k);
    k.roles = h->d_roles;
    return k;
}

extern "C" {

int mpcb200_version(void) { return MPCB200_VERSION; }

int mpcb200_default_config(mpcb200_config* c, int32_t N) {
    if (!c) return MPCB200_EINVAL;
    memset(c, 0, sizeof(*c));
    c->N = N;
    c->max_iter = 200;
    c->start_mode = MPCB200_START_ZERO;
    c->device = 0;
    c->dt = 0.20;          /* MKZMPCPathFollower.jl:33 */
    c->dt_control = 0.10;  /* :28 */
    c->L_a = 1.108;        /* :31 */
    c->L_b = 1.742;        /* :32 */
    c->v_min = 0.0;        /* :47 */
    c->v_max = 20.0;       /* :48 */
    c->a_max = 1.0;        /* :44 */
    c->steer_max = 0.5;    /* :41 */
    c->a_dmax = 1.5;       /* :45 */
    c->steer_dmax = 0.5;   /* :42 */
    c->tol = 1e-8;
    c->n_devices = 0;      /* one GPU: `device` */
    return MPCB200_OK;
}

/* The default rule of mpcb200_set_large_batch_path, from measurements on the B200 (profiles/r02_logs/r02_tpp_thresholds.log,
 * r02_tpp_frenet.log).  XY model from the all-zero start: iteration counts spread from 25 to 200, the streaming layout pays for
 * the tail of long solves: N <= 10 only, from 32,768 problems.  Frenet-frame variant: every solve takes 14-17 iterations, the
 * streaming layout wins at every horizon tried (N = 8: 1.3x / 1.9x / 2.1x at 16 K / 64 K / 256 K problems; N = 20: 0.94x / 1.47x /
 * 1.30x at 16 K / 32 K / 64 K; N = 31, 40: 1.29x at 64 K). */
static int64_t tpp_default_min_batch(int model, int N) {
    if (model) return (N <= 10) ? 16384 : 32768;
    return (N <= 10) ? 32768 : 0;
}

static int create_impl(mpcb200_handle** out, const mpcb200_config* cfg, int model);
static int create_multi(mpcb200_handle** out, const mpcb200_config* cfg, int model);
int mpcb200_create(mpcb200_handle** out, const mpcb200_config* cfg) { return create_multi(out, cfg, 0); }
int mpcb200_create_frenet(mpcb200_handle** out, const mpcb200_config* cfg) { return create_multi(out, cfg, 1); }

/* n_devices > 1: one full single-GPU handle per device; the first one is the handle the caller holds */
static int create_multi(mpcb200_handle** out, const mpcb200_config* cfg, int model) {
    if (!out || !cfg) return fail(nullptr, MPCB200_EINVAL, "mpcb200_create: NULL argument");
    *out = nullptr;
    if (cfg->n_devices < 0 || cfg->n_devices > 8) return fail(nullptr, MPCB200_EINVAL, "mpcb200_create: n_devices=%d outside [0,8]", cfg->n_devices);
    if (cfg->n_devices <= 1) {
        mpcb200_config c1 = *cfg;
        if (cfg->n_devices == 1) c1.device = cfg->devices[0];
        return create_impl(out, &c1, model);
    }
    for (int i = 0; i < cfg->n_devices; i++)
        for (int j = 0; j < i; j++)
            if (cfg->devices[i] == cfg->devices[j]) return fail(nullptr, MPCB200_EINVAL, "mpcb200_create: device %d listed twice", cfg->devices[i]);
    mpcb200_handle* first = nullptr;
    for (int i = 0; i < cfg->n_devices; i++) {
        mpcb200_config ci = *cfg;
        ci.device = cfg->devices[i];
        mpcb200_handle* hi = nullptr;
        const int rc = create_impl(&hi, &ci, model);
        if (rc) { if (first) mpcb200_destroy(first); return rc; }
        if (i == 0) first = hi; else first->sub[first->n_sub++] = hi;
    }
    first->cfg.n_devices = cfg->n_devices;
    *out = first;
    return MPCB200_OK;
}

static int create_impl(mpcb200_handle** out, const mpcb200_config* cfg, int model) {
    if (!out || !cfg) return fail(nullptr, MPCB200_EINVAL, "mpcb200_create: NULL argument");
    *out = nullptr;
    if (cfg->N < 3 || cfg->N > 95) return fail(nullptr, MPCB200_EINVAL, "mpcb200_create: horizon N=%d outside [3,95]", cfg->N);
    if (!(cfg->dt > 0) || !(cfg->dt_control > 0) || !(cfg->L_b > 0) || !(cfg->v_max > cfg->v_min) || !(cfg->a_max > 0) ||
        !(cfg->steer_max > 0 && cfg->steer_max < 1.5) || !(cfg->a_dmax > 0) || !(cfg->steer_dmax > 0) || !(cfg->tol > 0) ||
        cfg->max_iter < 0)
        return fail(nullptr, MPCB200_EINVAL, "mpcb200_create: invalid model constants");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(nullptr, MPCB200_ENODEVICE, "mpcb200_create: no CUDA device (%s); this library has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, MPCB200_EINVAL, "mpcb200_create: device %d of %d", cfg->device, ndev);
    mpcb200_handle* h = new (std::nothrow) mpcb200_handle();
    if (!h) return fail(nullptr, MPCB200_ENOMEM, "mpcb200_create: out of host memory");
    h->cfg = *cfg;
    h->device = cfg->device;
    h->err[0] = 0;
    memset(&h->stats, 0, sizeof(h->stats));
    /* defaults of MKZMPCPathFollower.jl:51-59 in update_cost order */
    const double w0[8] = {9.0, 9.0, 10.0, 0.0, 100.0, 1000.0, 0.0, 0.0};
    /* MKZMPCPathFollowerFrenet.jl:51-58 in the XY slots: nothing on s, C_ey, C_epsi, C_ev, C_dacc, C_ddf, C_acc, C_df */
    const double w1[8] = {0.0, 9.0, 10.0, 0.5, 100.0, 1000.0, 0.0, 0.0};
    memcpy(h->w, model ? w1 : w0, sizeof(w0));
    h->model = model;
#define TRY_OR_FREE(expr)                                                                 \
    do {                                                                                  \
        cudaError_t e2 = (expr);                                                          \
        if (e2 != cudaSuccess) {                                                          \
            fail(nullptr, MPCB200_ECUDA, "%s: %s", #expr, cudaGetErrorString(e2));        \
            mpcb200_destroy(h);   /* releases whatever was created so far */              \
            return MPCB200_ECUDA;                                                         \
        }                                                                                 \
    } while (0)
    TRY_OR_FREE(cudaSetDevice(h->device));
    cudaDeviceProp prop;
    TRY_OR_FREE(cudaGetDeviceProperties(&prop, h->device));
    h->num_sms = prop.multiProcessorCount;
    TRY_OR_FREE(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    TRY_OR_FREE(cudaEventCreate(&h->ev0));
    TRY_OR_FREE(cudaEventCreate(&h->ev1));
    TRY_OR_FREE(cudaMalloc((void**)&h->d_counter, sizeof(unsigned long long)));
    h->team_warps = team_warps(cfg->N);
    {
        int roles[32 * ROLE_STRIDE_F + 32 * RS_STRIDE];   /* shared-memory version, then the register version (MODEL 0) */
        const int rs = role_stride_of(model);
        for (int l = 0; l < 32; l++) riccati_roles(l, cfg->N, w_sd_of(h->team_warps, model), roles + l * rs, model);
        for (int l = 0; l < 32; l++) riccati_roles_shfl(l, cfg->N, w_sd_of(h->team_warps, model), roles + 32 * rs + l * RS_STRIDE);
        TRY_OR_FREE(cudaMalloc((void**)&h->d_roles, sizeof(roles)));
        TRY_OR_FREE(cudaMemcpy(h->d_roles, roles, sizeof(int) * (32 * rs + 32 * RS_STRIDE), cudaMemcpyHostToDevice));
    }
    if (model && h->team_warps > 1) {
        h->smem_bytes = (size_t)smem_doubles_per_team(cfg->N, 1) * sizeof(double);
        const void* fn = (h->team_warps == 2) ? (const void*)mpc_solve_long_kernel<2, 1> : (const void*)mpc_solve_long_kernel<3, 1>;
        TRY_OR_FREE(raise_dyn_smem((const void*)fn, h->device, (size_t)(h->smem_bytes)));
        TRY_OR_FREE(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TRY_OR_FREE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->blocks_per_sm, fn, h->team_warps * 32, h->smem_bytes));
    } else if (model) {
        const size_t team = (size_t)smem_doubles_per_team(cfg->N, 1) * sizeof(double);
        int b4 = 0, b3 = 0;
        TRY_OR_FREE(raise_dyn_smem((const void*)mpc_solve_frenet_kernel<4>, h->device, (size_t)(4 * team)));
        TRY_OR_FREE(cudaFuncSetAttribute(mpc_solve_frenet_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TRY_OR_FREE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b4, mpc_solve_frenet_kernel<4>, 4 * 32, 4 * team));
        TRY_OR_FREE(raise_dyn_smem((const void*)mpc_solve_frenet_kernel<3>, h->device, (size_t)(3 * team)));
        TRY_OR_FREE(cudaFuncSetAttribute(mpc_solve_frenet_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TRY_OR_FREE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b3, mpc_solve_frenet_kernel<3>, 3 * 32, 3 * team));
        /* measured (tools/frenet_bench.py): N = 8: 12 warps at 168 registers 6.06 M solves/s vs 8 warps at 255 registers 5.80 M;
         * N = 20: 9 warps 2.88 M vs 8 warps 3.23 M -- the spill code of the 168-register build pays only for >= 1.4x the warps */
        h->frenet_wpb = (10 * 3 * b3 >= 14 * 4 * b4) ? 3 : 4;
        if (const char* e = getenv("MPCB200_FRENET_WPB")) { int v = atoi(e); if (v == 3 || v == 4) h->frenet_wpb = v; }  /* tuning aid */
        h->blocks_per_sm = (h->frenet_wpb == 3) ? b3 : b4;
        h->smem_bytes = h->frenet_wpb * team;
        const size_t rb = WARPS_PER_BLOCK * team + WARPS_PER_BLOCK * ROLLOUT_PX * sizeof(double);
        TRY_OR_FREE(raise_dyn_smem((const void*)mpc_rollout_frenet_kernel, h->device, (size_t)(rb)));
        TRY_OR_FREE(cudaFuncSetAttribute(mpc_rollout_frenet_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TRY_OR_FREE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->frenet_rollout_blocks_per_sm, mpc_rollout_frenet_kernel, WARPS_PER_BLOCK * 32, rb));
        if (h->frenet_rollout_blocks_per_sm < 1) h->frenet_rollout_blocks_per_sm = 1;
    } else if (h->team_warps == 1) {
        h->smem_bytes = ((size_t)WARPS_PER_BLOCK * smem_doubles_per_team(cfg->N) + WARPS_PER_BLOCK * ROLLOUT_PX) * sizeof(double);
        TRY_OR_FREE(raise_dyn_smem((const void*)mpc_solve_kernel, h->device, (size_t)(h->smem_bytes)));
        TRY_OR_FREE(cudaFuncSetAttribute(mpc_solve_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TRY_OR_FREE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->blocks_per_sm, mpc_solve_kernel, WARPS_PER_BLOCK * 32, h->smem_bytes));
        TRY_OR_FREE(raise_dyn_smem((const void*)mpc_rollout_kernel, h->device, (size_t)(h->smem_bytes)));
        TRY_OR_FREE(cudaFuncSetAttribute(mpc_rollout_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TRY_OR_FREE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->rollout_blocks_per_sm, mpc_rollout_kernel, WARPS_PER_BLOCK * 32, h->smem_bytes));
        if (h->rollout_blocks_per_sm < 1) h->rollout_blocks_per_sm = 1;
    } else {
        h->smem_bytes = (size_t)smem_doubles_per_team(cfg->N) * sizeof(double);
        const void* fn = (h->team_warps == 2) ? (const void*)mpc_solve_long_kernel<2> : (const void*)mpc_solve_long_kernel<3>;
        TRY_OR_FREE(raise_dyn_smem((const void*)fn, h->device, (size_t)(h->smem_bytes)));
        TRY_OR_FREE(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TRY_OR_FREE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->blocks_per_sm, fn, h->team_warps * 32, h->smem_bytes));
        const void* fr = (h->team_warps == 2) ? (const void*)mpc_rollout_long_kernel<2> : (const void*)mpc_rollout_long_kernel<3>;
        const size_t rb = h->smem_bytes + ROLLOUT_PX * sizeof(double);
        TRY_OR_FREE(raise_dyn_smem((const void*)fr, h->device, (size_t)(rb)));
        TRY_OR_FREE(cudaFuncSetAttribute(fr, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TRY_OR_FREE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->rollout_blocks_per_sm, fr, h->team_warps * 32, rb));
        if (h->rollout_blocks_per_sm < 1) h->rollout_blocks_per_sm = 1;
    }
    {
        /* thread-per-problem path for large batches (any horizon): needs B >> resident lanes to keep them busy */
        const void* tk = model ? (const void*)mpc_solve_tpp_kernel<1> : (const void*)mpc_solve_tpp_kernel<0>;
        TRY_OR_FREE(cudaFuncSetAttribute(tk, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1));
        TRY_OR_FREE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->tpp_blocks_per_sm, tk, MPC_TPP_BLOCK, 0));
        if (h->tpp_blocks_per_sm < 1) h->tpp_blocks_per_sm = 1;
        if (const char* e = getenv("MPCB200_TPP_BLOCKS_PER_SM")) { int v = atoi(e); if (v >= 1 && v < h->tpp_blocks_per_sm) h->tpp_blocks_per_sm = v; }  /* tuning aid */
        if (const char* e = getenv("MPCB200_TPP_BLOCK")) { int v = atoi(e); if (v >= 32 && v <= MPC_TPP_BLOCK && v % 32 == 0) { h->tpp_block = v; h->tpp_block_forced = true; } }  /* tuning aid */
        /* measured (tools/tpp_ab.py, profiles/r02_logs/r02_tpp_thresholds.log): N = 8: 0.8x / 1.2x / 1.5x / 1.8x the warp-per-problem
         * kernel at 16 K / 32 K / 64 K / 128 K problems; N = 12, 16: break-even at ~128 K; N = 20: 0.65x at 64 K, 0.97x at 256 K */
        h->tpp_min_batch = tpp_default_min_batch(model, cfg->N);
        if (const char* e = getenv("MPCB200_TPP_MIN_BATCH")) { h->tpp_min_batch = atoll(e); h->tpp_default_rule = false; }   /* tuning aid; 0 switches the path off */
    }
    if (h->blocks_per_sm < 1) h->blocks_per_sm = 1;
    if (const char* e = getenv("MPCB200_BLOCKS_PER_SM")) { int v = atoi(e); if (v >= 1 && v < h->blocks_per_sm) h->blocks_per_sm = v; }  /* tuning aid */
#undef TRY_OR_FREE
    *out = h;
    return MPCB200_OK;
}

int mpcb200_destroy(mpcb200_handle* h) {
    if (!h) return MPCB200_EINVAL;
    for (int i = 0; i < h->n_sub; i++) if (h->sub[i]) mpcb200_destroy(h->sub[i]);
    h->n_sub = 0;
    cudaSetDevice(h->device);
    DevBuf* bufs[] = {&h->d_state, &h->d_ref, &h->d_vdes, &h->d_uprev, &h->d_warm, &h->d_u0, &h->d_cost, &h->d_status, &h->d_iters, &h->d_traj,
                      &h->d_path[0], &h->d_path[1], &h->d_path[2], &h->d_pose, &h->d_pathof, &h->d_log, &h->d_final, &h->d_stop, &h->d_stage,
                      &h->d_rec, &h->d_resto, &h->d_seed, &h->d_fit, &h->d_tpp_state, &h->d_tpp_filt, &h->d_veh, &h->d_vstop, &h->d_genref};
    for (DevBuf* b : bufs) if (b->p) cudaFree(b->p);
    if (h->d_counter) cudaFree(h->d_counter);
    if (h->d_roles) cudaFree(h->d_roles);
    drop_graphs(h);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return MPCB200_OK;
}

int mpcb200_set_cost(mpcb200_handle* h, const double w[8]) {
    if (!h || !w) return fail(h, MPCB200_EINVAL, "mpcb200_set_cost: NULL argument");
    for (int i = 0; i < 8; i++) if (!(w[i] >= 0.0)) return fail(h, MPCB200_EINVAL, "mpcb200_set_cost: weight %d is negative or NaN", i);
    memcpy(h->w, w, 8 * sizeof(double));
    drop_graphs(h);   /* the weights are kernel parameters */
    for (int i = 0; i < h->n_sub; i++) { memcpy(h->sub[i]->w, w, 8 * sizeof(double)); drop_graphs(h->sub[i]); }
    return MPCB200_OK;
}

int mpcb200_set_large_batch_path(mpcb200_handle* h, int64_t min_batch) {
    if (!h) return MPCB200_EINVAL;
    const int64_t v = (min_batch < 0) ? tpp_default_min_batch(h->model, h->cfg.N) : min_batch;
    h->tpp_min_batch = v; h->tpp_default_rule = (min_batch < 0);
    drop_graphs(h);   /* a captured small-batch graph holds the kernel it was captured with */
    for (int i = 0; i < h->n_sub; i++) { h->sub[i]->tpp_min_batch = v; h->sub[i]->tpp_default_rule = (min_batch < 0); drop_graphs(h->sub[i]); }
    return MPCB200_OK;
}

/* the stream of devices[0]; the other devices of a multi-GPU handle keep their own streams */
int mpcb200_set_stream(mpcb200_handle* h, void* s) {
    if (!h) return MPCB200_EINVAL;
    h->stream = s ? (cudaStream_t)s : h->own_stream;
    drop_graphs(h);
    return MPCB200_OK;
}

static int launch_solve(mpcb200_handle* h, int64_t B, const BatchPtrs& io, const RefGen& rg, unsigned long long* zeroed_counter = nullptr) {
    unsigned long long* counter = zeroed_counter ? zeroed_counter : h->d_counter;
    if (!zeroed_counter) CUDA_TRY(h, cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned long long), h->stream));
    /* (never for the packed small-batch path, which brings its own counter and may be inside a graph capture) */
    /* starts next to the solution (warm start, rollout start) take a handful of iterations each, all about the same number:
     * no tail of long solves, so the streaming layout already pays at half the batch (measured at N = 8, 16,384 problems:
     * 1.06x warm, 1.30x from the rollout start; profiles/r02_logs/r02_tpp_warm.log) */
    const int64_t tpp_from = (h->tpp_default_rule && !h->model && (io.warm || h->cfg.start_mode == MPCB200_START_ROLLOUT)) ? h->tpp_min_batch / 2 : h->tpp_min_batch;
    const bool tpp_path_ok = !rg.path_of || (!h->model && h->team_warps == 1);   /* waypoints by one-warp teams: N <= 31 */
    if (tpp_path_ok && !zeroed_counter && h->tpp_min_batch > 0 && B >= tpp_from) {
        /* thread-per-problem: as many slots as lanes can be resident, never more than problems.  The block size is the one that
         * fills the last wave best: when every solve takes about the same number of trips (Frenet variant, warm starts) a batch of
         * 1.73 x the resident lanes leaves a quarter of them idle for the second half of the launch (Frenet N = 20, 65,536
         * problems: 14.2 ms with 224-thread blocks = 1.98 waves, 15.5 ms with 256) */
        int tb = h->tpp_block;
        if (!h->tpp_block_forced) {
            long long best = -1;
            for (int cand = MPC_TPP_BLOCK; cand >= MPC_TPP_BLOCK - 96; cand -= 32) {
                const long long per_wave = (long long)h->num_sms * h->tpp_blocks_per_sm * cand;
                const long long cost = ((B + per_wave - 1) / per_wave) * cand;     /* lane-slots reserved per SM over the launch */
                if (best < 0 || cost < best) { best = cost; tb = cand; }
            }
        }
        long long blocks = (long long)h->num_sms * h->tpp_blocks_per_sm;
        const long long need = (B + tb - 1) / tb;
        if (blocks > need) blocks = need;
        const long long S = blocks * tb;
        int rc;
        if ((rc = ensure(h, h->d_tpp_state, tpp_state_doubles(h->cfg.N, (long)S, h->model) * sizeof(double)))) return rc;
        if ((rc = ensure(h, h->d_tpp_filt, tpp_filter_doubles((long)S) * sizeof(double)))) return rc;
        BatchPtrs io2 = io;
        if (rg.path_of) {
            double* ref = rg.ref_out;
            if (!ref) {
                if ((rc = ensure(h, h->d_genref, (size_t)B * 3 * ((size_t)h->cfg.N + 1) * sizeof(double)))) return rc;
                ref = (double*)h->d_genref.p;
            }
            const long long wblocks = (B + 3) / 4, wmax = (long long)h->num_sms * 16;
            on_path_ref_kernel<<<(int)(wblocks < wmax ? wblocks : wmax), 128, 0, h->stream>>>(make_kcfg(h), (long long)B, io.state, rg, ref);
            CUDA_TRY(h, cudaGetLastError());
            io2.ref = ref;
            h->stats.kernel_launches += 1;
        }
        const double v_des0 = rg.path_of ? rg.target_vel : 0.0;   /* des_speed when the caller gives no v_des, mpc_cmd_pub.jl:58-62,116 */
        if (h->model)
            mpc_solve_tpp_kernel<1><<<(int)blocks, tb, 0, h->stream>>>(make_kcfg(h), io2, (long long)B, (double*)h->d_tpp_state.p,
                                                                      (double*)h->d_tpp_filt.p, counter, v_des0);
        else
            mpc_solve_tpp_kernel<0><<<(int)blocks, tb, 0, h->stream>>>(make_kcfg(h), io2, (long long)B, (double*)h->d_tpp_state.p,
                                                                      (double*)h->d_tpp_filt.p, counter, v_des0);
        CUDA_TRY(h, cudaGetLastError());
        h->stats.kernel_launches += 1;
        if (rg.path_of) {
            on_path_invalid_kernel<<<(int)((B + 127) / 128), 128, 0, h->stream>>>((long long)B, rg, io);
            CUDA_TRY(h, cudaGetLastError());
            h->stats.kernel_launches += 1;
        }
        return 0;
    }
    const int teams_per_block = (h->team_warps != 1) ? 1 : (h->model ? h->frenet_wpb : WARPS_PER_BLOCK);
    long long blocks_needed = (B + teams_per_block - 1) / teams_per_block;
    long long max_blocks = (long long)h->num_sms * h->blocks_per_sm;
    int grid = (int)(blocks_needed < max_blocks ? blocks_needed : max_blocks);
    if (grid < 1) grid = 1;
    if (h->model && h->team_warps == 2)
        mpc_solve_long_kernel<2, 1><<<grid, 64, h->smem_bytes, h->stream>>>(make_kcfg(h), io, rg, (long long)B, counter);
    else if (h->model && h->team_warps == 3)
        mpc_solve_long_kernel<3, 1><<<grid, 96, h->smem_bytes, h->stream>>>(make_kcfg(h), io, rg, (long long)B, counter);
    else if (h->model && h->frenet_wpb == 3)
        mpc_solve_frenet_kernel<3><<<grid, 96, h->smem_bytes, h->stream>>>(make_kcfg(h), io, rg, (long long)B, counter);
    else if (h->model)
        mpc_solve_frenet_kernel<4><<<grid, 128, h->smem_bytes, h->stream>>>(make_kcfg(h), io, rg, (long long)B, counter);
    else if (h->team_warps == 1)
        mpc_solve_kernel<<<grid, WARPS_PER_BLOCK * 32, h->smem_bytes, h->stream>>>(make_kcfg(h), io, rg, (long long)B, counter);
    else if (h->team_warps == 2)
        mpc_solve_long_kernel<2><<<grid, 64, h->smem_bytes, h->stream>>>(make_kcfg(h), io, rg, (long long)B, counter);
    else
        mpc_solve_long_kernel<3><<<grid, 96, h->smem_bytes, h->stream>>>(make_kcfg(h), io, rg, (long long)B, counter);
    CUDA_TRY(h, cudaGetLastError());
    h->stats.kernel_launches += 1;
    return 0;
}

static void fill_paths(const mpcb200_handle* h, PathTable* paths) {
    for (int i = 0; i < 3; i++) {
        const double* base = (const double*)h->d_path[i].p;
        const int n = h->path_n[i];
        memset(&paths[i], 0, sizeof(PathTable));
        paths[i].n = n;
        if (n) { paths[i].t = base; paths[i].X = base + n; paths[i].Y = base + 2 * (size_t)n; paths[i].psi = base + 3 * (size_t)n; paths[i].s = base + 4 * (size_t)n;
                 paths[i].cb = base + 5 * (size_t)n; paths[i].nch = path_chunks(n); }
    }
}

/* the arguments of one solve call (host or device pointers) */
struct SolveArgs {
    const char* who;
    const double* state; const double* ref; const int32_t* path_of;
    int32_t track_using_time; double target_vel;
    const double* v_des; const double* u_prev; double* warm;
    double* u0; double* cost; int32_t* status; int32_t* iters; double* traj;
    double* ref_out; int32_t* stop; double* rec;
};

/* a contiguous slice [lo, lo + n) of a call's problems */
static SolveArgs slice_args(const SolveArgs& a, int64_t lo, int N, int model) {
    const size_t nt = 6 * (size_t)N + 4, nr = model ? 4 : 3 * ((size_t)N + 1);
    SolveArgs s = a;
    s.state = a.state + 4 * lo;
    if (a.ref) s.ref = a.ref + nr * lo;
    if (a.path_of) s.path_of = a.path_of + lo;
    if (a.v_des) s.v_des = a.v_des + lo;
    s.u_prev = a.u_prev + 2 * lo;
    if (a.warm) s.warm = a.warm + nt * lo;
    if (a.u0) s.u0 = a.u0 + 2 * lo;
    if (a.cost) s.cost = a.cost + lo;
    if (a.status) s.status = a.status + lo;
    if (a.iters) s.iters = a.iters + lo;
    if (a.traj) s.traj = a.traj + nt * lo;
    if (a.ref_out) s.ref_out = a.ref_out + 3 * ((size_t)N + 1) * lo;
    if (a.stop) s.stop = a.stop + lo;
    if (a.rec) s.rec = a.rec + 4 * lo;
    return s;
}

/* contiguous slice of `total` owned by part `i` of `parts`; sizes differ by at most one (same rule as sharding.shard_range) */
static void shard_range(int64_t total, int parts, int i, int64_t* lo, int64_t* hi) {
    const int64_t base = total / parts, rem = total % parts;
    *lo = i * base + (i < rem ? i : rem);
    *hi = *lo + base + (i < rem ? 1 : 0);
}

/* Host pointers, small batch (the control loop's batch of one): latency is the driver calls, not the bytes.  All
 * inputs are packed into one pinned buffer (with the zeroed problem counter in front) and go over in ONE copy, all
 * outputs come back in ONE copy: 7 driver calls instead of ~16. */
static const int64_t SMALL_BATCH = 64;
static int solve_small_host(mpcb200_handle* h, int64_t B, const SolveArgs& a) {
    const int N = h->cfg.N;
    const size_t nt = 6 * (size_t)N + 4, nr = h->model ? 4 : 3 * ((size_t)N + 1);
    /* layout in doubles: [counter, pad] state ref uprev vdes | warm | u0 cost traj status/iters(int32 pairs) resto rec */
    const size_t o_state = 2, o_ref = o_state + 4 * B, o_uprev = o_ref + nr * B, o_vdes = o_uprev + 2 * B, o_warm = o_vdes + B;
    const size_t o_u0 = o_warm + nt * B, o_cost = o_u0 + 2 * B, o_traj = o_cost + B, o_stat = o_traj + nt * B, o_iter = o_stat + (B + 1) / 2;
    const size_t o_resto = o_iter + (B + 1) / 2, o_rec = o_resto + (B + 1) / 2;
    const size_t total = o_rec + (a.rec ? 4 * B : 0);
    const size_t bytes = total * sizeof(double);
    int rc;
    if (bytes > h->d_stage.cap || bytes > h->h_stage_cap) drop_graphs(h);   /* the graphs hold the old addresses */
    if ((rc = ensure(h, h->d_stage, bytes))) return rc;
    if (bytes > h->h_stage_cap) {
        if (h->h_stage) CUDA_TRY(h, cudaFreeHost(h->h_stage));
        h->h_stage = nullptr; h->h_stage_cap = 0;
        CUDA_TRY(h, cudaHostAlloc(&h->h_stage, bytes + bytes / 4, cudaHostAllocDefault));
        h->h_stage_cap = bytes + bytes / 4;
    }
    double* hs = (double*)h->h_stage;
    double* ds = (double*)h->d_stage.p;
    hs[0] = 0.0; hs[1] = 0.0;   /* the problem counter (all-zero bits) */
    memcpy(hs + o_state, a.state, 4 * B * sizeof(double));
    memcpy(hs + o_ref, a.ref, nr * B * sizeof(double));
    memcpy(hs + o_uprev, a.u_prev, 2 * B * sizeof(double));
    if (a.v_des) memcpy(hs + o_vdes, a.v_des, B * sizeof(double));
    if (a.warm) memcpy(hs + o_warm, a.warm, nt * B * sizeof(double));
    const size_t in_doubles = a.warm ? o_u0 : o_warm;
    cudaStream_t s = h->stream;
    BatchPtrs io{ds + o_state, ds + o_ref, a.v_des ? ds + o_vdes : nullptr, ds + o_uprev, a.warm ? ds + o_warm : nullptr, ds + o_u0,
                 ds + o_cost, (int*)(ds + o_stat), (int*)(ds + o_iter), a.traj ? ds + o_traj : nullptr,
                 a.rec ? ds + o_rec : nullptr, (int*)(ds + o_resto)};
    RefGen rg;
    memset(&rg, 0, sizeof(rg));
    const size_t out_from = a.warm ? o_warm : o_u0;
    /* the sequence copy in -> kernel -> copy out; `ext`: events recorded from inside a graph need the external flag to be timed */
    auto enqueue = [&](bool ext) -> int {
        CUDA_TRY(h, cudaMemcpyAsync(ds, hs, in_doubles * sizeof(double), cudaMemcpyHostToDevice, s));
        CUDA_TRY(h, cudaEventRecordWithFlags(h->ev0, s, ext ? cudaEventRecordExternal : cudaEventRecordDefault));
        int r = launch_solve(h, B, io, rg, (unsigned long long*)ds);
        if (r) return r;
        CUDA_TRY(h, cudaEventRecordWithFlags(h->ev1, s, ext ? cudaEventRecordExternal : cudaEventRecordDefault));
        CUDA_TRY(h, cudaMemcpyAsync(hs + out_from, ds + out_from, (total - out_from) * sizeof(double), cudaMemcpyDeviceToHost, s));
        return 0;
    };
    bool launched = false;
    if (h->graphs_ok && s == h->own_stream) {
        const int flags = (a.v_des ? 1 : 0) | (a.warm ? 2 : 0) | (a.traj ? 4 : 0) | (a.rec ? 8 : 0);
        cudaGraphExec_t exec = nullptr;
        for (int i = 0; i < h->n_graphs; i++) if (h->graphs[i].B == B && h->graphs[i].flags == flags) exec = h->graphs[i].exec;
        if (!exec) {
            cudaGraph_t graph = nullptr;
            bool ok = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                const int r = enqueue(true);
                ok = (cudaStreamEndCapture(s, &graph) == cudaSuccess) && r == 0 && graph != nullptr;
                h->stats.kernel_launches = 0;   /* counted at the graph launch below */
            }
            if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
            if (ok) {
                if (h->n_graphs == 8) drop_graphs(h);
                h->graphs[h->n_graphs].B = B; h->graphs[h->n_graphs].flags = flags; h->graphs[h->n_graphs].exec = exec;
                h->n_graphs++;
            } else {
                exec = nullptr; h->graphs_ok = false; cudaGetLastError();   /* fall back to plain launches for good */
            }
        }
        if (exec) {
            if (cudaGraphLaunch(exec, s) == cudaSuccess) { launched = true; h->stats.kernel_launches += 1; }
            else { h->graphs_ok = false; cudaGetLastError(); drop_graphs(h); }
        }
    }
    if (!launched && (rc = enqueue(false))) return rc;
    h->stats.h2d_bytes += in_doubles * sizeof(double);
    h->stats.d2h_bytes += (total - out_from) * sizeof(double);
    CUDA_TRY(h, cudaStreamSynchronize(s));
    if (a.u0) memcpy(a.u0, hs + o_u0, 2 * B * sizeof(double));
    if (a.cost) memcpy(a.cost, hs + o_cost, B * sizeof(double));
    if (a.status) memcpy(a.status, hs + o_stat, B * sizeof(int32_t));
    if (a.iters) memcpy(a.iters, hs + o_iter, B * sizeof(int32_t));
    if (a.traj) memcpy(a.traj, hs + o_traj, nt * B * sizeof(double));
    if (a.warm) memcpy(a.warm, hs + o_warm, nt * B * sizeof(double));
    if (a.rec) memcpy(a.rec, hs + o_rec, 4 * B * sizeof(double));
    memcpy(h->small_resto, hs + o_resto, B * sizeof(int32_t));   /* where mpcb200_get_restorations looks after a small batch */
    h->resto_on_host = true;
    float ms = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->stats.kernel_ms = ms;
    return MPCB200_OK;
}

/* HOST pointers, one device: stage the slice through the handle's device buffers and launch (asynchronous) ... */
static int host_enqueue(mpcb200_handle* h, int64_t B, const SolveArgs& a, const RefGen& rg0) {
    CUDA_TRY(h, cudaSetDevice(h->device));
    const bool on_path = (a.path_of != nullptr);
    const int N = h->cfg.N;
    const size_t nt = 6 * (size_t)N + 4, nr = h->model ? 4 : 3 * ((size_t)N + 1);
    const size_t bs = B * 4 * sizeof(double), br = B * nr * sizeof(double), bu = B * 2 * sizeof(double), bt = B * nt * sizeof(double);
    int rc;
    if ((rc = ensure(h, h->d_state, bs))) return rc;
    if ((!on_path || a.ref_out) && (rc = ensure(h, h->d_ref, br))) return rc;
    if (on_path && (rc = ensure(h, h->d_pathof, B * sizeof(int32_t)))) return rc;
    if (on_path && a.stop && (rc = ensure(h, h->d_stop, B * sizeof(int32_t)))) return rc;
    if ((rc = ensure(h, h->d_uprev, bu))) return rc;
    if ((rc = ensure(h, h->d_u0, bu))) return rc;
    if (a.v_des && (rc = ensure(h, h->d_vdes, B * sizeof(double)))) return rc;
    if (a.warm && (rc = ensure(h, h->d_warm, bt))) return rc;
    if (a.cost && (rc = ensure(h, h->d_cost, B * sizeof(double)))) return rc;
    if (a.status && (rc = ensure(h, h->d_status, B * sizeof(int32_t)))) return rc;
    if (a.iters && (rc = ensure(h, h->d_iters, B * sizeof(int32_t)))) return rc;
    if (a.traj && (rc = ensure(h, h->d_traj, bt))) return rc;
    if (a.rec && (rc = ensure(h, h->d_rec, B * 4 * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->d_resto, B * sizeof(int32_t)))) return rc;
    cudaStream_t s = h->stream;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_state.p, a.state, bs, cudaMemcpyHostToDevice, s));
    h->stats.h2d_bytes += bs;
    if (on_path) { CUDA_TRY(h, cudaMemcpyAsync(h->d_pathof.p, a.path_of, B * sizeof(int32_t), cudaMemcpyHostToDevice, s)); h->stats.h2d_bytes += B * sizeof(int32_t); }
    else { CUDA_TRY(h, cudaMemcpyAsync(h->d_ref.p, a.ref, br, cudaMemcpyHostToDevice, s)); h->stats.h2d_bytes += br; }
    CUDA_TRY(h, cudaMemcpyAsync(h->d_uprev.p, a.u_prev, bu, cudaMemcpyHostToDevice, s));
    h->stats.h2d_bytes += bu;
    if (a.v_des) { CUDA_TRY(h, cudaMemcpyAsync(h->d_vdes.p, a.v_des, B * sizeof(double), cudaMemcpyHostToDevice, s)); h->stats.h2d_bytes += B * sizeof(double); }
    if (a.warm) { CUDA_TRY(h, cudaMemcpyAsync(h->d_warm.p, a.warm, bt, cudaMemcpyHostToDevice, s)); h->stats.h2d_bytes += bt; }
    BatchPtrs io{(const double*)h->d_state.p, on_path ? nullptr : (const double*)h->d_ref.p, a.v_des ? (const double*)h->d_vdes.p : nullptr,
                 (const double*)h->d_uprev.p, a.warm ? (double*)h->d_warm.p : nullptr, (double*)h->d_u0.p,
                 a.cost ? (double*)h->d_cost.p : nullptr, a.status ? (int*)h->d_status.p : nullptr,
                 a.iters ? (int*)h->d_iters.p : nullptr, a.traj ? (double*)h->d_traj.p : nullptr,
                 a.rec ? (double*)h->d_rec.p : nullptr, (int*)h->d_resto.p};
    RefGen rg = rg0;
    if (on_path) {
        fill_paths(h, rg.paths);   /* this device's copies of the tables */
        rg.path_of = (const int*)h->d_pathof.p;
        rg.ref_out = a.ref_out ? (double*)h->d_ref.p : nullptr;
        rg.stop = a.stop ? (int*)h->d_stop.p : nullptr;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev0, s));
    if ((rc = launch_solve(h, B, io, rg))) return rc;
    CUDA_TRY(h, cudaEventRecord(h->ev1, s));
    return MPCB200_OK;
}

/* ... and bring the slice's results back into the caller's buffers (synchronises the device's stream) */
static int host_collect(mpcb200_handle* h, int64_t B, const SolveArgs& a) {
    CUDA_TRY(h, cudaSetDevice(h->device));
    const bool on_path = (a.path_of != nullptr);
    const int N = h->cfg.N;
    const size_t nt = 6 * (size_t)N + 4, nr = h->model ? 4 : 3 * ((size_t)N + 1);
    const size_t br = B * nr * sizeof(double), bu = B * 2 * sizeof(double), bt = B * nt * sizeof(double);
    cudaStream_t s = h->stream;
    if (a.u0) { CUDA_TRY(h, cudaMemcpyAsync(a.u0, h->d_u0.p, bu, cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += bu; }
    if (a.cost) { CUDA_TRY(h, cudaMemcpyAsync(a.cost, h->d_cost.p, B * sizeof(double), cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += B * sizeof(double); }
    if (a.status) { CUDA_TRY(h, cudaMemcpyAsync(a.status, h->d_status.p, B * sizeof(int32_t), cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += B * sizeof(int32_t); }
    if (a.iters) { CUDA_TRY(h, cudaMemcpyAsync(a.iters, h->d_iters.p, B * sizeof(int32_t), cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += B * sizeof(int32_t); }
    if (a.traj) { CUDA_TRY(h, cudaMemcpyAsync(a.traj, h->d_traj.p, bt, cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += bt; }
    if (a.warm) { CUDA_TRY(h, cudaMemcpyAsync(a.warm, h->d_warm.p, bt, cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += bt; }
    if (a.rec) { CUDA_TRY(h, cudaMemcpyAsync(a.rec, h->d_rec.p, B * 4 * sizeof(double), cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += B * 4 * sizeof(double); }
    if (on_path && a.ref_out) { CUDA_TRY(h, cudaMemcpyAsync(a.ref_out, h->d_ref.p, br, cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += br; }
    if (on_path && a.stop) { CUDA_TRY(h, cudaMemcpyAsync(a.stop, h->d_stop.p, B * sizeof(int32_t), cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += B * sizeof(int32_t); }
    CUDA_TRY(h, cudaStreamSynchronize(s));
    float ms = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->stats.kernel_ms = ms;
    return MPCB200_OK;
}

/* a sub-handle's error text and counters surface through the handle the caller holds */
static int sub_fail(mpcb200_handle* h, mpcb200_handle* d, int rc) {
    if (d != h) snprintf(h->err, sizeof(h->err), "device %d: %.400s", d->device, d->err);
    return rc;
}

/* common body of mpcb200_solve_batch* (ref given) and mpcb200_solve_batch_on_path (path_of given) */
static int solve_batch_impl(mpcb200_handle* h, int64_t B, const SolveArgs& a, int32_t mem_space) {
    const char* who = a.who;
    if (!h) return MPCB200_EINVAL;
    if (B < 0) return fail(h, MPCB200_EINVAL, "%s: B=%lld", who, (long long)B);
    if (mem_space != MPCB200_HOST && mem_space != MPCB200_DEVICE) return fail(h, MPCB200_EINVAL, "%s: mem_space=%d", who, mem_space);
    memset(&h->stats, 0, sizeof(h->stats));
    h->last_B = B; h->last_sharded = false; h->resto_on_host = false;
    if (B == 0) return MPCB200_OK;
    const bool on_path = (a.path_of != nullptr);
    if (!a.state || (!a.ref && !on_path) || !a.u_prev || (!a.u0 && !a.rec))
        return fail(h, MPCB200_EINVAL, "%s: state, %s, u_prev and %s are required", who, on_path ? "path_of" : "ref", a.rec ? "rec" : "u0");
    if (on_path && h->model) return fail(h, MPCB200_EINVAL, "%s: not available for the Frenet-frame variant", who);
    CUDA_TRY(h, cudaSetDevice(h->device));
    RefGen rg;
    memset(&rg, 0, sizeof(rg));
    if (on_path) {
        fill_paths(h, rg.paths);
        rg.track_using_time = a.track_using_time;
        rg.target_vel = a.target_vel > 0.0 ? a.target_vel : 0.0;   /* des_speed, mpc_cmd_pub.jl:58-62 */
        if (mem_space == MPCB200_HOST) {
            for (int64_t b = 0; b < B; b++)
                if (a.path_of[b] < 0 || a.path_of[b] > 2 || h->path_n[a.path_of[b]] == 0)
                    return fail(h, MPCB200_EINVAL, "%s: problem %lld uses path %d, which was not set with mpcb200_set_path", who, (long long)b, a.path_of[b]);
        }
    }
    if (mem_space == MPCB200_DEVICE) {
        int rc;
        if ((rc = ensure(h, h->d_resto, B * sizeof(int32_t)))) return rc;
        BatchPtrs io{a.state, a.ref, a.v_des, a.u_prev, a.warm, a.u0, a.cost, a.status, a.iters, a.traj, a.rec, (int*)h->d_resto.p};
        rg.path_of = a.path_of; rg.ref_out = a.ref_out; rg.stop = a.stop;
        return launch_solve(h, B, io, rg);
    }
    if (!on_path && B <= SMALL_BATCH) return solve_small_host(h, B, a);
    /* host pointers: one contiguous slice per device, all devices working at once */
    const int parts = (B >= 2 * (int64_t)(h->n_sub + 1)) ? h->n_sub + 1 : 1;
    mpcb200_handle* dev[8];
    dev[0] = h;
    for (int i = 0; i < h->n_sub; i++) dev[i + 1] = h->sub[i];
    int rc;
    for (int i = 0; i < parts; i++) {
        shard_range(B, parts, i, &h->slice_lo[i], &h->slice_hi[i]);
        if (i) memset(&dev[i]->stats, 0, sizeof(mpcb200_stats));
        if ((rc = host_enqueue(dev[i], h->slice_hi[i] - h->slice_lo[i], slice_args(a, h->slice_lo[i], h->cfg.N, h->model), rg))) return sub_fail(h, dev[i], rc);
    }
    for (int i = 0; i < parts; i++)
        if ((rc = host_collect(dev[i], h->slice_hi[i] - h->slice_lo[i], slice_args(a, h->slice_lo[i], h->cfg.N, h->model)))) return sub_fail(h, dev[i], rc);
    for (int i = 1; i < parts; i++) {
        h->stats.kernel_launches += dev[i]->stats.kernel_launches;
        h->stats.h2d_bytes += dev[i]->stats.h2d_bytes; h->stats.d2h_bytes += dev[i]->stats.d2h_bytes;
        if (dev[i]->stats.kernel_ms > h->stats.kernel_ms) h->stats.kernel_ms = dev[i]->stats.kernel_ms;   /* the devices run concurrently */
    }
    h->last_sharded = parts > 1;
    for (int i = parts; i < 8; i++) h->slice_lo[i] = h->slice_hi[i] = 0;
    CUDA_TRY(h, cudaSetDevice(h->device));
    return MPCB200_OK;
}

int mpcb200_get_restorations(mpcb200_handle* h, int64_t B, int32_t* out) {
    if (!h || !out) return fail(h, MPCB200_EINVAL, "mpcb200_get_restorations: NULL argument");
    if (B != h->last_B) return fail(h, MPCB200_EINVAL, "mpcb200_get_restorations: B=%lld, the last solve had %lld problems", (long long)B, (long long)h->last_B);
    if (B == 0) return MPCB200_OK;
    if (h->resto_on_host) { memcpy(out, h->small_resto, B * sizeof(int32_t)); return MPCB200_OK; }
    if (!h->last_sharded) {
        CUDA_TRY(h, cudaSetDevice(h->device));
        CUDA_TRY(h, cudaMemcpyAsync(out, h->d_resto.p, B * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        return MPCB200_OK;
    }
    for (int i = 0; i <= h->n_sub; i++) {
        mpcb200_handle* d = i ? h->sub[i - 1] : h;
        const int64_t n = h->slice_hi[i] - h->slice_lo[i];
        if (n <= 0) continue;
        CUDA_TRY(h, cudaSetDevice(d->device));
        CUDA_TRY(h, cudaMemcpyAsync(out + h->slice_lo[i], d->d_resto.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, d->stream));
        CUDA_TRY(h, cudaStreamSynchronize(d->stream));
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    return MPCB200_OK;
}

int mpcb200_solve_batch_frenet(mpcb200_handle* h, int64_t B, const double* state, const double* k_coeffs, const double* v_des,
                               const double* u_prev, double* warm, double* u0, double* cost, int32_t* status, int32_t* iters,
                               double* traj, int32_t mem_space) {
    if (h && !h->model) return fail(h, MPCB200_EINVAL, "mpcb200_solve_batch_frenet: the handle was not created with mpcb200_create_frenet");
    const SolveArgs a{"mpcb200_solve_batch_frenet", state, k_coeffs, nullptr, 1, 0.0, v_des, u_prev, warm, u0, cost, status, iters, traj,
                      nullptr, nullptr, nullptr};
    return solve_batch_impl(h, B, a, mem_space);
}

int mpcb200_set_cost_frenet(mpcb200_handle* h, const double w[7]) {
    if (!h || !w) return fail(h, MPCB200_EINVAL, "mpcb200_set_cost_frenet: NULL argument");
    if (!h->model) return fail(h, MPCB200_EINVAL, "mpcb200_set_cost_frenet: the handle was not created with mpcb200_create_frenet");
    for (int i = 0; i < 7; i++) if (!(w[i] >= 0.0)) return fail(h, MPCB200_EINVAL, "mpcb200_set_cost_frenet: weight %d is negative or NaN", i);
    h->w[0] = 0.0;
    memcpy(h->w + 1, w, 7 * sizeof(double));
    drop_graphs(h);
    for (int i = 0; i < h->n_sub; i++) { memcpy(h->sub[i]->w, h->w, 8 * sizeof(double)); drop_graphs(h->sub[i]); }
    return MPCB200_OK;
}

int mpcb200_solve_batch(mpcb200_handle* h, int64_t B, const double* state, const double* ref, const double* v_des,
                        const double* u_prev, double* warm, double* u0, double* cost, int32_t* status, int32_t* iters,
                        double* traj, int32_t mem_space) {
    if (h && h->model) return fail(h, MPCB200_EINVAL, "mpcb200_solve_batch: Frenet handle; use mpcb200_solve_batch_frenet");
    const SolveArgs a{"mpcb200_solve_batch", state, ref, nullptr, 1, 0.0, v_des, u_prev, warm, u0, cost, status, iters, traj,
                      nullptr, nullptr, nullptr};
    return solve_batch_impl(h, B, a, mem_space);
}

int mpcb200_solve_batch_records(mpcb200_handle* h, int64_t B, const double* state, const double* ref, const double* v_des,
                                const double* u_prev, double* warm, double* rec, double* traj, int32_t mem_space) {
    if (h && !rec) return fail(h, MPCB200_EINVAL, "mpcb200_solve_batch_records: rec is required");
    /* u0 is part of the record; the kernel still wants a place for it in DEVICE mode: none is needed, it skips NULL */
    const SolveArgs a{"mpcb200_solve_batch_records", state, ref, nullptr, 1, 0.0, v_des, u_prev, warm, nullptr, nullptr, nullptr, nullptr, traj,
                      nullptr, nullptr, rec};
    return solve_batch_impl(h, B, a, mem_space);
}

int mpcb200_solve_batch_on_path(mpcb200_handle* h, int64_t B, const double* state, const int32_t* path_of, int32_t track_using_time,
                                double target_vel, const double* v_des, const double* u_prev, double* warm, double* u0, double* cost,
                                int32_t* status, int32_t* iters, double* traj, double* ref_out, int32_t* stop, int32_t mem_space) {
    if (h && !path_of) return fail(h, MPCB200_EINVAL, "mpcb200_solve_batch_on_path: path_of is required");
    const SolveArgs a{"mpcb200_solve_batch_on_path", state, nullptr, path_of, track_using_time, target_vel, v_des, u_prev, warm, u0, cost,
                      status, iters, traj, ref_out, stop, nullptr};
    return solve_batch_impl(h, B, a, mem_space);
}

int mpcb200_set_path(mpcb200_handle* h, int32_t path_id, int32_t n, const double* t, const double* X, const double* Y,
                     const double* psi, const double* s) {
    if (!h) return MPCB200_EINVAL;
    if (path_id < 0 || path_id > 2) return fail(h, MPCB200_EINVAL, "mpcb200_set_path: path_id %d outside [0,2]", path_id);
    if (n < 2 || !t || !X || !Y || !psi || !s) return fail(h, MPCB200_EINVAL, "mpcb200_set_path: need n >= 2 samples and five columns");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int nch = path_chunks(n);
    int rc = ensure(h, h->d_path[path_id], ((size_t)5 * n + (size_t)3 * nch) * sizeof(double));
    if (rc) return rc;
    const double* cols[5] = {t, X, Y, psi, s};
    for (int i = 0; i < 5; i++)
        CUDA_TRY(h, cudaMemcpyAsync((double*)h->d_path[path_id].p + (size_t)i * n, cols[i], (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    /* bounding circles of the chunks of samples the nearest-sample search skips (TeamSolver::nearest_sample) */
    double* cb = (double*)malloc((size_t)3 * nch * sizeof(double));
    if (!cb) return fail(h, MPCB200_ENOMEM, "mpcb200_set_path: out of host memory");
    path_chunk_bounds(n, X, Y, cb);
    const cudaError_t ce = cudaMemcpyAsync((double*)h->d_path[path_id].p + (size_t)5 * n, cb, (size_t)3 * nch * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    const cudaError_t cs = cudaStreamSynchronize(h->stream);
    free(cb);
    CUDA_TRY(h, ce);
    CUDA_TRY(h, cs);
    h->path_n[path_id] = n;
    for (int i = 0; i < h->n_sub; i++)   /* the tables are replicated on every device */
        if ((rc = mpcb200_set_path(h->sub[i], path_id, n, t, X, Y, psi, s))) return sub_fail(h, h->sub[i], rc);
    if (h->n_sub) CUDA_TRY(h, cudaSetDevice(h->device));
    return MPCB200_OK;
}

/* The start point of a vehicle's FIRST solve: the solution of the module-load solve of the default problem
 * (MKZMPCPathFollower.jl:36-39,75,82,110-113,126-128: zero state and previous command, x_ref = 15 t, y_ref = psi_ref = 0,
 * module-default weights, start = 0.0), which JuMP re-solves from.  Solved once per handle with this library. */
static int ensure_seed(mpcb200_handle* h) {
    if (h->seed_valid) return MPCB200_OK;
    const int N = h->cfg.N;
    const size_t nt = 6 * (size_t)N + 4;
    double* buf = (double*)calloc(4 + 3 * ((size_t)N + 1) + 2 + 1 + 2 + nt, sizeof(double));
    if (!buf) return fail(h, MPCB200_ENOMEM, "mpcb200_rollout: out of host memory");
    double *state = buf, *ref = state + 4, *u_prev = ref + 3 * (N + 1), *v_des = u_prev + 2, *u0 = v_des + 1, *traj = u0 + 2;
    for (int k = 0; k <= N; k++) ref[k] = 15.0 * ((double)k * h->cfg.dt);
    v_des[0] = 15.0;
    double w_keep[8];
    const double w0[8] = {9.0, 9.0, 10.0, 0.0, 100.0, 1000.0, 0.0, 0.0};
    memcpy(w_keep, h->w, sizeof(w_keep));
    memcpy(h->w, w0, sizeof(w0));
    const int start_keep = h->cfg.start_mode;
    h->cfg.start_mode = MPCB200_START_ZERO;
    drop_graphs(h);   /* weights and start mode are kernel parameters */
    const mpcb200_stats keep = h->stats;
    const SolveArgs a{"mpcb200_rollout (module-load solve)", state, ref, nullptr, 1, 0.0, v_des, u_prev, nullptr, u0, nullptr, nullptr, nullptr,
                      traj, nullptr, nullptr, nullptr};
    int rc = solve_small_host(h, 1, a);
    memcpy(h->w, w_keep, sizeof(w_keep));
    h->cfg.start_mode = start_keep;
    drop_graphs(h);
    h->stats = keep;
    if (!rc) rc = ensure(h, h->d_seed, nt * sizeof(double));
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(h->d_seed.p, traj, nt * sizeof(double), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail(h, MPCB200_ECUDA, "mpcb200_rollout: %s", cudaGetErrorString(e));
    }
    free(buf);
    if (!rc) h->seed_valid = true;
    return rc;
}

/* one device's share of a rollout: vehicles [lo, lo + n) of B; enqueue ... */
static int rollout_enqueue(mpcb200_handle* h, int64_t n, int32_t T, const double* pose0, const int32_t* path_of, int32_t track_using_time,
                           double target_vel, bool want_log, bool want_final) {
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc;
    if ((rc = ensure_seed(h))) return rc;
    const size_t bl = (size_t)T * n * 8 * sizeof(double), bf = (size_t)n * 8 * sizeof(double);
    if ((rc = ensure(h, h->d_pose, (size_t)n * 3 * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->d_pathof, (size_t)n * sizeof(int32_t)))) return rc;
    if (want_log && (rc = ensure(h, h->d_log, bl))) return rc;
    if (want_final && (rc = ensure(h, h->d_final, bf))) return rc;
    cudaStream_t s = h->stream;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_pose.p, pose0, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_pathof.p, path_of, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    h->stats.h2d_bytes += (size_t)n * (3 * sizeof(double) + sizeof(int32_t));
    RolloutArgs a;
    memset(&a, 0, sizeof(a));
    a.pose0 = (const double*)h->d_pose.p; a.path_of = (const int*)h->d_pathof.p;
    fill_paths(h, a.paths);
    a.T = T; a.track_using_time = track_using_time; a.target_vel = target_vel;
    a.log = want_log ? (double*)h->d_log.p : nullptr; a.final_state = want_final ? (double*)h->d_final.p : nullptr; a.B = (long)n;
    a.warm0 = (const double*)h->d_seed.p;
    /* large fleets at short horizons: one control period = plant / waypoints / thread-per-problem solve over the whole fleet.
     * Measured at N = 8, 200 periods (profiles/r02_logs/r02_rollout_thresholds.log): a period of the pipeline takes ~0.9 ms up to
     * 8,192 vehicles (the slowest vehicle's dozen solver trips) and grows slowly after that, the fused kernel takes 0.5 / 0.8 /
     * 1.4 / 2.5 ms at 2,048 / 4,096 / 8,192 / 16,384 vehicles: the default rule switches at a quarter of the cold-start batch */
    const int64_t roll_from = h->tpp_default_rule ? h->tpp_min_batch / 4 : h->tpp_min_batch;
    if (h->tpp_min_batch > 0 && n >= roll_from) {
        const long long Bp = (n + 31) / 32 * 32;
        /* blocks of 128 vehicles (two resident per SM): every vehicle solves at once and a block keeps its vehicles, so smaller
         * blocks mean less idle SMs in the last wave (16,384 vehicles: 0.89 instead of 1.07 ms per period; never slower) */
        const int tb = h->tpp_block < 128 ? h->tpp_block : 128;
        const long long sblocks = (n + tb - 1) / tb, S = sblocks * tb;
        if ((rc = ensure(h, h->d_tpp_state, tpp_state_doubles(h->cfg.N, (long)S) * sizeof(double)))) return rc;
        if ((rc = ensure(h, h->d_tpp_filt, tpp_filter_doubles((long)S) * sizeof(double)))) return rc;
        if ((rc = ensure(h, h->d_veh, (size_t)RV_NF * Bp * sizeof(double)))) return rc;
        if ((rc = ensure(h, h->d_vstop, (size_t)Bp * sizeof(int)))) return rc;
        if ((rc = ensure(h, h->d_u0, (size_t)n * 2 * sizeof(double)))) return rc;
        if ((rc = ensure(h, h->d_status, (size_t)n * sizeof(int32_t)))) return rc;
        if ((rc = ensure(h, h->d_iters, (size_t)n * sizeof(int32_t)))) return rc;
        BatchPtrs out;
        memset(&out, 0, sizeof(out));
        out.u0 = (double*)h->d_u0.p; out.status = (int*)h->d_status.p; out.iters = (int*)h->d_iters.p;
        RefGen rg;
        memset(&rg, 0, sizeof(rg));
        fill_paths(h, rg.paths);
        rg.track_using_time = track_using_time;
        const double des_speed = target_vel > 0.0 ? target_vel : 0.0;   /* mpc_cmd_pub.jl:58-62 */
        rg.target_vel = des_speed;
        const KCfg kc = make_kcfg(h);
        double* veh = (double*)h->d_veh.p;
        const int pgrid = (int)((n + 127) / 128);
        long long wblocks = (n + 3) / 4, wmax = (long long)h->num_sms * 16;
        const int wgrid = (int)(wblocks < wmax ? wblocks : wmax);
        CUDA_TRY(h, cudaEventRecord(h->ev0, s));
        for (int t = 0; t < T; t++) {
            rollout_plant_kernel<<<pgrid, 128, 0, s>>>((long long)n, Bp, veh, a.pose0, t == 0);
            rollout_waypoints_kernel<<<wgrid, 128, 0, s>>>(kc, (long long)n, Bp, veh, a.path_of, rg, (double*)h->d_tpp_state.p, (double*)h->d_tpp_filt.p,
                                                          (int*)h->d_vstop.p, t == 0);
            rollout_solve_tpp_kernel<0><<<(int)sblocks, tb, 0, s>>>(kc, (long long)n, Bp, veh, (const int*)h->d_vstop.p, (double*)h->d_tpp_state.p,
                                                                   (double*)h->d_tpp_filt.p, t == 0 ? a.warm0 : nullptr, des_speed, out,
                                                                   a.log ? a.log + (size_t)t * n * 8 : nullptr, nullptr, t == 0);
            CUDA_TRY(h, cudaGetLastError());
        }
        if (a.final_state) rollout_final_kernel<<<pgrid, 128, 0, s>>>((long long)n, Bp, veh, a.final_state);
        CUDA_TRY(h, cudaGetLastError());
        CUDA_TRY(h, cudaEventRecord(h->ev1, s));
        h->stats.kernel_launches += 3 * (int64_t)T + (a.final_state ? 1 : 0);
        return MPCB200_OK;
    }
    CUDA_TRY(h, cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned long long), s));
    const int per_block = (h->team_warps == 1) ? WARPS_PER_BLOCK : 1;
    long long blocks_needed = (n + per_block - 1) / per_block;
    long long max_blocks = (long long)h->num_sms * h->rollout_blocks_per_sm;
    int grid = (int)(blocks_needed < max_blocks ? blocks_needed : max_blocks);
    CUDA_TRY(h, cudaEventRecord(h->ev0, s));
    if (h->team_warps == 1)
        mpc_rollout_kernel<<<grid, WARPS_PER_BLOCK * 32, h->smem_bytes, s>>>(make_kcfg(h), a, h->d_counter);
    else if (h->team_warps == 2)
        mpc_rollout_long_kernel<2><<<grid, 64, h->smem_bytes + ROLLOUT_PX * sizeof(double), s>>>(make_kcfg(h), a, h->d_counter);
    else
        mpc_rollout_long_kernel<3><<<grid, 96, h->smem_bytes + ROLLOUT_PX * sizeof(double), s>>>(make_kcfg(h), a, h->d_counter);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaEventRecord(h->ev1, s));
    h->stats.kernel_launches += 1;
    return MPCB200_OK;
}

/* ... and copy the share's rows of log [T][B][8] (a strided block when the fleet is sharded) and final [B][8] back */
static int rollout_collect(mpcb200_handle* h, int64_t lo, int64_t n, int64_t B, int32_t T, double* log, double* final_state) {
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const size_t row = (size_t)n * 8 * sizeof(double);
    if (log) {
        CUDA_TRY(h, cudaMemcpy2DAsync(log + (size_t)lo * 8, (size_t)B * 8 * sizeof(double), h->d_log.p, row, row, (size_t)T, cudaMemcpyDeviceToHost, s));
        h->stats.d2h_bytes += row * T;
    }
    if (final_state) { CUDA_TRY(h, cudaMemcpyAsync(final_state + (size_t)lo * 8, h->d_final.p, row, cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += row; }
    CUDA_TRY(h, cudaStreamSynchronize(s));
    float ms = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->stats.kernel_ms = ms;
    return MPCB200_OK;
}

int mpcb200_rollout(mpcb200_handle* h, int64_t B, int32_t T, const double* pose0, const int32_t* path_of, int32_t track_using_time,
                    double target_vel, double* log, double* final_state) {
    if (!h) return MPCB200_EINVAL;
    if (B < 0 || T < 0) return fail(h, MPCB200_EINVAL, "mpcb200_rollout: B=%lld T=%d", (long long)B, T);
    if (h->model) return fail(h, MPCB200_EINVAL, "mpcb200_rollout: not available for the Frenet-frame variant");
    memset(&h->stats, 0, sizeof(h->stats));
    if (B == 0 || T == 0) return MPCB200_OK;
    if (!pose0 || !path_of) return fail(h, MPCB200_EINVAL, "mpcb200_rollout: pose0 and path_of are required");
    for (int64_t b = 0; b < B; b++)
        if (path_of[b] < 0 || path_of[b] > 2 || h->path_n[path_of[b]] == 0)
            return fail(h, MPCB200_EINVAL, "mpcb200_rollout: vehicle %lld uses path %d, which was not set with mpcb200_set_path", (long long)b, path_of[b]);
    /* a vehicle stays on one GPU for all T steps: contiguous slices of the fleet, every device at once */
    const int parts = (B >= 2 * (int64_t)(h->n_sub + 1)) ? h->n_sub + 1 : 1;
    mpcb200_handle* dev[8];
    dev[0] = h;
    for (int i = 0; i < h->n_sub; i++) dev[i + 1] = h->sub[i];
    int64_t lo[8], hi[8];
    int rc;
    for (int i = 0; i < parts; i++) {
        shard_range(B, parts, i, &lo[i], &hi[i]);
        if (i) memset(&dev[i]->stats, 0, sizeof(mpcb200_stats));
        if ((rc = rollout_enqueue(dev[i], hi[i] - lo[i], T, pose0 + 3 * lo[i], path_of + lo[i], track_using_time, target_vel, log != nullptr,
                                  final_state != nullptr)))
            return sub_fail(h, dev[i], rc);
    }
    for (int i = 0; i < parts; i++)
        if ((rc = rollout_collect(dev[i], lo[i], hi[i] - lo[i], B, T, log, final_state))) return sub_fail(h, dev[i], rc);
    for (int i = 1; i < parts; i++) {
        h->stats.kernel_launches += dev[i]->stats.kernel_launches;
        h->stats.h2d_bytes += dev[i]->stats.h2d_bytes; h->stats.d2h_bytes += dev[i]->stats.d2h_bytes;
        if (dev[i]->stats.kernel_ms > h->stats.kernel_ms) h->stats.kernel_ms = dev[i]->stats.kernel_ms;
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    return MPCB200_OK;
}

int mpcb200_rollout_frenet(mpcb200_handle* h, int64_t B, int32_t T, const double* pose0, const int32_t* path_of, double window,
                           double target_vel, int32_t ey_from_path, double* log, double* final_state) {
    if (!h) return MPCB200_EINVAL;
    if (B < 0 || T < 0) return fail(h, MPCB200_EINVAL, "mpcb200_rollout_frenet: B=%lld T=%d", (long long)B, T);
    if (!h->model) return fail(h, MPCB200_EINVAL, "mpcb200_rollout_frenet: the handle was not created with mpcb200_create_frenet");
    if (h->team_warps != 1) return fail(h, MPCB200_EINVAL, "mpcb200_rollout_frenet: N <= 31 (N=%d)", h->cfg.N);
    if (!(window >= 4.0 && window <= 500.0)) return fail(h, MPCB200_EINVAL, "mpcb200_rollout_frenet: window %.3g m outside [4, 500]", window);
    memset(&h->stats, 0, sizeof(h->stats));
    if (B == 0 || T == 0) return MPCB200_OK;
    if (!pose0 || !path_of) return fail(h, MPCB200_EINVAL, "mpcb200_rollout_frenet: pose0 and path_of are required");
    for (int64_t b = 0; b < B; b++)
        if (path_of[b] < 0 || path_of[b] > 2 || h->path_n[path_of[b]] == 0)
            return fail(h, MPCB200_EINVAL, "mpcb200_rollout_frenet: vehicle %lld uses path %d, which was not set with mpcb200_set_path", (long long)b, path_of[b]);
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    int rc;
    if (h->fit_window != window) {   /* grids of nav_msgs_path_frenet.py:63,79: arange(0, window, 0.5) and arange(0, window, 0.25) */
        const int n1 = (int)ceil(window / 0.5 - 1e-9), n2 = (int)ceil(window / 0.25 - 1e-9);
        double* P = (double*)malloc(sizeof(double) * 4 * ((size_t)n1 + n2));
        if (!P) return fail(h, MPCB200_ENOMEM, "mpcb200_rollout_frenet: out of host memory");
        cubic_fit_matrix(n1, 0.5, P); cubic_fit_matrix(n2, 0.25, P + 4 * (size_t)n1);
        rc = ensure(h, h->d_fit, sizeof(double) * 4 * ((size_t)n1 + n2));
        if (!rc) {
            cudaError_t e = cudaMemcpyAsync(h->d_fit.p, P, sizeof(double) * 4 * ((size_t)n1 + n2), cudaMemcpyHostToDevice, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) rc = fail(h, MPCB200_ECUDA, "mpcb200_rollout_frenet: %s", cudaGetErrorString(e));
        }
        free(P);
        if (rc) return rc;
        h->fit_window = window; h->fit_n1 = n1; h->fit_n2 = n2;
    }
    const size_t bl = (size_t)T * B * 8 * sizeof(double), bf = (size_t)B * 8 * sizeof(double);
    if ((rc = ensure(h, h->d_pose, (size_t)B * 3 * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->d_pathof, (size_t)B * sizeof(int32_t)))) return rc;
    if (log && (rc = ensure(h, h->d_log, bl))) return rc;
    if (final_state && (rc = ensure(h, h->d_final, bf))) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_pose.p, pose0, (size_t)B * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_pathof.p, path_of, (size_t)B * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    h->stats.h2d_bytes += (size_t)B * (3 * sizeof(double) + sizeof(int32_t));
    FrenetRolloutArgs a;
    memset(&a, 0, sizeof(a));
    a.pose0 = (const double*)h->d_pose.p; a.path_of = (const int*)h->d_pathof.p;
    fill_paths(h, a.paths);
    a.T = T; a.ey_from_path = ey_from_path; a.target_vel = target_vel;
    a.log = log ? (double*)h->d_log.p : nullptr; a.final_state = final_state ? (double*)h->d_final.p : nullptr; a.B = (long)B;
    a.P1 = (const double*)h->d_fit.p; a.P2 = a.P1 + 4 * (size_t)h->fit_n1; a.n1 = h->fit_n1; a.n2 = h->fit_n2;
    /* large fleets: one control period = plant / Frenet reference / thread-per-problem solve over the whole fleet (as mpcb200_rollout
     * does for the XY model); the default switch is a quarter of the batch rule */
    const int64_t roll_from = h->tpp_default_rule ? h->tpp_min_batch / 4 : h->tpp_min_batch;
    if (h->tpp_min_batch > 0 && B >= roll_from) {
        const long long Bp = (B + 31) / 32 * 32;
        const int tb = h->tpp_block < 128 ? h->tpp_block : 128;
        const long long sblocks = (B + tb - 1) / tb, S = sblocks * tb;
        if ((rc = ensure(h, h->d_tpp_state, tpp_state_doubles(h->cfg.N, (long)S, 1) * sizeof(double)))) return rc;
        if ((rc = ensure(h, h->d_tpp_filt, tpp_filter_doubles((long)S) * sizeof(double)))) return rc;
        if ((rc = ensure(h, h->d_veh, (size_t)(RV_NF + 6) * Bp * sizeof(double)))) return rc;
        if ((rc = ensure(h, h->d_u0, (size_t)B * 2 * sizeof(double)))) return rc;
        if ((rc = ensure(h, h->d_status, (size_t)B * sizeof(int32_t)))) return rc;
        if ((rc = ensure(h, h->d_iters, (size_t)B * sizeof(int32_t)))) return rc;
        BatchPtrs out;
        memset(&out, 0, sizeof(out));
        out.u0 = (double*)h->d_u0.p; out.status = (int*)h->d_status.p; out.iters = (int*)h->d_iters.p;
        const KCfg kc = make_kcfg(h);
        double* veh = (double*)h->d_veh.p;
        double* fref = veh + (size_t)RV_NF * Bp;
        const int pgrid = (int)((B + 127) / 128);
        long long wblocks = (B + 3) / 4, wmax = (long long)h->num_sms * 16;
        const int wgrid = (int)(wblocks < wmax ? wblocks : wmax);
        CUDA_TRY(h, cudaEventRecord(h->ev0, s));
        for (int t = 0; t < T; t++) {
            rollout_plant_kernel<<<pgrid, 128, 0, s>>>((long long)B, Bp, veh, a.pose0, t == 0);
            rollout_frenet_ref_kernel<<<wgrid, 128, 0, s>>>(kc, a, Bp, veh, fref);
            rollout_solve_tpp_kernel<1><<<(int)sblocks, tb, 0, s>>>(kc, (long long)B, Bp, veh, nullptr, (double*)h->d_tpp_state.p, (double*)h->d_tpp_filt.p,
                                                                   nullptr, target_vel, out, a.log ? a.log + (size_t)t * B * 8 : nullptr, fref, t == 0);
            CUDA_TRY(h, cudaGetLastError());
        }
        if (a.final_state) rollout_final_kernel<<<pgrid, 128, 0, s>>>((long long)B, Bp, veh, a.final_state);
        CUDA_TRY(h, cudaGetLastError());
        CUDA_TRY(h, cudaEventRecord(h->ev1, s));
        h->stats.kernel_launches += 3 * (int64_t)T + (a.final_state ? 1 : 0);
        if (log) { CUDA_TRY(h, cudaMemcpyAsync(log, h->d_log.p, bl, cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += bl; }
        if (final_state) { CUDA_TRY(h, cudaMemcpyAsync(final_state, h->d_final.p, bf, cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += bf; }
        CUDA_TRY(h, cudaStreamSynchronize(s));
        float ms2 = 0.f;
        CUDA_TRY(h, cudaEventElapsedTime(&ms2, h->ev0, h->ev1));
        h->stats.kernel_ms = ms2;
        return MPCB200_OK;
    }
    CUDA_TRY(h, cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned long long), s));
    long long blocks_needed = (B + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    long long max_blocks = (long long)h->num_sms * h->frenet_rollout_blocks_per_sm;
    int grid = (int)(blocks_needed < max_blocks ? blocks_needed : max_blocks);
    const size_t smem = (size_t)WARPS_PER_BLOCK * smem_doubles_per_team(h->cfg.N, 1) * sizeof(double) + WARPS_PER_BLOCK * ROLLOUT_PX * sizeof(double);
    CUDA_TRY(h, cudaEventRecord(h->ev0, s));
    mpc_rollout_frenet_kernel<<<grid, WARPS_PER_BLOCK * 32, smem, s>>>(make_kcfg(h), a, h->d_counter);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaEventRecord(h->ev1, s));
    h->stats.kernel_launches += 1;
    if (log) { CUDA_TRY(h, cudaMemcpyAsync(log, h->d_log.p, bl, cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += bl; }
    if (final_state) { CUDA_TRY(h, cudaMemcpyAsync(final_state, h->d_final.p, bf, cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += bf; }
    CUDA_TRY(h, cudaStreamSynchronize(s));
    float ms = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->stats.kernel_ms = ms;
    return MPCB200_OK;
}

int mpcb200_get_stats(mpcb200_handle* h, mpcb200_stats* out) {
    if (!h || !out) return MPCB200_EINVAL;
    *out = h->stats;
    return MPCB200_OK;
}

int mpcb200_fp64_peak(mpcb200_handle* h, double* tflops) {
    if (!h || !tflops) return MPCB200_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int threads = 256, blocks = h->num_sms * 8, iters = 1 << 15;
    double* d = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&d, (size_t)threads * blocks * sizeof(double)));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
        fp64_peak_kernel<<<blocks, threads, 0, h->stream>>>(d, iters);
        CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        float ms = 0.f;
        CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaFree(d);
    const double flops = 2.0 * 8.0 * (double)iters * threads * blocks;
    *tflops = flops / (best * 1e-3) / 1e12;
    return MPCB200_OK;
}

const char* mpcb200_last_error(mpcb200_handle* h) { return h ? h->err : g_err; }

}  // extern "C"
