// mpc_b200.cu -- C ABI (include/mpc_b200.h) and kernel launches for sm_100a.
//
// One handle = one GPU + one stream.  Problems are independent: a persistent grid of warps
// pulls problem indices from a device counter (iteration counts differ per problem, so a
// dynamic queue keeps the SMs level), solves each with mpc_kernel.cuh::solve_problem and
// writes the per-problem record.  No CPU path exists in this library.
#include "../../include/mpc_b200.h"
#include "mpc_kernel.cuh"

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

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
        rollout_group(cfg, args, (long)b0, smem, px, WARPS_PER_BLOCK);
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
    return MPCB200_OK;
}

static int create_impl(mpcb200_handle** out, const mpcb200_config* cfg, int model);
int mpcb200_create(mpcb200_handle** out, const mpcb200_config* cfg) { return create_impl(out, cfg, 0); }
int mpcb200_create_frenet(mpcb200_handle** out, const mpcb200_config* cfg) { return create_impl(out, cfg, 1); }

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
        int roles[32 * ROLE_STRIDE_F];
        const int rs = role_stride_of(model);
        for (int l = 0; l < 32; l++) riccati_roles(l, cfg->N, w_sd_of(h->team_warps, model), roles + l * rs, model);
        TRY_OR_FREE(cudaMalloc((void**)&h->d_roles, sizeof(roles)));
        TRY_OR_FREE(cudaMemcpy(h->d_roles, roles, sizeof(int) * 32 * rs, cudaMemcpyHostToDevice));
    }
    if (model && h->team_warps > 1) {
        h->smem_bytes = (size_t)smem_doubles_per_team(cfg->N, 1) * sizeof(double);
        const void* fn = (h->team_warps == 2) ? (const void*)mpc_solve_long_kernel<2, 1> : (const void*)mpc_solve_long_kernel<3, 1>;
        TRY_OR_FREE(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
        TRY_OR_FREE(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TRY_OR_FREE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->blocks_per_sm, fn, h->team_warps * 32, h->smem_bytes));
    } else if (model) {
        const size_t team = (size_t)smem_doubles_per_team(cfg->N, 1) * sizeof(double);
        int b4 = 0, b3 = 0;
        TRY_OR_FREE(cudaFuncSetAttribute(mpc_solve_frenet_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * team)));
        TRY_OR_FREE(cudaFuncSetAttribute(mpc_solve_frenet_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TRY_OR_FREE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b4, mpc_solve_frenet_kernel<4>, 4 * 32, 4 * team));
        TRY_OR_FREE(cudaFuncSetAttribute(mpc_solve_frenet_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(3 * team)));
        TRY_OR_FREE(cudaFuncSetAttribute(mpc_solve_frenet_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TRY_OR_FREE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b3, mpc_solve_frenet_kernel<3>, 3 * 32, 3 * team));
        /* measured (tools/frenet_bench.py): N = 8: 12 warps at 168 registers 6.06 M solves/s vs 8 warps at 255 registers 5.80 M;
         * N = 20: 9 warps 2.88 M vs 8 warps 3.23 M -- the spill code of the 168-register build pays only for >= 1.4x the warps */
        h->frenet_wpb = (10 * 3 * b3 >= 14 * 4 * b4) ? 3 : 4;
        if (const char* e = getenv("MPCB200_FRENET_WPB")) { int v = atoi(e); if (v == 3 || v == 4) h->frenet_wpb = v; }  /* tuning aid */
        h->blocks_per_sm = (h->frenet_wpb == 3) ? b3 : b4;
        h->smem_bytes = h->frenet_wpb * team;
    } else if (h->team_warps == 1) {
        h->smem_bytes = ((size_t)WARPS_PER_BLOCK * smem_doubles_per_team(cfg->N) + WARPS_PER_BLOCK * ROLLOUT_PX) * sizeof(double);
        TRY_OR_FREE(cudaFuncSetAttribute(mpc_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
        TRY_OR_FREE(cudaFuncSetAttribute(mpc_solve_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TRY_OR_FREE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->blocks_per_sm, mpc_solve_kernel, WARPS_PER_BLOCK * 32, h->smem_bytes));
        TRY_OR_FREE(cudaFuncSetAttribute(mpc_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
        TRY_OR_FREE(cudaFuncSetAttribute(mpc_rollout_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TRY_OR_FREE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->rollout_blocks_per_sm, mpc_rollout_kernel, WARPS_PER_BLOCK * 32, h->smem_bytes));
        if (h->rollout_blocks_per_sm < 1) h->rollout_blocks_per_sm = 1;
    } else {
        h->smem_bytes = (size_t)smem_doubles_per_team(cfg->N) * sizeof(double);
        const void* fn = (h->team_warps == 2) ? (const void*)mpc_solve_long_kernel<2> : (const void*)mpc_solve_long_kernel<3>;
        TRY_OR_FREE(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
        TRY_OR_FREE(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        TRY_OR_FREE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->blocks_per_sm, fn, h->team_warps * 32, h->smem_bytes));
    }
    if (h->blocks_per_sm < 1) h->blocks_per_sm = 1;
    if (const char* e = getenv("MPCB200_BLOCKS_PER_SM")) { int v = atoi(e); if (v >= 1 && v < h->blocks_per_sm) h->blocks_per_sm = v; }  /* tuning aid */
#undef TRY_OR_FREE
    *out = h;
    return MPCB200_OK;
}

int mpcb200_destroy(mpcb200_handle* h) {
    if (!h) return MPCB200_EINVAL;
    cudaSetDevice(h->device);
    DevBuf* bufs[] = {&h->d_state, &h->d_ref, &h->d_vdes, &h->d_uprev, &h->d_warm, &h->d_u0, &h->d_cost, &h->d_status, &h->d_iters, &h->d_traj,
                      &h->d_path[0], &h->d_path[1], &h->d_path[2], &h->d_pose, &h->d_pathof, &h->d_log, &h->d_final, &h->d_stop, &h->d_stage};
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
    return MPCB200_OK;
}

int mpcb200_set_stream(mpcb200_handle* h, void* s) {
    if (!h) return MPCB200_EINVAL;
    h->stream = s ? (cudaStream_t)s : h->own_stream;
    drop_graphs(h);
    return MPCB200_OK;
}

static int launch_solve(mpcb200_handle* h, int64_t B, const BatchPtrs& io, const RefGen& rg, unsigned long long* zeroed_counter = nullptr) {
    unsigned long long* counter = zeroed_counter ? zeroed_counter : h->d_counter;
    if (!zeroed_counter) CUDA_TRY(h, cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned long long), h->stream));
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
        if (n) { paths[i].t = base; paths[i].X = base + n; paths[i].Y = base + 2 * (size_t)n; paths[i].psi = base + 3 * (size_t)n; paths[i].s = base + 4 * (size_t)n; }
    }
}

/* Host pointers, small batch (the control loop's batch of one): latency is the driver calls, not the bytes.  All
 * inputs are packed into one pinned buffer (with the zeroed problem counter in front) and go over in ONE copy, all
 * outputs come back in ONE copy: 7 driver calls instead of ~16. */
static const int64_t SMALL_BATCH = 64;
static int solve_small_host(mpcb200_handle* h, int64_t B, const double* state, const double* ref, const double* v_des,
                            const double* u_prev, double* warm, double* u0, double* cost, int32_t* status, int32_t* iters,
                            double* traj) {
    const int N = h->cfg.N;
    const size_t nt = 6 * (size_t)N + 4, nr = h->model ? 4 : 3 * ((size_t)N + 1);
    /* layout in doubles: [counter, pad] state ref uprev vdes | warm | u0 cost traj status/iters(int32 pairs) */
    const size_t o_state = 2, o_ref = o_state + 4 * B, o_uprev = o_ref + nr * B, o_vdes = o_uprev + 2 * B, o_warm = o_vdes + B;
    const size_t o_u0 = o_warm + nt * B, o_cost = o_u0 + 2 * B, o_traj = o_cost + B, o_stat = o_traj + nt * B, o_iter = o_stat + (B + 1) / 2;
    const size_t total = o_iter + (B + 1) / 2;
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
    memcpy(hs + o_state, state, 4 * B * sizeof(double));
    memcpy(hs + o_ref, ref, nr * B * sizeof(double));
    memcpy(hs + o_uprev, u_prev, 2 * B * sizeof(double));
    if (v_des) memcpy(hs + o_vdes, v_des, B * sizeof(double));
    if (warm) memcpy(hs + o_warm, warm, nt * B * sizeof(double));
    const size_t in_doubles = warm ? o_u0 : o_warm;
    cudaStream_t s = h->stream;
    BatchPtrs io{ds + o_state, ds + o_ref, v_des ? ds + o_vdes : nullptr, ds + o_uprev, warm ? ds + o_warm : nullptr, ds + o_u0,
                 ds + o_cost, (int*)(ds + o_stat), (int*)(ds + o_iter), traj ? ds + o_traj : nullptr};
    RefGen rg;
    memset(&rg, 0, sizeof(rg));
    const size_t out_from = warm ? o_warm : o_u0;
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
        const int flags = (v_des ? 1 : 0) | (warm ? 2 : 0) | (traj ? 4 : 0);
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
    memcpy(u0, hs + o_u0, 2 * B * sizeof(double));
    if (cost) memcpy(cost, hs + o_cost, B * sizeof(double));
    if (status) memcpy(status, hs + o_stat, B * sizeof(int32_t));
    if (iters) memcpy(iters, hs + o_iter, B * sizeof(int32_t));
    if (traj) memcpy(traj, hs + o_traj, nt * B * sizeof(double));
    if (warm) memcpy(warm, hs + o_warm, nt * B * sizeof(double));
    float ms = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->stats.kernel_ms = ms;
    return MPCB200_OK;
}

/* common body of mpcb200_solve_batch (ref given) and mpcb200_solve_batch_on_path (path_of given) */
static int solve_batch_impl(mpcb200_handle* h, const char* who, int64_t B, const double* state, const double* ref, const int32_t* path_of,
                            int32_t track_using_time, double target_vel, const double* v_des, const double* u_prev, double* warm,
                            double* u0, double* cost, int32_t* status, int32_t* iters, double* traj, double* ref_out, int32_t* stop,
                            int32_t mem_space) {
    if (!h) return MPCB200_EINVAL;
    if (B < 0) return fail(h, MPCB200_EINVAL, "%s: B=%lld", who, (long long)B);
    if (mem_space != MPCB200_HOST && mem_space != MPCB200_DEVICE) return fail(h, MPCB200_EINVAL, "%s: mem_space=%d", who, mem_space);
    memset(&h->stats, 0, sizeof(h->stats));
    if (B == 0) return MPCB200_OK;
    const bool on_path = (path_of != nullptr);
    if (!state || (!ref && !on_path) || !u_prev || !u0) return fail(h, MPCB200_EINVAL, "%s: state, %s, u_prev and u0 are required", who, on_path ? "path_of" : "ref");
    if (on_path && h->model) return fail(h, MPCB200_EINVAL, "%s: not available for the Frenet-frame variant", who);
    if (on_path && h->team_warps != 1) return fail(h, MPCB200_EINVAL, "%s: on-device reference generation needs N <= 31 (N=%d)", who, h->cfg.N);
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int N = h->cfg.N;
    const size_t nt = 6 * (size_t)N + 4, nr = h->model ? 4 : 3 * ((size_t)N + 1);
    RefGen rg;
    memset(&rg, 0, sizeof(rg));
    if (on_path) {
        fill_paths(h, rg.paths);
        rg.track_using_time = track_using_time; rg.target_vel = target_vel;
        if (mem_space == MPCB200_HOST) {
            for (int64_t b = 0; b < B; b++)
                if (path_of[b] < 0 || path_of[b] > 2 || h->path_n[path_of[b]] == 0)
                    return fail(h, MPCB200_EINVAL, "%s: problem %lld uses path %d, which was not set with mpcb200_set_path", who, (long long)b, path_of[b]);
        } else {
            for (int i = 0; i < 3; i++)   /* device ids cannot be inspected here: every table must be present */
                if (h->path_n[i] == 0) return fail(h, MPCB200_EINVAL, "%s: with device pointers all three path tables must be set (path %d is not)", who, i);
        }
    }
    if (mem_space == MPCB200_DEVICE) {
        BatchPtrs io{state, ref, v_des, u_prev, warm, u0, cost, status, iters, traj};
        rg.path_of = path_of; rg.ref_out = ref_out; rg.stop = stop;
        return launch_solve(h, B, io, rg);
    }
    if (!on_path && B <= SMALL_BATCH) return solve_small_host(h, B, state, ref, v_des, u_prev, warm, u0, cost, status, iters, traj);
    /* host pointers: stage through the handle's device buffers */
    const size_t bs = B * 4 * sizeof(double), br = B * nr * sizeof(double), bu = B * 2 * sizeof(double), bt = B * nt * sizeof(double);
    int rc;
    if ((rc = ensure(h, h->d_state, bs))) return rc;
    if ((!on_path || ref_out) && (rc = ensure(h, h->d_ref, br))) return rc;
    if (on_path && (rc = ensure(h, h->d_pathof, B * sizeof(int32_t)))) return rc;
    if (on_path && stop && (rc = ensure(h, h->d_stop, B * sizeof(int32_t)))) return rc;
    if ((rc = ensure(h, h->d_uprev, bu))) return rc;
    if ((rc = ensure(h, h->d_u0, bu))) return rc;
    if (v_des && (rc = ensure(h, h->d_vdes, B * sizeof(double)))) return rc;
    if (warm && (rc = ensure(h, h->d_warm, bt))) return rc;
    if (cost && (rc = ensure(h, h->d_cost, B * sizeof(double)))) return rc;
    if (status && (rc = ensure(h, h->d_status, B * sizeof(int32_t)))) return rc;
    if (iters && (rc = ensure(h, h->d_iters, B * sizeof(int32_t)))) return rc;
    if (traj && (rc = ensure(h, h->d_traj, bt))) return rc;
    cudaStream_t s = h->stream;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_state.p, state, bs, cudaMemcpyHostToDevice, s));
    h->stats.h2d_bytes += bs;
    if (on_path) { CUDA_TRY(h, cudaMemcpyAsync(h->d_pathof.p, path_of, B * sizeof(int32_t), cudaMemcpyHostToDevice, s)); h->stats.h2d_bytes += B * sizeof(int32_t); }
    else { CUDA_TRY(h, cudaMemcpyAsync(h->d_ref.p, ref, br, cudaMemcpyHostToDevice, s)); h->stats.h2d_bytes += br; }
    CUDA_TRY(h, cudaMemcpyAsync(h->d_uprev.p, u_prev, bu, cudaMemcpyHostToDevice, s));
    h->stats.h2d_bytes += bu;
    if (v_des) { CUDA_TRY(h, cudaMemcpyAsync(h->d_vdes.p, v_des, B * sizeof(double), cudaMemcpyHostToDevice, s)); h->stats.h2d_bytes += B * sizeof(double); }
    if (warm) { CUDA_TRY(h, cudaMemcpyAsync(h->d_warm.p, warm, bt, cudaMemcpyHostToDevice, s)); h->stats.h2d_bytes += bt; }
    BatchPtrs io{(const double*)h->d_state.p, on_path ? nullptr : (const double*)h->d_ref.p, v_des ? (const double*)h->d_vdes.p : nullptr,
                 (const double*)h->d_uprev.p, warm ? (double*)h->d_warm.p : nullptr, (double*)h->d_u0.p,
                 cost ? (double*)h->d_cost.p : nullptr, status ? (int*)h->d_status.p : nullptr,
                 iters ? (int*)h->d_iters.p : nullptr, traj ? (double*)h->d_traj.p : nullptr};
    if (on_path) {
        rg.path_of = (const int*)h->d_pathof.p;
        rg.ref_out = ref_out ? (double*)h->d_ref.p : nullptr;
        rg.stop = stop ? (int*)h->d_stop.p : nullptr;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev0, s));
    if ((rc = launch_solve(h, B, io, rg))) return rc;
    CUDA_TRY(h, cudaEventRecord(h->ev1, s));
    CUDA_TRY(h, cudaMemcpyAsync(u0, h->d_u0.p, bu, cudaMemcpyDeviceToHost, s));
    h->stats.d2h_bytes += bu;
    if (cost) { CUDA_TRY(h, cudaMemcpyAsync(cost, h->d_cost.p, B * sizeof(double), cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += B * sizeof(double); }
    if (status) { CUDA_TRY(h, cudaMemcpyAsync(status, h->d_status.p, B * sizeof(int32_t), cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += B * sizeof(int32_t); }
    if (iters) { CUDA_TRY(h, cudaMemcpyAsync(iters, h->d_iters.p, B * sizeof(int32_t), cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += B * sizeof(int32_t); }
    if (traj) { CUDA_TRY(h, cudaMemcpyAsync(traj, h->d_traj.p, bt, cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += bt; }
    if (warm) { CUDA_TRY(h, cudaMemcpyAsync(warm, h->d_warm.p, bt, cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += bt; }
    if (on_path && ref_out) { CUDA_TRY(h, cudaMemcpyAsync(ref_out, h->d_ref.p, br, cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += br; }
    if (on_path && stop) { CUDA_TRY(h, cudaMemcpyAsync(stop, h->d_stop.p, B * sizeof(int32_t), cudaMemcpyDeviceToHost, s)); h->stats.d2h_bytes += B * sizeof(int32_t); }
    CUDA_TRY(h, cudaStreamSynchronize(s));
    float ms = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->stats.kernel_ms = ms;
    return MPCB200_OK;
}

int mpcb200_solve_batch_frenet(mpcb200_handle* h, int64_t B, const double* state, const double* k_coeffs, const double* v_des,
                               const double* u_prev, double* warm, double* u0, double* cost, int32_t* status, int32_t* iters,
                               double* traj, int32_t mem_space) {
    if (h && !h->model) return fail(h, MPCB200_EINVAL, "mpcb200_solve_batch_frenet: the handle was not created with mpcb200_create_frenet");
    return solve_batch_impl(h, "mpcb200_solve_batch_frenet", B, state, k_coeffs, nullptr, 1, 0.0, v_des, u_prev, warm, u0, cost, status, iters,
                            traj, nullptr, nullptr, mem_space);
}

int mpcb200_set_cost_frenet(mpcb200_handle* h, const double w[7]) {
    if (!h || !w) return fail(h, MPCB200_EINVAL, "mpcb200_set_cost_frenet: NULL argument");
    if (!h->model) return fail(h, MPCB200_EINVAL, "mpcb200_set_cost_frenet: the handle was not created with mpcb200_create_frenet");
    for (int i = 0; i < 7; i++) if (!(w[i] >= 0.0)) return fail(h, MPCB200_EINVAL, "mpcb200_set_cost_frenet: weight %d is negative or NaN", i);
    h->w[0] = 0.0;
    memcpy(h->w + 1, w, 7 * sizeof(double));
    drop_graphs(h);
    return MPCB200_OK;
}

int mpcb200_solve_batch(mpcb200_handle* h, int64_t B, const double* state, const double* ref, const double* v_des,
                        const double* u_prev, double* warm, double* u0, double* cost, int32_t* status, int32_t* iters,
                        double* traj, int32_t mem_space) {
    if (h && h->model) return fail(h, MPCB200_EINVAL, "mpcb200_solve_batch: Frenet handle; use mpcb200_solve_batch_frenet");
    return solve_batch_impl(h, "mpcb200_solve_batch", B, state, ref, nullptr, 1, 0.0, v_des, u_prev, warm, u0, cost, status, iters, traj,
                            nullptr, nullptr, mem_space);
}

int mpcb200_solve_batch_on_path(mpcb200_handle* h, int64_t B, const double* state, const int32_t* path_of, int32_t track_using_time,
                                double target_vel, const double* v_des, const double* u_prev, double* warm, double* u0, double* cost,
                                int32_t* status, int32_t* iters, double* traj, double* ref_out, int32_t* stop, int32_t mem_space) {
    if (h && !path_of) return fail(h, MPCB200_EINVAL, "mpcb200_solve_batch_on_path: path_of is required");
    return solve_batch_impl(h, "mpcb200_solve_batch_on_path", B, state, nullptr, path_of, track_using_time, target_vel, v_des, u_prev, warm,
                            u0, cost, status, iters, traj, ref_out, stop, mem_space);
}

int mpcb200_set_path(mpcb200_handle* h, int32_t path_id, int32_t n, const double* t, const double* X, const double* Y,
                     const double* psi, const double* s) {
    if (!h) return MPCB200_EINVAL;
    if (path_id < 0 || path_id > 2) return fail(h, MPCB200_EINVAL, "mpcb200_set_path: path_id %d outside [0,2]", path_id);
    if (n < 2 || !t || !X || !Y || !psi || !s) return fail(h, MPCB200_EINVAL, "mpcb200_set_path: need n >= 2 samples and five columns");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = ensure(h, h->d_path[path_id], (size_t)5 * n * sizeof(double));
    if (rc) return rc;
    const double* cols[5] = {t, X, Y, psi, s};
    for (int i = 0; i < 5; i++)
        CUDA_TRY(h, cudaMemcpyAsync((double*)h->d_path[path_id].p + (size_t)i * n, cols[i], (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->path_n[path_id] = n;
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
    if (h->team_warps != 1) return fail(h, MPCB200_EINVAL, "mpcb200_rollout: closed-loop rollouts need N <= 31 (N=%d)", h->cfg.N);
    for (int64_t b = 0; b < B; b++)
        if (path_of[b] < 0 || path_of[b] > 2 || h->path_n[path_of[b]] == 0)
            return fail(h, MPCB200_EINVAL, "mpcb200_rollout: vehicle %lld uses path %d, which was not set with mpcb200_set_path", (long long)b, path_of[b]);
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc;
    const size_t bl = (size_t)T * B * 8 * sizeof(double), bf = (size_t)B * 8 * sizeof(double);
    if ((rc = ensure(h, h->d_pose, (size_t)B * 3 * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->d_pathof, (size_t)B * sizeof(int32_t)))) return rc;
    if (log && (rc = ensure(h, h->d_log, bl))) return rc;
    if (final_state && (rc = ensure(h, h->d_final, bf))) return rc;
    cudaStream_t s = h->stream;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_pose.p, pose0, (size_t)B * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_pathof.p, path_of, (size_t)B * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    h->stats.h2d_bytes += (size_t)B * (3 * sizeof(double) + sizeof(int32_t));
    RolloutArgs a;
    memset(&a, 0, sizeof(a));
    a.pose0 = (const double*)h->d_pose.p; a.path_of = (const int*)h->d_pathof.p;
    fill_paths(h, a.paths);
    a.T = T; a.track_using_time = track_using_time; a.target_vel = target_vel;
    a.log = log ? (double*)h->d_log.p : nullptr; a.final_state = final_state ? (double*)h->d_final.p : nullptr; a.B = (long)B;
    CUDA_TRY(h, cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned long long), s));
    long long blocks_needed = (B + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    long long max_blocks = (long long)h->num_sms * h->rollout_blocks_per_sm;
    int grid = (int)(blocks_needed < max_blocks ? blocks_needed : max_blocks);
    CUDA_TRY(h, cudaEventRecord(h->ev0, s));
    mpc_rollout_kernel<<<grid, WARPS_PER_BLOCK * 32, h->smem_bytes, s>>>(make_kcfg(h), a, h->d_counter);
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
