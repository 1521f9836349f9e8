// mpc_kernel.cuh -- one MPC problem per warp: interior-point iteration with a Riccati KKT solve.
//
// What it replaces (reference file:line):
//   NLP transcription   scripts/mpc_utils/MKZMPCPathFollower.jl:65-123  (JuMP AD -> analytic
//                       stage Jacobians A_k, B_k and the (psi,v,df) Lagrangian-Hessian block)
//   solve(mdl)          scripts/mpc_utils/MKZMPCPathFollower.jl:176      (Ipopt: barrier update,
//                       fraction-to-boundary, filter line search, inertia correction; MUMPS LDL^T
//                       -> block-tridiagonal Riccati recursion on the condensed KKT system)
//
// Lane k of the warp owns stage k of the horizon (k = 0..N, N <= 31): its state (x,y,psi,v),
// input (acc,df), the multipliers of the equality rows that define s_k, its bound multipliers
// and the rate row that ends at u_k.  Stage-parallel work (model evaluation, residuals,
// multiplier updates, norms) runs with lanes = stages; the serial Riccati recursion switches
// to lanes = matrix entries, with per-stage data staged in shared memory.
//
// The iteration follows oracle/mpc_oracle.c step for step (same formulas, same constants);
// only the linear algebra differs (condensed Riccati here, full-space Bunch-Kaufman there).
#pragma once
#include "warp_prims.cuh"

namespace mpcb200 {

struct KCfg {
    int N, max_iter, start_mode, pad_;
    double dt, dtc, La, Lb, vmin, vmax, amax, smax, admax, sdmax, tol;
    double w[8];  // cx, cy, cpsi, cv, cdacc, cddf, cacc, cdf
};

// ---- Ipopt 3.12 default constants (same values as oracle/mpc_oracle.c) ----
#define K_KAPPA_EPS 10.0
#define K_KAPPA_MU 0.2
#define K_THETA_MU 1.5
#define K_MU_INIT 0.1
#define K_TAU_MIN 0.99
#define K_BOUND_PUSH 1e-2
#define K_BOUND_FRAC 1e-2
#define K_BOUND_RELAX 1e-8
#define K_S_MAX 100.0
#define K_KAPPA_SIGMA 1e10
#define K_GAMMA_THETA 1e-5
#define K_GAMMA_PHI 1e-8
#define K_ETA_PHI 1e-8
#define K_DELTA 1.0
#define K_S_THETA 1.1
#define K_S_PHI 2.3
#define K_ALPHA_MIN_FRAC 0.05
#define K_MAX_SOC 4
#define K_KAPPA_SOC 0.99
#define K_OBJ_MAX_INC 5.0
#define K_DW_INIT 1e-4
#define K_DW_MIN 1e-20
#define K_DW_MAX 1e20
#define K_DW_INC_FIRST 100.0
#define K_DW_INC 8.0
#define K_DW_DEC (1.0 / 3.0)
#define K_SCALE_MAX_GRAD 100.0
#define K_Y_INIT_MAX 1e3
#define K_ACCEPT_TOL 1e-6
#define K_ACCEPT_ITER 15
#define K_EPS 2.220446049250313e-16

// ---- shared-memory layout of one warp, in doubles ----
// work area
#define W_P 0       // 6x6 cost-to-go Hessian, full symmetric storage, row stride 6
#define W_PV 36     // 6   cost-to-go gradient
#define W_T 42      // 6x6: columns psi,v,a,df of P*M and column 4 = P*r + p
#define W_F 78      // 6x6 stage Hessian F over (x,y,psi,v,a,df)
#define W_FV 114    // 6   stage gradient f
#define W_EX 120    // 6   unit vector e_x
#define W_EY 126    // 6   unit vector e_y
#define W_Z 132     // constant 0.0
#define W_DUMMY 133 // sink for inactive lanes
#define W_SD 134    // stage records start here
// stage record (dense)
#define SD_MT 0     // 5 columns x 6: d(next state, next prev-input)/d(psi | v | a | df) and residual r
#define SD_HXX 30
#define SD_HYY 31
#define SD_HPP 32
#define SD_HPV 33
#define SD_HVV 34
#define SD_HPD 35
#define SD_HVD 36
#define SD_HAA 37
#define SD_HDD 38
#define SD_CA 39
#define SD_CD 40
#define SD_NCA 41
#define SD_ZERO2 42
#define SD_NCD 43
#define SD_GX 44  // GX,GY,GP,GV,GA,GD
#define SD_NF 50
#define SDS 51      // odd stride: lane-k writes of one field are bank-conflict free
#define KST_STRIDE 14

#ifdef MPC_HOST_EMU
#define MPC_HD inline
#else
#define MPC_HD __host__ __device__ __forceinline__
#endif
MPC_HD int smem_doubles_per_warp(int N) { return W_SD + (N + 1) * SDS + N * KST_STRIDE + 2; }

MPC_DEV double warp_sum(double v) {
    for (int o = 16; o; o >>= 1) v += shfl_xor(v, o);
    return v;
}
MPC_DEV double warp_max(double v) {
    for (int o = 16; o; o >>= 1) { double t = shfl_xor(v, o); v = (t > v || t != t) ? t : v; }
    return v;
}
MPC_DEV double warp_min(double v) {
    for (int o = 16; o; o >>= 1) { double t = shfl_xor(v, o); v = (t < v) ? t : v; }
    return v;
}
// inclusive suffix sum over lanes: out_k = sum_{j >= k} v_j
MPC_DEV double warp_suffix_sum(double v) {
    const int l = lane_id();
    for (int o = 1; o < 32; o <<= 1) { double t = shfl_down(v, o); if (l + o < 32) v += t; }
    return v;
}
MPC_DEV double dmax_(double a, double b) { return a > b ? a : b; }
MPC_DEV double dmin_(double a, double b) { return a < b ? a : b; }

MPC_DEV void push_interior(double& v, double lo, double hi) {
    double pl = dmin_(K_BOUND_PUSH * dmax_(1.0, fabs(lo)), K_BOUND_FRAC * (hi - lo));
    double pu = dmin_(K_BOUND_PUSH * dmax_(1.0, fabs(hi)), K_BOUND_FRAC * (hi - lo));
    if (v < lo + pl) v = lo + pl;
    if (v > hi - pu) v = hi - pu;
}
MPC_DEV bool cmp_le(double lhs, double rhs, double bas) { return lhs - rhs <= 10.0 * K_EPS * fabs(bas); }

// Stage model f(s,u) (MKZMPCPathFollower.jl:115-123) with the trig terms kept for derivatives.
struct StageTrig {
    double cs, sn;   // cos, sin (psi + beta)
    double cb, sb;   // cos, sin beta
    double b1, b2;   // beta', beta''
};
MPC_DEV void stage_trig(const KCfg& c, double psi, double df, StageTrig& t) {
    // beta = atan(r tan df): sin/cos beta in closed form, valid for |df| < pi/2 (bounds keep |df| <= 0.5)
    const double r = c.Lb / (c.La + c.Lb);
    double sd, cd, sp, cp;
    mpc_sincos(df, &sd, &cd);
    mpc_sincos(psi, &sp, &cp);
    const double D = cd * cd + r * r * sd * sd;
    const double inv = 1.0 / sqrt(D);
    t.cb = cd * inv;
    t.sb = r * sd * inv;
    t.cs = cp * t.cb - sp * t.sb;
    t.sn = sp * t.cb + cp * t.sb;
    t.b1 = r / D;
    t.b2 = r * (1.0 - r * r) * (2.0 * sd * cd) / (D * D);
}

// Everything one lane (= one stage) carries through the iteration.
struct LaneState {
    // primal
    double sx, sy, sp, sv, ua, ud;
    // equality multipliers of the rows that define s_k (init rows for k = 0, dynamics rows k-1 -> k)
    double yx, yy, yp, yv;
    // bound multipliers
    double zvL, zvU, zaL, zaU, zdL, zdU;
    // rate row ending at u_k: [0] steering, [1] acceleration: slack, y_d, slack-bound multipliers
    double rs[2], ryd[2], rvL[2], rvU[2];
};

struct StepState {
    double dsx, dsy, dsp, dsv, dua, dud;  // primal step
    double drs[2];                        // slack step
    double nyx, nyy, nyp, nyv;            // NEW equality multipliers (y + dy)
    double nyd[2];                        // NEW range multipliers
};

// The solver for one problem; all 32 lanes of the warp call it together.
struct WarpSolver {
    const KCfg& c;
    double* sm;    // this warp's shared memory
    const int k;   // lane = stage
    const int N;
    const bool isS, isU, isR;  // lane owns a state / an input / a rate row
    // bounds (relaxed by bound_relax_factor)
    double vLo, vHi, aLo, aHi, dLo, dHi, rLo[2], rHi[2];
    // problem data
    double xr, yr, pr, vdes, st0[4], uprev[2];
    double wx, wy, wp, wv;  // stage cost weights (0 where the sum does not run)
    double sigma;           // objective scaling
    LaneState L;
    StepState D;
    // per-iteration evaluation products
    StageTrig tg;
    double rdyn[4];   // f(s_k,u_k) - s_{k+1}   (= -c of the rows entering k+1), lanes k < N
    double rinit[4];  // state - s_0            (= -c of the init rows), lane 0
    double dres[2];   // d(x) - slack of this lane's rate row
    double gx, gy, gp, gv, ga, gd;  // scaled objective gradient of this stage's variables

    MPC_DEV WarpSolver(const KCfg& cfg, double* smem)
        : c(cfg), sm(smem), k(lane_id()), N(cfg.N),
          isS(lane_id() <= cfg.N), isU(lane_id() < cfg.N),
          isR((lane_id() == 0 || lane_id() >= 2) && lane_id() < cfg.N) {}

    // ------------------------------------------------------------------
    // model evaluation at (s,u) given per lane; fills rd[4] (dynamics defect of rows k -> k+1),
    // ri[4] (init defect, lane 0) and dr[2] (rate-row defect d(x) - slack)
    // ------------------------------------------------------------------
    MPC_DEV void eval_defects(double sx, double sy, double sp, double sv, double ua, double ud,
                              const double* slack, StageTrig& t, double* rd, double* ri, double* dr) const {
        stage_trig(c, sp, ud, t);
        const double fx = sx + c.dt * (sv * t.cs);
        const double fy = sy + c.dt * (sv * t.sn);
        const double fp = sp + c.dt * (sv / c.Lb * t.sb);
        const double fv = sv + c.dt * ua;
        const double nx = shfl_down(sx, 1), ny = shfl_down(sy, 1), np = shfl_down(sp, 1), nv = shfl_down(sv, 1);
        rd[0] = isU ? fx - nx : 0.0;
        rd[1] = isU ? fy - ny : 0.0;
        rd[2] = isU ? fp - np : 0.0;
        rd[3] = isU ? fv - nv : 0.0;
        const bool l0 = (k == 0);
        ri[0] = l0 ? st0[0] - sx : 0.0;
        ri[1] = l0 ? st0[1] - sy : 0.0;
        ri[2] = l0 ? st0[2] - sp : 0.0;
        ri[3] = l0 ? st0[3] - sv : 0.0;
        // rate rows: k = 0: u_0 - u_prev ; k >= 2: u_k - u_{k-1}
        const double pa = shfl_up(ua, 1), pd = shfl_up(ud, 1);
        const double ba = l0 ? uprev[1] : pa, bd = l0 ? uprev[0] : pd;
        dr[0] = isR ? (ud - bd) - slack[0] : 0.0;
        dr[1] = isR ? (ua - ba) - slack[1] : 0.0;
    }

    MPC_DEV double theta_of(const double* rd, const double* ri, const double* dr) const {
        double t = fabs(rd[0]) + fabs(rd[1]) + fabs(rd[2]) + fabs(rd[3]) + fabs(ri[0]) + fabs(ri[1]) + fabs(ri[2]) +
                   fabs(ri[3]) + fabs(dr[0]) + fabs(dr[1]);
        return warp_sum(t);
    }

    // unscaled objective (MKZMPCPathFollower.jl:97-103) at a per-lane point
    MPC_DEV double objective(double sx, double sy, double sp, double sv, double ua, double ud) const {
        const double na = shfl_down(ua, 1), nd = shfl_down(ud, 1);
        double f = 0.0;
        const double ex = sx - xr, ey = sy - yr, ep = sp - pr, ev = sv - vdes;
        f += wx * ex * ex + wy * ey * ey + wp * ep * ep + wv * ev * ev;
        if (isU) {
            f += c.w[6] * ua * ua + c.w[7] * ud * ud;
            if (k + 1 < N) { const double da = na - ua, dd = nd - ud; f += c.w[4] * da * da + c.w[5] * dd * dd; }
        }
        return warp_sum(f);
    }

    // sum of log slacks of this lane's bounds at a per-lane point
    MPC_DEV double log_barrier(double sv, double ua, double ud, const double* slack) const {
        double p = 1.0;
        if (isS) p *= (sv - vLo) * (vHi - sv);
        if (isU) p *= (ua - aLo) * (aHi - ua) * (ud - dLo) * (dHi - ud);
        if (isR) p *= (slack[0] - rLo[0]) * (rHi[0] - slack[0]) * (slack[1] - rLo[1]) * (rHi[1] - slack[1]);
        return warp_sum(log(p));
    }

    // scaled objective gradient at the current iterate -> gx..gd
    MPC_DEV void objective_gradient() {
        const double na = shfl_down(L.ua, 1), nd = shfl_down(L.ud, 1);
        const double pa = shfl_up(L.ua, 1), pd = shfl_up(L.ud, 1);
        gx = 2.0 * sigma * wx * (L.sx - xr);
        gy = 2.0 * sigma * wy * (L.sy - yr);
        gp = 2.0 * sigma * wp * (L.sp - pr);
        gv = 2.0 * sigma * wv * (L.sv - vdes);
        ga = 0.0; gd = 0.0;
        if (isU) {
            ga = 2.0 * sigma * c.w[6] * L.ua;
            gd = 2.0 * sigma * c.w[7] * L.ud;
            if (k >= 1) { ga += 2.0 * sigma * c.w[4] * (L.ua - pa); gd += 2.0 * sigma * c.w[5] * (L.ud - pd); }
            if (k + 1 < N) { ga -= 2.0 * sigma * c.w[4] * (na - L.ua); gd -= 2.0 * sigma * c.w[5] * (nd - L.ud); }
        }
    }

    // ------------------------------------------------------------------
    // Stage record assembly.  mode 0: least-squares multiplier system (identity Hessian,
    // unit slack weights); mode 1: primal-dual system with Lagrangian Hessian, Sigma, delta_w.
    // gs*/gu*: condensed gradient of this stage's variables; rr/ri: residual r_k and init step.
    // ------------------------------------------------------------------
    struct Cond {       // per-lane condensed quantities reused after the solve
        double Hxx, Hyy, Hpp, Hpv, Hvv, Hpd, Hvd;  // state/cross blocks (without the +C input parts)
        double Haa, Hdd;                           // input diagonal INCLUDING coupling C of this stage
        double Ca, Cd;                             // coupling with the previous input
        double gsx, gsy, gsp, gsv, gua, gud;       // condensed gradient
        double SrW[2];                             // (Sigma_s + delta_w) of this lane's rate row
        double br[2];                              // -mu/ss_L + mu/ss_U of the rate row
    };

    MPC_DEV void write_record(const Cond& q, const double* rr) {
        double* r = sm + W_SD + k * SDS;
        const double A02 = isU ? -c.dt * L.sv * tg.sn : 0.0, A03 = isU ? c.dt * tg.cs : 0.0;
        const double A12 = isU ? c.dt * L.sv * tg.cs : 0.0, A13 = isU ? c.dt * tg.sn : 0.0;
        const double A23 = isU ? c.dt * tg.sb / c.Lb : 0.0;
        const double b0 = A02 * tg.b1, b1v = A12 * tg.b1, b2v = isU ? c.dt * L.sv * tg.cb * tg.b1 / c.Lb : 0.0;
        // columns of [A B; 0 I] over rows (x,y,psi,v,prev_a,prev_df)
        r[SD_MT + 0] = A02; r[SD_MT + 1] = A12; r[SD_MT + 2] = 1.0; r[SD_MT + 3] = 0.0; r[SD_MT + 4] = 0.0; r[SD_MT + 5] = 0.0;      // psi
        r[SD_MT + 6] = A03; r[SD_MT + 7] = A13; r[SD_MT + 8] = A23; r[SD_MT + 9] = 1.0; r[SD_MT + 10] = 0.0; r[SD_MT + 11] = 0.0;    // v
        r[SD_MT + 12] = 0.0; r[SD_MT + 13] = 0.0; r[SD_MT + 14] = 0.0; r[SD_MT + 15] = c.dt; r[SD_MT + 16] = 1.0; r[SD_MT + 17] = 0.0; // a
        r[SD_MT + 18] = b0; r[SD_MT + 19] = b1v; r[SD_MT + 20] = b2v; r[SD_MT + 21] = 0.0; r[SD_MT + 22] = 0.0; r[SD_MT + 23] = 1.0; // df
        r[SD_MT + 24] = rr[0]; r[SD_MT + 25] = rr[1]; r[SD_MT + 26] = rr[2]; r[SD_MT + 27] = rr[3]; r[SD_MT + 28] = 0.0; r[SD_MT + 29] = 0.0;
        r[SD_HXX] = q.Hxx; r[SD_HYY] = q.Hyy; r[SD_HPP] = q.Hpp; r[SD_HPV] = q.Hpv; r[SD_HVV] = q.Hvv;
        r[SD_HPD] = q.Hpd; r[SD_HVD] = q.Hvd; r[SD_HAA] = q.Haa; r[SD_HDD] = q.Hdd;
        r[SD_CA] = q.Ca; r[SD_CD] = q.Cd; r[SD_NCA] = -q.Ca; r[SD_ZERO2] = 0.0; r[SD_NCD] = -q.Cd;
        r[SD_GX + 0] = q.gsx; r[SD_GX + 1] = q.gsy; r[SD_GX + 2] = q.gsp; r[SD_GX + 3] = q.gsv;
        r[SD_GX + 4] = q.gua; r[SD_GX + 5] = q.gud;
    }

    // only the pieces that change between inertia-correction attempts / SOC right-hand sides
    MPC_DEV void patch_record_diag(const Cond& q) {
        double* r = sm + W_SD + k * SDS;
        r[SD_HXX] = q.Hxx; r[SD_HYY] = q.Hyy; r[SD_HPP] = q.Hpp; r[SD_HVV] = q.Hvv; r[SD_HAA] = q.Haa; r[SD_HDD] = q.Hdd;
        r[SD_CA] = q.Ca; r[SD_CD] = q.Cd; r[SD_NCA] = -q.Ca; r[SD_NCD] = -q.Cd;
        r[SD_GX + 4] = q.gua; r[SD_GX + 5] = q.gud;
    }
    MPC_DEV void patch_record_rhs(const Cond& q, const double* rr) {
        double* r = sm + W_SD + k * SDS;
        r[SD_MT + 24] = rr[0]; r[SD_MT + 25] = rr[1]; r[SD_MT + 26] = rr[2]; r[SD_MT + 27] = rr[3];
        r[SD_GX + 4] = q.gua; r[SD_GX + 5] = q.gud;
    }

    // condensed Hessian/gradient for the primal-dual system at the current iterate.
    // rdr: rate-row residual to use (dres, or the SOC-accumulated one)
    MPC_DEV void build_cond_pd(Cond& q, double mu, double dw, const double* rdr) {
        // multipliers of the rows leaving this stage (held by lane k+1)
        const double y1x = shfl_down(L.yx, 1), y1y = shfl_down(L.yy, 1), y1p = shfl_down(L.yp, 1);
        double hpp = 0.0, hpv = 0.0, hpd = 0.0, hvd = 0.0, hdd = 0.0;
        if (isU) {
            const double v = L.sv, dt = c.dt;
            const double e1 = y1x * tg.cs + y1y * tg.sn;  // -> d2/dpsi2 direction
            const double e2 = y1x * tg.sn - y1y * tg.cs;
            hpp = dt * v * e1;
            hpv = dt * e2;
            hpd = tg.b1 * hpp;
            hvd = tg.b1 * hpv - y1p * dt * tg.cb * tg.b1 / c.Lb;
            hdd = tg.b1 * tg.b1 * hpp + v * tg.b2 * hpv - y1p * (dt * v / c.Lb) * (tg.cb * tg.b2 - tg.sb * tg.b1 * tg.b1);
        }
        double Sv = 0.0, bv = 0.0, Sa = 0.0, ba = 0.0, Sd = 0.0, bd = 0.0;
        if (isS) {
            const double sl = L.sv - vLo, su = vHi - L.sv;
            Sv = L.zvL / sl + L.zvU / su; bv = -mu / sl + mu / su;
        }
        if (isU) {
            double sl = L.ua - aLo, su = aHi - L.ua;
            Sa = L.zaL / sl + L.zaU / su; ba = -mu / sl + mu / su;
            sl = L.ud - dLo; su = dHi - L.ud;
            Sd = L.zdL / sl + L.zdU / su; bd = -mu / sl + mu / su;
        }
        double wrow[2] = {0.0, 0.0};
        q.SrW[0] = q.SrW[1] = 0.0; q.br[0] = q.br[1] = 0.0;
        if (isR) {
            for (int i = 0; i < 2; i++) {
                const double sl = L.rs[i] - rLo[i], su = rHi[i] - L.rs[i];
                q.SrW[i] = L.rvL[i] / sl + L.rvU[i] / su + dw;
                q.br[i] = -mu / sl + mu / su;
                wrow[i] = q.br[i] + q.SrW[i] * rdr[i];
            }
        }
        const double wn0 = shfl_down(wrow[0], 1), wn1 = shfl_down(wrow[1], 1);  // row ending at u_{k+1}
        q.Hxx = (isS ? 2.0 * sigma * wx : 0.0) + dw;
        q.Hyy = (isS ? 2.0 * sigma * wy : 0.0) + dw;
        q.Hpp = (isS ? 2.0 * sigma * wp : 0.0) + hpp + dw;
        q.Hpv = hpv;
        q.Hvv = (isS ? 2.0 * sigma * wv : 0.0) + Sv + dw;
        q.Hpd = hpd; q.Hvd = hvd;
        // coupling with the previous input: rate cost for k >= 1, rate row for k == 0 or k >= 2
        q.Ca = 0.0; q.Cd = 0.0;
        if (isU) {
            if (k >= 1) { q.Ca = 2.0 * sigma * c.w[4]; q.Cd = 2.0 * sigma * c.w[5]; }
            if (isR) { q.Ca += q.SrW[1]; q.Cd += q.SrW[0]; }
        }
        q.Haa = isU ? 2.0 * sigma * c.w[6] + Sa + dw + q.Ca : 1.0;
        q.Hdd = isU ? 2.0 * sigma * c.w[7] + Sd + dw + q.Cd + hdd : 1.0;
        q.gsx = gx; q.gsy = gy; q.gsp = gp; q.gsv = gv + bv;
        q.gua = isU ? ga + ba + wrow[1] - ((k + 1 < N) ? wn1 : 0.0) : 0.0;
        q.gud = isU ? gd + bd + wrow[0] - ((k + 1 < N) ? wn0 : 0.0) : 0.0;
    }

    // least-squares multiplier system: Hessian = I, slack weights = 1, no residuals
    MPC_DEV void build_cond_ls(Cond& q) {
        double wrow[2] = {0.0, 0.0};
        q.SrW[0] = q.SrW[1] = 0.0; q.br[0] = q.br[1] = 0.0;
        if (isR) for (int i = 0; i < 2; i++) { q.SrW[i] = 1.0; q.br[i] = -(L.rvL[i] - L.rvU[i]); wrow[i] = q.br[i]; }
        const double wn0 = shfl_down(wrow[0], 1), wn1 = shfl_down(wrow[1], 1);
        q.Hxx = q.Hyy = q.Hpp = q.Hvv = 1.0; q.Hpv = q.Hpd = q.Hvd = 0.0;
        q.Ca = (isU && isR) ? 1.0 : 0.0; q.Cd = q.Ca;
        q.Haa = 1.0 + q.Ca; q.Hdd = 1.0 + q.Cd;
        q.gsx = gx; q.gsy = gy; q.gsp = gp; q.gsv = gv + (isS ? (-L.zvL + L.zvU) : 0.0);
        q.gua = isU ? ga - L.zaL + L.zaU + wrow[1] - ((k + 1 < N) ? wn1 : 0.0) : 0.0;
        q.gud = isU ? gd - L.zdL + L.zdU + wrow[0] - ((k + 1 < N) ? wn0 : 0.0) : 0.0;
    }

    // ------------------------------------------------------------------
    // Riccati backward recursion over the stage records; lanes = matrix entries.
    // Returns false if some stage's reduced input Hessian is not positive definite
    // (the inertia of the full KKT matrix is wrong).  Gains go to the KST area.
    // ------------------------------------------------------------------
    struct Roles {
        int a_pb, a_mb, a_ex, a_out;
        int b_mb, b_mk, b_tb, b_h, b_hk, b_o1, b_o2;
        int d_xb, d_xk, d_r, d_out;
        int e_f0, e_f0k, e_f1, e_f1k, e_kb, e_o1, e_o2;
    };
    Roles R;

    MPC_DEV void init_roles() {
        const int l = k;
        // Round A: lanes 0..23 -> T[i][cc], lanes 24..29 -> T[i][4] (vector)
        {
            const int i = (l < 24) ? l / 4 : (l < 30 ? l - 24 : 0);
            const int cc = (l < 24) ? l % 4 : 4;
            R.a_pb = W_P + 6 * i;
            R.a_mb = W_SD + SD_MT + 6 * cc;
            R.a_ex = (l >= 24 && l < 30) ? W_PV + i : W_Z;
            R.a_out = (l < 30) ? W_T + 6 * i + cc : W_DUMMY;
        }
        // Round B: 21 symmetric pairs (c1 <= c2) then 6 vector entries
        {
            int c1 = 0, c2 = 0, vec = 0, act = 1;
            if (l < 21) { int t = l; c1 = 0; while (t >= 6 - c1) { t -= 6 - c1; c1++; } c2 = c1 + t; }
            else if (l < 27) { c1 = l - 21; vec = 1; }
            else act = 0;
            // coefficient column c1 of [A B; 0 I]: unit vectors for x,y, stage record otherwise
            if (c1 < 2) { R.b_mb = (c1 == 0) ? W_EX : W_EY; R.b_mk = 0; }
            else { R.b_mb = W_SD + SD_MT + 6 * (c1 - 2); R.b_mk = SDS; }
            if (vec) R.b_tb = W_T + 4;
            else R.b_tb = (c2 < 2) ? W_P + c2 : W_T + (c2 - 2);
            // H entry
            int h = -1;
            if (vec) h = SD_GX + c1;
            else if (c1 == c2) { const int dg[6] = {SD_HXX, SD_HYY, SD_HPP, SD_HVV, SD_HAA, SD_HDD}; h = dg[c1]; }
            else if (c1 == 2 && c2 == 3) h = SD_HPV;
            else if (c1 == 2 && c2 == 5) h = SD_HPD;
            else if (c1 == 3 && c2 == 5) h = SD_HVD;
            if (h >= 0) { R.b_h = W_SD + h; R.b_hk = SDS; } else { R.b_h = W_Z; R.b_hk = 0; }
            if (!act) { R.b_o1 = R.b_o2 = W_DUMMY; }
            else if (vec) { R.b_o1 = R.b_o2 = W_FV + c1; }
            else { R.b_o1 = W_F + 6 * c1 + c2; R.b_o2 = W_F + 6 * c2 + c1; }
        }
        // Round D: lanes 0..13 -> K[r][cidx], cidx over (x,y,psi,v,prev_a,prev_df,const)
        {
            const int r = (l < 14) ? l / 7 : 0, ci = (l < 14) ? l % 7 : 0;
            R.d_r = r;
            if (ci < 4) { R.d_xb = W_F + 6 * ci + 4; R.d_xk = 0; }
            else if (ci == 4) { R.d_xb = W_SD + SD_NCA; R.d_xk = SDS; }
            else if (ci == 5) { R.d_xb = W_SD + SD_ZERO2; R.d_xk = SDS; }
            else { R.d_xb = W_FV + 4; R.d_xk = 0; }
            R.d_out = (l < 14) ? r * 7 + ci : -1;
        }
        // Round E: 21 symmetric pairs over (x,y,psi,v,prev_a,prev_df) then 6 vector entries
        {
            int i = 0, j = 0, vec = 0, act = 1;
            if (l < 21) { int t = l; i = 0; while (t >= 6 - i) { t -= 6 - i; i++; } j = i + t; }
            else if (l < 27) { i = l - 21; vec = 1; }
            else act = 0;
            // F8[i][j]
            if (vec) { if (i < 4) { R.e_f0 = W_FV + i; R.e_f0k = 0; } else { R.e_f0 = W_Z; R.e_f0k = 0; } }
            else if (i < 4 && j < 4) { R.e_f0 = W_F + 6 * i + j; R.e_f0k = 0; }
            else if (i == 4 && j == 4) { R.e_f0 = W_SD + SD_CA; R.e_f0k = SDS; }
            else if (i == 5 && j == 5) { R.e_f0 = W_SD + SD_CD; R.e_f0k = SDS; }
            else { R.e_f0 = W_Z; R.e_f0k = 0; }
            // (F8[i][a], F8[i][df])
            if (i < 4) { R.e_f1 = W_F + 6 * i + 4; R.e_f1k = 0; }
            else if (i == 4) { R.e_f1 = W_SD + SD_NCA; R.e_f1k = SDS; }
            else { R.e_f1 = W_SD + SD_ZERO2; R.e_f1k = SDS; }
            R.e_kb = vec ? 6 : j;
            if (!act) { R.e_o1 = R.e_o2 = W_DUMMY; }
            else if (vec) { R.e_o1 = R.e_o2 = W_PV + i; }
            else { R.e_o1 = W_P + 6 * i + j; R.e_o2 = W_P + 6 * j + i; }
        }
        // constants in the work area
        if (l < 6) { sm[W_EX + l] = (l == 0) ? 1.0 : 0.0; sm[W_EY + l] = (l == 1) ? 1.0 : 0.0; }
        if (l == 0) { sm[W_Z] = 0.0; sm[W_DUMMY] = 0.0; }
        syncwarp();
    }

    MPC_DEV bool riccati_backward() {
        const int l = k;
        double* kst = sm + W_SD + (N + 1) * SDS;
        // terminal cost-to-go from record N
        {
            const double* rN = sm + W_SD + N * SDS;
            for (int e = l; e < 36; e += 32) {
                const int i = e / 6, j = e % 6;
                double v = 0.0;
                if (i == j && i < 4) v = rN[SD_HXX + (i == 0 ? 0 : i == 1 ? 1 : i == 2 ? 2 : 4)];
                sm[W_P + e] = v;
            }
            if (l < 6) sm[W_PV + l] = (l < 4) ? rN[SD_GX + l] : 0.0;
        }
        syncwarp();
        bool ok = true;
        for (int s = N - 1; s >= 0; s--) {
            const int so = s * SDS;
            // ---- Round A: T = P * [A B; 0 I](:, psi|v|a|df),  t = P r + p
            {
                const double* p = sm + R.a_pb;
                const double* m = sm + R.a_mb + so;
                const double t0 = p[0] * m[0] + p[1] * m[1] + p[2] * m[2];
                const double t1 = p[3] * m[3] + p[4] * m[4] + p[5] * m[5] + sm[R.a_ex];
                sm[R.a_out] = t0 + t1;
            }
            syncwarp();
            // ---- Round B: F = H + M' T,  f = g + M' t
            {
                const double* m = sm + R.b_mb + (R.b_mk ? so : 0);
                const double* t = sm + R.b_tb;
                const double h = sm[R.b_h + (R.b_hk ? so : 0)];
                const double t0 = m[0] * t[0] + m[1] * t[6] + m[2] * t[12];
                const double t1 = m[3] * t[18] + m[4] * t[24] + m[5] * t[30] + h;
                const double o = t0 + t1;
                sm[R.b_o1] = o; sm[R.b_o2] = o;
            }
            syncwarp();
            // ---- Round C: 2x2 input block, inertia check, inverse (all lanes redundantly)
            const double faa = sm[W_F + 28], fad = sm[W_F + 29], fdd = sm[W_F + 35];
            const double det = faa * fdd - fad * fad;
            if (!(faa > 0.0) || !(det > 0.0)) { ok = false; break; }
            const double idet = 1.0 / det;
            const double i00 = fdd * idet, i01 = -fad * idet, i11 = faa * idet;
            // ---- Round D: gains K = -Fuu^{-1} [F_u,xi | f_u]
            {
                const double* x = sm + R.d_xb + (R.d_xk ? so : 0);
                const double x0 = x[0], x1 = x[1];
                const double g = (R.d_r == 0) ? -(i00 * x0 + i01 * x1) : -(i01 * x0 + i11 * x1);
                if (R.d_out >= 0) kst[s * KST_STRIDE + R.d_out] = g;
            }
            syncwarp();
            // ---- Round E: P <- F_xixi + F_xi,u K ; p <- f_xi + F_xi,u k   (not needed for s = 0)
            if (s > 0) {
                const double f0 = sm[R.e_f0 + (R.e_f0k ? so : 0)];
                const double* f1 = sm + R.e_f1 + (R.e_f1k ? so : 0);
                const double* kk = kst + s * KST_STRIDE + R.e_kb;
                const double o = f0 + (f1[0] * kk[0] + f1[1] * kk[7]);
                sm[R.e_o1] = o; sm[R.e_o2] = o;
                syncwarp();
            }
        }
        return ok;
    }

    // forward sweep (every lane runs the scalar recursion; lane k keeps stage k)
    MPC_DEV void riccati_forward(const double* ds0) {
        const double* kst = sm + W_SD + (N + 1) * SDS;
        double s0 = shfl(ds0[0], 0), s1 = shfl(ds0[1], 0), s2 = shfl(ds0[2], 0), s3 = shfl(ds0[3], 0);
        double pa = 0.0, pd = 0.0;
        D.dsx = D.dsy = D.dsp = D.dsv = D.dua = D.dud = 0.0;
        for (int s = 0; s < N; s++) {
            const double* K = kst + s * KST_STRIDE;
            const double* r = sm + W_SD + s * SDS;
            const double ua = (K[0] * s0 + K[1] * s1 + K[2] * s2) + (K[3] * s3 + K[4] * pa + K[5] * pd) + K[6];
            const double ud = (K[7] * s0 + K[8] * s1 + K[9] * s2) + (K[10] * s3 + K[11] * pa + K[12] * pd) + K[13];
            if (s == k) { D.dsx = s0; D.dsy = s1; D.dsp = s2; D.dsv = s3; D.dua = ua; D.dud = ud; }
            // next state: columns psi (0..5), v (6..11), a (12..17), df (18..23), r (24..29)
            const double n0 = s0 + r[SD_MT + 0] * s2 + r[SD_MT + 6] * s3 + r[SD_MT + 18] * ud + r[SD_MT + 24];
            const double n1 = s1 + r[SD_MT + 1] * s2 + r[SD_MT + 7] * s3 + r[SD_MT + 19] * ud + r[SD_MT + 25];
            const double n2 = s2 + r[SD_MT + 8] * s3 + r[SD_MT + 20] * ud + r[SD_MT + 26];
            const double n3 = s3 + r[SD_MT + 15] * ua + r[SD_MT + 27];
            s0 = n0; s1 = n1; s2 = n2; s3 = n3; pa = ua; pd = ud;
        }
        if (k == N) { D.dsx = s0; D.dsy = s1; D.dsp = s2; D.dsv = s3; }
    }

    // new equality multipliers from stationarity in the state variables (parallel suffix sums),
    // new rate-row multipliers and slack steps.  rdr = rate-row residual used in the solve.
    MPC_DEV void recover_duals(const Cond& q, const double* rdr) {
        // local_k = -(H_ss ds + H_su du + g_s)
        double lx = 0.0, ly = 0.0, lp = 0.0, lv = 0.0;
        if (isS) {
            lx = -(q.Hxx * D.dsx + q.gsx);
            ly = -(q.Hyy * D.dsy + q.gsy);
            lp = -(q.Hpp * D.dsp + q.Hpv * D.dsv + q.Hpd * D.dud + q.gsp);
            lv = -(q.Hpv * D.dsp + q.Hvv * D.dsv + q.Hvd * D.dud + q.gsv);
        }
        // y_k = local_k + A_k' y_{k+1}
        const double* r = sm + W_SD + (isS ? k : 0) * SDS;
        const double A02 = r[SD_MT + 0], A12 = r[SD_MT + 1], A03 = r[SD_MT + 6], A13 = r[SD_MT + 7], A23 = r[SD_MT + 8];
        D.nyx = warp_suffix_sum(lx);
        D.nyy = warp_suffix_sum(ly);
        const double nx1 = shfl_down(D.nyx, 1), ny1 = shfl_down(D.nyy, 1);
        const double hasn = isU ? 1.0 : 0.0;
        D.nyp = warp_suffix_sum(lp + hasn * (A02 * nx1 + A12 * ny1));
        const double np1 = shfl_down(D.nyp, 1);
        D.nyv = warp_suffix_sum(lv + hasn * (A03 * nx1 + A13 * ny1 + A23 * np1));
        // rate rows
        const double pa = shfl_up(D.dua, 1), pd = shfl_up(D.dud, 1);
        D.drs[0] = D.drs[1] = 0.0; D.nyd[0] = D.nyd[1] = 0.0;
        if (isR) {
            const double ba = (k == 0) ? 0.0 : pa, bd = (k == 0) ? 0.0 : pd;
            D.drs[0] = (D.dud - bd) + rdr[0];
            D.drs[1] = (D.dua - ba) + rdr[1];
            D.nyd[0] = q.SrW[0] * D.drs[0] + q.br[0];
            D.nyd[1] = q.SrW[1] * D.drs[1] + q.br[1];
        }
    }

    // fraction-to-the-boundary for the primal step
    MPC_DEV double alpha_primal(double tau) const {
        double a = 1.0;
        if (isS) {
            if (D.dsv < 0.0) a = dmin_(a, -tau * (L.sv - vLo) / D.dsv);
            if (D.dsv > 0.0) a = dmin_(a, tau * (vHi - L.sv) / D.dsv);
        }
        if (isU) {
            if (D.dua < 0.0) a = dmin_(a, -tau * (L.ua - aLo) / D.dua);
            if (D.dua > 0.0) a = dmin_(a, tau * (aHi - L.ua) / D.dua);
            if (D.dud < 0.0) a = dmin_(a, -tau * (L.ud - dLo) / D.dud);
            if (D.dud > 0.0) a = dmin_(a, tau * (dHi - L.ud) / D.dud);
        }
        if (isR) for (int i = 0; i < 2; i++) {
            if (D.drs[i] < 0.0) a = dmin_(a, -tau * (L.rs[i] - rLo[i]) / D.drs[i]);
            if (D.drs[i] > 0.0) a = dmin_(a, tau * (rHi[i] - L.rs[i]) / D.drs[i]);
        }
        return warp_min(a);
    }

    // bound-multiplier steps: dz = mu/sl - z - (z/sl) dx  (lower),  mu/su - z + (z/su) dx (upper)
    MPC_DEV static double dzl(double mu, double z, double sl, double dx) { return mu / sl - z - z / sl * dx; }
    MPC_DEV static double dzu(double mu, double z, double su, double dx) { return mu / su - z + z / su * dx; }

    // ------------------------------------------------------------------
    // the whole solve
    // ------------------------------------------------------------------
    struct Result { int status; int iters; double cost; };

    MPC_DEV bool nlp_feasible() const {
        const double e = 1e-8;
        const double lim_d = c.sdmax * c.dtc, lim_a = c.admax * c.dtc;
        if (st0[3] < c.vmin - e * dmax_(1.0, fabs(c.vmin))) return false;
        if (st0[3] > c.vmax + e * dmax_(1.0, fabs(c.vmax))) return false;
        if (uprev[0] - lim_d > c.smax + 2 * e || uprev[0] + lim_d < -c.smax - 2 * e) return false;
        if (uprev[1] - lim_a > c.amax + 2 * e || uprev[1] + lim_a < -c.amax - 2 * e) return false;
        return true;
    }

    MPC_DEV Result solve() {
        Result res; res.status = 4; res.iters = 0; res.cost = 0.0;
        // ---- bounds, relaxed by bound_relax_factor
        vLo = c.vmin - K_BOUND_RELAX * dmax_(1.0, fabs(c.vmin)); vHi = c.vmax + K_BOUND_RELAX * dmax_(1.0, fabs(c.vmax));
        aLo = -c.amax - K_BOUND_RELAX * dmax_(1.0, c.amax); aHi = c.amax + K_BOUND_RELAX * dmax_(1.0, c.amax);
        dLo = -c.smax - K_BOUND_RELAX * dmax_(1.0, c.smax); dHi = c.smax + K_BOUND_RELAX * dmax_(1.0, c.smax);
        {
            const double h = (k == 0) ? c.dtc : c.dt;
            const double ld = c.sdmax * h, la = c.admax * h;
            rLo[0] = -ld - K_BOUND_RELAX * dmax_(1.0, ld); rHi[0] = ld + K_BOUND_RELAX * dmax_(1.0, ld);
            rLo[1] = -la - K_BOUND_RELAX * dmax_(1.0, la); rHi[1] = la + K_BOUND_RELAX * dmax_(1.0, la);
        }
        wx = (k >= 1 && k <= N) ? c.w[0] : 0.0;
        wy = (k >= 1 && k <= N) ? c.w[1] : 0.0;
        wp = (k >= 1 && k <= N) ? c.w[2] : 0.0;
        wv = (k >= 1 && k <= N - 1) ? c.w[3] : 0.0;

        if (!nlp_feasible()) {
            res.status = 1; res.iters = 0;
            res.cost = objective(L.sx, L.sy, L.sp, L.sv, L.ua, L.ud);
            return res;
        }

        // ---- objective scaling from the gradient at the user's start point
        sigma = 1.0;
        objective_gradient();
        {
            double gm = dmax_(dmax_(fabs(gx), fabs(gy)), dmax_(fabs(gp), fabs(gv)));
            gm = dmax_(gm, dmax_(fabs(ga), fabs(gd)));
            gm = warp_max(gm);
            sigma = (gm > K_SCALE_MAX_GRAD) ? dmax_(K_SCALE_MAX_GRAD / gm, 1e-8) : 1.0;
        }
        // ---- push into the interior, slacks, bound multipliers
        if (isS) push_interior(L.sv, vLo, vHi);
        if (isU) { push_interior(L.ua, aLo, aHi); push_interior(L.ud, dLo, dHi); }
        {
            double zero2[2] = {0.0, 0.0};
            eval_defects(L.sx, L.sy, L.sp, L.sv, L.ua, L.ud, zero2, tg, rdyn, rinit, dres);
            L.rs[0] = dres[0]; L.rs[1] = dres[1];
            if (isR) { push_interior(L.rs[0], rLo[0], rHi[0]); push_interior(L.rs[1], rLo[1], rHi[1]); }
        }
        L.zvL = L.zvU = isS ? 1.0 : 0.0;
        L.zaL = L.zaU = L.zdL = L.zdU = isU ? 1.0 : 0.0;
        L.rvL[0] = L.rvL[1] = L.rvU[0] = L.rvU[1] = isR ? 1.0 : 0.0;
        L.ryd[0] = L.ryd[1] = 0.0;
        L.yx = L.yy = L.yp = L.yv = 0.0;

        Cond q;
        // ---- least-squares equality multipliers
        {
            objective_gradient();
            eval_defects(L.sx, L.sy, L.sp, L.sv, L.ua, L.ud, L.rs, tg, rdyn, rinit, dres);
            build_cond_ls(q);
            const double zr[4] = {0.0, 0.0, 0.0, 0.0};
            const double zd[2] = {0.0, 0.0};
            if (isS) write_record(q, zr);
            syncwarp();
            const bool ok = riccati_backward();
            bool use = ok;
            if (ok) {
                riccati_forward(zr);
                recover_duals(q, zd);
                double ym = dmax_(dmax_(fabs(D.nyx), fabs(D.nyy)), dmax_(fabs(D.nyp), fabs(D.nyv)));
                ym = dmax_(ym, dmax_(fabs(D.nyd[0]), fabs(D.nyd[1])));
                ym = isS ? ym : 0.0;
                ym = warp_max(ym);
                use = (ym <= K_Y_INIT_MAX);
            }
            if (use && isS) { L.yx = D.nyx; L.yy = D.nyy; L.yp = D.nyp; L.yv = D.nyv; L.ryd[0] = D.nyd[0]; L.ryd[1] = D.nyd[1]; }
            syncwarp();
        }

        double mu = K_MU_INIT, tau = dmax_(K_TAU_MIN, 1.0 - K_MU_INIT);
        const double mu_min = dmin_(c.tol, 1e-4) / (K_KAPPA_EPS + 1.0);
        double dw_last = 0.0, theta_max = -1.0, theta_min = -1.0;
        double f_phi = 0.0, f_theta = 0.0;  // this lane's filter entry (entry e lives in lane e)
        int nfilt = 0, accept_count = 0, iter = 0, ret = -1;
        bool tiny_last = false;
        const int nz = 2 * (3 * N + 1) + 4 * (N - 1);  // bound multipliers: x-bounds + slack bounds
        const int my = 4 * (N + 1) + 2 * (N - 1);      // equality multipliers y_c, y_d

        for (;;) {
            // ---- evaluate at the current iterate
            objective_gradient();
            eval_defects(L.sx, L.sy, L.sp, L.sv, L.ua, L.ud, L.rs, tg, rdyn, rinit, dres);
            const double theta = theta_of(rdyn, rinit, dres);
            // stage Jacobian entries (also used for the record)
            const double A02 = isU ? -c.dt * L.sv * tg.sn : 0.0, A03 = isU ? c.dt * tg.cs : 0.0;
            const double A12 = isU ? c.dt * L.sv * tg.cs : 0.0, A13 = isU ? c.dt * tg.sn : 0.0;
            const double A23 = isU ? c.dt * tg.sb / c.Lb : 0.0;
            const double b0 = A02 * tg.b1, b1v = A12 * tg.b1, b2v = isU ? c.dt * L.sv * tg.cb * tg.b1 / c.Lb : 0.0;
            // dual infeasibility of this stage's variables
            double di, cv, sumy, sumz;
            {
                const double y1x = shfl_down(L.yx, 1), y1y = shfl_down(L.yy, 1), y1p = shfl_down(L.yp, 1), y1v = shfl_down(L.yv, 1);
                const double yd1_0 = shfl_down(L.ryd[0], 1), yd1_1 = shfl_down(L.ryd[1], 1);
                const double hn = isU ? 1.0 : 0.0;
                const double glx = gx + L.yx - hn * y1x;
                const double gly = gy + L.yy - hn * y1y;
                const double glp = gp + L.yp - hn * (A02 * y1x + A12 * y1y + y1p);
                const double glv = gv + L.yv - hn * (A03 * y1x + A13 * y1y + A23 * y1p + y1v) - L.zvL + L.zvU;
                const double nd0 = (k + 1 < N) ? yd1_0 : 0.0, nd1 = (k + 1 < N) ? yd1_1 : 0.0;
                const double gla = ga - hn * c.dt * y1v + (L.ryd[1] - nd1) - L.zaL + L.zaU;
                const double gld = gd - hn * (b0 * y1x + b1v * y1y + b2v * y1p) + (L.ryd[0] - nd0) - L.zdL + L.zdU;
                di = dmax_(dmax_(fabs(glx), fabs(gly)), dmax_(fabs(glp), fabs(glv)));
                di = dmax_(di, dmax_(fabs(gla), fabs(gld)));
                di = dmax_(di, dmax_(fabs(-L.ryd[0] - L.rvL[0] + L.rvU[0]), fabs(-L.ryd[1] - L.rvL[1] + L.rvU[1])));
                di = isS ? di : 0.0;
                cv = dmax_(dmax_(fabs(rdyn[0]), fabs(rdyn[1])), dmax_(fabs(rdyn[2]), fabs(rdyn[3])));
                cv = dmax_(cv, dmax_(dmax_(fabs(rinit[0]), fabs(rinit[1])), dmax_(fabs(rinit[2]), fabs(rinit[3]))));
                cv = dmax_(cv, dmax_(fabs(dres[0]), fabs(dres[1])));
                sumy = isS ? fabs(L.yx) + fabs(L.yy) + fabs(L.yp) + fabs(L.yv) + fabs(L.ryd[0]) + fabs(L.ryd[1]) : 0.0;
                sumz = L.zvL + L.zvU + L.zaL + L.zaU + L.zdL + L.zdU + L.rvL[0] + L.rvL[1] + L.rvU[0] + L.rvU[1];
                di = warp_max(di); cv = warp_max(cv); sumy = warp_sum(sumy); sumz = warp_sum(sumz);
            }
            const double sd = dmax_(K_S_MAX, (sumy + sumz) / (double)(my + nz)) / K_S_MAX;
            const double sc = dmax_(K_S_MAX, sumz / (double)nz) / K_S_MAX;
            // complementarity for target t: max |sl*z - t|
            auto compl_err = [&](double t) {
                double m = 0.0;
                if (isS) { m = dmax_(m, fabs((L.sv - vLo) * L.zvL - t)); m = dmax_(m, fabs((vHi - L.sv) * L.zvU - t)); }
                if (isU) {
                    m = dmax_(m, fabs((L.ua - aLo) * L.zaL - t)); m = dmax_(m, fabs((aHi - L.ua) * L.zaU - t));
                    m = dmax_(m, fabs((L.ud - dLo) * L.zdL - t)); m = dmax_(m, fabs((dHi - L.ud) * L.zdU - t));
                }
                if (isR) for (int i = 0; i < 2; i++) {
                    m = dmax_(m, fabs((L.rs[i] - rLo[i]) * L.rvL[i] - t)); m = dmax_(m, fabs((rHi[i] - L.rs[i]) * L.rvU[i] - t));
                }
                return warp_max(m);
            };
            // ---- convergence
            {
                const double cm0 = compl_err(0.0);
                const double E0 = dmax_(dmax_(di / sd, cv), cm0 / sc);
                const double du = di / sigma, mc = cm0 / sigma;
                if (E0 <= c.tol && du <= 1.0 && cv <= 1e-4 && mc <= 1e-4) { ret = 0; break; }
                if (E0 <= K_ACCEPT_TOL && du <= 1e10 && cv <= 1e-2 && mc <= 1e-2) {
                    if (++accept_count >= K_ACCEPT_ITER) { ret = 1; break; }
                } else accept_count = 0;
            }
            if (iter >= c.max_iter) { ret = -1; break; }
            {
                double xm = dmax_(dmax_(fabs(L.sx), fabs(L.sy)), dmax_(fabs(L.sp), fabs(L.sv)));
                xm = isS ? xm : 0.0;
                xm = warp_max(xm);
                if (!(xm <= 1e20)) { ret = -5; break; }
            }
            // ---- monotone barrier update
            for (;;) {
                const double cm = compl_err(mu);
                const double Emu = dmax_(dmax_(di / sd, cv), cm / sc);
                if (!(Emu <= K_KAPPA_EPS * mu) && !tiny_last) break;
                const double nm = dmax_(mu_min, dmin_(K_KAPPA_MU * mu, pow(mu, K_THETA_MU)));
                if (nm >= mu) { if (tiny_last) ret = -3; break; }
                mu = nm; tau = dmax_(K_TAU_MIN, 1.0 - mu);
                nfilt = 0;
                if (tiny_last) { tiny_last = false; break; }
            }
            if (ret == -3) break;

            // ---- primal-dual system: assemble, factorise with inertia correction, solve
            double dw = 0.0;
            {
                bool ok = false;
                const double rr[4] = {rdyn[0], rdyn[1], rdyn[2], rdyn[3]};
                build_cond_pd(q, mu, dw, dres);
                if (isS) write_record(q, rr);
                syncwarp();
                for (;;) {
                    ok = riccati_backward();
                    if (ok) break;
                    if (dw == 0.0) dw = (dw_last == 0.0) ? K_DW_INIT : dmax_(K_DW_MIN, dw_last * K_DW_DEC);
                    else dw = (dw_last == 0.0 || 1e5 * dw_last < dw) ? K_DW_INC_FIRST * dw : K_DW_INC * dw;
                    if (dw > K_DW_MAX) break;
                    syncwarp();
                    build_cond_pd(q, mu, dw, dres);
                    if (isS) patch_record_diag(q);
                    syncwarp();
                }
                if (!ok) { ret = -4; break; }
                if (dw > 0.0) dw_last = dw;
            }
            riccati_forward(rinit);
            recover_duals(q, dres);

            // ---- line search quantities at the current point
            const double fcur = objective(L.sx, L.sy, L.sp, L.sv, L.ua, L.ud);
            const double phi = sigma * fcur - mu * log_barrier(L.sv, L.ua, L.ud, L.rs);
            double gBd;
            {
                double t = gx * D.dsx + gy * D.dsy + gp * D.dsp + ga * D.dua + gd * D.dud;
                if (isS) t += (gv - mu / (L.sv - vLo) + mu / (vHi - L.sv)) * D.dsv;
                if (isU) {
                    t += (-mu / (L.ua - aLo) + mu / (aHi - L.ua)) * D.dua;
                    t += (-mu / (L.ud - dLo) + mu / (dHi - L.ud)) * D.dud;
                }
                if (isR) t += q.br[0] * D.drs[0] + q.br[1] * D.drs[1];
                t = isS ? t : 0.0;
                gBd = warp_sum(t);
            }
            if (theta_max < 0.0) { theta_max = 1e4 * dmax_(1.0, theta); theta_min = 1e-4 * dmax_(1.0, theta); }
            double alpha_max = alpha_primal(tau);
            double alpha_min = K_GAMMA_THETA;
            if (gBd < 0.0) {
                alpha_min = dmin_(K_GAMMA_THETA, K_GAMMA_PHI * theta / (-gBd));
                if (theta <= theta_min) alpha_min = dmin_(alpha_min, K_DELTA * pow(theta, K_S_THETA) / pow(-gBd, K_S_PHI));
            }
            alpha_min *= K_ALPHA_MIN_FRAC;
            bool tiny;
            {
                double m = 0.0;
                if (isS) {
                    m = dmax_(dmax_(fabs(D.dsx) / (1.0 + fabs(L.sx)), fabs(D.dsy) / (1.0 + fabs(L.sy))),
                              dmax_(fabs(D.dsp) / (1.0 + fabs(L.sp)), fabs(D.dsv) / (1.0 + fabs(L.sv))));
                    m = dmax_(m, dmax_(fabs(D.dua) / (1.0 + fabs(L.ua)), fabs(D.dud) / (1.0 + fabs(L.ud))));
                    m = dmax_(m, dmax_(fabs(D.drs[0]) / (1.0 + fabs(L.rs[0])), fabs(D.drs[1]) / (1.0 + fabs(L.rs[1]))));
                }
                m = warp_max(m);
                tiny = (m < 10.0 * K_EPS) && (theta < 1e-4);
            }

            // trial point storage
            double tx, ty, tp, tv, ta, td, ts[2];
            double alpha = alpha_max;
            bool accepted = false, ftype_arm = false;
            const double pw_gbd = (gBd < 0.0) ? pow(-gBd, K_S_PHI) : 0.0;
            const double pw_th = K_DELTA * pow(theta, K_S_THETA);
            auto make_trial = [&](double a) {
                tx = L.sx + a * D.dsx; ty = L.sy + a * D.dsy; tp = L.sp + a * D.dsp; tv = L.sv + a * D.dsv;
                ta = L.ua + a * D.dua; td = L.ud + a * D.dud; ts[0] = L.rs[0] + a * D.drs[0]; ts[1] = L.rs[1] + a * D.drs[1];
            };
            auto is_ftype = [&](double a) { return gBd < 0.0 && a * pw_gbd > pw_th; };
            auto armijo = [&](double a, double pht) { return cmp_le(pht - phi, K_ETA_PHI * a * gBd, phi); };
            auto accept_test = [&](double a, double tht, double pht) {
                bool ok = (tht == tht) && (pht == pht);
                if (ok && !cmp_le(tht, theta_max, theta)) ok = false;
                if (ok) {
                    if (is_ftype(a) && theta <= theta_min) ok = armijo(a, pht);
                    else {
                        if (pht > phi) {
                            double bas = 1.0; if (fabs(phi) > 10.0) bas = log10(fabs(phi));
                            if (log10(pht - phi) > K_OBJ_MAX_INC + bas) ok = false;
                        }
                        if (ok) ok = cmp_le(tht, (1.0 - K_GAMMA_THETA) * theta, theta) || cmp_le(pht - phi, -K_GAMMA_PHI * theta, phi);
                    }
                }
                // filter: entry e lives in lane e
                const bool mine = (k >= nfilt) || cmp_le(pht, f_phi, f_phi) || cmp_le(tht, f_theta, f_theta);
                const bool fok = warp_all(mine);
                return ok && fok;
            };
            StageTrig ttg;
            double trd[4], tri[4], tdr[4];
            if (tiny) { make_trial(alpha); accepted = true; }
            int nsteps = 0;
            while (!accepted && (alpha > alpha_min || nsteps == 0)) {
                make_trial(alpha);
                eval_defects(tx, ty, tp, tv, ta, td, ts, ttg, trd, tri, tdr);
                const double th_t = theta_of(trd, tri, tdr);
                const double ph_t = sigma * objective(tx, ty, tp, tv, ta, td) - mu * log_barrier(tv, ta, td, ts);
                if (accept_test(alpha, th_t, ph_t)) { accepted = true; ftype_arm = is_ftype(alpha) && armijo(alpha, ph_t); break; }
                // ---- second-order correction on the first trial
                if (nsteps == 0 && theta <= th_t && K_MAX_SOC > 0) {
                    double csoc[4] = {rdyn[0], rdyn[1], rdyn[2], rdyn[3]};
                    double isoc[4] = {rinit[0], rinit[1], rinit[2], rinit[3]};
                    double dsoc[2] = {dres[0], dres[1]};
                    double th_old = 0.0, th_soc = th_t, a_soc = alpha;
                    const StepState D0 = D;
                    int cnt = 0;
                    while (cnt < K_MAX_SOC && !accepted && (cnt == 0 || th_soc <= K_KAPPA_SOC * th_old)) {
                        th_old = th_soc;
                        for (int i = 0; i < 4; i++) { csoc[i] = a_soc * csoc[i] + trd[i]; isoc[i] = a_soc * isoc[i] + tri[i]; }
                        for (int i = 0; i < 2; i++) dsoc[i] = a_soc * dsoc[i] + tdr[i];
                        syncwarp();
                        build_cond_pd(q, mu, dw, dsoc);
                        if (isS) patch_record_rhs(q, csoc);
                        syncwarp();
                        (void)riccati_backward();
                        riccati_forward(isoc);
                        recover_duals(q, dsoc);
                        a_soc = alpha_primal(tau);
                        make_trial(a_soc);
                        eval_defects(tx, ty, tp, tv, ta, td, ts, ttg, trd, tri, tdr);
                        th_soc = theta_of(trd, tri, tdr);
                        const double ph_s = sigma * objective(tx, ty, tp, tv, ta, td) - mu * log_barrier(tv, ta, td, ts);
                        if (accept_test(alpha, th_soc, ph_s)) {
                            accepted = true; ftype_arm = is_ftype(alpha) && armijo(alpha, ph_s);
                            alpha = a_soc;
                        } else cnt++;
                    }
                    if (accepted) break;
                    D = D0;  // back to the uncorrected direction for the backtracking steps
                    // the record's right-hand side is restored on the next assembly
                }
                alpha *= 0.5; nsteps++;
            }
            if (!accepted) { ret = -2; break; }  // Ipopt would enter the restoration phase

            // ---- filter augmentation
            if (!tiny && !ftype_arm) {
                if (nfilt < 32) {
                    if (k == nfilt) { f_phi = phi - K_GAMMA_PHI * theta; f_theta = (1.0 - K_GAMMA_THETA) * theta; }
                    nfilt++;
                }
            }
            // ---- accept the trial point: duals
            double az = 1.0;
            {
                // bound multiplier steps with the (possibly corrected) primal step
                double dzvL = 0, dzvU = 0, dzaL = 0, dzaU = 0, dzdL = 0, dzdU = 0, drvL[2] = {0, 0}, drvU[2] = {0, 0};
                if (isS) { dzvL = dzl(mu, L.zvL, L.sv - vLo, D.dsv); dzvU = dzu(mu, L.zvU, vHi - L.sv, D.dsv); }
                if (isU) {
                    dzaL = dzl(mu, L.zaL, L.ua - aLo, D.dua); dzaU = dzu(mu, L.zaU, aHi - L.ua, D.dua);
                    dzdL = dzl(mu, L.zdL, L.ud - dLo, D.dud); dzdU = dzu(mu, L.zdU, dHi - L.ud, D.dud);
                }
                if (isR) for (int i = 0; i < 2; i++) {
                    drvL[i] = dzl(mu, L.rvL[i], L.rs[i] - rLo[i], D.drs[i]); drvU[i] = dzu(mu, L.rvU[i], rHi[i] - L.rs[i], D.drs[i]);
                }
                auto lim = [&](double a, double z, double dz) { return (dz < 0.0) ? dmin_(a, -tau * z / dz) : a; };
                az = lim(az, L.zvL, dzvL); az = lim(az, L.zvU, dzvU); az = lim(az, L.zaL, dzaL); az = lim(az, L.zaU, dzaU);
                az = lim(az, L.zdL, dzdL); az = lim(az, L.zdU, dzdU);
                az = lim(az, L.rvL[0], drvL[0]); az = lim(az, L.rvU[0], drvU[0]); az = lim(az, L.rvL[1], drvL[1]); az = lim(az, L.rvU[1], drvU[1]);
                az = warp_min(az);
                // primal
                L.sx = tx; L.sy = ty; L.sp = tp; L.sv = tv; L.ua = ta; L.ud = td; L.rs[0] = ts[0]; L.rs[1] = ts[1];
                // equality multipliers with the primal step size
                L.yx += alpha * (D.nyx - L.yx); L.yy += alpha * (D.nyy - L.yy); L.yp += alpha * (D.nyp - L.yp); L.yv += alpha * (D.nyv - L.yv);
                L.ryd[0] += alpha * (D.nyd[0] - L.ryd[0]); L.ryd[1] += alpha * (D.nyd[1] - L.ryd[1]);
                auto upd = [&](double& z, double dz, double sl) {
                    z += az * dz;
                    z = dmax_(dmin_(z, K_KAPPA_SIGMA * mu / sl), mu / (K_KAPPA_SIGMA * sl));
                };
                if (isS) { upd(L.zvL, dzvL, L.sv - vLo); upd(L.zvU, dzvU, vHi - L.sv); }
                if (isU) {
                    upd(L.zaL, dzaL, L.ua - aLo); upd(L.zaU, dzaU, aHi - L.ua);
                    upd(L.zdL, dzdL, L.ud - dLo); upd(L.zdU, dzdU, dHi - L.ud);
                }
                if (isR) for (int i = 0; i < 2; i++) { upd(L.rvL[i], drvL[i], L.rs[i] - rLo[i]); upd(L.rvU[i], drvU[i], rHi[i] - L.rs[i]); }
            }
            tiny_last = tiny;
            iter++;
            syncwarp();
        }
        res.iters = iter;
        res.status = (ret == 0 || ret == 1) ? 0 : (ret == -1 ? 3 : (ret == -5 ? 2 : 4));
        // honor_original_bounds
        if (isS) L.sv = dmin_(dmax_(L.sv, c.vmin), c.vmax);
        if (isU) { L.ua = dmin_(dmax_(L.ua, -c.amax), c.amax); L.ud = dmin_(dmax_(L.ud, -c.smax), c.smax); }
        res.cost = objective(L.sx, L.sy, L.sp, L.sv, L.ua, L.ud);
        return res;
    }
};

// Problem I/O for one warp: problem-major layouts of include/mpc_b200.h.
struct BatchPtrs {
    const double* state;   // [B][4]
    const double* ref;     // [B][3][N+1]
    const double* v_des;   // [B] or null
    const double* u_prev;  // [B][2]
    double* warm;          // [B][6N+4] or null
    double* u0;            // [B][2]
    double* cost;          // [B] or null
    int* status;           // [B] or null
    int* iters;            // [B] or null
    double* traj;          // [B][6N+4] or null
};

MPC_DEV void solve_problem(const KCfg& cfg, const BatchPtrs& io, long b, double* smem) {
    WarpSolver S(cfg, smem);
    const int k = S.k, N = cfg.N;
    const long nr = 3L * (N + 1), nt = 6L * N + 4;
    S.init_roles();
    // ---- coalesced loads: lane k takes stage k's reference sample; lanes 0..3 the state
    const double* rf = io.ref + nr * b;
    S.xr = (k <= N) ? rf[k] : 0.0;
    S.yr = (k <= N) ? rf[(N + 1) + k] : 0.0;
    S.pr = (k <= N) ? rf[2 * (N + 1) + k] : 0.0;
    {
        const double sv = (k < 4) ? io.state[4 * b + k] : 0.0;
        for (int i = 0; i < 4; i++) S.st0[i] = shfl(sv, i);
        const double uv = (k < 2) ? io.u_prev[2 * b + k] : 0.0;
        S.uprev[0] = shfl(uv, 0); S.uprev[1] = shfl(uv, 1);
    }
    S.vdes = io.v_des ? io.v_des[b] : 0.0;
    // ---- start point
    S.L.sx = S.L.sy = S.L.sp = S.L.sv = S.L.ua = S.L.ud = 0.0;
    if (io.warm) {
        const double* w = io.warm + nt * b;
        if (k <= N) { S.L.sx = w[k]; S.L.sy = w[(N + 1) + k]; S.L.sv = w[2 * (N + 1) + k]; S.L.sp = w[3 * (N + 1) + k]; }
        if (k < N) { S.L.ud = w[4 * (N + 1) + k]; S.L.ua = w[4 * (N + 1) + N + k]; }
    }
    WarpSolver::Result r = S.solve();
    // ---- results
    const double a0 = shfl(S.L.ua, 0), d0 = shfl(S.L.ud, 0);
    if (k == 0) {
        io.u0[2 * b] = a0; io.u0[2 * b + 1] = d0;
        if (io.cost) io.cost[b] = r.cost;
        if (io.status) io.status[b] = r.status;
        if (io.iters) io.iters[b] = r.iters;
    }
    for (int pass = 0; pass < 2; pass++) {
        double* t = (pass == 0) ? io.traj : io.warm;
        if (!t) continue;
        t += nt * b;
        if (k <= N) { t[k] = S.L.sx; t[(N + 1) + k] = S.L.sy; t[2 * (N + 1) + k] = S.L.sv; t[3 * (N + 1) + k] = S.L.sp; }
        if (k < N) { t[4 * (N + 1) + k] = S.L.ud; t[4 * (N + 1) + N + k] = S.L.ua; }
    }
}

}  // namespace mpcb200
