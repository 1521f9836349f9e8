/*
 * mpc_oracle.h -- CPU oracle for the kinematic-bicycle path-following NLP.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by
 * or executed from the shipped library (libmpc_b200.so) or the host package;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker / CPU baseline.
 *
 * PARITY UNPINNED: the reference delegates the arithmetic of this path to
 * un-vendored third-party code (JuMP <= 0.18 + Ipopt.jl + Ipopt 3.12.x/MUMPS,
 * versions unpinned, reference README.md:16-24) and holds no golden vectors,
 * known-answer tests or recorded solver outputs for it.  Neither Julia nor any
 * Ipopt build is reachable in this environment.  This file restates
 *   (1) the NLP exactly as scripts/mpc_utils/MKZMPCPathFollower.jl:28-123
 *       declares it, and
 *   (2) the published Ipopt algorithm (Waechter & Biegler, Math. Prog. 106,
 *       2006) with the 3.12 default options, as recalled;
 * and is validated intrinsically (finite-difference derivative checks, KKT
 * residuals of the unscaled NLP, agreement with scipy SLSQP/trust-constr).
 * One piece of Ipopt is NOT restated: its restoration phase.  Where the filter
 * line search fails, the iterate is replaced by the rollout of its own projected
 * inputs and the iteration is re-initialised there (ipm_rollout_restore; the
 * CUDA kernel does the same, step for step).  MPC_ORACLE_NO_RESTO=1 switches it
 * off (status Error instead), MPC_ORACLE_TRACE=1 prints the iteration log.
 */
#ifndef MPC_ORACLE_H
#define MPC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int    N;            /* horizon,                MKZMPCPathFollower.jl:34 */
    double dt;           /* model step td,          :33 */
    double dt_control;   /* control period ts,      :28 */
    double L_a, L_b;     /* CoG->front / rear axle, :31-32 */
    double v_min, v_max; /* :47-48 */
    double a_max;        /* :44 */
    double steer_max;    /* :41 */
    double a_dmax;       /* :45 */
    double steer_dmax;   /* :42 */
    /* update_cost order (:158-169): C_x, C_y, C_psi, C_v, C_dacc, C_ddf, C_acc, C_df */
    double w[8];
    /* Ipopt options the reference leaves at default, plus the iteration cap that
     * stands in for max_cpu_time (:29), which is not reproducible. */
    double tol;          /* 1e-8 */
    int    max_iter;     /* cap */
    /* model 0: XY kinematic bicycle (MKZMPCPathFollower.jl).  model 1: the Frenet-frame variant
     * (MKZMPCPathFollowerFrenet.jl): states (s, e_y, e_psi, v) in the slots of (x, y, psi, v),
     * curvature K(s) = kpoly[0] s^3 + kpoly[1] s^2 + kpoly[2] s + kpoly[3] (:40-41, :111), cost on
     * e_y, e_psi, v only (:95-101) = the XY cost with w = (0, C_ey, C_epsi, C_ev, ...) and a zero reference. */
    int    model;
    double kpoly[4];
} mpc_oracle_cfg;

/* status codes = JuMP symbols returned by solve_model (MKZMPCPathFollower.jl:176-182)
 * through Ipopt.jl's status mapping */
enum {
    MPC_ORACLE_OPTIMAL    = 0, /* Solve_Succeeded or Solved_To_Acceptable_Level */
    MPC_ORACLE_INFEASIBLE = 1, /* Infeasible_Problem_Detected */
    MPC_ORACLE_UNBOUNDED  = 2, /* Diverging_Iterates */
    MPC_ORACLE_USERLIMIT  = 3, /* Maximum_Iterations_Exceeded (stands in for Maximum_CpuTime_Exceeded) */
    MPC_ORACLE_ERROR      = 4  /* everything else (restoration needed/failed, tiny step, ...) */
};

void mpc_oracle_default_cfg(mpc_oracle_cfg* cfg, int N);

/* Diagnostics written per solve (all optional / may be NULL in batch call). */
typedef struct {
    double dual_inf, constr_viol, compl_inf; /* unscaled, at the returned point */
    double mu_final, obj_scale;
    int    n_inertia_corr;   /* factorizations that needed delta_w > 0 */
    int    n_soc;            /* accepted second-order corrections */
    int    n_backtrack;      /* total step halvings */
    int    ipopt_status;     /* 0 success, 1 acceptable, -1 maxiter, -2 restoration needed, -3 tiny step, -4 pert fail, 2 infeasible */
    int    n_resto;          /* restorations by rollout (see ipm_rollout_restore) */
} mpc_oracle_diag;

/*
 * One solve.  Layouts:
 *   state[4]  = x, y, psi, v        (update_init_cond order, :132)
 *   ref       = x_ref[N+1], y_ref[N+1], psi_ref[N+1]   (update_reference, :142)
 *   u_prev[2] = d_f_current, acc_current   (update_current_input order: steering first, :151)
 *   warm      = NULL (start=0.0, :65-72) or 6N+4 doubles in traj order
 *   traj[6N+4]= x[N+1], y[N+1], v[N+1], psi[N+1], d_f[N], acc[N]  (get_solver_results order, :188-206)
 *   u0[2]     = acc_opt[1], d_f_opt[1]  (solve_model return order, :182)
 *   mult      = NULL or 4+4N equality multipliers (init rows, then dynamics rows x,y,psi,v per stage)
 */
int mpc_oracle_solve(const mpc_oracle_cfg* cfg, const double* state, const double* ref,
                     double v_des, const double* u_prev, const double* warm,
                     double* traj, double* u0, double* cost, int* status, int* iters,
                     mpc_oracle_diag* diag);

/* The start point of MPCB200_START_ROLLOUT (include/mpc_b200.h): previous command held over the
 * horizon, projected onto box / rate rows / speed bounds, model rolled out from `state`. */
int mpc_oracle_rollout_start(const mpc_oracle_cfg* cfg, const double* state, const double* u_prev, double* traj);

/* Batch, problem-major layouts identical to include/mpc_b200.h; OpenMP over problems
 * when n_threads > 1.  v_des may be NULL (then 0), warm/traj may be NULL. */
int mpc_oracle_solve_batch(const mpc_oracle_cfg* cfg, long B, const double* state,
                           const double* ref, const double* v_des, const double* u_prev,
                           double* warm, double* u0, double* cost, int* status, int* iters,
                           double* traj, int n_threads);

/* the same, also reporting per problem how many restorations by rollout the solve went through (n_resto [B] or NULL) */
int mpc_oracle_solve_batch_resto(const mpc_oracle_cfg* cfg, long B, const double* state,
                                 const double* ref, const double* v_des, const double* u_prev,
                                 double* warm, double* u0, double* cost, int* status, int* iters,
                                 double* traj, int* n_resto, int n_threads);

/* Frenet-frame variant (MKZMPCPathFollowerFrenet.jl): cfg->model must be 1 (mpc_oracle_default_cfg_frenet);
 * state = (s0, ey0, epsi0, v0) (update_init_cond :132-138), kpoly = 4 curvature coefficients per problem,
 * highest degree first (update_reference :142-147); traj = s[N+1], ey[N+1], v[N+1], epsi[N+1], d_f[N], acc[N]
 * (get_solver_results :188-206).  cfg->kpoly is ignored here (overwritten per problem). */
void mpc_oracle_default_cfg_frenet(mpc_oracle_cfg* cfg, int N);
int mpc_oracle_solve_batch_frenet(const mpc_oracle_cfg* cfg, long B, const double* state,
                                  const double* kpoly, const double* v_des, const double* u_prev,
                                  double* warm, double* u0, double* cost, int* status, int* iters,
                                  double* traj, int n_threads);

/* NLP pieces exported for derivative / KKT tests.  z is in INTERNAL order:
 * stage-major (x,y,psi,v,acc,df) for k<N, then (x,y,psi,v) for k=N; n = 6N+4. */
int    mpc_oracle_nvar(const mpc_oracle_cfg* cfg);
int    mpc_oracle_ncon(const mpc_oracle_cfg* cfg);   /* 4 + 4N equality rows */
int    mpc_oracle_nrange(const mpc_oracle_cfg* cfg); /* 2(N-1) range rows   */
double mpc_oracle_eval_f(const mpc_oracle_cfg* cfg, const double* ref, double v_des, const double* z);
void   mpc_oracle_eval_grad_f(const mpc_oracle_cfg* cfg, const double* ref, double v_des, const double* z, double* g);
void   mpc_oracle_eval_c(const mpc_oracle_cfg* cfg, const double* state, const double* z, double* c);
void   mpc_oracle_eval_d(const mpc_oracle_cfg* cfg, const double* u_prev, const double* z, double* d);
/* dense row-major Jacobians: Jc is ncon x nvar, Jd is nrange x nvar */
void   mpc_oracle_eval_jac(const mpc_oracle_cfg* cfg, const double* z, double* Jc, double* Jd);
/* dense nvar x nvar Hessian of sigma*f + yc'c */
void   mpc_oracle_eval_hess(const mpc_oracle_cfg* cfg, const double* z, double sigma, const double* yc, double* H);
/* traj-order <-> internal-order helpers */
void   mpc_oracle_traj_to_z(const mpc_oracle_cfg* cfg, const double* traj, double* z);
void   mpc_oracle_z_to_traj(const mpc_oracle_cfg* cfg, const double* z, double* traj);

/* Reference generator (ref_gps_traj.py:131-218) and plant (vehicle_simulator.py:58-112)
 * restated in C for the closed-loop oracle. */
typedef struct {
    int n;
    const double *t, *X, *Y, *psi, *s; /* columns 0,4,5,3,6 of the trajectory table, ref_gps_traj.py:106 */
} mpc_oracle_path;

/* time mode (v_target < 0 means None) or distance mode; returns stop_cmd */
int mpc_oracle_get_waypoints(const mpc_oracle_path* p, int horizon, double traj_dt,
                             double X, double Y, double yaw, int use_vtarget, double v_target,
                             double* ref /* 3(N+1) */);

/* plant state: X,Y,psi,vx,vy,wz,acc,df ; one call = one 100 Hz publish period =
 * disc_steps(10) Euler sub-steps of 1 ms (vehicle_simulator.py:58-106) */
void mpc_oracle_plant_step(double* st /*8*/, double acc_des, double df_des);

/* Solution (traj order, 6N+4) of the module-load solve of the default problem (MKZMPCPathFollower.jl:36-39,126-128):
 * the start point of the node's first solve.  Returns its status. */
int mpc_oracle_module_load_solution(const mpc_oracle_cfg* cfg, double* traj);

/* Closed loop of mpc_cmd_pub.jl:86-157 + plant: T control steps (10 Hz), each
 * followed by 10 plant publishes; with warm_start the first solve starts from the module-load solution,
 * every later one from the previous solution.  Returns number of steps executed before the
 * stop latch (or T).  log: per step [x,y,psi,v, acc_cmd, df_cmd, status, iters] (8 doubles). */
int mpc_oracle_closed_loop(const mpc_oracle_cfg* cfg, const mpc_oracle_path* p,
                           const double* pose0 /* X0,Y0,Psi0 */, int T, int track_using_time,
                           double target_vel, int warm_start, double* log);

#ifdef __cplusplus
}
#endif
#endif
