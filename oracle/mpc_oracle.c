/*
 * mpc_oracle.c -- CPU oracle (plain C, FP64) for the MPC hot path.
 *
 * TEST INFRASTRUCTURE ONLY -- see mpc_oracle.h.  PARITY UNPINNED (no Ipopt,
 * no golden solver outputs in the reference); see the header comment there.
 *
 * What is restated, and from where:
 *   NLP            scripts/mpc_utils/MKZMPCPathFollower.jl:28-123
 *   step driver    scripts/mpc_cmd_pub.jl:86-157
 *   reference gen  scripts/gps_utils/ref_gps_traj.py:131-218
 *   plant          scripts/vehicle_simulator.py:58-112
 *   solver         Ipopt 3.12 defaults [external, recalled]: Waechter & Biegler 2006.
 *
 * The linear algebra is deliberately NOT the GPU's: the full (x, s, y_c, y_d)
 * augmented system is assembled as one symmetric matrix and factorised with a
 * Bunch-Kaufman LDL^T (inertia from D), with exact-zero skipping inside a
 * monotone profile so that the cost is O(n b^2) rather than O(n^3).
 */
#include "mpc_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define NX 4
#define NU 2
#define NSTG 6 /* x,y,psi,v,acc,df per stage, internal order */
enum { JX = 0, JY = 1, JPSI = 2, JV = 3, JACC = 4, JDF = 5 };
#define IX(k, j) (NSTG * (k) + (j))

/* ------------------------------------------------------------------------- */
/* configuration                                                             */
/* ------------------------------------------------------------------------- */
void mpc_oracle_default_cfg(mpc_oracle_cfg* c, int N) {
    c->N = N;
    c->dt = 0.20;         /* MKZMPCPathFollower.jl:33 */
    c->dt_control = 0.10; /* :28 */
    c->L_a = 1.108;       /* :31 */
    c->L_b = 1.742;       /* :32 */
    c->v_min = 0.0;       /* :47 */
    c->v_max = 20.0;      /* :48 */
    c->a_max = 1.0;       /* :44 */
    c->steer_max = 0.5;   /* :41 */
    c->a_dmax = 1.5;      /* :45 */
    c->steer_dmax = 0.5;  /* :42 */
    /* :51-59 and mpc_cmd_pub.jl:49 : x,y,psi,v,dacc,ddf,acc,df */
    c->w[0] = 9.0; c->w[1] = 9.0; c->w[2] = 10.0; c->w[3] = 0.0;
    c->w[4] = 100.0; c->w[5] = 1000.0; c->w[6] = 0.0; c->w[7] = 0.0;
    c->tol = 1e-8;
    c->max_iter = 200;
    c->model = 0;
    c->kpoly[0] = c->kpoly[1] = c->kpoly[2] = c->kpoly[3] = 0.0;
}

/* MKZMPCPathFollowerFrenet.jl:28-59: same vehicle, horizon, bounds and rate limits; cost C_ey = 9,
 * C_epsi = 10, C_ev = 0.5, C_dacc = 100, C_ddf = 1000, C_acc = C_df = 0; nothing on s. */
void mpc_oracle_default_cfg_frenet(mpc_oracle_cfg* c, int N) {
    mpc_oracle_default_cfg(c, N);
    c->model = 1;
    c->w[0] = 0.0; c->w[1] = 9.0; c->w[2] = 10.0; c->w[3] = 0.5;
}

int mpc_oracle_nvar(const mpc_oracle_cfg* c) { return NSTG * c->N + NX; }
int mpc_oracle_ncon(const mpc_oracle_cfg* c) { return NX + NX * c->N; }
int mpc_oracle_nrange(const mpc_oracle_cfg* c) { return 2 * (c->N - 1); }

void mpc_oracle_traj_to_z(const mpc_oracle_cfg* c, const double* t, double* z) {
    int N = c->N, k;
    const double *x = t, *y = t + (N + 1), *v = t + 2 * (N + 1), *psi = t + 3 * (N + 1);
    const double *df = t + 4 * (N + 1), *acc = df + N;
    for (k = 0; k <= N; k++) {
        z[IX(k, JX)] = x[k]; z[IX(k, JY)] = y[k]; z[IX(k, JPSI)] = psi[k]; z[IX(k, JV)] = v[k];
        if (k < N) { z[IX(k, JACC)] = acc[k]; z[IX(k, JDF)] = df[k]; }
    }
}
void mpc_oracle_z_to_traj(const mpc_oracle_cfg* c, const double* z, double* t) {
    int N = c->N, k;
    double *x = t, *y = t + (N + 1), *v = t + 2 * (N + 1), *psi = t + 3 * (N + 1);
    double *df = t + 4 * (N + 1), *acc = df + N;
    for (k = 0; k <= N; k++) {
        x[k] = z[IX(k, JX)]; y[k] = z[IX(k, JY)]; psi[k] = z[IX(k, JPSI)]; v[k] = z[IX(k, JV)];
        if (k < N) { acc[k] = z[IX(k, JACC)]; df[k] = z[IX(k, JDF)]; }
    }
}

/* ------------------------------------------------------------------------- */
/* NLP: objective  (MKZMPCPathFollower.jl:97-103), 0-based stage index k      */
/* ------------------------------------------------------------------------- */
double mpc_oracle_eval_f(const mpc_oracle_cfg* c, const double* ref, double v_des, const double* z) {
    int N = c->N, k;
    const double *xr = ref, *yr = ref + (N + 1), *pr = ref + 2 * (N + 1);
    double f = 0.0;
    for (k = 1; k <= N; k++) { /* i = 2:(N+1)  (Q6: index 1 of the reference never used) */
        double ex = z[IX(k, JX)] - xr[k], ey = z[IX(k, JY)] - yr[k], ep = z[IX(k, JPSI)] - pr[k];
        f += c->w[0] * ex * ex + c->w[1] * ey * ey + c->w[2] * ep * ep;
    }
    for (k = 1; k <= N - 1; k++) { /* i = 2:N  (Q2: terminal speed uncosted) */
        double ev = z[IX(k, JV)] - v_des;
        f += c->w[3] * ev * ev;
    }
    for (k = 0; k < N; k++) {
        double a = z[IX(k, JACC)], d = z[IX(k, JDF)];
        f += c->w[6] * a * a + c->w[7] * d * d;
    }
    for (k = 0; k < N - 1; k++) { /* i = 1:(N-1) */
        double da = z[IX(k + 1, JACC)] - z[IX(k, JACC)], dd = z[IX(k + 1, JDF)] - z[IX(k, JDF)];
        f += c->w[4] * da * da + c->w[5] * dd * dd;
    }
    return f;
}

void mpc_oracle_eval_grad_f(const mpc_oracle_cfg* c, const double* ref, double v_des, const double* z, double* g) {
    int N = c->N, k, n = mpc_oracle_nvar(c);
    const double *xr = ref, *yr = ref + (N + 1), *pr = ref + 2 * (N + 1);
    memset(g, 0, sizeof(double) * n);
    for (k = 1; k <= N; k++) {
        g[IX(k, JX)] += 2.0 * c->w[0] * (z[IX(k, JX)] - xr[k]);
        g[IX(k, JY)] += 2.0 * c->w[1] * (z[IX(k, JY)] - yr[k]);
        g[IX(k, JPSI)] += 2.0 * c->w[2] * (z[IX(k, JPSI)] - pr[k]);
    }
    for (k = 1; k <= N - 1; k++) g[IX(k, JV)] += 2.0 * c->w[3] * (z[IX(k, JV)] - v_des);
    for (k = 0; k < N; k++) {
        g[IX(k, JACC)] += 2.0 * c->w[6] * z[IX(k, JACC)];
        g[IX(k, JDF)] += 2.0 * c->w[7] * z[IX(k, JDF)];
    }
    for (k = 0; k < N - 1; k++) {
        double da = z[IX(k + 1, JACC)] - z[IX(k, JACC)], dd = z[IX(k + 1, JDF)] - z[IX(k, JDF)];
        g[IX(k + 1, JACC)] += 2.0 * c->w[4] * da; g[IX(k, JACC)] -= 2.0 * c->w[4] * da;
        g[IX(k + 1, JDF)] += 2.0 * c->w[5] * dd;  g[IX(k, JDF)] -= 2.0 * c->w[5] * dd;
    }
}

/* ------------------------------------------------------------------------- */
/* NLP: stage map f(s,u)  (MKZMPCPathFollower.jl:115-123) and derivatives     */
/* ------------------------------------------------------------------------- */
typedef struct {
    double f[4];      /* next state x,y,psi,v */
    double A[4][4];   /* d f / d (x,y,psi,v) */
    double Bd[4];     /* d f / d df ; d f / d acc = (0,0,0,dt) */
    /* second derivatives of f_x, f_y, f_psi over (psi, v, df): pp, pv, pd, vd, dd  (vv = 0) */
    double hx[5], hy[5], hp[5];
    /* Frenet model: full second derivatives of f_s, f_ey, f_epsi over (s, ey, epsi, v, df) */
    double hf[3][5][5];
} stage_eval;

/* MKZMPCPathFollowerFrenet.jl:111-123.  g = ds/dt = v cos(epsi + beta) / (1 - ey K(s)) */
static void stage_map_frenet(const mpc_oracle_cfg* c, const double* st, double acc, double df, stage_eval* e, int order) {
    const double dt = c->dt, Lb = c->L_b, r = c->L_b / (c->L_a + c->L_b);
    const double s = st[0], ey = st[1], ep = st[2], v = st[3];
    const double K = ((c->kpoly[0] * s + c->kpoly[1]) * s + c->kpoly[2]) * s + c->kpoly[3];   /* :111 */
    const double bta = atan(r * tan(df));                                                    /* :112 */
    const double C = cos(ep + bta), S = sin(ep + bta), sb = sin(bta), cb = cos(bta);
    const double q = 1.0 / (1.0 - ey * K);
    const double g = v * C * q;                                                              /* :113 */
    e->f[0] = s + dt * g;                            /* :117 */
    e->f[1] = ey + dt * (v * S);                     /* :118 */
    e->f[2] = ep + dt * (v / Lb * sb - g * K);       /* :119 */
    e->f[3] = v + dt * acc;                          /* :120 */
    if (order < 1) return;
    {
        const double K1 = (3.0 * c->kpoly[0] * s + 2.0 * c->kpoly[1]) * s + c->kpoly[2];
        const double K2 = 6.0 * c->kpoly[0] * s + 2.0 * c->kpoly[1];
        const double cd = cos(df), sd = sin(df);
        const double D = cd * cd + r * r * sd * sd;
        const double b1 = r / D, b2 = r * (1.0 - r * r) * (2.0 * sd * cd) / (D * D);
        const double q2 = q * q, q3 = q2 * q;
        const double qs = ey * K1 * q2, qe = K * q2;
        /* gradient of g over (s, ey, epsi, v, df) */
        const double G[5] = {v * C * qs, v * C * qe, -v * S * q, C * q, -v * S * b1 * q};
        int i, j;
        for (i = 0; i < 4; i++) for (j = 0; j < 4; j++) e->A[i][j] = (i == j) ? 1.0 : 0.0;
        for (j = 0; j < 4; j++) e->A[0][j] += dt * G[j];
        e->Bd[0] = dt * G[4];
        e->A[1][2] += dt * v * C; e->A[1][3] += dt * S; e->Bd[1] = dt * v * C * b1;
        /* h = g K:  h_s = g_s K + g K', h_x = g_x K */
        e->A[2][0] += -dt * (G[0] * K + g * K1);
        e->A[2][1] += -dt * G[1] * K;
        e->A[2][2] += -dt * G[2] * K;
        e->A[2][3] += dt * (sb / Lb - G[3] * K);
        e->Bd[2] = dt * (v * cb * b1 / Lb - G[4] * K);
        e->Bd[3] = 0.0;
        if (order < 2) return;
        {
            const double qss = ey * K2 * q2 + 2.0 * ey * ey * K1 * K1 * q3;
            const double qse = K1 * q2 + 2.0 * ey * K * K1 * q3;
            const double qee = 2.0 * K * K * q3;
            double GG[5][5];
            memset(GG, 0, sizeof(GG)); memset(e->hf, 0, sizeof(e->hf));
            GG[0][0] = v * C * qss; GG[0][1] = v * C * qse; GG[0][2] = -v * S * qs; GG[0][3] = C * qs; GG[0][4] = -v * S * b1 * qs;
            GG[1][1] = v * C * qee; GG[1][2] = -v * S * qe; GG[1][3] = C * qe; GG[1][4] = -v * S * b1 * qe;
            GG[2][2] = -v * C * q;  GG[2][3] = -S * q;      GG[2][4] = -v * C * b1 * q;
            GG[3][4] = -S * b1 * q;
            GG[4][4] = -v * q * (C * b1 * b1 + S * b2);
            for (i = 0; i < 5; i++) for (j = 0; j < i; j++) GG[i][j] = GG[j][i];
            for (i = 0; i < 5; i++) for (j = 0; j < 5; j++) {
                double hh = GG[i][j] * K;                       /* second derivatives of h = g K */
                if (i == 0) hh += G[j] * K1;
                if (j == 0) hh += G[i] * K1;
                if (i == 0 && j == 0) hh += g * K2;
                e->hf[0][i][j] = dt * GG[i][j];
                e->hf[2][i][j] = -dt * hh;
            }
            e->hf[1][2][2] = -dt * v * S; e->hf[1][2][3] = e->hf[1][3][2] = dt * C;
            e->hf[1][2][4] = e->hf[1][4][2] = -dt * v * S * b1; e->hf[1][3][4] = e->hf[1][4][3] = dt * C * b1;
            e->hf[1][4][4] = dt * v * (-S * b1 * b1 + C * b2);
            e->hf[2][3][4] += dt * cb * b1 / Lb; e->hf[2][4][3] += dt * cb * b1 / Lb;
            e->hf[2][4][4] += dt * v / Lb * (-sb * b1 * b1 + cb * b2);
        }
    }
}

static void stage_map(const mpc_oracle_cfg* c, const double* s, double acc, double df, stage_eval* e, int order) {
    if (c->model == 1) { stage_map_frenet(c, s, acc, df, e, order); return; }
    {
    const double dt = c->dt, Lb = c->L_b, r = c->L_b / (c->L_a + c->L_b);
    const double x = s[JX], y = s[JY], psi = s[JPSI], v = s[JV];
    const double bta = atan(r * tan(df)); /* :115 */
    const double cs = cos(psi + bta), sn = sin(psi + bta), sb = sin(bta), cb = cos(bta);
    e->f[JX] = x + dt * (v * cs);        /* :119 */
    e->f[JY] = y + dt * (v * sn);        /* :120 */
    e->f[JPSI] = psi + dt * (v / Lb * sb); /* :121 */
    e->f[JV] = v + dt * acc;             /* :122 */
    if (order < 1) return;
    {
        const double cd = cos(df), sd = sin(df);
        const double D = cd * cd + r * r * sd * sd;
        const double b1 = r / D;                                          /* beta'  */
        const double b2 = r * (1.0 - r * r) * (2.0 * sd * cd) / (D * D);  /* beta'' */
        int i, j;
        for (i = 0; i < 4; i++) for (j = 0; j < 4; j++) e->A[i][j] = (i == j) ? 1.0 : 0.0;
        e->A[JX][JPSI] = -dt * v * sn; e->A[JX][JV] = dt * cs;
        e->A[JY][JPSI] = dt * v * cs;  e->A[JY][JV] = dt * sn;
        e->A[JPSI][JV] = dt * sb / Lb;
        e->Bd[JX] = -dt * v * sn * b1;
        e->Bd[JY] = dt * v * cs * b1;
        e->Bd[JPSI] = dt * v * cb * b1 / Lb;
        e->Bd[JV] = 0.0;
        if (order < 2) return;
        /* order: pp, pv, pd, vd, dd */
        e->hx[0] = -dt * v * cs; e->hx[1] = -dt * sn; e->hx[2] = -dt * v * cs * b1;
        e->hx[3] = -dt * sn * b1; e->hx[4] = -dt * v * (cs * b1 * b1 + sn * b2);
        e->hy[0] = -dt * v * sn; e->hy[1] = dt * cs;  e->hy[2] = -dt * v * sn * b1;
        e->hy[3] = dt * cs * b1;  e->hy[4] = dt * v * (-sn * b1 * b1 + cs * b2);
        e->hp[0] = 0.0; e->hp[1] = 0.0; e->hp[2] = 0.0;
        e->hp[3] = dt * cb * b1 / Lb; e->hp[4] = dt * v / Lb * (-sb * b1 * b1 + cb * b2);
    }
}
}

/* equality rows: 0..3  s_0 - state (:110-113); 4+4k+j  s_{k+1,j} - f_j(s_k,u_k) (:119-122) */
void mpc_oracle_eval_c(const mpc_oracle_cfg* c, const double* state, const double* z, double* cv) {
    int N = c->N, k, j;
    stage_eval e;
    for (j = 0; j < 4; j++) cv[j] = z[IX(0, j)] - state[j];
    for (k = 0; k < N; k++) {
        stage_map(c, &z[IX(k, 0)], z[IX(k, JACC)], z[IX(k, JDF)], &e, 0);
        for (j = 0; j < 4; j++) cv[4 + 4 * k + j] = z[IX(k + 1, j)] - e.f[j];
    }
}

/* range rows (:75-86), q = 0..N-2, two per q: [2q] steering, [2q+1] acceleration.
 * q = 0: first move minus previous command; q >= 1: u_{q+1} - u_q  (Q1: pair (1,0) has no row) */
void mpc_oracle_eval_d(const mpc_oracle_cfg* c, const double* u_prev, const double* z, double* d) {
    int N = c->N, q;
    d[0] = z[IX(0, JDF)] - u_prev[0];
    d[1] = z[IX(0, JACC)] - u_prev[1];
    for (q = 1; q <= N - 2; q++) {
        d[2 * q] = z[IX(q + 1, JDF)] - z[IX(q, JDF)];
        d[2 * q + 1] = z[IX(q + 1, JACC)] - z[IX(q, JACC)];
    }
}

static void range_bounds(const mpc_oracle_cfg* c, int row, double* lo, double* hi) {
    int q = row / 2, ch = row % 2;
    double h = (q == 0) ? c->dt_control : c->dt;
    double lim = (ch == 0 ? c->steer_dmax : c->a_dmax) * h;
    *lo = -lim; *hi = lim;
}

void mpc_oracle_eval_jac(const mpc_oracle_cfg* c, const double* z, double* Jc, double* Jd) {
    int N = c->N, n = mpc_oracle_nvar(c), mc = mpc_oracle_ncon(c), md = mpc_oracle_nrange(c);
    int k, i, j, q;
    stage_eval e;
    memset(Jc, 0, sizeof(double) * mc * n);
    memset(Jd, 0, sizeof(double) * md * n);
    for (j = 0; j < 4; j++) Jc[j * n + IX(0, j)] = 1.0;
    for (k = 0; k < N; k++) {
        stage_map(c, &z[IX(k, 0)], z[IX(k, JACC)], z[IX(k, JDF)], &e, 1);
        for (i = 0; i < 4; i++) {
            double* row = Jc + (size_t)(4 + 4 * k + i) * n;
            row[IX(k + 1, i)] += 1.0;
            for (j = 0; j < 4; j++) row[IX(k, j)] -= e.A[i][j];
            row[IX(k, JDF)] -= e.Bd[i];
        }
        Jc[(size_t)(4 + 4 * k + JV) * n + IX(k, JACC)] -= c->dt;
    }
    Jd[0 * n + IX(0, JDF)] = 1.0;
    Jd[1 * n + IX(0, JACC)] = 1.0;
    for (q = 1; q <= N - 2; q++) {
        Jd[(size_t)(2 * q) * n + IX(q + 1, JDF)] = 1.0;  Jd[(size_t)(2 * q) * n + IX(q, JDF)] = -1.0;
        Jd[(size_t)(2 * q + 1) * n + IX(q + 1, JACC)] = 1.0; Jd[(size_t)(2 * q + 1) * n + IX(q, JACC)] = -1.0;
    }
}

/* dense Hessian of sigma*f + yc'c  (range rows are linear) */
void mpc_oracle_eval_hess(const mpc_oracle_cfg* c, const double* z, double sigma, const double* yc, double* H) {
    int N = c->N, n = mpc_oracle_nvar(c), k;
    stage_eval e;
    memset(H, 0, sizeof(double) * n * n);
#define HH(i, j) H[(size_t)(i) * n + (j)]
    for (k = 1; k <= N; k++) {
        HH(IX(k, JX), IX(k, JX)) += 2.0 * sigma * c->w[0];
        HH(IX(k, JY), IX(k, JY)) += 2.0 * sigma * c->w[1];
        HH(IX(k, JPSI), IX(k, JPSI)) += 2.0 * sigma * c->w[2];
    }
    for (k = 1; k <= N - 1; k++) HH(IX(k, JV), IX(k, JV)) += 2.0 * sigma * c->w[3];
    for (k = 0; k < N; k++) {
        HH(IX(k, JACC), IX(k, JACC)) += 2.0 * sigma * c->w[6];
        HH(IX(k, JDF), IX(k, JDF)) += 2.0 * sigma * c->w[7];
    }
    for (k = 0; k < N - 1; k++) {
        int a0 = IX(k, JACC), a1 = IX(k + 1, JACC), d0 = IX(k, JDF), d1 = IX(k + 1, JDF);
        double wa = 2.0 * sigma * c->w[4], wd = 2.0 * sigma * c->w[5];
        HH(a0, a0) += wa; HH(a1, a1) += wa; HH(a0, a1) -= wa; HH(a1, a0) -= wa;
        HH(d0, d0) += wd; HH(d1, d1) += wd; HH(d0, d1) -= wd; HH(d1, d0) -= wd;
    }
    for (k = 0; k < N; k++) {
        const double* y = yc + 4 + 4 * k; /* c = s+ - f  =>  Hessian term is -y_j * hess f_j */
        int ip = IX(k, JPSI), iv = IX(k, JV), id = IX(k, JDF);
        double pp, pv, pd, vd, dd;
        stage_map(c, &z[IX(k, 0)], z[IX(k, JACC)], z[IX(k, JDF)], &e, 2);
        if (c->model == 1) {
            static const int col[5] = {0, 1, 2, 3, JDF};
            int a, b, rw;
            for (a = 0; a < 5; a++) for (b = 0; b < 5; b++) {
                double t = 0.0;
                for (rw = 0; rw < 3; rw++) t += y[rw] * e.hf[rw][a][b];
                HH(IX(k, col[a]), IX(k, col[b])) -= t;
            }
            continue;
        }
        pp = -(y[JX] * e.hx[0] + y[JY] * e.hy[0] + y[JPSI] * e.hp[0]);
        pv = -(y[JX] * e.hx[1] + y[JY] * e.hy[1] + y[JPSI] * e.hp[1]);
        pd = -(y[JX] * e.hx[2] + y[JY] * e.hy[2] + y[JPSI] * e.hp[2]);
        vd = -(y[JX] * e.hx[3] + y[JY] * e.hy[3] + y[JPSI] * e.hp[3]);
        dd = -(y[JX] * e.hx[4] + y[JY] * e.hy[4] + y[JPSI] * e.hp[4]);
        HH(ip, ip) += pp; HH(ip, iv) += pv; HH(iv, ip) += pv;
        HH(ip, id) += pd; HH(id, ip) += pd; HH(iv, id) += vd; HH(id, iv) += vd; HH(id, id) += dd;
    }
#undef HH
}

/* ------------------------------------------------------------------------- */
/* Symmetric indefinite LDL^T, Bunch-Kaufman partial pivoting (lower), with   */
/* exact-zero skipping inside a monotone row profile hi[].                    */
/* A: dim x dim row-major, lower triangle referenced (A[i][j], i >= j).       */
/* ------------------------------------------------------------------------- */
typedef struct {
    int dim;
    double* A;   /* dim*dim */
    int* ipiv;   /* LAPACK dsytf2 convention, 1-based with sign */
    int* hi;     /* monotone profile: column j has no nonzero below row hi[j] */
    int npos, nneg, nzero;
} ldl_t;

#define LA(i, j) A[(size_t)(i) * dim + (j)]

static void ldl_swap(double* A, int dim, int k, int kk, int kp, int top) {
    /* symmetric interchange of rows/cols kk and kp (kp > kk) in trailing matrix A[k:,k:],
     * rows considered up to 'top' (inclusive) */
    int i;
    double t;
    for (i = kp + 1; i <= top; i++) { t = LA(i, kk); LA(i, kk) = LA(i, kp); LA(i, kp) = t; }
    for (i = kk + 1; i < kp; i++) { t = LA(i, kk); LA(i, kk) = LA(kp, i); LA(kp, i) = t; }
    t = LA(kk, kk); LA(kk, kk) = LA(kp, kp); LA(kp, kp) = t;
    if (kk == k + 1) { t = LA(kk, k); LA(kk, k) = LA(kp, k); LA(kp, k) = t; }
}

static int ldl_factor(ldl_t* F) {
    const double alpha = (1.0 + sqrt(17.0)) / 8.0;
    double* A = F->A;
    int dim = F->dim, *hi = F->hi, *ipiv = F->ipiv;
    int k = 0, i, j;
    F->npos = F->nneg = F->nzero = 0;
    /* initial profile */
    for (j = 0; j < dim; j++) {
        int last = j;
        for (i = dim - 1; i > j; i--) if (LA(i, j) != 0.0) { last = i; break; }
        hi[j] = last;
    }
    for (j = 1; j < dim; j++) if (hi[j] < hi[j - 1]) hi[j] = hi[j - 1];

    while (k < dim) {
        int kstep = 1, kp = k, imax = k, top = hi[k];
        double absakk = fabs(LA(k, k)), colmax = 0.0;
        for (i = k + 1; i <= top; i++) { double a = fabs(LA(i, k)); if (a > colmax) { colmax = a; imax = i; } }
        if (absakk == 0.0 && colmax == 0.0) {
            F->nzero++; ipiv[k] = k + 1; k++; continue;
        }
        if (absakk >= alpha * colmax) {
            kp = k;
        } else {
            double rowmax = 0.0;
            for (j = k; j < imax; j++) { double a = fabs(LA(imax, j)); if (a > rowmax) rowmax = a; }
            for (i = imax + 1; i <= hi[imax]; i++) { double a = fabs(LA(i, imax)); if (a > rowmax) rowmax = a; }
            if (absakk >= alpha * colmax * (colmax / rowmax)) kp = k;
            else if (fabs(LA(imax, imax)) >= alpha * rowmax) kp = imax;
            else { kp = imax; kstep = 2; }
        }
        {
            int kk = k + kstep - 1;
            if (kp != kk) {
                int newtop = hi[kp];
                for (j = k; j <= kp; j++) hi[j] = newtop; /* keep profile monotone and safe */
                top = newtop;
                ldl_swap(A, dim, k, kk, kp, top);
            } else if (kstep == 2) {
                if (hi[k + 1] > top) top = hi[k + 1];
                hi[k] = top;
            }
        }
        if (kstep == 1) {
            double d = LA(k, k), r1 = 1.0 / d;
            if (d > 0) F->npos++; else F->nneg++;
            for (j = k + 1; j <= top; j++) {
                double ljk = LA(j, k);
                if (ljk != 0.0) {
                    double t = ljk * r1;
                    for (i = j; i <= top; i++) LA(i, j) -= LA(i, k) * t;
                }
            }
            for (i = k + 1; i <= top; i++) LA(i, k) *= r1;
            ipiv[k] = kp + 1;
        } else {
            /* 2x2 pivot D = [[a, b],[b, c]] at (k,k+1) */
            double a = LA(k, k), b = LA(k + 1, k), cc = LA(k + 1, k + 1);
            double det = a * cc - b * b;
            double i00 = cc / det, i01 = -b / det, i11 = a / det;
            F->npos++; F->nneg++; /* BK 2x2 pivots are always indefinite */
            for (j = k + 2; j <= top; j++) {
                double wk = LA(j, k) * i00 + LA(j, k + 1) * i01;
                double wk1 = LA(j, k) * i01 + LA(j, k + 1) * i11;
                if (wk != 0.0 || wk1 != 0.0) {
                    for (i = j; i <= top; i++) LA(i, j) -= LA(i, k) * wk + LA(i, k + 1) * wk1;
                }
                /* store L after the column's updates are done: defer */
            }
            for (j = k + 2; j <= top; j++) {
                double lk = LA(j, k), lk1 = LA(j, k + 1);
                LA(j, k) = lk * i00 + lk1 * i01;
                LA(j, k + 1) = lk * i01 + lk1 * i11;
            }
            ipiv[k] = -(kp + 1); ipiv[k + 1] = -(kp + 1);
        }
        k += kstep;
    }
    return 0;
}

/* Careful: in the 2x2 update above the trailing update for column j uses LA(i,k), LA(i,k+1)
 * for i >= j which are still the ORIGINAL (unscaled) columns because scaling is deferred. */

static void ldl_solve(const ldl_t* F, double* b) {
    const double* A = F->A;
    int dim = F->dim, k, i;
    const int *ipiv = F->ipiv, *hi = F->hi;
    /* forward: L then D */
    k = 0;
    while (k < dim) {
        if (ipiv[k] > 0) {
            int kp = ipiv[k] - 1, top = hi[k];
            double bk;
            if (kp != k) { double t = b[k]; b[k] = b[kp]; b[kp] = t; }
            bk = b[k];
            if (bk != 0.0) for (i = k + 1; i <= top; i++) b[i] -= LA(i, k) * bk;
            if (LA(k, k) != 0.0) b[k] = bk / LA(k, k);
            k += 1;
        } else {
            int kp = -ipiv[k] - 1, top = hi[k];
            double bk, bk1, a, bb, cc, det;
            if (kp != k + 1) { double t = b[k + 1]; b[k + 1] = b[kp]; b[kp] = t; }
            bk = b[k]; bk1 = b[k + 1];
            for (i = k + 2; i <= top; i++) b[i] -= LA(i, k) * bk + LA(i, k + 1) * bk1;
            a = LA(k, k); bb = LA(k + 1, k); cc = LA(k + 1, k + 1);
            det = a * cc - bb * bb;
            b[k] = (cc * bk - bb * bk1) / det;
            b[k + 1] = (-bb * bk + a * bk1) / det;
            k += 2;
        }
    }
    /* backward: L^T, with interchanges undone in reverse */
    k = dim - 1;
    while (k >= 0) {
        if (ipiv[k] > 0) {
            int kp = ipiv[k] - 1, top = hi[k];
            double s = b[k];
            for (i = k + 1; i <= top; i++) s -= LA(i, k) * b[i];
            b[k] = s;
            if (kp != k) { double t = b[k]; b[k] = b[kp]; b[kp] = t; }
            k -= 1;
        } else {
            int kp = -ipiv[k] - 1, k0 = k - 1, top = hi[k0];
            double s0 = b[k0], s1 = b[k];
            for (i = k + 1; i <= top; i++) { s0 -= LA(i, k0) * b[i]; s1 -= LA(i, k) * b[i]; }
            b[k0] = s0; b[k] = s1;
            if (kp != k) { double t = b[k]; b[k] = b[kp]; b[kp] = t; }
            k -= 2;
        }
    }
}
#undef LA

/* ------------------------------------------------------------------------- */
/* Interior point (Ipopt 3.12 defaults, recalled)                             */
/* ------------------------------------------------------------------------- */
#define IPM_KAPPA_EPS 10.0       /* barrier_tol_factor */
#define IPM_KAPPA_MU 0.2         /* mu_linear_decrease_factor */
#define IPM_THETA_MU 1.5         /* mu_superlinear_decrease_power */
#define IPM_MU_INIT 0.1
#define IPM_TAU_MIN 0.99
#define IPM_BOUND_PUSH 1e-2
#define IPM_BOUND_FRAC 1e-2
#define IPM_BOUND_RELAX 1e-8
#define IPM_S_MAX 100.0
#define IPM_KAPPA_SIGMA 1e10
#define IPM_GAMMA_THETA 1e-5
#define IPM_GAMMA_PHI 1e-8
#define IPM_ETA_PHI 1e-8
#define IPM_DELTA 1.0
#define IPM_S_THETA 1.1
#define IPM_S_PHI 2.3
#define IPM_ALPHA_MIN_FRAC 0.05
#define IPM_MAX_SOC 4
#define IPM_KAPPA_SOC 0.99
#define IPM_OBJ_MAX_INC 5.0
#define IPM_DW_INIT 1e-4
#define IPM_DW_MIN 1e-20
#define IPM_DW_MAX 1e20
#define IPM_DW_INC_FIRST 100.0
#define IPM_DW_INC 8.0
#define IPM_DW_DEC (1.0 / 3.0)
#define IPM_SCALE_MAX_GRAD 100.0
#define IPM_Y_INIT_MAX 1e3
#define IPM_ACCEPT_TOL 1e-6
#define IPM_ACCEPT_ITER 15
#define IPM_FILTER_MAX 256

typedef struct {
    const mpc_oracle_cfg* cfg;
    const double *state, *ref, *u_prev;
    double v_des;
    int N, n, mc, md, dim;
    /* bounds (relaxed); unbounded entries flagged by has_b == 0 */
    double *xL, *xU; int* has_b;
    double *sL, *sU;
    /* iterate */
    double *x, *s, *yc, *yd, *zL, *zU, *vL, *vU;
    /* step */
    double *dx, *ds, *dyc, *dyd, *dzL, *dzU, *dvL, *dvU;
    /* work */
    double *g, *c, *dms, *Jc, *Jd, *H, *rhs, *xt, *st, *ct, *dt_, *csoc, *dsoc;
    int* perm; /* KKT unknown index -> position */
    ldl_t F;
    double sigma_f;
    int nb_x, nb_s;
} ipm_t;

static double dmax(double a, double b) { return a > b ? a : b; }
static double dmin(double a, double b) { return a < b ? a : b; }

static void* xcalloc(size_t n, size_t sz) { void* p = calloc(n ? n : 1, sz); return p; }

static void ipm_alloc(ipm_t* P, const mpc_oracle_cfg* cfg) {
    int N = cfg->N, n = mpc_oracle_nvar(cfg), mc = mpc_oracle_ncon(cfg), md = mpc_oracle_nrange(cfg);
    int dim = n + md + mc + md;
    memset(P, 0, sizeof(*P));
    P->cfg = cfg; P->N = N; P->n = n; P->mc = mc; P->md = md; P->dim = dim;
#define AL(name, cnt) P->name = (double*)xcalloc((cnt), sizeof(double))
    AL(xL, n); AL(xU, n); AL(sL, md); AL(sU, md);
    AL(x, n); AL(s, md); AL(yc, mc); AL(yd, md); AL(zL, n); AL(zU, n); AL(vL, md); AL(vU, md);
    AL(dx, n); AL(ds, md); AL(dyc, mc); AL(dyd, md); AL(dzL, n); AL(dzU, n); AL(dvL, md); AL(dvU, md);
    AL(g, n); AL(c, mc); AL(dms, md); AL(Jc, (size_t)mc * n); AL(Jd, (size_t)md * n); AL(H, (size_t)n * n);
    AL(rhs, dim); AL(xt, n); AL(st, md); AL(ct, mc); AL(dt_, md); AL(csoc, mc); AL(dsoc, md);
#undef AL
    P->has_b = (int*)xcalloc(n, sizeof(int));
    P->perm = (int*)xcalloc(dim, sizeof(int));
    P->F.dim = dim;
    P->F.A = (double*)xcalloc((size_t)dim * dim, sizeof(double));
    P->F.ipiv = (int*)xcalloc(dim, sizeof(int));
    P->F.hi = (int*)xcalloc(dim, sizeof(int));
    /* Stage-interleaved elimination order so the KKT matrix is banded:
     *   [y_c rows entering stage k][s_k][u_k][range slack, y_d of the rows whose later input is u_k] */
    {
        int pos = 0, k, j;
        int ox = 0, os = n, oc = n + md, od = n + md + mc; /* unknown index offsets */
        for (k = 0; k <= N; k++) {
            for (j = 0; j < 4; j++) P->perm[oc + 4 * k + j] = pos++;
            for (j = 0; j < (k < N ? 6 : 4); j++) P->perm[ox + IX(k, j)] = pos++;
            if (k < N) {
                int q = (k == 0) ? 0 : (k >= 2 ? k - 1 : -1);
                if (q >= 0 && q <= N - 2) {
                    for (j = 0; j < 2; j++) P->perm[os + 2 * q + j] = pos++;
                    for (j = 0; j < 2; j++) P->perm[od + 2 * q + j] = pos++;
                }
            }
        }
    }
}

static void ipm_free(ipm_t* P) {
    free(P->xL); free(P->xU); free(P->sL); free(P->sU);
    free(P->x); free(P->s); free(P->yc); free(P->yd); free(P->zL); free(P->zU); free(P->vL); free(P->vU);
    free(P->dx); free(P->ds); free(P->dyc); free(P->dyd); free(P->dzL); free(P->dzU); free(P->dvL); free(P->dvU);
    free(P->g); free(P->c); free(P->dms); free(P->Jc); free(P->Jd); free(P->H);
    free(P->rhs); free(P->xt); free(P->st); free(P->ct); free(P->dt_); free(P->csoc); free(P->dsoc);
    free(P->has_b); free(P->perm); free(P->F.A); free(P->F.ipiv); free(P->F.hi);
}

/* constraint violation theta (1-norm, Ipopt constr_viol_normtype default) at (x, s) */
static double ipm_theta(ipm_t* P, const double* x, const double* s, double* c, double* dms) {
    int i; double th = 0.0;
    mpc_oracle_eval_c(P->cfg, P->state, x, c);
    mpc_oracle_eval_d(P->cfg, P->u_prev, x, dms);
    for (i = 0; i < P->md; i++) dms[i] -= s[i];
    for (i = 0; i < P->mc; i++) th += fabs(c[i]);
    for (i = 0; i < P->md; i++) th += fabs(dms[i]);
    return th;
}

/* barrier objective phi_mu at (x, s) */
static double ipm_barrier(ipm_t* P, const double* x, const double* s, double mu) {
    int i; double phi = P->sigma_f * mpc_oracle_eval_f(P->cfg, P->ref, P->v_des, x), lb = 0.0;
    for (i = 0; i < P->n; i++) if (P->has_b[i]) lb += log(x[i] - P->xL[i]) + log(P->xU[i] - x[i]);
    for (i = 0; i < P->md; i++) lb += log(s[i] - P->sL[i]) + log(P->sU[i] - s[i]);
    return phi - mu * lb;
}

/* assemble + factorise the augmented system with W (or identity), Sigma, delta_w.
 * unknown order: [x (n)] [s (md)] [y_c (mc)] [y_d (md)]
 *   [ W+Sx+dw   0        Jc'   Jd' ]
 *   [ 0         Ss+dw    0     -I  ]
 *   [ Jc        0        0     0   ]
 *   [ Jd        -I       0     0   ]                                         */
static void ipm_factor(ipm_t* P, int use_hess, const double* Sx, const double* Ss, double dw) {
    int n = P->n, md = P->md, mc = P->mc, dim = P->dim, i, j;
    double* A = P->F.A;
    const int* pm = P->perm;
    int os = n, oc = n + md, od = n + md + mc;
    memset(A, 0, sizeof(double) * (size_t)dim * dim);
#define SET(u, v, val) do { int pu = pm[u], pv = pm[v]; if (pu >= pv) A[(size_t)pu * dim + pv] += (val); else A[(size_t)pv * dim + pu] += (val); } while (0)
    for (i = 0; i < n; i++) {
        if (use_hess) { for (j = 0; j <= i; j++) { double h = P->H[(size_t)i * n + j]; if (h != 0.0) SET(i, j, h); } }
        SET(i, i, (use_hess ? 0.0 : 1.0) + (Sx ? Sx[i] : 0.0) + dw);
    }
    for (i = 0; i < md; i++) SET(os + i, os + i, (Ss ? Ss[i] : 1.0) + dw);
    for (i = 0; i < mc; i++) for (j = 0; j < n; j++) { double v = P->Jc[(size_t)i * n + j]; if (v != 0.0) SET(oc + i, j, v); }
    for (i = 0; i < md; i++) {
        for (j = 0; j < n; j++) { double v = P->Jd[(size_t)i * n + j]; if (v != 0.0) SET(od + i, j, v); }
        SET(od + i, os + i, -1.0);
    }
#undef SET
    ldl_factor(&P->F);
}

static void ipm_solve(ipm_t* P, const double* rx, const double* rs, const double* rc, const double* rd,
                      double* ox, double* os_, double* oyc, double* oyd) {
    int n = P->n, md = P->md, mc = P->mc, i;
    const int* pm = P->perm;
    double* b = P->rhs;
    for (i = 0; i < n; i++) b[pm[i]] = rx[i];
    for (i = 0; i < md; i++) b[pm[n + i]] = rs[i];
    for (i = 0; i < mc; i++) b[pm[n + md + i]] = rc[i];
    for (i = 0; i < md; i++) b[pm[n + md + mc + i]] = rd[i];
    ldl_solve(&P->F, b);
    for (i = 0; i < n; i++) ox[i] = b[pm[i]];
    for (i = 0; i < md; i++) os_[i] = b[pm[n + i]];
    for (i = 0; i < mc; i++) oyc[i] = b[pm[n + md + i]];
    for (i = 0; i < md; i++) oyd[i] = b[pm[n + md + mc + i]];
}

typedef struct { double phi, theta; } filt_entry;

static int cmp_le(double lhs, double rhs, double bas) { return lhs - rhs <= 10.0 * DBL_EPSILON * fabs(bas); }

typedef struct {
    double dual_inf, constr_viol, compl_inf; /* scaled-NLP quantities, max-norm */
    double sd, sc;
} err_t;

/* E_mu pieces (Ipopt curr_nlp_error / curr_barrier_error) at the current iterate */
static void ipm_errors(ipm_t* P, double mu, err_t* e) {
    int n = P->n, mc = P->mc, md = P->md, i, j;
    double di = 0.0, cv = 0.0, cm = 0.0, sumy = 0.0, sumz = 0.0;
    int nz = 2 * P->nb_x + 2 * md;
    for (j = 0; j < n; j++) {
        double gl = P->g[j];
        for (i = 0; i < mc; i++) { double v = P->Jc[(size_t)i * n + j]; if (v != 0.0) gl += v * P->yc[i]; }
        for (i = 0; i < md; i++) { double v = P->Jd[(size_t)i * n + j]; if (v != 0.0) gl += v * P->yd[i]; }
        if (P->has_b[j]) gl += -P->zL[j] + P->zU[j];
        di = dmax(di, fabs(gl));
    }
    for (i = 0; i < md; i++) di = dmax(di, fabs(-P->yd[i] - P->vL[i] + P->vU[i]));
    for (i = 0; i < mc; i++) { cv = dmax(cv, fabs(P->c[i])); sumy += fabs(P->yc[i]); }
    for (i = 0; i < md; i++) { cv = dmax(cv, fabs(P->dms[i])); sumy += fabs(P->yd[i]); }
    for (j = 0; j < n; j++) if (P->has_b[j]) {
        cm = dmax(cm, fabs((P->x[j] - P->xL[j]) * P->zL[j] - mu));
        cm = dmax(cm, fabs((P->xU[j] - P->x[j]) * P->zU[j] - mu));
        sumz += fabs(P->zL[j]) + fabs(P->zU[j]);
    }
    for (i = 0; i < md; i++) {
        cm = dmax(cm, fabs((P->s[i] - P->sL[i]) * P->vL[i] - mu));
        cm = dmax(cm, fabs((P->sU[i] - P->s[i]) * P->vU[i] - mu));
        sumz += fabs(P->vL[i]) + fabs(P->vU[i]);
    }
    e->dual_inf = di; e->constr_viol = cv; e->compl_inf = cm;
    e->sd = dmax(IPM_S_MAX, (sumy + sumz) / (double)(mc + md + nz)) / IPM_S_MAX;
    e->sc = dmax(IPM_S_MAX, sumz / (double)nz) / IPM_S_MAX;
}
static double err_total(const err_t* e) { return dmax(dmax(e->dual_inf / e->sd, e->constr_viol), e->compl_inf / e->sc); }

static double frac_to_bound_primal(ipm_t* P, double tau, const double* dx, const double* ds) {
    int i; double a = 1.0;
    for (i = 0; i < P->n; i++) if (P->has_b[i]) {
        if (dx[i] < 0.0) a = dmin(a, -tau * (P->x[i] - P->xL[i]) / dx[i]);
        if (dx[i] > 0.0) a = dmin(a, tau * (P->xU[i] - P->x[i]) / dx[i]);
    }
    for (i = 0; i < P->md; i++) {
        if (ds[i] < 0.0) a = dmin(a, -tau * (P->s[i] - P->sL[i]) / ds[i]);
        if (ds[i] > 0.0) a = dmin(a, tau * (P->sU[i] - P->s[i]) / ds[i]);
    }
    return a;
}

static double frac_to_bound_dual(ipm_t* P, double tau) {
    int i; double a = 1.0;
    for (i = 0; i < P->n; i++) if (P->has_b[i]) {
        if (P->dzL[i] < 0.0) a = dmin(a, -tau * P->zL[i] / P->dzL[i]);
        if (P->dzU[i] < 0.0) a = dmin(a, -tau * P->zU[i] / P->dzU[i]);
    }
    for (i = 0; i < P->md; i++) {
        if (P->dvL[i] < 0.0) a = dmin(a, -tau * P->vL[i] / P->dvL[i]);
        if (P->dvU[i] < 0.0) a = dmin(a, -tau * P->vU[i] / P->dvU[i]);
    }
    return a;
}

static void push_interior(double* v, double lo, double hi) {
    /* Ipopt DefaultIterateInitializer::push_variables, two-sided bounds */
    double pl = dmin(IPM_BOUND_PUSH * dmax(1.0, fabs(lo)), IPM_BOUND_FRAC * (hi - lo));
    double pu = dmin(IPM_BOUND_PUSH * dmax(1.0, fabs(hi)), IPM_BOUND_FRAC * (hi - lo));
    if (*v < lo + pl) *v = lo + pl;
    if (*v > hi - pu) *v = hi - pu;
}

/* a-priori feasibility of the bound/equality/range system (stands in for the
 * restoration phase's Infeasible_Problem_Detected): the only ways this NLP can be
 * infeasible are v0 outside [v_min, v_max] or the previous command further from the
 * input box than one first-move rate step. */
static int nlp_feasible(const mpc_oracle_cfg* c, const double* state, const double* u_prev) {
    const double e = 1e-8;
    double lim_d = c->steer_dmax * c->dt_control, lim_a = c->a_dmax * c->dt_control;
    if (state[3] < c->v_min - e * dmax(1.0, fabs(c->v_min))) return 0;
    if (state[3] > c->v_max + e * dmax(1.0, fabs(c->v_max))) return 0;
    if (u_prev[0] - lim_d > c->steer_max + 2 * e || u_prev[0] + lim_d < -c->steer_max - 2 * e) return 0;
    if (u_prev[1] - lim_a > c->a_max + 2 * e || u_prev[1] + lim_a < -c->a_max - 2 * e) return 0;
    return 1;
}

/* Interior start from the primal point in P->x (DefaultIterateInitializer): push x inside its
 * bounds, slacks = d(x) pushed, bound multipliers 1, least-squares equality multipliers.
 * Used for the user's start point and after a restoration. */
static void ipm_init_point(ipm_t* P, double* rx, double* rs, double* rc, double* rd) {
    const mpc_oracle_cfg* cfg = P->cfg;
    int n = P->n, mc = P->mc, md = P->md, i;
    for (i = 0; i < n; i++) if (P->has_b[i]) push_interior(&P->x[i], P->xL[i], P->xU[i]);
    mpc_oracle_eval_d(cfg, P->u_prev, P->x, P->s);
    for (i = 0; i < md; i++) push_interior(&P->s[i], P->sL[i], P->sU[i]);
    for (i = 0; i < n; i++) { P->zL[i] = P->has_b[i] ? 1.0 : 0.0; P->zU[i] = P->has_b[i] ? 1.0 : 0.0; }
    for (i = 0; i < md; i++) { P->vL[i] = 1.0; P->vU[i] = 1.0; }

    /* ---- least-squares equality multipliers ---- */
    mpc_oracle_eval_grad_f(cfg, P->ref, P->v_des, P->x, P->g);
    for (i = 0; i < n; i++) P->g[i] *= P->sigma_f;
    mpc_oracle_eval_jac(cfg, P->x, P->Jc, P->Jd);
    ipm_factor(P, 0, NULL, NULL, 0.0);
    for (i = 0; i < n; i++) rx[i] = -(P->g[i] - P->zL[i] + P->zU[i]);
    for (i = 0; i < md; i++) rs[i] = -(-P->vL[i] + P->vU[i]);
    memset(rc, 0, sizeof(double) * mc); memset(rd, 0, sizeof(double) * md);
    ipm_solve(P, rx, rs, rc, rd, P->dx, P->ds, P->yc, P->yd);
    {
        double ym = 0.0;
        for (i = 0; i < mc; i++) ym = dmax(ym, fabs(P->yc[i]));
        for (i = 0; i < md; i++) ym = dmax(ym, fabs(P->yd[i]));
        if (!(ym <= IPM_Y_INIT_MAX)) { memset(P->yc, 0, sizeof(double) * mc); memset(P->yd, 0, sizeof(double) * md); }
    }

}

/* ---- restoration by rollout ------------------------------------------------------------------
 * Stands in for Ipopt's restoration phase (MinC_1NrmRestorationPhase), which is NOT restated: when
 * the filter line search fails, the iterate is replaced by a point that satisfies the equality rows
 * by construction -- the inputs of the current iterate, projected stage by stage onto the input box
 * and the rate rows, and the states obtained by rolling the bicycle model out from the measured
 * state -- and the interior-point iteration is re-initialised there (push into the interior, slacks,
 * bound multipliers 1, least-squares y) with mu and the filter kept.  Like Ipopt's restoration it
 * returns a point with (almost) no constraint violation that must be acceptable to the filter. */
#define IPM_MAX_RESTO 3
static double clipd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
static void ipm_rollout_restore(ipm_t* P) {
    const mpc_oracle_cfg* c = P->cfg;
    int N = P->N, k, j;
    double pa = P->u_prev[1], pd = P->u_prev[0];
    stage_eval e;
    for (j = 0; j < 4; j++) P->x[IX(0, j)] = P->state[j];
    for (k = 0; k < N; k++) {
        double ua = P->x[IX(k, JACC)], ud = P->x[IX(k, JDF)];
        if (k != 1) {   /* rate rows exist for the first move and for pairs (k, k-1), k >= 2 (Q1) */
            const double h = (k == 0) ? c->dt_control : c->dt;
            const double la = 0.98 * c->a_dmax * h, ld = 0.98 * c->steer_dmax * h;   /* 1 % of the range inside, like bound_frac */
            ua = clipd(ua, pa - la, pa + la);
            ud = clipd(ud, pd - ld, pd + ld);
        }
        ua = clipd(ua, -c->a_max, c->a_max);
        ud = clipd(ud, -c->steer_max, c->steer_max);
        /* keep the speed inside its bounds: v_{k+1} = v_k + dt * acc */
        {
            const double v = P->x[IX(k, JV)];
            ua = clipd(ua, (c->v_min - v) / c->dt, (c->v_max - v) / c->dt);
        }
        P->x[IX(k, JACC)] = ua; P->x[IX(k, JDF)] = ud;
        stage_map(c, &P->x[IX(k, 0)], ua, ud, &e, 0);
        for (j = 0; j < 4; j++) P->x[IX(k + 1, j)] = e.f[j];
        pa = ua; pd = ud;
    }
}

static int ipm_run(ipm_t* P, const double* warm, int* iters_out, mpc_oracle_diag* dg) {
    const mpc_oracle_cfg* cfg = P->cfg;
    int N = P->N, n = P->n, mc = P->mc, md = P->md, i, j, k, iter = 0;
    double mu = IPM_MU_INIT, tau = dmax(IPM_TAU_MIN, 1.0 - IPM_MU_INIT);
    const double mu_min = dmin(cfg->tol, 1e-4) / (IPM_KAPPA_EPS + 1.0);
    double dw_last = 0.0, theta_max = -1.0, theta_min = -1.0;
    filt_entry filt[IPM_FILTER_MAX]; int nfilt = 0;
    int accept_count = 0, ret = -1, tiny_last = 0, n_resto = 0, cur_acceptable = 0, had_acceptable = 0;
    const int use_resto = getenv("MPC_ORACLE_NO_RESTO") == NULL;
    double *Sx = (double*)xcalloc(n, sizeof(double)), *Ss = (double*)xcalloc(md, sizeof(double));
    double *rx = (double*)xcalloc(n, sizeof(double)), *rs = (double*)xcalloc(md, sizeof(double));
    double *rc = (double*)xcalloc(mc, sizeof(double)), *rd = (double*)xcalloc(md, sizeof(double));

    dg->n_inertia_corr = dg->n_soc = dg->n_backtrack = 0;

    /* ---- bounds, relaxed by bound_relax_factor ---- */
    for (i = 0; i < n; i++) P->has_b[i] = 0;
    P->nb_x = 0;
    for (k = 0; k <= N; k++) {
        int iv = IX(k, JV);
        P->has_b[iv] = 1; P->xL[iv] = cfg->v_min; P->xU[iv] = cfg->v_max; P->nb_x++;
        if (k < N) {
            int ia = IX(k, JACC), id = IX(k, JDF);
            P->has_b[ia] = 1; P->xL[ia] = -cfg->a_max; P->xU[ia] = cfg->a_max;
            P->has_b[id] = 1; P->xL[id] = -cfg->steer_max; P->xU[id] = cfg->steer_max;
            P->nb_x += 2;
        }
    }
    for (i = 0; i < n; i++) if (P->has_b[i]) {
        P->xL[i] -= IPM_BOUND_RELAX * dmax(1.0, fabs(P->xL[i]));
        P->xU[i] += IPM_BOUND_RELAX * dmax(1.0, fabs(P->xU[i]));
    }
    for (i = 0; i < md; i++) {
        range_bounds(cfg, i, &P->sL[i], &P->sU[i]);
        P->sL[i] -= IPM_BOUND_RELAX * dmax(1.0, fabs(P->sL[i]));
        P->sU[i] += IPM_BOUND_RELAX * dmax(1.0, fabs(P->sU[i]));
    }
    P->nb_s = md;

    /* ---- starting point: start=0.0 (MKZMPCPathFollower.jl:65-72) or previous solution ---- */
    if (warm) mpc_oracle_traj_to_z(cfg, warm, P->x); else memset(P->x, 0, sizeof(double) * n);

    /* gradient-based objective scaling at the user's start point (nlp_scaling_max_gradient = 100);
     * every constraint row has max |gradient entry| <= max(1, dt*v*...) << 100, so rows are unscaled */
    mpc_oracle_eval_grad_f(cfg, P->ref, P->v_des, P->x, P->g);
    {
        double gm = 0.0;
        for (i = 0; i < n; i++) gm = dmax(gm, fabs(P->g[i]));
        P->sigma_f = (gm > IPM_SCALE_MAX_GRAD) ? dmax(IPM_SCALE_MAX_GRAD / gm, 1e-8) : 1.0;
    }
    dg->obj_scale = P->sigma_f;

    ipm_init_point(P, rx, rs, rc, rd);

    for (;;) {
        err_t e0, em;
        double theta, phi, gBd, alpha_max, alpha, alpha_min, alpha_z;
        int accepted = 0, tiny = 0;

        /* ---- evaluate at current iterate ---- */
        mpc_oracle_eval_grad_f(cfg, P->ref, P->v_des, P->x, P->g);
        for (i = 0; i < n; i++) P->g[i] *= P->sigma_f;
        theta = ipm_theta(P, P->x, P->s, P->c, P->dms);
        mpc_oracle_eval_jac(cfg, P->x, P->Jc, P->Jd);

        /* ---- convergence (OptimalityErrorConvergenceCheck) ---- */
        ipm_errors(P, 0.0, &e0);
        {
            double E0 = err_total(&e0);
            double du = e0.dual_inf / P->sigma_f, cu = e0.constr_viol, mu_c = e0.compl_inf / P->sigma_f;
            dg->dual_inf = du; dg->constr_viol = cu; dg->compl_inf = mu_c; dg->mu_final = mu;
            if (E0 <= cfg->tol && du <= 1.0 && cu <= 1e-4 && mu_c <= 1e-4) { ret = 0; break; }
            cur_acceptable = (E0 <= IPM_ACCEPT_TOL && du <= 1e10 && cu <= 1e-2 && mu_c <= 1e-2);
            if (cur_acceptable) {
                had_acceptable = 1;   /* BacktrackingLineSearch::StoreAcceptablePoint */
                if (++accept_count >= IPM_ACCEPT_ITER) { ret = 1; break; }
            } else accept_count = 0;
        }
        if (iter >= cfg->max_iter) { ret = -1; break; }
        /* Diverging_Iterates: diverging_iterates_tol = 1e20 */
        {
            double xm = 0.0; for (i = 0; i < n; i++) xm = dmax(xm, fabs(P->x[i]));
            if (!(xm <= 1e20)) { ret = -5; break; }
        }

        /* ---- monotone barrier update ---- */
        for (;;) {
            double nm;
            ipm_errors(P, mu, &em);
            if (!(err_total(&em) <= IPM_KAPPA_EPS * mu) && !tiny_last) break;
            nm = dmax(mu_min, dmin(IPM_KAPPA_MU * mu, pow(mu, IPM_THETA_MU)));
            if (nm >= mu) { if (tiny_last) { ret = -3; } break; }
            mu = nm; tau = dmax(IPM_TAU_MIN, 1.0 - mu);
            nfilt = 0; /* filter reset on mu change */
            if (tiny_last) { tiny_last = 0; break; }
        }
        if (ret == -3) break;

        /* ---- primal-dual system ---- */
        mpc_oracle_eval_hess(cfg, P->x, P->sigma_f, P->yc, P->H);
        for (i = 0; i < n; i++) Sx[i] = P->has_b[i] ? P->zL[i] / (P->x[i] - P->xL[i]) + P->zU[i] / (P->xU[i] - P->x[i]) : 0.0;
        for (i = 0; i < md; i++) Ss[i] = P->vL[i] / (P->s[i] - P->sL[i]) + P->vU[i] / (P->sU[i] - P->s[i]);
        {
            double dw = 0.0; int ok = 0, tries = 0;
            for (;;) {
                ipm_factor(P, 1, Sx, Ss, dw);
                if (P->F.nzero == 0 && P->F.npos == n + md && P->F.nneg == mc + md) { ok = 1; break; }
                /* PDPerturbationHandler::PerturbForWrongInertia */
                if (dw == 0.0) dw = (dw_last == 0.0) ? IPM_DW_INIT : dmax(IPM_DW_MIN, dw_last * IPM_DW_DEC);
                else dw = (dw_last == 0.0 || 1e5 * dw_last < dw) ? IPM_DW_INC_FIRST * dw : IPM_DW_INC * dw;
                if (dw > IPM_DW_MAX) break;
                tries++;
            }
            if (!ok) { ret = -4; break; }
            if (dw > 0.0) { dw_last = dw; dg->n_inertia_corr++; }
        }
        /* rhs (negated residuals of the barrier KKT conditions) */
        for (j = 0; j < n; j++) {
            double gl = P->g[j];
            for (i = 0; i < mc; i++) { double v = P->Jc[(size_t)i * n + j]; if (v != 0.0) gl += v * P->yc[i]; }
            for (i = 0; i < md; i++) { double v = P->Jd[(size_t)i * n + j]; if (v != 0.0) gl += v * P->yd[i]; }
            if (P->has_b[j]) gl += -mu / (P->x[j] - P->xL[j]) + mu / (P->xU[j] - P->x[j]);
            rx[j] = -gl;
        }
        for (i = 0; i < md; i++) rs[i] = -(-P->yd[i] - mu / (P->s[i] - P->sL[i]) + mu / (P->sU[i] - P->s[i]));
        for (i = 0; i < mc; i++) rc[i] = -P->c[i];
        for (i = 0; i < md; i++) rd[i] = -P->dms[i];
        ipm_solve(P, rx, rs, rc, rd, P->dx, P->ds, P->dyc, P->dyd);
        for (j = 0; j < n; j++) if (P->has_b[j]) {
            double sl = P->x[j] - P->xL[j], su = P->xU[j] - P->x[j];
            P->dzL[j] = mu / sl - P->zL[j] - P->zL[j] / sl * P->dx[j];
            P->dzU[j] = mu / su - P->zU[j] + P->zU[j] / su * P->dx[j];
        }
        for (i = 0; i < md; i++) {
            double sl = P->s[i] - P->sL[i], su = P->sU[i] - P->s[i];
            P->dvL[i] = mu / sl - P->vL[i] - P->vL[i] / sl * P->ds[i];
            P->dvU[i] = mu / su - P->vU[i] + P->vU[i] / su * P->ds[i];
        }

        /* ---- line search (BacktrackingLineSearch + FilterLSAcceptor) ---- */
        phi = ipm_barrier(P, P->x, P->s, mu);
        gBd = 0.0;
        for (j = 0; j < n; j++) {
            double gb = P->g[j];
            if (P->has_b[j]) gb += -mu / (P->x[j] - P->xL[j]) + mu / (P->xU[j] - P->x[j]);
            gBd += gb * P->dx[j];
        }
        for (i = 0; i < md; i++) gBd += (-mu / (P->s[i] - P->sL[i]) + mu / (P->sU[i] - P->s[i])) * P->ds[i];
        if (theta_max < 0.0) { theta_max = 1e4 * dmax(1.0, theta); theta_min = 1e-4 * dmax(1.0, theta); }

        alpha_max = frac_to_bound_primal(P, tau, P->dx, P->ds);
        alpha_min = IPM_GAMMA_THETA;
        if (gBd < 0.0) {
            alpha_min = dmin(IPM_GAMMA_THETA, IPM_GAMMA_PHI * theta / (-gBd));
            if (theta <= theta_min) alpha_min = dmin(alpha_min, IPM_DELTA * pow(theta, IPM_S_THETA) / pow(-gBd, IPM_S_PHI));
        }
        alpha_min *= IPM_ALPHA_MIN_FRAC;

        /* tiny step (BacktrackingLineSearch::DetectTinyStep) */
        {
            double m = 0.0;
            for (j = 0; j < n; j++) m = dmax(m, fabs(P->dx[j]) / (1.0 + fabs(P->x[j])));
            for (i = 0; i < md; i++) m = dmax(m, fabs(P->ds[i]) / (1.0 + fabs(P->s[i])));
            tiny = (m < 10.0 * DBL_EPSILON) && (theta < 1e-4);
        }

        alpha = alpha_max;
        {
            int nsteps = 0, ftype_arm = 0;
            double *udx = P->dx, *uds = P->ds; /* direction actually used (may become the SOC one) */
            double alpha_test = alpha_max;
            if (tiny) {
                for (j = 0; j < n; j++) P->xt[j] = P->x[j] + alpha * P->dx[j];
                for (i = 0; i < md; i++) P->st[i] = P->s[i] + alpha * P->ds[i];
                accepted = 1;
            }
            while (!accepted && (alpha > alpha_min || nsteps == 0)) {
                double th_t, ph_t; int ok;
                alpha_test = alpha;
                for (j = 0; j < n; j++) P->xt[j] = P->x[j] + alpha * udx[j];
                for (i = 0; i < md; i++) P->st[i] = P->s[i] + alpha * uds[i];
                th_t = ipm_theta(P, P->xt, P->st, P->ct, P->dt_);
                ph_t = ipm_barrier(P, P->xt, P->st, mu);
#define IS_FTYPE(a) (gBd < 0.0 && (a) * pow(-gBd, IPM_S_PHI) > IPM_DELTA * pow(theta, IPM_S_THETA))
#define ARMIJO(a, pt) cmp_le((pt) - phi, IPM_ETA_PHI * (a) * gBd, phi)
#define ACCEPT_TEST(ok_, a_, tht_, pht_) do { \
                    ok_ = 1; \
                    if (!(tht_ == tht_) || !(pht_ == pht_)) ok_ = 0; \
                    if (ok_ && !cmp_le(tht_, theta_max, theta)) ok_ = 0; \
                    if (ok_) { \
                        if (IS_FTYPE(a_) && theta <= theta_min) ok_ = ARMIJO(a_, pht_); \
                        else { \
                            if (pht_ > phi) { double bas = 1.0; if (fabs(phi) > 10.0) bas = log10(fabs(phi)); \
                                if (log10(pht_ - phi) > IPM_OBJ_MAX_INC + bas) ok_ = 0; } \
                            if (ok_) ok_ = cmp_le(tht_, (1.0 - IPM_GAMMA_THETA) * theta, theta) || cmp_le(pht_ - phi, -IPM_GAMMA_PHI * theta, phi); \
                        } \
                    } \
                    if (ok_) { int f_; for (f_ = 0; f_ < nfilt; f_++) \
                        if (!(cmp_le(pht_, filt[f_].phi, filt[f_].phi) || cmp_le(tht_, filt[f_].theta, filt[f_].theta))) { ok_ = 0; break; } } \
                } while (0)
                ACCEPT_TEST(ok, alpha_test, th_t, ph_t);
                if (ok) { accepted = 1; ftype_arm = IS_FTYPE(alpha_test) && ARMIJO(alpha_test, ph_t); break; }

                /* second-order correction, first trial only */
                if (nsteps == 0 && theta <= th_t && IPM_MAX_SOC > 0) {
                    double th_old = 0.0, th_soc = th_t, a_soc = alpha;
                    int cnt = 0;
                    memcpy(P->csoc, P->c, sizeof(double) * mc);
                    memcpy(P->dsoc, P->dms, sizeof(double) * md);
                    while (cnt < IPM_MAX_SOC && !accepted && (cnt == 0 || th_soc <= IPM_KAPPA_SOC * th_old)) {
                        double* sx = P->dzL; /* borrow: dz are recomputed after acceptance */
                        (void)sx;
                        th_old = th_soc;
                        for (i = 0; i < mc; i++) P->csoc[i] = a_soc * P->csoc[i] + P->ct[i];
                        for (i = 0; i < md; i++) P->dsoc[i] = a_soc * P->dsoc[i] + P->dt_[i];
                        for (i = 0; i < mc; i++) rc[i] = -P->csoc[i];
                        for (i = 0; i < md; i++) rd[i] = -P->dsoc[i];
                        {
                            double *sdx = (double*)xcalloc(n, sizeof(double)), *sds = (double*)xcalloc(md, sizeof(double));
                            double *syc = (double*)xcalloc(mc, sizeof(double)), *syd = (double*)xcalloc(md, sizeof(double));
                            double ph_s;
                            ipm_solve(P, rx, rs, rc, rd, sdx, sds, syc, syd);
                            a_soc = frac_to_bound_primal(P, tau, sdx, sds);
                            for (j = 0; j < n; j++) P->xt[j] = P->x[j] + a_soc * sdx[j];
                            for (i = 0; i < md; i++) P->st[i] = P->s[i] + a_soc * sds[i];
                            th_soc = ipm_theta(P, P->xt, P->st, P->ct, P->dt_);
                            ph_s = ipm_barrier(P, P->xt, P->st, mu);
                            ACCEPT_TEST(ok, alpha_test, th_soc, ph_s);
                            if (ok) {
                                accepted = 1; dg->n_soc++;
                                memcpy(P->dx, sdx, sizeof(double) * n); memcpy(P->ds, sds, sizeof(double) * md);
                                memcpy(P->dyc, syc, sizeof(double) * mc); memcpy(P->dyd, syd, sizeof(double) * md);
                                alpha = a_soc;
                                ftype_arm = IS_FTYPE(alpha_test) && ARMIJO(alpha_test, ph_s);
                                /* bound-multiplier steps follow the corrected primal step */
                                for (j = 0; j < n; j++) if (P->has_b[j]) {
                                    double sl = P->x[j] - P->xL[j], su = P->xU[j] - P->x[j];
                                    P->dzL[j] = mu / sl - P->zL[j] - P->zL[j] / sl * P->dx[j];
                                    P->dzU[j] = mu / su - P->zU[j] + P->zU[j] / su * P->dx[j];
                                }
                                for (i = 0; i < md; i++) {
                                    double sl = P->s[i] - P->sL[i], su = P->sU[i] - P->s[i];
                                    P->dvL[i] = mu / sl - P->vL[i] - P->vL[i] / sl * P->ds[i];
                                    P->dvU[i] = mu / su - P->vU[i] + P->vU[i] / su * P->ds[i];
                                }
                            } else cnt++;
                            free(sdx); free(sds); free(syc); free(syd);
                        }
                    }
                    if (accepted) break;
                }
                alpha *= 0.5; nsteps++; dg->n_backtrack++;
            }
            if (!accepted && getenv("MPC_ORACLE_TRACE")) {
                int jb = -1; double ab = 1.0;
                for (j = 0; j < n; j++) if (P->has_b[j]) {
                    double a1 = P->dx[j] < 0 ? -tau * (P->x[j] - P->xL[j]) / P->dx[j] : (P->dx[j] > 0 ? tau * (P->xU[j] - P->x[j]) / P->dx[j] : 1.0);
                    if (a1 < ab) { ab = a1; jb = j; }
                }
                fprintf(stderr, "LS FAIL it=%d amax=%.3e amin=%.3e theta=%.3e gBd=%.3e limiting var stage %d comp %d x=%.6e dx=%.3e nsteps=%d\n",
                        iter, alpha_max, alpha_min, theta, gBd, jb / 6, jb % 6, jb >= 0 ? P->x[jb] : 0.0, jb >= 0 ? P->dx[jb] : 0.0, nsteps);
            }
            /* BacktrackingLineSearch: "Restoration phase called at acceptable point" -> Solved_To_Acceptable_Level.
             * Near the optimum the line search fails on rounding alone; the point is then returned, not restored. */
            if (!accepted && cur_acceptable) { ret = 1; break; }
            /* "Restoration phase is called at point that is almost feasible": theta <= 1e-2 tol.  Ipopt then goes back
             * to the last acceptable point if it stored one (Solved_To_Acceptable_Level), else gives up
             * (Restoration_Failed).  The stored point is not kept here: with theta at rounding level the iterate has
             * not left it (the line search has been failing on rounding), so the current point is returned. */
            if (!accepted && theta <= 1e-2 * cfg->tol) { ret = had_acceptable ? 1 : -2; break; }
            if (!accepted && use_resto && n_resto < IPM_MAX_RESTO) {
                /* restoration: filter augmented with the point that is left (PrepareRestoPhaseStart) */
                double th_r, ph_r; int ok;
                n_resto++;
                if (nfilt < IPM_FILTER_MAX) { filt[nfilt].phi = phi - IPM_GAMMA_PHI * theta; filt[nfilt].theta = (1.0 - IPM_GAMMA_THETA) * theta; nfilt++; }
                had_acceptable = 0;
                ipm_rollout_restore(P);
                ipm_init_point(P, rx, rs, rc, rd);
                th_r = ipm_theta(P, P->x, P->s, P->ct, P->dt_);
                ph_r = ipm_barrier(P, P->x, P->s, mu);
                ok = (th_r == th_r) && (ph_r == ph_r) && cmp_le(th_r, theta_max, theta_max);
                if (ok) { int f_; for (f_ = 0; f_ < nfilt; f_++)
                    if (!(cmp_le(ph_r, filt[f_].phi, filt[f_].phi) || cmp_le(th_r, filt[f_].theta, filt[f_].theta))) { ok = 0; break; } }
                iter++;   /* the restoration counts as one iteration */
                if (!ok) { ret = -2; break; }
                dg->n_resto++;
                tiny_last = 0;
                continue;
            }
            if (!accepted) { ret = -2; break; } /* restoration budget spent */

            /* filter augmentation (FilterLSAcceptor::UpdateForNextIteration) */
            if (!tiny && !ftype_arm) {
                if (nfilt < IPM_FILTER_MAX) {
                    filt[nfilt].phi = phi - IPM_GAMMA_PHI * theta;
                    filt[nfilt].theta = (1.0 - IPM_GAMMA_THETA) * theta;
                    nfilt++;
                }
            }
        }

        /* ---- accept: primal with alpha, y with alpha (alpha_for_y=primal), z with alpha_z ---- */
        alpha_z = frac_to_bound_dual(P, tau);
        memcpy(P->x, P->xt, sizeof(double) * n);
        memcpy(P->s, P->st, sizeof(double) * md);
        for (i = 0; i < mc; i++) P->yc[i] += alpha * P->dyc[i];
        for (i = 0; i < md; i++) P->yd[i] += alpha * P->dyd[i];
        for (j = 0; j < n; j++) if (P->has_b[j]) {
            double sl = P->x[j] - P->xL[j], su = P->xU[j] - P->x[j];
            P->zL[j] += alpha_z * P->dzL[j]; P->zU[j] += alpha_z * P->dzU[j];
            P->zL[j] = dmax(dmin(P->zL[j], IPM_KAPPA_SIGMA * mu / sl), mu / (IPM_KAPPA_SIGMA * sl));
            P->zU[j] = dmax(dmin(P->zU[j], IPM_KAPPA_SIGMA * mu / su), mu / (IPM_KAPPA_SIGMA * su));
        }
        for (i = 0; i < md; i++) {
            double sl = P->s[i] - P->sL[i], su = P->sU[i] - P->s[i];
            P->vL[i] += alpha_z * P->dvL[i]; P->vU[i] += alpha_z * P->dvU[i];
            P->vL[i] = dmax(dmin(P->vL[i], IPM_KAPPA_SIGMA * mu / sl), mu / (IPM_KAPPA_SIGMA * sl));
            P->vU[i] = dmax(dmin(P->vU[i], IPM_KAPPA_SIGMA * mu / su), mu / (IPM_KAPPA_SIGMA * su));
        }
        if (getenv("MPC_ORACLE_TRACE")) {
            double dn = 0.0; for (j = 0; j < n; j++) dn = dmax(dn, fabs(P->dx[j]));
            fprintf(stderr, "%3d f=%.8e th=%.2e du=%.2e lg(mu)=%5.1f |d|=%.2e dw=%.1e az=%.2e a=%.2e gBd=%.2e phi=%.6e nf=%d\n",
                    iter, mpc_oracle_eval_f(cfg, P->ref, P->v_des, P->x), theta, e0.dual_inf, log10(mu), dn, dw_last, alpha_z, alpha, gBd, phi, nfilt);
        }
        tiny_last = tiny;
        iter++;
    }
    *iters_out = iter;
    dg->ipopt_status = ret;
    free(Sx); free(Ss); free(rx); free(rs); free(rc); free(rd);
    return ret;
}

static int map_status(int r) {
    switch (r) {
        case 0: case 1: return MPC_ORACLE_OPTIMAL;
        case 2: return MPC_ORACLE_INFEASIBLE;
        case -1: return MPC_ORACLE_USERLIMIT;
        case -5: return MPC_ORACLE_UNBOUNDED;
        default: return MPC_ORACLE_ERROR;
    }
}

static int solve_with(ipm_t* P, const double* state, const double* ref, double v_des, const double* u_prev,
                      const double* warm, double* traj, double* u0, double* cost, int* status, int* iters,
                      mpc_oracle_diag* diag) {
    const mpc_oracle_cfg* cfg = P->cfg;
    mpc_oracle_diag dloc; int r, it = 0, i;
    if (!diag) diag = &dloc;
    memset(diag, 0, sizeof(*diag));
    P->state = state; P->ref = ref; P->u_prev = u_prev; P->v_des = v_des;
    if (!nlp_feasible(cfg, state, u_prev)) {
        /* report the start point, like the node that publishes whatever the solver holds */
        if (warm) mpc_oracle_traj_to_z(cfg, warm, P->x); else memset(P->x, 0, sizeof(double) * P->n);
        r = 2;
        diag->ipopt_status = 2;
    } else {
        r = ipm_run(P, warm, &it, diag);
    }
    /* honor_original_bounds: project the returned primal onto the original bounds */
    for (i = 0; i < P->n; i++) if (P->has_b[i] && r != 2) {
        int j = i % NSTG;
        double lo = (j == JV) ? cfg->v_min : (j == JACC ? -cfg->a_max : -cfg->steer_max);
        double hi = (j == JV) ? cfg->v_max : (j == JACC ? cfg->a_max : cfg->steer_max);
        if (P->x[i] < lo) P->x[i] = lo;
        if (P->x[i] > hi) P->x[i] = hi;
    }
    if (traj) mpc_oracle_z_to_traj(cfg, P->x, traj);
    if (u0) { u0[0] = P->x[IX(0, JACC)]; u0[1] = P->x[IX(0, JDF)]; }
    if (cost) *cost = mpc_oracle_eval_f(cfg, ref, v_des, P->x);
    if (status) *status = map_status(r);
    if (iters) *iters = it;
    return 0;
}

/* MPCB200_START_ROLLOUT of include/mpc_b200.h: the previous command held over the horizon (projected
 * like ipm_rollout_restore does) and the model rolled out from the measured state; traj order. */
int mpc_oracle_rollout_start(const mpc_oracle_cfg* cfg, const double* state, const double* u_prev, double* traj) {
    ipm_t P; int k;
    if (cfg->N < 3 || cfg->N > 512) return -1;
    ipm_alloc(&P, cfg);
    P.state = state; P.u_prev = u_prev;
    memset(P.x, 0, sizeof(double) * P.n);
    for (k = 0; k < cfg->N; k++) { P.x[IX(k, JACC)] = u_prev[1]; P.x[IX(k, JDF)] = u_prev[0]; }
    ipm_rollout_restore(&P);
    mpc_oracle_z_to_traj(cfg, P.x, traj);
    ipm_free(&P);
    return 0;
}

int mpc_oracle_solve(const mpc_oracle_cfg* cfg, const double* state, const double* ref, double v_des,
                     const double* u_prev, const double* warm, double* traj, double* u0, double* cost,
                     int* status, int* iters, mpc_oracle_diag* diag) {
    ipm_t P;
    if (cfg->N < 3 || cfg->N > 512) return -1;
    ipm_alloc(&P, cfg);
    solve_with(&P, state, ref, v_des, u_prev, warm, traj, u0, cost, status, iters, diag);
    ipm_free(&P);
    return 0;
}

/* n_resto: optional [B] restorations by rollout per problem (see mpc_oracle_diag) */
int mpc_oracle_solve_batch_resto(const mpc_oracle_cfg* cfg, long B, const double* state, const double* ref,
                                 const double* v_des, const double* u_prev, double* warm, double* u0,
                                 double* cost, int* status, int* iters, double* traj, int* n_resto, int n_threads) {
    int N = cfg->N, nt = 6 * N + 4, nr = 3 * (N + 1);
    if (cfg->N < 3 || cfg->N > 512) return -1;
    if (n_threads < 1) n_threads = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads)
#endif
    {
        ipm_t P; long b; mpc_oracle_diag dg;
        double* tbuf = (double*)malloc(sizeof(double) * nt);
        ipm_alloc(&P, cfg);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (b = 0; b < B; b++) {
            solve_with(&P, state + 4 * b, ref + (size_t)nr * b, v_des ? v_des[b] : 0.0, u_prev + 2 * b,
                       warm ? warm + (size_t)nt * b : NULL, tbuf, u0 ? u0 + 2 * b : NULL,
                       cost ? cost + b : NULL, status ? status + b : NULL, iters ? iters + b : NULL, &dg);
            if (n_resto) n_resto[b] = dg.n_resto;
            if (traj) memcpy(traj + (size_t)nt * b, tbuf, sizeof(double) * nt);
            if (warm) memcpy(warm + (size_t)nt * b, tbuf, sizeof(double) * nt);
        }
        ipm_free(&P); free(tbuf);
    }
    return 0;
}

int mpc_oracle_solve_batch(const mpc_oracle_cfg* cfg, long B, const double* state, const double* ref,
                           const double* v_des, const double* u_prev, double* warm, double* u0,
                           double* cost, int* status, int* iters, double* traj, int n_threads) {
    return mpc_oracle_solve_batch_resto(cfg, B, state, ref, v_des, u_prev, warm, u0, cost, status, iters, traj, NULL, n_threads);
}

/* Frenet-frame variant: one private copy of cfg per thread carries the problem's curvature polynomial;
 * the cost is the XY cost against a zero reference (see mpc_oracle.h). */
int mpc_oracle_solve_batch_frenet(const mpc_oracle_cfg* cfg, long B, const double* state, const double* kpoly,
                                  const double* v_des, const double* u_prev, double* warm, double* u0,
                                  double* cost, int* status, int* iters, double* traj, int n_threads) {
    int N = cfg->N, nt = 6 * N + 4, nr = 3 * (N + 1);
    if (cfg->N < 3 || cfg->N > 512 || cfg->model != 1) return -1;
    if (n_threads < 1) n_threads = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads)
#endif
    {
        ipm_t P; long b;
        mpc_oracle_cfg lc = *cfg;
        double* tbuf = (double*)malloc(sizeof(double) * nt);
        double* zref = (double*)xcalloc(nr, sizeof(double));
        ipm_alloc(&P, &lc);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (b = 0; b < B; b++) {
            memcpy(lc.kpoly, kpoly + 4 * b, sizeof(double) * 4);
            solve_with(&P, state + 4 * b, zref, v_des ? v_des[b] : 0.0, u_prev + 2 * b,
                       warm ? warm + (size_t)nt * b : NULL, tbuf, u0 ? u0 + 2 * b : NULL,
                       cost ? cost + b : NULL, status ? status + b : NULL, iters ? iters + b : NULL, NULL);
            if (traj) memcpy(traj + (size_t)nt * b, tbuf, sizeof(double) * nt);
            if (warm) memcpy(warm + (size_t)nt * b, tbuf, sizeof(double) * nt);
        }
        ipm_free(&P); free(tbuf); free(zref);
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* reference generator, ref_gps_traj.py:131-218                               */
/* ------------------------------------------------------------------------- */
static double np_interp(double xq, const double* xp, const double* fp, int n) {
    /* numpy.interp semantics: clamp outside, linear inside; xp increasing */
    int lo = 0, hi = n - 1;
    if (xq <= xp[0]) return fp[0];
    if (xq >= xp[n - 1]) return fp[n - 1];
    while (hi - lo > 1) { int mid = (lo + hi) / 2; if (xp[mid] <= xq) lo = mid; else hi = mid; }
    {
        /* numpy: slope = (fp[j+1]-fp[j])/(xp[j+1]-xp[j]); result = slope*(x - xp[j]) + fp[j] */
        double slope = (fp[lo + 1] - fp[lo]) / (xp[lo + 1] - xp[lo]);
        return slope * (xq - xp[lo]) + fp[lo];
    }
}

int mpc_oracle_get_waypoints(const mpc_oracle_path* p, int horizon, double traj_dt, double X, double Y,
                             double yaw, int use_vtarget, double v_target, double* ref) {
    int n = p->n, i, best = 0, h, np1 = horizon + 1;
    double bd = INFINITY;
    double *xr = ref, *yr = ref + np1, *pr = ref + 2 * np1;
    const double* absc = use_vtarget ? p->s : p->t;
    double start, chk1 = 0.0, chk2 = 0.0;
    for (i = 0; i < n; i++) { /* :136-137, argmin returns the first minimum */
        double dx = p->X[i] - X, dy = p->Y[i] - Y, d = dx * dx + dy * dy;
        if (d < bd) { bd = d; best = i; }
    }
    start = absc[best];
    for (h = 0; h < np1; h++) {
        /* distance mode :175 uses x = 1..N+1; time mode :191 uses h = 0..N */
        double q = use_vtarget ? ((double)(h + 1) * traj_dt * v_target + start) : ((double)h * traj_dt + start);
        xr[h] = np_interp(q, absc, p->X, n);
        yr[h] = np_interp(q, absc, p->Y, n);
        pr[h] = np_interp(q, absc, p->psi, n);
    }
    /* :204-218 heading wrap-around fix */
    for (h = 0; h + 1 < np1; h++) chk1 = dmax(chk1, fabs(pr[h + 1] - pr[h]));
    for (h = 0; h < np1; h++) chk2 = dmax(chk2, fabs(pr[h] - yaw));
    if (!(chk1 < M_PI && chk2 < M_PI)) {
        for (h = 0; h < np1; h++) {
            double c0 = pr[h], c1 = pr[h] + 2.0 * M_PI, c2 = pr[h] - 2.0 * M_PI;
            double b = c0, e = fabs(c0 - yaw);
            if (fabs(c1 - yaw) < e) { e = fabs(c1 - yaw); b = c1; }
            if (fabs(c2 - yaw) < e) { b = c2; }
            pr[h] = b;
        }
    }
    return (xr[np1 - 1] == p->X[n - 1] && yr[np1 - 1] == p->Y[n - 1]) ? 1 : 0; /* :198-200 */
}

/* ------------------------------------------------------------------------- */
/* plant, vehicle_simulator.py:58-112                                         */
/* ------------------------------------------------------------------------- */
static double py_mod(double a, double m) { double r = fmod(a, m); if (r != 0.0 && ((r < 0.0) != (m < 0.0))) r += m; return r; }

void mpc_oracle_plant_step(double* st, double acc_des, double df_des) {
    const double lf = 1.152, lr = 1.693, m = 1840.0, Iz = 3477.0, Caf = 4.0703e4, Car = 6.4495e4; /* :61-67 */
    const double deltaT = 0.01 / 10.0; /* :69 */
    const double inv_m_vx = (double)(1 / 1840); /* = 0, see :86 below */
    int i;
    for (i = 0; i < 10; i++) {
        double X = st[0], Y = st[1], psi = st[2], vx = st[3], vy = st[4], wz = st[5], acc = st[6], df = st[7];
        double af = 0.0, ar = 0.0, Fyf, Fyr, vx_n, vy_n, wz_n, psi_n, X_n, Y_n;
        if (fabs(vx) > 1e-6) {
            af = df - atan2(vy + lf * wz, vx);
            ar = -atan2(vy - lf * wz, vx); /* :77 uses lf (sic) */
        }
        Fyf = Caf * af; Fyr = Car * ar;
        /* :86 is `1/m*Fyf*np.sin(self.df)` in a Python 2 node with the int m = 1840: integer division, 1/m == 0 */
        vx_n = vx + deltaT * (acc - inv_m_vx * Fyf * sin(df) + wz * vy);
        if (vx_n < 0.0) vx_n = 0.0;
        if (vx_n > 1e-6) {
            vy_n = vy + deltaT * (1.0 / m * (Fyf * cos(df) + Fyr) - wz * vx);
            wz_n = wz + deltaT * (1.0 / Iz * (lf * Fyf * cos(df) - lr * Fyr));
        } else { vy_n = 0.0; wz_n = 0.0; }
        psi_n = psi + deltaT * wz;
        X_n = X + deltaT * (vx * cos(psi) - vy * sin(psi));
        Y_n = Y + deltaT * (vx * sin(psi) + vy * cos(psi));
        st[0] = X_n; st[1] = Y_n;
        st[2] = py_mod(psi_n + M_PI, 2.0 * M_PI) - M_PI; /* :101 */
        st[3] = vx_n; st[4] = vy_n; st[5] = wz_n;
        st[6] = 5.0 * (acc_des - acc) * deltaT + acc; /* :111 */
        st[7] = 5.0 * (df_des - df) * deltaT + df;    /* :112 */
    }
}

/* ------------------------------------------------------------------------- */
/* closed loop, mpc_cmd_pub.jl:86-157 driving the plant at 100 Hz             */
/* ------------------------------------------------------------------------- */
/* The solution of the module-load solve (MKZMPCPathFollower.jl:126-128) of the default problem: zero state and
 * previous command (:75,82,110-113), x_ref = 15 t, y_ref = psi_ref = 0, v_target = 15 (:36-39,93), module-default
 * weights (:51-59), start = 0.0 (:65-72).  JuMP <= 0.18 re-solves from the previous primal values, so this is the
 * start point of the node's FIRST solve. */
int mpc_oracle_module_load_solution(const mpc_oracle_cfg* cfg, double* traj) {
    mpc_oracle_cfg c = *cfg;
    const double w0[8] = {9.0, 9.0, 10.0, 0.0, 100.0, 1000.0, 0.0, 0.0};
    int N = cfg->N, k, status = 0, iters = 0;
    double state[4] = {0, 0, 0, 0}, u_prev[2] = {0, 0}, u0[2], cost;
    double* ref = (double*)calloc(3 * (N + 1), sizeof(double));
    memcpy(c.w, w0, sizeof(w0));
    for (k = 0; k <= N; k++) ref[k] = 15.0 * ((double)k * cfg->dt);
    mpc_oracle_solve(&c, state, ref, 15.0, u_prev, NULL, traj, u0, &cost, &status, &iters, NULL);
    free(ref);
    return status;
}

int mpc_oracle_closed_loop(const mpc_oracle_cfg* cfg, const mpc_oracle_path* p, const double* pose0, int T,
                           int track_using_time, double target_vel, int warm_start, double* log) {
    int N = cfg->N, nt = 6 * N + 4, step, i, stop = 0;
    double st[8] = {pose0[0], pose0[1], pose0[2], 0, 0, 0, 0, 0};
    double acc_des = 0.0, df_des = 0.0;     /* vehicle_simulator.py:20-21 */
    double u_prev[2] = {0.0, 0.0};          /* d_f_current, acc_current  MKZMPCPathFollower.jl:75,82 */
    double des_speed = target_vel > 0.0 ? target_vel : 0.0; /* mpc_cmd_pub.jl:58-62 */
    double* ref = (double*)malloc(sizeof(double) * 3 * (N + 1));
    double* warm = (double*)calloc(nt, sizeof(double));
    double* traj = (double*)malloc(sizeof(double) * nt);
    ipm_t P;
    if (warm_start) mpc_oracle_module_load_solution(cfg, warm);
    ipm_alloc(&P, cfg);
    for (step = 0; step < T; step++) {
        double state[4], u0[2] = {0, 0}, cost = 0;
        int status = 0, iters = 0, sc;
        /* the node samples the latest state_est published by the plant (10 publishes per control period) */
        for (i = 0; i < 10; i++) mpc_oracle_plant_step(st, acc_des, df_des);
        state[0] = st[0]; state[1] = st[1]; state[2] = st[2]; state[3] = st[3];
        sc = mpc_oracle_get_waypoints(p, N, cfg->dt, state[0], state[1], state[2], !track_using_time, des_speed, ref);
        if (sc) stop = 1; /* latch, mpc_cmd_pub.jl:102-111 */
        if (!stop) {
            solve_with(&P, state, ref, des_speed, u_prev, warm_start ? warm : NULL, traj, u0, &cost, &status, &iters, NULL);
            if (warm_start) memcpy(warm, traj, sizeof(double) * nt);
            acc_des = u0[0]; df_des = u0[1];       /* published whatever the status, :129-132 */
            u_prev[0] = u0[1]; u_prev[1] = u0[0];  /* update_current_input(df_opt, a_opt), :140 */
        } else {
            acc_des = -1.0; df_des = 0.0;          /* :148-153 */
            status = -1;
        }
        if (log) {
            double* l = log + 8 * step;
            l[0] = state[0]; l[1] = state[1]; l[2] = state[2]; l[3] = state[3];
            l[4] = acc_des; l[5] = df_des; l[6] = (double)status; l[7] = (double)iters;
        }
    }
    ipm_free(&P); free(ref); free(warm); free(traj);
    return step;
}
