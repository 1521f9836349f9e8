/* TEST ONLY: a plain-C caller of libmpc_b200.so doing exactly what julia/MKZMPCPathFollower.jl does behind
 * scripts/mpc_cmd_pub.jl:45-51,115-141 -- a non-Python process driving the C ABI with a batch of one:
 *
 *   module load      mpcb200_default_config -> mpcb200_create -> solve of the default problem (MKZMPCPathFollower.jl:126-128)
 *   mpc_cmd_pub.jl:49   mpcb200_set_cost(9, 9, 10, 0, 100, 1000, 0, 0)
 *   per control step    update_init_cond / update_reference (:115-116) = the step's fixture row,
 *                       solve_model (:121) = mpcb200_solve_batch(B = 1, warm in/out),
 *                       update_current_input(df_opt, a_opt) (:140)
 *
 * usage: mpc_cmd_loop <fixture.bin>     fixture = int32 N, int32 steps, then per step 4 + 3(N+1) + 1 doubles
 *                                       (state x,y,psi,v; x_ref, y_ref, psi_ref; des_speed)
 * prints one line per step: step status iters acc df cost   (floats as C99 hex, so the comparison is bit for bit)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mpc_b200.h"

static void die(mpcb200_handle* h, const char* what, int rc) {
    fprintf(stderr, "%s failed (%d): %s\n", what, rc, mpcb200_last_error(h));
    exit(2);
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s fixture.bin\n", argv[0]); return 1; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    int32_t hdr[2];
    if (fread(hdr, sizeof(int32_t), 2, f) != 2) return 1;
    const int N = hdr[0], steps = hdr[1], nr = 3 * (N + 1), nt = 6 * N + 4;

    mpcb200_config cfg;
    mpcb200_handle* h = NULL;
    int rc;
    if ((rc = mpcb200_default_config(&cfg, N))) die(NULL, "mpcb200_default_config", rc);
    if ((rc = mpcb200_create(&h, &cfg))) die(NULL, "mpcb200_create", rc);

    double state[4] = {0, 0, 0, 0}, u_curr[2] = {0, 0}, v_target = 15.0, u0[2], cost;
    int32_t status, iters;
    double* ref = (double*)calloc(nr, sizeof(double));
    double* warm = (double*)calloc(nt, sizeof(double));      /* start = 0.0; afterwards the last solution */
    for (int k = 0; k <= N; k++) ref[k] = 15.0 * ((double)k * cfg.dt);   /* x_ref = v_ref * (0:dt:N*dt), :36-37 */
    /* "MPC: Initial solve ..." */
    if ((rc = mpcb200_solve_batch(h, 1, state, ref, &v_target, u_curr, warm, u0, &cost, &status, &iters, NULL, MPCB200_HOST)))
        die(h, "mpcb200_solve_batch (module load)", rc);
    printf("load %d %d %a %a %a\n", status, iters, u0[0], u0[1], cost);

    const double w[8] = {9.0, 9.0, 10.0, 0.0, 100.0, 1000.0, 0.0, 0.0};   /* mpc_cmd_pub.jl:49 */
    if ((rc = mpcb200_set_cost(h, w))) die(h, "mpcb200_set_cost", rc);

    for (int t = 0; t < steps; t++) {
        if (fread(state, sizeof(double), 4, f) != 4 || fread(ref, sizeof(double), nr, f) != (size_t)nr ||
            fread(&v_target, sizeof(double), 1, f) != 1) { fprintf(stderr, "short fixture\n"); return 1; }
        if ((rc = mpcb200_solve_batch(h, 1, state, ref, &v_target, u_curr, warm, u0, &cost, &status, &iters, NULL, MPCB200_HOST)))
            die(h, "mpcb200_solve_batch", rc);
        printf("%d %d %d %a %a %a\n", t, status, iters, u0[0], u0[1], cost);
        u_curr[0] = u0[1]; u_curr[1] = u0[0];   /* update_current_input(df_opt, a_opt): steering first */
    }
    mpcb200_stats st;
    mpcb200_get_stats(h, &st);
    printf("stats %lld %lld %lld\n", (long long)st.kernel_launches, (long long)st.h2d_bytes, (long long)st.d2h_bytes);
    mpcb200_destroy(h);
    free(ref); free(warm); fclose(f);
    return 0;
}
