/* TEST ONLY: prints the layout of the structs that cross the C ABI (include/mpc_b200.h), one "name offset size" line
 * per field, so that the ctypes binding (capi.Config / capi.Stats) and the Julia shim's `Config` can be checked against
 * what a C compiler lays out. */
#include <stddef.h>
#include <stdio.h>
#include "mpc_b200.h"

#define F(T, f) printf(#T "." #f " %zu %zu\n", offsetof(T, f), sizeof(((T*)0)->f))
int main(void) {
    printf("mpcb200_config sizeof %zu\n", sizeof(mpcb200_config));
    F(mpcb200_config, N); F(mpcb200_config, max_iter); F(mpcb200_config, start_mode); F(mpcb200_config, device);
    F(mpcb200_config, dt); F(mpcb200_config, dt_control); F(mpcb200_config, L_a); F(mpcb200_config, L_b);
    F(mpcb200_config, v_min); F(mpcb200_config, v_max); F(mpcb200_config, a_max); F(mpcb200_config, steer_max);
    F(mpcb200_config, a_dmax); F(mpcb200_config, steer_dmax); F(mpcb200_config, tol);
    F(mpcb200_config, n_devices); F(mpcb200_config, devices);
    printf("mpcb200_stats sizeof %zu\n", sizeof(mpcb200_stats));
    F(mpcb200_stats, kernel_launches); F(mpcb200_stats, h2d_bytes); F(mpcb200_stats, d2h_bytes); F(mpcb200_stats, kernel_ms);
    printf("MPCB200_VERSION %d\n", MPCB200_VERSION);
    return 0;
}
