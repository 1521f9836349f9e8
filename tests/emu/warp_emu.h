// warp_emu.h -- TEST ONLY.  Single-threaded emulation of one 32-lane warp with ucontext
// coroutines, so that mkz_mpc_path_follower_b200/csrc/mpc_kernel.cuh can be compiled by g++
// and stepped on a CPU.  Every collective is a rendezvous of all 32 lanes (full mask), checked
// by call-site id.  Never part of the shipped library.
#pragma once
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#define MPC_DEV inline
#define MPC_DEV_NOINLINE inline
#define MPC_FULL 0xffffffffu

namespace mpcb200 {
namespace emu {
struct Warp {
    ucontext_t main_ctx, ctx[32];
    char* stacks;
    int cur;          // running lane
    int done[32];
    double dslot[2][32];
    int islot[2][32];
    int site[2][32];
    int epoch[32];    // per-lane collective counter
    void (*fn)(int lane, void* arg);
    void* arg;
};
extern thread_local Warp* W;

inline void yield_next() {
    Warp* w = W;
    int from = w->cur;
    int to = from;
    for (int i = 1; i <= 32; i++) { int c = (from + i) & 31; if (!w->done[c]) { to = c; break; } }
    if (to == from) return;
    w->cur = to;
    swapcontext(&w->ctx[from], &w->ctx[to]);
}
// rendezvous: publish value, let every other lane reach the same collective, then read
inline void publish_d(double v, int site) {
    Warp* w = W; int l = w->cur; int b = w->epoch[l] & 1;
    w->dslot[b][l] = v; w->site[b][l] = site;
    yield_next();
}
inline void publish_i(int v, int site) {
    Warp* w = W; int l = w->cur; int b = w->epoch[l] & 1;
    w->islot[b][l] = v; w->site[b][l] = site;
    yield_next();
}
inline void check_site(int b, int site) {
    Warp* w = W;
    for (int i = 0; i < 32; i++) if (w->site[b][i] != site) {
        fprintf(stderr, "warp_emu: divergent collective (lane %d at site %d, lane %d at site %d)\n", w->cur, site, i, w->site[b][i]);
        abort();
    }
}
void run_warp(void (*fn)(int, void*), void* arg);
}  // namespace emu

MPC_DEV int lane_id() { return emu::W->cur; }
#define MPC_EMU_COLLECTIVE_D(expr_src)                                   \
    emu::Warp* w = emu::W; int l = w->cur; int b = w->epoch[l] & 1;      \
    emu::publish_d(v, __LINE__); emu::check_site(b, __LINE__);           \
    w->epoch[l]++; int s_ = (expr_src);
MPC_DEV double shfl(double v, int src) { MPC_EMU_COLLECTIVE_D(src & 31) return w->dslot[b][s_]; }
MPC_DEV double shfl_down(double v, int d) { MPC_EMU_COLLECTIVE_D(l + d) return s_ < 32 ? w->dslot[b][s_] : v; }
MPC_DEV double shfl_up(double v, int d) { MPC_EMU_COLLECTIVE_D(l - d) return s_ >= 0 ? w->dslot[b][s_] : v; }
MPC_DEV double shfl_xor(double v, int m) { MPC_EMU_COLLECTIVE_D(l ^ m) return w->dslot[b][s_]; }
MPC_DEV int shfl(int v, int src) {
    emu::Warp* w = emu::W; int l = w->cur; int b = w->epoch[l] & 1;
    emu::publish_i(v, __LINE__); emu::check_site(b, __LINE__); w->epoch[l]++;
    return w->islot[b][src & 31];
}
MPC_DEV void syncwarp() {
    emu::Warp* w = emu::W; int l = w->cur; int b = w->epoch[l] & 1;
    emu::publish_i(0, __LINE__); emu::check_site(b, __LINE__); w->epoch[l]++;
}
MPC_DEV bool warp_all(bool p) {
    emu::Warp* w = emu::W; int l = w->cur; int b = w->epoch[l] & 1;
    emu::publish_i(p ? 1 : 0, __LINE__); emu::check_site(b, __LINE__); w->epoch[l]++;
    int r = 1; for (int i = 0; i < 32; i++) r &= w->islot[b][i]; return r != 0;
}
MPC_DEV bool warp_any(bool p) {
    emu::Warp* w = emu::W; int l = w->cur; int b = w->epoch[l] & 1;
    emu::publish_i(p ? 1 : 0, __LINE__); emu::check_site(b, __LINE__); w->epoch[l]++;
    int r = 0; for (int i = 0; i < 32; i++) r |= w->islot[b][i]; return r != 0;
}
MPC_DEV void mpc_sincos(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }
MPC_DEV float fast_log2(float x) { return log2f(x); }
MPC_DEV float fast_exp2(float x) { return exp2f(x); }
// shared memory: a plain double array, offsets in doubles
typedef double* smem_t;
#define SO(x) (x)
MPC_DEV smem_t smem_base(double* p) { return p; }
MPC_DEV double lds(smem_t b, int off) { return b[off]; }
struct d2 { double x, y; };
MPC_DEV d2 lds2(smem_t b, int off) { d2 r; r.x = b[off]; r.y = b[off + 1]; return r; }
MPC_DEV void sts(smem_t b, int off, double v) { b[off] = v; }
MPC_DEV int launder(int v) { return v; }
}  // namespace mpcb200
