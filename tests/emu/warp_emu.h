// warp_emu.h -- TEST ONLY.  Single-threaded emulation of one thread block of 1..4 warps with
// ucontext coroutines, so that mkz_mpc_path_follower_b200/csrc/mpc_kernel.cuh can be compiled by
// g++ and stepped on a CPU.  Every warp collective is a rendezvous of the 32 lanes of the calling
// lane's warp (full mask), checked by call-site id; block barriers wait for every lane of the
// block.  Lanes run round-robin, each up to its next collective.  Never part of the shipped library.
#pragma once
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>
#include <unordered_map>

#define MPC_DEV inline
#define MPC_DEV_NOINLINE inline
#define MPC_FULL 0xffffffffu

namespace mpcb200 {
namespace emu {
#define MPC_EMU_MAX_LANES 128
struct Warp {
    ucontext_t main_ctx, ctx[MPC_EMU_MAX_LANES];
    char* stacks;
    int nl;           // lanes in the block (32 * warps)
    int cur;          // running lane (= thread index in the block)
    int done[MPC_EMU_MAX_LANES];
    double dslot[2][MPC_EMU_MAX_LANES];
    int islot[2][MPC_EMU_MAX_LANES];
    int site[2][MPC_EMU_MAX_LANES];
    int epoch[MPC_EMU_MAX_LANES];    // per-lane warp-collective counter
    long arrive[MPC_EMU_MAX_LANES];  // per-lane block-barrier counter
    long wsync[MPC_EMU_MAX_LANES];   // per-lane bar.warp.sync counter
    int bpred[2][MPC_EMU_MAX_LANES];
    void (*fn)(int lane, void* arg);
    void* arg;
};
extern thread_local Warp* W;

inline void yield_next() {
    Warp* w = W;
    int from = w->cur;
    int to = from;
    for (int i = 1; i <= w->nl; i++) { int c = (from + i) % w->nl; if (!w->done[c]) { to = c; break; } }
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
    const int w0 = w->cur & ~31;
    for (int i = w0; i < w0 + 32; i++) if (w->site[b][i] != site) {
        fprintf(stderr, "warp_emu: divergent collective (lane %d at site %d, lane %d at site %d)\n", w->cur, site, i, w->site[b][i]);
        abort();
    }
}
void run_warp(void (*fn)(int, void*), void* arg, int warps = 1);

// ---- shared-memory race check (stands in for compute-sanitizer racecheck, which is closed on the GPU
// pool): a lane may read or overwrite a word last written by ANOTHER lane only if a bar.warp.sync (same
// warp) or a block barrier separates the two accesses; shuffles are rendezvous points, not fences.
// Several lanes storing the same value to a word is not a race.  Words registered as benign (sinks of inactive lanes, the slot shared by threads that own no stage)
// are skipped.
struct ShadowWord { int lane; long wsync; long arrive; };
struct BenignSet { const double* base; long stride; int count; };   // words base + i * stride, i < count
struct RaceState {
    std::unordered_map<const double*, ShadowWord> last_write;
    BenignSet benign[64]; int n_benign;
    long races; int enabled;
};
extern thread_local RaceState RS;
inline void race_reset() { RS.last_write.clear(); RS.n_benign = 0; }
inline void race_benign(const double* base, long stride, int count) {
    if (RS.n_benign < 64) RS.benign[RS.n_benign++] = BenignSet{base, stride, count};
}
inline bool race_skip(const double* a) {
    for (int i = 0; i < RS.n_benign; i++) {
        const BenignSet& b = RS.benign[i];
        const long d = a - b.base;
        if (d >= 0 && d % b.stride == 0 && d / b.stride < b.count) return true;
    }
    return false;
}
inline void race_check(const double* a, bool is_write, double v = 0.0) {
    if (!RS.enabled || race_skip(a)) return;
    if (is_write && memcmp(a, &v, sizeof(double)) == 0) return;   // several lanes storing the same value: idempotent
    Warp* w = W; const int l = w->cur;
    auto it = RS.last_write.find(a);
    if (it != RS.last_write.end() && it->second.lane != l) {
        const ShadowWord& s = it->second;
        const bool same_warp = (s.lane >> 5) == (l >> 5);
        const bool fenced = (w->arrive[l] > s.arrive) || (same_warp && w->wsync[l] > s.wsync);
        if (!fenced) {
            if (RS.races < 10) fprintf(stderr, "warp_emu: shared-memory race on %p: lane %d %s a word lane %d wrote with no barrier in between\n",
                                       (const void*)a, l, is_write ? "overwrites" : "reads", s.lane);
            RS.races++;
        }
    }
    if (is_write) RS.last_write[a] = ShadowWord{l, w->wsync[l], w->arrive[l]};
}
// block barrier: every lane of the block arrives; lanes of other warps keep taking turns meanwhile
inline long block_arrive() {
    Warp* w = W; const int l = w->cur;
    const long gen = ++w->arrive[l];
    for (;;) {
        bool all = true;
        for (int i = 0; i < w->nl; i++) if (w->arrive[i] < gen) { all = false; break; }
        if (all) break;
        yield_next();
    }
    return gen;
}
}  // namespace emu

MPC_DEV int lane_id() { return emu::W->cur & 31; }
MPC_DEV int thread_in_block() { return emu::W->cur; }
MPC_DEV void block_sync() { emu::block_arrive(); }
MPC_DEV bool block_all(bool p) {
    emu::Warp* w = emu::W; const int l = w->cur;
    const int b = (int)((w->arrive[l] + 1) & 1);
    w->bpred[b][l] = p ? 1 : 0;
    emu::block_arrive();
    int r = 1; for (int i = 0; i < w->nl; i++) r &= w->bpred[b][i];
    return r != 0;
}
// l = lane within the warp, w0 = first lane of the warp
#define MPC_EMU_COLLECTIVE_D(expr_src)                                          \
    emu::Warp* w = emu::W; int t_ = w->cur; int b = w->epoch[t_] & 1;           \
    const int w0 = t_ & ~31; const int l = t_ & 31; (void)l;                    \
    emu::publish_d(v, __LINE__); emu::check_site(b, __LINE__);                  \
    w->epoch[t_]++; int s_ = (expr_src);
MPC_DEV double shfl(double v, int src) { MPC_EMU_COLLECTIVE_D(src & 31) return w->dslot[b][w0 + s_]; }
MPC_DEV double shfl_down(double v, int d) { MPC_EMU_COLLECTIVE_D(l + d) return s_ < 32 ? w->dslot[b][w0 + s_] : v; }
MPC_DEV double shfl_up(double v, int d) { MPC_EMU_COLLECTIVE_D(l - d) return s_ >= 0 ? w->dslot[b][w0 + s_] : v; }
MPC_DEV double shfl_xor(double v, int m) { MPC_EMU_COLLECTIVE_D(l ^ m) return w->dslot[b][w0 + s_]; }
MPC_DEV int shfl(int v, int src) {
    emu::Warp* w = emu::W; int l = w->cur; int b = w->epoch[l] & 1;
    emu::publish_i(v, __LINE__); emu::check_site(b, __LINE__); w->epoch[l]++;
    return w->islot[b][(l & ~31) + (src & 31)];
}
MPC_DEV void syncwarp() {
    emu::Warp* w = emu::W; int l = w->cur; int b = w->epoch[l] & 1;
    emu::publish_i(0, __LINE__); emu::check_site(b, __LINE__); w->epoch[l]++;
    w->wsync[l]++;
}
MPC_DEV bool warp_all(bool p) {
    emu::Warp* w = emu::W; int l = w->cur; int b = w->epoch[l] & 1;
    emu::publish_i(p ? 1 : 0, __LINE__); emu::check_site(b, __LINE__); w->epoch[l]++;
    int r = 1; for (int i = (l & ~31); i < (l & ~31) + 32; i++) r &= w->islot[b][i]; return r != 0;
}
MPC_DEV bool warp_any(bool p) {
    emu::Warp* w = emu::W; int l = w->cur; int b = w->epoch[l] & 1;
    emu::publish_i(p ? 1 : 0, __LINE__); emu::check_site(b, __LINE__); w->epoch[l]++;
    int r = 0; for (int i = (l & ~31); i < (l & ~31) + 32; i++) r |= w->islot[b][i]; return r != 0;
}
MPC_DEV void mpc_sincos(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }
MPC_DEV double fast_rcp(double x) { return 1.0 / x; }
MPC_DEV float fast_log2(float x) { return log2f(x); }
MPC_DEV float fast_exp2(float x) { return exp2f(x); }
MPC_DEV double mul_rn(double a, double b) { volatile double r = a * b; return r; }   // the emulator is built with -ffp-contract=off
MPC_DEV double add_rn(double a, double b) { volatile double r = a + b; return r; }
MPC_DEV void st_global_v2(double* p, double x, double y) { p[0] = x; p[1] = y; }
MPC_DEV double int2_as_double(int lo, int hi) { double d; int v[2] = {lo, hi}; memcpy(&d, v, 8); return d; }
// shared memory: a plain double array, offsets in doubles
typedef double* smem_t;
#define SO(x) (x)
MPC_DEV smem_t smem_base(double* p) { return p; }
MPC_DEV double lds(smem_t b, int off) { emu::race_check(b + off, false); return b[off]; }
struct d2 { double x, y; };
MPC_DEV d2 lds2(smem_t b, int off) { emu::race_check(b + off, false); emu::race_check(b + off + 1, false); d2 r; r.x = b[off]; r.y = b[off + 1]; return r; }
MPC_DEV void sts(smem_t b, int off, double v) { emu::race_check(b + off, true, v); b[off] = v; }
MPC_DEV void sts2(smem_t b, int off, double x, double y) { sts(b, off, x); sts(b, off + 1, y); }
MPC_DEV int launder(int v) { return v; }
MPC_DEV void ld_roles(const int* p, int* out, int n = 20) { for (int i = 0; i < n; i++) out[i] = p[i]; }
}  // namespace mpcb200
