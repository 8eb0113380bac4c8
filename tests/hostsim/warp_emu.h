/*
 * warp_emu.h -- TEST INFRASTRUCTURE.  Runs device code written against warp collectives (csrc/obm_warp_core.h) on
 * the host: 32 fibers (ucontext) in lock step.  A collective deposits the caller's value, yields until all 32 lanes
 * have arrived, then every lane reads the exchanged values -- the semantics of the *_sync intrinsics with a full
 * mask.  Fibers only switch inside collectives, so everything between two collectives runs lane after lane, in any
 * interleaving the device could show for properly synchronised code.
 */
#pragma once
#include <ucontext.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

namespace wemu {

struct Warp {
    ucontext_t main_ctx, ctx[32];
    std::vector<char> stacks[32];
    bool done[32];
    int cur = 0;
    uint64_t buf[2][32];
    uint32_t arrived = 0, gen = 0;
    std::function<void()> body;
};
inline Warp *&current() { static Warp *w = nullptr; return w; }

inline void trampoline() {
    Warp *w = current();
    w->body();
    w->done[w->cur] = true;
    swapcontext(&w->ctx[w->cur], &w->main_ctx);
}
/* runs body() once per lane, lanes switching at collectives */
inline void run(Warp &w, std::function<void()> body) {
    Warp *saved = current();
    current() = &w;
    w.body = std::move(body);
    w.arrived = 0; w.gen = 0;
    for (int l = 0; l < 32; l++) {
        if (w.stacks[l].empty()) w.stacks[l].resize(512 * 1024);
        w.done[l] = false;
        getcontext(&w.ctx[l]);
        w.ctx[l].uc_stack.ss_sp = w.stacks[l].data();
        w.ctx[l].uc_stack.ss_size = w.stacks[l].size();
        w.ctx[l].uc_link = &w.main_ctx;
        makecontext(&w.ctx[l], (void (*)())trampoline, 0);
    }
    for (;;) {
        bool any = false;
        for (int l = 0; l < 32; l++) {
            if (w.done[l]) continue;
            any = true;
            w.cur = l;
            swapcontext(&w.main_ctx, &w.ctx[l]);
        }
        if (!any) break;
        bool all_done = true, none_done = true;
        for (int l = 0; l < 32; l++) { all_done &= w.done[l]; none_done &= !w.done[l]; }
        if (!all_done && !none_done) {
            /* some lanes finished while others wait in a collective: a divergent collective (a bug in the kernel) */
            bool waiting = w.arrived != 0;
            if (waiting) { fprintf(stderr, "warp_emu: collective reached by only part of the warp\n"); abort(); }
        }
    }
    current() = saved;
}
inline uint32_t lane() { return (uint32_t)current()->cur; }
/* the exchange primitive: returns the 32 deposited values of this collective */
inline const uint64_t *exchange(uint64_t v) {
    Warp *w = current();
    const uint32_t g = w->gen;
    w->buf[g & 1][w->cur] = v;
    if (++w->arrived == 32) { w->arrived = 0; w->gen++; }
    else while (w->gen == g) swapcontext(&w->ctx[w->cur], &w->main_ctx);
    return w->buf[g & 1];
}
inline uint32_t ballot(bool p) { const uint64_t *a = exchange(p ? 1 : 0); uint32_t m = 0; for (int l = 0; l < 32; l++) m |= (uint32_t)(a[l] & 1) << l; return m; }
inline uint32_t shfl(uint32_t v, uint32_t src) { const uint64_t *a = exchange(v); return (uint32_t)a[src & 31]; }
inline uint32_t shfl_up(uint32_t v, uint32_t d) { const uint32_t me = lane(); const uint64_t *a = exchange(v); return me >= d ? (uint32_t)a[me - d] : v; }
inline void sync() { exchange(0); }

} /* namespace wemu */
