// wemu.cpp -- host lane emulator behind lastz_b200/csrc/cuda/warp_ops.cuh (TEST INFRASTRUCTURE).
// 32 ucontext coroutines are the 32 lanes of one warp; a collective is a rendezvous of all 32 and
// aborts if the lanes did not arrive from the same source line or if a lane has already returned
// (both are undefined behaviour for full-mask *_sync intrinsics on the device).
#include <ucontext.h>
#include <stdio.h>
#include <stdlib.h>
#include <functional>
#include "wemu.h"

static const int NL = 32;
static ucontext_t g_main, g_ctx[NL];
static char* g_stack[NL];
static bool g_done[NL];
static int g_cur = -1, g_arrived = 0, g_line[NL];
static unsigned long g_generation = 0;
static unsigned long long g_slot[NL];
static std::function<void(int)> g_body;
unsigned long long wemu_collectives = 0;

int wemu_lane(void) { return g_cur; }

static void yield_to_main(void) { int me = g_cur; swapcontext(&g_ctx[me], &g_main); g_cur = me; }

static void rendezvous(int line) {
    const int me = g_cur;
    for (int l = 0; l < NL; l++) if (g_done[l]) { fprintf(stderr, "wemu: lane %d returned while lane %d waits in a collective (line %d)\n", l, me, line); abort(); }
    g_line[me] = line;
    const unsigned long gen = g_generation;
    if (++g_arrived == NL) {
        for (int l = 0; l < NL; l++) if (g_line[l] != line) { fprintf(stderr, "wemu: divergent collective: lane %d at line %d, lane %d at line %d\n", me, line, l, g_line[l]); abort(); }
        g_arrived = 0; g_generation++;
    } else while (g_generation == gen) yield_to_main();
}

unsigned long long wemu_exchange(unsigned long long v, int src, int line) {
    g_slot[g_cur] = v; rendezvous(line);
    const unsigned long long out = g_slot[src & 31];
    rendezvous(-line);
    wemu_collectives++;
    return out;
}

unsigned wemu_ballot(int pred, int line) {
    g_slot[g_cur] = pred ? 1 : 0; rendezvous(line);
    unsigned m = 0; for (int l = 0; l < NL; l++) if (g_slot[l]) m |= 1u << l;
    rendezvous(-line);
    wemu_collectives++;
    return m;
}

static void trampoline(void) { const int me = g_cur; g_body(me); g_done[me] = true; }

void wemu_run(const std::function<void(int)>& body) {
    g_body = body; g_arrived = 0;
    for (int l = 0; l < NL; l++) {
        if (!g_stack[l]) g_stack[l] = (char*)malloc(1 << 20);
        g_done[l] = false; getcontext(&g_ctx[l]);
        g_ctx[l].uc_stack.ss_sp = g_stack[l]; g_ctx[l].uc_stack.ss_size = 1 << 20; g_ctx[l].uc_link = &g_main;
        makecontext(&g_ctx[l], trampoline, 0);
    }
    for (;;) {
        bool any = false;
        for (int l = 0; l < NL; l++) if (!g_done[l]) { any = true; g_cur = l; swapcontext(&g_main, &g_ctx[l]); }
        if (!any) break;
    }
    if (g_arrived) { fprintf(stderr, "wemu: warp finished with %d lanes parked in a collective\n", g_arrived); abort(); }
    g_cur = -1;
}
