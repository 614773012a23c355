// cuda_emu.cpp -- scheduler of the host block emulator (TEST INFRASTRUCTURE), see cuda_emu.h
#include <ucontext.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "cuda_emu.h"

emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
unsigned long long emu_collectives = 0;

static const int MAXT = 1024;
static ucontext_t g_main, g_ctx[MAXT];
static char* g_stack[MAXT];
static bool g_done[MAXT];
static int g_cur = -1, g_nthreads = 0;
static std::function<void()> g_kernel;

struct group { int arrived; unsigned long gen; int line; unsigned long long slot[32]; };
static group g_block, g_warp[MAXT / 32];

int emu_lane(void) { return g_cur & 31; }

static void yield_to_main(void) { const int me = g_cur; swapcontext(&g_ctx[me], &g_main); g_cur = me; threadIdx.x = (unsigned)me; }

// wait until `need` threads of the group have arrived from the same line
static void rendezvous(group& g, int need, int line, const char* what) {
    if (g.arrived == 0) g.line = line;
    else if (g.line != line) { fprintf(stderr, "cuda_emu: divergent %s: thread %d at line %d, others at line %d\n", what, g_cur, line, g.line); abort(); }
    const unsigned long gen = g.gen;
    if (++g.arrived >= need) { g.arrived = 0; g.gen++; }
    else while (g.gen == gen) yield_to_main();
}

void emu_syncthreads(int line) {
    int live = 0; for (int t = 0; t < g_nthreads; t++) if (!g_done[t]) live++;
    rendezvous(g_block, live, line, "__syncthreads");
    emu_collectives++;
}

static void warp_check(int line) {
    const int w = g_cur >> 5;
    for (int l = 0; l < 32; l++) if (g_done[w * 32 + l]) { fprintf(stderr, "cuda_emu: lane %d of warp %d returned before a full-mask collective (line %d)\n", l, w, line); abort(); }
}

unsigned long long emu_warp_exchange(unsigned long long v, int src, int line) {
    warp_check(line);
    group& g = g_warp[g_cur >> 5];
    g.slot[g_cur & 31] = v; rendezvous(g, 32, line, "warp collective");
    const unsigned long long out = g.slot[src & 31];
    rendezvous(g, 32, -line, "warp collective");
    emu_collectives++;
    return out;
}

unsigned emu_warp_ballot(int pred, int line) {
    warp_check(line);
    group& g = g_warp[g_cur >> 5];
    g.slot[g_cur & 31] = pred ? 1 : 0; rendezvous(g, 32, line, "warp collective");
    unsigned m = 0; for (int l = 0; l < 32; l++) if (g.slot[l]) m |= 1u << l;
    rendezvous(g, 32, -line, "warp collective");
    emu_collectives++;
    return m;
}

static void trampoline(void) { const int me = g_cur; g_kernel(); g_done[me] = true; }

void emu_launch(unsigned grid, unsigned block, const std::function<void()>& kernel) {
    if (block > (unsigned)MAXT || (block & 31)) { fprintf(stderr, "cuda_emu: block size %u not supported\n", block); abort(); }
    g_kernel = kernel; g_nthreads = (int)block;
    gridDim = { grid, 1, 1 }; blockDim = { block, 1, 1 };
    for (unsigned b = 0; b < grid; b++) {
        blockIdx = { b, 0, 0 };
        g_block.arrived = 0; for (auto& w : g_warp) w.arrived = 0;
        for (int t = 0; t < g_nthreads; t++) {
            if (!g_stack[t]) g_stack[t] = (char*)malloc(512 << 10);
            g_done[t] = false; getcontext(&g_ctx[t]);
            g_ctx[t].uc_stack.ss_sp = g_stack[t]; g_ctx[t].uc_stack.ss_size = 512 << 10; g_ctx[t].uc_link = &g_main;
            makecontext(&g_ctx[t], trampoline, 0);
        }
        for (;;) {
            bool any = false;
            for (int t = 0; t < g_nthreads; t++) if (!g_done[t]) { any = true; g_cur = t; threadIdx = { (unsigned)t, 0, 0 }; swapcontext(&g_main, &g_ctx[t]); }
            if (!any) break;
        }
        if (g_block.arrived) { fprintf(stderr, "cuda_emu: block %u ended with %d threads parked in __syncthreads\n", b, g_block.arrived); abort(); }
        for (auto& w : g_warp) if (w.arrived) { fprintf(stderr, "cuda_emu: block %u ended with lanes parked in a warp collective\n", b); abort(); }
    }
    g_cur = -1;
}
