// cuda_emu.cpp -- scheduler of the host block emulator (TEST INFRASTRUCTURE), see cuda_emu.h
#include <ucontext.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "cuda_emu.h"

emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
unsigned long long emu_collectives = 0;

static const int MAXT = 1024;
static char* g_stack[MAXT];
static const size_t STACK_BYTES = 512 << 10;

// Context switching.  glibc's swapcontext makes a sigprocmask system call per switch (~0.3 us) and the
// Y-drop kernel test performs some 10^8 switches, so x86-64 gets a bare callee-saved-register switch.
#if defined(__x86_64__)
extern "C" void emu_switch(void** saveSp, void* newSp);
asm(".text\n.globl emu_switch\n.type emu_switch,@function\nemu_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size emu_switch,.-emu_switch\n");
static void* g_mainSp; static void* g_sp[MAXT];
static void thread_entry(void);
static void ctx_make(int t) {
    void** top = (void**)(((uintptr_t)(g_stack[t] + STACK_BYTES)) & ~(uintptr_t)15);
    top[-1] = NULL;                      // where a return address would sit: keeps the entry frame 16-byte aligned
    top[-2] = (void*)thread_entry;       // `ret` target
    for (int k = 3; k <= 8; k++) top[-k] = NULL;   // rbp rbx r12 r13 r14 r15
    g_sp[t] = &top[-8];
}
static inline void ctx_to_thread(int t) { emu_switch(&g_mainSp, g_sp[t]); }
static inline void ctx_to_main(int me) { emu_switch(&g_sp[me], g_mainSp); }
#else
static ucontext_t g_main, g_ctx[MAXT];
static void thread_entry(void);
static void ctx_make(int t) {
    getcontext(&g_ctx[t]);
    g_ctx[t].uc_stack.ss_sp = g_stack[t]; g_ctx[t].uc_stack.ss_size = STACK_BYTES; g_ctx[t].uc_link = &g_main;
    makecontext(&g_ctx[t], thread_entry, 0);
}
static inline void ctx_to_thread(int t) { swapcontext(&g_main, &g_ctx[t]); }
static inline void ctx_to_main(int me) { swapcontext(&g_ctx[me], &g_main); }
#endif
static bool g_done[MAXT];
static int g_cur = -1, g_nthreads = 0;
static std::function<void()> g_kernel;
static unsigned long long g_progress = 0;        // bumped on every arrival and every thread exit

struct group { int arrived; unsigned long gen; int line; unsigned long long slot[32]; };
static group g_block, g_warp[MAXT / 32];
static group* g_waitGroup[MAXT]; static unsigned long g_waitGen[MAXT];   // what a parked thread waits for (the scheduler skips it until then)

int emu_lane(void) { return g_cur & 31; }

static void yield_to_main(void) { const int me = g_cur; ctx_to_main(me); g_cur = me; threadIdx.x = (unsigned)me; }

// wait until `need` threads of the group have arrived from the same line
static void rendezvous(group& g, int need, int line, const char* what) {
    if (g.arrived == 0) g.line = line;
    else if (g.line != line) { fprintf(stderr, "cuda_emu: divergent %s: thread %d at line %d, others at line %d\n", what, g_cur, line, g.line); abort(); }
    const unsigned long gen = g.gen; g_progress++;
    if (++g.arrived >= need) { g.arrived = 0; g.gen++; }
    else { const int me = g_cur; g_waitGroup[me] = &g; g_waitGen[me] = gen; while (g.gen == gen) yield_to_main(); g_waitGroup[me] = NULL; }
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

static void thread_entry(void) { const int me = g_cur; g_kernel(); g_done[me] = true; g_progress++; for (;;) ctx_to_main(me); }   // never returns

void emu_launch(unsigned grid, unsigned block, const std::function<void()>& kernel) {
    if (block > (unsigned)MAXT || (block & 31)) { fprintf(stderr, "cuda_emu: block size %u not supported\n", block); abort(); }
    g_kernel = kernel; g_nthreads = (int)block;
    gridDim = { grid, 1, 1 }; blockDim = { block, 1, 1 };
    for (unsigned b = 0; b < grid; b++) {
        blockIdx = { b, 0, 0 };
        g_block.arrived = 0; for (auto& w : g_warp) w.arrived = 0;
        for (int t = 0; t < g_nthreads; t++) {
            if (!g_stack[t]) g_stack[t] = (char*)malloc(STACK_BYTES);
            g_done[t] = false; g_waitGroup[t] = NULL; ctx_make(t);
        }
        for (;;) {
            bool any = false; const unsigned long long before = g_progress;
            for (int t = 0; t < g_nthreads; t++) if (!g_done[t]) {
                any = true;
                if (g_waitGroup[t] && g_waitGroup[t]->gen == g_waitGen[t]) continue;      // still parked
                g_cur = t; threadIdx = { (unsigned)t, 0, 0 }; ctx_to_thread(t);
            }
            if (!any) break;
            if (g_progress == before) { fprintf(stderr, "cuda_emu: deadlock in block %u: every live thread is parked (a barrier some thread never reaches)\n", b); abort(); }
        }
        if (g_block.arrived) { fprintf(stderr, "cuda_emu: block %u ended with %d threads parked in __syncthreads\n", b, g_block.arrived); abort(); }
        for (auto& w : g_warp) if (w.arrived) { fprintf(stderr, "cuda_emu: block %u ended with lanes parked in a warp collective\n", b); abort(); }
    }
    g_cur = -1;
}
