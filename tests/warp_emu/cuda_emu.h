// cuda_emu.h -- host BLOCK emulator for hand-written CUDA kernels (TEST INFRASTRUCTURE).
//
// Compiles a kernel's source as plain C++ and runs one thread block at a time: every CUDA thread is a
// ucontext coroutine; __syncthreads() is a rendezvous of the block's live threads, the *_sync warp
// intrinsics a rendezvous of the 32 lanes of the calling warp.  Each rendezvous checks that all
// participants arrived from the SAME source line and that no participant of a full-mask warp collective
// has already returned -- the conditions the hardware leaves undefined.  Shared memory is `static`
// storage (one block runs at a time).  This is how the Y-drop kernel is unit-tested bit for bit without
// a GPU (tests/test_ydrop_emu.py); it is not a CPU path of the product.
#ifndef CUDA_EMU_H
#define CUDA_EMU_H
#include <stdint.h>
#include <string.h>
#include <functional>

struct emu_dim3 { unsigned x, y, z; };
extern emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)

void emu_syncthreads(int line);
unsigned long long emu_warp_exchange(unsigned long long v, int srcLane, int line);   // value passed by lane srcLane
unsigned emu_warp_ballot(int pred, int line);
int emu_lane(void);
void emu_launch(unsigned grid, unsigned block, const std::function<void()>& kernel); // blocks run one after another
extern unsigned long long emu_collectives;

template <class T> static inline T emu_shfl(T v, int src, int line) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    unsigned long long raw = 0; memcpy(&raw, &v, sizeof(T));
    raw = emu_warp_exchange(raw, src & 31, line);
    T out; memcpy(&out, &raw, sizeof(T)); return out;
}
#define __syncthreads()               emu_syncthreads(__LINE__)
#define __threadfence()               ((void)0)
#define __syncwarp()                  ((void)emu_warp_ballot(1, __LINE__))     /* a warp rendezvous */
#define __shfl_sync(m, v, src)        emu_shfl((v), (int)(src), __LINE__)
#define __shfl_up_sync(m, v, d)       emu_shfl((v), emu_lane() >= (int)(d) ? emu_lane() - (int)(d) : emu_lane(), __LINE__)
#define __shfl_down_sync(m, v, d)     emu_shfl((v), emu_lane() + (int)(d) < 32 ? emu_lane() + (int)(d) : emu_lane(), __LINE__)
#define __shfl_xor_sync(m, v, x)      emu_shfl((v), emu_lane() ^ (int)(x), __LINE__)
#define __ballot_sync(m, p)           emu_warp_ballot((p) ? 1 : 0, __LINE__)

// warp reductions: every lane contributes, every lane gets the result
template <class T, class F> static inline T emu_reduce(T v, F f, int line) {
    for (int d = 16; d > 0; d >>= 1) { T o = emu_shfl(v, emu_lane() ^ d, line); v = f(v, o); }
    return v;
}
#define __reduce_max_sync(m, v)  emu_reduce((v), [](decltype(v) a, decltype(v) b) { return a > b ? a : b; }, __LINE__)
#define __reduce_min_sync(m, v)  emu_reduce((v), [](decltype(v) a, decltype(v) b) { return a < b ? a : b; }, __LINE__)
#define __reduce_add_sync(m, v)  emu_reduce((v), [](decltype(v) a, decltype(v) b) { return (decltype(v))(a + b); }, __LINE__)

static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }

// CUDA's integer min/max overloads
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }

// vector types the kernels use
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 v = { x, y, z, w }; return v; }
#endif
