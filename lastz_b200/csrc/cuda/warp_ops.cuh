/*
 * warp_ops.cuh -- the warp-collective vocabulary of the warp-cooperative device functions.
 *
 * Under nvcc these are the sm_100a intrinsics.  With -DLZB_WARP_EMU (tests/warp_emu, plain g++)
 * the same source runs on the host: 32 coroutines stand in for the 32 lanes and every collective
 * is a rendezvous that also checks that all lanes arrived from the SAME source line -- what the
 * hardware requires of full-mask *_sync intrinsics.  This lets the warp-cooperative x-drop code
 * be tested bit for bit without a GPU (tests/test_xdrop_emu.py); it is not a CPU path of the
 * product: nothing in liblastz_b200.so is compiled with LZB_WARP_EMU.
 */
#ifndef LZB_WARP_OPS_CUH
#define LZB_WARP_OPS_CUH

#ifdef LZB_WARP_EMU
#include <stdint.h>
unsigned long long wemu_exchange(unsigned long long v, int srcLane, int line);   /* value lane srcLane passed */
unsigned wemu_ballot(int pred, int line);
int wemu_lane(void);
#define W_DEV static inline
#define W_FULL 0xFFFFFFFFu
template <class T> static inline T wemu_shfl(T v, int src, int line) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    unsigned long long raw = 0; __builtin_memcpy(&raw, &v, sizeof(T));
    raw = wemu_exchange(raw, src & 31, line);
    T out; __builtin_memcpy(&out, &raw, sizeof(T)); return out;
}
#define W_SHFL(v, src)      wemu_shfl((v), (int)(src), __LINE__)
/* like the hardware: a lane whose source is out of range keeps its own value */
#define W_SHFL_UP(v, d)     wemu_shfl((v), wemu_lane() >= (int)(d) ? wemu_lane() - (int)(d) : wemu_lane(), __LINE__)
#define W_SHFL_XOR(v, m)    wemu_shfl((v), wemu_lane() ^ (int)(m), __LINE__)
#define W_BALLOT(p)         wemu_ballot((p) ? 1 : 0, __LINE__)
#define W_FFS(x)            __builtin_ffs((int)(x))
#define W_POPC(x)           __builtin_popcount((unsigned)(x))
#define W_ATOMIC_ADD_ULL(p, v) ([&]() { unsigned long long o_ = *(p); *(p) += (v); return o_; }())
#else
#define W_DEV __device__ __forceinline__
#define W_FULL 0xFFFFFFFFu
#define W_SHFL(v, src)      __shfl_sync(W_FULL, (v), (int)(src))
#define W_SHFL_UP(v, d)     __shfl_up_sync(W_FULL, (v), (unsigned)(d))
#define W_SHFL_XOR(v, m)    __shfl_xor_sync(W_FULL, (v), (int)(m))
#define W_BALLOT(p)         __ballot_sync(W_FULL, (p))
#define W_FFS(x)            __ffs((int)(x))
#define W_POPC(x)           __popc((unsigned)(x))
#define W_ATOMIC_ADD_ULL(p, v) atomicAdd((p), (unsigned long long)(v))
#endif

template <class T> W_DEV T w_min(T a, T b) { return a < b ? a : b; }
template <class T> W_DEV T w_max(T a, T b) { return a > b ? a : b; }

#endif
