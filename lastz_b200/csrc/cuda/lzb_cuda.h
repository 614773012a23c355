/*
 * lzb_cuda.h -- internals shared by the sm_100a implementation of include/lastz_b200.h.
 * Nothing in here crosses the C-ABI.
 */
#ifndef LZB_CUDA_H
#define LZB_CUDA_H

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../../include/lastz_b200.h"

#include "lzb_types.h"

int lzb_fail(const char* fmt, ...);         /* sets lzb_last_error(), returns -1 */

#define CUDA_TRY(expr)                                                                    \
    do { cudaError_t e_ = (expr);                                                         \
         if (e_ != cudaSuccess) { lzb_fail("%s failed: %s (%s:%d)", #expr,                \
                                           cudaGetErrorString(e_), __FILE__, __LINE__);   \
                                  return -1; } } while (0)
#define CUDA_TRYP(expr)                                                                   \
    do { cudaError_t e_ = (expr);                                                         \
         if (e_ != cudaSuccess) { lzb_fail("%s failed: %s (%s:%d)", #expr,                \
                                           cudaGetErrorString(e_), __FILE__, __LINE__);   \
                                  return NULL; } } while (0)

struct lzb_ctx {
    int device;
    cudaStream_t stream;
    int smCount;
    bool haveScoring;
    s32* hostSub;  s32* hostMsub;            /* host copies (entropy/rescoring bookkeeping) */
    lzb_scoring_dev sc;                      /* host mirror */
    lzb_scoring_dev* d_sc;                   /* device copy */
    u64 launches;                            /* kernels launched by this context */
    void* gappedCache;                       /* gapped.cu: speculation lanes kept across calls */
    void* peaksBuf; size_t peaksCap;         /* gapped.cu: anchor table of lzb_reduce_to_points */
    void* seedScratch;                       /* seed_search.cu: device scratch kept across calls */
    unsigned candCapWanted;                  /* seed_search.cu: HSP candidate capacity a call found it needed (0: the default) */
    struct { u8* seq; u8* cls; size_t cap; } qpool[4];   /* device buffers of freed queries, reused by the next load
                                                (cudaFree is a device-wide synchronisation: 0.1 s with 32 lanes' buffers live) */
};

void lzb_gapped_cache_free(lzb_ctx*);        /* gapped.cu */
void lzb_seed_scratch_free(lzb_ctx*);        /* seed_search.cu */

struct lzb_target {
    lzb_ctx* ctx;
    u8* h_seq;  u32 len;                     /* host copy, NUL terminated */
    u8* d_seq;                               /* ASCII bytes in HBM, len+1 (+pad) */
    u8* d_cls;                               /* class codes in HBM */
    u32 start, end, step;
    int wordBits, seedLength;
    u32* d_off;                              /* CSR offsets [2^wordBits + 1] */
    u32* d_pos;                              /* positions, per word in DEcreasing order */
    u64 npos;
};

struct lzb_query {
    lzb_ctx* ctx;
    u8* h_seq;  u32 len;
    u8* d_seq;  u8* d_cls;  size_t cap;
};

static inline void seed_to_dev(seed_dev* d, const lzb_seed* s) {
    d->length = s->length; d->numParts = s->numParts;
    for (int i = 0; i < s->numParts; i++) { d->shift[i] = s->shift[i]; d->mask[i] = s->mask[i]; }
}

int  lzb_upload_classes(lzb_ctx*, const u8* h_seq, u32 len, u8** d_seq, u8** d_cls, size_t* cap);

#endif
