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

typedef uint8_t  u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef int32_t  s32;
typedef uint64_t u64;
typedef int64_t  s64;

#define LZB_NEG_INF ((s32)-1932735283)      /* dna_utilities.h:138 negInfinity */
#define LZB_MAX_CLASSES 32                  /* byte equivalence classes of the two score matrices */

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

/* score matrices reduced to byte classes: two bytes share a class iff their rows and columns
 * agree in BOTH scoring->sub and maskedScoring->sub, so sub[a][b] == subC[cls[a]][cls[b]] exactly */
struct lzb_scoring_dev {
    int  numClasses;
    u8   cls[256];
    s32  subC[LZB_MAX_CLASSES * LZB_MAX_CLASSES];
    s32  msubC[LZB_MAX_CLASSES * LZB_MAX_CLASSES];
    s32  gapOpen, gapExtend;
};

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
    void* seedScratch;                       /* seed_search.cu: device scratch kept across calls */
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

/* device helpers */
struct seed_dev {                            /* lzb_seed, flattened for kernels */
    int length, numParts;
    int shift[LZB_MAX_SEED_PARTS];
    u32 mask[LZB_MAX_SEED_PARTS];
};
static inline void seed_to_dev(seed_dev* d, const lzb_seed* s) {
    d->length = s->length; d->numParts = s->numParts;
    for (int i = 0; i < s->numParts; i++) { d->shift[i] = s->shift[i]; d->mask[i] = s->mask[i]; }
}

int  lzb_upload_classes(lzb_ctx*, const u8* h_seq, u32 len, u8** d_seq, u8** d_cls, size_t* cap);

#endif
