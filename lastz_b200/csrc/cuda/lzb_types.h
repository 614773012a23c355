/*
 * lzb_types.h -- integer shorthands, sentinels and the class-reduced score set shared by every kernel.
 * No CUDA headers: the kernels' device code is also compiled for the host emulators under tests/warp_emu.
 */
#ifndef LZB_TYPES_H
#define LZB_TYPES_H
#include <stdint.h>
#include "../../../include/lastz_b200.h"

typedef uint8_t  u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef int32_t  s32;
typedef uint64_t u64;
typedef int64_t  s64;

#define LZB_NEG_INF ((s32)-1932735283)      /* dna_utilities.h:138 negInfinity */
#define LZB_MAX_CLASSES 32                  /* byte equivalence classes of the two score matrices */

/* score matrices reduced to byte classes: two bytes share a class iff their rows and columns
 * agree in BOTH scoring->sub and maskedScoring->sub, so sub[a][b] == subC[cls[a]][cls[b]] exactly */
struct lzb_scoring_dev {
    int  numClasses;
    u8   cls[256];
    s32  subC[LZB_MAX_CLASSES * LZB_MAX_CLASSES];
    s32  msubC[LZB_MAX_CLASSES * LZB_MAX_CLASSES];
    s32  gapOpen, gapExtend;
};

/* device helpers */
struct seed_dev {                            /* lzb_seed, flattened for kernels */
    int length, numParts;
    int shift[LZB_MAX_SEED_PARTS];
    u32 mask[LZB_MAX_SEED_PARTS];
};

#endif
