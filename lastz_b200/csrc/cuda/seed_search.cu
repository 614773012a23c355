/*
 * seed_search.cu -- K2/K3: seed-hit enumeration, diagonal-hash filtering and x-drop extension.
 *
 * Replaces seed_hit_search / private_hit_search / find_table_matches (seed_search.c:322, :464,
 * :810), the default hit processor process_for_simple_hit (:1056) with the diag-hash protocol of
 * diag_hash.h:61-101, xdrop_extend_seed_hit (:2528) and the collect_hsps reporter
 * (lastz.c:3991).
 *
 * The reference handles hits one at a time in discovery order (query position up, seed variant,
 * target position down) and threads every hit through diagEnd[(pos1-pos2) & 0xFFFF].  That state
 * is the ONLY coupling between hits, and it couples only hits in the same bucket.  So:
 *
 *   k_query_words   pack the seed word at every query position                    (hot loop B)
 *   k_count_hits    hits per 1024 query positions -> host cuts the query into chunks that fit
 *   per chunk:
 *     k_slot_count  hits per (position, variant) slot; cub exclusive scan
 *     k_expand      CSR lists -> (pos1,pos2) records in discovery order, key = bucket  (loop C)
 *     cub radix sort on the bucket bits, STABLE => each bucket keeps discovery order
 *     k_bucket_bounds
 *     k_extend      one warp per bucket: 32 hits at a time; right x-drop scans in parallel (they
 *                   do not depend on the bucket state), a 32-step register walk that replays the
 *                   diagEnd test/update exactly, then left scans for the surviving hits (loop D)
 *
 * diagEnd persists in HBM across chunks.  HSP candidates come back with the integer match counts
 * the entropy factor needs; the double/libm part of entropy (dna_utilities.c:2926-2936) and the
 * final ordering are O(#HSPs) host work in lzb_seed_hit_search.
 */
#include <cub/cub.cuh>
#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <vector>
#include "lzb_cuda.h"

#define POS_PER_BLOCK 1024            /* granularity of the chunk planner */
#define MAX_VARIANTS 512

struct ctb_dev2 { int8_t v[256]; };

struct sp_dev {
    u32 qstart, qend, len1, len2;
    int L, V, hashBits;
    int selfCompare, sameStrand;
    s32 xDrop, K;
    int gfExtend, plain, entropy;
};

#include "xdrop_warp.cuh"             /* cand_rec + the warp-cooperative bucket replay */

struct search_counters {
    unsigned long long words, extensions, bpExtended, ncand, overflow;
};

/* ---- K2a: the packed seed word at every query position (invalid => 0xFFFFFFFF) ---- */
__global__ void k_query_words(const u8* __restrict__ seq, sp_dev P, seed_dev sd, ctb_dev2 ctb,
                              u32* __restrict__ qword, search_counters* cnt) {
    u32 n = P.qend - P.qstart;
    unsigned long long valid = 0;
    for (u64 k = blockIdx.x * (u64)blockDim.x + threadIdx.x; k < n; k += (u64)gridDim.x * blockDim.x) {
        u32 i = P.qstart + (u32)k;                 /* index of the last base of the window */
        u32 word = 0xFFFFFFFFu;
        if (i + 1 >= P.qstart + (u32)sd.length) {
            u64 w = 0; bool ok = true;
            u32 first = i + 1 - (u32)sd.length;
            for (int j = 0; j < sd.length; j++) {
                int b = ctb.v[seq[first + j]];
                ok = ok && (b >= 0);
                w = (w << 2) | (u64)(b & 3);
            }
            if (ok) {
                word = 0;
                for (int p = 0; p < sd.numParts; p++) word |= (u32)(w >> sd.shift[p]) & sd.mask[p];
                valid++;
            }
        }
        qword[k] = word;
    }
    valid = __reduce_add_sync(0xFFFFFFFFu, (unsigned)valid);
    if ((threadIdx.x & 31) == 0 && valid) atomicAdd(&cnt->words, valid);
}

/* the part of a word's position list a hit at pos2 may use.  Lists are in decreasing position
 * order; --self keeps only positions below a limit (seed_hit_below_diagonal seed_search.c:2182),
 * i.e. a suffix of the list, found by binary search. */
__device__ __forceinline__ void list_range(const u32* __restrict__ off, const u32* __restrict__ pos,
                                           u32 key, u32 pos2, const sp_dev& P, u32& lo, u32& n) {
    u32 a = off[key], b = off[key + 1];
    if (P.selfCompare && b > a) {
        s64 limit = P.sameStrand ? (s64)pos2 : (s64)P.len2 - 1 - (s64)pos2 + 2 * (s64)P.L;
        u32 x = a, y = b;                           /* first index whose position < limit */
        while (x < y) { u32 m = (x + y) >> 1; if ((s64)pos[m] < limit) y = m; else x = m + 1; }
        a = x;
    }
    lo = a; n = b - a;
}

/* ---- K2b: hits per POS_PER_BLOCK query positions (for the chunk planner) ---- */
__global__ void k_count_hits(const u32* __restrict__ qword, const u32* __restrict__ off,
                             const u32* __restrict__ pos, const u32* __restrict__ flips, sp_dev P,
                             unsigned long long* __restrict__ blkcnt) {
    u32 n = P.qend - P.qstart;
    u32 base = blockIdx.x * POS_PER_BLOCK;
    unsigned long long mine = 0;
    for (u32 k = base + threadIdx.x; k < base + POS_PER_BLOCK && k < n; k += blockDim.x) {
        u32 w = qword[k];
        if (w == 0xFFFFFFFFu) continue;
        u32 pos2 = P.qstart + k + 1;
        for (int v = 0; v < P.V; v++) { u32 lo, c; list_range(off, pos, w ^ flips[v], pos2, P, lo, c); mine += c; }
    }
    typedef cub::BlockReduce<unsigned long long, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    unsigned long long tot = BR(tmp).Sum(mine);
    if (threadIdx.x == 0) blkcnt[blockIdx.x] = tot;
}

/* ---- K2c: hits per slot inside one chunk; slot = (k - k0) * V + v ---- */
__global__ void k_slot_count(const u32* __restrict__ qword, const u32* __restrict__ off,
                             const u32* __restrict__ pos, const u32* __restrict__ flips, sp_dev P,
                             u32 k0, u32 nslots, u32* __restrict__ slotcnt) {
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += gridDim.x * blockDim.x) {
        u32 k = k0 + s / (u32)P.V; int v = (int)(s % (u32)P.V);
        u32 w = qword[k], c = 0;
        if (w != 0xFFFFFFFFu) { u32 lo; list_range(off, pos, w ^ flips[v], P.qstart + k + 1, P, lo, c); }
        slotcnt[s] = c;
    }
}

/* ---- K2d: expand the CSR lists into hit records, in discovery order ----
 * A warp owns 32 consecutive slots and copies their lists as one concatenated array, so lanes
 * stay busy whatever the individual list lengths are.  key = diag-hash bucket. */
__global__ void k_expand(const u32* __restrict__ qword, const u32* __restrict__ off,
                         const u32* __restrict__ pos, const u32* __restrict__ flips, sp_dev P,
                         u32 k0, u32 nslots, const u32* __restrict__ slotoff,
                         u32* __restrict__ keys, u64* __restrict__ vals) {
    const u32 lane = threadIdx.x & 31;
    const u32 hmask = (1u << P.hashBits) - 1;
    u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 s0 = warp * 32; s0 < nslots; s0 += nwarps * 32) {
        u32 s = s0 + lane;
        u32 base = 0, lo = 0, pos2 = 0, c = 0;
        if (s < nslots) {
            u32 k = k0 + s / (u32)P.V; int v = (int)(s % (u32)P.V);
            u32 w = qword[k];
            pos2 = P.qstart + k + 1;
            base = slotoff[s];
            if (w != 0xFFFFFFFFu) list_range(off, pos, w ^ flips[v], pos2, P, lo, c);
        }
        u32 wbase = __shfl_sync(0xFFFFFFFFu, base, 0);
        u32 last = min(s0 + 31, nslots - 1) - s0;
        u32 total = __shfl_sync(0xFFFFFFFFu, base + c, last) - wbase;
        u32 rel = (s < nslots) ? base - wbase : 0xFFFFFFFFu;   /* start of my list in the warp's array */
        for (u32 t = lane; t < ((total + 31) & ~31u); t += 32) {
            /* owner = last lane whose rel <= t (rel is non-decreasing over lanes) */
            u32 owner = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                u32 cand = owner + step;
                u32 r = __shfl_sync(0xFFFFFFFFu, rel, cand & 31);
                if (cand < 32 && r <= t) owner = cand;
            }
            u32 orel = __shfl_sync(0xFFFFFFFFu, rel, owner);
            u32 olo  = __shfl_sync(0xFFFFFFFFu, lo, owner);
            u32 op2  = __shfl_sync(0xFFFFFFFFu, pos2, owner);
            if (t < total) {
                u32 p1 = pos[olo + (t - orel)];
                keys[wbase + t] = (p1 - op2) & hmask;
                vals[wbase + t] = ((u64)op2 << 32) | p1;
            }
        }
    }
}

/* ---- K3a: first record of every bucket in the sorted array ---- */
__global__ void k_bucket_bounds(const u32* __restrict__ keys, u32 nhits, u32 nbuckets, u32* __restrict__ bstart) {
    for (u32 h = blockIdx.x * blockDim.x + threadIdx.x; h <= nbuckets; h += gridDim.x * blockDim.x) {
        u32 x = 0, y = nhits;                       /* lower_bound(keys, h) */
        while (x < y) { u32 m = (x + y) >> 1; if (keys[m] < h) x = m + 1; else y = m; }
        bstart[h] = x;
    }
}

/* ---- K3b: replay each bucket in discovery order; x-drop extension ----
 *
 * The scans read the class-coded sequences eight bases at a time (two aligned 64-bit loads and a
 * funnel shift per sequence) instead of one byte per column: a warp's 32 lanes sit on 32
 * unrelated diagonals, so every load instruction costs up to 32 L1 wavefronts whatever its width,
 * and the wavefront pipe -- not HBM -- was the first limit (profiles/r01_k_extend_before.txt).
 * With at most 16 byte classes (any DNA scoring set) the eight (row, column) class pairs of a
 * chunk are formed by one shift+or and looked up in a 256-entry table. */
__device__ __forceinline__ u64 ld8(const u8* __restrict__ p, u32 idx) {
    const u64* w = (const u64*)(p + (idx & ~7u));
    u32 sh = (idx & 7u) * 8u;
    u64 lo = w[0], hi = w[1];
    return sh ? (lo >> sh) | (hi << (64u - sh)) : lo;
}

template <bool SMALL>
__device__ __forceinline__ s32 pair_score(const s32* __restrict__ lut, u64 x1, u64 x2, u64 pr, int i) {
    if (SMALL) return lut[(u32)(pr >> (8 * i)) & 255u];
    return lut[((u32)(x1 >> (8 * i)) & 255u) * LZB_MAX_CLASSES + ((u32)(x2 >> (8 * i)) & 255u)];
}

template <bool SMALL>
__global__ void __launch_bounds__(256, 3)
k_extend(const u64* __restrict__ hits, const u32* __restrict__ bstart, u32 nbuckets,
         const u8* __restrict__ cls1, const u8* __restrict__ cls2,
         const u8* __restrict__ asc1, const u8* __restrict__ asc2,
         const lzb_scoring_dev* __restrict__ sc, sp_dev P, u32* __restrict__ diagEnd,
         cand_rec* __restrict__ cand, u32 candCap, search_counters* cnt) {
    __shared__ s32 lut[SMALL ? 256 : LZB_MAX_CLASSES * LZB_MAX_CLASSES];
    if (SMALL) { for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = sc->msubC[(i >> 4) * LZB_MAX_CLASSES + (i & 15)]; }
    else { for (int i = threadIdx.x; i < LZB_MAX_CLASSES * LZB_MAX_CLASSES; i += blockDim.x) lut[i] = sc->msubC[i]; }
    __syncthreads();
    const u32 lane = threadIdx.x & 31;
    const u32 L = (u32)P.L;
    const s32 xDrop = P.xDrop;
    u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long nExt = 0, nBp = 0;
    for (u32 h = warp; h < nbuckets; h += nwarps) {
        u32 b0 = bstart[h], b1 = bstart[h + 1];
        if (b0 == b1) continue;
        u32 E = diagEnd[h];
        for (u32 base = b0; base < b1; base += 32) {
            u32 idx = base + lane;
            bool have = idx < b1;
            u64 rec = have ? hits[idx] : 0;
            u32 pos1 = (u32)rec, pos2 = (u32)(rec >> 32);
            s64 diag = (s64)pos1 - (s64)pos2;
            if (P.plain) {                          /* process_for_plain_hit seed_search.c:995 */
                if (have) {
                    u32 slot = (u32)atomicAdd(&cnt->ncand, 1ull);
                    if (slot < candCap) { cand_rec r = { pos1, pos2, pos1 - L, pos2 - L, L, 0, 0, 0, 0, 0 }; cand[slot] = r; }
                }
                continue;
            }
            /* right scan (seed_search.c:2663-2693): independent of the bucket state.  Hits the
             * bucket has already passed (diagEnd only grows) are skipped outright. */
            bool maybe = have && !(E > pos2 - L);
            u32 rightLen = 0, rightCols = 0; s32 rightScore = 0;        /* best prefix length, columns examined */
            if (maybe && P.gfExtend == LZB_GFEX_XDROP) {
                s64 lim = (s64)P.len2 + diag;
                u32 rstop = ((s64)P.len1 <= lim) ? P.len1 : (u32)lim;
                u32 avail = rstop > pos1 ? rstop - pos1 : 0;
                s32 run = 0; bool going = true;
                while (going && rightCols < avail) {
                    u64 x1 = ld8(cls1, pos1 + rightCols), x2 = ld8(cls2, pos2 + rightCols);
                    u64 pr = (x1 << 4) | x2;
                    u32 n = avail - rightCols; if (n > 8) n = 8;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        if (going && (u32)i < n) {
                            if (run >= rightScore - xDrop) {
                                run += pair_score<SMALL>(lut, x1, x2, pr, i);
                                rightCols++;
                                if (run > rightScore) { rightScore = run; rightLen = rightCols; }
                            } else going = false;
                        }
                    }
                }
            }
            u32 rightBlock = pos1 + rightCols;
            u32 ext = (P.gfExtend == LZB_GFEX_XDROP) ? (u32)((s64)rightBlock - diag) : pos2;
            /* replay process_for_simple_hit's test/update (seed_search.c:1113, :2785-2789) in
             * discovery order: lane k sees the bucket exactly as hit k would have */
            bool live = false; u32 myStop = 0;
            u32 active = __ballot_sync(0xFFFFFFFFu, maybe);
            while (active) {
                int k = __ffs(active) - 1; active &= active - 1;
                u32 p2 = __shfl_sync(0xFFFFFFFFu, pos2, k);
                u32 ex = __shfl_sync(0xFFFFFFFFu, ext, k);
                bool lv = !(E > p2 - L);
                if ((int)lane == k) { live = lv; myStop = E; }
                if (lv && ex > E) E = ex;
            }
            if (!live) continue;
            if (P.gfExtend != LZB_GFEX_XDROP) {     /* --nogfextend with gapped stage: raw hit, score 0 */
                u32 slot = (u32)atomicAdd(&cnt->ncand, 1ull);
                if (slot < candCap) { cand_rec r = { pos1, pos2, pos1 - L, pos2 - L, L, 0, 0, 0, 0, 0 }; cand[slot] = r; }
                continue;
            }
            /* left scan (seed_search.c:2598-2632), blocked by the bucket's previous extent */
            s64 blk = (s64)myStop + diag;
            u32 stop = blk > 0 ? (u32)blk : 0;
            u32 leftCols = 0, leftLen = 0; s32 leftScore = 0;
            {
                u32 avail = pos1 > stop ? pos1 - stop : 0;
                s32 run = 0; bool going = true;
                while (going && leftCols < avail) {
                    u32 a = pos1 - leftCols, b = pos2 - leftCols;       /* columns a-1, a-2, ... */
                    u32 n = avail - leftCols; if (n > 8) n = 8;
                    if (a >= 8 && b >= 8) {
                        u64 x1 = ld8(cls1, a - 8), x2 = ld8(cls2, b - 8);
                        u64 pr = (x1 << 4) | x2;
#pragma unroll
                        for (int i = 7; i >= 0; i--) {
                            if (going && (u32)(7 - i) < n) {
                                if (run >= leftScore - xDrop) {
                                    run += pair_score<SMALL>(lut, x1, x2, pr, i);
                                    leftCols++;
                                    if (run > leftScore) { leftScore = run; leftLen = leftCols; }
                                } else going = false;
                            }
                        }
                    } else {                                              /* within 8 bases of a sequence start */
                        for (u32 i = 0; i < n && going; i++) {
                            if (run >= leftScore - xDrop) {
                                u32 c1 = cls1[a - 1 - i], c2 = cls2[b - 1 - i];
                                run += SMALL ? lut[(c1 << 4) | c2] : lut[c1 * LZB_MAX_CLASSES + c2];
                                leftCols++;
                                if (run > leftScore) { leftScore = run; leftLen = leftCols; }
                            } else going = false;
                        }
                    }
                }
            }
            nExt++; nBp += rightCols + leftCols;
            s32 sim = leftScore + rightScore;
            if (sim < P.K) continue;                /* entropy can only lower the score */
            cand_rec r;
            r.hit1 = pos1; r.hit2 = pos2; r.pos1 = pos1 - leftLen; r.pos2 = pos2 - leftLen;
            r.length = leftLen + rightLen; r.score = sim; r.cA = r.cC = r.cG = r.cT = 0;
            if (P.entropy && sim <= 3 * P.K) {      /* match counts for entropy(), dna_utilities.c:2905-2915 */
                for (u32 i = 0; i < r.length; i++) {
                    u8 x = asc1[r.pos1 + i];
                    if (x == asc2[r.pos2 + i]) { r.cA += x == 'A'; r.cC += x == 'C'; r.cG += x == 'G'; r.cT += x == 'T'; }
                }
            }
            u32 slot = (u32)atomicAdd(&cnt->ncand, 1ull);
            if (slot < candCap) cand[slot] = r;
        }
        if (lane == 0) diagEnd[h] = E;
    }
    nExt = __reduce_add_sync(0xFFFFFFFFu, (unsigned)nExt);
    unsigned long long bpw = nBp;
    for (int o = 16; o > 0; o >>= 1) bpw += __shfl_down_sync(0xFFFFFFFFu, bpw, o);
    if (lane == 0) { if (nExt) atomicAdd(&cnt->extensions, nExt); if (bpw) atomicAdd(&cnt->bpExtended, bpw); }
}

/* ---- K3c: the default extension kernel (x-drop, <= 16 byte classes): xdrop_warp.cuh ----
 * Buckets are handed out largest first from a global counter (longest-processing-time order): the
 * few buckets that hold a homologous diagonal take far longer than the rest, and with a static
 * assignment they finished long after every other warp had run dry. */
__global__ void k_bucket_sizes(const u32* __restrict__ bstart, u32 nbuckets, u32* __restrict__ cnt, u32* __restrict__ ids) {
    for (u32 h = blockIdx.x * blockDim.x + threadIdx.x; h < nbuckets; h += gridDim.x * blockDim.x) { cnt[h] = bstart[h + 1] - bstart[h]; ids[h] = h; }
}

__global__ void __launch_bounds__(256, 4)
k_extend2(const u64* __restrict__ hits, const u32* __restrict__ bstart, const u32* __restrict__ order, u32 nbuckets,
          const u8* __restrict__ cls1, const u8* __restrict__ cls2, const u8* __restrict__ asc1, const u8* __restrict__ asc2,
          const lzb_scoring_dev* __restrict__ sc, sp_dev P, u32* __restrict__ diagEnd,
          cand_rec* __restrict__ cand, u32 candCap, search_counters* cnt, u32* __restrict__ nextBucket) {
    __shared__ s32 lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const int c1 = i / (int)XD_LUT_STRIDE, c2 = i % (int)XD_LUT_STRIDE;
        lut[i] = (c1 < XD_LUT_MAX_CLASSES && c2 < XD_LUT_MAX_CLASSES) ? sc->msubC[c1 * LZB_MAX_CLASSES + c2] : 0;
    }
    __syncthreads();
    xd_env e;
    e.cls1 = cls1; e.cls2 = cls2; e.asc1 = asc1; e.asc2 = asc2; e.lut = lut;
    e.len1 = P.len1; e.len2 = P.len2; e.L = (u32)P.L; e.xDrop = P.xDrop; e.K = P.K; e.entropy = P.entropy;
    e.cand = cand; e.candCap = candCap; e.ncand = &cnt->ncand;
    const u32 lane = threadIdx.x & 31;
    unsigned long long nExt = 0, nBp = 0;
    for (;;) {
        u32 slot = 0;
        if (lane == 0) slot = atomicAdd(nextBucket, 1u);
        slot = __shfl_sync(0xFFFFFFFFu, slot, 0);
        if (slot >= nbuckets) break;
        const u32 h = order[slot];
        const u32 b0 = bstart[h], b1 = bstart[h + 1];
        if (b0 == b1) break;                            /* sorted by size: everything after is empty too */
        u32 E = diagEnd[h];
        xd_bucket(e, lane, hits, b0, b1, E, nExt, nBp);
        if (lane == 0) diagEnd[h] = E;
    }
    for (int o = 16; o > 0; o >>= 1) { nExt += __shfl_down_sync(0xFFFFFFFFu, nExt, o); nBp += __shfl_down_sync(0xFFFFFFFFu, nBp, o); }
    if (lane == 0) { if (nExt) atomicAdd(&cnt->extensions, nExt); if (nBp) atomicAdd(&cnt->bpExtended, nBp); }
}

/* ---- K3d: --exact=N and --mismatch=M,N extension (match_extend_seed_hit seed_search.c:3018-3254,
 * mismatch_extend_seed_hit :3450-3778).  In the mismatch mode the extent a hit leaves in its bucket
 * depends on the left scan (:3740), which depends on the bucket -- right scans cannot run ahead of the
 * replay -- so these non-default modes take the plain route: ONE THREAD PER BUCKET walks its hits in
 * discovery order (2^hashBits independent threads).  Bases are compared with the case-insensitive
 * nuc_to_bits table (dna_utilities.c:56; params->charToBits lastz.c:353), not with the score matrix. */
__device__ __forceinline__ int alt_bits(u8 c) {
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return -1; }
}
__device__ __forceinline__ bool alt_mism(u8 a, u8 b) { const int x = alt_bits(a), y = alt_bits(b); return x != y || x < 0 || y < 0; }

__global__ void __launch_bounds__(128)
k_extend_alt(const u64* __restrict__ hits, const u32* __restrict__ bstart, u32 nbuckets,
             const u8* __restrict__ v1, const u8* __restrict__ v2, sp_dev P, int mismatches, u32* __restrict__ diagEnd,
             cand_rec* __restrict__ cand, u32 candCap, search_counters* cnt) {
    const u32 L = (u32)P.L, K = (u32)P.K, INACTIVE = 0xFFFFFFFFu;
    unsigned long long nExt = 0;
    for (u32 h = blockIdx.x * blockDim.x + threadIdx.x; h < nbuckets; h += gridDim.x * blockDim.x) {
        const u32 b0 = bstart[h], b1 = bstart[h + 1];
        if (b0 == b1) continue;
        u32 E = diagEnd[h];
        for (u32 idx = b0; idx < b1; idx++) {
            const u64 rec = hits[idx];
            const u32 pos1 = (u32)rec, pos2 = (u32)(rec >> 32);
            if (E > pos2 - L) continue;                              /* process_for_simple_hit :1113 */
            const s64 diag = (s64)pos1 - (s64)pos2;
            const s64 blk = (s64)E + diag, stop = blk > 0 ? blk : 0;
            const s64 lim = (s64)P.len2 + diag, rstop = ((s64)P.len1 <= lim) ? (s64)P.len1 : lim;
            u32 extent = INACTIVE; int inHit = 0; bool reject = false;
            for (u32 k = 1; k <= L; k++)                             /* mismatches inside the hit, right to left */
                if (alt_mism(v1[pos1 - k], v2[pos2 - k])) { extent = pos2 - k; if (++inHit > mismatches) { reject = true; break; } }
            if (reject) { if (extent > E) E = extent; continue; }   /* hit_isnt_a_match :3232, :3760 */
            s64 s1 = (s64)pos1 - L, s2 = (s64)pos2 - L, left, right;
            if (mismatches == 0) {                                   /* ---- exact ---- */
                if (s1 < stop) { s1--; s2--; }
                else while (s1 >= stop) {
                    if (s1 == stop) { s1--; s2--; break; }
                    const u8 n1 = v1[--s1], n2 = v2[--s2];
                    if (n1 == 0 || n2 == 0 || alt_mism(n1, n2)) break;
                }
                left = s1;
                s1 = (s64)pos1 - 1; s2 = (s64)pos2 - 1;
                while (s1 < rstop) {
                    const u8 n1 = v1[++s1], n2 = v2[++s2];
                    if (n1 == 0 || n2 == 0 || alt_mism(n1, n2)) break;
                }
                right = s1;
                extent = (u32)(right - diag);
            } else {                                                 /* ---- up to `mismatches` mismatches ---- */
                s64 mmLoc[LZB_GFEX_MISMATCH_MAX + 1];
                int mmScan = mismatches + 1 - inHit; const int mmStop = mmScan;
                if (s1 < stop) { s1--; s2--; }
                else while (s1 >= stop) {
                    if (s1 == stop) { s1--; s2--; break; }
                    const u8 n1 = v1[--s1], n2 = v2[--s2];
                    if (n1 == 0 || n2 == 0) break;
                    if (alt_mism(n1, n2)) { mmLoc[--mmScan] = s1; if (mmScan == 0) break; }
                }
                if (mmScan > 0) mmLoc[--mmScan] = s1;
                int shortfall = mmScan;
                s1 = (s64)pos1 - 1; s2 = (s64)pos2 - 1;
                s64 bestLength = 0; left = right = -2; bool have = false;
                while (s1 < rstop) {
                    const u8 n1 = v1[++s1], n2 = v2[++s2];
                    if (n1 == 0 || n2 == 0) break;
                    if (alt_mism(n1, n2)) {
                        if (extent == INACTIVE) extent = (u32)s2;
                        if (shortfall > 0) { shortfall--; continue; }
                        const s64 thisLength = s1 - mmLoc[mmScan];
                        if (thisLength > bestLength) { bestLength = thisLength; left = mmLoc[mmScan]; right = s1; have = true; }
                        if (++mmScan == mmStop) break;
                    }
                }
                if (mmScan < mmStop) {
                    if (extent == INACTIVE) extent = (u32)s2;
                    const s64 thisLength = s1 - mmLoc[mmScan];
                    if (thisLength > bestLength) { left = mmLoc[mmScan]; right = s1; have = true; }
                }
                if (!have) continue;
                if ((u32)(right - (left + 1)) >= K) extent = (u32)(right + 1 - diag);
            }
            nExt++;
            if (extent > E) E = extent;
            const u32 length = (u32)(right - (left + 1));
            if (length < K) continue;
            const u32 slot = (u32)atomicAdd(&cnt->ncand, 1ull);
            if (slot < candCap) { cand_rec r = { pos1, pos2, (u32)(left + 1), (u32)(left + 1 - diag), length, (s32)length, 0, 0, 0, 0 }; cand[slot] = r; }
        }
        diagEnd[h] = E;
    }
    if (nExt) atomicAdd(&cnt->extensions, nExt);
}

/* ------------------------------------------------------------------------------------------ */

#include "xdrop_split.cuh"

static inline u32 host_pack(const lzb_seed* sd, u64 w) {
    u32 p = 0;
    for (int i = 0; i < sd->numParts; i++) p |= (u32)(w >> sd->shift[i]) & sd->mask[i];
    return p;
}
static u32 host_word_at(const u8* v, u32 endPos, const lzb_seed* sd, const int8_t* ctb) {
    u64 w = 0;
    for (int j = 0; j < sd->length; j++) w = (w << 2) | (u64)(ctb[v[endPos - sd->length + j]] & 3);
    return host_pack(sd, w);
}

struct ev_pair { cudaEvent_t a, b; int which; };

/* Device scratch of the seed stage, kept in the context across calls: at 50 Mbp x 50 Mbp the hit buffers are
 * 3.5 GB and cudaMalloc + cudaFree of them cost 0.1-0.7 s of host time per call (LZB_SEED_TRACE) against
 * 0.34 s of device work.  A buffer only ever grows. */
#define SEED_SCRATCH_SLOTS 24
struct seed_scratch { void* p[SEED_SCRATCH_SLOTS]; size_t cap[SEED_SCRATCH_SLOTS]; };
static int scratch_get(lzb_ctx* c, int slot, size_t bytes, void** out) {
    if (!c->seedScratch) c->seedScratch = calloc(1, sizeof(seed_scratch));
    seed_scratch* sc = (seed_scratch*)c->seedScratch;
    if (bytes == 0) bytes = 16;
    if (sc->cap[slot] < bytes) {
        if (sc->p[slot]) { cudaStreamSynchronize(c->stream); cudaFree(sc->p[slot]); sc->p[slot] = NULL; sc->cap[slot] = 0; }
        CUDA_TRY(cudaMalloc(&sc->p[slot], bytes));
        sc->cap[slot] = bytes;
    }
    *out = sc->p[slot];
    return 0;
}
#define SCRATCH(slot_, var_, size_) do { if (scratch_get(c, (slot_), (size_), (void**)&var_)) return -1; } while (0)
void lzb_seed_scratch_free(lzb_ctx* c) {
    seed_scratch* sc = (seed_scratch*)c->seedScratch;
    if (!sc) return;
    for (int i = 0; i < SEED_SCRATCH_SLOTS; i++) cudaFree(sc->p[i]);
    free(sc); c->seedScratch = NULL;
}

extern "C" int lzb_seed_hit_search(lzb_ctx* c, lzb_target* t, lzb_query* q, const lzb_seed* seed,
                                   const int8_t ctb[256], const lzb_seed_params* prm,
                                   lzb_segment** segs, uint64_t* nsegs, lzb_seed_stats* stats) {
    cudaSetDevice(c->device);
    if (!c->haveScoring) return lzb_fail("lzb_set_scoring has not been called");
    u32 qstart = prm->start, qend = prm->end ? prm->end : q->len;
    if (qend <= qstart) return lzb_fail("in seed_hit_search(), interval is void (%u-%u)", qstart, qend);
    if (qend > q->len) return lzb_fail("in seed_hit_search(), interval end is bad (%u>%u)", qend, q->len);
    if (seed->length < 2) return lzb_fail("seed length must be at least two (yours is %d)", seed->length);
    if (prm->gfExtend == LZB_GFEX_MISMATCH && (prm->gfMismatches < 1 || prm->gfMismatches > LZB_GFEX_MISMATCH_MAX))
        return lzb_fail("%d is out of range for N-mismatch (valid range is 1..%d)", prm->gfMismatches, LZB_GFEX_MISMATCH_MAX);
    if (seed->weight != t->wordBits || seed->length != t->seedLength)
        return lzb_fail("the seed does not match the one the target index was built with");
    int hashBits = prm->hashBits ? prm->hashBits : 16;
    if (hashBits < 4 || hashBits > 26) return lzb_fail("diag hash size 2^%d is not supported", hashBits);
    *segs = NULL; *nsegs = 0;
    if (stats) memset(stats, 0, sizeof *stats);
    cudaStream_t st = c->stream;
    /* LZB_SEED_TRACE=1: host wall clock of the call's phases on stderr (what the CUDA-event time does not see) */
    const bool wtrace = getenv("LZB_SEED_TRACE") != NULL;
    const auto w0 = std::chrono::steady_clock::now(); double wt[8] = {0}; int wk = 0;
#define WMARK() do { if (wtrace && wk < 8) wt[wk++] = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count(); } while (0)

    /* seed variants in probing order (seed_search.c:522-549) */
    std::vector<u32> flips; flips.push_back(0);
    if (seed->withTrans == 1) for (int f = 0; f < seed->numFlips; f++) flips.push_back(seed->transFlips[f]);
    else if (seed->withTrans >= 2)
        for (int f = 0; f < seed->numFlips; f++) {
            flips.push_back(seed->transFlips[f]);
            for (int g = f + 1; g < seed->numFlips; g++) flips.push_back(seed->transFlips[f] ^ seed->transFlips[g]);
        }
    if (flips.size() > MAX_VARIANTS) return lzb_fail("too many seed variants (%zu)", flips.size());
    sp_dev P; memset(&P, 0, sizeof P);
    P.qstart = qstart; P.qend = qend; P.len1 = t->len; P.len2 = q->len; P.L = seed->length; P.V = (int)flips.size();
    P.hashBits = hashBits; P.selfCompare = prm->selfCompare; P.sameStrand = prm->sameStrand;
    P.xDrop = prm->xDrop; P.K = prm->hspThreshold; P.gfExtend = prm->gfExtend; P.plain = prm->plainHits; P.entropy = prm->entropy;
    seed_dev sd; seed_to_dev(&sd, seed);
    ctb_dev2 cd; memcpy(cd.v, ctb, 256);

    u32 n = qend - qstart;
    u32 nblk = (n + POS_PER_BLOCK - 1) / POS_PER_BLOCK;
    u32 nbuckets = 1u << hashBits;
    u32 *d_qword = NULL, *d_flips = NULL, *d_E = NULL, *d_bstart = NULL;
    unsigned long long* d_blkcnt = NULL; search_counters* d_cnt = NULL;
    std::vector<ev_pair> evs;
    cudaEvent_t evBegin, evMid, evMid2, evEnd;       /* device time = [begin,mid] + [mid2,end]; planning + cudaMalloc sit between */
    CUDA_TRY(cudaEventCreate(&evBegin)); CUDA_TRY(cudaEventCreate(&evEnd));
    CUDA_TRY(cudaEventCreate(&evMid)); CUDA_TRY(cudaEventCreate(&evMid2));
    SCRATCH(0, d_qword, (size_t)n * 4 + 16);
    SCRATCH(1, d_flips, flips.size() * 4);
    SCRATCH(2, d_blkcnt, (size_t)nblk * 8);
    SCRATCH(3, d_cnt, sizeof(search_counters));
    SCRATCH(4, d_E, (size_t)nbuckets * 4);
    SCRATCH(5, d_bstart, ((size_t)nbuckets + 2) * 4);
    CUDA_TRY(cudaMemcpyAsync(d_flips, flips.data(), flips.size() * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(d_cnt, 0, sizeof(search_counters), st));
    CUDA_TRY(cudaMemsetAsync(d_E, 0, (size_t)nbuckets * 4, st));        /* empty_diag_hash diag_hash.c:125 */
    CUDA_TRY(cudaEventRecord(evBegin, st));
    int grid = c->smCount * 8;
#define TIMED(which_, launch_)                                                          \
    do { ev_pair e_; e_.which = (which_); cudaEventCreate(&e_.a); cudaEventCreate(&e_.b); \
         cudaEventRecord(e_.a, st); launch_; cudaEventRecord(e_.b, st); evs.push_back(e_); \
         c->launches++; } while (0)
    TIMED(0, (k_query_words<<<grid, 256, 0, st>>>(q->d_seq, P, sd, cd, d_qword, d_cnt)));
    TIMED(1, (k_count_hits<<<nblk, 256, 0, st>>>(d_qword, t->d_off, t->d_pos, d_flips, P, d_blkcnt)));
    CUDA_TRY(cudaGetLastError());
    std::vector<unsigned long long> blkcnt(nblk);
    CUDA_TRY(cudaMemcpyAsync(blkcnt.data(), d_blkcnt, (size_t)nblk * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(evMid, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    WMARK();                                            /* [0] small allocations + words + hit counts + sync */
    u64 totalHits = 0, maxBlk = 0;
    for (u32 b = 0; b < nblk; b++) { totalHits += blkcnt[b]; if (blkcnt[b] > maxBlk) maxBlk = blkcnt[b]; }

    const char* capEnv = getenv("LZB_HIT_CAP");
    u64 hitCap = capEnv ? strtoull(capEnv, 0, 10) : (1ull << 27);
    if (hitCap > 0xFFFFFFF0ull) hitCap = 0xFFFFFFF0ull;
    if (maxBlk > hitCap) hitCap = maxBlk;
    if (hitCap > 0xFFFFFFF0ull) return lzb_fail("%llu seed hits within %d query positions; use --step or mask repeats", (unsigned long long)maxBlk, POS_PER_BLOCK);
    if (totalHits < hitCap) hitCap = totalHits ? totalHits : 1;
    u64 slotCap = 1ull << 26;
    { u64 allSlots = (u64)nblk * POS_PER_BLOCK * P.V; if (allSlots < slotCap) slotCap = allSlots; }
    if (slotCap < (u64)POS_PER_BLOCK * P.V) slotCap = (u64)POS_PER_BLOCK * P.V;
    u32 candCap = prm->plainHits || prm->gfExtend != LZB_GFEX_XDROP ? (u32)std::min<u64>(totalHits + 1, 1ull << 26) : (1u << 22);

    u32 *d_slotcnt = NULL, *d_slotoff = NULL, *keysA = NULL, *keysB = NULL; u64 *valsA = NULL, *valsB = NULL;
    cand_rec* d_cand = NULL; void* d_tmp = NULL; size_t tmpBytes = 0, tmpScan = 0;
    SCRATCH(6, d_slotcnt, (slotCap + 2) * 4); SCRATCH(7, d_slotoff, (slotCap + 2) * 4);
    SCRATCH(8, keysA, hitCap * 4); SCRATCH(9, keysB, hitCap * 4);
    SCRATCH(10, valsA, hitCap * 8); SCRATCH(11, valsB, hitCap * 8);
    SCRATCH(12, d_cand, (size_t)candCap * sizeof(cand_rec));
    /* the three-kernel extension (xdrop_split.cuh) measured SLOWER than the fused kernel (0.70 s vs 0.62 s
     * at 50 Mbp x 50 Mbp: the extra passes over the hit records cost more than the lockstep idling they
     * remove), so it is opt-in: LZB_SPLIT_EXTEND=1.  Kept because it bounds the work on long repeats. */
    const bool splitExtend = prm->gfExtend == LZB_GFEX_XDROP && !prm->plainHits && c->sc.numClasses <= 16 &&
                             getenv("LZB_SPLIT_EXTEND") && atoi(getenv("LZB_SPLIT_EXTEND"));
    /* the warp-cooperative kernel (xdrop_warp.cuh) is the default x-drop path; LZB_EXTEND_V1=1 keeps the first one */
    const bool coopExtend = prm->gfExtend == LZB_GFEX_XDROP && !prm->plainHits && c->sc.numClasses <= XD_LUT_MAX_CLASSES && !splitExtend &&
                            !(getenv("LZB_EXTEND_V1") && atoi(getenv("LZB_EXTEND_V1")));
    u32 *d_bcnt = NULL, *d_bcnt2 = NULL, *d_bid = NULL, *d_border = NULL, *d_next = NULL; size_t tmpOrder = 0;
    if (coopExtend) {
        SCRATCH(13, d_bcnt, (size_t)nbuckets * 4); SCRATCH(14, d_bcnt2, (size_t)nbuckets * 4);
        SCRATCH(15, d_bid, (size_t)nbuckets * 4); SCRATCH(16, d_border, (size_t)nbuckets * 4);
        SCRATCH(17, d_next, 4);
        CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(NULL, tmpOrder, d_bcnt, d_bcnt2, d_bid, d_border, nbuckets, 0, 32, st));
    }
    right_rec* d_right = NULL; live_rec* d_live = NULL; unsigned long long* d_nlive = NULL;
    if (splitExtend) {
        SCRATCH(18, d_right, hitCap * sizeof(right_rec));
        SCRATCH(19, d_live, hitCap * sizeof(live_rec));
        SCRATCH(20, d_nlive, 8);
    }
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(NULL, tmpBytes, keysA, keysB, valsA, valsB, hitCap, 0, hashBits, st));
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(NULL, tmpScan, d_slotcnt, d_slotoff, slotCap + 1, st));
    if (tmpScan > tmpBytes) tmpBytes = tmpScan;
    if (tmpOrder > tmpBytes) tmpBytes = tmpOrder;
    SCRATCH(21, d_tmp, tmpBytes);

    /* chunk loop */
    WMARK();                                            /* [1] hit/slot/candidate buffers allocated */
    CUDA_TRY(cudaEventRecord(evMid2, st));
    u64 chunks = 0;
    for (u32 b0 = 0; b0 < nblk;) {
        u64 h = 0; u32 b1 = b0;
        while (b1 < nblk && h + blkcnt[b1] <= hitCap && (u64)(b1 + 1 - b0) * POS_PER_BLOCK * P.V <= slotCap) { h += blkcnt[b1]; b1++; }
        if (b1 == b0) return lzb_fail("internal error: seed-hit chunk planner made no progress");
        u32 k0 = b0 * POS_PER_BLOCK, k1 = std::min<u64>((u64)b1 * POS_PER_BLOCK, n);
        if (h > 0) {
            u32 nslots = (k1 - k0) * (u32)P.V, nh = (u32)h;
            size_t tb = tmpBytes;
            TIMED(2, (k_slot_count<<<grid, 256, 0, st>>>(d_qword, t->d_off, t->d_pos, d_flips, P, k0, nslots, d_slotcnt)));
            TIMED(3, cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_slotcnt, d_slotoff, nslots + 1, st));
            TIMED(4, (k_expand<<<grid, 256, 0, st>>>(d_qword, t->d_off, t->d_pos, d_flips, P, k0, nslots, d_slotoff, keysA, valsA)));
            tb = tmpBytes;
            TIMED(5, cub::DeviceRadixSort::SortPairs(d_tmp, tb, keysA, keysB, valsA, valsB, nh, 0, hashBits, st));
            TIMED(6, (k_bucket_bounds<<<(nbuckets + 256) / 256, 256, 0, st>>>(keysB, nh, nbuckets, d_bstart)));
            if (splitExtend) {
                CUDA_TRY(cudaMemsetAsync(d_nlive, 0, 8, st));
                TIMED(8, (k_right<<<c->smCount * 8, 256, 0, st>>>(valsB, nh, t->d_cls, q->d_cls, c->d_sc, P, d_E, d_right)));
                TIMED(9, (k_replay<<<grid, 256, 0, st>>>(valsB, d_bstart, nbuckets, d_right, t->d_cls, q->d_cls, c->d_sc, P, d_E, d_live, d_nlive)));
                TIMED(10, (k_left<<<c->smCount * 8, 256, 0, st>>>(d_live, d_nlive, t->d_cls, q->d_cls, t->d_seq, q->d_seq, c->d_sc, P, d_cand, candCap, d_cnt)));
            } else if (prm->gfExtend == LZB_GFEX_EXACT || prm->gfExtend == LZB_GFEX_MISMATCH) {
                TIMED(7, (k_extend_alt<<<(nbuckets + 127) / 128, 128, 0, st>>>(valsB, d_bstart, nbuckets, t->d_seq, q->d_seq, P,
                                                                                 prm->gfExtend == LZB_GFEX_EXACT ? 0 : prm->gfMismatches,
                                                                                 d_E, d_cand, candCap, d_cnt)));
            } else if (coopExtend) {
                int bits = 1; while (bits < 32 && (nh >> bits)) bits++;          /* bucket sizes are <= nh */
                size_t tbo = tmpBytes;
                CUDA_TRY(cudaMemsetAsync(d_next, 0, 4, st));
                TIMED(11, (k_bucket_sizes<<<(nbuckets + 255) / 256, 256, 0, st>>>(d_bstart, nbuckets, d_bcnt, d_bid)));
                TIMED(11, cub::DeviceRadixSort::SortPairsDescending(d_tmp, tbo, d_bcnt, d_bcnt2, d_bid, d_border, nbuckets, 0, bits, st));
                TIMED(7, (k_extend2<<<c->smCount * 4, 256, 0, st>>>(valsB, d_bstart, d_border, nbuckets, t->d_cls, q->d_cls, t->d_seq, q->d_seq,
                                                                     c->d_sc, P, d_E, d_cand, candCap, d_cnt, d_next)));
            } else if (c->sc.numClasses <= 16)
                TIMED(7, (k_extend<true><<<grid, 256, 0, st>>>(valsB, d_bstart, nbuckets, t->d_cls, q->d_cls, t->d_seq, q->d_seq,
                                                               c->d_sc, P, d_E, d_cand, candCap, d_cnt)));
            else
                TIMED(7, (k_extend<false><<<grid, 256, 0, st>>>(valsB, d_bstart, nbuckets, t->d_cls, q->d_cls, t->d_seq, q->d_seq,
                                                                c->d_sc, P, d_E, d_cand, candCap, d_cnt)));
            CUDA_TRY(cudaGetLastError());
            chunks++;
        }
        b0 = b1;
    }
    CUDA_TRY(cudaEventRecord(evEnd, st));
    WMARK();                                            /* [2] chunk loop enqueued */
    search_counters hc;
    CUDA_TRY(cudaMemcpyAsync(&hc, d_cnt, sizeof hc, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    WMARK();                                            /* [3] device work finished */
    if (hc.ncand > candCap) {
        lzb_fail("%llu HSP candidates exceed the %u-entry result buffer; raise the threshold (--hspthresh)", hc.ncand, candCap);
        goto cleanup_fail;
    }
    {
        std::vector<cand_rec> cand(hc.ncand);
        if (hc.ncand) CUDA_TRY(cudaMemcpy(cand.data(), d_cand, (size_t)hc.ncand * sizeof(cand_rec), cudaMemcpyDeviceToHost));
        /* discovery order = (query position, variant index, target position descending).  Candidates
         * that share a query position are rare, so: sort (hit2, index) as plain 64-bit keys, then order
         * each run of equal hit2 by the seed variant that produced the hit (recomputed from the bytes) */
        struct keyed { u32 variant; u32 hit1; u32 ix; };
        std::vector<u64> order(cand.size());
        for (size_t i = 0; i < cand.size(); i++) order[i] = ((u64)cand[i].hit2 << 32) | (u32)i;
        std::sort(order.begin(), order.end());
        for (size_t i = 0; i < order.size();) {
            size_t j = i + 1;
            while (j < order.size() && (order[j] >> 32) == (order[i] >> 32)) j++;
            if (j - i > 1) {
                std::vector<keyed> run(j - i);
                for (size_t k = i; k < j; k++) {
                    const cand_rec& r = cand[(u32)order[k]];
                    u32 x = host_word_at(q->h_seq, r.hit2, seed, ctb) ^ host_word_at(t->h_seq, r.hit1, seed, ctb), v = 0;
                    for (; v < flips.size(); v++) if (flips[v] == x) break;
                    run[k - i] = { v, r.hit1, (u32)order[k] };
                }
                std::sort(run.begin(), run.end(), [](const keyed& a, const keyed& b) {
                    if (a.variant != b.variant) return a.variant < b.variant;
                    return a.hit1 > b.hit1;
                });
                for (size_t k = i; k < j; k++) order[k] = (order[k] & 0xFFFFFFFF00000000ull) | run[k - i].ix;
            }
            i = j;
        }
        lzb_segment* out = (lzb_segment*)malloc((cand.size() + 1) * sizeof(lzb_segment));
        u64 m = 0;
        for (size_t i = 0; i < order.size(); i++) {
            const cand_rec& r = cand[(u32)order[i]];
            s32 sim = r.score;
            if (!prm->plainHits && prm->gfExtend == LZB_GFEX_XDROP && prm->entropy &&
                sim >= prm->hspThreshold && sim <= 3 * prm->hspThreshold) {
                /* entropy() dna_utilities.c:2918-2936 on the device's integer counts */
                double qf = 1.0;
                if (r.cA + r.cC + r.cG + r.cT >= 20) {
                    double len = (double)(int)r.length;
                    double pA = (double)(int)r.cA / len, pC = (double)(int)r.cC / len, pG = (double)(int)r.cG / len, pT = (double)(int)r.cT / len;
                    double qA = r.cA ? log(pA) : 0.0, qC = r.cC ? log(pC) : 0.0, qG = r.cG ? log(pG) : 0.0, qT = r.cT ? log(pT) : 0.0;
                    qf = -(pA * qA + pC * qC + pG * qG + pT * qT) / log(4.0);
                }
                sim = (s32)((double)sim * qf);
                if (sim < prm->hspThreshold) continue;
            }
            lzb_segment* g = &out[m++];
            memset(g, 0, sizeof *g);
            g->pos1 = r.pos1; g->pos2 = r.pos2; g->length = r.length; g->s = sim; g->id = prm->strandId; g->scoreCov = r.length;
        }
        *segs = out; *nsegs = m;
        WMARK();                                        /* [4] candidates copied back, entropy, ordering */
        if (stats) {
            stats->wordsInQuery = hc.words; stats->rawSeedHits = totalHits; stats->extensions = hc.extensions;
            stats->bpExtended = hc.bpExtended; stats->hsps = m;
            float ms = 0, ms2 = 0; cudaEventElapsedTime(&ms, evBegin, evMid); cudaEventElapsedTime(&ms2, evMid2, evEnd);
            stats->seconds = (ms + ms2) / 1e3;
            for (auto& e : evs) { float x = 0; cudaEventElapsedTime(&x, e.a, e.b); stats->kernelSeconds[e.which] += x / 1e3; stats->kernelLaunches[e.which]++; }
        }
    }
    for (auto& e : evs) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    cudaEventDestroy(evBegin); cudaEventDestroy(evEnd); cudaEventDestroy(evMid); cudaEventDestroy(evMid2);
    WMARK();                                            /* [5] buffers freed */
    if (wtrace) fprintf(stderr, "[seed trace] plan=%.4f alloc=%.4f enqueue=%.4f device_wait=%.4f post=%.4f free=%.4f total=%.4f s (chunks=%llu hits=%llu cand=%llu)\n",
                        wt[0], wt[1] - wt[0], wt[2] - wt[1], wt[3] - wt[2], wt[4] - wt[3], wt[5] - wt[4], wt[5],
                        (unsigned long long)chunks, (unsigned long long)totalHits, (unsigned long long)hc.ncand);
    return 0;
cleanup_fail:
    for (auto& e : evs) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    return -1;
}
