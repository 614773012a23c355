/*
 * seed_kernels.cuh -- device code of the seed stage (K2/K3): query words, hit counting, expansion of the
 * position lists into hit records, bucket bounds/sizes, and the extension kernels (the first-generation
 * k_extend, k_extend2 = xdrop_warp.cuh, k_extend_alt for --exact/--mismatch).  Host orchestration lives in
 * seed_search.cu.  Kept free of runtime calls so that the kernels also compile for the host block emulator
 * (tests/warp_emu/cuda_emu.h, tests/test_seed_kernels_emu.py).
 */
#ifndef LZB_SEED_KERNELS_CUH
#define LZB_SEED_KERNELS_CUH

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

/* ---- K3r: process_for_recoverable_hit (seed_search.c:1221-1443, --recoverseeds) with x-drop or no extension ----
 * One thread replays one bucket in discovery order.  The bucket remembers the ACTUAL diagonal of its last fresh hit
 * beside the extent: a hit on another diagonal of the same bucket is extended instead of lost (:1298-1336), one on the
 * same diagonal that starts inside the extent is dropped and moves the extent up to its own end (:1341-1360).  Left
 * extension is not blocked by the bucket (unblockedLeftExtension, :2612) and the extent only ever grows (:2785-2789),
 * so overlapping HSPs can come out; the caller merges them (merge_segments segment.c:1527).  A non-default option:
 * the scans walk one column at a time. */
__global__ void __launch_bounds__(128)
k_extend_recover(const u64* __restrict__ hits, const u32* __restrict__ bstart, u32 nbuckets,
                 const u8* __restrict__ cls1, const u8* __restrict__ cls2,
                 const u8* __restrict__ asc1, const u8* __restrict__ asc2,
                 const lzb_scoring_dev* __restrict__ sc, sp_dev P, u32* __restrict__ diagEnd, s32* __restrict__ diagActual,
                 cand_rec* __restrict__ cand, u32 candCap, search_counters* cnt) {
    const u32 L = (u32)P.L;
    const s32 xDrop = P.xDrop;
    unsigned long long nExt = 0, nBp = 0;
    for (u32 h = blockIdx.x * blockDim.x + threadIdx.x; h < nbuckets; h += gridDim.x * blockDim.x) {
        const u32 b0 = bstart[h], b1 = bstart[h + 1];
        if (b0 == b1) continue;
        u32 E = diagEnd[h]; s32 A = diagActual[h];       /* an untouched bucket reads as extent 0, which lets every hit through */
        for (u32 idx = b0; idx < b1; idx++) {
            const u64 rec = hits[idx];
            const u32 pos1 = (u32)rec, pos2 = (u32)(rec >> 32);
            const s32 diag = (s32)(pos1 - pos2);
            if (diag == A && pos2 - L < E) { if (pos2 > E) E = pos2; continue; }
            A = diag;                                                    /* fresh_hit :1373 */
            if (P.gfExtend != LZB_GFEX_XDROP) {                         /* :1417-1421 */
                if (pos2 > E) E = pos2;
                const u32 slot = (u32)atomicAdd(&cnt->ncand, 1ull);
                if (slot < candCap) { cand_rec r = { pos1, pos2, pos1 - L, pos2 - L, L, 0, 0, 0, 0, 0 }; cand[slot] = r; }
                continue;
            }
            /* left scan :2598-2632 from the hit's right end, stopping only at the start of either sequence */
            const u32 stop = diag > 0 ? (u32)diag : 0u;
            u32 a = pos1, b = pos2, leftLen = 0, leftCols = 0; s32 run = 0, leftScore = 0;
            while (a > stop && run >= leftScore - xDrop) {
                --a; --b;
                run += sc->msubC[(u32)cls1[a] * LZB_MAX_CLASSES + cls2[b]];
                leftCols++;
                if (run > leftScore) { leftScore = run; leftLen = leftCols; }
            }
            /* right scan :2663-2693 to the end of either sequence */
            const s64 lim = (s64)P.len2 + diag;
            const u32 rstop = ((s64)P.len1 <= lim) ? P.len1 : (u32)lim;
            u32 rightLen = 0, rightCols = 0; s32 rightScore = 0; run = 0; a = pos1; b = pos2;
            while (a < rstop && run >= rightScore - xDrop) {
                run += sc->msubC[(u32)cls1[a] * LZB_MAX_CLASSES + cls2[b]];
                a++; b++; rightCols++;
                if (run > rightScore) { rightScore = run; rightLen = rightCols; }
            }
            nExt++; nBp += rightCols + leftCols;
            const u32 extent = (u32)((s64)a - diag);                    /* where the right scan stopped, in sequence 2 */
            if (extent > E) E = extent;
            const s32 sim = leftScore + rightScore;
            if (sim < P.K) continue;                                     /* entropy can only lower the score */
            cand_rec r;
            r.hit1 = pos1; r.hit2 = pos2; r.pos1 = pos1 - leftLen; r.pos2 = pos2 - leftLen;
            r.length = leftLen + rightLen; r.score = sim; r.cA = r.cC = r.cG = r.cT = 0;
            if (P.entropy && sim <= 3 * P.K) {
                for (u32 i = 0; i < r.length; i++) {
                    const u8 x = asc1[r.pos1 + i];
                    if (x == asc2[r.pos2 + i]) { r.cA += x == 'A'; r.cC += x == 'C'; r.cG += x == 'G'; r.cT += x == 'T'; }
                }
            }
            const u32 slot = (u32)atomicAdd(&cnt->ncand, 1ull);
            if (slot < candCap) cand[slot] = r;
        }
        diagEnd[h] = E; diagActual[h] = A;
    }
    if (nExt) atomicAdd(&cnt->extensions, nExt);
    if (nBp) atomicAdd(&cnt->bpExtended, nBp);
}

/* ---- K3t: process_for_twin_hit (seed_search.c:1814-2046, the seed-hit-queue version; --twins=<min>..<max>) ----
 * A hit is extended only when an earlier hit of its diagonal lies minSpan..maxSpan columns back.  The reference keeps
 * the recent hits -- and a "block" entry at the extent of every extension -- in a queue chained per hash bucket
 * (diag_hash.c:276-322) and walks a bucket's chain newest first until an entry is more than maxSpan columns back
 * (:1904-1947).  Here one thread replays one bucket in discovery order and the chain is the bucket's own entry list:
 * entries made in this chunk sit in ent[b0 ...] (at most one per hit), entries that can still matter to later chunks
 * are carried per bucket in carry[h][carryCap] (carryCap = 2 * maxSpan + 16: along a homologous diagonal every column
 * can leave an entry).  An entry stops mattering once every later hit sees it more than
 * maxSpan columns back: it ends the walk, exactly like the end of the chain does (both fall through to "queue this
 * hit", :1955-1958), so it and everything older are dropped.  The queue's CAPACITY is global state: the reference
 * forgets an entry after seedHitQueueSize newer ones in ALL buckets (and warns "seed hit queue shortfall" when that loses
 * a hit within maxSpan columns).  A forgotten entry ends the walk like the end of the chain, so it only matters when the
 * walk would have gone THROUGH it: a hit entry within maxSpan columns (the host refuses inputs whose hit density makes
 * that possible) or a block entry that is consulted long after it was made.  For the latter every block remembers the
 * query position it was born at, and consulting one that MAY have been forgotten -- at least queueSize raw hits between
 * its birth and now, by the per-block prefix sums of the hit counts -- raises cnt->overflow: the call then fails instead
 * of returning a table that might differ from the reference's. */
#define TWIN_CARRY_CAP(maxSpan_) (2u * (u32)(maxSpan_) + 16u)
struct twin_ent { u32 pos2; s32 diag; u32 isBlock; u32 born; };

__global__ void __launch_bounds__(128)
k_extend_twin(const u64* __restrict__ hits, const u32* __restrict__ bstart, u32 nbuckets,
              const u8* __restrict__ cls1, const u8* __restrict__ cls2,
              const u8* __restrict__ asc1, const u8* __restrict__ asc2,
              const lzb_scoring_dev* __restrict__ sc, sp_dev P, u32 minSpan, u32 maxSpan,
              const unsigned long long* __restrict__ blkPrefix, u32 nblk, u32 queueSize,
              u32* __restrict__ diagEnd, twin_ent* __restrict__ ent, twin_ent* __restrict__ carry, u32 carryCap, u32* __restrict__ ncarry,
              cand_rec* __restrict__ cand, u32 candCap, search_counters* cnt) {
    const u32 L = (u32)P.L;
    const s32 xDrop = P.xDrop;
    unsigned long long nExt = 0, nBp = 0;
    for (u32 h = blockIdx.x * blockDim.x + threadIdx.x; h < nbuckets; h += gridDim.x * blockDim.x) {
        const u32 b0 = bstart[h], b1 = bstart[h + 1];
        if (b0 == b1) continue;
        u32 E = diagEnd[h];
        twin_ent* const old = carry + (size_t)h * carryCap; const u32 nold = ncarry[h];
        twin_ent* const mine = ent + b0; u32 nnew = 0;
        u32 lastPos2 = 0;
        for (u32 idx = b0; idx < b1; idx++) {
            const u64 rec = hits[idx];
            const u32 pos1 = (u32)rec, pos2 = (u32)(rec >> 32);
            const s32 diag = (s32)(pos1 - pos2);
            lastPos2 = pos2;
            /* walk the chain, newest entry first (:1904-1947) */
            int verdict = 0;                                   /* 0 queue the hit, 1 blocked, 2 twin */
            u32 span = 0;
            for (u32 k = nnew + nold; k-- > 0;) {
                const twin_ent q = k >= nold ? mine[k - nold] : old[k];
                span = pos2 - (q.pos2 - L);
                if (span > maxSpan) break;
                if (q.isBlock) {                                /* may the reference have forgotten this block by now? */
                    const u32 bl = (q.born - P.qstart) / POS_PER_BLOCK, bh = (pos2 - P.qstart) / POS_PER_BLOCK;
                    const unsigned long long lo = blkPrefix[bl > 0 ? bl - 1 : 0], hi = blkPrefix[bh + 2 < nblk ? bh + 2 : nblk];
                    if (hi - lo + 2 >= queueSize) atomicAdd(&cnt->overflow, 1ull);
                }
                if (q.diag != diag) continue;
                if (q.isBlock) { if (pos2 - L <= q.pos2) verdict = 1; break; }
                if (span < minSpan) continue;
                verdict = 2; break;
            }
            if (verdict == 1) continue;
            if (verdict == 0) { twin_ent e = { pos2, diag, 0u, pos2 }; mine[nnew++] = e; continue; }
            if (P.gfExtend != LZB_GFEX_XDROP) {                 /* :2021-2026: the twin itself, span columns long */
                E = pos2;
                twin_ent e = { pos2, diag, 1u, pos2 }; mine[nnew++] = e;
                const u32 slot = (u32)atomicAdd(&cnt->ncand, 1ull);
                if (slot < candCap) { cand_rec r = { pos1, pos2, pos1 - span, pos2 - span, span, 0, 0, 0, 0, 0 }; cand[slot] = r; }
                continue;
            }
            /* xdrop_extend_seed_hit :2528-2959 from the hit's right end; the left scan stops at the bucket's extent */
            const s64 blk = (s64)E + diag;
            const u32 stop = blk > 0 ? (u32)blk : 0u;
            u32 a = pos1, b = pos2, leftLen = 0, leftCols = 0; s32 run = 0, leftScore = 0;
            while (a > stop && run >= leftScore - xDrop) {
                --a; --b;
                run += sc->msubC[(u32)cls1[a] * LZB_MAX_CLASSES + cls2[b]];
                leftCols++;
                if (run > leftScore) { leftScore = run; leftLen = leftCols; }
            }
            const s64 lim = (s64)P.len2 + diag;
            const u32 rstop = ((s64)P.len1 <= lim) ? P.len1 : (u32)lim;
            u32 rightLen = 0, rightCols = 0; s32 rightScore = 0; run = 0; a = pos1; b = pos2;
            while (a < rstop && run >= rightScore - xDrop) {
                run += sc->msubC[(u32)cls1[a] * LZB_MAX_CLASSES + cls2[b]];
                a++; b++; rightCols++;
                if (run > rightScore) { rightScore = run; rightLen = rightCols; }
            }
            nExt++; nBp += rightCols + leftCols;
            const u32 extent = (u32)((s64)a - diag);
            if (extent > E) { E = extent; twin_ent e = { extent, diag, 1u, pos2 }; mine[nnew++] = e; }     /* :1996-2002 */
            const s32 sim = leftScore + rightScore;
            if (sim < P.K) continue;
            cand_rec r;
            r.hit1 = pos1; r.hit2 = pos2; r.pos1 = pos1 - leftLen; r.pos2 = pos2 - leftLen;
            r.length = leftLen + rightLen; r.score = sim; r.cA = r.cC = r.cG = r.cT = 0;
            if (P.entropy && sim <= 3 * P.K) {
                for (u32 i = 0; i < r.length; i++) {
                    const u8 x = asc1[r.pos1 + i];
                    if (x == asc2[r.pos2 + i]) { r.cA += x == 'A'; r.cC += x == 'C'; r.cG += x == 'G'; r.cT += x == 'T'; }
                }
            }
            const u32 slot = (u32)atomicAdd(&cnt->ncand, 1ull);
            if (slot < candCap) cand[slot] = r;
        }
        diagEnd[h] = E;
        /* what later chunks may still need: the newest entries down to the first one that every later hit (pos2 >= the last
         * one seen here) finds more than maxSpan columns back */
        u32 keep = 0;
        for (u32 k = nnew + nold; k-- > 0; keep++) {
            const twin_ent q = k >= nold ? mine[k - nold] : old[k];
            if ((s64)lastPos2 - ((s64)q.pos2 - (s64)L) > (s64)maxSpan) break;
        }
        if (keep > carryCap) { atomicAdd(&cnt->overflow, 1ull); keep = carryCap; }
        /* the last `keep` entries of (old, mine) move to the front of old: every read is at or behind the write, in order */
        for (u32 j = 0; j < keep; j++) { const u32 k = nnew + nold - keep + j; const twin_ent q = k >= nold ? mine[k - nold] : old[k]; old[j] = q; }
        ncarry[h] = keep;
    }
    if (nExt) atomicAdd(&cnt->extensions, nExt);
    if (nBp) atomicAdd(&cnt->bpExtended, nBp);
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

#endif
