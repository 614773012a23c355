/*
 * xdrop_split.cuh -- K3 as three kernels (included by seed_search.cu).
 *
 * k_extend fuses the right scan, the bucket replay and the left scan, 32 hits of one bucket at a
 * time, in lockstep; ncu showed 19 of 32 lanes active on average because the scan lengths in a batch
 * differ (profiles/r01_k_extend_after.txt).  The split keeps every lane busy:
 *
 *   k_right   lane-independent state machine: each lane fetches its next hit as soon as it has
 *             finished the previous one and advances it one 8-base chunk per loop trip.  Right scans do
 *             not depend on the bucket state.  Hits the bucket had already passed when the chunk began
 *             are skipped (diagEnd only grows); a scan that is still going after XS_CAP columns is
 *             flagged and finished later, only if the hit turns out to be live, so repeats cannot
 *             make the work quadratic.
 *   k_replay  one warp per bucket, discovery order: the diagEnd test/update of
 *             process_for_simple_hit (seed_search.c:1113, :2785) from the precomputed extents; live hits
 *             are appended (with their left-stop) to a dense list.
 *   k_left    the same state machine over the live list: left scan, score, threshold, candidates.
 */
#define XS_CAP 512u
#define XS_FLAG_CAPPED 0x80000000u

struct right_rec { u32 cols; s32 score; u32 len; };          /* columns examined (| flag), best prefix */
struct live_rec { u32 pos1, pos2, stop; u32 rcols; s32 rscore; u32 rlen; };

__device__ __forceinline__ s32 xs_pair(const s32* __restrict__ lut, u64 pr, int i) { return lut[(u32)(pr >> (8 * i)) & 255u]; }

/* one right scan from (pos1,pos2), continuing from state; stops at cap columns if cap != 0 */
__device__ __forceinline__ void xs_scan_right(const s32* __restrict__ lut, const u8* __restrict__ cls1,
                                              const u8* __restrict__ cls2, u32 pos1, u32 pos2, u32 avail, s32 xDrop,
                                              u32& cols, s32& run, s32& best, u32& bestLen, bool& going) {
    u64 x1 = ld8(cls1, pos1 + cols), x2 = ld8(cls2, pos2 + cols);
    u64 pr = (x1 << 4) | x2;
    u32 n = avail - cols; if (n > 8) n = 8;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (going && (u32)i < n) {
            if (run >= best - xDrop) { run += xs_pair(lut, pr, i); cols++; if (run > best) { best = run; bestLen = cols; } }
            else going = false;
        }
    }
}

__global__ void __launch_bounds__(256, 4)
k_right(const u64* __restrict__ hits, u32 nhits, const u8* __restrict__ cls1, const u8* __restrict__ cls2,
        const lzb_scoring_dev* __restrict__ sc, sp_dev P, const u32* __restrict__ diagEnd, right_rec* __restrict__ out) {
    __shared__ s32 lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = sc->msubC[(i >> 4) * LZB_MAX_CLASSES + (i & 15)];
    __syncthreads();
    const u32 stride = gridDim.x * blockDim.x;
    const u32 hmask = (1u << P.hashBits) - 1, L = (u32)P.L;
    const s32 xDrop = P.xDrop;
    u32 next = blockIdx.x * blockDim.x + threadIdx.x;      /* my next hit */
    bool active = false;
    u32 cur = 0, pos1 = 0, pos2 = 0, avail = 0, cols = 0, bestLen = 0; s32 run = 0, best = 0; bool going = false;
    while (true) {
        if (!active) {
            if (next >= nhits) break;
            cur = next; next += stride;
            const u64 rec = hits[cur];
            pos1 = (u32)rec; pos2 = (u32)(rec >> 32);
            const s64 diag = (s64)pos1 - (s64)pos2;
            const s64 lim = (s64)P.len2 + diag;
            const u32 rstop = ((s64)P.len1 <= lim) ? P.len1 : (u32)lim;
            avail = rstop > pos1 ? rstop - pos1 : 0;
            cols = 0; run = 0; best = 0; bestLen = 0; going = true; active = true;
            /* the bucket had already passed this hit when the chunk began: it can never be live */
            if (diagEnd[(u32)diag & hmask] > pos2 - L) { right_rec r = { 0, 0, 0 }; out[cur] = r; active = false; continue; }
        }
        if (going && cols < avail && cols < XS_CAP) xs_scan_right(lut, cls1, cls2, pos1, pos2, avail, xDrop, cols, run, best, bestLen, going);
        if (!going || cols >= avail || cols >= XS_CAP) {
            const bool capped = going && cols < avail;     /* stopped only because of the cap */
            right_rec r = { cols | (capped ? XS_FLAG_CAPPED : 0u), best, bestLen };
            out[cur] = r;
            active = false;
        }
    }
}

__global__ void __launch_bounds__(256)
k_replay(const u64* __restrict__ hits, const u32* __restrict__ bstart, u32 nbuckets, right_rec* __restrict__ right,
         const u8* __restrict__ cls1, const u8* __restrict__ cls2, const lzb_scoring_dev* __restrict__ sc, sp_dev P,
         u32* __restrict__ diagEnd, live_rec* __restrict__ live, unsigned long long* __restrict__ nlive) {
    __shared__ s32 lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = sc->msubC[(i >> 4) * LZB_MAX_CLASSES + (i & 15)];
    __syncthreads();
    const u32 lane = threadIdx.x & 31, L = (u32)P.L, FULL = 0xFFFFFFFFu;
    u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 h = warp; h < nbuckets; h += nwarps) {
        const u32 b0 = bstart[h], b1 = bstart[h + 1];
        if (b0 == b1) continue;
        u32 E = diagEnd[h];
        for (u32 base = b0; base < b1; base += 32) {
            const u32 idx = base + lane;
            const bool have = idx < b1;
            const u64 rec = have ? hits[idx] : 0;
            const u32 pos1 = (u32)rec, pos2 = (u32)(rec >> 32);
            right_rec rr = { 0, 0, 0 };
            if (have) rr = right[idx];
            bool maybe = have && !(E > pos2 - L);
            /* a capped scan must be completed before its extent can enter the bucket state; do it only
             * for hits that can still be live (rare: long identical stretches) */
            if (maybe && (rr.cols & XS_FLAG_CAPPED)) {
                const s64 diag = (s64)pos1 - (s64)pos2;
                const s64 lim = (s64)P.len2 + diag;
                const u32 rstop = ((s64)P.len1 <= lim) ? P.len1 : (u32)lim;
                const u32 avail = rstop > pos1 ? rstop - pos1 : 0;
                u32 cols = 0, bestLen = 0; s32 run = 0, best = 0; bool going = true;
                while (going && cols < avail) xs_scan_right(lut, cls1, cls2, pos1, pos2, avail, P.xDrop, cols, run, best, bestLen, going);
                rr.cols = cols; rr.score = best; rr.len = bestLen;
            }
            const u32 rcols = rr.cols & ~XS_FLAG_CAPPED;
            const u32 ext = pos2 + rcols;                  /* seq-2 coordinate where the right scan stopped */
            bool lv = false; u32 myStop = 0;
            u32 act = __ballot_sync(FULL, maybe);
            while (act) {
                const int k = __ffs(act) - 1; act &= act - 1;
                const u32 p2 = __shfl_sync(FULL, pos2, k), ex = __shfl_sync(FULL, ext, k);
                const bool l2 = !(E > p2 - L);
                if ((int)lane == k) { lv = l2; myStop = E; }
                if (l2 && ex > E) E = ex;
            }
            const u32 lm = __ballot_sync(FULL, lv);
            if (lm) {
                unsigned long long at = 0;
                if (lane == 0) at = atomicAdd(nlive, (unsigned long long)__popc(lm));
                at = __shfl_sync(FULL, at, 0);
                if (lv) {
                    live_rec o = { pos1, pos2, myStop, rcols, rr.score, rr.len };
                    live[at + __popc(lm & ((1u << lane) - 1))] = o;
                }
            }
        }
        if (lane == 0) diagEnd[h] = E;
    }
}

__global__ void __launch_bounds__(256, 4)
k_left(const live_rec* __restrict__ live, const unsigned long long* __restrict__ nlivePtr, const u8* __restrict__ cls1, const u8* __restrict__ cls2,
       const u8* __restrict__ asc1, const u8* __restrict__ asc2, const lzb_scoring_dev* __restrict__ sc, sp_dev P,
       cand_rec* __restrict__ cand, u32 candCap, search_counters* cnt) {
    __shared__ s32 lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = sc->msubC[(i >> 4) * LZB_MAX_CLASSES + (i & 15)];
    __syncthreads();
    const u32 stride = gridDim.x * blockDim.x;
    const s32 xDrop = P.xDrop;
    const u32 nlive = (u32)*nlivePtr;
    u32 next = blockIdx.x * blockDim.x + threadIdx.x;
    bool active = false;
    live_rec h = { 0, 0, 0, 0, 0, 0 };
    u32 avail = 0, cols = 0, bestLen = 0; s32 run = 0, best = 0; bool going = false;
    unsigned long long nExt = 0, nBp = 0;
    while (true) {
        if (!active) {
            if (next >= nlive) break;
            h = live[next]; next += stride;
            const s64 diag = (s64)h.pos1 - (s64)h.pos2;
            const s64 blk = (s64)h.stop + diag;            /* left stop: the bucket's previous extent on this diagonal (:2612-2616) */
            const u32 stop = blk > 0 ? (u32)blk : 0;
            avail = h.pos1 > stop ? h.pos1 - stop : 0;
            cols = 0; run = 0; best = 0; bestLen = 0; going = true; active = true;
        }
        if (going && cols < avail) {
            const u32 a = h.pos1 - cols, b = h.pos2 - cols;  /* next columns are a-1, a-2, ... */
            u32 n = avail - cols; if (n > 8) n = 8;
            if (a >= 8 && b >= 8) {
                const u64 x1 = ld8(cls1, a - 8), x2 = ld8(cls2, b - 8);
                const u64 pr = (x1 << 4) | x2;
#pragma unroll
                for (int i = 7; i >= 0; i--) {
                    if (going && (u32)(7 - i) < n) {
                        if (run >= best - xDrop) { run += xs_pair(lut, pr, i); cols++; if (run > best) { best = run; bestLen = cols; } }
                        else going = false;
                    }
                }
            } else {
                for (u32 i = 0; i < n && going; i++) {
                    if (run >= best - xDrop) { run += lut[((u32)cls1[a - 1 - i] << 4) | cls2[b - 1 - i]]; cols++; if (run > best) { best = run; bestLen = cols; } }
                    else going = false;
                }
            }
        }
        if (!going || cols >= avail) {
            active = false;
            nExt++; nBp += h.rcols + cols;
            const s32 sim = best + h.rscore;
            if (sim >= P.K) {
                cand_rec r;
                r.hit1 = h.pos1; r.hit2 = h.pos2; r.pos1 = h.pos1 - bestLen; r.pos2 = h.pos2 - bestLen;
                r.length = bestLen + h.rlen; r.score = sim; r.cA = r.cC = r.cG = r.cT = 0;
                if (P.entropy && sim <= 3 * P.K) {
                    for (u32 i = 0; i < r.length; i++) {
                        const u8 x = asc1[r.pos1 + i];
                        if (x == asc2[r.pos2 + i]) { r.cA += x == 'A'; r.cC += x == 'C'; r.cG += x == 'G'; r.cT += x == 'T'; }
                    }
                }
                const u32 slot = (u32)atomicAdd(&cnt->ncand, 1ull);
                if (slot < candCap) cand[slot] = r;
            }
        }
    }
    /* per-warp totals (all lanes reach this point) */
    const u32 lane = threadIdx.x & 31;
    for (int o = 16; o > 0; o >>= 1) { nExt += __shfl_down_sync(0xFFFFFFFFu, nExt, o); nBp += __shfl_down_sync(0xFFFFFFFFu, nBp, o); }
    if (lane == 0) { if (nExt) atomicAdd(&cnt->extensions, nExt); if (nBp) atomicAdd(&cnt->bpExtended, nBp); }
}
