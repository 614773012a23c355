/*
 * xdrop_warp.cuh -- K3b (second generation): one warp replays one diag-hash bucket.
 *
 * Semantics are those of process_for_simple_hit (seed_search.c:1056-1192) feeding
 * xdrop_extend_seed_hit (:2528-2959), hit by hit in discovery order, for the hits of ONE bucket
 * (the only state hits share is diagEnd[bucket]).  What changed against the first k_extend:
 *
 *   - lanes scan their own hit only for the first XD_CAP columns (lockstep, 8 columns per trip);
 *     a scan that is still alive after that -- a real HSP, hundreds of columns -- is finished by
 *     the WHOLE warp, 256 columns per trip: prefix sums and running maxima by warp scans, the first
 *     column at which the reference's loop test fails by ballot.  One lane walking a 600-column
 *     HSP at ~35 cycles per column was the critical path of the kernel (ncu: SMs 36 % active,
 *     profiles/r01_k_extend_tail.txt);
 *   - right scans are finished lazily inside the replay: in a bucket that holds a homologous
 *     diagonal the first live hit's extent kills the 31 hits behind it, so their long scans are
 *     never completed;
 *   - the entropy match counts (dna_utilities.c:2905-2915) are taken by the warp, not by one lane.
 *
 * All functions are called by all 32 lanes (full-mask collectives, see warp_ops.cuh); the same
 * source runs on the host lane emulator (tests/warp_emu) for bit-exact unit tests.
 * Needs: u8/u32/u64/s32/s64 typedefs, LZB_GFEX_*.
 */
#ifndef LZB_XDROP_WARP_CUH
#define LZB_XDROP_WARP_CUH
#include "warp_ops.cuh"

#ifndef XD_CAP
#define XD_CAP 128u                    /* columns a lane scans alone before the warp takes over */
#endif
#define XD_NEG ((s32)-0x3FFFFFFF)
/* pair table index = class1 * 20 + class2 (<= 13 classes, so < 256).  With A,C,G,T as classes 0..3
 * (context.cu numbers them first) the 16 common pairs fall into 16 different shared-memory banks;
 * at stride 16 they shared 8 and every lookup of a warp on 32 unrelated diagonals conflicted
 * (ncu: L1/shared pipe 77 % busy, 590 M bank conflicts per launch). */
#define XD_LUT_STRIDE 20u
#define XD_LUT_MAX_CLASSES 13

struct cand_rec {                      /* one HSP candidate, 40 bytes */
    u32 hit1, hit2;                    /* the seed hit (one past its end) that produced it */
    u32 pos1, pos2, length;            /* HSP start + length */
    s32 score;
    u32 cA, cC, cG, cT;                /* exact-match counts by base (entropy) */
};

struct xd_env {
    const u8* cls1; const u8* cls2;    /* class-coded sequences (<= 16 classes) */
    const u8* asc1; const u8* asc2;    /* the bytes themselves (entropy counts) */
    const s32* lut;                    /* 256 entries: maskedScoring by class1 * XD_LUT_STRIDE + class2 */
    u32 len1, len2, L;                 /* L = seed length */
    s32 xDrop, K; int entropy;
    cand_rec* cand; u32 candCap; unsigned long long* ncand;
};

/* eight consecutive bytes starting at any index (two aligned 64-bit loads + funnel shift) */
W_DEV u64 xd_ld8(const u8* p, u32 idx) {
    const u64* w = (const u64*)(p + (idx & ~7u));
    const u32 sh = (idx & 7u) * 8u;
    const u64 lo = w[0], hi = w[1];
    return sh ? (lo >> sh) | (hi << (64u - sh)) : lo;
}

/* scores of scan columns o .. o+7 (scan-relative; n of them exist) of the scan that starts at
 * (p1,p2): DIR=+1 reads p1+o, p1+o+1, ...; DIR=-1 reads p1-1-o, p1-2-o, ...  Missing columns score 0. */
template <int DIR, bool FULL = false>
W_DEV void xd_fetch8(const xd_env& e, u32 p1, u32 p2, u32 o, u32 n, s32 (&s)[8]) {
    if (DIR > 0 && FULL) {                                       /* all eight columns exist: no per-column predicates */
        const u64 x1 = xd_ld8(e.cls1, p1 + o), x2 = xd_ld8(e.cls2, p2 + o);
        const u64 pr = (x1 << 4) + (x1 << 2) + x2;
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = e.lut[(u32)(pr >> (8 * i)) & 255u];
        return;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = 0;
    if (n == 0) return;
    if (DIR > 0) {
        const u64 x1 = xd_ld8(e.cls1, p1 + o), x2 = xd_ld8(e.cls2, p2 + o);
        const u64 pr = (x1 << 4) + (x1 << 2) + x2;              /* bytewise class1 * 20 + class2; no byte overflows */
#pragma unroll
        for (int i = 0; i < 8; i++) if ((u32)i < n) s[i] = e.lut[(u32)(pr >> (8 * i)) & 255u];
    } else {
        const u32 a = p1 - o, b = p2 - o;                       /* columns a-1, a-2, ... */
        if (n == 8 && a >= 8 && b >= 8) {
            const u64 x1 = xd_ld8(e.cls1, a - 8), x2 = xd_ld8(e.cls2, b - 8);
            const u64 pr = (x1 << 4) + (x1 << 2) + x2;
#pragma unroll
            for (int i = 0; i < 8; i++) s[i] = e.lut[(u32)(pr >> (8 * (7 - i))) & 255u];
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) if ((u32)i < n) s[i] = e.lut[(u32)e.cls1[a - 1 - i] * XD_LUT_STRIDE + e.cls2[b - 1 - i]];
        }
    }
}

/* one lane's own scan state; the reference loop is
 *     while (cols < avail && run >= best - xDrop) { run += score(col); cols++; if (run > best) { best = run; bestLen = cols; } }
 * A scan is CLOSED when cols == avail: a failed loop test closes it by setting avail = cols. */
struct xd_scan { u32 avail, cols, bestLen; s32 run, best; };

W_DEV void xd_scan_init(xd_scan& z, u32 avail) { z.avail = avail; z.cols = 0; z.bestLen = 0; z.run = 0; z.best = 0; }
W_DEV bool xd_scan_open(const xd_scan& z) { return z.cols < z.avail; }

/* advance my (open) scan by up to 8 columns; written as straight-line selects so that the column
 * step compiles to predicated instructions (the first version's nested ifs became a BSSY/BRA/BSYNC
 * block per column: 16 instructions a column, profiles/r01_k_extend2_sass.txt) */
template <int DIR, bool FULL>
W_DEV void xd_scan_cols(const s32 (&s)[8], u32 n, s32 xDrop, xd_scan& z) {
    /* d = best - run >= 0 is the whole recurrence: the loop test is d <= xDrop, a new best is d - score < 0.
     * If column i is consumed so were all before it, hence cols = cols0 + i + 1 with a literal offset. */
    s32 d = z.best - z.run, best = z.best; const u32 cols0 = z.cols; u32 cols = cols0, bestLen = z.bestLen;
    bool alive = true;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const bool p = alive && (FULL || (u32)i < n) && d <= xDrop;
        d = p ? d - s[i] : d;
        const bool q = p && d < 0;
        best = q ? best - d : best;
        d = w_max(d, 0);
        cols = p ? cols0 + (u32)(i + 1) : cols;
        bestLen = q ? cols0 + (u32)(i + 1) : bestLen;
        alive = p;
    }
    z.run = best - d; z.best = best; z.cols = cols; z.bestLen = bestLen;
    if (!alive) z.avail = cols;                         /* the loop test failed (or the columns ran out): closed */
}
template <int DIR>
W_DEV void xd_scan_step8(const xd_env& e, u32 p1, u32 p2, s32 xDrop, xd_scan& z) {
    u32 n = z.avail - z.cols; if (n > 8) n = 8;
    s32 s[8];
    if (n == 8) { xd_fetch8<DIR, true>(e, p1, p2, z.cols, n, s); xd_scan_cols<DIR, true>(s, n, xDrop, z); }
    else { xd_fetch8<DIR, false>(e, p1, p2, z.cols, n, s); xd_scan_cols<DIR, false>(s, n, xDrop, z); }
}

/* The warp finishes ONE scan; every lane passes the same state and receives the same result. */
template <int DIR>
W_DEV void xd_coop_finish(const xd_env& e, u32 lane, u32 p1, u32 p2, xd_scan& z) {
    while (z.cols < z.avail) {
        const u32 o = z.cols + lane * 8u;
        const u32 n = o < z.avail ? w_min(z.avail - o, 8u) : 0u;
        s32 ps[8]; xd_fetch8<DIR>(e, p1, p2, o, n, ps);
#pragma unroll
        for (int i = 1; i < 8; i++) ps[i] += ps[i - 1];                       /* inclusive prefix inside the lane */
        const s32 tot = ps[7];
        s32 incl = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const s32 t = W_SHFL_UP(incl, d); if (lane >= (u32)d) incl += t; }
        const s32 R = z.run + incl - tot;                                      /* run before my first column */
        s32 m = XD_NEG;
#pragma unroll
        for (int i = 0; i < 8; i++) if ((u32)i < n) m = w_max(m, R + ps[i]);
        s32 minc = m;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const s32 t = W_SHFL_UP(minc, d); if (lane >= (u32)d) minc = w_max(minc, t); }
        s32 xm = W_SHFL_UP(minc, 1); if (lane == 0) xm = XD_NEG;
        /* first of my columns in front of which the loop test fails, had the scan reached it */
        u32 t = 8; s32 rb = R, bb = w_max(z.best, xm);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if ((u32)i < n) {
                if (t == 8 && rb < bb - e.xDrop) t = (u32)i;
                rb = R + ps[i]; bb = w_max(bb, rb);
            }
        }
        const u32 term = W_BALLOT(t < 8);
        u32 consumed = w_min(z.avail - z.cols, 256u);
        if (term) { const u32 f = (u32)W_FFS(term) - 1u; consumed = f * 8u + W_SHFL(t, f); }
        /* best run over the consumed columns: the first column that reaches the maximum */
        s32 bv = XD_NEG; u32 bi = 0xFFFFFFFFu;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const u32 gi = lane * 8u + (u32)i;
            if ((u32)i < n && gi < consumed) { const s32 r = R + ps[i]; if (r > bv) { bv = r; bi = gi; } }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const s32 ov = W_SHFL_XOR(bv, d); const u32 oi = W_SHFL_XOR(bi, d);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (bv > z.best) { z.best = bv; z.bestLen = z.cols + bi + 1u; }
        if (consumed > 0) {
            const u32 gl = consumed - 1u;
            s32 mine = R;
#pragma unroll
            for (int i = 0; i < 8; i++) if ((gl & 7u) == (u32)i) mine = R + ps[i];
            z.run = W_SHFL(mine, gl >> 3);
        }
        z.cols += consumed;
        if (term) z.avail = z.cols;                      /* closed by the loop test */
    }
}

/* broadcast lane k's scan, finish it with the whole warp, hand it back */
template <int DIR>
W_DEV void xd_finish_lane(const xd_env& e, u32 lane, int k, u32 pos1, u32 pos2, xd_scan& mine) {
    xd_scan z;
    const u32 p1 = W_SHFL(pos1, k), p2 = W_SHFL(pos2, k);
    z.avail = W_SHFL(mine.avail, k); z.cols = W_SHFL(mine.cols, k); z.bestLen = W_SHFL(mine.bestLen, k);
    z.run = W_SHFL(mine.run, k); z.best = W_SHFL(mine.best, k);
    xd_coop_finish<DIR>(e, lane, p1, p2, z);
    if ((int)lane == k) { mine = z; mine.avail = z.cols; }
}

/* all hits [b0,b1) of one bucket, in discovery order; E = diagEnd[bucket] in and out */
W_DEV void xd_bucket(const xd_env& e, u32 lane, const u64* hits, u32 b0, u32 b1, u32& E,
                     unsigned long long& nExt, unsigned long long& nBp) {
    const u32 L = e.L; const s32 xDrop = e.xDrop;
    for (u32 base = b0; base < b1; base += 32) {
        const u32 idx = base + lane;
        const bool have = idx < b1;
        const u64 rec = have ? hits[idx] : 0;
        const u32 pos1 = (u32)rec, pos2 = (u32)(rec >> 32);
        const s64 diag = (s64)pos1 - (s64)pos2;
        /* hits the bucket has already passed can never be live (diagEnd only grows) */
        const bool maybe = have && !(E > pos2 - L);
        u32 active = W_BALLOT(maybe);
        if (!active) continue;
        /* right scan (seed_search.c:2663-2693), independent of the bucket state */
        xd_scan rs; xd_scan_init(rs, 0);
        if (maybe) {
            const s64 lim = (s64)e.len2 + diag;
            const u32 rstop = ((s64)e.len1 <= lim) ? e.len1 : (u32)lim;
            xd_scan_init(rs, rstop > pos1 ? rstop - pos1 : 0);
            while (xd_scan_open(rs) && rs.cols < XD_CAP) xd_scan_step8<1>(e, pos1, pos2, xDrop, rs);
        }
        /* replay the test/update of process_for_simple_hit (:1113, :2785-2789) in discovery order */
        bool live = false; u32 myStop = 0;
        if (!W_BALLOT(maybe && xd_scan_open(rs))) {
            /* every extent is known.  Suppose all remaining candidates contribute their extent: the
             * bucket as hit k meets it is E0 max'ed with the extents in front of k (a prefix maximum).
             * Lanes in front of the first hit that comes out dead saw exactly the true bucket, so that
             * hit IS dead; drop its extent and look again.  One trip per dead hit (usually none or one). */
            const u32 ext = pos2 + rs.cols;                           /* where the right scan stopped, in seq 2 */
            bool contrib = maybe;
            for (;;) {
                u32 pm = contrib ? ext : 0u;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const u32 t = W_SHFL_UP(pm, d); if (lane >= (u32)d) pm = w_max(pm, t); }
                u32 before = W_SHFL_UP(pm, 1); if (lane == 0) before = 0u;
                before = w_max(before, E);
                const u32 dm = W_BALLOT(contrib && before > pos2 - L);
                if (!dm) { live = contrib; myStop = before; E = w_max(E, W_SHFL(pm, 31)); break; }
                if ((int)lane == W_FFS(dm) - 1) contrib = false;
            }
        } else {
            /* some scan outlived XD_CAP columns: walk the hits one by one and let the warp finish a
             * long scan only when its hit turns out to be live */
            while (active) {
                const int k = W_FFS(active) - 1; active &= active - 1;
                const u32 p2k = W_SHFL(pos2, k);
                if (E > p2k - L) continue;                             /* dead by now; its scan is never finished */
                if (W_SHFL(xd_scan_open(rs) ? 1 : 0, k)) xd_finish_lane<1>(e, lane, k, pos1, pos2, rs);
                const u32 ex = W_SHFL(pos2 + rs.cols, k);
                if ((int)lane == k) { live = true; myStop = E; }
                if (ex > E) E = ex;
            }
        }
        /* left scan (:2598-2632), blocked by the bucket's previous extent on this diagonal */
        xd_scan ls; xd_scan_init(ls, 0);
        if (live) {
            const s64 blk = (s64)myStop + diag;
            const u32 stop = blk > 0 ? (u32)blk : 0;
            xd_scan_init(ls, pos1 > stop ? pos1 - stop : 0);
            while (xd_scan_open(ls) && ls.cols < XD_CAP) xd_scan_step8<-1>(e, pos1, pos2, xDrop, ls);
        }
        u32 longLeft = W_BALLOT(live && xd_scan_open(ls));
        while (longLeft) {
            const int k = W_FFS(longLeft) - 1; longLeft &= longLeft - 1;
            xd_finish_lane<-1>(e, lane, k, pos1, pos2, ls);
        }
        s32 sim = 0; bool keep = false;
        cand_rec r;
        r.hit1 = pos1; r.hit2 = pos2; r.pos1 = pos1 - ls.bestLen; r.pos2 = pos2 - ls.bestLen;
        r.length = ls.bestLen + rs.bestLen; r.cA = r.cC = r.cG = r.cT = 0;
        if (live) {
            nExt++; nBp += rs.cols + ls.cols;
            sim = ls.best + rs.best;
            keep = sim >= e.K;                                         /* entropy can only lower the score */
        }
        r.score = sim;
        /* match counts for entropy(), dna_utilities.c:2905-2915: the warp counts, 32 columns per trip */
        u32 want = W_BALLOT(keep && e.entropy && sim <= 3 * e.K);
        while (want) {
            const int k = W_FFS(want) - 1; want &= want - 1;
            const u32 q1 = W_SHFL(r.pos1, k), q2 = W_SHFL(r.pos2, k), len = W_SHFL(r.length, k);
            u32 cA = 0, cC = 0, cG = 0, cT = 0;
            for (u32 i0 = 0; i0 < len; i0 += 32) {
                const u32 i = i0 + lane;
                u8 x = 0, y = 1;
                if (i < len) { x = e.asc1[q1 + i]; y = e.asc2[q2 + i]; }
                const bool eq = x == y;
                cA += (u32)W_POPC(W_BALLOT(eq && x == 'A')); cC += (u32)W_POPC(W_BALLOT(eq && x == 'C'));
                cG += (u32)W_POPC(W_BALLOT(eq && x == 'G')); cT += (u32)W_POPC(W_BALLOT(eq && x == 'T'));
            }
            if ((int)lane == k) { r.cA = cA; r.cC = cC; r.cG = cG; r.cT = cT; }
        }
        if (keep) {
            const u32 slot = (u32)W_ATOMIC_ADD_ULL(e.ncand, 1ull);
            if (slot < e.candCap) e.cand[slot] = r;
        }
    }
}

#endif
