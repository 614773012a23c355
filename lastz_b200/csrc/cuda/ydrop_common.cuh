/*
 * ydrop_common.cuh -- device-side vocabulary shared by the Y-drop kernels (k_ydrop in gapped.cu,
 * k_ydrop_warp, k_ydrop_mw): job descriptors, the device view of committed alignments, the affine
 * max-plus map of the masked insertion chain, the bound walks over neighbouring alignments
 * (next/prev_sweep_seg gapped_extend.c:4754-4853), active-segment stamps (build_active_seg :4989) and
 * the warp traceback (:3847-3859).  Kept free of host/runtime code so that the kernels can also be
 * compiled for the host block emulator (tests/warp_emu/cuda_emu.h).
 */
#ifndef LZB_YDROP_COMMON_CUH
#define LZB_YDROP_COMMON_CUH

enum { SEG_DIAG = 0, SEG_HORZ = 1, SEG_VERT = 2 };

struct dseg { u32 b1, b2, e1, e2; int type; };
struct segref { int al, sg; };
struct dalign {                       /* device view of a committed alignment (galign :214-245) */
    u32 pos1, end1;
    int segBegin, segCount;
    segref left1, right1, left2, right2;
};

enum { DP_OK = 0, DP_TRUNCATED = 1, DP_RING = 2, DP_TBROW = 3, DP_OPS = 4, DP_ACT = 5, DP_ABORTED = 6, DP_PAUSED = 7 };

/* jobs of one launch: block b runs jobs[ll.ix[b]] (the list travels in the kernel's parameter space, so a launch
 * needs no device-side table that would have to outlive it) */
#define LZB_LAUNCH_MAX 960
struct launch_list { u16 ix[LZB_LAUNCH_MAX]; };

/* checkpoint record of a sweep, written every ckptEvery rows (a multiple of 32): CK_HDR scalar words, the active
 * segments (5 ints each, at most CK_ACT; a row with more takes no checkpoint), then the band BY COLUMN: C and D of the
 * CK_COLS columns from the record's base column on (word rec[26], at or left of the band's first column).  Indexed by
 * column, a record does not depend on which kernel wrote it: a sweep started by the one-warp kernel (512-column
 * window) can be continued by the four-warp kernel (1024 columns), which is what a sweep that is about to run along an
 * earlier alignment needs -- the dead flank beside that alignment widens the band for a few hundred rows.
 * A sweep restarted from record k continues at row (k+1)*ckptEvery + 1 exactly as if it had never stopped, against a
 * LARGER set of earlier alignments whose first rows all lie beyond that row (see gapped_sched.hpp). */
#define CK_HDR 32
#define CK_ACT 8
#define CK_COLS 1024u
#define CK_RECORD_WORDS (CK_HDR + 5 * CK_ACT + 2 * CK_COLS)

struct dp_job {
    int reversed; u32 a1, a2, M, N;
    s32 L0, R0;
    segref leftSeg, rightSeg;
    const int* listv;                 /* alignments this sweep may run into, device table indices in sweep order
                                         (aboveList / belowList, gapped_extend.c:4043-4060), -1 terminated; NULL = none */
    int alignList;                    /* first pending entry of listv */
    u8* tb; u32 tbLen; u32* tbRow; u32 tbRowCap; u32* ops; u32 opsCap;
    int* act; u32 actCap;             /* 5 ints per active segment */
    const dalign* al;                 /* alignment table (append-only: a sweep only follows indices it was given) */
    u32* ckpt; u32 ckptCap, ckptEvery;/* checkpoint records (k_ydrop_mw only); ckptCap = 0: none taken */
    int resume;                       /* -1 fresh sweep, else the checkpoint record to continue from */
    u32 rowLimit;                     /* stop (DP_PAUSED) before this row: the host expects an earlier anchor's alignment there and
                                         continues the sweep from a checkpoint once that alignment is known (0: no limit) */
    int tbOnly;                       /* nonzero: the sweep is done (end1/end2/status below are its results); redo the traceback walk */
    u32* dbg; u32 dbgCap;             /* LZB_DP_DEBUG: per-row {LY, colEnd, best, used} for kernel-vs-kernel diffs */
    u32 token;                        /* written to `done` after everything else */
    volatile int abort;               /* set by the host while the sweep runs: its result is no longer wanted */
    /* results */
    s32 score; u32 end1, end2, nops, rows; int status; int opsOverflow; unsigned long long cells;
    u32 ckptCount;                    /* records 0 .. ckptCount-1 are valid */
    volatile u32 progRows, progUsed;  /* written with every checkpoint while the sweep runs: rows done, traceback bytes used */
    volatile u32 done;
};

struct xf { s32 A; s32 S; int r; };   /* x -> r ? A : max(A, x + S) */

__device__ __forceinline__ s32 satadd(s32 a, s32 s) { s32 v = a + s; return v < LZB_NEG_INF ? LZB_NEG_INF : v; }
__device__ __forceinline__ xf xf_then(xf f, xf g) {      /* apply f, then g */
    xf o;
    if (g.r) return g;
    o.A = max(g.A, satadd(f.A, g.S)); o.S = f.S + g.S; o.r = f.r;
    return o;
}

#define LINK_I 1
#define LINK_D 2
#define LINK_IEXT 4
#define LINK_DEXT 8
#define F_CAND 16
#define F_MASK 32

/* next_sweep_seg / prev_sweep_seg gapped_extend.c:4754-4853 */
__device__ s32 sweep_step(const dalign* al, const dseg* segs, int rev, int lookRight, segref* bp,
                          u32 row, u32 a1, u32 a2) {
    const dalign m = al[bp->al];
    if (!rev) {
        if (bp->sg + 1 < m.segCount) {
            bp->sg++;
            if (segs[m.segBegin + bp->sg].type == SEG_HORZ) bp->sg++;
            return (s32)(segs[m.segBegin + bp->sg].b2 - a2);
        }
        *bp = lookRight ? m.right2 : m.left2;
        if (bp->al < 0) return 0;
        const dseg s = segs[al[bp->al].segBegin + bp->sg];
        if (s.type == SEG_DIAG) return (s32)row + (s32)(s.b2 - a2) - (s32)(s.b1 - a1);
        return (s32)(s.b2 - a2);
    }
    if (bp->sg - 1 >= 0) {
        bp->sg--;
        if (segs[m.segBegin + bp->sg].type == SEG_HORZ) bp->sg--;
        return (s32)(a2 - segs[m.segBegin + bp->sg].e2);
    }
    *bp = lookRight ? m.right1 : m.left1;
    if (bp->al < 0) return 0;
    const dseg s = segs[al[bp->al].segBegin + bp->sg];
    if (s.type == SEG_DIAG) return (s32)row + (s32)(a2 - s.e2) - (s32)(a1 - s.e1);
    return (s32)(a2 - s.e2);
}

/* build_active_seg gapped_extend.c:4989-5040; act record = {al, sg, x, lastRow, type} */
__device__ void act_build(int* a, const dalign* al, const dseg* segs, int rev, u32* stamp, u32 msk,
                          u32 row, u32 a1, u32 a2, u32 LY, u32 RY) {
    const dseg s = segs[al[a[0]].segBegin + a[1]];
    a[4] = s.type;
    u32 x, lastRow;
    if (!rev) { x = s.b2 - a2; lastRow = s.e1 - a1; } else { x = a2 - s.e2; lastRow = a1 - s.b1; }
    a[2] = (int)x; a[3] = (int)lastRow;
    if (s.type != SEG_HORZ) { if (x >= LY && x <= RY) stamp[x & msk] = row; }
    else {
        u32 hend = !rev ? s.e2 - a2 : a2 - s.b2;
        u32 lo = x > LY ? x : LY, hi = hend < RY ? hend : RY;
        for (u32 i = lo; i <= hi && i >= lo; i++) stamp[i & msk] = row;
    }
}


/* row (counted from the anchor) at which entry li of a sweep's list starts: aboveList entries are taken when
 * pos1 - anchor1 == row, belowList entries when anchor1 - end1 == row (gapped_extend.c:4927-4945) */
__device__ __forceinline__ u32 list_row(const int* listv, int li, const dalign* al, int rev, u32 a1) {
    const int ix = listv ? listv[li] : -1;
    if (ix < 0) return 0xFFFFFFFFu;
    return !rev ? al[ix].pos1 - a1 : a1 - al[ix].end1;
}

/* update_active_segs gapped_extend.c:4885-4962, by ONE thread: advance the active segments to `row`, take the
 * alignments that start at this row off the list, drop the finished ones.  *li indexes listv; *nextRow = row at
 * which the next list entry starts (0xFFFFFFFF: none), so callers only come here when something is active. */
__device__ void active_update(int* act, int* pnact, u32 actCap, const int* listv, int* li, u32* nextRow, int* status,
                              const dalign* al, const dseg* segs, int rev, u32* stamp, u32 msk,
                              u32 row, u32 a1, u32 a2, u32 LY, u32 RY) {
    int nact = *pnact;
    for (int k = 0; k < nact; k++) {
        int* a = act + 5 * k;
        if ((u32)a[3] >= row) {
            if (a[4] == SEG_DIAG) a[2]++;
            u32 x = (u32)a[2];
            if (x >= LY && x <= RY) stamp[x & msk] = row;
        } else {
            int cnt = al[a[0]].segCount;
            bool more = !rev ? (a[1] + 1 < cnt) : (a[1] - 1 >= 0);
            if (more) {
                a[1] += !rev ? 1 : -1;
                act_build(a, al, segs, rev, stamp, msk, row, a1, a2, LY, RY);
                if (a[4] == SEG_HORZ) { a[1] += !rev ? 1 : -1; act_build(a, al, segs, rev, stamp, msk, row, a1, a2, LY, RY); }
            } else a[4] = -1;
        }
    }
    while (*nextRow == row) {
        if ((u32)nact >= actCap) { *status = DP_ACT; break; }
        const int ix = listv[*li];
        int* a = act + 5 * nact; nact++;
        a[0] = ix; a[1] = !rev ? 0 : al[ix].segCount - 1;
        act_build(a, al, segs, rev, stamp, msk, row, a1, a2, LY, RY);
        (*li)++;
        *nextRow = list_row(listv, *li, al, rev, a1);
    }
    int w = 0;
    for (int k = 0; k < nact; k++) if (act[5 * k + 4] >= 0) { if (w != k) for (int z = 0; z < 5; z++) act[5 * w + z] = act[5 * k + z]; w++; }
    *pnact = w;
}

/* publish a job's results: everything first, then the token the host polls for */
__device__ __forceinline__ void job_done(dp_job* J) {
#ifndef CUDA_EMU_H
    __threadfence_system();
#endif
    J->done = J->token;
}

/* traceback, gapped_extend.c:3847-3859, by one warp: lane t speculates that the path continues
 * diagonally and looks at (r-t, c-t); a ballot finds the first non-substitution, so a run of up to
 * 32 substitutions costs one dependent load.  Emits run-length ops (op | count<<2) in walk order. */
__device__ u32 traceback_walk(const u8* tb, const u32* tbRow, u32 end1, u32 end2, u32* ops, u32 opsCap,
                              u32 lane, bool* overflow) {
    const u32 FULL = 0xFFFFFFFFu;
    u32 nops = 0;
    u32 r = end1, c = end2; u32 prevOp = 0;
    u32 curOp = 0, curCnt = 0; bool ovf = false;
    while (r >= 1 || c > 0) {
        bool inb = (r >= lane) && (c >= lane) && ((r - lane) >= 1 || (c - lane) > 0);
        u32 link = 0;
        if (inb) link = tb[(u32)(tbRow[r - lane] + (c - lane))];
        u32 op = link & 3;
        if (lane == 0) {
            if (prevOp == LINK_I && (link & LINK_IEXT)) op = LINK_I;
            if (prevOp == LINK_D && (link & LINK_DEXT)) op = LINK_D;
        }
        /* lane t>0 assumes the step before it was a substitution, true iff all earlier lanes are
         * substitutions; a diagonal step needs r-t >= 1 and c-t >= 1 */
        bool isSub = inb && op == 0 && (r - lane) >= 1 && (c - lane) >= 1;
        u32 notSub = __ballot_sync(FULL, !isSub);
        u32 run = notSub ? (u32)(__ffs(notSub) - 1) : 32;
        if (run > 0) {
            if (curOp == LZB_OP_SUB) curCnt += run;
            else { if (curCnt) { if (nops < opsCap) { if (lane == 0) ops[nops] = curOp | (curCnt << 2); } else ovf = true; nops++; } curOp = LZB_OP_SUB; curCnt = run; }
            r -= run; c -= run; prevOp = 0;
            continue;
        }
        u32 op0 = __shfl_sync(FULL, op, 0);
        u32 eop;
        if (op0 == LINK_I) { c--; eop = LZB_OP_INS; }
        else if (op0 == LINK_D) { r--; eop = LZB_OP_DEL; }
        else { r--; c--; eop = LZB_OP_SUB; }
        if (curOp == eop) curCnt++;
        else { if (curCnt) { if (nops < opsCap) { if (lane == 0) ops[nops] = curOp | (curCnt << 2); } else ovf = true; nops++; } curOp = eop; curCnt = 1; }
        prevOp = op0;
    }
    if (curCnt) { if (nops < opsCap) { if (lane == 0) ops[nops] = curOp | (curCnt << 2); } else ovf = true; nops++; }
    *overflow = ovf;
    return nops;
}

/* far edge (e1 forward, b1 reversed) and type of a bounding segment; expects al, segs, rev in scope */
#define LOAD_BOUND(ref_, lim_, typ_) \
    do { if ((ref_).al >= 0) { const dseg s_ = segs[al[(ref_).al].segBegin + (ref_).sg]; lim_ = rev ? s_.b1 : s_.e1; typ_ = s_.type; } } while (0)

__device__ __forceinline__ u32 wg_lowmask(u32 n) { return n >= 32u ? 0xFFFFFFFFu : (1u << n) - 1u; }

#endif
