/*
 * ydrop_warp.cuh -- K5, the one-warp Y-drop sweep (included by gapped.cu).
 *
 * Same mathematics as k_ydrop (see gapped.cu) with the opposite placement.  A one-sided extension
 * is a strictly serial chain of rows, the gapped stage is a handful of very long such chains
 * (50 Mbp x 50 Mbp: 177 anchors x 2 sides x ~220 000 rows), so what counts is the LATENCY of one
 * row, not throughput.  k_ydrop spends ~700 instructions per warp per row at ~7 cycles each
 * (profiles/r01_k_ydrop_after.txt): eight warps that mostly wait on each other at four block
 * barriers and on shared-memory round trips.  Here ONE warp owns the whole band:
 *
 *   - lane l owns K consecutive columns; column block b (columns b*K .. b*K+K-1) lives in lane
 *     b mod 32, so the 32*K-column window slides right with the band without moving any data: a
 *     lane whose block has fallen off the left edge re-initialises itself 32 blocks further right;
 *   - C, D (previous row) and the query-side class codes stay in registers; the only memory
 *     traffic per row is the score-table lookups (shared), the traceback bytes (global, write
 *     only) and one class code of the target per row (prefetched 32 rows at a time);
 *   - the three row-wide dependencies (insertion chain, in-row bestScore threshold, live range)
 *     are warp scans/reductions in the rotated lane order -- no barrier anywhere;
 *   - every per-column step is select-based (no divergent branches), so the sixteen columns of a
 *     lane give the scheduler independent work: the warp issues most cycles instead of one in seven.
 *
 * Bands that outgrow the window return DP_RING and the host reruns that extension with a wider
 * window (K = 24) and then with the shared-memory kernel, so nothing is approximated.
 *
 * Dead cells: every column outside the live band holds LZB_NEG_INF in C and D after every row (pass
 * 3 writes all K columns), which is also what the reference's sentinel cell provides (:3822-3827).
 * Intermediate values of dead cells are not clamped (they may sit a few thousand below
 * LZB_NEG_INF); they lose every comparison against live values and are reset at the end of the row.
 */
template <int K> struct wg_state { s32 C[K], D[K]; u32 Bq[K / 4]; u32 cb; };

struct wg_in {
    u32 lane, LY, colEnd, row, M, N;
    const s32* subRow; const u32* stamp; u8* tbRowPtr;     /* tbRowPtr[col] = this row's traceback byte of column col */
    s32 gapE, gapOE, yDrop, best; int trim;
};
struct wg_out { u32 fa, la, uc, bc; s32 uv, bv, Iout; };

#define WG_SCAP 1024u                  /* stamp ring (columns), >= the widest window */


/* one row of the sweep, gapped_extend.c:3669-3774 */
template <int K, bool MASKING>
__device__ __forceinline__ void wg_sweep(wg_state<K>& S, const wg_in& in, wg_out& out) {
    const u32 FULL = 0xFFFFFFFFu, lane = in.lane, cb = S.cb;
    const s32 gapE = in.gapE, gapOE = in.gapOE;
    /* my in-band columns; the band's first column takes no diagonal (:3683) */
    const u32 lo = in.LY > cb ? min(in.LY - cb, (u32)K) : 0u;
    const u32 hi = in.colEnd > cb ? min(in.colEnd - cb, (u32)K) : 0u;
    const u32 im = hi > lo ? (wg_lowmask(hi) & ~wg_lowmask(lo)) : 0u;
    const u32 dmk = (in.LY >= cb && in.LY - cb < (u32)K) ? (im & ~(1u << (in.LY - cb))) : im;
    u32 mm = 0;                                            /* columns on an earlier alignment (:3690) */
    if (MASKING) {
#pragma unroll
        for (int s = 0; s < K; s++) if (((im >> s) & 1u) && in.stamp[(cb + s) & (WG_SCAP - 1)] == in.row) mm |= 1u << s;
    }
    const u32 lane0 = (in.LY / (u32)K) & 31u, li = (lane - lane0) & 31u;   /* my rank in column order */
    /* ---- pass 1: diagonal proposals, the insertion-open values a(j) and my piece of the chain ---- */
    s32 dg[K], av[K];
    s32 Iin, Iout;
    {
        s32 leftC = __shfl_sync(FULL, S.C[K - 1], (lane + 31u) & 31u);
        s32 vmax = LZB_NEG_INF;
        xf mine; mine.A = LZB_NEG_INF; mine.S = 0; mine.r = 0;
        const s32 ebase = gapE * ((s32)cb - (s32)in.LY + 1);
#pragma unroll
        for (int s = 0; s < K; s++) {
            const u32 bc = (S.Bq[s >> 2] >> (8 * (s & 3))) & 255u;
            const s32 diag = ((dmk >> s) & 1u) ? leftC + in.subRow[bc] : LZB_NEG_INF;
            leftC = S.C[s];
            const s32 d = ((im >> s) & 1u) ? S.D[s] : LZB_NEG_INF;
            S.D[s] = d;
            s32 a = diag >= d ? diag - gapOE : LZB_NEG_INF;
            if (MASKING && ((mm >> s) & 1u)) a = LZB_NEG_INF;
            dg[s] = diag; av[s] = a;
            if (!MASKING) vmax = max(vmax, a + ebase + gapE * s);
            else {
                xf g;                                      /* out-of-band columns are the identity: Iout is I at colEnd */
                if ((mm >> s) & 1u) { g.A = LZB_NEG_INF; g.S = 0; g.r = 1; }
                else if ((im >> s) & 1u) { g.A = a; g.S = -gapE; g.r = 0; }
                else { g.A = LZB_NEG_INF; g.S = 0; g.r = 0; }
                mine = xf_then(mine, g);
            }
        }
        if (!MASKING) {
            /* shifted form I'(j) = I(j) + e*(j-LY): the chain is a running max (see k_ydrop) */
            s32 inc = vmax;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const s32 u = __shfl_sync(FULL, inc, (lane - o) & 31u); if (li >= (u32)o) inc = max(inc, u); }
            s32 ex = __shfl_sync(FULL, inc, (lane + 31u) & 31u);
            if (li == 0) ex = LZB_NEG_INF;
            const s32 tot = __shfl_sync(FULL, inc, (lane0 + 31u) & 31u);
            Iin = ex - gapE * ((s32)cb - (s32)in.LY);
            Iout = tot - gapE * (s32)(in.colEnd > in.LY ? in.colEnd - in.LY : 0u);
            if (Iout < LZB_NEG_INF) Iout = LZB_NEG_INF;
        } else {
            xf inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                xf up; const u32 src = (lane - o) & 31u;
                up.A = __shfl_sync(FULL, inc.A, src); up.S = __shfl_sync(FULL, inc.S, src); up.r = __shfl_sync(FULL, inc.r, src);
                if (li >= (u32)o) inc = xf_then(up, inc);
            }
            xf exl; const u32 src = (lane + 31u) & 31u;
            exl.A = __shfl_sync(FULL, inc.A, src); exl.S = __shfl_sync(FULL, inc.S, src); exl.r = __shfl_sync(FULL, inc.r, src);
            if (li == 0) { exl.A = LZB_NEG_INF; exl.S = 0; exl.r = 0; }
            Iin = exl.A;
            Iout = __shfl_sync(FULL, inc.A, (lane0 + 31u) & 31u);
        }
    }
    /* ---- pass 2: cell values, links, next row's D; candidates for bestScore ---- */
    u32 fl[K / 4];
    s32 cand[K];                                           /* c where the diagonal won inside the band, else -inf */
#pragma unroll
    for (int w = 0; w < K / 4; w++) fl[w] = 0;
    s32 candMax = LZB_NEG_INF;
    {
        s32 I = Iin;
#pragma unroll
        for (int s = 0; s < K; s++) {
            const s32 diag = dg[s], d = S.D[s], a = av[s];
            const s32 m = max(d, I);
            const bool gap = m > diag;                     /* a gap beats the diagonal (ties: diagonal) */
            s32 c = gap ? m : diag;
            const s32 ii = I - gapE, dx = d - gapE;
            s32 In = max(a, ii);                           /* = ii whenever a gap won (a <= I - oe) */
            s32 Dn = gap ? dx : max(a, dx);
            u32 f = gap ? ((d >= I ? LINK_D : LINK_I) | LINK_IEXT | LINK_DEXT)
                        : ((a > dx ? 0u : LINK_DEXT) | (a > ii ? 0u : LINK_IEXT));
            s32 cd = gap ? LZB_NEG_INF : diag;             /* out-of-band columns: diag = -inf already */
            if (MASKING && ((mm >> s) & 1u)) { c = LZB_NEG_INF; Dn = LZB_NEG_INF; In = LZB_NEG_INF; f = 0; cd = LZB_NEG_INF; }
            candMax = max(candMax, cd);
            dg[s] = c; av[s] = Dn; cand[s] = cd; fl[s >> 2] |= f << (8 * (s & 3)); I = In;
        }
    }
    /* exclusive prefix max of the candidates in column order, seeded with bestScore */
    s32 B;
    {
        s32 pm = candMax;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const s32 u = __shfl_sync(FULL, pm, (lane - o) & 31u); if (li >= (u32)o) pm = max(pm, u); }
        B = __shfl_sync(FULL, pm, (lane + 31u) & 31u);
        if (li == 0) B = LZB_NEG_INF;
        B = max(B, in.best);
    }
    /* ---- pass 3: prune against the running bestScore (:3737-3745), band edges, best/end ----
     * a candidate that reaches B is alive by construction, so B is a plain running max; masked
     * and out-of-band cells carry -inf and fail the threshold test on their own */
    u32 am = 0; u32 upCol1 = 0;
#pragma unroll
    for (int s = 0; s < K; s++) {
        const bool alive = ((im >> s) & 1u) && dg[s] >= B - in.yDrop;
        if (alive) am |= 1u << s;
        if (cand[s] >= B) upCol1 = cb + s + 1;
        B = max(B, cand[s]);
    }
    const s32 upVal = upCol1 ? B : -1;
    u8* const tp = in.tbRowPtr + cb;
#pragma unroll
    for (int s = 0; s < K; s++) {
        const bool alive = (am >> s) & 1u;
        const u32 f = (fl[s >> 2] >> (8 * (s & 3))) & 15u;
        S.C[s] = alive ? dg[s] : LZB_NEG_INF; S.D[s] = alive ? av[s] : LZB_NEG_INF;
        if ((im >> s) & 1u) tp[s] = alive ? (u8)f : (u8)0;
    }
    s32 bVal = LZB_NEG_INF; u32 bCol1 = 0;
    if (!in.trim) {                                        /* boundaryScore :3747-3750 */
#pragma unroll
        for (int s = 0; s < K; s++) {
            if (((am >> s) & 1u) && cand[s] > LZB_NEG_INF && (in.row == in.M || cb + s == in.N) && S.C[s] >= bVal) { bVal = S.C[s]; bCol1 = cb + s + 1; }
        }
    }
    out.fa = __reduce_min_sync(FULL, am ? cb + (u32)__ffs(am) - 1u : 0xFFFFFFFFu);
    out.la = __reduce_max_sync(FULL, am ? cb + 32u - (u32)__clz(am) : 0u);
    out.uv = __reduce_max_sync(FULL, upVal);
    out.uc = __reduce_max_sync(FULL, (upVal == out.uv && upCol1) ? upCol1 : 0u);
    out.bv = LZB_NEG_INF; out.bc = 0;
    if (!in.trim) {
        out.bv = __reduce_max_sync(FULL, bVal);
        out.bc = __reduce_max_sync(FULL, (bVal == out.bv && bCol1) ? bCol1 : 0u);
    }
    out.Iout = Iout;
}

template <int K>
__global__ void __launch_bounds__(32)
k_ydrop_warp(dp_job* jobs, const launch_list ll, const dseg* __restrict__ segs,
             const u8* __restrict__ cls1, const u8* __restrict__ cls2, u32 len1, u32 len2,
             const lzb_scoring_dev* __restrict__ sc, s32 yDrop, int trim) {
    static_assert(K % 4 == 0 && K <= 32 && 32u * K <= WG_SCAP, "window must fit the stamp ring");
    constexpr u32 WIN = 32u * K, smsk = WG_SCAP - 1;
    __shared__ s32 subC[LZB_MAX_CLASSES * LZB_MAX_CLASSES];
    __shared__ u32 stamp[WG_SCAP];
    const u32 lane = threadIdx.x, FULL = 0xFFFFFFFFu;
    dp_job* J = &jobs[ll.ix[blockIdx.x]];
    const dalign* __restrict__ al = J->al;
    for (u32 i = lane; i < LZB_MAX_CLASSES * LZB_MAX_CLASSES; i += 32) subC[i] = sc->subC[i];
    for (u32 i = lane; i < WG_SCAP; i += 32) stamp[i] = 0;
    const int rev = J->reversed; const u32 a1 = J->a1, a2 = J->a2, M = J->M, N = J->N;
    const s32 gapE = sc->gapExtend, gapOE = sc->gapOpen + sc->gapExtend;
    const u8 cls0 = sc->cls[0];
    u8* tb = J->tb; const s64 tbLen = J->tbLen; u32* tbRow = J->tbRow;
    int status = DP_OK;
    s32 best = 0, bnd = LZB_NEG_INF; u32 end1 = 0, end2 = 0; int endIsBnd = 0;
    unsigned long long cells = 0; u32 row = 0;
    if (N == 0 || M == 0) {
        if (lane == 0) { J->score = 0; J->end1 = J->end2 = 0; J->nops = 0; J->rows = 0; J->cells = 0; J->status = DP_OK; J->opsOverflow = 0; J->ckptCount = 0; job_done(J); }
        return;
    }
    const s32 yTail = gapE != 0 ? yDrop / gapE + 6 : (N < 500000u ? (s32)N + 1 : 500000);
    s32 L = J->L0, R = J->R0;
    segref leftSeg = J->leftSeg, rightSeg = J->rightSeg;
    const int* const listv = J->listv;
    int alignList = J->alignList;                          /* index into listv */
    u32 nextActRow = list_row(listv, alignList, al, rev, a1);
    const int tbOnly = J->tbOnly, resume = J->resume;
    if (tbOnly) { status = J->status; end1 = J->end1; end2 = J->end2; }
    int* act = J->act; int nact = 0;
    constexpr u32 CKW = CK_RECORD_WORDS;
    u32* const ckpt = J->ckpt; const u32 ckptCap = J->ckptCap, ckptEvery = J->ckptEvery;
    u32 ckptCount = tbOnly ? J->ckptCount : 0;
    const u32 rowLimit = J->rowLimit ? J->rowLimit : 0xFFFFFFFFu;
    const u32 tbRowCap = J->tbRowCap, actCap = J->actCap;
    u32* const dbg = J->dbg; const u32 dbgCap = J->dbgCap;
    u32 lLim = 0, rLim = 0; int lTyp = 0, rTyp = 0;
    LOAD_BOUND(leftSeg, lLim, lTyp); LOAD_BOUND(rightSeg, rLim, rTyp);
    s64 used = 0;
    wg_state<K> S;
    /* query-side class codes of my block (B(col), gapped_extend.c:2512-2527) */
#define WG_LOAD_BLOCK()                                                                                   \
    do {                                                                                                  \
        _Pragma("unroll") for (int w_ = 0; w_ < K / 4; w_++) S.Bq[w_] = 0;                                \
        _Pragma("unroll") for (int s_ = 0; s_ < K; s_++) {                                                \
            const u32 col_ = S.cb + s_;                                                                   \
            const u32 code_ = (col_ <= N) ? (u32)cls2[!rev ? (a2 + col_) : (a2 + 1 - col_)] : (u32)cls0;  \
            S.Bq[s_ >> 2] |= code_ << (8 * (s_ & 3));                                                     \
        }                                                                                                 \
    } while (0)
    S.cb = lane * K;
    /* ---- first row, gapped_extend.c:3576-3591 -- or the state a checkpoint saved ---- */
    u32 LY = 0, RY = 0, row0 = 1;
    if (resume >= 0 && !tbOnly) {
        const u32* rec = ckpt + (size_t)resume * CKW;
        row0 = rec[0] + 1; LY = rec[1]; RY = rec[2]; L = (s32)rec[3]; R = (s32)rec[4];
        leftSeg.al = (int)rec[5]; leftSeg.sg = (int)rec[6]; rightSeg.al = (int)rec[7]; rightSeg.sg = (int)rec[8];
        lLim = rec[9]; rLim = rec[10]; lTyp = (int)rec[11]; rTyp = (int)rec[12]; nact = (int)rec[13];
        used = (s64)((u64)rec[14] | ((u64)rec[15] << 32));
        best = (s32)rec[16]; bnd = (s32)rec[17]; end1 = rec[18]; end2 = rec[19]; endIsBnd = (int)rec[20];
        cells = (u64)rec[21] | ((u64)rec[22] << 32);
        if (lane == 0) for (int k = 0; k < 5 * nact; k++) act[k] = (int)rec[CK_HDR + k];
        /* my columns of the band, wherever the writer kept them: first the block this thread owns at that row */
        S.cb = lane * K;
        while (S.cb + K <= LY) S.cb += WIN;                 /* the rule of the row loop: a block left of the band sits 32 blocks further right */
        const u32 c0 = rec[26];
        const u32* tv = rec + CK_HDR + 5 * CK_ACT;
#pragma unroll
        for (int s = 0; s < K; s++) {
            const u32 ix = S.cb + (u32)s - c0;
            const bool in = S.cb + (u32)s >= c0 && ix < CK_COLS;
            S.C[s] = in ? (s32)tv[ix] : LZB_NEG_INF; S.D[s] = in ? (s32)tv[CK_COLS + ix] : LZB_NEG_INF;
        }
        if ((u64)RY + 2 > (u64)(LY / (u32)K) * K + WIN) status = DP_RING;          /* written by a kernel with a wider window */
        ckptCount = (u32)resume + 1;
        row = row0;
    } else {
        u32 last = 1;
        if (gapE > 0) { if (yDrop >= gapOE) last = (u32)(((s64)yDrop - gapOE) / gapE) + 2; }
        else if (yDrop >= gapOE) last = N;
        if (last > N) last = N;
        if ((u64)last + 3 > WIN) status = DP_RING;
#pragma unroll
        for (int s = 0; s < K; s++) {
            const u32 col = S.cb + s;
            if (col <= last && status == DP_OK) {
                const s32 v = col == 0 ? 0 : -gapOE - (s32)(col - 1) * gapE;
                S.C[s] = v; S.D[s] = v - gapOE;
                if (!tbOnly) tb[col] = col == 0 ? 0 : LINK_I;
            } else { S.C[s] = LZB_NEG_INF; S.D[s] = LZB_NEG_INF; }
        }
        used = (s64)last + 1;
        RY = last + 1; cells = RY;                        /* the first row counts too, :3593 */
        if (lane == 0 && tbRowCap > 0 && !tbOnly) tbRow[0] = 0;
    }
    WG_LOAD_BLOCK();
    /* target-side class codes, 32 rows per load, one chunk ahead */
#define WG_ACODE(r_) ([&]() -> u32 { const u32 rr_ = (r_); if (rr_ > M) return (u32)cls0;                 \
                                      const s64 ai_ = !rev ? (s64)a1 + rr_ : (s64)a1 + 1 - (s64)rr_;      \
                                      return (ai_ < 0 || ai_ >= (s64)len1) ? (u32)cls0 : (u32)cls1[ai_]; }())
    /* (a checkpoint row is a multiple of 32: the loop's first iteration then shifts acvNext into acv) */
    u32 acv = WG_ACODE(((row0 - 1) & ~31u) + (row0 > 1 ? -31 : 1) + lane), acvNext = WG_ACODE(((row0 - 1) & ~31u) + (row0 > 1 ? 1 : 33) + lane);
    __syncwarp();
    u32 nextCk = ckptCap ? (ckptCount + 1u) * ckptEvery : 0xFFFFFFFFu;
    if (status == DP_OK && !tbOnly)
    for (row = row0; row <= M; row++) {
        if (row >= rowLimit) { status = DP_PAUSED; break; }
        if ((row & 255u) == 0 && J->abort) { status = DP_ABORTED; break; }   /* the anchor was retired (mapped host memory: looked at rarely) */
        /* ---- update_LR_bounds gapped_extend.c:4588-4724 (every lane, same values) ---- */
        if (!rev) {
            if (leftSeg.al >= 0) {
                if (lLim >= row + a1) { if (lTyp == SEG_DIAG) L++; }
                else { L = sweep_step(al, segs, 0, 0, &leftSeg, row, a1, a2) + 1; LOAD_BOUND(leftSeg, lLim, lTyp); }
            }
            if (leftSeg.al >= 0) { if ((s32)LY < L) LY = (u32)L; }
            if (rightSeg.al >= 0) {
                if (rLim >= row + a1) { if (rTyp == SEG_DIAG) R++; }
                else { R = sweep_step(al, segs, 0, 1, &rightSeg, row, a1, a2) - 1; LOAD_BOUND(rightSeg, rLim, rTyp); }
            }
            if (rightSeg.al >= 0) { if (R <= 0) RY = 0; else if ((u32)R < RY) RY = (u32)R; }
        } else {
            if (rightSeg.al >= 0) {
                if (rLim <= a1 - row) { if (rTyp == SEG_DIAG) L++; }
                else { L = sweep_step(al, segs, 1, 1, &rightSeg, row, a1, a2) + 1; LOAD_BOUND(rightSeg, rLim, rTyp); }
            }
            if (rightSeg.al >= 0) { if ((s32)LY < L) LY = (u32)L; }
            if (leftSeg.al >= 0) {
                if (lLim <= a1 - row) { if (lTyp == SEG_DIAG) R++; }
                else { R = sweep_step(al, segs, 1, 0, &leftSeg, row, a1, a2) - 1; LOAD_BOUND(leftSeg, lLim, lTyp); }
            }
            if (leftSeg.al >= 0) { if (R <= 0) RY = 0; else if ((u32)R < RY) RY = (u32)R; }
        }
        /* ---- update_active_segs gapped_extend.c:4885-4962 (lane 0; the list is tiny) ---- */
        if (nact > 0 || row == nextActRow) {
            if (lane == 0)
                active_update(act, &nact, actCap, listv, &alignList, &nextActRow, &status, al, segs, rev, stamp, smsk, row, a1, a2, LY, RY);
            __syncwarp();
            nact = __shfl_sync(FULL, nact, 0); alignList = __shfl_sync(FULL, alignList, 0); nextActRow = __shfl_sync(FULL, nextActRow, 0); status = __shfl_sync(FULL, status, 0);
            if (status != DP_OK) break;
        }
        /* ---- traceback capacity gapped_extend.c:3636-3662 ---- */
        if (RY < LY) RY = LY;
        const s64 need = (s64)(RY - LY) + yTail;
        if (used + need >= tbLen) { status = DP_TRUNCATED; break; }
        if (row >= tbRowCap) { status = DP_TBROW; break; }
        const u32 tbBase = (u32)((u64)used - (u64)LY);
        if (lane == 0) tbRow[row] = tbBase;
        /* ---- the sweep ---- */
        const u32 leftCol = LY;
        const u32 colEnd = RY < N + 1 ? RY : N + 1;
        if (((row - 1) & 31u) == 0 && row > 1) { acv = acvNext; acvNext = WG_ACODE(row + 32 + lane); }
        const u32 ac = __shfl_sync(FULL, acv, (row - 1) & 31u);
        wg_in in;
        in.lane = lane; in.LY = LY; in.colEnd = colEnd; in.row = row; in.M = M; in.N = N;
        in.subRow = subC + ac * LZB_MAX_CLASSES; in.stamp = stamp; in.tbRowPtr = tb + (used - (s64)LY);
        in.gapE = gapE; in.gapOE = gapOE; in.yDrop = yDrop; in.best = best; in.trim = trim;
        u8* const rowPtr = in.tbRowPtr;
        wg_out o;
        if (nact > 0) wg_sweep<K, true>(S, in, o); else wg_sweep<K, false>(S, in, o);
        /* bestScore moves to the LAST cell (row-major) that equalled the row's final best (:3742) */
        u32 bestCol = 0; bool bestMoved = false;
        if (o.uc) { best = o.uv; bestCol = o.uc - 1; bestMoved = true; }
        u32 bndCol = 0; bool bndMoved = false;
        if (!trim && o.bc && o.bv >= bnd) { bnd = o.bv; bndCol = o.bc - 1; bndMoved = true; }
        if (bestMoved && (!bndMoved || bestCol > bndCol)) { end1 = row; end2 = bestCol; endIsBnd = 0; }
        else if (bndMoved) { end1 = row; end2 = bndCol; endIsBnd = 1; }
        cells += colEnd - leftCol;
        used += colEnd - leftCol;
        if (dbg && lane == 0 && row < dbgCap) { u32* g = dbg + 4 * (size_t)row; g[0] = leftCol; g[1] = colEnd; g[2] = (u32)best; g[3] = (u32)used; }
        u32 npCol;
        if (o.la) { LY = o.fa; npCol = o.la - 1; } else { LY = colEnd; npCol = leftCol; }
        if (LY >= RY) break;
        /* ---- row end, gapped_extend.c:3789-3827 ---- */
        const s32 NN = (rightSeg.al >= 0 && R > 0) ? R - 1 : (s32)N;
        const u32 wcol = colEnd; u32 p = 0;
        if (RY > npCol + 1) RY = npCol + 1;
        else {
            const s32 thr = best - yDrop;
            if (o.Iout >= thr && (s32)RY <= NN) {
                const u32 room = (u32)(NN - (s32)RY) + 1;
                const u32 byScore = gapE > 0 ? (u32)((o.Iout - thr) / gapE) + 1 : room;
                p = byScore < room ? byScore : room;
            }
        }
        /* the window follows the band: [first block holding LY, +32 blocks) must hold the next row */
        if ((u64)(RY > wcol ? RY : wcol) + p + 2 > (u64)(LY / (u32)K) * K + WIN) { status = DP_RING; break; }
        while (S.cb + K <= LY) {
            S.cb += WIN;
#pragma unroll
            for (int s = 0; s < K; s++) { S.C[s] = LZB_NEG_INF; S.D[s] = LZB_NEG_INF; }
            WG_LOAD_BLOCK();
        }
        if (p) {
            /* prolong the row with insertions (:3801-3816): each lane patches the columns it owns */
            const u32 plo = wcol > S.cb ? min(wcol - S.cb, (u32)K) : 0u;
            const u32 phi = wcol + p > S.cb ? min(wcol + p - S.cb, (u32)K) : 0u;
            const u32 pmk = phi > plo ? (wg_lowmask(phi) & ~wg_lowmask(plo)) : 0u;
            if (pmk) {
                s32 v = o.Iout - ((s32)S.cb - (s32)wcol) * gapE;
                u8* const prow = rowPtr + S.cb;
#pragma unroll
                for (int s = 0; s < K; s++) {
                    if ((pmk >> s) & 1u) { S.C[s] = v; S.D[s] = v - gapOE; prow[s] = LINK_I; }
                    v -= gapE;
                }
            }
            RY += p; used += p;
        }
        if ((s32)RY <= NN) RY++;                             /* the sentinel column already holds LZB_NEG_INF */
        /* ---- checkpoint: everything the next row reads ---- */
        if (row == nextCk && ckptCount < ckptCap && nact <= CK_ACT) {     /* record k belongs to row (k+1)*ckptEvery; a record that cannot be written ends the series */
            u32* rec = ckpt + (size_t)ckptCount * CKW;
            if (lane == 0) {
                rec[0] = row; rec[1] = LY; rec[2] = RY; rec[3] = (u32)L; rec[4] = (u32)R;
                rec[5] = (u32)leftSeg.al; rec[6] = (u32)leftSeg.sg; rec[7] = (u32)rightSeg.al; rec[8] = (u32)rightSeg.sg;
                rec[9] = lLim; rec[10] = rLim; rec[11] = (u32)lTyp; rec[12] = (u32)rTyp; rec[13] = (u32)nact;
                rec[14] = (u32)(u64)used; rec[15] = (u32)((u64)used >> 32);
                rec[16] = (u32)best; rec[17] = (u32)bnd; rec[18] = end1; rec[19] = end2; rec[20] = (u32)endIsBnd;
                rec[21] = (u32)cells; rec[22] = (u32)(cells >> 32); rec[26] = LY & ~31u;
                for (int k = 0; k < 5 * nact; k++) rec[CK_HDR + k] = (u32)act[k];
                J->progUsed = (u32)(u64)used; J->progRows = row;      /* lets the host estimate where the traceback will run out */
            }
            const u32 c0 = LY & ~31u;                        /* base column of the record */
            u32* tv = rec + CK_HDR + 5 * CK_ACT;
            for (u32 i = lane; i < 2 * CK_COLS; i += 32u) tv[i] = (u32)LZB_NEG_INF;
            __syncwarp();
#pragma unroll
            for (int s = 0; s < K; s++) {
                const u32 ix = S.cb + (u32)s - c0;
                if (S.cb + (u32)s >= c0 && ix < CK_COLS) { tv[ix] = (u32)S.C[s]; tv[CK_COLS + ix] = (u32)S.D[s]; }
            }
            ckptCount++; nextCk += ckptEvery;
        }
    }
#undef WG_LOAD_BLOCK
#undef WG_ACODE
    /* ---- traceback, gapped_extend.c:3847-3859 ---- */
    __threadfence();
    __syncwarp();
    u32 nops = 0; bool ovf = false;
    if (status == DP_OK || status == DP_TRUNCATED)
        nops = traceback_walk(tb, tbRow, end1, end2, J->ops, J->opsCap, lane, &ovf);
    if (lane == 0) {
        if (!tbOnly) {
            J->score = endIsBnd ? bnd : best; J->end1 = end1; J->end2 = end2;
            J->rows = row; J->cells = cells; J->status = status; J->ckptCount = ckptCount;
        }
        J->nops = nops; J->opsOverflow = ovf ? 1 : 0;
        job_done(J);
    }
}
