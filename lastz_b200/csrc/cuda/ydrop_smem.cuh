/*
 * ydrop_smem.cuh -- K5a, the shared-memory Y-drop sweep (k_ydrop): the fallback for bands wider than the register
 * kernels' windows (ring of 4096 / 8192 columns in dynamic shared memory).  See gapped.cu's header for the
 * mathematics.  Device code only (also compiled for the host block emulator, tests/warp_emu/cuda_emu.h).
 */
#ifndef LZB_YDROP_SMEM_CUH
#define LZB_YDROP_SMEM_CUH

#ifndef LZB_DYNAMIC_SHARED                /* the emulator supplies its own (static storage) */
#define LZB_DYNAMIC_SHARED(name_) extern __shared__ __align__(16) unsigned char name_[]
#endif

#define DP_MAX_WARPS 8

#define DP_KC 5                       /* cells per thread whose inputs are kept in registers */

struct dp_shared {                    /* small block-wide exchange area */
    xf  wagg[DP_MAX_WARPS];
    s32 wmaxI[DP_MAX_WARPS], wmax[DP_MAX_WARPS];
    u32 wfa[DP_MAX_WARPS], wla[DP_MAX_WARPS], wuc[DP_MAX_WARPS], wbc[DP_MAX_WARPS];
    s32 wuv[DP_MAX_WARPS], wbv[DP_MAX_WARPS];
    int nact, alignList, status; u32 nextActRow;
};

template <int DP_THREADS>
__global__ void __launch_bounds__(DP_THREADS)
k_ydrop(dp_job* jobs, const launch_list ll, const dseg* __restrict__ segs,
        const u8* __restrict__ cls1, const u8* __restrict__ cls2, u32 len1, u32 len2,
        const lzb_scoring_dev* __restrict__ sc, s32 yDrop, int trim, u32 cap) {
    LZB_DYNAMIC_SHARED(smem_raw);
    s32* C0 = (s32*)smem_raw; s32* C1 = C0 + cap; s32* Dv = C1 + cap;
    u32* stamp = (u32*)(Dv + cap); s32* subC = (s32*)(stamp + cap);
    dp_shared* sh = (dp_shared*)(subC + LZB_MAX_CLASSES * LZB_MAX_CLASSES);
    u8* flg = (u8*)(sh + 1);
    constexpr int DP_WARPS = DP_THREADS / 32;
    const u32 msk = cap - 1;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 FULL = 0xFFFFFFFFu;
    dp_job* J = &jobs[ll.ix[blockIdx.x]];
    const dalign* __restrict__ al = J->al;
    for (u32 i = tid; i < LZB_MAX_CLASSES * LZB_MAX_CLASSES; i += DP_THREADS) subC[i] = sc->subC[i];
    for (u32 i = tid; i < cap; i += DP_THREADS) stamp[i] = 0;
    const int rev = J->reversed; const u32 a1 = J->a1, a2 = J->a2, M = J->M, N = J->N;
    const s32 gapE = sc->gapExtend, gapOE = sc->gapOpen + sc->gapExtend;
    const u8 cls0 = sc->cls[0];
    u8* tb = J->tb; const s64 tbLen = J->tbLen; u32* tbRow = J->tbRow;
    int status = DP_OK;
    s32 best = 0, bnd = LZB_NEG_INF; u32 end1 = 0, end2 = 0; int endIsBnd = 0;
    unsigned long long cells = 0; u32 row = 0;
    if (N == 0 || M == 0) {
        if (tid == 0) { J->score = 0; J->end1 = J->end2 = 0; J->nops = 0; J->rows = 0; J->cells = 0; J->status = DP_OK; J->opsOverflow = 0; J->ckptCount = 0; job_done(J); }
        return;
    }
    s32 yTail = gapE != 0 ? yDrop / gapE + 6 : (N < 500000u ? (s32)N + 1 : 500000);
    s32 L = J->L0, R = J->R0;
    segref leftSeg = J->leftSeg, rightSeg = J->rightSeg;
    const int* const listv = J->listv;
    int alignList = J->alignList;                          /* index into listv */
    u32 nextActRow = list_row(listv, alignList, al, rev, a1);
    const int tbOnly = J->tbOnly;
    if (tbOnly) { status = J->status; end1 = J->end1; end2 = J->end2; }
    int* act = J->act; int nact = 0;
    const u32 tbRowCap = J->tbRowCap, actCap = J->actCap;
    u32* const dbg = J->dbg; const u32 dbgCap = J->dbgCap;
    /* far edge (e1 forward, b1 reversed) and type of the two bounding segments */
    u32 lLim = 0, rLim = 0; int lTyp = 0, rTyp = 0;
    LOAD_BOUND(leftSeg, lLim, lTyp); LOAD_BOUND(rightSeg, rLim, rTyp);
    s64 used = 0;
    __syncthreads();
    /* ---- first row, gapped_extend.c:3576-3591 ---- */
    u32 LY = 0, RY;
    {
        /* C[0][c] = -(oe + (c-1)e); col c (>=1) exists iff c <= N and C[0][c-1] >= -yDrop */
        u32 last = 1;
        if (gapE > 0) { if (yDrop >= gapOE) last = (u32)(((s64)yDrop - gapOE) / gapE) + 2; }
        else if (yDrop >= gapOE) last = N;
        if (last > N) last = N;
        if ((s64)last + 1 + yTail + 40 >= (s64)cap) status = DP_RING;
        else {
            for (u32 c = tid; c <= last; c += DP_THREADS) {
                s32 v = c == 0 ? 0 : -gapOE - (s32)(c - 1) * gapE;
                C0[c & msk] = v; Dv[c & msk] = v - gapOE;
                tb[c] = c == 0 ? 0 : LINK_I;
            }
            used = (s64)last + 1;
        }
        RY = last + 1; cells = RY;                        /* the first row counts too, :3593 */
        if (tid == 0 && tbRowCap > 0) tbRow[0] = 0;
    }
    __syncthreads();
    s32* Cprev = C0; s32* Ccur = C1;
    if (status == DP_OK && !tbOnly)
    for (row = 1; row <= M; row++) {
        u32 prevLY = LY;
        if ((row & 255u) == 0 && J->abort) { status = DP_ABORTED; break; }   /* the anchor was retired (mapped host memory: looked at rarely) */
        /* ---- update_LR_bounds gapped_extend.c:4588-4724 (every thread, same values).  The bounding
         * segment's far edge and type sit in registers; HBM is touched only when the walk moves on ---- */
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
        if ((s64)(RY > prevLY ? RY - prevLY : 0) + yTail + 40 >= (s64)cap) { status = DP_RING; break; }
        /* ---- update_active_segs gapped_extend.c:4885-4962 (thread 0; the list is tiny) ---- */
        if (nact > 0 || row == nextActRow) {
            if (tid == 0) {
                active_update(act, &nact, actCap, listv, &alignList, &nextActRow, &status, al, segs, rev, stamp, msk, row, a1, a2, LY, RY);
                sh->nact = nact; sh->alignList = alignList; sh->nextActRow = nextActRow; sh->status = status;
            }
            __syncthreads();
            nact = sh->nact; alignList = sh->alignList; nextActRow = sh->nextActRow; status = sh->status;
            if (status != DP_OK) break;
        }
        /* ---- traceback capacity gapped_extend.c:3636-3662 ---- */
        if (RY < LY) RY = LY;
        s64 need = (s64)(RY - LY) + yTail;
        if (used + need >= tbLen) { status = DP_TRUNCATED; break; }
        if (row >= tbRowCap) { status = DP_TBROW; break; }
        const u32 tbBase = (u32)((u64)used - (u64)LY);            /* tbRow[row], modulo 2^32 like the walk */
        if (tid == 0) tbRow[row] = tbBase;
        /* ---- the sweep, gapped_extend.c:3669-3774 ---- */
        const u32 leftCol = LY;
        const u32 colEnd = RY < N + 1 ? RY : N + 1;
        const u32 width = colEnd > LY ? colEnd - LY : 0;
        const u32 k = ((width + DP_THREADS - 1) / DP_THREADS) | 1;
        const u32 j0 = LY + tid * k;
        const u32 j1 = (j0 + k < colEnd) ? j0 + k : colEnd;       /* may be <= j0: idle thread */
        const s32 ai = !rev ? (s32)(a1 + row) : (s32)(a1 + 1 - row);
        const u8 ac = (ai < 0 || (u32)ai >= len1) ? cls0 : cls1[ai];
        const s32* subRow = subC + ac * LZB_MAX_CLASSES;
        const bool masking = nact > 0;
        const bool cached = k <= DP_KC;                            /* cell inputs stay in registers between passes */
        s32 dg0 = 0, dg1 = 0, dg2 = 0, dg3 = 0, dg4 = 0, dd0 = 0, dd1 = 0, dd2 = 0, dd3 = 0, dd4 = 0;   /* DP_KC of each */
        /* pass 1: diagonal proposals and D inputs; this thread's piece of the insertion chain.
         * Unmasked rows use the shifted form I'(j) = I(j) + e*(j-LY): I'(j+1) = max(a_j + e*(j-LY+1), I'(j)),
         * a plain running max, so the chain is a prefix-max scan.  Rows with masked cells (earlier
         * alignments inside the band) need resets and take the general affine-map scan. */
        s32 Iin, Iout;
        {
            s32 pc = (j0 > LY && j0 < colEnd) ? Cprev[(j0 - 1) & msk] : LZB_NEG_INF;
            s32 vmax = LZB_NEG_INF;
            xf mine; mine.A = LZB_NEG_INF; mine.S = 0; mine.r = 0;
            if (cached) {
#define DP_P1(c_)                                                                                          \
                { const u32 j = j0 + (c_);                                                                   \
                  if (j < j1) {                                                                              \
                    const s32 bi = !rev ? (s32)(a2 + j) : (s32)(a2 + 1 - j);                                 \
                    const u8 bc = (bi < 0) ? cls0 : cls2[bi];                                                \
                    const s32 diag = (j == LY) ? LZB_NEG_INF : pc + subRow[bc];                              \
                    pc = Cprev[j & msk];                                                                     \
                    const s32 d = Dv[j & msk];                                                               \
                    dg##c_ = diag; dd##c_ = d;                                                               \
                    const s32 a = diag >= d ? satadd(diag, -gapOE) : LZB_NEG_INF;                            \
                    if (!masking) vmax = max(vmax, a + gapE * (s32)(j - LY + 1));                            \
                    else {                                                                                   \
                        xf g;                                                                                \
                        if (stamp[j & msk] == row) { g.A = LZB_NEG_INF; g.S = 0; g.r = 1; } else { g.A = a; g.S = -gapE; g.r = 0; } \
                        mine = xf_then(mine, g);                                                             \
                    } } }
                DP_P1(0) DP_P1(1) DP_P1(2) DP_P1(3) DP_P1(4)
#undef DP_P1
            } else {
                for (u32 j = j0; j < j1; j++) {
                    const s32 bi = !rev ? (s32)(a2 + j) : (s32)(a2 + 1 - j);
                    const u8 bc = (bi < 0) ? cls0 : cls2[bi];
                    const s32 diag = (j == LY) ? LZB_NEG_INF : pc + subRow[bc];
                    pc = Cprev[j & msk];
                    const s32 d = Dv[j & msk];
                    const s32 a = diag >= d ? satadd(diag, -gapOE) : LZB_NEG_INF;
                    if (!masking) vmax = max(vmax, a + gapE * (s32)(j - LY + 1));
                    else {
                        xf g;
                        if (stamp[j & msk] == row) { g.A = LZB_NEG_INF; g.S = 0; g.r = 1; } else { g.A = a; g.S = -gapE; g.r = 0; }
                        mine = xf_then(mine, g);
                    }
                }
            }
            if (!masking) {
                /* block-wide exclusive prefix max of vmax */
                s32 inc = vmax;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { s32 u = __shfl_up_sync(FULL, inc, o); if ((int)lane >= o) inc = max(inc, u); }
                if (lane == 31) sh->wmaxI[warp] = inc;
                s32 ex = __shfl_up_sync(FULL, inc, 1);
                if (lane == 0) ex = LZB_NEG_INF;
                __syncthreads();
                s32 tot = LZB_NEG_INF;
#pragma unroll
                for (int w = 0; w < DP_WARPS; w++) { const s32 v = sh->wmaxI[w]; if (w < (int)warp) ex = max(ex, v); tot = max(tot, v); }
                Iin = satadd(ex, -gapE * (s32)(j0 < colEnd ? j0 - LY : 0));
                Iout = satadd(tot, -gapE * (s32)width);
            } else {
                xf inc = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    xf up; up.A = __shfl_up_sync(FULL, inc.A, o); up.S = __shfl_up_sync(FULL, inc.S, o); up.r = __shfl_up_sync(FULL, inc.r, o);
                    if ((int)lane >= o) inc = xf_then(up, inc);
                }
                if (lane == 31) sh->wagg[warp] = inc;
                xf exl; exl.A = __shfl_up_sync(FULL, inc.A, 1); exl.S = __shfl_up_sync(FULL, inc.S, 1); exl.r = __shfl_up_sync(FULL, inc.r, 1);
                if (lane == 0) { exl.A = LZB_NEG_INF; exl.S = 0; exl.r = 0; }
                __syncthreads();
                xf pre; pre.A = LZB_NEG_INF; pre.S = 0; pre.r = 0;
                xf tot = pre;
#pragma unroll
                for (int w = 0; w < DP_WARPS; w++) { xf a = sh->wagg[w]; if (w < (int)warp) pre = xf_then(pre, a); tot = xf_then(tot, a); }
                Iin = xf_then(pre, exl).A;
                Iout = tot.A;
            }
        }
        /* pass 2: cell values, links, next row's D; candidates for bestScore */
        s32 candMax = LZB_NEG_INF;
        {
            s32 I = Iin;
#define DP_CELL(j_, c_, d_)                                                                              \
            {   s32 c = (c_); const s32 d = (d_); u32 f; s32 Dn, In;                                       \
                if (masking && stamp[(j_) & msk] == row) { f = F_MASK; c = LZB_NEG_INF; Dn = LZB_NEG_INF; In = LZB_NEG_INF; } \
                else if (d > c || I > c) {                                                                 \
                    if (d >= I) { c = d; f = LINK_D | LINK_IEXT | LINK_DEXT; } else { c = I; f = LINK_I | LINK_IEXT | LINK_DEXT; } \
                    In = satadd(I, -gapE); Dn = satadd(d, -gapE);                                          \
                } else {                                                                                   \
                    const s32 open = satadd(c, -gapOE), dx = satadd(d, -gapE), ii = satadd(I, -gapE);      \
                    if (open > dx) { Dn = open; f = 0; } else { Dn = dx; f = LINK_DEXT; }                  \
                    if (open > ii) In = open; else { In = ii; f |= LINK_IEXT; }                            \
                    f |= F_CAND; candMax = max(candMax, c);                                                \
                }                                                                                          \
                Ccur[(j_) & msk] = c; Dv[(j_) & msk] = Dn; flg[(j_) & msk] = (u8)f; I = In; }
            if (cached) {
                if (j0 + 0 < j1) DP_CELL(j0 + 0, dg0, dd0)
                if (j0 + 1 < j1) DP_CELL(j0 + 1, dg1, dd1)
                if (j0 + 2 < j1) DP_CELL(j0 + 2, dg2, dd2)
                if (j0 + 3 < j1) DP_CELL(j0 + 3, dg3, dd3)
                if (j0 + 4 < j1) DP_CELL(j0 + 4, dg4, dd4)
            } else {
                s32 pc = (j0 > LY && j0 < colEnd) ? Cprev[(j0 - 1) & msk] : LZB_NEG_INF;
                for (u32 j = j0; j < j1; j++) {
                    const s32 bi = !rev ? (s32)(a2 + j) : (s32)(a2 + 1 - j);
                    const u8 bc = (bi < 0) ? cls0 : cls2[bi];
                    const s32 diag = (j == LY) ? LZB_NEG_INF : pc + subRow[bc];
                    pc = Cprev[j & msk];
                    DP_CELL(j, diag, Dv[j & msk])
                }
            }
#undef DP_CELL
        }
        /* block-wide exclusive prefix max of the candidates, seeded with bestScore */
        s32 B;
        {
            s32 pm = candMax;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { s32 u = __shfl_up_sync(FULL, pm, o); if ((int)lane >= o) pm = max(pm, u); }
            if (lane == 31) sh->wmax[warp] = pm;
            B = __shfl_up_sync(FULL, pm, 1);
            if (lane == 0) B = LZB_NEG_INF;
            __syncthreads();
#pragma unroll
            for (int w = 0; w < DP_WARPS; w++) if (w < (int)warp) B = max(B, sh->wmax[w]);
            B = max(B, best);
        }
        /* pass 3: prune, band edges, best/end */
        u32 firstAlive = 0xFFFFFFFFu, lastAlive1 = 0;               /* lastAlive1 = last alive column + 1 */
        s32 upVal = -1; u32 upCol1 = 0;                             /* bestScore updates are >= best >= 0 */
        s32 bVal = LZB_NEG_INF; u32 bCol1 = 0;
        for (u32 j = j0; j < j1; j++) {
            const s32 c = Ccur[j & msk]; const u32 f = flg[j & msk];
            const bool alive = !(f & F_MASK) && c >= B - yDrop;
            if (!alive) { Ccur[j & msk] = LZB_NEG_INF; Dv[j & msk] = LZB_NEG_INF; tb[(u32)(tbBase + j)] = 0; continue; }
            tb[(u32)(tbBase + j)] = (u8)(f & 15);
            if (firstAlive == 0xFFFFFFFFu) firstAlive = j;
            lastAlive1 = j + 1;
            if (f & F_CAND) {
                if (c >= B) { B = c; upVal = c; upCol1 = j + 1; }
                if (!trim && (row == M || j == N) && c >= bVal) { bVal = c; bCol1 = j + 1; }
            }
        }
        /* block reductions with the warp-reduce unit; ties go to the later cell (:3742, :3749) */
        {
            const u32 fa = __reduce_min_sync(FULL, firstAlive), la = __reduce_max_sync(FULL, lastAlive1);
            const s32 uv = __reduce_max_sync(FULL, upVal);
            const u32 uc = __reduce_max_sync(FULL, (upVal == uv && upCol1) ? upCol1 : 0u);
            if (lane == 0) { sh->wfa[warp] = fa; sh->wla[warp] = la; sh->wuv[warp] = uv; sh->wuc[warp] = uc; }
            if (!trim) {
                const s32 bv = __reduce_max_sync(FULL, bVal);
                const u32 bc = __reduce_max_sync(FULL, (bVal == bv && bCol1) ? bCol1 : 0u);
                if (lane == 0) { sh->wbv[warp] = bv; sh->wbc[warp] = bc; }
            }
        }
        __syncthreads();
        u32 fa = 0xFFFFFFFFu, la = 0; s32 uv = -1; u32 uc = 0; s32 bv = LZB_NEG_INF; u32 bc = 0;
#pragma unroll
        for (int w = 0; w < DP_WARPS; w++) {
            fa = min(fa, sh->wfa[w]); la = max(la, sh->wla[w]);
            const s32 v = sh->wuv[w]; const u32 cc = sh->wuc[w];
            if (v > uv || (v == uv && cc > uc)) { uv = v; uc = cc; }
            if (!trim) { const s32 v2 = sh->wbv[w]; const u32 c2 = sh->wbc[w]; if (v2 > bv || (v2 == bv && c2 > bc)) { bv = v2; bc = c2; } }
        }
        /* bestScore moves to the LAST cell (row-major) that equalled the row's final best (:3742) */
        u32 bestCol = 0; bool bestMoved = false;
        if (uc) { best = uv; bestCol = uc - 1; bestMoved = true; }
        /* boundaryScore (:3747-3750, only without y-drop trimming) */
        u32 bndCol = 0; bool bndMoved = false;
        if (!trim && bc && bv >= bnd) { bnd = bv; bndCol = bc - 1; bndMoved = true; }
        /* the later event in row-major order owns the end cell; in one cell the boundary test runs second */
        if (bestMoved && (!bndMoved || bestCol > bndCol)) { end1 = row; end2 = bestCol; endIsBnd = 0; }
        else if (bndMoved) { end1 = row; end2 = bndCol; endIsBnd = 1; }
        cells += colEnd - leftCol;
        used += colEnd - leftCol;
        if (dbg && tid == 0 && row < dbgCap) { u32* g = dbg + 4 * (size_t)row; g[0] = leftCol; g[1] = colEnd; g[2] = (u32)best; g[3] = (u32)used; }
        u32 npCol;
        if (la) { LY = fa; npCol = la - 1; } else { LY = colEnd; npCol = leftCol; }
        s32* t = Cprev; Cprev = Ccur; Ccur = t;
        if (LY >= RY) break;
        /* ---- row end, gapped_extend.c:3789-3827 ---- */
        s32 NN = (rightSeg.al >= 0 && R > 0) ? R - 1 : (s32)N;
        u32 wcol = colEnd;
        if (RY > npCol + 1) RY = npCol + 1;
        else {
            s32 thr = best - yDrop; u32 p = 0;
            if (Iout >= thr && (s32)RY <= NN) {
                u32 room = (u32)(NN - (s32)RY) + 1;
                u32 byScore = gapE > 0 ? (u32)((Iout - thr) / gapE) + 1 : room;
                p = byScore < room ? byScore : room;
            }
            if ((s64)(RY + p + 2 - LY) + 8 >= (s64)cap) { status = DP_RING; break; }
            for (u32 q = tid; q < p; q += DP_THREADS) {
                s32 v = Iout - (s32)q * gapE;
                Cprev[(wcol + q) & msk] = v; Dv[(wcol + q) & msk] = v - gapOE;
                tb[(u32)(tbBase + wcol + q)] = LINK_I;
            }
            wcol += p; RY += p; used += p;
        }
        if ((s32)RY <= NN) {
            if (tid == 0) { Cprev[wcol & msk] = LZB_NEG_INF; Dv[wcol & msk] = LZB_NEG_INF; }
            RY++;
        }
        __syncthreads();
    }
    /* ---- traceback, gapped_extend.c:3847-3859: warp 0, 32 diagonal steps per iteration ---- */
    __threadfence();
    __syncthreads();
    if (warp != 0) return;
    u32 nops = 0; bool ovf = false;
    if (status == DP_OK || status == DP_TRUNCATED)
        nops = traceback_walk(tb, tbRow, end1, end2, J->ops, J->opsCap, lane, &ovf);
    if (lane == 0) {
        if (!tbOnly) {
            J->score = endIsBnd ? bnd : best; J->end1 = end1; J->end2 = end2;
            J->rows = row; J->cells = cells; J->status = status; J->ckptCount = 0;
        }
        J->nops = nops; J->opsOverflow = ovf ? 1 : 0;
        job_done(J);
}
}

#endif
