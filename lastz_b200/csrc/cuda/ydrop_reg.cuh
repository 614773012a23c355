/*
 * ydrop_reg.cuh -- K5, register-resident variant of the Y-drop sweep (included by gapped.cu).
 *
 * Same mathematics as k_ydrop (see gapped.cu), different data placement.  k_ydrop keeps the sweep
 * row in shared memory and re-deals columns to threads every row; ncu showed ~700 warp instructions
 * per row per warp, most of them address arithmetic, shared-memory traffic and per-row fixed costs
 * (profiles/r01_k_ydrop_after.txt).  Here every thread OWNS K fixed columns
 *      column(t, s) = base + t*K + s,   t = 0..255, s = 0..K-1
 * and keeps their C and D values and the query-side class codes in registers across rows; only
 * the left neighbour's last column crosses threads (one shared-memory word per thread per row).
 * The band drifts right as the rows advance: when its left edge has passed a whole warp's columns,
 * `base` advances by 32*K and every warp adopts the registers of the warp to its right (a shared
 * memory hand-over every >=32*K rows or so).  Bands wider than 256*K - 32*K columns return DP_RING and
 * the host reruns that alignment with the shared-memory kernel, so nothing is approximated.
 */
#define RG_THREADS 256
#define RG_WARPS 8

template <int K>
struct rg_shared {
    s32 edge[RG_THREADS];                 /* C of each thread's last column after the previous row */
    s32 xchC[RG_THREADS * K], xchD[RG_THREADS * K];
    u8  xchB[RG_THREADS * K];
    s32 wmaxI[RG_WARPS], wmax[RG_WARPS], wuv[RG_WARPS], wbv[RG_WARPS];
    u32 wfa[RG_WARPS], wla[RG_WARPS], wuc[RG_WARPS], wbc[RG_WARPS];
    xf  wagg[RG_WARPS];
    int nact, alignList, status;
};

template <int K>
__global__ void __launch_bounds__(RG_THREADS)
k_ydrop_reg(dp_job* jobs, const dseg* __restrict__ segs,
            const u8* __restrict__ cls1, const u8* __restrict__ cls2, u32 len1, u32 len2,
            const lzb_scoring_dev* __restrict__ sc, s32 yDrop, int trim) {
    constexpr u32 CAP = RG_THREADS * K, SHIFT = 32 * K, msk = CAP - 1;
    __shared__ s32 subC[LZB_MAX_CLASSES * LZB_MAX_CLASSES];
    __shared__ u32 stamp[CAP];
    __shared__ rg_shared<K> shs;
    rg_shared<K>* sh = &shs;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 FULL = 0xFFFFFFFFu;
    dp_job* J = &jobs[blockIdx.x];
    if (J->skip) return;
    const dalign* __restrict__ al = J->al;
    for (u32 i = tid; i < LZB_MAX_CLASSES * LZB_MAX_CLASSES; i += RG_THREADS) subC[i] = sc->subC[i];
    for (u32 i = tid; i < CAP; i += RG_THREADS) stamp[i] = 0;
    const int rev = J->reversed; const u32 a1 = J->a1, a2 = J->a2, M = J->M, N = J->N;
    const s32 gapE = sc->gapExtend, gapOE = sc->gapOpen + sc->gapExtend;
    const u8 cls0 = sc->cls[0];
    u8* tb = J->tb; const s64 tbLen = J->tbLen; u32* tbRow = J->tbRow;
    int status = DP_OK;
    s32 best = 0, bnd = LZB_NEG_INF; u32 end1 = 0, end2 = 0; int endIsBnd = 0;
    unsigned long long cells = 0; u32 row = 0;
    if (N == 0 || M == 0) {
        if (tid == 0) { J->score = 0; J->end1 = J->end2 = 0; J->nops = 0; J->rows = 0; J->cells = 0; J->status = DP_OK; }
        return;
    }
    const s32 yTail = yDrop / gapE + 6;                       /* the host only uses this kernel when gapE > 0 */
    s32 L = J->L0, R = J->R0;
    segref leftSeg = J->leftSeg, rightSeg = J->rightSeg;
    int alignList = J->alignList;
    int* act = J->act; int nact = 0;
    const u32 tbRowCap = J->tbRowCap, actCap = J->actCap;
    u32 lLim = 0, rLim = 0; int lTyp = 0, rTyp = 0;
    LOAD_BOUND(leftSeg, lLim, lTyp); LOAD_BOUND(rightSeg, rLim, rTyp);
    s64 used = 0;
    u32 base = 0;
    s32 C[K], D[K]; u32 Bc[K];
    /* query-side class code of column `col` (B(col), gapped_extend.c:2512-2527) */
#define RG_BCODE(col_) (((col_) <= N) ? (u32)cls2[!rev ? (a2 + (col_)) : (a2 + 1 - (col_))] : (u32)cls0)
    /* ---- first row, gapped_extend.c:3576-3591 ---- */
    u32 LY = 0, RY;
    {
        u32 last = 1;
        if (yDrop >= gapOE) last = (u32)(((s64)yDrop - gapOE) / gapE) + 2;
        if (last > N) last = N;
        if ((s64)last + 1 + yTail + 8 >= (s64)(CAP - SHIFT)) status = DP_RING;
#pragma unroll
        for (int s = 0; s < K; s++) {
            const u32 col = tid * K + s;
            Bc[s] = RG_BCODE(col);
            if (col <= last && status == DP_OK) {
                const s32 v = col == 0 ? 0 : -gapOE - (s32)(col - 1) * gapE;
                C[s] = v; D[s] = v - gapOE;
                tb[col] = col == 0 ? 0 : LINK_I;
            } else { C[s] = LZB_NEG_INF; D[s] = LZB_NEG_INF; }
        }
        used = (s64)last + 1;
        RY = last + 1;
        if (tid == 0 && tbRowCap > 0) tbRow[0] = 0;
        sh->edge[tid] = C[K - 1];
    }
    __syncthreads();
    if (status == DP_OK)
    for (row = 1; row <= M; row++) {
        /* ---- update_LR_bounds gapped_extend.c:4588-4724 ---- */
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
        /* ---- keep the band inside the register window: hand columns over to the left ---- */
        while (LY >= base + SHIFT) {
#pragma unroll
            for (int s = 0; s < K; s++) { sh->xchC[tid * K + s] = C[s]; sh->xchD[tid * K + s] = D[s]; sh->xchB[tid * K + s] = (u8)Bc[s]; }
            const s32 myEdge = sh->edge[tid];
            __syncthreads();
            base += SHIFT;
            const u32 src = tid + 32;
            s32 newEdge = LZB_NEG_INF;
#pragma unroll
            for (int s = 0; s < K; s++) {
                if (src < RG_THREADS) { C[s] = sh->xchC[src * K + s]; D[s] = sh->xchD[src * K + s]; Bc[s] = sh->xchB[src * K + s]; }
                else { C[s] = LZB_NEG_INF; D[s] = LZB_NEG_INF; const u32 col = base + tid * K + s; Bc[s] = RG_BCODE(col); }
            }
            if (src < RG_THREADS) newEdge = sh->edge[src];
            (void)myEdge;
            __syncthreads();
            sh->edge[tid] = newEdge;
            __syncthreads();
        }
        if ((s64)(RY > LY ? RY : LY) + yTail + 8 >= (s64)base + (s64)CAP) { status = DP_RING; break; }
        /* ---- update_active_segs gapped_extend.c:4885-4962 (thread 0; the list is tiny) ---- */
        if (nact > 0 || alignList >= 0) {
            if (tid == 0) {
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
                while (alignList >= 0) {
                    const dalign x = al[alignList];
                    if (!rev) { if (x.pos1 - a1 != row) break; } else { if (a1 - x.end1 != row) break; }
                    if ((u32)nact >= actCap) { status = DP_ACT; break; }
                    int* a = act + 5 * nact; nact++;
                    a[0] = alignList; a[1] = !rev ? 0 : x.segCount - 1;
                    act_build(a, al, segs, rev, stamp, msk, row, a1, a2, LY, RY);
                    alignList = !rev ? x.next : x.prev;
                }
                int w = 0;
                for (int k = 0; k < nact; k++) if (act[5 * k + 4] >= 0) { if (w != k) for (int z = 0; z < 5; z++) act[5 * w + z] = act[5 * k + z]; w++; }
                nact = w;
                sh->nact = nact; sh->alignList = alignList; sh->status = status;
            }
            __syncthreads();
            nact = sh->nact; alignList = sh->alignList; status = sh->status;
            if (status != DP_OK) break;
        }
        /* ---- traceback capacity gapped_extend.c:3636-3662 ---- */
        if (RY < LY) RY = LY;
        const s64 need = (s64)(RY - LY) + yTail;
        if (used + need >= tbLen) { status = DP_TRUNCATED; break; }
        if (row >= tbRowCap) { status = DP_TBROW; break; }
        const u32 tbBase = (u32)((u64)used - (u64)LY);
        if (tid == 0) tbRow[row] = tbBase;
        /* ---- the sweep, gapped_extend.c:3669-3774 ---- */
        const u32 leftCol = LY;
        const u32 colEnd = RY < N + 1 ? RY : N + 1;
        const u32 width = colEnd > LY ? colEnd - LY : 0;
        const u32 cb = base + tid * K;                             /* my first column */
        const s32 ai = !rev ? (s32)(a1 + row) : (s32)(a1 + 1 - row);
        const u8 ac = (ai < 0 || (u32)ai >= len1) ? cls0 : cls1[ai];
        const s32* subRow = subC + ac * LZB_MAX_CLASSES;
        const bool masking = nact > 0;
        /* pass 1: diagonal proposals; my piece of the insertion chain (see k_ydrop for the algebra) */
        s32 dg[K];
        s32 Iin, Iout;
        {
            s32 leftC = tid > 0 ? sh->edge[tid - 1] : LZB_NEG_INF;
            s32 vmax = LZB_NEG_INF;
            xf mine; mine.A = LZB_NEG_INF; mine.S = 0; mine.r = 0;
#pragma unroll
            for (int s = 0; s < K; s++) {
                const u32 col = cb + s;
                const bool inb = col >= LY && col < colEnd;
                const s32 diag = (inb && col != LY) ? leftC + subRow[Bc[s]] : LZB_NEG_INF;
                leftC = C[s];
                dg[s] = diag;
                if (inb) {
                    const s32 a = diag >= D[s] ? satadd(diag, -gapOE) : LZB_NEG_INF;
                    if (!masking) vmax = max(vmax, a + gapE * (s32)(col - LY + 1));
                    else {
                        xf g;
                        if (stamp[col & msk] == row) { g.A = LZB_NEG_INF; g.S = 0; g.r = 1; } else { g.A = a; g.S = -gapE; g.r = 0; }
                        mine = xf_then(mine, g);
                    }
                }
            }
            if (!masking) {
                s32 inc = vmax;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { s32 u = __shfl_up_sync(FULL, inc, o); if ((int)lane >= o) inc = max(inc, u); }
                if (lane == 31) sh->wmaxI[warp] = inc;
                s32 ex = __shfl_up_sync(FULL, inc, 1);
                if (lane == 0) ex = LZB_NEG_INF;
                __syncthreads();
                s32 tot = LZB_NEG_INF;
#pragma unroll
                for (int w = 0; w < RG_WARPS; w++) { const s32 v = sh->wmaxI[w]; if (w < (int)warp) ex = max(ex, v); tot = max(tot, v); }
                /* I entering my first in-band column */
                const u32 firstIn = cb > LY ? cb : LY;
                Iin = satadd(ex, -gapE * (s32)(firstIn - LY));
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
                for (int w = 0; w < RG_WARPS; w++) { xf a = sh->wagg[w]; if (w < (int)warp) pre = xf_then(pre, a); tot = xf_then(tot, a); }
                Iin = xf_then(pre, exl).A;
                Iout = tot.A;
            }
        }
        /* pass 2: cell values, links, next row's D; candidates for bestScore */
        s32 nC[K], nD[K]; u32 fl[K];
        s32 candMax = LZB_NEG_INF;
        {
            s32 I = Iin;
#pragma unroll
            for (int s = 0; s < K; s++) {
                const u32 col = cb + s;
                const bool inb = col >= LY && col < colEnd;
                nC[s] = C[s]; nD[s] = D[s]; fl[s] = 0;
                if (inb) {
                    s32 c = dg[s]; const s32 d = D[s]; u32 f; s32 Dn, In;
                    if (masking && stamp[col & msk] == row) { f = F_MASK; c = LZB_NEG_INF; Dn = LZB_NEG_INF; In = LZB_NEG_INF; }
                    else if (d > c || I > c) {
                        if (d >= I) { c = d; f = LINK_D | LINK_IEXT | LINK_DEXT; } else { c = I; f = LINK_I | LINK_IEXT | LINK_DEXT; }
                        In = satadd(I, -gapE); Dn = satadd(d, -gapE);
                    } else {
                        const s32 open = satadd(c, -gapOE), dx = satadd(d, -gapE), ii = satadd(I, -gapE);
                        if (open > dx) { Dn = open; f = 0; } else { Dn = dx; f = LINK_DEXT; }
                        if (open > ii) In = open; else { In = ii; f |= LINK_IEXT; }
                        f |= F_CAND; candMax = max(candMax, c);
                    }
                    nC[s] = c; nD[s] = Dn; fl[s] = f; I = In;
                }
            }
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
            for (int w = 0; w < RG_WARPS; w++) if (w < (int)warp) B = max(B, sh->wmax[w]);
            B = max(B, best);
        }
        /* pass 3: prune, band edges, best/end */
        u32 firstAlive = 0xFFFFFFFFu, lastAlive1 = 0;
        s32 upVal = -1; u32 upCol1 = 0;
        s32 bVal = LZB_NEG_INF; u32 bCol1 = 0;
#pragma unroll
        for (int s = 0; s < K; s++) {
            const u32 col = cb + s;
            const bool inb = col >= LY && col < colEnd;
            if (inb) {
                const s32 c = nC[s]; const u32 f = fl[s];
                const bool alive = !(f & F_MASK) && c >= B - yDrop;
                if (!alive) { C[s] = LZB_NEG_INF; D[s] = LZB_NEG_INF; tb[(u32)(tbBase + col)] = 0; }
                else {
                    C[s] = c; D[s] = nD[s];
                    tb[(u32)(tbBase + col)] = (u8)(f & 15);
                    if (firstAlive == 0xFFFFFFFFu) firstAlive = col;
                    lastAlive1 = col + 1;
                    if (f & F_CAND) {
                        if (c >= B) { B = c; upVal = c; upCol1 = col + 1; }
                        if (!trim && (row == M || col == N) && c >= bVal) { bVal = c; bCol1 = col + 1; }
                    }
                }
            }
        }
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
        for (int w = 0; w < RG_WARPS; w++) {
            fa = min(fa, sh->wfa[w]); la = max(la, sh->wla[w]);
            const s32 v = sh->wuv[w]; const u32 cc = sh->wuc[w];
            if (v > uv || (v == uv && cc > uc)) { uv = v; uc = cc; }
            if (!trim) { const s32 v2 = sh->wbv[w]; const u32 c2 = sh->wbc[w]; if (v2 > bv || (v2 == bv && c2 > bc)) { bv = v2; bc = c2; } }
        }
        u32 bestCol = 0; bool bestMoved = false;
        if (uc) { best = uv; bestCol = uc - 1; bestMoved = true; }
        u32 bndCol = 0; bool bndMoved = false;
        if (!trim && bc && bv >= bnd) { bnd = bv; bndCol = bc - 1; bndMoved = true; }
        if (bestMoved && (!bndMoved || bestCol > bndCol)) { end1 = row; end2 = bestCol; endIsBnd = 0; }
        else if (bndMoved) { end1 = row; end2 = bndCol; endIsBnd = 1; }
        cells += colEnd - leftCol;
        used += colEnd - leftCol;
        u32 npCol;
        if (la) { LY = fa; npCol = la - 1; } else { LY = colEnd; npCol = leftCol; }
        if (LY >= RY) break;
        /* ---- row end, gapped_extend.c:3789-3827: every thread patches the columns it owns ---- */
        const s32 NN = (rightSeg.al >= 0 && R > 0) ? R - 1 : (s32)N;
        u32 wcol = colEnd, p = 0;
        if (RY > npCol + 1) RY = npCol + 1;
        else {
            const s32 thr = best - yDrop;
            if (Iout >= thr && (s32)RY <= NN) {
                const u32 room = (u32)(NN - (s32)RY) + 1;
                const u32 byScore = (u32)((Iout - thr) / gapE) + 1;
                p = byScore < room ? byScore : room;
            }
        }
        const bool sentinel = (s32)(RY + p) <= NN;
        const u32 sentCol = wcol + p;
        if (p | (u32)sentinel) {
#pragma unroll
            for (int s = 0; s < K; s++) {
                const u32 col = cb + s;
                if (col >= wcol && col < wcol + p) {
                    const s32 v = Iout - (s32)(col - wcol) * gapE;
                    C[s] = v; D[s] = v - gapOE;
                    tb[(u32)(tbBase + col)] = LINK_I;
                }
                if (sentinel && col == sentCol) { C[s] = LZB_NEG_INF; D[s] = LZB_NEG_INF; }
            }
        }
        RY += p; used += p;
        if (sentinel) RY++;
        sh->edge[tid] = C[K - 1];
        __syncthreads();
    }
#undef RG_BCODE
    /* ---- traceback, gapped_extend.c:3847-3859: warp 0, 32 diagonal steps per iteration ---- */
    __threadfence();
    __syncthreads();
    if (warp != 0) return;
    u32 nops = 0;
    if (status == DP_OK || status == DP_TRUNCATED) {
        bool ovf = false;
        nops = traceback_walk(tb, tbRow, end1, end2, J->ops, J->opsCap, lane, &ovf);
        if (ovf) status = DP_OPS;
    }
    if (lane == 0) {
        J->score = endIsBnd ? bnd : best; J->end1 = end1; J->end2 = end2; J->nops = nops;
        J->rows = row; J->cells = cells; J->status = status;
    }
}
