/*
 * ydrop_mw.cuh -- K5b, the register-resident Y-drop sweep on NW warps (included by gapped.cu).
 *
 * Same mathematics as k_ydrop (gapped.cu) and the same register placement as k_ydrop_warp
 * (ydrop_warp.cuh), spread over NW warps so that each of the SM's schedulers owns one:
 *
 *   - thread g = 32*warp + lane owns K consecutive columns; column block b lives in thread
 *     b mod (32*NW).  A WARP whose 32 blocks have all fallen off the left edge of the band
 *     re-initialises itself 32*NW blocks further right, so inside a warp lane order is column
 *     order and only the order of the warps rotates (w0 = the warp that holds LY);
 *   - C, D and the query-side class codes stay in registers; per row the warps meet three times
 *     (insertion chain, in-row bestScore threshold, live range + bestScore), each time through one
 *     16-byte shared-memory slot per warp and one block barrier;
 *   - the left neighbour's last column crosses warps through shared memory in the third exchange.
 *
 * From LY the window always reaches at least 32*(NW-1)*K + 1 columns (769 for K = 8, NW = 4);
 * bands that outgrow it return DP_RING and rerun on the shared-memory kernel.
 */
#define MW_SCAP 1024u                  /* stamp ring, columns */

template <int NW> struct mw_shared {
    s32 subC[LZB_MAX_CLASSES * LZB_MAX_CLASSES];
    u32 stamp[MW_SCAP];
    s32 xI[NW];                        /* exchange 1: each warp's chain aggregate (unmasked rows) */
    xf  xIf[NW];                       /*             ... masked rows */
    s32 xB[NW];                        /* exchange 2: each warp's best candidate */
    s32 edgeC[NW];                     /* C of each warp's last column after the row */
    u32 xfa[NW], xla[NW], xuc[NW], xbc[NW]; s32 xuv[NW], xbv[NW];   /* exchange 3 */
    int nact, alignList, status; u32 nextActRow;
};

struct mw_in {
    u32 lane, warp, LY, colEnd, row, M, N, cb, pWcol, pCnt;
    const s32* subRow; u8* rowPtr;
    s32 gapE, gapOE, yDrop, best, pIout; int trim;
};
struct mw_out { u32 fa, la, uc, bc; s32 uv, bv, Iout; };

/* one row of the sweep, gapped_extend.c:3669-3774; MASKING = some earlier alignment crosses the band */
template <int K, int NW, bool MASKING>
__device__ __forceinline__ void mw_sweep(s32 (&C)[K], s32 (&D)[K], const u32 (&Bq)[(K + 3) / 4], mw_shared<NW>& sh, const mw_in& in, mw_out& out) {
    constexpr u32 WSPAN = 32u * K, smsk = MW_SCAP - 1;
    constexpr int KW = (K + 3) / 4;
    const u32 FULL = 0xFFFFFFFFu, lane = in.lane, warp = in.warp, LY = in.LY, colEnd = in.colEnd, row = in.row, cb = in.cb;
    const u32 M = in.M, N = in.N, pWcol = in.pWcol, pCnt = in.pCnt;
    const s32 gapE = in.gapE, gapOE = in.gapOE, yDrop = in.yDrop, best = in.best, pIout = in.pIout;
    const int trim = in.trim;
    const s32* subRow = in.subRow; u8* const rowPtr = in.rowPtr;
    /* my in-band columns; the band's first column takes no diagonal (:3683) */
    const u32 lo = LY > cb ? min(LY - cb, (u32)K) : 0u;
    const u32 hi = colEnd > cb ? min(colEnd - cb, (u32)K) : 0u;
    const u32 im = hi > lo ? (wg_lowmask(hi) & ~wg_lowmask(lo)) : 0u;
    const u32 dmk = (LY >= cb && LY - cb < (u32)K) ? (im & ~(1u << (LY - cb))) : im;
    u32 mm = 0;                                        /* columns on an earlier alignment (:3690) */
    if (MASKING) {
#pragma unroll
        for (int s = 0; s < K; s++) if (((im >> s) & 1u) && sh.stamp[(cb + s) & smsk] == row) mm |= 1u << s;
    }
    const u32 w0 = (LY / WSPAN) & (NW - 1), lw = (warp - w0) & (NW - 1);      /* my warp's rank in column order */
    /* ---- pass 1: diagonal proposals, insertion-open values a(j), my piece of the chain ---- */
    s32 dg[K], av[K];
    s32 Iin, Iout;
    {
        s32 leftC = __shfl_up_sync(FULL, C[K - 1], 1);
        if (lane == 0) {
            const u32 pc = cb - 1;                     /* a prolonged cell of the previous row never reached the owner's edge slot */
            leftC = sh.edgeC[(warp + NW - 1) & (NW - 1)];
            if (pCnt && pc >= pWcol && pc - pWcol < pCnt) leftC = pIout - (s32)(pc - pWcol) * gapE;
        }
        s32 t[K];
        const s32 ebase = gapE * ((s32)cb - (s32)LY + 1);
        xf mine; mine.A = LZB_NEG_INF; mine.S = 0; mine.r = 0;
#pragma unroll
        for (int s = 0; s < K; s++) {
            const u32 bc = (Bq[s >> 2] >> (8 * (s & 3))) & 255u;
            const s32 diag = ((dmk >> s) & 1u) ? leftC + subRow[bc] : LZB_NEG_INF;
            leftC = C[s];
            const s32 d = ((im >> s) & 1u) ? D[s] : LZB_NEG_INF;
            D[s] = d;
            s32 a = diag >= d ? diag - gapOE : LZB_NEG_INF;
            if (MASKING && ((mm >> s) & 1u)) a = LZB_NEG_INF;
            dg[s] = diag; av[s] = a;
            t[s] = a + ebase + gapE * s;
            if (MASKING) {
                xf g;                                  /* out-of-band columns are the identity: Iout is I at colEnd */
                if ((mm >> s) & 1u) { g.A = LZB_NEG_INF; g.S = 0; g.r = 1; }
                else if ((im >> s) & 1u) { g.A = a; g.S = -gapE; g.r = 0; }
                else { g.A = LZB_NEG_INF; g.S = 0; g.r = 0; }
                mine = xf_then(mine, g);
            }
        }
        if (!MASKING) {
            /* shifted form I'(j) = I(j) + e*(j-LY): the chain is a running max (see k_ydrop) */
#pragma unroll
            for (int st = 1; st < K; st <<= 1)
#pragma unroll
                for (int i = 0; i + st < K; i += 2 * st) t[i] = max(t[i], t[i + st]);
            s32 inc = t[0];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const s32 u = __shfl_up_sync(FULL, inc, o); if (lane >= (u32)o) inc = max(inc, u); }
            if (lane == 31) sh.xI[warp] = inc;
            s32 ex = __shfl_up_sync(FULL, inc, 1);
            if (lane == 0) ex = LZB_NEG_INF;
            __syncthreads();
            s32 tot = LZB_NEG_INF;
#pragma unroll
            for (int w = 0; w < NW; w++) { const s32 v = sh.xI[w]; if (((u32)(w - (int)w0) & (NW - 1)) < lw) ex = max(ex, v); tot = max(tot, v); }
            Iin = ex - gapE * ((s32)cb - (s32)LY);
            Iout = tot - gapE * (s32)(colEnd > LY ? colEnd - LY : 0u);
            if (Iout < LZB_NEG_INF) Iout = LZB_NEG_INF;
        } else {
            xf inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                xf up; up.A = __shfl_up_sync(FULL, inc.A, o); up.S = __shfl_up_sync(FULL, inc.S, o); up.r = __shfl_up_sync(FULL, inc.r, o);
                if (lane >= (u32)o) inc = xf_then(up, inc);
            }
            if (lane == 31) sh.xIf[warp] = inc;
            xf exl; exl.A = __shfl_up_sync(FULL, inc.A, 1); exl.S = __shfl_up_sync(FULL, inc.S, 1); exl.r = __shfl_up_sync(FULL, inc.r, 1);
            if (lane == 0) { exl.A = LZB_NEG_INF; exl.S = 0; exl.r = 0; }
            __syncthreads();
            xf pre; pre.A = LZB_NEG_INF; pre.S = 0; pre.r = 0;
            xf tot = pre;
#pragma unroll
            for (int k = 0; k < NW; k++) {             /* warps in column order */
                const xf a = sh.xIf[(w0 + k) & (NW - 1)];
                if ((u32)k < lw) pre = xf_then(pre, a);
                tot = xf_then(tot, a);
            }
            Iin = xf_then(pre, exl).A;
            Iout = tot.A;
        }
    }
    /* ---- pass 2: cell values, links, next row's D; candidates for bestScore ---- */
    u32 fl[KW];
    s32 cand[K];                                       /* c where the diagonal won inside the band, else -inf */
#pragma unroll
    for (int w = 0; w < KW; w++) fl[w] = 0;
    {
        s32 I = Iin;
#pragma unroll
        for (int s = 0; s < K; s++) {
            const s32 diag = dg[s], d = D[s], a = av[s];
            const s32 m = max(d, I);
            const bool gap = m > diag;                 /* a gap beats the diagonal (ties: diagonal) */
            s32 c = gap ? m : diag;
            const s32 ii = I - gapE, dx = d - gapE;
            s32 In = max(a, ii);                       /* = ii whenever a gap won (a <= I - oe) */
            s32 Dn = gap ? dx : max(a, dx);
            u32 f = gap ? ((d >= I ? LINK_D : LINK_I) | LINK_IEXT | LINK_DEXT)
                        : ((a > dx ? 0u : LINK_DEXT) | (a > ii ? 0u : LINK_IEXT));
            s32 cd = gap ? LZB_NEG_INF : diag;         /* out-of-band columns: diag = -inf already */
            if (MASKING && ((mm >> s) & 1u)) { c = LZB_NEG_INF; Dn = LZB_NEG_INF; In = LZB_NEG_INF; f = 0; cd = LZB_NEG_INF; }
            dg[s] = c; av[s] = Dn; cand[s] = cd; fl[s >> 2] |= f << (8 * (s & 3)); I = In;
        }
    }
    /* exclusive prefix max of the candidates in column order, seeded with bestScore */
    s32 B;
    {
        s32 u_[K];
#pragma unroll
        for (int s = 0; s < K; s++) u_[s] = cand[s];
#pragma unroll
        for (int st = 1; st < K; st <<= 1)
#pragma unroll
            for (int i = 0; i + st < K; i += 2 * st) u_[i] = max(u_[i], u_[i + st]);
        s32 pm = u_[0];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const s32 u = __shfl_up_sync(FULL, pm, o); if (lane >= (u32)o) pm = max(pm, u); }
        if (lane == 31) sh.xB[warp] = pm;
        B = __shfl_up_sync(FULL, pm, 1);
        if (lane == 0) B = LZB_NEG_INF;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < NW; w++) { const s32 v = sh.xB[w]; if (((u32)(w - (int)w0) & (NW - 1)) < lw) B = max(B, v); }
        B = max(B, best);
    }
    /* ---- pass 3: prune against the running bestScore (:3737-3745), band edges, best/end ---- */
    u32 am = 0; u32 upCol1 = 0;
#pragma unroll
    for (int s = 0; s < K; s++) {
        const bool alive = ((im >> s) & 1u) && dg[s] >= B - yDrop;
        if (alive) am |= 1u << s;
        if (cand[s] >= B) upCol1 = cb + s + 1;
        B = max(B, cand[s]);
    }
    const s32 upVal = upCol1 ? B : -1;
    {
        u8* const tp = rowPtr + cb;
#pragma unroll
        for (int s = 0; s < K; s++) {
            const bool alive = (am >> s) & 1u;
            const u32 f = (fl[s >> 2] >> (8 * (s & 3))) & 15u;
            C[s] = alive ? dg[s] : LZB_NEG_INF; D[s] = alive ? av[s] : LZB_NEG_INF;
            if ((im >> s) & 1u) tp[s] = alive ? (u8)f : (u8)0;
        }
    }
    s32 bVal = LZB_NEG_INF; u32 bCol1 = 0;
    if (!trim) {                                       /* boundaryScore :3747-3750 */
#pragma unroll
        for (int s = 0; s < K; s++)
            if (((am >> s) & 1u) && cand[s] > LZB_NEG_INF && (row == M || cb + s == N) && C[s] >= bVal) { bVal = C[s]; bCol1 = cb + s + 1; }
    }
    {
        const u32 fa = __reduce_min_sync(FULL, am ? cb + (u32)__ffs(am) - 1u : 0xFFFFFFFFu);
        const u32 la = __reduce_max_sync(FULL, am ? cb + 32u - (u32)__clz(am) : 0u);
        const s32 uv = __reduce_max_sync(FULL, upVal);
        const u32 uc = __reduce_max_sync(FULL, (upVal == uv && upCol1) ? upCol1 : 0u);
        if (lane == 0) { sh.xfa[warp] = fa; sh.xla[warp] = la; sh.xuv[warp] = uv; sh.xuc[warp] = uc; }
        if (!trim) {
            const s32 bv = __reduce_max_sync(FULL, bVal);
            const u32 bc = __reduce_max_sync(FULL, (bVal == bv && bCol1) ? bCol1 : 0u);
            if (lane == 0) { sh.xbv[warp] = bv; sh.xbc[warp] = bc; }
        }
        if (lane == 31) sh.edgeC[warp] = C[K - 1];
    }
    __syncthreads();
    u32 fa = 0xFFFFFFFFu, la = 0; s32 uv = -1; u32 uc = 0; s32 bv = LZB_NEG_INF; u32 bc = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        fa = min(fa, sh.xfa[w]); la = max(la, sh.xla[w]);
        const s32 v = sh.xuv[w]; const u32 cc = sh.xuc[w];
        if (v > uv || (v == uv && cc > uc)) { uv = v; uc = cc; }
        if (!trim) { const s32 v2 = sh.xbv[w]; const u32 c2 = sh.xbc[w]; if (v2 > bv || (v2 == bv && c2 > bc)) { bv = v2; bc = c2; } }
    }
    out.fa = fa; out.la = la; out.uv = uv; out.uc = uc; out.bv = bv; out.bc = bc; out.Iout = Iout;
}

template <int K, int NW, int MINB = 1>
__global__ void __launch_bounds__(32 * NW, MINB)
k_ydrop_mw(dp_job* jobs, const launch_list ll, const dseg* __restrict__ segs,
           const u8* __restrict__ cls1, const u8* __restrict__ cls2, u32 len1, u32 len2,
           const lzb_scoring_dev* __restrict__ sc, s32 yDrop, int trim) {
    constexpr u32 NT = 32u * NW, WIN = NT * K, WSPAN = 32u * K, smsk = MW_SCAP - 1;
    constexpr int KW = (K + 3) / 4;
    static_assert(WIN <= MW_SCAP && K <= 32 && (NW & (NW - 1)) == 0, "window must fit the stamp ring");
    __shared__ mw_shared<NW> sh;
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, FULL = 0xFFFFFFFFu;
    dp_job* J = &jobs[ll.ix[blockIdx.x]];
    const dalign* __restrict__ al = J->al;
    constexpr u32 CKW = CK_RECORD_WORDS;
    for (u32 i = tid; i < LZB_MAX_CLASSES * LZB_MAX_CLASSES; i += NT) sh.subC[i] = sc->subC[i];
    for (u32 i = tid; i < MW_SCAP; i += NT) sh.stamp[i] = 0;
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
    const s32 yTail = gapE != 0 ? yDrop / gapE + 6 : (N < 500000u ? (s32)N + 1 : 500000);
    s32 L = J->L0, R = J->R0;
    segref leftSeg = J->leftSeg, rightSeg = J->rightSeg;
    const int* const listv = J->listv;
    int alignList = J->alignList;                          /* index into listv */
    u32 nextActRow = list_row(listv, alignList, al, rev, a1);
    int* act = J->act; int nact = 0;
    u32* const ckpt = J->ckpt; const u32 ckptCap = J->ckptCap, ckptEvery = J->ckptEvery;
    u32 ckptCount = 0;
    const int resume = J->resume, tbOnly = J->tbOnly;
    const u32 rowLimit = J->rowLimit ? J->rowLimit : 0xFFFFFFFFu;
    const u32 tbRowCap = J->tbRowCap, actCap = J->actCap;
    u32* const dbg = J->dbg; const u32 dbgCap = J->dbgCap;
    u32 lLim = 0, rLim = 0; int lTyp = 0, rTyp = 0;
    LOAD_BOUND(leftSeg, lLim, lTyp); LOAD_BOUND(rightSeg, rLim, rTyp);
    s64 used = 0;
    s32 C[K], D[K]; u32 Bq[KW];
    u32 cb = tid * K;                                      /* my first column */
    /* query-side class codes of my block (B(col), gapped_extend.c:2512-2527) */
#define MW_LOAD_BLOCK()                                                                                   \
    do {                                                                                                  \
        _Pragma("unroll") for (int w_ = 0; w_ < KW; w_++) Bq[w_] = 0;                                     \
        _Pragma("unroll") for (int s_ = 0; s_ < K; s_++) {                                                \
            const u32 col_ = cb + s_;                                                                     \
            const u32 code_ = (col_ <= N) ? (u32)cls2[!rev ? (a2 + col_) : (a2 + 1 - col_)] : (u32)cls0;  \
            Bq[s_ >> 2] |= code_ << (8 * (s_ & 3));                                                       \
        }                                                                                                 \
    } while (0)
    /* ---- first row, gapped_extend.c:3576-3591 -- or the state a checkpoint saved ---- */
    u32 LY = 0, RY = 0, row0 = 1;
    u32 pWcol = 0, pCnt = 0; s32 pIout = LZB_NEG_INF;       /* the previous row's prolongation (:3801-3816), needed to read a left neighbour's last column */
    if (tbOnly) {
        status = J->status; end1 = J->end1; end2 = J->end2; row = J->rows; ckptCount = J->ckptCount;
#pragma unroll
        for (int s = 0; s < K; s++) { C[s] = LZB_NEG_INF; D[s] = LZB_NEG_INF; }
    } else if (resume >= 0) {
        const u32* rec = ckpt + (size_t)resume * CKW;
        row0 = rec[0] + 1; LY = rec[1]; RY = rec[2]; L = (s32)rec[3]; R = (s32)rec[4];
        leftSeg.al = (int)rec[5]; leftSeg.sg = (int)rec[6]; rightSeg.al = (int)rec[7]; rightSeg.sg = (int)rec[8];
        lLim = rec[9]; rLim = rec[10]; lTyp = (int)rec[11]; rTyp = (int)rec[12]; nact = (int)rec[13];
        used = (s64)((u64)rec[14] | ((u64)rec[15] << 32));
        best = (s32)rec[16]; bnd = (s32)rec[17]; end1 = rec[18]; end2 = rec[19]; endIsBnd = (int)rec[20];
        cells = (u64)rec[21] | ((u64)rec[22] << 32);
        if (tid == 0) for (int k = 0; k < 5 * nact; k++) act[k] = (int)rec[CK_HDR + k];
        /* my columns of the band, wherever the writer kept them: first the block this thread owns at that row */
        cb = tid * K;
        while (cb - lane * K + WSPAN <= LY) cb += WIN;      /* the rule of the row loop: a warp wholly left of the band sits at the right end */
        const u32 c0 = rec[26];
        const u32* tv = rec + CK_HDR + 5 * CK_ACT;
#pragma unroll
        for (int s = 0; s < K; s++) {
            const u32 ix = cb + (u32)s - c0;
            const bool in = cb + (u32)s >= c0 && ix < CK_COLS;
            C[s] = in ? (s32)tv[ix] : LZB_NEG_INF; D[s] = in ? (s32)tv[CK_COLS + ix] : LZB_NEG_INF;
        }
        pWcol = 0; pCnt = 0; pIout = LZB_NEG_INF;             /* the restored C already holds the prolonged cells (edgeC is rebuilt from it below) */
        if ((u64)RY + 2 > (u64)(LY / WSPAN) * WSPAN + WIN) status = DP_RING;      /* written by a kernel with a wider window */
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
            const u32 col = cb + s;
            if (col <= last && status == DP_OK) {
                const s32 v = col == 0 ? 0 : -gapOE - (s32)(col - 1) * gapE;
                C[s] = v; D[s] = v - gapOE;
                tb[col] = col == 0 ? 0 : LINK_I;
            } else { C[s] = LZB_NEG_INF; D[s] = LZB_NEG_INF; }
        }
        used = (s64)last + 1;
        RY = last + 1; cells = RY;                        /* the first row counts too, :3593 */
        if (tid == 0 && tbRowCap > 0) tbRow[0] = 0;
    }
    MW_LOAD_BLOCK();
    /* target-side class codes, 32 rows per load, one chunk ahead (every warp keeps its own copy) */
#define MW_ACODE(r_) ([&]() -> u32 { const u32 rr_ = (r_); if (rr_ > M) return (u32)cls0;                 \
                                      const s64 ai_ = !rev ? (s64)a1 + rr_ : (s64)a1 + 1 - (s64)rr_;      \
                                      return (ai_ < 0 || ai_ >= (s64)len1) ? (u32)cls0 : (u32)cls1[ai_]; }())
    /* rows row0 .. row0+31 in acv when row0 - 1 is not a multiple of 32 (never after a checkpoint), else the 32 rows
     * before row0 in acv and the loop's first iteration shifts; a fresh sweep has row0 = 1 and shifts at row 33 */
    u32 acv = MW_ACODE(((row0 - 1) & ~31u) + (row0 > 1 ? -31 : 1) + lane), acvNext = MW_ACODE(((row0 - 1) & ~31u) + (row0 > 1 ? 1 : 33) + lane);
    if (lane == 31) sh.edgeC[warp] = C[K - 1];
    __syncthreads();
    u32 nextCk = ckptCap ? (ckptCount + 1u) * ckptEvery : 0xFFFFFFFFu;
    if (status == DP_OK && !tbOnly)
    for (row = row0; row <= M; row++) {
        if (row >= rowLimit) { status = DP_PAUSED; break; }
        if ((row & 255u) == 0 && J->abort) { status = DP_ABORTED; break; }   /* the anchor was retired (mapped host memory: looked at rarely) */
        /* ---- update_LR_bounds gapped_extend.c:4588-4724 (every thread, same values) ---- */
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
        /* ---- update_active_segs gapped_extend.c:4885-4962 (thread 0; the list is tiny) ---- */
        if (nact > 0 || row == nextActRow) {
            if (tid == 0) {
                active_update(act, &nact, actCap, listv, &alignList, &nextActRow, &status, al, segs, rev, sh.stamp, smsk, row, a1, a2, LY, RY);
                sh.nact = nact; sh.alignList = alignList; sh.nextActRow = nextActRow; sh.status = status;
            }
            __syncthreads();
            nact = sh.nact; alignList = sh.alignList; nextActRow = sh.nextActRow; status = sh.status;
            if (status != DP_OK) break;
        }
        /* ---- traceback capacity gapped_extend.c:3636-3662 ---- */
        if (RY < LY) RY = LY;
        const s64 need = (s64)(RY - LY) + yTail;
        if (used + need >= tbLen) { status = DP_TRUNCATED; break; }
        if (row >= tbRowCap) { status = DP_TBROW; break; }
        if (tid == 0) tbRow[row] = (u32)((u64)used - (u64)LY);
        u8* const rowPtr = tb + (used - (s64)LY);            /* rowPtr[col] = this row's traceback byte of column col */
        /* ---- the sweep, gapped_extend.c:3669-3774 ---- */
        const u32 leftCol = LY;
        const u32 colEnd = RY < N + 1 ? RY : N + 1;
        if (((row - 1) & 31u) == 0 && row > 1) { acv = acvNext; acvNext = MW_ACODE(row + 32 + lane); }
        const u32 ac = __shfl_sync(FULL, acv, (row - 1) & 31u);
        const s32* subRow = sh.subC + ac * LZB_MAX_CLASSES;
        mw_in in;
        in.lane = lane; in.warp = warp; in.LY = LY; in.colEnd = colEnd; in.row = row; in.M = M; in.N = N; in.cb = cb;
        in.pWcol = pWcol; in.pCnt = pCnt; in.pIout = pIout; in.subRow = subRow; in.rowPtr = rowPtr;
        in.gapE = gapE; in.gapOE = gapOE; in.yDrop = yDrop; in.best = best; in.trim = trim;
        mw_out o;
        if (nact > 0) mw_sweep<K, NW, true>(C, D, Bq, sh, in, o); else mw_sweep<K, NW, false>(C, D, Bq, sh, in, o);
        const u32 fa = o.fa, la = o.la, uc = o.uc, bc = o.bc; const s32 uv = o.uv, bv = o.bv, Iout = o.Iout;
        /* bestScore moves to the LAST cell (row-major) that equalled the row's final best (:3742) */
        u32 bestCol = 0; bool bestMoved = false;
        if (uc) { best = uv; bestCol = uc - 1; bestMoved = true; }
        u32 bndCol = 0; bool bndMoved = false;
        if (!trim && bc && bv >= bnd) { bnd = bv; bndCol = bc - 1; bndMoved = true; }
        if (bestMoved && (!bndMoved || bestCol > bndCol)) { end1 = row; end2 = bestCol; endIsBnd = 0; }
        else if (bndMoved) { end1 = row; end2 = bndCol; endIsBnd = 1; }
        cells += colEnd - leftCol;
        used += colEnd - leftCol;
        if (dbg && tid == 0 && row < dbgCap) { u32* g = dbg + 4 * (size_t)row; g[0] = leftCol; g[1] = colEnd; g[2] = (u32)best; g[3] = (u32)used; }
        u32 npCol;
        if (la) { LY = fa; npCol = la - 1; } else { LY = colEnd; npCol = leftCol; }
        if (LY >= RY) break;
        /* ---- row end, gapped_extend.c:3789-3827 ---- */
        const s32 NN = (rightSeg.al >= 0 && R > 0) ? R - 1 : (s32)N;
        const u32 wcol = colEnd; u32 p = 0;
        if (RY > npCol + 1) RY = npCol + 1;
        else {
            const s32 thr = best - yDrop;
            if (Iout >= thr && (s32)RY <= NN) {
                const u32 room = (u32)(NN - (s32)RY) + 1;
                const u32 byScore = gapE > 0 ? (u32)((Iout - thr) / gapE) + 1 : room;
                p = byScore < room ? byScore : room;
            }
        }
        /* the window follows the band: [first column of the warp holding LY, + 32*NW blocks) must hold the next row */
        if ((u64)(RY > wcol ? RY : wcol) + p + 2 > (u64)(LY / WSPAN) * WSPAN + WIN) { status = DP_RING; break; }
        while (cb - lane * K + WSPAN <= LY) {                /* my whole warp is left of the band: move it to the right end */
            cb += WIN;
#pragma unroll
            for (int s = 0; s < K; s++) { C[s] = LZB_NEG_INF; D[s] = LZB_NEG_INF; }
            MW_LOAD_BLOCK();
        }
        pWcol = wcol; pCnt = p; pIout = Iout;
        if (p) {
            /* prolong the row with insertions (:3801-3816): each thread patches the columns it owns */
            const u32 plo = wcol > cb ? min(wcol - cb, (u32)K) : 0u;
            const u32 phi = wcol + p > cb ? min(wcol + p - cb, (u32)K) : 0u;
            const u32 pmk = phi > plo ? (wg_lowmask(phi) & ~wg_lowmask(plo)) : 0u;
            if (pmk) {
                s32 v = Iout - ((s32)cb - (s32)wcol) * gapE;
                u8* const prow = rowPtr + cb;
#pragma unroll
                for (int s = 0; s < K; s++) {
                    if ((pmk >> s) & 1u) { C[s] = v; D[s] = v - gapOE; prow[s] = LINK_I; }
                    v -= gapE;
                }
            }
            RY += p; used += p;
        }
        if ((s32)RY <= NN) RY++;                             /* the sentinel column already holds LZB_NEG_INF */
        /* ---- checkpoint: everything the next row reads ---- */
        if (row == nextCk && ckptCount < ckptCap && nact <= CK_ACT) {     /* record k belongs to row (k+1)*ckptEvery; a record that cannot be written ends the series */
            u32* rec = ckpt + (size_t)ckptCount * CKW;
            if (tid == 0) {
                rec[0] = row; rec[1] = LY; rec[2] = RY; rec[3] = (u32)L; rec[4] = (u32)R;
                rec[5] = (u32)leftSeg.al; rec[6] = (u32)leftSeg.sg; rec[7] = (u32)rightSeg.al; rec[8] = (u32)rightSeg.sg;
                rec[9] = lLim; rec[10] = rLim; rec[11] = (u32)lTyp; rec[12] = (u32)rTyp; rec[13] = (u32)nact;
                rec[14] = (u32)(u64)used; rec[15] = (u32)((u64)used >> 32);
                rec[16] = (u32)best; rec[17] = (u32)bnd; rec[18] = end1; rec[19] = end2; rec[20] = (u32)endIsBnd;
                rec[21] = (u32)cells; rec[22] = (u32)(cells >> 32);
                rec[26] = LY & ~31u;
                for (int k = 0; k < 5 * nact; k++) rec[CK_HDR + k] = (u32)act[k];
                J->progUsed = (u32)(u64)used; J->progRows = row;      /* lets the host estimate where the traceback will run out */
            }
            const u32 c0 = LY & ~31u;                        /* base column of the record */
            u32* tv = rec + CK_HDR + 5 * CK_ACT;
            for (u32 i = tid; i < 2 * CK_COLS; i += NT) tv[i] = (u32)LZB_NEG_INF;
            __syncthreads();
#pragma unroll
            for (int s = 0; s < K; s++) {
                const u32 ix = cb + (u32)s - c0;
                if (cb + (u32)s >= c0 && ix < CK_COLS) { tv[ix] = (u32)C[s]; tv[CK_COLS + ix] = (u32)D[s]; }
            }
            ckptCount++; nextCk += ckptEvery;
        }
    }
#undef MW_LOAD_BLOCK
#undef MW_ACODE
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
            J->rows = row; J->cells = cells; J->status = status; J->ckptCount = ckptCount;
        }
        J->nops = nops; J->opsOverflow = ovf ? 1 : 0;
        job_done(J);
    }
}
