/*
 * lzb_oracle.c -- CPU restatement of LASTZ's seed-and-extend hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker: tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load liblzb_oracle.so; nothing under lastz_b200/ may.  It
 * implements include/lastz_b200.h on the host so the CUDA library can be compared call by call.
 *
 * Parity is PINNED: oracle/lastz_oracle (the host front end linked against this library)
 * reproduces the reference's golden files test_data/base_test.{default,hits,hsp,seeded,
 * extended,...} byte for byte, and tests/ diff it against oracle/_ref/lastz (the unmodified
 * reference compiled from /root/reference/src by oracle/build_ref.sh) on synthetic inputs.
 *
 * Every function cites the reference lines (lastz 1.04.58, paths relative to the reference
 * tree) whose behaviour it restates.  Data structures are our own (CSR index instead of
 * last/prev chains, index-based alignment records, ring-buffered sweep rows); only the
 * observable semantics follow the reference.
 */
#include "../include/lastz_b200.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef uint8_t  u8;
typedef uint32_t u32;
typedef int32_t  s32;
typedef uint64_t u64;
typedef int64_t  s64;

_Static_assert(sizeof(lzb_segment) == 48, "segment.h:46-64 layout");
_Static_assert(offsetof(lzb_segment, s) == 20 && offsetof(lzb_segment, scoreCov) == 32 &&
               offsetof(lzb_segment, filter) == 40, "segment.h:46-64 layout");
_Static_assert(sizeof(lzb_alignel) == 64 && offsetof(lzb_alignel, s) == 28 &&
               offsetof(lzb_alignel, script) == 48, "edit_script.h:30-41 layout");
_Static_assert(offsetof(lzb_editscript, op) == 12, "edit_script.h:55-61 layout");

/* dna_utilities.h:130-139 */
#define NEG_INF   ((s32)-1932735283)
#define WORST_SCORE ((s32)(-0x7FFFFFFF - 1))

static char g_err[512];
static int fail(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
    return -1;
}
const char* lzb_last_error(void) { return g_err; }
const char* lzb_backend(void) { return "oracle-cpu"; }

struct lzb_ctx {
    s32* sub;        /* scoring->sub, 256x256 */
    s32* msub;       /* maskedScoring->sub */
    s32  gapOpen, gapExtend;
};
struct lzb_target {
    u8* v; u32 len;              /* NUL terminated copy */
    u32 start, end, step;
    int wordBits;
    u32* off;                    /* CSR offsets, 2^wordBits+1 */
    u32* pos;                    /* positions, per word in DEcreasing order */
    u64 npos;
};
struct lzb_query { u8* v; u32 len; };

lzb_ctx* lzb_open(int device) {
    (void)device;
    lzb_ctx* c = calloc(1, sizeof *c);
    return c;
}
void lzb_close(lzb_ctx* c) { if (c) { free(c->sub); free(c->msub); free(c); } }
void lzb_free(void* p) { free(p); }
uint64_t lzb_launch_count(lzb_ctx* c) { (void)c; return 0; }

int lzb_set_scoring(lzb_ctx* c, const int32_t* sub, const int32_t* msub, int32_t go, int32_t ge) {
    free(c->sub); free(c->msub);
    c->sub = malloc(65536 * 4); c->msub = malloc(65536 * 4);
    memcpy(c->sub, sub, 65536 * 4); memcpy(c->msub, msub, 65536 * 4);
    c->gapOpen = go; c->gapExtend = ge;
    return 0;
}

/* apply_seed, seeds.c:1335-1378 (non-complementing seeds) */
static inline u32 pack_word(const lzb_seed* sd, u64 w) {
    u32 p = 0;
    for (int i = 0; i < sd->numParts; i++) p |= (u32)(w >> sd->shift[i]) & sd->mask[i];
    return p;
}

/*
 * Index.  record_seed_positions pos_table.c:396-476 + add_word :1326: every window of
 * seed->length valid characters inside [start,end) whose END position p (index of the base after
 * the window, pos_table.h:72-77) is a multiple of step is recorded under its packed word; a word's
 * positions are later visited largest-first (prepend at :1341-1344, walk at seed_search.c:832).
 * The step>length skip-ahead (:419-424,:466-470) only avoids examining bases outside recorded
 * windows, so the recorded set is exactly "all-valid windows ending on a step multiple".
 */
lzb_target* lzb_target_build(lzb_ctx* c, const uint8_t* seq1, uint32_t len1, uint32_t start,
                             uint32_t end, const int8_t ctb[256], const lzb_seed* sd, uint32_t step) {
    (void)c;
    if (step < 1) { fail("in build_seed_position_table(), step can't be %u", step); return NULL; }
    if (end == 0) end = len1;
    if (end <= start || end > len1) { fail("in build_seed_position_table(), interval is bad"); return NULL; }
    if (sd->weight > 28) { fail("new_position_table can't support >28 seed bits (%d requested)", sd->weight); return NULL; }
    lzb_target* t = calloc(1, sizeof *t);
    t->v = malloc((size_t)len1 + 1); memcpy(t->v, seq1, len1); t->v[len1] = 0;
    t->len = len1; t->start = start; t->end = end; t->step = step; t->wordBits = sd->weight;
    u64 nw = 1ull << sd->weight;
    t->off = calloc(nw + 1, 4);
    int L = sd->length;
    t->pos = malloc(4);
    if (len1 >= (u32)L) {
        /* pass 0 counts, pass 1 fills; each word's slot is filled from its end backwards while
         * scanning forwards, which leaves the positions largest-first */
        u32* cur = NULL;
        for (int pass = 0; pass < 2; pass++) {
            u64 w = 0; int run = 0;
            for (u32 i = start; i < end; i++) {
                int b = ctb[t->v[i]];
                if (b < 0) { run = 0; w = 0; continue; }
                w = (w << 2) | (u64)b; run++;
                u32 p = i + 1;
                if (run < L || p % step != 0) continue;
                u32 word = pack_word(sd, w);
                if (pass == 0) t->off[word + 1]++; else t->pos[--cur[word]] = p;
            }
            if (pass == 0) {
                for (u64 k = 0; k < nw; k++) t->off[k + 1] += t->off[k];
                t->npos = t->off[nw];
                free(t->pos); t->pos = malloc((t->npos ? t->npos : 1) * 4);
                cur = malloc(nw * 4); memcpy(cur, t->off + 1, nw * 4);
            }
        }
        free(cur);
    }
    return t;
}
void lzb_target_free(lzb_target* t) { if (t) { free(t->v); free(t->off); free(t->pos); free(t); } }

int64_t lzb_target_export_index(lzb_target* t, uint32_t* counts, uint32_t* positions) {
    u64 nw = 1ull << t->wordBits;
    for (u64 k = 0; k < nw; k++) counts[k] = t->off[k + 1] - t->off[k];
    if (positions) memcpy(positions, t->pos, t->npos * 4);
    return (int64_t)t->npos;
}

/* limit_position_table pos_table.c:1763-1920 with maxChasm == 0: the position lists of words that occur more than
 * `limit` times are emptied (:1896-1915: last[w] = 0) */
int lzb_target_limit(lzb_target* t, uint32_t limit) {
    u64 nw = 1ull << t->wordBits, w = 0, kept = 0;
    for (w = 0; w < nw; w++) {
        u32 a = t->off[w], b = t->off[w + 1];
        t->off[w] = (u32)kept;
        if (b - a <= limit) { memmove(t->pos + kept, t->pos + a, (size_t)(b - a) * 4); kept += b - a; }
    }
    t->off[nw] = (u32)kept; t->npos = kept;
    return 0;
}

lzb_query* lzb_query_load(lzb_ctx* c, const uint8_t* seq2, uint32_t len2) {
    (void)c;
    lzb_query* q = calloc(1, sizeof *q);
    q->v = malloc((size_t)len2 + 1); memcpy(q->v, seq2, len2); q->v[len2] = 0; q->len = len2;
    return q;
}
void lzb_query_free(lzb_query* q) { if (q) { free(q->v); free(q); } }

/* ------------------------------------------------------------------------------------------
 * seed stage
 * ---------------------------------------------------------------------------------------- */

typedef struct {
    lzb_ctx* c; lzb_target* t; lzb_query* q;
    const lzb_seed* sd; const lzb_seed_params* p;
    u32* E;  u32 hmask;                 /* diagEnd[], diag_hash.h:68 */
    s32* A;                             /* diagActual[], diag_hash.h:70 (recoverable processor only) */
    /* the seed hit queue of the twin processor (diag_hash.h:104-160, diag_hash.c:276-322): a ring of the last Qsize
     * enqueued hits/blocks, chained per hash bucket through prevHit; entry numbers start at Qsize */
    struct twq { u32 pos2; s32 diag; u32 prev; int isBlock; }* Q; u32 Qsize, Qnum; u32* Qlast;
    lzb_segment* out; u64 n, cap;
    lzb_seed_stats st;
} search;

/* compute_entropy, dna_utilities.c:2892-2940 (lowerOk == false) */
static double hsp_entropy(const u8* s, const u8* t, int len) {
    int cA = 0, cC = 0, cG = 0, cT = 0;
    for (int i = 0; i < len; i++) if (s[i] == t[i]) {
        switch (s[i]) { case 'A': cA++; break; case 'C': cC++; break; case 'G': cG++; break; case 'T': cT++; break; default: break; }
    }
    if (cA + cC + cG + cT < 20) return 1.0;
    double pA = (double)cA / (double)len, pC = (double)cC / (double)len;
    double pG = (double)cG / (double)len, pT = (double)cT / (double)len;
    double qA = cA ? log(pA) : 0.0, qC = cC ? log(pC) : 0.0;
    double qG = cG ? log(pG) : 0.0, qT = cT ? log(pT) : 0.0;
    return -(pA * qA + pC * qC + pG * qG + pT * qT) / log(4.0);
}

static void emit_hsp(search* S, u32 end1, u32 end2, u32 len, s32 score) {
    if (S->n == S->cap) { S->cap = S->cap * 2 + 1024; S->out = realloc(S->out, S->cap * sizeof(lzb_segment)); }
    lzb_segment* g = &S->out[S->n++];
    memset(g, 0, sizeof *g);
    g->pos1 = end1 - len; g->pos2 = end2 - len; g->length = len; g->s = score;
    g->id = S->p->strandId; g->scoreCov = len;
    S->st.hsps++;
}

/* nuc_to_bits dna_utilities.c:56-74: the case-insensitive table the exact/mismatch extensions compare with
 * (params->charToBits, lastz.c:353, :2902) */
static inline int n2b(u8 c) {
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return -1; }
}
static inline int mism(u8 a, u8 b) { int x = n2b(a), y = n2b(b); return x != y || x < 0 || y < 0; }

/* match_extend_seed_hit seed_search.c:3018-3254 (--exact=N); positions are indices, s64 where the
 * reference steps a pointer one below the sequence start */
static void exact_hit(search* S, u32 pos1, u32 pos2, u32 h) {
    const u8* v1 = S->t->v; const u8* v2 = S->q->v;
    u32 len = (u32)S->sd->length; s64 diag = (s64)pos1 - (s64)pos2;
    u32 extent;
    for (u32 k = 1; k <= len; k++)                                  /* :3076-3090 the hit itself must match */
        if (mism(v1[pos1 - k], v2[pos2 - k])) { extent = pos2 - k; goto not_a_match; }
    {
        s64 s1 = (s64)pos1 - len, s2 = (s64)pos2 - len;             /* :3096-3140 left */
        s64 blk = (s64)S->E[h] + diag, stop = blk > 0 ? blk : 0;
        if (s1 < stop) { s1--; s2--; }
        else while (s1 >= stop) {
            if (s1 == stop) { s1--; s2--; break; }
            u8 n1 = v1[--s1], n2 = v2[--s2];
            if (n1 == 0 || n2 == 0 || mism(n1, n2)) break;
        }
        s64 left = s1;
        s1 = (s64)pos1 - 1; s2 = (s64)pos2 - 1;                     /* :3146-3176 right */
        s64 lim = (s64)S->q->len + diag, rstop = ((s64)S->t->len <= lim) ? (s64)S->t->len : lim;
        while (s1 < rstop) {
            u8 n1 = v1[++s1], n2 = v2[++s2];
            if (n1 == 0 || n2 == 0 || mism(n1, n2)) break;
        }
        s64 right = s1;
        S->st.extensions++;
        extent = (u32)(right - diag);                               /* :3182-3200 */
        if (extent > S->E[h]) S->E[h] = extent;
        u32 length = (u32)(right - (left + 1));
        if (length < (u32)S->p->hspThreshold) return;
        emit_hsp(S, (u32)right, (u32)(right - diag), length, (s32)length);
        return;
    }
not_a_match:
    if (extent > S->E[h]) S->E[h] = extent;                         /* :3232-3250 */
}

/* mismatch_extend_seed_hit seed_search.c:3450-3778 (--mismatch=M,N) */
static void mismatch_hit(search* S, u32 pos1, u32 pos2, u32 h) {
    const u8* v1 = S->t->v; const u8* v2 = S->q->v;
    u32 len = (u32)S->sd->length; s64 diag = (s64)pos1 - (s64)pos2;
    int M = S->p->gfMismatches, Ecnt = 0;
    const u32 INACTIVE = 0xFFFFFFFFu;                               /* hashInactiveEnd diag_hash.h:101 */
    u32 extent = INACTIVE;
    s64 mmLoc[LZB_GFEX_MISMATCH_MAX + 1];
    for (u32 k = 1; k <= len; k++)                                  /* :3520-3538 mismatches inside the hit */
        if (mism(v1[pos1 - k], v2[pos2 - k])) { extent = pos2 - k; if (++Ecnt > M) goto not_a_match; }
    {
        s64 s1 = (s64)pos1 - len, s2 = (s64)pos2 - len;             /* :3548-3612 left, collecting mismatch positions */
        s64 blk = (s64)S->E[h] + diag, stop = blk > 0 ? blk : 0;
        int mmScan = M + 1 - Ecnt; const int mmStop = mmScan;
        if (s1 < stop) { s1--; s2--; }
        else while (s1 >= stop) {
            if (s1 == stop) { s1--; s2--; break; }
            u8 n1 = v1[--s1], n2 = v2[--s2];
            if (n1 == 0 || n2 == 0) break;
            if (mism(n1, n2)) { mmLoc[--mmScan] = s1; if (mmScan == 0) break; }
        }
        if (mmScan > 0) mmLoc[--mmScan] = s1;                       /* :3618-3621 */
        int mmShortfall = mmScan;
        s1 = (s64)pos1 - 1; s2 = (s64)pos2 - 1;                     /* :3630-3700 right */
        s64 lim = (s64)S->q->len + diag, rstop = ((s64)S->t->len <= lim) ? (s64)S->t->len : lim;
        s64 bestLength = 0, left = -2, right = -2; int have = 0;
        while (s1 < rstop) {
            u8 n1 = v1[++s1], n2 = v2[++s2];
            if (n1 == 0 || n2 == 0) break;
            if (mism(n1, n2)) {
                if (extent == INACTIVE) extent = (u32)s2;
                if (mmShortfall > 0) { mmShortfall--; continue; }
                s64 thisLength = s1 - mmLoc[mmScan];
                if (thisLength > bestLength) { bestLength = thisLength; left = mmLoc[mmScan]; right = s1; have = 1; }
                if (++mmScan == mmStop) break;
            }
        }
        if (mmScan < mmStop) {                                      /* :3706-3720 the stopping point is an endpoint too */
            if (extent == INACTIVE) extent = (u32)s2;
            s64 thisLength = s1 - mmLoc[mmScan];
            if (thisLength > bestLength) { left = mmLoc[mmScan]; right = s1; have = 1; }
        }
        if (!have) return;                                          /* the reference dies here (":3723 found no interval") */
        S->st.extensions++;
        u32 length = (u32)(right - (left + 1));                     /* :3730-3745 */
        if (length >= (u32)S->p->hspThreshold) extent = (u32)(right + 1 - diag);
        if (extent > S->E[h]) S->E[h] = extent;
        if (length < (u32)S->p->hspThreshold) return;
        emit_hsp(S, (u32)right, (u32)(right - diag), length, (s32)length);
        return;
    }
not_a_match:
    if (extent > S->E[h]) S->E[h] = extent;
}

/* _enqueue_seed_hit diag_hash.c:276-322 */
static void twin_enqueue(search* S, u32 pos1, u32 pos2, int isBlock) {
    s32 diag = (s32)(pos1 - pos2);
    u32 h = (u32)diag & S->hmask;
    S->Qnum++;
    struct twq* q = &S->Q[S->Qnum % S->Qsize];
    q->prev = (S->Qlast[h] <= S->Qnum - S->Qsize) ? 0 : S->Qlast[h];       /* a stale predecessor is no longer in the queue */
    S->Qlast[h] = S->Qnum;
    q->isBlock = isBlock; q->pos2 = pos2; q->diag = diag;
}

static int xdrop_from_hit(search* S, u32 pos1, u32 pos2, u32 h, int recover);

/* process_for_twin_hit seed_search.c:1814-2046 (seed hit queue version), x-drop or no extension */
static void twin_hit(search* S, u32 pos1, u32 pos2) {
    const lzb_seed_params* P = S->p;
    const u32 len = (u32)S->sd->length, minSpan = (u32)P->twinMinSpan, maxSpan = (u32)P->twinMaxSpan;
    s32 diag = (s32)(pos1 - pos2);
    u32 h = (u32)diag & S->hmask;
    /* :1873-1892: the first hit of a bucket is only queued.  (An untouched bucket has an empty chain, so the scan below
     * finds nothing and queues the hit just the same.) */
    for (u32 num = S->Qlast[h]; num > S->Qnum - S->Qsize; ) {                  /* :1904-1947 */
        const struct twq* q = &S->Q[num % S->Qsize];
        num = q->prev;
        const u32 span = pos2 - (q->pos2 - len);
        if (span > maxSpan) break;
        if (q->diag != diag) continue;
        if (q->isBlock) { if (pos2 - len <= q->pos2) return; break; }          /* inside / right of an earlier extension */
        if (span < minSpan) continue;
        /* twin_hit :1969-2026 */
        if (P->gfExtend == LZB_GFEX_NONE) {
            S->E[h] = pos2;
            twin_enqueue(S, pos1, pos2, 1);
            emit_hsp(S, pos1, pos2, span, 0);
            return;
        }
        const u32 old = S->E[h];
        xdrop_from_hit(S, pos1, pos2, h, 0);
        if (S->E[h] != old) { const u32 extent = S->E[h]; twin_enqueue(S, (u32)((s64)diag + extent), extent, 1); }
        return;
    }
    twin_enqueue(S, pos1, pos2, 0);                                            /* no twin yet :1957 */
}

/*
 * One seed hit: process_for_simple_hit seed_search.c:1056-1192 followed by
 * xdrop_extend_seed_hit :2528-2959.  pos1/pos2 = one past the hit end.
 */
static void one_hit(search* S, u32 pos1, u32 pos2) {
    const lzb_seed_params* P = S->p;
    u32 len = (u32)S->sd->length;
    S->st.rawSeedHits++;
    if (P->twinMinSpan > 0) { twin_hit(S, pos1, pos2); return; }     /* the twin processor goes before every other (lastz.c:2787-2805) */
    if (P->plainHits) { emit_hsp(S, pos1, pos2, len, 0); return; }   /* :995-1030 */
    s32 diag = (s32)(pos1 - pos2);
    u32 h = (u32)diag & S->hmask;
    const int recover = P->recoverSeeds;
    if (!recover) {
        /* :1097-1113 -- inactive buckets read as 0; discard if an earlier extension on this
         * hash-equivalent diagonal already passed the hit's start */
        if (S->E[h] > pos2 - len) return;
    } else {
        /* process_for_recoverable_hit :1283-1373.  An untouched bucket reads as extent 0 here, and extent 0 lets every
         * hit through both tests below, like the reference's explicit "inactive => fresh hit" (:1285-1290); a touched
         * bucket never has extent 0.  A hit on a different actual diagonal is accepted as fresh (:1298-1336); one on the
         * same diagonal that starts inside the recorded extent is dropped, moving the extent up to its end (:1341-1360). */
        if (diag == S->A[h] && pos2 - len < S->E[h]) { if (pos2 > S->E[h]) S->E[h] = pos2; return; }
        S->A[h] = diag;                                                /* fresh_hit :1373 */
    }
    if (P->gfExtend == LZB_GFEX_NONE) {            /* :1163-1178, :1417-1421 */
        if (!recover || pos2 > S->E[h]) S->E[h] = pos2;
        emit_hsp(S, pos1, pos2, len, 0);
        return;
    }
    if (P->gfExtend == LZB_GFEX_EXACT) { exact_hit(S, pos1, pos2, h); return; }          /* :1146-1151 */
    if (P->gfExtend == LZB_GFEX_MISMATCH) { mismatch_hit(S, pos1, pos2, h); return; }    /* :1158-1164 */
    xdrop_from_hit(S, pos1, pos2, h, recover);
}

/* xdrop_extend_seed_hit seed_search.c:2528-2959 for the hit that ends at (pos1, pos2); returns 1 if an HSP came out */
static int xdrop_from_hit(search* S, u32 pos1, u32 pos2, u32 h, int recover) {
    const lzb_seed_params* P = S->p;
    u32 len = (u32)S->sd->length;
    s32 diag = (s32)(pos1 - pos2);
    S->st.extensions++;
    const u8* v1 = S->t->v; const u8* v2 = S->q->v;
    const s32* sub = S->c->msub;
    s32 xDrop = P->xDrop;
    /* left scan :2598-2632; stop = max(0, diagEnd + diag) in seq1 coordinates; the recoverable processor extends
     * unblocked (oldDiagEnd = 0, :2612) */
    s64 blk = (s64)(recover ? 0u : S->E[h]) + diag;
    u32 stop = blk > 0 ? (u32)blk : 0;
    u32 a = pos1, b = pos2, leftStart = pos1;
    s32 run = 0, leftScore = 0;
    while (a > stop && run >= leftScore - xDrop) {
        --a; --b;
        run += sub[(u32)v1[a] * 256 + v2[b]];
        if (run > leftScore) { leftStart = a; leftScore = run; }
    }
    u32 leftScanned = a;
    u32 hitLeft = pos1 - len;                       /* :2637-2639 */
    if (leftStart > hitLeft) len -= leftStart - hitLeft;
    /* right scan :2663-2693; bounded by the end of either sequence */
    s64 lim = (s64)S->q->len + diag;
    u32 rstop = ((s64)S->t->len <= lim) ? S->t->len : (u32)lim;
    a = pos1; b = pos2;
    u32 rightStop = pos1; s32 rightScore = 0; run = 0;
    while (a < rstop && run >= rightScore - xDrop) {
        run += sub[(u32)v1[a] * 256 + v2[b]];
        a++; b++;
        if (run > rightScore) { rightStop = a; rightScore = run; }
    }
    u32 rightBlock = a;
    S->st.bpExtended += rightBlock - leftScanned;
    /* :2785-2789 -- the bucket remembers where the right SCAN stopped (and on which diagonal) */
    u32 extent = (u32)((s64)rightBlock - diag);
    if (extent > S->E[h]) { S->E[h] = extent; if (recover) S->A[h] = diag; }
    s32 sim = leftScore + rightScore;
    u32 e1 = rightStop, e2 = (u32)((s64)e1 - diag), hl = rightStop - leftStart;
    (void)len;
    /* :2851-2874 entropy adjustment for scores in [K, 3K] */
    if (P->entropy && sim >= P->hspThreshold && sim <= 3 * P->hspThreshold) {
        double q = hsp_entropy(v1 + e1 - hl, v2 + e2 - hl, (int)hl);
        sim = (s32)((double)sim * q);
    }
    if (sim < P->hspThreshold) return 0;            /* :2907 */
    emit_hsp(S, e1, e2, hl, sim);
    return 1;
}

/* find_table_matches seed_search.c:810-875 (+ seed_hit_below_diagonal :2182-2237, unpartitioned) */
static void probe(search* S, u32 word, u32 pos2) {
    const lzb_target* t = S->t;
    for (u32 k = t->off[word]; k < t->off[word + 1]; k++) {
        u32 pos1 = t->pos[k];
        if (S->p->selfCompare) {
            if (S->p->sameStrand) { if (pos1 >= pos2) continue; }
            else {
                u32 a = pos1 - (u32)S->sd->length, b = pos2 - (u32)S->sd->length;
                b = (S->q->len - 1) - b;
                if (a >= b) continue;
            }
        }
        one_hit(S, pos1, pos2);
    }
}

/* seed_hit_search seed_search.c:322-448 + private_hit_search :464-574 */
int lzb_seed_hit_search(lzb_ctx* c, lzb_target* t, lzb_query* q, const lzb_seed* sd,
                        const int8_t ctb[256], const lzb_seed_params* P,
                        lzb_segment** segs, uint64_t* nsegs, lzb_seed_stats* stats) {
    if (!c->sub) return fail("lzb_set_scoring has not been called");
    u32 start = P->start, end = P->end ? P->end : q->len;
    if (end <= start) return fail("in seed_hit_search(), interval is void (%u-%u)", start, end);
    if (end > q->len) return fail("in seed_hit_search(), interval end is bad (%u>%u)", end, q->len);
    if (sd->length < 2) return fail("seed length must be at least two (yours is %d)", sd->length);
    if (P->gfExtend == LZB_GFEX_MISMATCH && (P->gfMismatches < 1 || P->gfMismatches > LZB_GFEX_MISMATCH_MAX))
        return fail("%d is out of range for N-mismatch (valid range is 1..%d)", P->gfMismatches, LZB_GFEX_MISMATCH_MAX);
    search S; memset(&S, 0, sizeof S);
    S.c = c; S.t = t; S.q = q; S.sd = sd; S.p = P;
    int hb = P->hashBits ? P->hashBits : 16;
    S.hmask = (1u << hb) - 1;
    S.E = calloc((size_t)1 << hb, 4);
    if (P->recoverSeeds) {
        if (P->gfExtend != LZB_GFEX_XDROP && P->gfExtend != LZB_GFEX_NONE) return fail("recoverSeeds is built for x-drop extension and --nogfextend only");
        S.A = calloc((size_t)1 << hb, 4);
    }
    if (P->twinMinSpan > 0) {
        if (P->gfExtend != LZB_GFEX_XDROP && P->gfExtend != LZB_GFEX_NONE) return fail("twins are built for x-drop extension and --nogfextend only");
        if (P->twinMaxSpan < P->twinMinSpan) return fail("maxGap for twins can't be less than min gap");
        S.Qsize = P->seedQueueSize > 0 ? (u32)P->seedQueueSize : 256u * 1024u;       /* defaultSeedHitQueueSize diag_hash.h:112 */
        S.Qnum = S.Qsize;                                                             /* diag_hash.c:168 */
        S.Q = calloc(S.Qsize, sizeof *S.Q);
        S.Qlast = calloc((size_t)1 << hb, 4);
    }
    int L = sd->length;
    if (q->len >= (u32)L) {
        u64 w = 0; int run = 0;
        for (u32 i = start; i < end; i++) {
            int b = ctb[q->v[i]];
            if (b < 0) { run = 0; w = 0; continue; }
            w = (w << 2) | (u64)b; run++;
            if (run < L) continue;
            u32 pos2 = i + 1;
            u32 packed = pack_word(sd, w);
            S.st.wordsInQuery++;
            probe(&S, packed, pos2);                                   /* :522 */
            if (sd->withTrans == 1) {                                  /* :526-534 */
                for (int f = 0; f < sd->numFlips; f++) probe(&S, packed ^ sd->transFlips[f], pos2);
            } else if (sd->withTrans >= 2) {                           /* :535-549 */
                for (int f = 0; f < sd->numFlips; f++) {
                    probe(&S, packed ^ sd->transFlips[f], pos2);
                    for (int g = f + 1; g < sd->numFlips; g++)
                        probe(&S, packed ^ sd->transFlips[f] ^ sd->transFlips[g], pos2);
                }
            }
            if (P->searchLimit > 0 && P->twinMinSpan <= 0 && S.st.hsps > P->searchLimit) break;   /* searchToGo < 0, :551; the twin processor never counts (:1814-2046 has no searchToGo--) */
        }
    }
    free(S.E); free(S.A); free(S.Q); free(S.Qlast);
    *segs = S.out; *nsegs = S.n;
    if (stats) *stats = S.st;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * gapped stage
 * ---------------------------------------------------------------------------------------- */

/* segment_peak gapped_extend.c:515-559 */
static u32 peak_offset(const u8* s1, const u8* s2, u32 len, const s32* sub) {
    if (len <= 31) return len / 2;
    s32 sum = 0;
    for (u32 i = 0; i < 31; i++) sum += sub[(u32)s1[i] * 256 + s2[i]];
    s32 best = sum; u32 peak = 15;
    for (u32 i = 31; i < len; i++) {
        sum -= sub[(u32)s1[i - 31] * 256 + s2[i - 31]];
        sum += sub[(u32)s1[i] * 256 + s2[i]];
        if (sum > best) { best = sum; peak = i - 15; }
    }
    return peak;
}

/* reduce_to_points gapped_extend.c:463-512 */
int lzb_reduce_to_points(lzb_ctx* c, lzb_target* t, lzb_query* q, lzb_segment* a, uint64_t n) {
    for (u64 i = 0; i < n; i++) {
        u32 pk = peak_offset(t->v + a[i].pos1, q->v + a[i].pos2, a[i].length, c->sub);
        a[i].pos1 += pk; a[i].pos2 += pk; a[i].length = 0;
    }
    return 0;
}

/* --- edit scripts, edit_script.c:261-330 --- */
#define MAX_RPT ((1u << 30) - 1)
static lzb_editscript* es_new(void) {
    lzb_editscript* s = calloc(1, sizeof(lzb_editscript) + 49 * 4);
    s->size = 50; return s;
}
static void es_room(lzb_editscript** ps, u32 extra) {
    lzb_editscript* s = *ps;
    if (s->len + extra <= s->size) return;
    u32 nsz = s->len + extra; nsz += nsz / 2 + 16;
    s = realloc(s, sizeof(lzb_editscript) + (size_t)(nsz - 1) * 4);
    s->size = nsz; *ps = s;
}
static void es_put(lzb_editscript** ps, u32 op, u32 rpt) {
    es_room(ps, 1);
    lzb_editscript* s = *ps;
    s->op[s->len++] = (op & 3) | (rpt << 2);
    s->tailOp = op;
}
static void es_add(lzb_editscript** ps, u32 op, u32 rpt) {
    lzb_editscript* s = *ps;
    if ((s->tailOp & 3) == op) {
        u32* tail = &s->op[s->len - 1];
        u32 tr = *tail >> 2;
        if ((u64)tr + rpt <= MAX_RPT) { *tail += rpt << 2; return; }
        *tail = (op & 3) | (MAX_RPT << 2);
        rpt = tr + rpt - MAX_RPT;
    }
    while (rpt > MAX_RPT) { es_put(ps, op, MAX_RPT); rpt -= MAX_RPT; }
    es_put(ps, op, rpt);
}
/* edit_script_reverse :417-423 (tailOp intentionally untouched, as in the reference) */
static void es_reverse(lzb_editscript* s) {
    if (s->len < 2) return;
    for (u32 i = 0, j = s->len - 1; i < j; i++, j--) { u32 t = s->op[i]; s->op[i] = s->op[j]; s->op[j] = t; }
}
/* edit_script_append :346-386 */
static void es_append(lzb_editscript** pd, lzb_editscript* src) {
    if (src->len == 0) return;
    es_room(pd, src->len + 1);
    lzb_editscript* d = *pd;
    u32 si = 0, n = src->len;
    u32 sop = src->op[0] & 3;
    if (sop == d->tailOp) {
        u32 dr = d->op[d->len - 1] >> 2, sr = src->op[0] >> 2;
        if ((u64)dr + sr <= MAX_RPT) d->op[d->len - 1] += sr << 2;
        else { d->op[d->len - 1] = sop | (MAX_RPT << 2); d->op[d->len] = sop | ((dr + sr - MAX_RPT) << 2); d->len++; }
        si = 1; n--;
    }
    memcpy(&d->op[d->len], &src->op[si], (size_t)n * 4);
    d->len += n; d->tailOp = src->tailOp;
}

/* --- alignment bookkeeping (galign/aliseg gapped_extend.c:204-245, as index-based records) --- */
enum { SEG_DIAG = 0, SEG_HORZ = 1, SEG_VERT = 2 };
typedef struct { int type; u32 b1, b2, e1, e2; } aseg;
typedef struct { int al, sg; } segref;               /* al < 0 => NULL */
typedef struct {
    u32 pos1, pos2, end1, end2; u64 hspId;
    aseg* segs; int nsegs;
    segref left1, right1, left2, right2;              /* leftSeg1/leftAlign1 ... */
    lzb_alignel* align;
    int next, prev;                                   /* obi / oed links (indices, -1 = NULL) */
} galn;

typedef struct {
    lzb_ctx* c; const u8* s1; const u8* s2; u32 len1, len2;
    const lzb_gapped_params* P;
    galn* al; int nal;
    int obi, oed;                                     /* list heads */
    u8* tb; u32 tbLen;
    lzb_gapped_stats st;
} genv;

static const segref NOSEG = { -1, -1 };
static inline aseg* SEG(genv* G, segref r) { return &G->al[r.al].segs[r.sg]; }

/* msp_left_right gapped_extend.c:3953-4040 */
static int anchor_neighbours(genv* G, galn* m) {
    u32 pos1 = m->pos1, pos2 = m->pos2;
    u32 right = 0xFFFFFFFFu, left = 0xFFFFFFFFu;
    segref R = NOSEG, Lf = NOSEG;
    for (int o = G->obi; o >= 0 && G->al[o].pos1 <= pos1; o = G->al[o].next) {
        galn* x = &G->al[o];
        if (x->end1 < pos1) continue;
        int k = 0;
        while (k < x->nsegs && x->segs[k].e1 < pos1) k++;
        if (k == x->nsegs) continue;
        aseg* bp = &x->segs[k];
        s32 d;
        if (bp->type == SEG_DIAG) d = (s32)(bp->b2 - pos2) + (s32)(pos1 - bp->b1);
        else                      d = (s32)(bp->b2 - pos2);
        if (d == 0) return 0;
        if (d > 0 && (u32)d < right)      { right = (u32)d;  R.al = o;  R.sg = k; }
        else if (d < 0 && (u32)-d < left) { left  = (u32)-d; Lf.al = o; Lf.sg = k; }
    }
    m->right1 = m->right2 = R; m->left1 = m->left2 = Lf;
    return 1;
}

/* align_left_right gapped_extend.c:4078-4175 */
static void alignment_neighbours(genv* G, galn* m) {
    u32 pos1 = m->pos1, pos2 = m->pos2, end1 = m->end1, end2 = m->end2;
    u32 rB = 0xFFFFFFFFu, rT = rB, lB = rB, lT = rB;
    segref RB = NOSEG, RT = NOSEG, LB = NOSEG, LT = NOSEG;
    for (int o = G->obi; o >= 0; o = G->al[o].next) {
        galn* x = &G->al[o];
        if (x->pos1 > end1 || x->end1 < pos1) continue;
        int k = 0;
        while (k < x->nsegs && !(x->segs[k].type != SEG_HORZ && x->segs[k].e1 >= pos1)) k++;
        if (k < x->nsegs && x->segs[k].b1 <= pos1) {
            aseg* bp = &x->segs[k]; s32 d;
            if (bp->type == SEG_DIAG) d = (s32)(bp->b2 - pos2) + (s32)(pos1 - bp->b1);
            else                      d = (s32)(bp->b2 - pos2);
            if (d > 0 && (u32)d < rB)       { rB = (u32)d;  RB.al = o; RB.sg = k; }
            else if (d < 0 && (u32)-d < lB) { lB = (u32)-d; LB.al = o; LB.sg = k; }
        }
        while (k < x->nsegs && !(x->segs[k].type != SEG_HORZ && x->segs[k].e1 >= end1)) k++;
        if (k < x->nsegs) {
            aseg* bp = &x->segs[k]; s32 d;
            if (bp->type == SEG_DIAG) d = (s32)(bp->b2 - end2) + (s32)(end1 - bp->b1);
            else                      d = (s32)(bp->b2 - end2);
            if (d > 0 && (u32)d < rT)       { rT = (u32)d;  RT.al = o; RT.sg = k; }
            else if (d < 0 && (u32)-d < lT) { lT = (u32)-d; LT.al = o; LT.sg = k; }
        }
    }
    m->right1 = RB; m->right2 = RT; m->left1 = LB; m->left2 = LT;
}

/* insert_align gapped_extend.c:4210-4240 */
static void list_insert(genv* G, int mi) {
    galn* m = &G->al[mi];
    int q = -1, p = G->obi;
    while (p >= 0 && G->al[p].pos1 < m->pos1) { q = p; p = G->al[p].next; }
    if (q >= 0) { G->al[q].next = mi; m->next = p; } else { m->next = G->obi; G->obi = mi; }
    q = -1; p = G->oed;
    while (p >= 0 && G->al[p].end1 > m->end1) { q = p; p = G->al[p].prev; }
    if (q >= 0) { G->al[q].prev = mi; m->prev = p; } else { m->prev = G->oed; G->oed = mi; }
}

/* save_seg gapped_extend.c:5220-5262: diagonal pieces joined by a horizontal or vertical piece */
static void add_diag(galn* m, u32 b1, u32 b2, u32 e1, u32 e2) {
    m->segs = realloc(m->segs, (size_t)(m->nsegs + 2) * sizeof(aseg));
    if (m->nsegs > 0) {
        aseg* last = &m->segs[m->nsegs - 1];
        aseg g; g.type = (b1 == last->e1 + 1) ? SEG_HORZ : SEG_VERT;
        g.b1 = last->e1 + 1; g.b2 = last->e2 + 1; g.e1 = b1 - 1; g.e2 = b2 - 1;
        m->segs[m->nsegs++] = g;
    }
    aseg d = { SEG_DIAG, b1, b2, e1, e2 };
    m->segs[m->nsegs++] = d;
}

/* per-DP walking state for the bounding segments (update_LR_bounds :4588, next/prev_sweep_seg :4754) */
typedef struct { segref seg; } bwalk;

static s32 sweep_next(genv* G, int lookRight, segref* bp, u32 row, u32 a1, u32 a2) {
    galn* m = &G->al[bp->al];
    if (bp->sg + 1 < m->nsegs) {
        bp->sg++;
        if (m->segs[bp->sg].type == SEG_HORZ) bp->sg++;   /* a horizontal piece is never last */
        return (s32)(m->segs[bp->sg].b2 - a2);
    }
    *bp = lookRight ? m->right2 : m->left2;
    if (bp->al < 0) return 0;
    aseg* s = SEG(G, *bp);
    if (s->type == SEG_DIAG) return (s32)row + (s32)(s->b2 - a2) - (s32)(s->b1 - a1);
    return (s32)(s->b2 - a2);
}
static s32 sweep_prev(genv* G, int lookRight, segref* bp, u32 row, u32 a1, u32 a2) {
    galn* m = &G->al[bp->al];
    if (bp->sg - 1 >= 0) {
        bp->sg--;
        if (m->segs[bp->sg].type == SEG_HORZ) bp->sg--;   /* nor first */
        return (s32)(a2 - m->segs[bp->sg].e2);
    }
    *bp = lookRight ? m->right1 : m->left1;
    if (bp->al < 0) return 0;
    aseg* s = SEG(G, *bp);
    if (s->type == SEG_DIAG) return (s32)row + (s32)(a2 - s->e2) - (s32)(a1 - s->e1);
    return (s32)(a2 - s->e2);
}

/* active segments (update_active_segs :4885-4962, build_active_seg :4989-5040) */
typedef struct { segref seg; u32 x, lastRow; int type, dead; } actseg;

typedef struct {
    s32* C; s32* D; u32* K; u32 cap, msk;       /* ring-buffered sweep row + mask stamps */
} sweeprow;

static void row_grow(sweeprow* r, u32 lo, u32 hi, u32 need) {
    if (need + 8 < r->cap) return;
    u32 ncap = r->cap ? r->cap : 1024;
    while (ncap <= need + 8) ncap *= 2;
    s32* C = malloc((size_t)ncap * 4); s32* D = malloc((size_t)ncap * 4); u32* K = calloc(ncap, 4);
    if (r->cap) for (u32 x = lo; x <= hi; x++) {
        C[x & (ncap - 1)] = r->C[x & r->msk]; D[x & (ncap - 1)] = r->D[x & r->msk]; K[x & (ncap - 1)] = r->K[x & r->msk];
    }
    free(r->C); free(r->D); free(r->K);
    r->C = C; r->D = D; r->K = K; r->cap = ncap; r->msk = ncap - 1;
}

static void act_build(genv* G, int rev, actseg* a, sweeprow* R, u32 row, u32 a1, u32 a2, u32 LY, u32 RY) {
    aseg* s = SEG(G, a->seg);
    a->type = s->type;
    if (!rev) { a->x = s->b2 - a2; a->lastRow = s->e1 - a1; }
    else      { a->x = a2 - s->e2; a->lastRow = a1 - s->b1; }
    if (a->type != SEG_HORZ) {
        if (a->x >= LY && a->x <= RY) R->K[a->x & R->msk] = row;
    } else {
        u32 hend = !rev ? s->e2 - a2 : a2 - s->b2;
        u32 lo = a->x > LY ? a->x : LY, hi = hend < RY ? hend : RY;
        for (u32 i = lo; i <= hi && i >= lo; i++) R->K[i & R->msk] = row;
    }
}
static int act_step_seg(genv* G, int rev, segref* r) {      /* next_seg :4883 */
    if (!rev) { if (r->sg + 1 < G->al[r->al].nsegs) { r->sg++; return 1; } return 0; }
    if (r->sg - 1 >= 0) { r->sg--; return 1; }
    return 0;
}

/*
 * ydrop_one_sided_align gapped_extend.c:3388-3868.  Sequence access: forward A(k)=seq1[a1+k],
 * B(k)=seq2[a2+k]; reversed A(k)=seq1[a1+1-k], B(k)=seq2[a2+1-k] (the reference reads the same
 * bytes through rev1/rev2, :2512-2527); one step past either end reads the NUL terminator.
 */
static inline u8 at(const u8* v, u32 len, s64 i) { return (i < 0 || i >= (s64)len) ? 0 : v[i]; }

static s32 one_sided(genv* G, int rev, u32 a1, u32 a2, u32 M, u32 N,
                     segref leftSeg, segref rightSeg, int alignList,
                     lzb_editscript** script, u32* oend1, u32* oend2) {
    if (N == 0 || M == 0) { *oend1 = *oend2 = 0; return 0; }
    const s32* subm = G->c->sub;
    s32 gapE = G->c->gapExtend, gapOE = G->c->gapOpen + gapE, yDrop = G->P->yDrop;
    int trim = G->P->trimToPeak;
    s64 tbLen = G->tbLen; u8* tb = G->tb;
    s32 yTail;
    if (gapE != 0) yTail = yDrop / gapE + 6;
    else yTail = (N < 500000u) ? (s32)N + 1 : 500000;
    /* initial bounds :3500-3543 */
    s32 L = 0, R = (s32)(N + 1);
    if (leftSeg.al >= 0) { aseg* s = SEG(G, leftSeg); L = (s32)(s->b2 - a2); if (s->type == SEG_DIAG) L -= (s32)(s->b1 - a1); }
    if (rightSeg.al >= 0) { aseg* s = SEG(G, rightSeg); R = (s32)(s->b2 - a2); if (s->type == SEG_DIAG) R -= (s32)(s->b1 - a1); }
    if (rev) {
        if (leftSeg.al < 0 && rightSeg.al >= 0) { L = -R + 1; R = (s32)(N + 1); }
        else if (leftSeg.al >= 0 && rightSeg.al < 0) { R = -L - 1; L = 0; }
        else if (leftSeg.al >= 0 && rightSeg.al >= 0) { s32 t = -L - 1; L = -R + 1; R = t; }
    }
    actseg* act = NULL; int nact = 0, capact = 0;
    s64* tbRow = NULL; u32 capRows = 0;
    s64 used = 0;
    if (yTail > tbLen) { fail("not enough space in trace_back array"); return 0; }
    sweeprow Rw; memset(&Rw, 0, sizeof Rw);
    row_grow(&Rw, 0, 0, (u32)yTail + 1024);
    /* first row :3576-3591 */
    capRows = 1 << 16; tbRow = malloc((size_t)capRows * 8);
    tbRow[0] = 0;
    Rw.C[0] = 0; Rw.D[0] = -gapOE; tb[used++] = 0;
    s32 c = -gapOE, ct = 0; u32 col;
    for (col = 1; col <= N && ct >= -yDrop; col++) {
        row_grow(&Rw, 0, col - 1, col + 2);
        Rw.C[col & Rw.msk] = ct = c; Rw.D[col & Rw.msk] = c - gapOE; c -= gapE; tb[used++] = 1;
    }
    u32 LY = 0, RY = col;
    u32 end1 = 0, end2 = 0; s32 best = 0, bnd = NEG_INF; int endIsBnd = 0;
    u32 row; u64 cells = col;                                 /* the first row counts too, :3593 */
    for (row = 1; row <= M; row++) {
        u32 prevLY = LY;
        /* update_LR_bounds :4588-4724 */
        if (!rev) {
            if (leftSeg.al >= 0) {
                aseg* s = SEG(G, leftSeg);
                if (s->e1 >= row + a1) { if (s->type == SEG_DIAG) L++; }
                else L = sweep_next(G, 0, &leftSeg, row, a1, a2) + 1;
            }
            if (leftSeg.al >= 0) { if ((s32)LY < L) LY = (u32)L; }
            if (rightSeg.al >= 0) {
                aseg* s = SEG(G, rightSeg);
                if (s->e1 >= row + a1) { if (s->type == SEG_DIAG) R++; }
                else R = sweep_next(G, 1, &rightSeg, row, a1, a2) - 1;
            }
            if (rightSeg.al >= 0) { if (R <= 0) RY = 0; else if ((u32)R < RY) RY = (u32)R; }
        } else {
            if (rightSeg.al >= 0) {
                aseg* s = SEG(G, rightSeg);
                if (s->b1 <= a1 - row) { if (s->type == SEG_DIAG) L++; }
                else L = sweep_prev(G, 1, &rightSeg, row, a1, a2) + 1;
            }
            if (rightSeg.al >= 0) { if ((s32)LY < L) LY = (u32)L; }
            if (leftSeg.al >= 0) {
                aseg* s = SEG(G, leftSeg);
                if (s->b1 <= a1 - row) { if (s->type == SEG_DIAG) R++; }
                else R = sweep_prev(G, 0, &leftSeg, row, a1, a2) - 1;
            }
            if (leftSeg.al >= 0) { if (R <= 0) RY = 0; else if ((u32)R < RY) RY = (u32)R; }
        }
        /* update_active_segs :4885-4962 */
        row_grow(&Rw, prevLY, RY > prevLY ? RY : prevLY, (RY > prevLY ? RY - prevLY : 0) + (u32)yTail + 16);
        for (int k = 0; k < nact; k++) {
            actseg* a = &act[k];
            if (a->lastRow >= row) {
                if (a->type == SEG_DIAG) a->x++;
                if (a->x >= LY && a->x <= RY) Rw.K[a->x & Rw.msk] = row;
            } else if (act_step_seg(G, rev, &a->seg)) {
                act_build(G, rev, a, &Rw, row, a1, a2, LY, RY);
                if (a->type == SEG_HORZ) { act_step_seg(G, rev, &a->seg); act_build(G, rev, a, &Rw, row, a1, a2, LY, RY); }
            } else a->dead = 1;
        }
        for (;;) {
            if (alignList < 0) break;
            galn* x = &G->al[alignList];
            if (!rev) { if (x->pos1 - a1 != row) break; }
            else      { if (a1 - x->end1 != row) break; }
            if (nact == capact) { capact = capact * 2 + 8; act = realloc(act, (size_t)capact * sizeof(actseg)); }
            /* add_new_active prepends; order of the list does not influence the stamps */
            actseg* a = &act[nact++]; a->dead = 0;
            a->seg.al = alignList; a->seg.sg = !rev ? 0 : x->nsegs - 1;
            act_build(G, rev, a, &Rw, row, a1, a2, LY, RY);
            alignList = !rev ? x->next : x->prev;
        }
        { int w = 0; for (int k = 0; k < nact; k++) if (!act[k].dead) act[w++] = act[k]; nact = w; }
        /* traceback capacity :3636-3662 */
        if (row + 1 >= capRows) { capRows *= 2; tbRow = realloc(tbRow, (size_t)capRows * 8); }
        if (RY < LY) RY = LY;
        s64 need = (s64)(RY - LY) + yTail;
        if (used + need >= tbLen) { G->st.truncated++; break; }
        tbRow[row] = used - (s64)LY;
        /* sweep :3669-3774 */
        u8 arow = !rev ? at(G->s1, G->len1, (s64)a1 + row) : at(G->s1, G->len1, (s64)a1 + 1 - (s64)row);
        const s32* sub = subm + (u32)arow * 256;
        col = LY; u32 leftCol = LY, npCol = LY, wcol = LY;
        s32 i = NEG_INF; c = NEG_INF;
        for (; col < RY && col <= N; col++) {
            u8 bnext = !rev ? at(G->s2, G->len2, (s64)a2 + col + 1) : at(G->s2, G->len2, (s64)a2 + 1 - (s64)(col + 1));
            s32 d = Rw.D[col & Rw.msk];
            s32 prevC = Rw.C[col & Rw.msk];
            u8 link; int dead = 0;
            if (nact > 0 && Rw.K[col & Rw.msk] == row) dead = 1;
            else if (d > c || i > c) {
                if (d >= i) { c = d; link = 2 | 4 | 8; } else { c = i; link = 1 | 4 | 8; }
                if (c < best - yDrop) dead = 1;
                else { i -= gapE; Rw.D[col & Rw.msk] = d - gapE; }
            } else {
                if (c < best - yDrop) dead = 1;
                else {
                    if (c >= best) { best = c; end1 = row; end2 = col; endIsBnd = 0; }
                    if (!trim && c >= bnd && (row == M || col == N)) { bnd = c; end1 = row; end2 = col; endIsBnd = 1; }
                    s32 open = c - gapOE;
                    d -= gapE;
                    if (open > d) { Rw.D[col & Rw.msk] = open; link = 0; } else { Rw.D[col & Rw.msk] = d; link = 8; }
                    i -= gapE;
                    if (open > i) i = open; else link |= 4;
                }
            }
            if (dead) {                                   /* prune macro :2977-2987 */
                c = prevC + sub[bnext];
                if (col == LY) { LY++; wcol = LY; }
                else { i = NEG_INF; Rw.D[col & Rw.msk] = Rw.C[col & Rw.msk] = NEG_INF; wcol = col + 1; }
                tb[used++] = 0;
                continue;
            }
            npCol = col;
            s32 cn = prevC + sub[bnext];
            Rw.C[col & Rw.msk] = c; wcol = col + 1;
            c = cn;
            tb[used++] = link;
        }
        cells += col - leftCol;
        if (LY >= RY) break;
        s32 NN = (rightSeg.al >= 0 && R > 0) ? R - 1 : (s32)N;
        if (RY > npCol + 1) RY = npCol + 1;
        else {
            while (i >= best - yDrop && (s32)RY <= NN) {     /* row prolongation :3801-3811 */
                row_grow(&Rw, LY, wcol, wcol - prevLY + 8);
                Rw.C[wcol & Rw.msk] = i; Rw.D[wcol & Rw.msk] = i - gapOE; wcol++;
                i -= gapE; tb[used++] = 1; RY++;
            }
        }
        if ((s32)RY <= NN) {                                  /* dead sentinel :3819-3827 */
            row_grow(&Rw, LY, wcol, wcol - prevLY + 8);
            Rw.D[wcol & Rw.msk] = Rw.C[wcol & Rw.msk] = NEG_INF; RY++;
        }
    }
    G->st.dpCells += cells; G->st.dpCellsComputed += cells; G->st.dpRows += row;
    /* traceback :3847-3859 */
    u32 r = end1, cc = end2; u8 prevOp = 0, op;
    *oend1 = end1; *oend2 = end2;
    for (; r >= 1 || cc > 0; prevOp = op) {
        u8 link = tb[tbRow[r] + (s64)cc];
        op = link & 3;
        if (prevOp == 1 && (link & 4)) op = 1;
        if (prevOp == 2 && (link & 8)) op = 2;
        if (op == 1)      { cc--;      es_add(script, LZB_OP_INS, 1); }
        else if (op == 2) { r--;       es_add(script, LZB_OP_DEL, 1); }
        else              { r--; cc--; es_add(script, LZB_OP_SUB, 1); }
    }
    free(act); free(tbRow); free(Rw.C); free(Rw.D); free(Rw.K);
    return endIsBnd ? bnd : best;
}

/* score_alignment gapped_extend.c:5631-5690 */
static s32 rescore(genv* G, u32 p1, u32 p2, lzb_editscript* s) {
    const s32* sub = G->c->sub; s32 sim = 0;
    for (u32 k = 0; k < s->len; k++) {
        u32 rpt = s->op[k] >> 2, op = s->op[k] & 3;
        if (!rpt) continue;
        if (op == LZB_OP_SUB) { for (u32 j = 0; j < rpt; j++) sim += sub[(u32)G->s1[p1 + j] * 256 + G->s2[p2 + j]]; p1 += rpt; p2 += rpt; }
        else if (op == LZB_OP_INS) { sim -= G->c->gapOpen + (s32)rpt * G->c->gapExtend; p2 += rpt; }
        else { sim -= G->c->gapOpen + (s32)rpt * G->c->gapExtend; p1 += rpt; }
    }
    return sim;
}

typedef struct { s32 s; u32 start1, start2, stop1, stop2; lzb_editscript* script; } ydres;

/* ydrop_align gapped_extend.c:2459-2584 + lop_initial/final_indels :2589-2683 */
/* the partition of a [multi] sequence that holds position a (gapped_extend.c:1357-1372: io.low = sepBefore + 1,
 * io.high = sepAfter); partitions are delimited by NUL bytes (sequences.h:188-191), an ordinary sequence has none */
static void partition_limits(const u8* v, u32 len, u32 a, u32* low, u32* high) {
    s64 i = a; while (i >= 0 && v[i] != 0) i--;
    u32 j = a; while (j < len && v[j] != 0) j++;
    *low = (u32)(i + 1); *high = j;
}

static void two_sided(genv* G, galn* m, int above, int below, ydres* o) {
    u32 a1 = m->pos1, a2 = m->pos2, e1, e2, low1, high1, low2, high2;
    partition_limits(G->s1, G->len1, a1, &low1, &high1);
    partition_limits(G->s2, G->len2, a2, &low2, &high2);
    lzb_editscript* sl = es_new();
    s32 left = one_sided(G, 1, a1, a2, a1 + 1 - low1, a2 + 1 - low2, m->left1, m->right1, below, &sl, &e1, &e2);
    o->start1 = a1 + 1 - e1; o->start2 = a2 + 1 - e2;
    lzb_editscript* sr = es_new();
    s32 right = one_sided(G, 0, a1, a2, high1 - (a1 + 1), high2 - (a2 + 1), m->left1, m->right1, above, &sr, &e1, &e2);
    o->stop1 = a1 + e1; o->stop2 = a2 + e2;
    es_reverse(sr); es_append(&sl, sr); free(sr);
    o->s = left + right; o->script = sl;
    if (sl->len != 0) {
        if ((sl->op[0] & 3) != LZB_OP_SUB) {
            u32 p1 = o->start1, p2 = o->start2, k = 0;
            for (; k < sl->len; k++) {
                u32 op = sl->op[k] & 3, rpt = sl->op[k] >> 2;
                if (op == LZB_OP_SUB) break;
                if (op == LZB_OP_INS) p2 += rpt; else p1 += rpt;
            }
            if (k == sl->len) { o->s = WORST_SCORE; return; }
            o->start1 = p1; o->start2 = p2;
            sl->len -= k; memmove(sl->op, sl->op + k, (size_t)sl->len * 4);
            o->s = rescore(G, o->start1, o->start2, sl);
        }
        if ((sl->op[sl->len - 1] & 3) != LZB_OP_SUB) {
            u32 p1 = o->stop1, p2 = o->stop2, k = sl->len;
            while (k > 0) {
                k--;
                u32 op = sl->op[k] & 3, rpt = sl->op[k] >> 2;
                if (op == LZB_OP_SUB) { k++; break; }
                if (op == LZB_OP_INS) p2 -= rpt; else p1 -= rpt;
            }
            if (k == 0) { o->s = WORST_SCORE; return; }
            o->stop1 = p1; o->stop2 = p2; sl->len = k;
            o->s = rescore(G, o->start1, o->start2, sl);
        }
    }
}

/* qSegmentsByDecreasingScore segment.c:1748-1771 */
static int cmp_anchor(const void* A, const void* B) {
    const lzb_segment* a = A; const lzb_segment* b = B;
    if (a->s < b->s) return 1;  if (a->s > b->s) return -1;
    if (a->length < b->length) return -1;  if (a->length > b->length) return 1;
    if (a->pos2 < b->pos2) return -1;  if (a->pos2 > b->pos2) return 1;
    if (a->pos1 < b->pos1) return -1;  if (a->pos1 > b->pos1) return 1;
    if (a->id < b->id) return -1;  if (a->id > b->id) return 1;
    return 0;
}

/* identical_sequences gapped_extend.c:1886-1930 */
static int same_sequences(genv* G, s32* score) {
    if (G->len1 != G->len2) return 0;
    s32 s = 0;
    for (u32 i = 0; i < G->len1; i++) {
        u8 a = G->s1[i], b = G->s2[i];
        if (a >= 'a' && a <= 'z') a -= 32;
        if (b >= 'a' && b <= 'z') b -= 32;
        if (a != b) return 0;
        s32 v = G->c->sub[(u32)a * 256 + b];
        if (s == 0x7FFFFFFF) ;
        else if (v <= 0 || s < 0x7FFFFFFF - v) s += v;
        else s = 0x7FFFFFFF;
    }
    *score = s; return 1;
}

/* identical_partition_of_sequence gapped_extend.c:2034-2120 + score_identical_partition_of :2217: the first partition of
 * a [multi] target that equals the (unpartitioned) query, case-insensitively; its NUL-to-NUL bounds and score */
static int same_as_partition(genv* G, u32* sepBefore, u32* sepAfter, s32* score) {
    if (memchr(G->s2, 0, G->len2) || !memchr(G->s1, 0, G->len1)) return 0;
    for (u32 before = 0; before < G->len1;) {
        if (G->s1[before] != 0) return 0;                    /* a partitioned sequence starts with a separator */
        u32 after = before + 1; while (after < G->len1 && G->s1[after] != 0) after++;
        if (after - (before + 1) == G->len2) {
            s32 s = 0; u32 i = 0;
            for (; i < G->len2; i++) {
                u8 a = G->s1[before + 1 + i], b = G->s2[i];
                if (a >= 'a' && a <= 'z') a -= 32;
                if (b >= 'a' && b <= 'z') b -= 32;
                if (a != b) break;
                s32 v = G->c->sub[(u32)a * 256 + b];
                if (s == 0x7FFFFFFF) ;
                else if (v <= 0 || s < 0x7FFFFFFF - v) s += v;
                else s = 0x7FFFFFFF;
            }
            if (i == G->len2) { *sepBefore = before; *sepAfter = after; *score = s; return 1; }
        }
        before = after;
    }
    return 0;
}

/* gapped_extend gapped_extend.c:1012-1604 */
int lzb_gapped_extend(lzb_ctx* c, lzb_target* t, lzb_query* q, const uint8_t* h1, const uint8_t* h2,
                      lzb_segment* anchors, uint64_t n, const lzb_gapped_params* P,
                      lzb_alignel** list, lzb_gapped_stats* stats) {
    if (!c->sub) return fail("lzb_set_scoring has not been called");
    if (P->tracebackBytes < 8) return fail("in new_traceback(), size can't be %u", P->tracebackBytes);
    genv G; memset(&G, 0, sizeof G);
    G.c = c; G.s1 = t->v; G.s2 = q->v; G.len1 = t->len; G.len2 = q->len; G.P = P;
    G.obi = G.oed = -1;
    G.tbLen = 1 + (P->tracebackBytes - 8);           /* new_traceback :2272-2290 */
    G.tb = malloc(P->tracebackBytes);
    G.st.anchors = n;
    if (n) qsort(anchors, n, sizeof(lzb_segment), cmp_anchor);   /* batched_segments :1675 */
    G.al = calloc(n + 1, sizeof(galn)); G.nal = (int)n + 1;
    for (u64 i = 0; i < n; i++) { G.al[i].pos1 = anchors[i].pos1; G.al[i].pos2 = anchors[i].pos2; G.al[i].hspId = anchors[i].hspId; }
    s32 ts; u64 pairedBases = 0;
    /* identical_sequences also requires equal revCompFlags (:1905); the caller vouches for that
     * through identityCheck */
    if (P->identityCheck && same_sequences(&G, &ts)) {
        galn* m = &G.al[n];                              /* :1113-1151 trivial self alignment */
        m->pos1 = m->pos2 = 0; m->end1 = m->end2 = G.len1 - 1;
        m->left1 = m->left2 = m->right1 = m->right2 = NOSEG;
        add_diag(m, 0, 0, m->end1, m->end2);
        list_insert(&G, (int)n);
        lzb_alignel* a = calloc(1, sizeof *a);
        a->script = es_new(); es_add(&a->script, LZB_OP_SUB, G.len1);
        a->beg1 = a->beg2 = 1; a->end1 = a->end2 = G.len1;
        a->seq1 = h1; a->seq2 = h2;
        a->s = ts < P->scoreThreshold ? P->scoreThreshold : ts;
        a->isTrivial = 1; m->align = a;
    }
    u32 sepB = 0, sepA = 0;
    if (P->identityCheck && !G.al[n].align && same_as_partition(&G, &sepB, &sepA, &ts)) {
        galn* m = &G.al[n];                              /* :1185-1230 the query is one partition of the target */
        m->pos1 = sepB + 1; m->pos2 = 0; m->end1 = sepA - 1; m->end2 = G.len2 - 1;
        m->left1 = m->left2 = m->right1 = m->right2 = NOSEG;
        add_diag(m, m->pos1, m->pos2, m->end1, m->end2);
        list_insert(&G, (int)n);
        lzb_alignel* a = calloc(1, sizeof *a);
        a->script = es_new(); es_add(&a->script, LZB_OP_SUB, G.len2);
        a->beg1 = sepB + 2; a->beg2 = 1; a->end1 = sepA; a->end2 = G.len2;
        a->seq1 = h1; a->seq2 = h2;
        a->s = ts < P->scoreThreshold ? P->scoreThreshold : ts;
        a->isTrivial = 1; m->align = a;
    }
    for (u64 i = 0; i < n; i++) {
        galn* m = &G.al[i];
        if (!anchor_neighbours(&G, m)) continue;
        /* get_above_below :4043-4060 */
        int below = G.oed; while (below >= 0 && !(G.al[below].end1 < m->pos1)) below = G.al[below].prev;
        int above = G.obi; while (above >= 0 && !(G.al[above].pos1 > m->pos1)) above = G.al[above].next;
        ydres r; two_sided(&G, m, above, below, &r);
        G.st.anchorsExtended++;
        /* format_alignment :5153-5198 */
        u32 beg1 = r.start1 + 1, end1 = r.stop1 + 1, beg2 = r.start2 + 1, end2 = r.stop2 + 1;
        u32 height = end1 - beg1 + 1, width = end2 - beg2 + 1, k = 0;
        for (u32 ii = 0, jj = 0; ii < height || jj < width;) {
            u32 si = ii, sj = jj, run = 0;
            while (k < r.script->len && (r.script->op[k] & 3) == LZB_OP_SUB) { run += r.script->op[k] >> 2; k++; }
            ii += run; jj += run;
            add_diag(m, beg1 + si - 1, beg2 + sj - 1, beg1 + ii - 2, beg2 + jj - 2);
            if (ii < height || jj < width) {
                if (k < r.script->len) {
                    u32 op = r.script->op[k] & 3, rpt = r.script->op[k] >> 2;
                    if (op == LZB_OP_INS) jj += rpt; else if (op == LZB_OP_DEL) ii += rpt;
                    k++;
                }
            }
        }
        lzb_alignel* a = calloc(1, sizeof *a);
        a->script = r.script; a->beg1 = beg1; a->beg2 = beg2; a->end1 = end1; a->end2 = end2;
        a->seq1 = h1; a->seq2 = h2; a->s = r.s; a->hspId = m->hspId;
        m->align = a;
        m->pos1 = r.start1; m->pos2 = r.start2; m->end1 = r.stop1; m->end2 = r.stop2;
        if (m->nsegs == 0) continue;                     /* empty alignment (leaks like :1404) */
        if (!P->allBounds && a->s < P->scoreThreshold) {
            free(a->script); free(a); m->align = NULL; free(m->segs); m->segs = NULL; m->nsegs = 0;
            continue;
        }
        alignment_neighbours(&G, m);
        list_insert(&G, (int)i);
        if (P->maxPairedBases > 0) {                      /* :1444-1459, count_paired_bases :5695 */
            for (int sg = 0; sg < m->nsegs; sg++) if (m->segs[sg].type == SEG_DIAG) pairedBases += (u64)m->segs[sg].e1 + 1 - m->segs[sg].b1;
            if (pairedBases > P->maxPairedBases) { G.st.overlyPaired = 1; break; }
        }
    }
    lzb_alignel* head = NULL, *last = NULL;
    for (int o = G.obi; o >= 0; o = G.al[o].next) {
        galn* m = &G.al[o];
        int drop = m->align->s < P->scoreThreshold || (P->inhibitTrivial && m->align->isTrivial);
        if (G.st.overlyPaired && !P->overlyPairedKeep) drop = 1;          /* discard_alignments :1580 */
        if (drop) { free(m->align->script); free(m->align); }
        else { if (!head) head = last = m->align; else { last->next = m->align; last = m->align; } }
    }
    for (int k = 0; k < G.nal; k++) free(G.al[k].segs);
    free(G.al); free(G.tb);
    *list = head;
    if (stats) *stats = G.st;
    return 0;
}

void lzb_free_align_list(lzb_alignel* a) {
    while (a) { lzb_alignel* nx = a->next; free(a->script); free(a); a = nx; }
}
