/*
 * mirror.c -- --self with gapped extension: reflect the alignments across the main diagonal.
 *
 * Host front-end logic of the reference driver (mirror_alignments lastz.c:4229-4408 with the
 * edit-script helpers edit_script.c:221-258, :352-411, :413-440, :442-470, :487-531, :573-690,
 * :820-850 and score_alignment gapped_extend.c:5631-5690), unpartitioned sequences only.  It runs
 * on the finished alignment list, after the hot path; nothing here touches the device.
 * The reference's quirks are kept on purpose: tailOp is NOT refreshed by reverse/mirror/truncate, and
 * edit_script_append merges on that (possibly stale) tailOp.
 */
#include <stdlib.h>
#include <string.h>
#include "lzb_host.h"

#define OP(s, k)  ((s)->op[k] & 3u)
#define RPT(s, k) ((s)->op[k] >> 2)
#define MAXRPT    ((1u << 30) - 1)                      /* maxEditopRepeat edit_script.h:52 */
#define POS_INF   0xFFFFFFFFu                           /* seqposInfinity */

static lzb_editscript* es_copy(const lzb_editscript* s) {          /* edit_script_copy */
    uint32_t n = s->len; size_t bytes = sizeof(lzb_editscript) + (size_t)(n > 0 ? n - 1 : 0) * 4;
    lzb_editscript* c = calloc(1, bytes + 4);
    memcpy(c, s, sizeof(lzb_editscript) - 4 + (size_t)n * 4);
    c->size = n;
    return c;
}
static void es_reverse(lzb_editscript* s) {                         /* edit_script_reverse */
    if (s->len < 2) return;
    for (uint32_t i = 0, j = s->len - 1; i < j; i++, j--) { uint32_t t = s->op[i]; s->op[i] = s->op[j]; s->op[j] = t; }
}
static void es_mirror(lzb_editscript* s) {                          /* edit_script_mirror */
    for (uint32_t i = 0; i < s->len; i++) {
        uint32_t op = OP(s, i), rpt = RPT(s, i);
        if (op == LZB_OP_INS) s->op[i] = LZB_OP_DEL | (rpt << 2);
        else if (op == LZB_OP_DEL) s->op[i] = LZB_OP_INS | (rpt << 2);
    }
}
static void es_trim_head(lzb_editscript* s, uint32_t len) {         /* edit_script_trim_head */
    if (s->len == 0 || len == 0) return;
    uint32_t i, rpt = 0; int shortScript = 1;
    for (i = 0; i < s->len; i++) { rpt = RPT(s, i); if (rpt > len) { shortScript = 0; break; } len -= rpt; }
    if (shortScript) { s->len = 0; return; }
    uint32_t op = s->op[i];
    if (i > 0) { for (uint32_t j = i; j < s->len; j++) s->op[j - i] = s->op[j]; s->len -= i; }
    if (len > 0) s->op[0] = (op & 3u) | ((rpt - len) << 2);
}
static void es_append(lzb_editscript** pd, const lzb_editscript* src) {   /* edit_script_append */
    if (src->len == 0) return;
    lzb_editscript* d = *pd;
    if (d->len + src->len + 1 > d->size) {
        uint32_t nsz = d->len + src->len + 16;
        d = realloc(d, sizeof(lzb_editscript) + (size_t)(nsz - 1) * 4); d->size = nsz; *pd = d;
    }
    uint32_t k = 0, toCopy = src->len;
    if (d->len > 0 && OP(src, 0) == d->tailOp) {
        uint32_t dr = RPT(d, d->len - 1), sr = RPT(src, 0), op = OP(src, 0);
        if ((uint64_t)dr + sr <= MAXRPT) d->op[d->len - 1] += sr << 2;
        else { d->op[d->len - 1] = op | (MAXRPT << 2); d->op[d->len] = op | ((dr + sr - MAXRPT) << 2); d->len++; }
        k = 1; toCopy--;
    }
    memcpy(d->op + d->len, src->op + k, (size_t)toCopy * 4);
    d->len += toCopy; d->tailOp = src->tailOp;
}
static void es_overall(const lzb_editscript* s, uint32_t* pi, uint32_t* pj) {   /* edit_script_overall_len */
    uint32_t i = 0, j = 0;
    for (uint32_t k = 0; k < s->len; k++) {
        uint32_t rpt = RPT(s, k);
        switch (OP(s, k)) { case LZB_OP_SUB: i += rpt; j += rpt; break; case LZB_OP_INS: j += rpt; break; case LZB_OP_DEL: i += rpt; break; }
    }
    *pi = i; *pj = j;
}
/* edit_script_upper_truncate: keep the part of an opposite-strand alignment above the main diagonal */
static int es_upper_truncate(lzb_editscript* s, uint32_t* p1, uint32_t* p2) {
    if (s->len == 0) return 0;
    uint32_t pos1 = *p1, pos2 = *p2, prev1 = 0, prev2 = 0, limit = 0, i, op = 0, rpt;
    if (pos1 > pos2) { s->len = 0; *p1 = *p2 = POS_INF; return 1; }
    int reaches = 0;
    for (i = 0; i < s->len; i++) {
        prev1 = pos1; prev2 = pos2; op = OP(s, i); rpt = RPT(s, i);
        switch (op) {
            case LZB_OP_SUB: pos1 += rpt; pos2 -= rpt; limit = pos2 + 1; break;
            case LZB_OP_INS: pos2 -= rpt; limit = pos2; break;
            case LZB_OP_DEL: pos1 += rpt; limit = pos2; break;
        }
        if (pos1 >= limit) { reaches = 1; break; }
    }
    if (!reaches) return 0;
    s->len = i + 1;
    if (pos1 > pos2) {
        switch (op) {
            case LZB_OP_SUB: rpt = (prev2 + 1 - prev1) / 2; s->op[i] = LZB_OP_SUB | (rpt << 2); pos1 = prev1 + rpt; pos2 = prev2 - rpt; break;
            case LZB_OP_INS: rpt = prev2 - prev1; s->op[i] = LZB_OP_INS | (rpt << 2); pos1 = prev1; pos2 = prev2 - rpt; break;
            case LZB_OP_DEL: rpt = prev2 - prev1; s->op[i] = LZB_OP_DEL | (rpt << 2); pos1 = prev1 + rpt; pos2 = prev2; break;
        }
    }
    *p1 = pos1; *p2 = pos2;
    return 1;
}
static int32_t score_script(const lzb_scoreset* ss, const uint8_t* s1, const uint8_t* s2, const lzb_editscript* sc) {
    int32_t sim = 0;
    for (uint32_t k = 0; k < sc->len; k++) {
        uint32_t rpt = RPT(sc, k);
        if (!rpt) continue;
        switch (OP(sc, k)) {
            case LZB_OP_SUB: for (uint32_t j = 0; j < rpt; j++) sim += ss->sub[(uint32_t)s1[j] * 256 + s2[j]]; s1 += rpt; s2 += rpt; break;
            case LZB_OP_INS: sim -= ss->gapOpen + (int32_t)rpt * ss->gapExtend; s2 += rpt; break;
            case LZB_OP_DEL: sim -= ss->gapOpen + (int32_t)rpt * ss->gapExtend; s1 += rpt; break;
        }
    }
    return sim;
}

lzb_alignel* lzb_mirror_alignments(lzb_alignel* list, const lzb_seq* seq1, const lzb_seq* seq2, const lzb_scoreset* ss) {
    uint32_t seqLen = seq1->len;
    if (seq2->len != seqLen) lzb_die("internal error (for mirroring), sequence lengths differ %u vs %u", seqLen, seq2->len);
    int sameStrand = seq1->revCompFlags == seq2->revCompFlags;
    lzb_alignel* newList = NULL, *bTail = NULL, *aPrev = NULL, *aTail = NULL, *aNext, *b;
    for (lzb_alignel* a = list; a; a = aNext) {
        aPrev = aTail; aNext = a->next; aTail = a;
        uint32_t pos1 = a->beg1 - 1, end1 = a->end1, pos2 = a->beg2 - 1, end2 = a->end2;
        if (sameStrand) {
            b = calloc(1, sizeof *b);
            b->beg1 = pos2 + 1; b->end1 = end2; b->beg2 = pos1 + 1; b->end2 = end1; b->s = a->s;
            b->seq1 = a->seq1; b->seq2 = a->seq2; b->script = es_copy(a->script); es_mirror(b->script);
        } else {
            uint32_t inPos2 = pos2, inEnd2 = end2, invert1 = seqLen, invert2 = seqLen;
            pos2 = invert2 - inPos2; end2 = invert2 - inEnd2;       /* end2 < pos2 */
            int discard = pos1 == pos2;                             /* starts on the diagonal */
            if (!discard && end1 >= end2) {                         /* touches or crosses the diagonal */
                uint32_t x = pos1, y = pos2;
                int truncated = es_upper_truncate(a->script, &x, &y);
                if (truncated && x == POS_INF) discard = 1;
                else {
                    int overlap = 0;
                    if (truncated) {
                        int dontMirror = 0;
                        if (x < y || x > y + 1) {
                            fprintf(stderr, "WARNING.  Internal error in mirror_alignments().\n"
                                            "  An alignment crosses the main diagonal in an unexpected way.\n");
                            dontMirror = 1;
                        }
                        a->end1 = end1 = x; a->end2 = inEnd2 = invert2 - y; end2 = y;
                        if (dontMirror) continue;
                        if (x == y + 1) overlap = 1;
                    }
                    lzb_editscript* t = es_copy(a->script);
                    es_reverse(t); es_mirror(t);
                    if (overlap) es_trim_head(t, 1);
                    es_append(&a->script, t);
                    free(t);
                    es_overall(a->script, &x, &y);
                    a->end1 = pos1 + x; a->end2 = inPos2 + y;
                    a->s = score_script(ss, seq1->v + pos1, seq2->v + inPos2, a->script);
                    continue;                                        /* the mirror image now is the alignment's second half */
                }
            }
            if (discard) {
                free(a->script); free(a);
                if (!aPrev) { list = aNext; aTail = NULL; } else { aPrev->next = aNext; aTail = aPrev; }
                continue;
            }
            b = calloc(1, sizeof *b);
            b->beg1 = (invert2 - inEnd2) + 1; b->end1 = invert2 - inPos2;
            b->beg2 = (invert1 - end1) + 1;   b->end2 = invert1 - pos1;
            b->s = a->s; b->seq1 = a->seq1; b->seq2 = a->seq2;
            b->script = es_copy(a->script); es_reverse(b->script); es_mirror(b->script);
        }
        if (!bTail) newList = b; else bTail->next = b;
        bTail = b;
    }
    if (!aTail) list = newList; else aTail->next = newList;
    return list;
}
