/*
 * adaptive.c -- the HSP table behind an adaptive threshold (K=top<N>% or K=top<bases>).
 *
 * The reference keeps "the best HSPs that together cover at least <limit> bases": add_segment
 * (segment.c:981-1180) appends until the covered bases reach the limit, then turns the table into a
 * binary min-heap on score in which every node also carries the bases of the equal-scoring part of
 * its own subtree (scoreCov, record_tie_score segment.c:1290-1330), so that a whole group of tied
 * lowest scores is dropped together, and only while the rest still meets the limit
 * (remove_root segment.c:1186-1250).  The table's ARRAY ORDER is what later stages and the
 * writers see, so the heap is reproduced move for move here; lzb_segment carries scoreCov at the
 * reference's offset for this purpose.
 *
 * Host-side plumbing on the reporter side of the hot path (collect_hsps lastz.c:3991): the segments
 * fed in come from lzb_seed_hit_search in discovery order.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "lzb_host.h"

void lzb_hsptable_init(lzb_hsptable* t, uint64_t coverageLimit) {
    memset(t, 0, sizeof *t);
    t->limit = coverageLimit; t->lowScore = INT32_MIN;          /* worstPossibleScore, new_segment_table segment.c:127 */
}
void lzb_hsptable_free(lzb_hsptable* t) { free(t->seg); memset(t, 0, sizeof *t); }

/* bases of the tied-score part of the subtree under ix; 1 when the stored value changed */
static int tie_cover(lzb_hsptable* t, uint32_t ix) {
    lzb_segment* g = &t->seg[ix];
    uint64_t cov = g->length;
    for (uint32_t kid = 2 * ix + 1; kid <= 2 * ix + 2 && kid < t->len; kid++)
        if (t->seg[kid].s == g->s) cov += t->seg[kid].scoreCov;
    if (cov == g->scoreCov) return 0;
    g->scoreCov = cov; return 1;
}

static int by_rising_score(const void* pa, const void* pb) {    /* qSegmentsByIncreasingScore segment.c:1774 */
    const lzb_segment* a = pa; const lzb_segment* b = pb;
    if (a->s != b->s) return a->s < b->s ? -1 : 1;
    if (a->length != b->length) return a->length < b->length ? -1 : 1;
    if (a->pos2 != b->pos2) return a->pos2 < b->pos2 ? -1 : 1;
    if (a->pos1 != b->pos1) return a->pos1 < b->pos1 ? -1 : 1;
    if (a->id != b->id) return a->id < b->id ? -1 : 1;
    return 0;
}

/* drop the root: the last element takes its place and sinks (remove_root segment.c:1186) */
static void drop_lowest(lzb_hsptable* t) {
    t->coverage -= t->seg[0].length;
    if (t->len <= 1) { t->len = 0; return; }
    lzb_segment moved = t->seg[--t->len];
    if (t->len == 1) { t->seg[0] = moved; return; }
    for (uint32_t ix = (t->len - 1) / 2; ix > 0; ix = (ix - 1) / 2)          /* the old parent chain of the detached leaf */
        if (!tie_cover(t, ix)) break;
    uint32_t ix = 0;
    for (;;) {
        uint32_t kid = 2 * ix + 1;
        if (kid >= t->len) break;
        if (kid + 1 < t->len && t->seg[kid + 1].s < t->seg[kid].s) kid++;
        if (moved.s <= t->seg[kid].s) break;
        t->seg[ix] = t->seg[kid]; ix = kid;
    }
    t->seg[ix] = moved;
    for (; ix > 0; ix = (ix - 1) / 2) tie_cover(t, ix);
    tie_cover(t, 0);
}

void lzb_hsptable_add(lzb_hsptable* t, const lzb_segment* g) {
    if (t->limit != 0 && t->coverage >= t->limit && t->len > 0 && g->s < t->lowScore) return;
    if (t->len >= t->cap) {
        t->cap = t->cap + 100 + t->cap / 3;
        t->seg = realloc(t->seg, (size_t)t->cap * sizeof *t->seg);
        if (!t->seg) lzb_die("out of memory for %u HSPs", t->cap);
    }
    lzb_segment* n = &t->seg[t->len++];
    *n = *g; n->filter = 0; n->scoreCov = g->length;
    t->coverage += g->length;
    if (t->len == 1 || g->s < t->lowScore) t->lowScore = g->s;
    if (t->limit == 0 || t->coverage < t->limit) return;         /* still a plain list */

    if (t->coverage - g->length < t->limit) {                    /* the limit is met for the first time: list -> heap */
        qsort(t->seg, t->len, sizeof *t->seg, by_rising_score);
        for (uint32_t ix = t->len; ix-- > 0;) tie_cover(t, ix);
    } else {                                                     /* sift the newcomer up, keeping the tie sums right */
        uint32_t ix = t->len - 1; int tied = 0;
        while (ix > 0) {
            uint32_t up = (ix - 1) / 2;
            if (t->seg[ix].s >= t->seg[up].s) { tied = t->seg[ix].s == t->seg[up].s; break; }
            lzb_segment tmp = t->seg[ix]; t->seg[ix] = t->seg[up]; t->seg[up] = tmp;
            tie_cover(t, ix);
            ix = up;
        }
        tie_cover(t, ix);
        if (tied) {
            int stopped = 0;
            for (ix = (ix - 1) / 2; ix > 0; ix = (ix - 1) / 2)
                if (!tie_cover(t, ix)) { stopped = 1; break; }
            if (!stopped) tie_cover(t, 0);
        }
    }
    /* drop whole groups of tied lowest scores while what remains still meets the limit (segment.c:1133-1160) */
    if (t->coverage - t->seg[0].scoreCov < t->limit) return;
    while (t->coverage - t->seg[0].scoreCov >= t->limit) {
        int32_t s = t->seg[0].s;
        while (t->seg[0].s == s) drop_lowest(t);
    }
    t->lowScore = t->seg[0].s;
}

/* keep the segments whose id is `id`, in place and in array order; the others go to `rest` in array order
 * (split_segment_table segment.c:1352, used by split_anchors lastz.c:3618 to separate the two strands) */
void lzb_hsptable_split(lzb_hsptable* t, int id, lzb_hsptable* rest) {
    uint64_t cov = 0; int32_t low = INT32_MIN; uint32_t kept = 0;
    for (uint32_t k = 0; k < t->len; k++) {
        lzb_segment* g = &t->seg[k];
        if (g->id != id) { lzb_hsptable_add(rest, g); continue; }
        cov += g->length;
        if (kept == 0 || g->s < low) low = g->s;
        t->seg[kept++] = *g;
    }
    t->len = kept; t->coverage = cov; t->lowScore = low;
}

/* entropy of the matched bases of an ungapped segment, as a fraction of 2 bits (compute_entropy
 * dna_utilities.c:2892, upper case only); 1.0 when fewer than 20 bases match */
double lzb_hsp_entropy(const uint8_t* s, const uint8_t* t, uint32_t len) {
    int n[4] = { 0, 0, 0, 0 };
    for (uint32_t k = 0; k < len; k++) {
        if (s[k] != t[k]) continue;
        switch (s[k]) { case 'A': n[0]++; break; case 'C': n[1]++; break; case 'G': n[2]++; break; case 'T': n[3]++; break; default: break; }
    }
    if (n[0] + n[1] + n[2] + n[3] < 20) return 1.0;
    double h = 0;
    double pA = (double)n[0] / (double)(int)len, pC = (double)n[1] / (double)(int)len, pG = (double)n[2] / (double)(int)len, pT = (double)n[3] / (double)(int)len;
    double qA = n[0] ? log(pA) : 0.0, qC = n[1] ? log(pC) : 0.0, qG = n[2] ? log(pG) : 0.0, qT = n[3] ? log(pT) : 0.0;
    h = -(pA * qA + pC * qC + pG * qG + pT * qT) / log(4.0);
    return h;
}
