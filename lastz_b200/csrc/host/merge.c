/* merge.c -- merge_segments (segment.c:1527-1603): HSPs that share positions on one diagonal become one.
 *
 * The reference merges its anchor table whenever the hit processor may report overlapping HSPs (the recoverable and
 * twin processors, lastz.c:2811) and whenever anchors come from a file (lastz.c:757), before chaining, filtering and
 * gapped extension (finish_one_strand lastz.c:3296).  Segments are ordered by diagonal, then start (then length, id,
 * score: qSegmentsByDiag segment.c:1819-1847, a total order on everything the result depends on); a run of segments on
 * one diagonal in which each starts before the running end is replaced by their union with the highest score of the
 * run.  Adjoining segments are not merged.  Results are written over the front of the sorted table, so the fields that
 * are not recomputed (id, hspId, scoreCov, filter) are whatever the sort left in that slot, as in the reference. */
#include <stdlib.h>
#include "lzb_host.h"

static int by_diagonal(const void* pa, const void* pb) {
    const lzb_segment* a = pa; const lzb_segment* b = pb;
    const int64_t da = (int64_t)a->pos1 - (int64_t)a->pos2, db = (int64_t)b->pos1 - (int64_t)b->pos2;
    if (da != db) return da < db ? -1 : 1;
    if (a->pos2 != b->pos2) return a->pos2 < b->pos2 ? -1 : 1;
    if (a->length != b->length) return a->length < b->length ? -1 : 1;
    if (a->id != b->id) return a->id < b->id ? -1 : 1;
    if (a->s != b->s) return a->s < b->s ? -1 : 1;
    return 0;
}

void lzb_merge_segments(lzb_segment* seg, uint64_t* n) {
    if (*n < 2) return;
    qsort(seg, *n, sizeof *seg, by_diagonal);
    uint64_t out = 0;
    uint32_t pos2 = seg[0].pos2, end2 = pos2 + seg[0].length;
    int64_t diag = (int64_t)seg[0].pos1 - (int64_t)pos2;
    int32_t s = seg[0].s;
    for (uint64_t k = 1; k <= *n; k++) {
        if (k < *n) {
            const int64_t d = (int64_t)seg[k].pos1 - (int64_t)seg[k].pos2;
            if (d == diag && seg[k].pos2 < end2) {                      /* overlaps the run: widen it */
                const uint32_t e = seg[k].pos2 + seg[k].length;
                if (e > end2) end2 = e;
                if (seg[k].s > s) s = seg[k].s;
                continue;
            }
        }
        lzb_segment* g = &seg[out++];                                   /* (out <= k: never ahead of the reader) */
        g->pos1 = (uint32_t)(diag + pos2); g->pos2 = pos2; g->length = end2 - pos2; g->s = s;
        if (k < *n) { pos2 = seg[k].pos2; end2 = pos2 + seg[k].length; diag = (int64_t)seg[k].pos1 - (int64_t)pos2; s = seg[k].s; }
    }
    *n = out;
}
