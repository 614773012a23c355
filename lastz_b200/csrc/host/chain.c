/*
 * chain.c -- best-chain reduction of an HSP table (--chain, C=1, C=2).  Host C, between the seed
 * stage and the gapped stage, as in finish_one_strand (lastz.c:3341-3358).
 *
 * Reference: reduce_to_chain chain.c:497-606, build_kd_tree :623-681, partition_segments :807-869,
 * best_predecessor :921-990, propagate_max_score :1006-1021, chain_connect_penalty lastz.c:3687-3757.
 * Result = the highest-scoring chain (each segment starts strictly before the next in both
 * sequences).  Chains of equal score are possible, so the choice depends on the ORDER in which
 * candidate predecessors are met: the first one with a strictly better contribution wins
 * (chain.c:952).  That order is the K-d tree's, so the tree is built (median-of-three partition,
 * buckets of 3, axes alternating between seq-2 position and diagonal) and walked (upper son first
 * on the seq-2 axis, nearer son first on the diagonal axis) the same way.  The reference passes
 * (lowerBound, axis) to the recursive calls on the seq-2 axis in swapped order (chain.c:963-964);
 * that affects only pruning strength for ordinary penalties, and is reproduced literally so the walk
 * is identical for any penalties.
 */
#include <stdlib.h>
#include <string.h>
#include "lzb_host.h"

typedef struct {
    int leaf; uint32_t lo, hi;          /* leaf: perm[lo..hi];  inner: hi = index of the cut element */
    int64_t cut; double best;           /* cut value on this node's axis; best chain score below this node */
    int loSon, hiSon;
} knode;

typedef struct {
    lzb_segment* seg; uint32_t* perm; double* chainScore;
    knode* nodes; int nnodes, capnodes;
    int32_t diagPen, antiPen, scale, subAA;
    /* the segment being added */
    const lzb_segment* q; uint32_t x, y; int64_t diag;
} kd;

static int by_pos1(const void* A, const void* B) {            /* qSegmentsByPos1 segment.c:1657 */
    const lzb_segment* a = A; const lzb_segment* b = B;
    if (a->pos1 != b->pos1) return a->pos1 < b->pos1 ? -1 : 1;
    if (a->length != b->length) return a->length < b->length ? -1 : 1;
    if (a->pos2 != b->pos2) return a->pos2 < b->pos2 ? -1 : 1;
    if (a->id != b->id) return a->id < b->id ? -1 : 1;
    if (a->s != b->s) return a->s < b->s ? -1 : 1;
    return 0;
}

static int64_t proj(const kd* k, uint32_t i, int axis) {
    const lzb_segment* s = &k->seg[k->perm[i]];
    return axis == 0 ? (int64_t)s->pos1 - (int64_t)s->pos2 : (int64_t)s->pos2;
}
static void swp(kd* k, uint32_t a, uint32_t b) { uint32_t t = k->perm[a]; k->perm[a] = k->perm[b]; k->perm[b] = t; }

/* partition_segments chain.c:807-869: median of (first, middle, last) to the front as pivot, two
 * inward scans (<= pivot from the left, > pivot from the right) with swaps, the final swap undone,
 * pivot dropped at the meeting point; if everything was <= pivot the range is retried without its
 * last element */
static uint32_t split(kd* k, uint32_t lo, uint32_t hi, int axis) {
    for (;;) {
        uint32_t mid = (lo + hi) / 2;
        int64_t a = proj(k, lo, axis), b = proj(k, mid, axis), c = proj(k, hi, axis), pivot;
        if ((a <= b && b <= c) || (c <= b && b <= a)) { swp(k, lo, mid); pivot = b; }
        else if ((a <= c && c <= b) || (b <= c && c <= a)) { swp(k, lo, hi); pivot = c; }
        else pivot = a;
        uint32_t i = lo, j = hi + 1;
        while (i < j) {
            for (i++; i <= hi && proj(k, i, axis) <= pivot; i++) ;
            for (j--; j >= lo && proj(k, j, axis) > pivot; j--) ;
            swp(k, i, j);
        }
        swp(k, i, j);
        swp(k, lo, j);
        if (j < hi) return j;
        if (hi - lo == 2) return hi - 1;
        hi--;
    }
}

static int build(kd* k, uint32_t lo, uint32_t hi, int axis) {
    if (k->nnodes == k->capnodes) { k->capnodes = k->capnodes * 2 + 64; k->nodes = realloc(k->nodes, (size_t)k->capnodes * sizeof(knode)); }
    int me = k->nnodes++;
    knode nd; memset(&nd, 0, sizeof nd);
    if (hi + 1 - lo <= 3) { nd.leaf = 1; nd.lo = lo; nd.hi = hi; nd.loSon = nd.hiSon = -1; k->nodes[me] = nd; return me; }
    uint32_t m = split(k, lo, hi, axis);
    nd.leaf = 0; nd.cut = proj(k, m, axis); nd.hi = m;
    k->nodes[me] = nd;
    int l = build(k, lo, m, 1 - axis);
    int h = build(k, m + 1, hi, 1 - axis);
    k->nodes[me].loSon = l; k->nodes[me].hiSon = h;
    return me;
}

/* chain_connect_penalty lastz.c:3687-3757 */
static int32_t connect_penalty(const kd* k, const lzb_segment* s1, const lzb_segment* s2) {
    uint32_t xEnd = s1->pos1 + s1->length - 1, yEnd = s1->pos2 + s1->length - 1;
    int32_t d1 = (int32_t)(s1->pos1 - s1->pos2), d2 = (int32_t)(s2->pos1 - s2->pos2);
    int32_t dd = d2 - d1, numSubs;
    if (dd >= 0) numSubs = (int32_t)s2->pos2 - (int32_t)yEnd - 1;
    else { numSubs = (int32_t)s2->pos1 - (int32_t)xEnd - 1; dd = -dd; }
    double pen = (double)(dd * k->diagPen);                   /* int*int, then double, as in the reference */
    if (numSubs >= 0) pen += (double)(numSubs * k->antiPen);
    else pen += (double)((-numSubs) * k->scale * k->subAA);
    if (pen > 2147483647.0) return 0x7FFFFFFF;
    return (int32_t)pen;
}

typedef struct { uint32_t num; double contrib; } pred;
#define NO_PRED 0xFFFFFFFFu

/* best_predecessor chain.c:921-990 */
static pred best_pred(kd* k, int node, int axis, double lowerBound, pred bp) {
    const knode* nd = &k->nodes[node];
    if (bp.contrib >= nd->best - lowerBound) return bp;
    if (nd->leaf) {
        for (uint32_t i = nd->lo; i <= nd->hi; i++) {
            uint32_t j = k->perm[i];
            const lzb_segment* s = &k->seg[j];
            if (s->pos1 >= k->x || s->pos2 >= k->y) continue;
            double v = k->chainScore[j] - (double)connect_penalty(k, s, k->q);
            if (v > bp.contrib) { bp.contrib = v; bp.num = j; }
        }
    } else if (axis == 1) {
        /* (lowerBound, axis) swapped exactly as at chain.c:963-964 */
        if ((int64_t)k->y >= nd->cut) bp = best_pred(k, nd->hiSon, (int)lowerBound, (double)(1 - axis), bp);
        bp = best_pred(k, nd->loSon, (int)lowerBound, (double)(1 - axis), bp);
    } else {
        double diff = (double)(k->diag - nd->cut);
        if (diff >= 0) {
            bp = best_pred(k, nd->hiSon, 1 - axis, lowerBound, bp);
            bp = best_pred(k, nd->loSon, 1 - axis, diff * k->diagPen, bp);
        } else {
            bp = best_pred(k, nd->loSon, 1 - axis, lowerBound, bp);
            bp = best_pred(k, nd->hiSon, 1 - axis, -diff * k->antiPen, bp);
        }
    }
    return bp;
}

/* reduce_to_chain chain.c:497-606; returns the chain score, rewrites segs/n to the chain */
int32_t lzb_reduce_to_chain(lzb_segment* segs, uint64_t* pn, int32_t diagPen, int32_t antiPen, int32_t scale, int32_t subAA) {
    uint32_t n = (uint32_t)*pn;
    if (n == 0) return 0;
    qsort(segs, n, sizeof *segs, by_pos1);
    kd k; memset(&k, 0, sizeof k);
    k.seg = segs; k.diagPen = diagPen; k.antiPen = antiPen; k.scale = scale; k.subAA = subAA;
    k.perm = malloc((size_t)n * 4); k.chainScore = calloc(n, sizeof(double));
    uint32_t* inv = malloc((size_t)n * 4); uint32_t* link = malloc((size_t)n * 4);
    for (uint32_t i = 0; i < n; i++) k.perm[i] = i;
    int root = build(&k, 0, n - 1, 1);
    for (uint32_t i = 0; i < n; i++) inv[k.perm[i]] = i;
    double best = 0; uint32_t bestEnd = NO_PRED;
    for (uint32_t i = 0; i < n; i++) {
        k.q = &segs[i]; k.x = segs[i].pos1; k.y = segs[i].pos2; k.diag = (int64_t)(int32_t)(k.x - k.y);
        pred bp = { NO_PRED, 0 };
        bp = best_pred(&k, root, 1, 0, bp);
        k.chainScore[i] = (double)segs[i].s * (double)scale + bp.contrib;
        if (k.chainScore[i] > best) { best = k.chainScore[i]; bestEnd = i; }
        link[i] = bp.num;
        /* propagate_max_score chain.c:1006-1021 */
        for (int nd = root; nd >= 0;) {
            knode* p = &k.nodes[nd];
            if (k.chainScore[i] > p->best) p->best = k.chainScore[i];
            nd = (inv[i] <= p->hi) ? p->loSon : p->hiSon;
        }
    }
    for (uint32_t i = 0; i < n; i++) segs[i].filter = 1;
    for (uint32_t i = bestEnd; i != NO_PRED; i = link[i]) segs[i].filter = 0;
    uint64_t m = 0;
    for (uint32_t i = 0; i < n; i++) if (!segs[i].filter) segs[m++] = segs[i];     /* filter_marked_segments segment.c:1617 */
    *pn = m;
    best = best / scale + 0.5;
    free(k.perm); free(k.chainScore); free(k.nodes); free(inv); free(link);
    return best > 2147483647.0 ? 0x7FFFFFFF : (int32_t)best;
}

/* try_reduce_to_chain chain.c:224-392: with [multi] sequences the segments are chained separately for every pair of
 * partitions (a segment never spans two), batches in partition order; the caller's sort by pos1 (lastz.c:3352) follows */
static uint32_t partition_index(const lzb_seq* s, uint32_t pos0) {
    if (s->npart == 0) return 0;
    uint32_t lo = 0, hi = s->npart;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) / 2; if (s->part[mid].sepBefore <= pos0) lo = mid; else hi = mid; }
    return lo;
}
typedef struct { uint32_t p1, p2; lzb_segment g; } batched;
static int by_batch(const void* pa, const void* pb) {
    const batched* a = pa; const batched* b = pb;
    if (a->p1 != b->p1) return a->p1 < b->p1 ? -1 : 1;
    if (a->p2 != b->p2) return a->p2 < b->p2 ? -1 : 1;
    return by_pos1(&a->g, &b->g);
}
int32_t lzb_reduce_to_chains(const lzb_seq* s1, const lzb_seq* s2, lzb_segment* segs, uint64_t* pn,
                             int32_t diagPen, int32_t antiPen, int32_t scale, int32_t subAA) {
    int32_t best = 0;
    if (s1->npart == 0 && s2->npart == 0) best = lzb_reduce_to_chain(segs, pn, diagPen, antiPen, scale, subAA);
    else if (*pn) {
        uint64_t n = *pn, kept = 0;
        batched* b = malloc(n * sizeof *b);
        for (uint64_t i = 0; i < n; i++) { b[i].p1 = partition_index(s1, segs[i].pos1); b[i].p2 = partition_index(s2, segs[i].pos2); b[i].g = segs[i]; }
        qsort(b, n, sizeof *b, by_batch);
        for (uint64_t lo = 0; lo < n;) {
            uint64_t hi = lo; while (hi < n && b[hi].p1 == b[lo].p1 && b[hi].p2 == b[lo].p2) hi++;
            uint64_t m = hi - lo;
            for (uint64_t i = 0; i < m; i++) segs[kept + i] = b[lo + i].g;
            int32_t sc = lzb_reduce_to_chain(segs + kept, &m, diagPen, antiPen, scale, subAA);
            if (sc > best) best = sc;
            kept += m; lo = hi;
        }
        free(b); *pn = kept;
    }
    if (*pn) qsort(segs, *pn, sizeof *segs, by_pos1);
    return best;
}
