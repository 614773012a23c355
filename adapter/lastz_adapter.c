/*
 * lastz_adapter.c -- the binding a lastz maintainer would add: the reference's OWN entry points for the hot path,
 * with the reference's own signatures, implemented over the C-ABI of include/lastz_b200.h.
 *
 *   build_seed_position_table   pos_table.h:230      (call sites lastz.c:1205,1211,1302)
 *   free_position_table         pos_table.h:245      (lastz.c:1300,1908,1969)
 *   seed_hit_search             seed_search.h:265    (lastz.c:3089-3123)
 *   reduce_to_points            gapped_extend.h:151  (lastz.c:3401)
 *   gapped_extend               gapped_extend.h:153  (lastz.c:3419)
 *
 * This file is compiled against the reference's headers where they lie (-I/root/reference/src; nothing of the
 * reference is copied here) and linked with the reference's own objects, in which those five symbols have been renamed
 * ref_* by objcopy (adapter/build.sh).  Everything else of lastz -- command line, sequence files, chaining, writers --
 * is the reference's code, unchanged.  Linked against oracle/liblzb_oracle.so this is a CPU-only check of the boundary
 * (tests/test_adapter.py); linked against lastz_b200/csrc/liblastz_b200.so it is lastz running its hot path on a B200.
 *
 * What the library does not implement the adapter refuses loudly (suicide): adaptive HSP thresholds, positional filters,
 * bandWidth, density filtering, half-weight / overweight / reverse-complement seeds.  (The simple, plain, recoverable and
 * twin hit processors, searchLimit and maxPairedBases are passed through.)  There is no fallback to the renamed originals.
 */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include "utilities.h"
#include "dna_utilities.h"
#include "sequences.h"
#include "seeds.h"
#include "pos_table.h"
#include "segment.h"
#include "edit_script.h"
#include "seed_search.h"
#include "gapped_extend.h"
#include "lastz.h"
#include "lastz_b200.h"

static lzb_ctx*    g_ctx;
static lzb_target* g_target;          /* lastz keeps one target position table at a time (lastz.c:660) */
static postable*   g_targetKey;
static lzb_seed    g_seed;
static lzb_query*  g_query;           /* the query strand of the last seed_hit_search: reused by the gapped stage */
static const u8*   g_queryBytes; static unspos g_queryLen;
static const void* g_scoringKey; static const void* g_maskedKey;
extern control* currParams;            /* lastz.c:327 (output.c reads it the same way) */

static lzb_ctx* ctx(void) {
    if (!g_ctx) {
        const char* dev = getenv("LZB_DEVICE");
        g_ctx = lzb_open(dev ? atoi(dev) : 0);
        if (!g_ctx) suicidef("%s", lzb_last_error());
    }
    return g_ctx;
}

/* scoreset.sub is score[256][256] (dna_utilities.h:176-212); the seed stage scores with the (masked) set in the hit
 * processor's info, the gapped stage with the set it is handed -- the library takes both at once */
static void use_scoring(scoreset* gappedSet, scoreset* maskedSet) {
    if (!gappedSet && !maskedSet) {                       /* plain seed hits carry no scoring set (lastz.c:2789): nothing is scored */
        static int32_t zero[256 * 256];
        if (g_scoringKey != NULL) return;                 /* some set is in place already: good enough for hits that are not scored */
        if (lzb_set_scoring(ctx(), zero, zero, 0, 0)) suicidef("%s", lzb_last_error());
        g_scoringKey = g_maskedKey = zero;
        return;
    }
    /* one call gives the library both sets; whichever stage asks first, the other set comes from the run's parameters */
    if (!gappedSet) gappedSet = (currParams && currParams->scoring) ? currParams->scoring : maskedSet;
    if (!maskedSet) maskedSet = (currParams && currParams->maskedScoring) ? currParams->maskedScoring : gappedSet;
    if (g_scoringKey == (const void*)gappedSet && g_maskedKey == (const void*)maskedSet) return;
    if (sizeof(score) != 4) suicide("the lastz_b200 adapter needs the integer-score build (score_type=I)");
    if (lzb_set_scoring(ctx(), (const int32_t*)&gappedSet->sub[0][0], (const int32_t*)&maskedSet->sub[0][0], gappedSet->gapOpen, gappedSet->gapExtend))
        suicidef("%s", lzb_last_error());
    g_scoringKey = gappedSet; g_maskedKey = maskedSet;
}

static void seed_to_lzb(const seed* s, lzb_seed* o) {
    if (s->type == 'R' || s->resolvingMask != 0) suicide("the lastz_b200 adapter does not support overweight seeds");
    if (s->isHalfweight) suicide("the lastz_b200 adapter does not support half-weight seeds");
    if (s->revComp) suicide("the lastz_b200 adapter does not support reverse-complement seed packing");
    if (s->numParts > LZB_MAX_SEED_PARTS) suicide("seed has too many shift/mask parts for lastz_b200");
    memset(o, 0, sizeof *o);
    o->length = s->length; o->weight = s->weight; o->numParts = s->numParts; o->withTrans = s->withTrans;
    for (int i = 0; i < s->numParts; i++) { o->shift[i] = s->shift[i]; o->mask[i] = s->mask[i]; }
    if (s->withTrans && s->transFlips)
        for (const u32* f = s->transFlips; *f != 0; f++) {
            if (o->numFlips >= LZB_MAX_SEED_FLIPS) suicide("seed has too many transition positions for lastz_b200");
            o->transFlips[o->numFlips++] = *f;
        }
}

/* ---- pos_table.h:230 ---- */
postable* build_seed_position_table(seq* s, unspos start, unspos end, const s8 upperCharToBits[], seed* hitSeed, u32 step) {
    if (s->fileType == seq_type_qdna) suicide("the lastz_b200 adapter does not support quantum DNA");
    /* the library classifies the bases of a sequence by the scoring sets in force when it is loaded */
    if (currParams) use_scoring(currParams->scoring, currParams->maskedScoring);
    seed_to_lzb(hitSeed, &g_seed);
    if (g_target) { lzb_target_free(g_target); g_target = NULL; }
    g_target = lzb_target_build(ctx(), s->v, (uint32_t)s->len, (uint32_t)start, (uint32_t)end, (const int8_t*)upperCharToBits, &g_seed, step);
    if (!g_target) suicidef("%s", lzb_last_error());
    /* lastz.c treats the table as opaque unless --maxwordcount / --masking / --stats look inside; it gets a zeroed
     * record that says which words the table covers */
    postable* pt = (postable*)zalloc_or_die("lastz_b200 adapter position table handle", sizeof(postable));
    pt->wordBits = hitSeed->weight; pt->start = start; pt->end = end; pt->step = step;
    g_targetKey = pt;
    return pt;
}

/* ---- pos_table.h:245 ---- */
void free_position_table(postable* pt) {
    if (pt == NULL) return;
    if (pt == g_targetKey) { lzb_target_free(g_target); g_target = NULL; g_targetKey = NULL; }
    free(pt);
}

/* ---- seed_search.h:265 ---- */
u64 seed_hit_search(seq* seq1, postable* pt, seq* seq2, unspos start, unspos end, int selfCompare,
                    const s8 charToBits[], seed* hitSeed, u32 searchLimit, u32 reportSearchLimit,
#ifdef densityFiltering
                    double maxDensity,
#endif
#ifndef forbidBandWidth
                    u32 bandWidth,
#endif
                    hitprocessor processor, void* processorInfo) {
    (void)hitSeed;
    if (pt != g_targetKey || !g_target) suicide("the lastz_b200 adapter was handed a position table it did not build");
    if (processor != process_for_simple_hit && processor != process_for_plain_hit && processor != process_for_recoverable_hit
     && processor != process_for_twin_hit)
        suicide("the lastz_b200 adapter was handed a hit processor it does not know");
#ifdef densityFiltering
    if (maxDensity != 0) suicide("the lastz_b200 adapter does not support --maxdensity");
#endif
#ifndef forbidBandWidth
    if (bandWidth != 0) suicide("the lastz_b200 adapter does not support --band");
#endif
    hitprocinfo* hp = (hitprocinfo*)processorInfo;          /* first member of every hit processor's info (seed_search.h:112-156) */
    if (hp->posFilter) suicide("the lastz_b200 adapter does not support positional hit filters");
    if (hp->minMatches >= 0) suicide("the lastz_b200 adapter does not support --filter=<transv>,<matches>");
    if (processor != process_for_plain_hit && hp->gfExtend != gfexNoExtend && hp->hspThreshold.t != 'S')
        suicide("the lastz_b200 adapter does not support adaptive HSP thresholds");
    use_scoring(NULL, hp->scoring);
    if (g_query) { lzb_query_free(g_query); g_query = NULL; }
    g_query = lzb_query_load(ctx(), seq2->v, (uint32_t)seq2->len);
    if (!g_query) suicidef("%s", lzb_last_error());
    g_queryBytes = seq2->v; g_queryLen = seq2->len;
    lzb_seed_params sp; memset(&sp, 0, sizeof sp);
    sp.start = (uint32_t)start; sp.end = (uint32_t)end;
    sp.plainHits = processor == process_for_plain_hit;
    sp.recoverSeeds = processor == process_for_recoverable_hit;   /* --recoverseeds; the reference merges the table itself (lastz.c:3296) */
    if (processor == process_for_twin_hit) {                      /* --twins: hitproctwin = hitprocinfo + the two spans (seed_search.h:144-156) */
        hitproctwin* tw = (hitproctwin*)processorInfo;
        sp.twinMinSpan = (int32_t)tw->minSpan; sp.twinMaxSpan = (int32_t)tw->maxSpan;
        sp.seedQueueSize = (int32_t)seedHitQueueSize;             /* diag_hash.h:108 */
    }
    sp.gfExtend = sp.plainHits ? LZB_GFEX_NONE
                : hp->gfExtend == gfexXDrop ? LZB_GFEX_XDROP
                : hp->gfExtend == gfexExact ? LZB_GFEX_EXACT
                : hp->gfExtend >= gfexMismatch_min ? LZB_GFEX_MISMATCH : LZB_GFEX_NONE;
    sp.gfMismatches = hp->gfExtend >= gfexMismatch_min ? hp->gfExtend : 0;
    sp.xDrop = hp->xDrop; sp.hspThreshold = hp->hspThreshold.s; sp.entropy = hp->entropicHsp;
    sp.hashBits = 16;                                        /* diag_hash_size 65536 in the stock build (diag_hash.h:43-59) */
    sp.selfCompare = selfCompare;
    sp.sameStrand = selfCompare && seq1->revCompFlags == seq2->revCompFlags;
    sp.strandId = seq2->revCompFlags;
    sp.searchLimit = searchLimit;                            /* --queryhsplimit: the table ends with the query position that took it past the limit */
    lzb_segment* segs = NULL; uint64_t n = 0; u64 bases = 0;
    if (lzb_seed_hit_search(ctx(), g_target, g_query, &g_seed, (const int8_t*)charToBits, &sp, &segs, &n, NULL)) suicidef("%s", lzb_last_error());
    /* the library returns the table in discovery order; the reporter sees what process_for_simple_hit shows it
     * (seed_search.c:1183): the position just past the HSP's end in both sequences, its length and its score */
    /* an x-drop extension that yields an HSP marks the table being collected as scored (seed_search.c:2951-2952) */
    if (n > 0 && sp.gfExtend == LZB_GFEX_XDROP && hp->anchors != NULL) (*(hp->anchors))->haveScores = 1;
    for (uint64_t i = 0; i < n; i++)
        bases += (*hp->reporter)(hp->reporterInfo, (unspos)segs[i].pos1 + segs[i].length, (unspos)segs[i].pos2 + segs[i].length, segs[i].length, segs[i].s);
    lzb_free(segs);
    if (searchLimit > 0 && n > searchLimit) {                /* warn_for_search_limit seed_search.c:3790-3815 (a static function of the replaced file) */
        seed_search_dbgSearchLimitExceeded++;
        if (reportSearchLimit != 0)
            fprintf(stderr, "WARNING. Query \"%s\" contains more than %u HSPs.\n", seq2->useFullNames ? seq2->header : seq2->shortHeader, (unsigned)reportSearchLimit);
    }
    return bases;
}

/* the gapped stage alone (--anchors=, --segments=: lastz.c:1186 builds no position table then): the library still wants
 * the target resident, so it gets one with an index of a single word */
static const u8* g_bareTargetBytes;
static void need_target(seq* seq1) {
    if (g_target && (g_targetKey != NULL || g_bareTargetBytes == seq1->v)) return;
    if (g_target) lzb_target_free(g_target);
    if (currParams) use_scoring(currParams->scoring, currParams->maskedScoring);
    lzb_seed one; memset(&one, 0, sizeof one);
    one.length = 12; one.weight = 24; one.numParts = 1; one.mask[0] = 0xFFFFFF;
    int8_t ctb[256]; memset(ctb, -1, sizeof ctb); ctb['A'] = ctb['a'] = 0; ctb['C'] = ctb['c'] = 1; ctb['G'] = ctb['g'] = 2; ctb['T'] = ctb['t'] = 3;
    const uint32_t len = (uint32_t)seq1->len;
    g_target = lzb_target_build(ctx(), seq1->v, len, 0, 0, ctb, &one, len > 12 ? len : 13);
    if (!g_target) suicidef("%s", lzb_last_error());
    g_targetKey = NULL; g_bareTargetBytes = seq1->v;
}

static void need_query(seq* seq2) {
    if (g_query && g_queryBytes == seq2->v && g_queryLen == seq2->len) return;
    if (g_query) lzb_query_free(g_query);
    g_query = lzb_query_load(ctx(), seq2->v, (uint32_t)seq2->len);
    if (!g_query) suicidef("%s", lzb_last_error());
    g_queryBytes = seq2->v; g_queryLen = seq2->len;
}

/* ---- gapped_extend.h:151 ---- */
void reduce_to_points(seq* seq1, seq* seq2, scoreset* scoring, segtable* anchors) {
    use_scoring(scoring, NULL);
    need_target(seq1);
    need_query(seq2);
    if (sizeof(segment) != sizeof(lzb_segment)) suicide("segment layout differs from lzb_segment");
    if (lzb_reduce_to_points(ctx(), g_target, g_query, (lzb_segment*)anchors->seg, anchors->len)) suicidef("%s", lzb_last_error());
}

/* ---- gapped_extend.h:153 ---- */
alignel* gapped_extend(seq* seq1, u8* rev1, seq* seq2, u8* rev2, int inhibitTrivial, scoreset* scoring, segtable* anchors, tback* tb,
                       int allBounds, score yDrop, int trimToPeak, sthresh scoreThresh, u64 maxPairedBases, int overlyPairedWarn, int overlyPairedKeep) {
    (void)rev1; (void)rev2;                                            /* the reversed copies and the traceback live on the device */
    if (scoreThresh.t != 'S') suicide("the lastz_b200 adapter does not support adaptive gapped thresholds");
    use_scoring(scoring, NULL);
    need_target(seq1);
    need_query(seq2);
    lzb_gapped_params gp; memset(&gp, 0, sizeof gp);
    gp.yDrop = yDrop; gp.trimToPeak = trimToPeak; gp.scoreThreshold = scoreThresh.s;
    gp.allBounds = allBounds; gp.inhibitTrivial = inhibitTrivial;
    gp.identityCheck = seq1->revCompFlags == seq2->revCompFlags;      /* identical_sequences gapped_extend.c:1905 */
    gp.tracebackBytes = tb->size + 8 - 1;                              /* new_traceback gapped_extend.c:2272-2290: size = bytes - sizeof header + 1 */
    gp.speculation = 256;
    gp.maxPairedBases = maxPairedBases; gp.overlyPairedKeep = overlyPairedKeep;       /* --querydepth= (gapped_extend.c:1444-1459) */
    lzb_alignel* list = NULL; lzb_gapped_stats gst;
    if (sizeof(alignel) != sizeof(lzb_alignel)) suicide("alignel layout differs from lzb_alignel");
    if (lzb_gapped_extend(ctx(), g_target, g_query, seq1->v, seq2->v, (lzb_segment*)anchors->seg, anchors->len, &gp, &list, &gst)) suicidef("%s", lzb_last_error());
    if (gst.overlyPaired && overlyPairedWarn)                          /* (the reference's own warning is a static function of the replaced file) */
        fprintf(stderr, "WARNING. Query %s (%c strand) contains more than %llu paired bases.\n",
                seq2->partition.p != NULL ? "seq2" : seq2->useFullNames ? seq2->header : seq2->shortHeader,
                (seq2->revCompFlags & rcf_rev) == 0 ? '+' : '-', (unsigned long long)maxPairedBases);
    return (alignel*)list;                                             /* same 64-byte records; free_align_list (edit_script.c:53) frees them */
}
