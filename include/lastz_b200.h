/*
 * lastz_b200.h -- C-ABI drop-in boundary for LASTZ's seed-and-extend hot path on B200.
 *
 * Plain C, plain pointers and sizes, no C++/torch types.  Every entry point below replaces one
 * reference interface (cited as file:line relative to the reference tree, lastz 1.04.58).  The
 * reference has no plugin/FFI layer; its boundary is the C function surface exported by
 * pos_table.c, diag_hash.c, seed_search.c and gapped_extend.c.  Those functions take the
 * reference-internal aggregates `seq`, `scoreset`, `seed`, `postable`; only a handful of their
 * fields matter to the path, so the ABI takes those fields flat and returns reference-layout
 * records (48-byte `segment`, `alignel` + `editscript`).  INTEGRATION.md shows the adapter a
 * reference maintainer would write around each call site.
 *
 * Two libraries implement this header:
 *   lastz_b200/csrc/liblastz_b200.so   the product: hand-written sm_100a kernels, no CPU fallback
 *   oracle/liblzb_oracle.so            TEST INFRASTRUCTURE: CPU restatement of the reference
 * lzb_backend() tells them apart.
 *
 * Error convention: the reference prints "FAILURE: ..." and exit(1)s (utilities.c:1859-1885).  A
 * library cannot exit on behalf of its caller, so every call returns 0 on success / NULL on
 * failure and leaves the message in lzb_last_error(); the lastz_b200 CLI turns that into the
 * reference's FAILURE + exit(1).
 *
 * Threading: like the reference (seed_search.c:364-365) a context is not re-entrant; use one
 * context per host thread / per GPU.
 */
#ifndef LASTZ_B200_H
#define LASTZ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- records crossing the boundary (layouts pinned by static asserts in the sources) ---- */

/* segment.h:46-64 -- 48 bytes.  pos1/pos2 are 0-based starts; after lzb_reduce_to_points
 * length is 0 and pos1/pos2 are the anchor point. */
typedef struct lzb_segment {
    uint64_t hspId;      /* @0  */
    uint32_t pos1;       /* @8  */
    uint32_t pos2;       /* @12 */
    uint32_t length;     /* @16 */
    int32_t  s;          /* @20 */
    int32_t  id;         /* @24  strand flags of the query (collect_hsps lastz.c:3991) */
    uint32_t pad0;       /* @28 */
    uint64_t scoreCov;   /* @32 */
    int32_t  filter;     /* @40 */
    uint32_t pad1;       /* @44 */
} lzb_segment;

/* edit_script.h:55-61.  op = code | (repeat << 2); codes ins=1 del=2 sub=3 (edit_script.h:44-50) */
typedef struct lzb_editscript {
    uint32_t size;       /* entries allocated for op[] */
    uint32_t len;        /* entries used */
    uint32_t tailOp;     /* most recent operation added */
    uint32_t op[1];      /* @12, variable length */
} lzb_editscript;

#define LZB_OP_INS 1u
#define LZB_OP_DEL 2u
#define LZB_OP_SUB 3u

/* edit_script.h:30-41 -- 64 bytes.  beg/end are 1-based, inclusive. */
typedef struct lzb_alignel {
    struct lzb_alignel* next;   /* @0  */
    int32_t  isTrivial;         /* @8  */
    uint32_t beg1, beg2;        /* @12 @16 */
    uint32_t end1, end2;        /* @20 @24 */
    int32_t  s;                 /* @28 */
    const uint8_t* seq1;        /* @32 */
    const uint8_t* seq2;        /* @40 */
    lzb_editscript* script;     /* @48 */
    uint64_t hspId;             /* @56 */
} lzb_alignel;

/* seeds.h:37-76, the fields apply_seed (seeds.c:1335) and private_hit_search
 * (seed_search.c:464) read.  Match ('1') positions contribute both bits of a base, transition ('T') positions
 * the purine/pyrimidine bit; transFlips[] holds the packed single-bit masks in the reference's order, the
 * rightmost seed position first (seeds.c:165,603-613: built with maintainFlippedBitOrder). */
#define LZB_MAX_SEED_PARTS 32
#define LZB_MAX_SEED_FLIPS 32
typedef struct lzb_seed {
    int32_t  length;                         /* seed span in bases (<=31) */
    int32_t  weight;                         /* index bits (<=28, pos_table.c:1057) */
    int32_t  numParts;
    int32_t  withTrans;                      /* 0, 1 or 2 transitions allowed */
    int32_t  numFlips;
    int32_t  shift[LZB_MAX_SEED_PARTS];
    uint32_t mask[LZB_MAX_SEED_PARTS];
    uint32_t transFlips[LZB_MAX_SEED_FLIPS];
} lzb_seed;

/* parameters of the seed stage: seed_hit_search (seed_search.h:265-276) arguments plus the
 * hitprocinfo fields (seed_search.h:112-156) process_for_simple_hit / xdrop_extend_seed_hit use */
#define LZB_GFEX_NONE  0                      /* --nogfextend: raw hits */
#define LZB_GFEX_XDROP 1                      /* default x-drop extension */
#define LZB_GFEX_EXACT 2                      /* --exact=N: match_extend_seed_hit seed_search.c:3018; hspThreshold = N (a length) */
#define LZB_GFEX_MISMATCH 3                   /* --mismatch=M,N / --<M>mismatch=N: mismatch_extend_seed_hit :3450; gfMismatches = M */
#define LZB_GFEX_MISMATCH_MAX 50              /* gfexMismatch_max seed_search.h:168 */
typedef struct lzb_seed_params {
    uint32_t start, end;       /* query interval to scan; end==0 => whole query */
    int32_t  gfExtend;         /* LZB_GFEX_* */
    int32_t  xDrop;            /* default 10*sub[A][A] = 910 (lastz.c:9313) */
    int32_t  hspThreshold;     /* K, fixed score threshold (default 3000) */
    int32_t  entropy;          /* nonzero => entropy adjustment (seed_search.c:2851-2874) */
    int32_t  hashBits;         /* log2(diag hash size): 16 stock, 22 for lastz_32 (diag_hash.h:43-59) */
    int32_t  selfCompare;      /* --self: drop hits on/below the diagonal (seed_search.c:2182) */
    int32_t  sameStrand;       /* selfCompare && both sequences on the same strand */
    int32_t  strandId;         /* copied to segment.id */
    int32_t  plainHits;        /* process_for_plain_hit (seed_search.c:995): report every raw hit,
                                  no diag-hash filter; chosen by the reference when neither gap-free
                                  nor gapped extension is requested (lastz.c:2789) */
    int32_t  gfMismatches;     /* LZB_GFEX_MISMATCH: mismatches allowed, 1..50 */
    int32_t  recoverSeeds;     /* --recoverseeds: process_for_recoverable_hit (seed_search.c:1221) instead of the simple
                                  processor -- a hit on another diagonal of the same hash bucket is extended rather than
                                  lost, left extension is not blocked by the bucket, overlapping HSPs may be reported
                                  (the caller merges them: merge_segments segment.c:1527 = lzb_merge_segments of the host
                                  library).  With LZB_GFEX_XDROP or LZB_GFEX_NONE only. */
    int32_t  twinMinSpan;      /* --twins=<minGap>..<maxGap>: process_for_twin_hit (seed_search.c:1814, the seed-hit-queue
                                  version): a hit is extended only when an earlier hit on its diagonal lies minSpan..maxSpan
                                  columns back (span = start of the earlier hit to the end of this one = 2*seedLength + gap,
                                  lastz.c:9835); 0 = off.  With LZB_GFEX_XDROP or LZB_GFEX_NONE only; the caller merges the
                                  table like for recoverSeeds. */
    int32_t  twinMaxSpan;
    int32_t  seedQueueSize;    /* --seedqueue=<entries>, default 256K (diag_hash.h:112); 0 = default */
    uint32_t searchLimit;      /* seed_hit_search's searchLimit (seed_search.h:265; --queryhsplimit): the scan of the query stops
                                  at the end of the query position whose hits brought the number of HSPs reported to
                                  searchLimit + 1 (searchToGo < 0, seed_search.c:551).  The table returned holds exactly the
                                  HSPs the reference has collected at that point; the caller decides what a query that passed
                                  the limit means (lastz.c:3139-3151).  0 = no limit; ignored by the twin processor, which never counts its HSPs.  (The product still enumerates every hit;
                                  its rawSeedHits is then the full count, not the count up to the stop.) */
    int32_t  extendCtasPerSm;  /* tuning, no effect on results: persistent CTAs of the x-drop extension kernel per SM, 1..4
                                  (0 = the default 4).  A caller that runs this stage beside another context's Y-drop sweeps
                                  asks for fewer so that those keep their issue slots (bench.py). */
} lzb_seed_params;

typedef struct lzb_seed_stats {
    uint64_t wordsInQuery;     /* seedSearchStats.wordsInSequence */
    uint64_t rawSeedHits;      /* seedSearchStats.rawSeedHits (seed_search.c:865) */
    uint64_t extensions;       /* hits that passed the diag-hash test and were extended */
    uint64_t bpExtended;       /* seedSearchStats.bpExtended (seed_search.c:2839) */
    uint64_t hsps;             /* HSPs reported */
    double   seconds;          /* device time of the call (CUDA events), 0 for the oracle */
    double   kernelSeconds[12];/* CUDA-event time per kernel: 0 words 1 count 2 slots 3 scan 4 expand 5 sort
                                  6 bounds 7 extend 8 right 9 replay 10 left 11 bucket order */
    uint64_t kernelLaunches[12];/* launches behind each kernelSeconds entry */
} lzb_seed_stats;

/* gapped_extend (gapped_extend.h:153-159) arguments */
typedef struct lzb_gapped_params {
    int32_t  yDrop;            /* default gapOpen + 300*gapExtend = 9400 */
    int32_t  trimToPeak;       /* !--noytrim; default 1 */
    int32_t  scoreThreshold;   /* L, default 3000 */
    int32_t  allBounds;        /* --allgappedbounds */
    int32_t  inhibitTrivial;   /* --notrivial */
    int32_t  identityCheck;    /* nonzero => run identical_sequences (gapped_extend.c:1886) and, for a partitioned
                                  target and a plain query, identical_partition_of_sequence (:2034); the caller
                                  sets it when both sequences carry the same strand flags (:1905, :2058) */
    uint32_t tracebackBytes;   /* --allocate:traceback, default 80 MiB; changes results */
    int32_t  speculation;      /* product only: max anchors extended speculatively in parallel */
    int32_t  overlyPairedKeep; /* with maxPairedBases: keep what was found before the limit was passed (--querydepth=keep:) */
    uint64_t maxPairedBases;   /* gapped_extend's maxPairedBases (gapped_extend.c:1444-1459; --querydepth=<d> = d x query length,
                                  lastz.c:3414-3417): when the aligned (substitution) columns of the alignments kept so far
                                  exceed it, no further anchor is extended and, unless overlyPairedKeep, the call returns no
                                  alignment at all.  0 = no limit.  stats->overlyPaired tells the caller (who warns). */
} lzb_gapped_params;

typedef struct lzb_gapped_stats {
    uint64_t anchors;          /* anchors given */
    uint64_t anchorsExtended;  /* anchors for which ydrop_align ran */
    uint64_t dpCells;          /* gappedExtendStats.dpCellsVisited (gapped_extend.c:3593,3776): DPs whose result was used */
    uint64_t dpRows;
    uint64_t truncated;        /* one-sided DPs stopped by traceback capacity */
    uint64_t speculated;       /* product: DPs launched speculatively */
    uint64_t redone;           /* product: speculative DPs invalidated and recomputed */
    double   seconds;          /* wall time of the call (host clock around the device work) */
    double   kernelSeconds[4]; /* [0] Y-drop DP kernel, [1] traceback kernel (CUDA events), see DESIGN.md */
    uint64_t launches;         /* kernels launched by this call */
    uint64_t dpCellsComputed;  /* product: every cell the device computed, discarded speculation included (>= dpCells) */
    uint64_t overlyPaired;     /* 1 if maxPairedBases was exceeded */
} lzb_gapped_stats;

typedef struct lzb_ctx    lzb_ctx;     /* one device + stream + scratch */
typedef struct lzb_target lzb_target;  /* target bytes + seed position index, resident in HBM */
typedef struct lzb_query  lzb_query;   /* one query strand, resident in HBM */

/* ---- entry points ---- */

const char* lzb_backend(void);          /* "cuda-sm_100a" or "oracle-cpu" */
const char* lzb_last_error(void);

/* device == cuda ordinal.  Fails (NULL) when no sm_100 device is present: there is no CPU path. */
lzb_ctx* lzb_open(int device);
void     lzb_close(lzb_ctx*);

/* scoring->sub and maskedScoring->sub (dna_utilities.h:176-212, 256x256 row-major s32), gap
 * penalties as in scoreset.gapOpen/gapExtend.  Must be called before any search/extend call. */
int lzb_set_scoring(lzb_ctx*, const int32_t* sub, const int32_t* maskedSub,
                    int32_t gapOpen, int32_t gapExtend);

/* build_seed_position_table (pos_table.h:230; pos_table.c:144).  end==0 => len1.  The bytes are
 * copied to the device; seq1[len1] need not be addressable. */
lzb_target* lzb_target_build(lzb_ctx*, const uint8_t* seq1, uint32_t len1,
                             uint32_t start, uint32_t end,
                             const int8_t charToBits[256], const lzb_seed* seed, uint32_t step);
void lzb_target_free(lzb_target*);

/* test hook (the reference exposes the same data through dump_position_table pos_table.h:250):
 * counts[w] for every packed word and the positions of each word in DEcreasing order
 * (the order find_table_matches walks them, seed_search.c:832), concatenated by word.
 * positions may be NULL to get counts only; returns total positions or -1. */
int64_t lzb_target_export_index(lzb_target*, uint32_t* counts, uint32_t* positions);

/* limit_position_table (pos_table.h:248; pos_table.c:1763) with maxChasm == 0: every seed word that occurs more than
 * `limit` times in the target is dropped from the table (--maxwordcount=<limit>, lastz.c:1219).  The percentage form
 * (--maxwordcount=<p>%, find_position_table_limit pos_table.c:2000) is a host computation on the word counts of
 * lzb_target_export_index that ends in this call. */
int lzb_target_limit(lzb_target*, uint32_t limit);

lzb_query* lzb_query_load(lzb_ctx*, const uint8_t* seq2, uint32_t len2);
void       lzb_query_free(lzb_query*);

/* seed_hit_search (seed_search.h:265) with process_for_simple_hit (seed_search.c:1056) as the
 * processor and collect_hsps (lastz.c:3991) as the reporter: returns the HSP table in discovery
 * order.  *segs is malloc'd by the library; release with lzb_free. */
int lzb_seed_hit_search(lzb_ctx*, lzb_target*, lzb_query*, const lzb_seed* seed,
                        const int8_t charToBits[256], const lzb_seed_params*,
                        lzb_segment** segs, uint64_t* nsegs, lzb_seed_stats* stats);

/* reduce_to_points (gapped_extend.h:151; gapped_extend.c:463), in place */
int lzb_reduce_to_points(lzb_ctx*, lzb_target*, lzb_query*, lzb_segment* anchors, uint64_t n);

/* gapped_extend (gapped_extend.h:153; gapped_extend.c:1012).  Reorders anchors in place exactly
 * as the reference does (sort by decreasing score).  Result list is ordered by start in seq1;
 * release with lzb_free_align_list.  seq1/seq2 in each alignel point at the host bytes given
 * here (may be NULL).
 * Partitioned ([multi]) sequences: bytes of value 0 INSIDE either sequence separate partitions (sequences.h:188-191);
 * seeds and x-drop scans never cross them and every DP sweep ends at the separators around its anchor (:1357-1372).
 */
int lzb_gapped_extend(lzb_ctx*, lzb_target*, lzb_query*,
                      const uint8_t* hostSeq1, const uint8_t* hostSeq2,
                      lzb_segment* anchors, uint64_t n, const lzb_gapped_params*,
                      lzb_alignel** list, lzb_gapped_stats* stats);

void lzb_free_align_list(lzb_alignel*);   /* free_align_list edit_script.c:53 */

/* kernels launched so far through this context (0 for the oracle): bench.py's gpu_launches */
uint64_t lzb_launch_count(lzb_ctx*);
void lzb_free(void*);

#ifdef __cplusplus
}
#endif
#endif /* LASTZ_B200_H */
