/*
 * lzb_host.h -- host-side (plain C) front end of lastz_b200: the data contracts on either side
 * of the hot path (sequence files, scoring sets, seeds, output writers) and the `lastz` command
 * line subset.  Everything here is reference-compatible plumbing; all seed-and-extend work goes
 * through the C-ABI in include/lastz_b200.h.
 */
#ifndef LZB_HOST_H
#define LZB_HOST_H

#include <stdint.h>
#include <stdio.h>
#include "../../../include/lastz_b200.h"

/* strand flags, sequences.h:345-350 */
#define LZB_RCF_FORWARD 0
#define LZB_RCF_COMP    1
#define LZB_RCF_REV     2
#define LZB_RCF_REVCOMP 3

/* the subset of `seq` (sequences.h:381-588) the path and the writers need */
typedef struct lzb_seq {
    uint8_t* v;          /* bases, NUL terminated (v[len] == 0) */
    uint8_t* vq;         /* base qualities parallel to v (FASTQ input), or NULL; only the SAM writer reads them */
    uint32_t len;        /* bases held in v */
    uint32_t startLoc;   /* 1-based position of v[0] in the full sequence */
    uint32_t trueLen;    /* length of the full sequence */
    int      revCompFlags;
    uint32_t contig;     /* 1-based ordinal within the file */
    char*    filename;   /* as given on the command line, actions stripped */
    char*    header;     /* full header line (FASTA: including '>') */
    char*    shortHeader;/* first word of the header */
    /* [multi]: several sequences in one vector, v = NUL seq0 NUL seq1 NUL ... (sequences.h:188-191 note 2): v[0] and
     * every separator are NUL, len counts them, part[i] describes sequence i; npart == 0 for an ordinary sequence */
    uint32_t npart;
    struct lzb_partition* part;
} lzb_seq;
typedef struct lzb_partition {
    uint32_t sepBefore, sepAfter;   /* indices of the NULs around the sequence */
    uint32_t contig, startLoc, trueLen;
    char* header; char* shortHeader;
} lzb_partition;
/* name, offset in v, startLoc, length and full length of the (part of the) sequence that holds position pos0 */
typedef struct lzb_seqview { const char* name; uint32_t offset, startLoc, len, trueLen, contig; } lzb_seqview;
void lzb_seq_view(const lzb_seq* s, uint32_t pos0, lzb_seqview* out);

/* a sequence file + bracketed actions; load successive sequences with lzb_seqfile_next */
typedef struct lzb_seqfile lzb_seqfile;
lzb_seqfile* lzb_seqfile_open(const char* spec);       /* "file[actions]" or "file.2bit/name[...]" */
int          lzb_seqfile_next(lzb_seqfile*, lzb_seq* out);   /* 1 = loaded, 0 = no more */
void         lzb_seqfile_close(lzb_seqfile*);
void         lzb_seq_free(lzb_seq*);
void         lzb_seq_revcomp(lzb_seq*);                /* rev_comp_sequence sequences.c:7511 */
void         lzb_die(const char* fmt, ...);            /* suicidef utilities.c:1866 */

/* scoreset (dna_utilities.h:176-212): the two matrices + gap penalties + embedded parameters */
typedef struct lzb_scoreset {
    int32_t sub[256 * 256];
    int32_t masked[256 * 256];
    int32_t gapOpen, gapExtend;
    int gapOpenSet, gapExtendSet;
    int hspThresholdSet, gappedThresholdSet, xDropSet, yDropSet, stepSet;
    int32_t hspThreshold, gappedThreshold, xDrop, yDrop; uint32_t step;
} lzb_scoreset;
void lzb_scores_default(lzb_scoreset*);                        /* HOXD70, dna_utilities.c:137-148 */
void lzb_scores_from_template(lzb_scoreset*, int32_t tmpl[4][4], int32_t bad, int32_t fill,
                              int32_t gapOpen, int32_t gapExtend);   /* new_dna_score_set :215 */
void lzb_scores_mask(lzb_scoreset*);                           /* masked_score_set :497 */
void lzb_scores_read_file(lzb_scoreset*, const char* path);    /* read_score_set_by_name :657 */

/* seeds (seeds.c:321-632) */
#define LZB_SEED_12OF19 "1110100110010101111"
#define LZB_SEED_14OF22 "1110101100110010101111"
void lzb_seed_parse(lzb_seed* out, const char* pattern, int withTrans);
extern const int8_t lzb_upper_nuc_to_bits[256];                /* dna_utilities.c:76-94 */
extern const int8_t lzb_nuc_to_bits[256];                      /* dna_utilities.c:56-74 */

/* best-chain reduction (chain.c:497, penalties lastz.c:3687); rewrites segs[0..*n) to the chain, sorted by pos1 */
int32_t lzb_reduce_to_chain(lzb_segment* segs, uint64_t* n, int32_t diagPen, int32_t antiPen, int32_t scale, int32_t subAA);
void lzb_merge_segments(lzb_segment* segs, uint64_t* n);                                     /* merge_segments segment.c:1527 (merge.c) */
/* ... per pair of partitions when either sequence is [multi] (try_reduce_to_chain chain.c:224), result sorted by pos1 */
int32_t lzb_reduce_to_chains(const lzb_seq* s1, const lzb_seq* s2, lzb_segment* segs, uint64_t* n,
                             int32_t diagPen, int32_t antiPen, int32_t scale, int32_t subAA);

/* adaptive.c -- HSP table of an adaptive threshold (add_segment segment.c:981): a list until the covered bases
 * reach the limit, then a min-heap on score that drops whole groups of tied lowest scores */
typedef struct lzb_hsptable { lzb_segment* seg; uint32_t len, cap; uint64_t coverage, limit; int32_t lowScore; } lzb_hsptable;
void   lzb_hsptable_init(lzb_hsptable*, uint64_t coverageLimit);     /* limit 0 = keep everything, in arrival order */
void   lzb_hsptable_add(lzb_hsptable*, const lzb_segment*);
void   lzb_hsptable_split(lzb_hsptable*, int id, lzb_hsptable* rest); /* split_segment_table segment.c:1352 */
void   lzb_hsptable_free(lzb_hsptable*);
double lzb_hsp_entropy(const uint8_t* s, const uint8_t* t, uint32_t len);   /* entropy dna_utilities.c:2883 */

/* output writers */
void lzb_lav_job_header(FILE*, const char* prog, const char* name1, const char* name2,
                        const char* args, const lzb_scoreset*, const char* K, const char* L);  /* lav.c:40; thresholds as text (score_thresh_to_string) */
void lzb_lav_strand_header(FILE*, const lzb_seq* s1, const lzb_seq* s2);               /* lav.c:101 */
void lzb_lav_align(FILE*, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel*);   /* lav.c:187 */
void lzb_lav_match(FILE*, const lzb_seq* s1, const lzb_seq* s2, const lzb_segment*);   /* lav.c:327 */
void lzb_lav_footer(FILE*);                                                            /* m stanza + #:eof */
/* general.c -- --format=general[-][:fields] / mapping[-] / cigar (genpaf.c, cigar.c): rows with named columns */
typedef struct lzb_fieldlist lzb_fieldlist;
lzb_fieldlist* lzb_fieldlist_parse(const char* commaSeparatedNames);      /* parse_genpaf_keys genpaf.c:1945 */
lzb_fieldlist* lzb_fieldlist_standard(void);                              /* --format=general */
lzb_fieldlist* lzb_fieldlist_mapping(void);                               /* --format=mapping */
lzb_fieldlist* lzb_fieldlist_paf(int wfmash);                             /* --format=paf[:minimap2|:wfmash] */
lzb_fieldlist* lzb_fieldlist_blastn(void);                                /* --format=blastn[-] */
void lzb_blastn_header(FILE*, const char* prog, const char* args, const char* databaseFile, const lzb_seq* query);
void lzb_fieldlist_header(FILE*, const lzb_fieldlist*);
void lzb_fieldlist_align(FILE*, const lzb_fieldlist*, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel*, uint64_t* number);
void lzb_fieldlist_match(FILE*, const lzb_fieldlist*, const lzb_seq* s1, const lzb_seq* s2, const lzb_segment*, uint64_t* number);
typedef struct lzb_alignstats { uint64_t idNumer, idDenom, covNumer, covDenom, conNumer, conDenom, gapNumer, gapDenom, ngap; } lzb_alignstats;
void lzb_align_stats(const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel*, lzb_alignstats*);
/* --filter=identity|coverage|continuity|nmatch|nmismatch|ngap|cgap (lastz.c:6672-6950); fractions, not percentages */
typedef struct lzb_filters { float minIdentity, maxIdentity, minCoverage, maxCoverage, minContinuity, maxContinuity;
                             uint32_t minMatchCount; int32_t maxMismatchCount, maxSeparateGaps, maxGapColumns; } lzb_filters;
void lzb_filters_init(lzb_filters*);
int  lzb_filters_active(const lzb_filters*);
int  lzb_filters_reject(const lzb_filters*, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel*, int isSegment);
void lzb_cigar_align(FILE*, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel*);
void lzb_cigar_match(FILE*, const lzb_seq* s1, const lzb_seq* s2, const lzb_segment*);
typedef struct lzb_rdotplot { char prev1[256], prev2[256];                      /* the name pair last announced */
                              int limited; uint32_t blocksLeft; } lzb_rdotplot;  /* --queryhsplimit: blocks this query may still print (output.c:744-747) */
void lzb_rdotplot_align(FILE*, lzb_rdotplot*, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel*, const lzb_scoreset*, int withScore);
void lzb_rdotplot_match(FILE*, lzb_rdotplot*, const lzb_seq* s1, const lzb_seq* s2, const lzb_segment*, const lzb_scoreset*, int withScore);
void lzb_sam_header(FILE*, const lzb_seq* s1);                             /* sam.c:196-232 */
void lzb_sam_align(FILE*, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel*, int markMismatches, int softClip);
void lzb_sam_match(FILE*, const lzb_seq* s1, const lzb_seq* s2, const lzb_segment*, int markMismatches, int softClip);
void lzb_maf_align(FILE*, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel*);    /* maf.c:271, --format=maf- */
/* gfa.c:95-330, --format=gfa */
void lzb_gfa_job_header(FILE*, const char* prog, const char* name1, const char* name2, const char* seedPattern, int withTrans, uint32_t step);
void lzb_gfa_strand_header(FILE*, const lzb_seq* s1, const lzb_seq* s2);
void lzb_gfa_match(FILE*, const lzb_seq* s1, const lzb_seq* s2, const lzb_segment*);
void lzb_gfa_align(FILE*, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel*, const lzb_scoreset*);
void lzb_axt_header(FILE*, const char* prog, const char* args, const lzb_scoreset*, const char* K, const char* L, int32_t X, int32_t Y);
void lzb_axt_align(FILE*, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel*, uint64_t* number);   /* axt.c:96, --format=axt */

/* mirror.c -- mirror_alignments lastz.c:4229 (--self with gapped extension) */
lzb_alignel* lzb_mirror_alignments(lzb_alignel* list, const lzb_seq* seq1, const lzb_seq* seq2, const lzb_scoreset* ss);

#endif
