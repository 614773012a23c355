/*
 * general.c -- --format=general[-][:<fields>], --format=mapping[-] and --format=cigar: one row per alignment (or HSP),
 * the columns chosen by name.  Reference: genpaf.c (field names genpaf.h:150-250, header :87-190, rows :545-1490),
 * cigar.c:135-360 and :528-605, the statistics behind identity / coverage / continuity / gap rate
 * (identity_dist.c, coverage_dist.c:132, continuity_dist.c:282-349).
 *
 * Host-side plumbing on the output side of the hot path.  An HSP is printed as an alignment whose edit script is one
 * run of substitutions, so both kinds of record go through the same code.  Fields that need data this front end does
 * not carry (base qualities, chore ids, hashes, BLAST statistics) are refused by name.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "lzb_host.h"

enum {
    F_NA, F_NAME1, F_NUMBER1, F_STRAND1, F_SIZE1, F_START1, F_ZSTART1, F_END1, F_LENGTH1, F_TEXT1,
    F_NAME2, F_NUMBER2, F_STRAND2, F_SIZE2, F_START2, F_ZSTART2, F_START2P, F_ZSTART2P, F_END2, F_END2P, F_LENGTH2, F_TEXT2,
    F_NMATCH, F_NMISMATCH, F_NPAIR, F_NCOLUMN, F_NGAP, F_CGAP, F_DIFF, F_CIGAR, F_CIGARL, F_CIGARX, F_CIGARXL, F_CIGARX1, F_CIGARX1L,
    F_DIAGONAL, F_SHINGLE, F_SCORE, F_IDENTITY, F_IDFRAC, F_IDPCT, F_BLASTIDPCT, F_COVERAGE, F_COVFRAC, F_COVPCT,
    F_CONTINUITY, F_CONFRAC, F_CONPCT, F_GAPRATE, F_NUMBER, F_ZNUMBER,
    /* only inside the canned lists of --format=paf / blastn (genpaf.h:119-125) */
    F_MAPQUAL, F_ASTAG, F_CGTAG_X, F_CGTAG_M, F_BSTART1, F_BEND1, F_EVALUE, F_BITSCORE
};
static const struct { const char* name; int code; const char* heading; } FIELDS[] = {
    { "name1", F_NAME1, "name1" }, { "number1", F_NUMBER1, "number1" }, { "strand1", F_STRAND1, "strand1" }, { "size1", F_SIZE1, "size1" },
    { "start1", F_START1, "start1" }, { "zstart1", F_ZSTART1, "zstart1" }, { "end1", F_END1, "end1" }, { "length1", F_LENGTH1, "length1" },
    { "align1", F_TEXT1, "align1" }, { "text1", F_TEXT1, "text1" },
    { "name2", F_NAME2, "name2" }, { "number2", F_NUMBER2, "number2" }, { "strand2", F_STRAND2, "strand2" }, { "size2", F_SIZE2, "size2" },
    { "start2", F_START2, "start2" }, { "zstart2", F_ZSTART2, "zstart2" }, { "start2+", F_START2P, "start2+" }, { "zstart2+", F_ZSTART2P, "zstart2+" },
    { "end2", F_END2, "end2" }, { "end2+", F_END2P, "end2+" }, { "length2", F_LENGTH2, "length2" }, { "align2", F_TEXT2, "align2" }, { "text2", F_TEXT2, "text2" },
    { "nmatch", F_NMATCH, "nmatch" }, { "nmismatch", F_NMISMATCH, "nmismatch" }, { "npair", F_NPAIR, "npair" }, { "ncolumn", F_NCOLUMN, "ncolumn" },
    { "ngap", F_NGAP, "ngap" }, { "cgap", F_CGAP, "cgap" }, { "diff", F_DIFF, "diff" },
    { "cigar", F_CIGAR, "cigar" }, { "cigar-", F_CIGARL, "cigar-" }, { "cigarx", F_CIGARX, "cigarx" }, { "cigarx-", F_CIGARXL, "cigarx-" },
    { "cigarx1", F_CIGARX1, "cigarx1" }, { "cigarx1-", F_CIGARX1L, "cigarx1-" },
    { "diagonal", F_DIAGONAL, "diagonal" }, { "shingle", F_SHINGLE, "shingle" }, { "score", F_SCORE, "score" },
    { "identity", F_IDENTITY, "identity\tidPct" }, { "idfrac", F_IDFRAC, "idfrac" }, { "id%", F_IDPCT, "id%" }, { "blastid%", F_BLASTIDPCT, "blastid%" },
    { "coverage", F_COVERAGE, "coverage\tcovPct" }, { "covfrac", F_COVFRAC, "covfrac" }, { "cov%", F_COVPCT, "cov%" },
    { "continuity", F_CONTINUITY, "continuity\tconPct" }, { "confrac", F_CONFRAC, "confrac" }, { "con%", F_CONPCT, "con%" },
    { "gaprate", F_GAPRATE, "gaprate\tgapPct" }, { "number", F_NUMBER, "number" }, { "znumber", F_ZNUMBER, "znumber" }, { "NA", F_NA, "" },
    /* the short aliases, genpaf.h:221-250 */
    { "n1", F_NAME1, "name1" }, { "s1", F_START1, "start1" }, { "z1", F_ZSTART1, "zstart1" }, { "e1", F_END1, "end1" }, { "l1", F_LENGTH1, "length1" },
    { "a1", F_TEXT1, "align1" }, { "t1", F_TEXT1, "text1" }, { "n2", F_NAME2, "name2" }, { "s2", F_START2, "start2" }, { "z2", F_ZSTART2, "zstart2" },
    { "s2+", F_START2P, "start2+" }, { "z2+", F_ZSTART2P, "zstart2+" }, { "e2", F_END2, "end2" }, { "e2+", F_END2P, "end2+" }, { "l2", F_LENGTH2, "length2" },
    { "a2", F_TEXT2, "align2" }, { "t2", F_TEXT2, "text2" }, { "d", F_DIAGONAL, "diagonal" }, { "diag", F_DIAGONAL, "diagonal" }, { "s", F_SCORE, "score" },
    { "id", F_IDENTITY, "identity\tidPct" }, { "ident", F_IDENTITY, "identity\tidPct" }, { "cov", F_COVERAGE, "coverage\tcovPct" },
    { "con", F_CONTINUITY, "continuity\tconPct" }, { "gap", F_GAPRATE, "gaprate\tgapPct" },
    { NULL, 0, NULL }
};
static const char* const NOT_CARRIED[] = { "qalign1", "qalign2", "nucs1", "quals1", "nucs2", "quals2", "chore", "entropy1", "entropy2",
                                           "hspid", "phash", "ahash", "~", NULL };

struct lzb_fieldlist { int n; int code[64]; const char* heading[64]; };

/* parse_genpaf_keys genpaf.c:1945: comma-separated field names */
lzb_fieldlist* lzb_fieldlist_parse(const char* spec) {
    lzb_fieldlist* fl = calloc(1, sizeof *fl);
    char* copy = strdup(spec);
    for (char* tok = strtok(copy, ","); tok; tok = strtok(NULL, ",")) {
        int k = 0;
        for (; FIELDS[k].name && strcmp(FIELDS[k].name, tok); k++) ;
        if (!FIELDS[k].name) {
            for (int u = 0; NOT_CARRIED[u]; u++)
                if (!strcmp(NOT_CARRIED[u], tok)) lzb_die("lastz_b200 does not carry what the field \"%s\" needs (--format=general)", tok);
            lzb_die("unrecognized field name (for --format=general): \"%s\"", tok);
        }
        if (fl->n >= 64) lzb_die("too many fields for --format=general");
        fl->code[fl->n] = FIELDS[k].code; fl->heading[fl->n] = FIELDS[k].heading; fl->n++;
    }
    free(copy);
    if (fl->n == 0) lzb_die("empty keys string for --format=general:");
    return fl;
}
lzb_fieldlist* lzb_fieldlist_standard(void) {      /* genpafStandardKeys genpaf.h:117 */
    return lzb_fieldlist_parse("score,name1,strand1,size1,zstart1,end1,name2,strand2,size2,zstart2,end2,identity,coverage");
}
lzb_fieldlist* lzb_fieldlist_mapping(void) {       /* genpafMappingKeys genpaf.h:118 */
    return lzb_fieldlist_parse("name1,zstart1,end1,name2,strand2,zstart2+,end2+,identity,coverage,cigarx-");
}
static lzb_fieldlist* canned(const int* codes, int n) {
    lzb_fieldlist* fl = calloc(1, sizeof *fl);
    for (int k = 0; k < n; k++) { fl->code[k] = codes[k]; fl->heading[k] = ""; }
    fl->n = n;
    return fl;
}
lzb_fieldlist* lzb_fieldlist_paf(int wfmash) {     /* genpafPafMinimap2Keys "ns>,dNSZEuW{|." / genpafPafWfMashKeys "...}" */
    const int codes[14] = { F_NAME2, F_SIZE2, F_ZSTART2P, F_END2P, F_STRAND2, F_NAME1, F_SIZE1, F_ZSTART1, F_END1, F_NMATCH, F_NCOLUMN, F_MAPQUAL, F_ASTAG,
                            wfmash ? F_CGTAG_X : F_CGTAG_M };
    return canned(codes, 14);
}
lzb_fieldlist* lzb_fieldlist_blastn(void) {        /* genpafBlastKeys "nNmWvy<,QR%$" */
    const int codes[12] = { F_NAME2, F_NAME1, F_BLASTIDPCT, F_NCOLUMN, F_NMISMATCH, F_NGAP, F_START2P, F_END2P, F_BSTART1, F_BEND1, F_EVALUE, F_BITSCORE };
    return canned(codes, 12);
}
/* print_blast_header genpaf.c:228: one comment block per query */
void lzb_blastn_header(FILE* f, const char* prog, const char* args, const char* databaseFile, const lzb_seq* query) {
    const char* name = query->shortHeader && query->shortHeader[0] ? query->shortHeader : "query";
    fprintf(f, "# %s %s\n# Query: %s\n# Database: %s\n", prog, args, name, databaseFile);
    fprintf(f, "# Fields: query id, subject id, %% identity, alignment length, mismatches, gap opens, q. start, q. end, s. start, s. end, evalue, bit score\n");
}
void lzb_fieldlist_header(FILE* f, const lzb_fieldlist* fl) {
    for (int k = 0; k < fl->n; k++) fprintf(f, "%s%s", k ? "\t" : "#", fl->heading[k]);
    fputc('\n', f);
}

/* ---- walking an edit script: runs of substitutions separated by gaps ---- */
typedef struct { const lzb_editscript* sc; uint32_t k, i, j, height, width; } walker;
static void walk_start(walker* w, const lzb_alignel* a) { w->sc = a->script; w->k = 0; w->i = w->j = 0; w->height = a->end1 - a->beg1 + 1; w->width = a->end2 - a->beg2 + 1; }
static int walk_more(const walker* w) { return w->i < w->height || w->j < w->width; }
static uint32_t walk_subs(walker* w) {             /* edit_script_run_of_subs edit_script.c */
    uint32_t run = 0;
    while (w->k < w->sc->len && (w->sc->op[w->k] & 3) == LZB_OP_SUB) { run += w->sc->op[w->k] >> 2; w->k++; }
    return run;
}
static void walk_gap(walker* w, uint32_t* di, uint32_t* dj) {   /* edit_script_indel_len edit_script.c:778: ONE op */
    *di = *dj = 0;
    if (w->k < w->sc->len) {
        uint32_t op = w->sc->op[w->k] & 3, rpt = w->sc->op[w->k] >> 2; w->k++;
        if (op == LZB_OP_INS) *dj = rpt; else if (op == LZB_OP_DEL) *di = rpt;
    }
    w->i += *di; w->j += *dj;
}

static char printable(uint8_t c) { return (c >= 0x20 && c < 0x7F) ? (char)c : '*'; }       /* dna_toprint dna_utilities.h:305 */

/* a run of aligned bases as =/X runs (print_cigar_mismatchy_run cigar.c:528) */
static void mismatchy_run(FILE* f, const uint8_t* a, const uint8_t* b, uint32_t len, int letterAfter, int withSpaces, int hideSingles, int lower) {
    const char chX = lower ? 'x' : 'X';
    int runIsMm = 0; uint32_t runLen = 0;
    for (uint32_t x = 0; x <= len; x++) {
        int flush = x == len, mm = 0;
        if (!flush) { int p = lzb_nuc_to_bits[a[x]], q = lzb_nuc_to_bits[b[x]]; mm = !(p == q && p >= 0); }
        if (!flush && mm == runIsMm) { runLen++; continue; }
        if (runLen > 0) {
            char ch = runIsMm ? chX : '=';
            if (!letterAfter && !withSpaces) fprintf(f, "%c%u", ch, runLen);
            else if (!letterAfter) fprintf(f, " %c %u", ch, runLen);
            else if (hideSingles && runLen == 1) fprintf(f, "%c", ch);
            else fprintf(f, "%u%c", runLen, ch);
        }
        runIsMm = mm; runLen = 1;
    }
}

/* the operations of an alignment in CIGAR-like text (the loop of print_cigar_align cigar.c:300-350) */
static void cigar_ops(FILE* f, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel* a, int markMismatches, int letterAfter, int withSpaces, int hideSingles, int lower) {
    const char chM = lower ? 'm' : 'M', chD = lower ? 'd' : 'D', chI = lower ? 'i' : 'I';
    walker w; walk_start(&w, a);
    while (walk_more(&w)) {
        uint32_t run = walk_subs(&w);
        if (run > 0) {
            if (markMismatches) mismatchy_run(f, s1->v + a->beg1 - 1 + w.i, s2->v + a->beg2 - 1 + w.j, run, letterAfter, withSpaces, hideSingles, lower);
            else if (letterAfter) fprintf(f, "%u%c", run, chM);
            else fprintf(f, " %c %u", chM, run);
            w.i += run; w.j += run;
        }
        if (!walk_more(&w)) break;
        if (w.k >= w.sc->len) break;
        uint32_t di, dj; walk_gap(&w, &di, &dj);
        if (di) { if (!letterAfter) fprintf(f, " %c %u", chD, di); else if (hideSingles && di == 1) fprintf(f, "%c", chD); else fprintf(f, "%u%c", di, chD); }
        if (dj) { if (!letterAfter) fprintf(f, " %c %u", chI, dj); else if (hideSingles && dj == 1) fprintf(f, "%c", chI); else fprintf(f, "%u%c", dj, chI); }
    }
}

/* one character of the diff field (diff_char genpaf.c:1857, info string ".:x--X") */
static char diff_char(uint8_t p, uint8_t q) {
    int a = lzb_nuc_to_bits[p], b = lzb_nuc_to_bits[q];
    if (a < 0 || b < 0) return 'X';
    if (a == b) return '.';
    return ((a ^ b) == 2) ? ':' : 'x';                              /* A<->G, C<->T differ in bit 1 only */
}

/* matches / aligned pairs (alignment_identity identity_dist.c:184), shorter sequence covered (alignment_coverage
 * coverage_dist.c:132), gap-free columns / columns and gapped bases / gap-free columns (continuity_dist.c:282-349) */
void lzb_align_stats(const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel* a, lzb_alignstats* st) {
    lzb_seqview w1, w2; lzb_seq_view(s1, a->beg1 - 1, &w1); lzb_seq_view(s2, a->beg2 - 1, &w2);
    const uint32_t beg1 = a->beg1, beg2 = a->beg2, height = a->end1 - beg1 + 1, width = a->end2 - beg2 + 1;
    uint64_t idNumer = 0, idDenom = 0, subs = 0, ngap = 0;
    walker w; walk_start(&w, a);
    while (walk_more(&w)) {
        uint32_t run = walk_subs(&w);
        const uint8_t* p = s1->v + beg1 - 1 + w.i; const uint8_t* q = s2->v + beg2 - 1 + w.j;
        for (uint32_t x = 0; x < run; x++) { int u = lzb_nuc_to_bits[p[x]], v = lzb_nuc_to_bits[q[x]]; if (u >= 0 && v >= 0) { idDenom++; if (u == v) idNumer++; } }
        subs += run; w.i += run; w.j += run;
        if (!walk_more(&w) || w.k >= w.sc->len) break;
        uint32_t di, dj; walk_gap(&w, &di, &dj); ngap++;
    }
    if (idDenom == 0) idNumer = 0;
    memset(st, 0, sizeof *st);
    st->idNumer = idNumer; st->idDenom = idDenom; st->ngap = ngap;
    if (subs) { st->gapNumer = (height - subs) + (width - subs); st->gapDenom = subs; }
    st->conNumer = st->gapDenom; st->conDenom = st->gapDenom + st->gapNumer;
    if (w1.trueLen < w2.trueLen) { st->covNumer = height; st->covDenom = w1.trueLen; } else { st->covNumer = width; st->covDenom = w2.trueLen; }
}

/* the --filter= family (lastz.c:3430-3462 for alignments, :3312-3332 for HSPs); single-precision products as in the
 * reference (identity_dist.c:102-104, coverage_dist.c:86-89, continuity_dist.c:92-95) */
void lzb_filters_init(lzb_filters* f) {
    memset(f, 0, sizeof *f);
    f->maxIdentity = f->maxCoverage = f->maxContinuity = 1; f->maxMismatchCount = f->maxSeparateGaps = f->maxGapColumns = -1;
}
int lzb_filters_active(const lzb_filters* f) {
    return f->minIdentity > 0 || f->maxIdentity < 1 || f->minCoverage > 0 || f->maxCoverage < 1 || f->minContinuity > 0 || f->maxContinuity < 1 ||
           f->minMatchCount > 0 || f->maxMismatchCount >= 0 || f->maxSeparateGaps >= 0 || f->maxGapColumns >= 0;
}
int lzb_filters_reject(const lzb_filters* f, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel* a, int isSegment) {
    lzb_alignstats st; lzb_align_stats(s1, s2, a, &st);
    const uint32_t idN = (uint32_t)st.idNumer, idD = (uint32_t)st.idDenom, covN = (uint32_t)st.covNumer, covD = (uint32_t)st.covDenom,
                   conN = (uint32_t)st.conNumer, conD = (uint32_t)st.conDenom;
    if (f->minIdentity > 0 || f->maxIdentity < 1)
        if (idD == 0 || idN < idD * f->minIdentity || idN > idD * f->maxIdentity) return 1;
    if (f->minCoverage > 0 || f->maxCoverage < 1) {
        float lo = covD * f->minCoverage, hi = covD * f->maxCoverage;
        if (covN < lo || covN > hi) return 1;
    }
    if (!isSegment && (f->minContinuity > 0 || f->maxContinuity < 1)) {
        float lo = conD * f->minContinuity, hi = conD * f->maxContinuity;
        if (conN < lo || conN > hi) return 1;
    }
    if (f->minMatchCount > 0 && (idD == 0 || idN < f->minMatchCount)) return 1;
    if (f->maxMismatchCount >= 0 && (idD == 0 || idD - idN > (uint32_t)f->maxMismatchCount)) return 1;
    if (!isSegment && f->maxSeparateGaps >= 0 && (int64_t)st.ngap > f->maxSeparateGaps) return 1;
    if (!isSegment && f->maxGapColumns >= 0 && (conD == 0 || conD - conN > (uint32_t)f->maxGapColumns)) return 1;
    return 0;
}

static const char* const RCF_SUFFIX[4] = { "", "~", "~", "" };      /* genpaf.c / cigar.c:191 */

void lzb_fieldlist_align(FILE* f, const lzb_fieldlist* fl, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel* a, uint64_t* number) {
    lzb_seqview w1, w2; lzb_seq_view(s1, a->beg1 - 1, &w1); lzb_seq_view(s2, a->beg2 - 1, &w2);
    const char* name1 = w1.name && w1.name[0] ? w1.name : "seq1"; const char* name2 = w2.name && w2.name[0] ? w2.name : "seq2";
    const uint32_t beg1 = a->beg1, beg2 = a->beg2, height = a->end1 - beg1 + 1, width = a->end2 - beg2 + 1;
    uint32_t start1, start2; char strand1, strand2;
    if (!(s1->revCompFlags & LZB_RCF_REV)) { start1 = beg1 - 1 - w1.offset + w1.startLoc; strand1 = '+'; }
    else { start1 = beg1 - 1 - w1.offset + w1.trueLen + 2 - (w1.startLoc + w1.len); strand1 = '-'; }
    if (!(s2->revCompFlags & LZB_RCF_REV)) { start2 = beg2 - 1 - w2.offset + w2.startLoc; strand2 = '+'; }
    else { start2 = beg2 - 1 - w2.offset + w2.trueLen + 2 - (w2.startLoc + w2.len); strand2 = '-'; }
    lzb_alignstats st; lzb_align_stats(s1, s2, a, &st);
    const uint64_t idNumer = st.idNumer, idDenom = st.idDenom, covNumer = st.covNumer, covDenom = st.covDenom, conNumer = st.conNumer,
                   conDenom = st.conDenom, gapNumer = st.gapNumer, gapDenom = st.gapDenom, ngap = st.ngap;
    walker w;
    const uint64_t ordinal = (*number)++;

    for (int k = 0; k < fl->n; k++) {
        if (k) fputc('\t', f);
        switch (fl->code[k]) {
            case F_NA: fprintf(f, "NA"); break;
            case F_NAME1: fprintf(f, "%s%s", name1, RCF_SUFFIX[s1->revCompFlags & 3]); break;
            case F_NUMBER1: fprintf(f, "%u", w1.contig - 1); break;
            case F_STRAND1: fputc(strand1, f); break;
            case F_SIZE1: fprintf(f, "%u", w1.trueLen); break;
            case F_START1: fprintf(f, "%u", start1); break;
            case F_ZSTART1: fprintf(f, "%u", start1 - 1); break;
            case F_END1: fprintf(f, "%u", start1 + height - 1); break;
            case F_LENGTH1: fprintf(f, "%u", height); break;
            case F_NAME2: fprintf(f, "%s%s", name2, RCF_SUFFIX[s2->revCompFlags & 3]); break;
            case F_NUMBER2: fprintf(f, "%u", w2.contig - 1); break;
            case F_STRAND2: fputc(strand2, f); break;
            case F_SIZE2: fprintf(f, "%u", w2.trueLen); break;
            case F_START2: fprintf(f, "%u", start2); break;
            case F_ZSTART2: fprintf(f, "%u", start2 - 1); break;
            case F_START2P: fprintf(f, "%u", strand2 == '-' ? w2.trueLen + 2 - start2 - width : start2); break;
            case F_ZSTART2P: fprintf(f, "%u", strand2 == '-' ? w2.trueLen + 1 - start2 - width : start2 - 1); break;
            case F_END2: fprintf(f, "%u", start2 + width - 1); break;
            case F_END2P: fprintf(f, "%u", strand2 == '-' ? w2.trueLen + 1 - start2 : start2 + width - 1); break;
            case F_LENGTH2: fprintf(f, "%u", width); break;
            case F_TEXT1: case F_TEXT2: case F_DIFF: {
                const int which = fl->code[k];
                walk_start(&w, a);
                while (walk_more(&w)) {
                    uint32_t run = walk_subs(&w);
                    const uint8_t* p = s1->v + beg1 - 1 + w.i; const uint8_t* q = s2->v + beg2 - 1 + w.j;
                    for (uint32_t x = 0; x < run; x++) fputc(which == F_TEXT1 ? printable(p[x]) : which == F_TEXT2 ? printable(q[x]) : diff_char(p[x], q[x]), f);
                    w.i += run; w.j += run;
                    if (!walk_more(&w) || w.k >= w.sc->len) break;
                    p += run; q += run;
                    uint32_t di, dj; walk_gap(&w, &di, &dj);
                    for (uint32_t x = 0; x < di; x++) fputc(which == F_TEXT1 ? printable(p[x]) : '-', f);
                    for (uint32_t x = 0; x < dj; x++) fputc(which == F_TEXT2 ? printable(q[x]) : '-', f);
                }
                break;
            }
            case F_NMATCH: fprintf(f, "%llu", (unsigned long long)idNumer); break;
            case F_NMISMATCH: fprintf(f, "%llu", (unsigned long long)(idDenom - idNumer)); break;
            case F_NPAIR: fprintf(f, "%llu", (unsigned long long)idDenom); break;
            case F_NCOLUMN: fprintf(f, "%llu", (unsigned long long)conDenom); break;
            case F_NGAP: fprintf(f, "%llu", (unsigned long long)ngap); break;
            case F_CGAP: fprintf(f, "%llu", (unsigned long long)(conDenom - conNumer)); break;
            case F_CIGAR: cigar_ops(f, s1, s2, a, 0, 1, 0, 0, 0); break;
            case F_CIGARL: cigar_ops(f, s1, s2, a, 0, 1, 0, 0, 1); break;
            case F_CIGARX: cigar_ops(f, s1, s2, a, 1, 1, 1, 1, 0); break;
            case F_CIGARXL: cigar_ops(f, s1, s2, a, 1, 1, 1, 1, 1); break;
            case F_CIGARX1: cigar_ops(f, s1, s2, a, 1, 1, 1, 0, 0); break;
            case F_CIGARX1L: cigar_ops(f, s1, s2, a, 1, 1, 1, 0, 0); break;     /* the reference tests the wrong key here, so cigarx1- prints upper case (genpaf.c:1142) */
            case F_DIAGONAL: fprintf(f, "%d", (int32_t)(start1 - start2)); break;
            case F_SHINGLE: {
                int64_t diag = (int64_t)start1 - (int64_t)start2, diagSE = (int64_t)w1.len - diag, diagNW = (int64_t)w2.len + diag;
                if (diag < 0) diag = (diagNW < 0 || (uint32_t)diagNW < w1.len) ? -diagNW : 0;
                else if (diag > 0) diag = (diagSE < 0 || (uint32_t)diagSE < w2.len) ? diagSE : 0;
                if (diag == 0) fprintf(f, "NA"); else fprintf(f, "%lld", (long long)diag);
                break;
            }
            case F_SCORE: fprintf(f, "%d", a->s); break;
            case F_ZNUMBER: fprintf(f, "%llu", (unsigned long long)ordinal); break;
            case F_MAPQUAL: fprintf(f, "255"); break;
            case F_ASTAG: fprintf(f, "AS:i:%d", a->s); break;
            case F_CGTAG_X: fprintf(f, "cg:Z:"); cigar_ops(f, s1, s2, a, 1, 1, 1, 0, 0); break;
            case F_CGTAG_M: fprintf(f, "cg:Z:"); cigar_ops(f, s1, s2, a, 0, 1, 1, 0, 0); break;
            case F_BSTART1: fprintf(f, "%u", strand2 == strand1 ? start1 : start1 + height - 1); break;      /* genpafStart1Blast genpaf.c:748 */
            case F_BEND1: fprintf(f, "%u", strand2 == strand1 ? start1 + height - 1 : start1); break;
            case F_EVALUE: fprintf(f, "%.2g", 3.0e9 * exp(-(a->s * 0.0205) * log(2))); break;                 /* blastz_score_to_ncbi_expectation dna_utilities.c:2346 */
            case F_BITSCORE: fprintf(f, "%.1f", a->s * 0.0205); break;
            case F_NUMBER: fprintf(f, "%llu", (unsigned long long)ordinal + 1); break;
#define FRACTION(n, d) fprintf(f, "%llu/%llu", (unsigned long long)(n), (unsigned long long)(d))
#define PERCENT(n, d) do { if (d) fprintf(f, "%.1f%%", (100.0 * (n)) / (d)); else fprintf(f, "NA"); } while (0)
            case F_IDENTITY: FRACTION(idNumer, idDenom); fputc('\t', f); PERCENT(idNumer, idDenom); break;
            case F_IDFRAC: FRACTION(idNumer, idDenom); break;
            case F_IDPCT: PERCENT(idNumer, idDenom); break;
            case F_BLASTIDPCT: if (conDenom) fprintf(f, "%.2f", (100.0 * idNumer) / conDenom); else fprintf(f, "NA"); break;
            case F_COVERAGE: FRACTION(covNumer, covDenom); fputc('\t', f); PERCENT(covNumer, covDenom); break;
            case F_COVFRAC: FRACTION(covNumer, covDenom); break;
            case F_COVPCT: PERCENT(covNumer, covDenom); break;
            case F_CONTINUITY: FRACTION(conNumer, conDenom); fputc('\t', f); PERCENT(conNumer, conDenom); break;
            case F_CONFRAC: FRACTION(conNumer, conDenom); break;
            case F_CONPCT: PERCENT(conNumer, conDenom); break;
            case F_GAPRATE: FRACTION(gapNumer, gapDenom); fputc('\t', f); PERCENT(gapNumer, gapDenom); break;
        }
    }
    fputc('\n', f);
}

void lzb_fieldlist_match(FILE* f, const lzb_fieldlist* fl, const lzb_seq* s1, const lzb_seq* s2, const lzb_segment* g, uint64_t* number) {
    lzb_editscript es = { 1, 1, LZB_OP_SUB, { LZB_OP_SUB | (g->length << 2) } };
    lzb_alignel al; memset(&al, 0, sizeof al);
    al.beg1 = g->pos1 + 1; al.end1 = g->pos1 + g->length; al.beg2 = g->pos2 + 1; al.end2 = g->pos2 + g->length; al.s = g->s; al.script = &es;
    lzb_fieldlist_align(f, fl, s1, s2, &al, number);
}

/* --format=cigar (print_cigar_align with info, letters before counts, cigar.c:135; output.c:658) */
void lzb_cigar_align(FILE* f, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel* a) {
    lzb_seqview w1, w2; lzb_seq_view(s1, a->beg1 - 1, &w1); lzb_seq_view(s2, a->beg2 - 1, &w2);
    const char* name1 = w1.name && w1.name[0] ? w1.name : "seq1"; const char* name2 = w2.name && w2.name[0] ? w2.name : "seq2";
    const uint32_t beg1 = a->beg1 - 1, beg2 = a->beg2 - 1, height = a->end1 - beg1, width = a->end2 - beg2;
    uint32_t start1, end1, start2, end2; char strand1, strand2;
    if (!(s1->revCompFlags & LZB_RCF_REV)) { start1 = beg1 - 1 - w1.offset + w1.startLoc; end1 = start1 + height; strand1 = '+'; }
    else { start1 = w1.startLoc + w1.len + w1.offset - (beg1 + 1); end1 = start1 - height; strand1 = '-'; }
    if (!(s2->revCompFlags & LZB_RCF_REV)) { start2 = beg2 - 1 - w2.offset + w2.startLoc; end2 = start2 + width; strand2 = '+'; }
    else { start2 = w2.startLoc + w2.len + w2.offset - (beg2 + 1); end2 = start2 - width; strand2 = '-'; }
    fprintf(f, "cigar: %s%s %u %u %c %s%s %u %u %c %d", name2, RCF_SUFFIX[s2->revCompFlags & 3], start2, end2, strand2,
            name1, RCF_SUFFIX[s1->revCompFlags & 3], start1, end1, strand1, a->s);
    cigar_ops(f, s1, s2, a, 0, 0, 1, 0, 0);
    fputc('\n', f);
}
void lzb_cigar_match(FILE* f, const lzb_seq* s1, const lzb_seq* s2, const lzb_segment* g) {
    lzb_editscript es = { 1, 1, LZB_OP_SUB, { LZB_OP_SUB | (g->length << 2) } };
    lzb_alignel al; memset(&al, 0, sizeof al);
    al.beg1 = g->pos1 + 1; al.end1 = g->pos1 + g->length; al.beg2 = g->pos2 + 1; al.end2 = g->pos2 + g->length; al.s = g->s; al.script = &es;
    lzb_cigar_align(f, s1, s2, &al);
}

/* ---- --format=sam / softsam [+eqx] [-] (sam.c:196-560, :680-760) ---- */
static int samSqPending = 0;                /* @SQ lines go out before the first record, not up front (sam.c:213-250: print_sam_header is reached from print_align_list / print_match, output.c:561, :748) */
void lzb_sam_header(FILE* f, const lzb_seq* s1) {
    (void)s1;
    fprintf(f, "@HD\tVN:1.0\tSO:unsorted\n");
    samSqPending = 1;
}

static void sam_sequence_lines(FILE* f, const lzb_seq* s1) {
    samSqPending = 0;
    if (s1->npart == 0) fprintf(f, "@SQ\tSN:%s\tLN:%u\n", s1->shortHeader && s1->shortHeader[0] ? s1->shortHeader : "seq1", s1->trueLen);
    else for (uint32_t k = 0; k < s1->npart; k++) fprintf(f, "@SQ\tSN:%s\tLN:%u\n", s1->part[k].shortHeader, s1->part[k].trueLen);
}

void lzb_sam_align(FILE* f, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel* a, int markMismatches, int softMasked) {
    if (s1->revCompFlags != LZB_RCF_FORWARD) lzb_die("attempt to print - strand or complement for sequence 1 in print_sam_align");
    if (samSqPending) sam_sequence_lines(f, s1);
    lzb_seqview w1, w2; lzb_seq_view(s1, a->beg1 - 1, &w1); lzb_seq_view(s2, a->beg2 - 1, &w2);
    const char* name1 = w1.name && w1.name[0] ? w1.name : "seq1"; const char* name2 = w2.name && w2.name[0] ? w2.name : "seq2";
    const uint32_t beg1 = a->beg1, beg2 = a->beg2, width = a->end2 - beg2 + 1;
    const uint32_t start1 = beg1 - 1 - w1.offset + w1.startLoc;
    uint32_t start2, end2; int flag;
    if (!(s2->revCompFlags & LZB_RCF_REV)) { start2 = beg2 - 1 - w2.offset + w2.startLoc; end2 = start2 - 1 + width; flag = 0; }
    else { start2 = w2.startLoc + w2.offset + (w2.len - beg2) - (width - 1); end2 = w2.startLoc + w2.offset + (w2.len - beg2); flag = 16; }   /* BAM_FREVERSE */
    fprintf(f, "%s\t%d\t%s\t%u\t%d\t", name2, flag, name1, start1, 255);
    const char clip = softMasked ? 'S' : 'H';
    uint32_t pre = start2 > 1 ? start2 - 1 : 0, post = end2 < w2.trueLen ? w2.trueLen - end2 : 0;
    if (s2->revCompFlags & LZB_RCF_REV) { uint32_t t = pre; pre = post; post = t; }
    if (pre) fprintf(f, "%u%c", pre, clip);
    walker w; walk_start(&w, a);
    while (walk_more(&w)) {
        uint32_t run = walk_subs(&w);
        if (run > 0) {
            if (markMismatches) mismatchy_run(f, s1->v + beg1 - 1 + w.i, s2->v + beg2 - 1 + w.j, run, 1, 0, 0, 0);
            else fprintf(f, "%uM", run);
            w.i += run; w.j += run;
        }
        if (!walk_more(&w) || w.k >= w.sc->len) break;
        uint32_t di, dj; walk_gap(&w, &di, &dj);
        if (di) fprintf(f, "%uD", di);
        if (dj) fprintf(f, "%uI", dj);
    }
    if (post) fprintf(f, "%u%c", post, clip);
    fprintf(f, "\t*\t0\t0\t");
    /* print_query_bases sam.c:680: the aligned bases in upper case; with soft clipping the rest of the read around them in lower case */
    const uint32_t pos2 = beg2 - 1, qstart = pos2 - w2.offset + w2.startLoc, qend = qstart - 1 + width;
    if (softMasked && qstart > 1) {
        if (qstart - 1 > pos2 - w2.offset) lzb_die("softsam needs the whole read: %s was loaded as a subrange", name2);
        for (uint32_t x = 0; x < qstart - 1; x++) { uint8_t c = s2->v[pos2 - (qstart - 1) + x]; fputc(c >= 'A' && c <= 'Z' ? c + 32 : c, f); }
    }
    for (uint32_t x = 0; x < width; x++) { uint8_t c = s2->v[pos2 + x]; fputc(c >= 'a' && c <= 'z' ? c - 32 : c, f); }
    if (softMasked && qend < w2.trueLen) {
        if (w2.trueLen - (qstart - 1) > w2.len - (pos2 - w2.offset)) lzb_die("softsam needs the whole read: %s was loaded as a subrange", name2);
        for (uint32_t x = width; x < w2.trueLen - (qstart - 1); x++) { uint8_t c = s2->v[pos2 + x]; fputc(c >= 'A' && c <= 'Z' ? c + 32 : c, f); }
    }
    if (!s2->vq) fprintf(f, "\t*\n");
    else {                                                       /* print_query_quals sam.c:764: the same stretch of the quality string, as is */
        fputc('\t', f);
        if (softMasked && qstart > 1) for (uint32_t x = 0; x < qstart - 1; x++) fputc(s2->vq[pos2 - (qstart - 1) + x], f);
        for (uint32_t x = 0; x < width; x++) fputc(s2->vq[pos2 + x], f);
        if (softMasked && qend < w2.trueLen) for (uint32_t x = width; x < w2.trueLen - (qstart - 1); x++) fputc(s2->vq[pos2 + x], f);
        fputc('\n', f);
    }
}
void lzb_sam_match(FILE* f, const lzb_seq* s1, const lzb_seq* s2, const lzb_segment* g, int markMismatches, int softMasked) {
    lzb_editscript es = { 1, 1, LZB_OP_SUB, { LZB_OP_SUB | (g->length << 2) } };
    lzb_alignel al; memset(&al, 0, sizeof al);
    al.beg1 = g->pos1 + 1; al.end1 = g->pos1 + g->length; al.beg2 = g->pos2 + 1; al.end2 = g->pos2 + g->length; al.s = g->s; al.script = &es;
    lzb_sam_align(f, s1, s2, &al, markMismatches, softMasked);
}

/* ---- --format=rdotplot[+score] and --rdotplot[+score]=<file>: every gap-free block of an alignment as a line segment
 * for R's plot(): "start1 start2 / end1 end2 / NA NA" (genpafRDotplotKeys "02!13!XX" genpaf.h:121, blocks from
 * print_genpaf_align_list_segments genpaf.c:433, coordinates from print_genpaf_match :1458-1490), preceded by the
 * two sequence names whenever they change (print_header output.c:459-478) ---- */
static void rdotplot_rows(FILE* f, lzb_rdotplot* st, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel* a, const lzb_scoreset* ss, int withScore, int ownScore) {
    const char* name1 = s1->npart == 0 && s1->shortHeader && s1->shortHeader[0] ? s1->shortHeader : "seq1";
    const char* name2 = s2->npart == 0 && s2->shortHeader && s2->shortHeader[0] ? s2->shortHeader : "seq2";
    if (st->limited && st->blocksLeft == 0) return;             /* print_match returns before anything is printed, the header included (output.c:744-750) */
    if (strcmp(name1, st->prev1) || strcmp(name2, st->prev2)) {
        fprintf(f, withScore ? "%s\t%s\tscore\n" : "%s\t%s\n", name1, name2);
        snprintf(st->prev1, sizeof st->prev1, "%s", name1); snprintf(st->prev2, sizeof st->prev2, "%s", name2);
    }
    walker w; walk_start(&w, a);
    while (walk_more(&w)) {
        if (st->limited) { if (st->blocksLeft == 0) break; st->blocksLeft--; }
        const uint32_t pos1 = a->beg1 - 1 + w.i, pos2 = a->beg2 - 1 + w.j;
        uint32_t run = walk_subs(&w);
        w.i += run; w.j += run;
        if (walk_more(&w) && w.k < w.sc->len) { uint32_t di, dj; walk_gap(&w, &di, &dj); }
        else if (walk_more(&w)) { w.i = w.height; w.j = w.width; }
        lzb_seqview w1, w2; lzb_seq_view(s1, pos1, &w1); lzb_seq_view(s2, pos2, &w2);
        uint32_t d1, e1, d2, e2;
        if (!(s1->revCompFlags & LZB_RCF_REV)) { d1 = s1->npart == 0 ? pos1 - w1.offset + w1.startLoc : pos1 + 1; e1 = d1 + run - 1; }
        else { d1 = s1->npart == 0 ? (w1.startLoc + w1.len + w1.offset - pos1) - 1 : (w1.offset - 1) + (w1.offset + w1.len) + 1 - pos1; e1 = d1 - run + 1; }
        if (!(s2->revCompFlags & LZB_RCF_REV)) { d2 = s2->npart == 0 ? pos2 - w2.offset + w2.startLoc : pos2 + 1; e2 = d2 + run - 1; }
        else { d2 = s2->npart == 0 ? (w2.startLoc + w2.len + w2.offset - pos2) - 1 : (w2.offset - 1) + (w2.offset + w2.len) + 1 - pos2; e2 = d2 - run + 1; }
        if (!withScore) fprintf(f, "%u\t%u\n%u\t%u\nNA\tNA\n", d1, d2, e1, e2);
        else {
            int32_t sc = ownScore ? a->s : 0;                    /* an HSP keeps its score (output.c:930); a block of an alignment is scored, score_match sequences.c:9682 */
            for (uint32_t x = 0; x < run && !ownScore; x++) sc += ss->sub[(uint32_t)s1->v[pos1 + x] * 256 + s2->v[pos2 + x]];
            fprintf(f, "%u\t%u\t%d\n%u\t%u\t%d\nNA\tNA\tNA\n", d1, d2, sc, e1, e2, sc);
        }
    }
}
void lzb_rdotplot_match(FILE* f, lzb_rdotplot* st, const lzb_seq* s1, const lzb_seq* s2, const lzb_segment* g, const lzb_scoreset* ss, int withScore) {
    lzb_editscript es = { 1, 1, LZB_OP_SUB, { LZB_OP_SUB | (g->length << 2) } };
    lzb_alignel al; memset(&al, 0, sizeof al);
    al.beg1 = g->pos1 + 1; al.end1 = g->pos1 + g->length; al.beg2 = g->pos2 + 1; al.end2 = g->pos2 + g->length; al.s = g->s; al.script = &es;
    rdotplot_rows(f, st, s1, s2, &al, ss, withScore, 1);
}
void lzb_rdotplot_align(FILE* f, lzb_rdotplot* st, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel* a, const lzb_scoreset* ss, int withScore) {
    rdotplot_rows(f, st, s1, s2, a, ss, withScore, 0);
}
