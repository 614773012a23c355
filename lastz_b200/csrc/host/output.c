/*
 * output.c -- the wire formats on the output side of the hot path: LAV (lav.c:40-360) and the
 * segments file (--format=segments; output.c fmtSegments -> genpaf.c with the fields
 * name1 start1 end1 name2 start2 end2 strand2 score).
 */
#include <ctype.h>
#include <string.h>
#include "lzb_host.h"

void lzb_lav_job_header(FILE* f, const char* prog, const char* name1, const char* name2,
                        const char* args, const lzb_scoreset* ss, const char* K, const char* L) {
    const char* nuc = "ACGT";
    fprintf(f, "#:lav\nd {\n  \"%s %s %s %s\n", prog, name1, name2, args);
    /* print_score_matrix dna_utilities.c: column header, then one row per nucleotide */
    fprintf(f, "  ");
    for (int c = 0; c < 4; c++) fprintf(f, "%s%4c", c ? " " : "", nuc[c]);
    fprintf(f, "\n");
    for (int r = 0; r < 4; r++) {
        fprintf(f, "  ");
        for (int c = 0; c < 4; c++) fprintf(f, "%s%4d", c ? " " : "", ss->sub[nuc[r] * 256 + nuc[c]]);
        fprintf(f, "\n");
    }
    fprintf(f, "  O = %d, E = %d, K = %s, L = %s, M = %d\"\n}\n", ss->gapOpen, ss->gapExtend, K, L, 0);
}

void lzb_lav_strand_header(FILE* f, const lzb_seq* s1, const lzb_seq* s2) {
    if (s1->npart || s2->npart) lzb_die("lav format can't handle multi-sequences");
    static const char* shortSfx[4] = { "", "~", "~-", "-" };
    static const char* longSfx[4] = { "", "~", "~ (reverse complement)", " (reverse complement)" };
    fprintf(f, "#:lav\ns {\n");
    fprintf(f, "  \"%s%s\" %u %u %d %u\n", s1->filename, shortSfx[s1->revCompFlags], s1->startLoc,
            s1->startLoc + s1->len - 1, (s1->revCompFlags & LZB_RCF_REV) ? 1 : 0, s1->contig);
    fprintf(f, "  \"%s%s\" %u %u %d %u\n", s2->filename, shortSfx[s2->revCompFlags], s2->startLoc,
            s2->startLoc + s2->len - 1, (s2->revCompFlags & LZB_RCF_REV) ? 1 : 0, s2->contig);
    fprintf(f, "}\nh {\n   \"%s%s\"\n   \"%s%s\"\n}\n", s1->header, longSfx[s1->revCompFlags],
            s2->header, longSfx[s2->revCompFlags]);
}

/* print_lav_align lav.c:187-245: one l-line per gap-free run, percent identity rounded */
void lzb_lav_align(FILE* f, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel* a) {
    uint32_t beg1 = a->beg1, beg2 = a->beg2, end1 = a->end1, end2 = a->end2;
    uint32_t height = end1 - beg1 + 1, width = end2 - beg2 + 1;
    const lzb_editscript* sc = a->script;
    fprintf(f, "a {\n  s %d\n  b %u %u\n  e %u %u\n", a->s, beg1, beg2, end1, end2);
    uint32_t k = 0;
    for (uint32_t i = 0, j = 0; i < height || j < width;) {
        uint32_t pi = i, pj = j, run = 0, match = 0;
        const uint8_t* p = s1->v + beg1 + i - 1; const uint8_t* q = s2->v + beg2 + j - 1;
        while (k < sc->len && (sc->op[k] & 3) == LZB_OP_SUB) {
            uint32_t rpt = sc->op[k] >> 2; k++; run += rpt;
            while (rpt-- > 0) { if (toupper(*p) == toupper(*q)) match++; p++; q++; }
        }
        i += run; j += run;
        int pct = run ? (int)((200ull * match + run) / (2ull * run)) : 0;
        fprintf(f, "  l %u %u %u %u %d\n", beg1 + pi, beg2 + pj, beg1 + i - 1, beg2 + j - 1, pct);
        if (i < height || j < width) {
            if (k < sc->len) {
                uint32_t op = sc->op[k] & 3, rpt = sc->op[k] >> 2;
                if (op == LZB_OP_INS) j += rpt; else if (op == LZB_OP_DEL) i += rpt;
                k++;
            }
        }
    }
    fprintf(f, "}\n");
}

/* percent_identical sequences.c:9623-9659 */
static int pct_identical(const uint8_t* a, const uint8_t* b, uint32_t len) {
    uint32_t m = 0, d = 0;
    for (uint32_t i = 0; i < len; i++) {
        int x = lzb_nuc_to_bits[a[i]], y = lzb_nuc_to_bits[b[i]];
        if (x >= 0 && y >= 0) { if (x == y) m++; d++; }
    }
    return d ? (int)((200ull * m + d) / (2ull * d)) : 0;
}

void lzb_lav_match(FILE* f, const lzb_seq* s1, const lzb_seq* s2, const lzb_segment* g) {
    uint32_t e1 = g->pos1 + g->length, e2 = g->pos2 + g->length;
    int pct = g->length ? pct_identical(s1->v + g->pos1, s2->v + g->pos2, g->length) : 0;
    fprintf(f, "a {\n  s %d\n  b %u %u\n  e %u %u\n  l %u %u %u %u %d\n}\n", g->s,
            g->pos1 + 1, g->pos2 + 1, e1, e2, g->pos1 + 1, g->pos2 + 1, e1, e2, pct);
}

void lzb_lav_footer(FILE* f) { fprintf(f, "m {\n  n 0\n}\n#:eof\n"); }

/* (--format=general / mapping / cigar / segments / sam / paf / blastn / rdotplot: general.c) */

/* ---- --format=maf- (MAF blocks without the parameter header), print_maf_align maf.c:271-470,
 * unpartitioned sequences ---- */
static int digits_of(uint32_t a, uint32_t b) { uint32_t m = a > b ? a : b; int d = 1; while (m >= 10) { m /= 10; d++; } return d; }
static char toprint(uint8_t c) { return (c >= 0x20 && c < 0x7F) ? (char)c : '*'; }      /* dna_toprint dna_utilities.h:305 */

/* one text row of an alignment (row 0: sequence 1, row 1: sequence 2), gaps as '-' (maf.c:392-470, axt.c:196-253) */
static void align_text_row(FILE* f, int row, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel* a) {
    const lzb_editscript* sc = a->script;
    uint32_t beg1 = a->beg1, beg2 = a->beg2, height = a->end1 - beg1 + 1, width = a->end2 - beg2 + 1, k = 0;
    for (uint32_t i = 0, j = 0; i < height || j < width;) {
        uint32_t run = 0;
        while (k < sc->len && (sc->op[k] & 3) == LZB_OP_SUB) { run += sc->op[k] >> 2; k++; }
        const uint8_t* p = (row == 0 ? s1->v + beg1 - 1 + i : s2->v + beg2 - 1 + j);
        for (uint32_t x = 0; x < run; x++) fputc(toprint(p[x]), f);
        i += run; j += run;
        if (i < height || j < width) {
            if (k >= sc->len) break;
            uint32_t op = sc->op[k] & 3, rpt = sc->op[k] >> 2; k++;
            if (op == LZB_OP_DEL) { for (uint32_t x = 0; x < rpt; x++) fputc(row == 0 ? toprint(s1->v[beg1 - 1 + i + x]) : '-', f); i += rpt; }
            else if (op == LZB_OP_INS) { for (uint32_t x = 0; x < rpt; x++) fputc(row == 0 ? '-' : toprint(s2->v[beg2 - 1 + j + x]), f); j += rpt; }
        }
    }
    fputc('\n', f);
}

void lzb_maf_align(FILE* f, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel* a) {
    lzb_seqview w1, w2; lzb_seq_view(s1, a->beg1 - 1, &w1); lzb_seq_view(s2, a->beg2 - 1, &w2);
    const char* name1 = w1.name ? w1.name : "seq1"; const char* name2 = w2.name ? w2.name : "seq2";
    static const char* rcfSuffix[4] = { "", "~", "~", "" };                             /* maf.c / cigar.c:191 */
    const char* suff1 = rcfSuffix[s1->revCompFlags & 3]; const char* suff2 = rcfSuffix[s2->revCompFlags & 3];
    uint32_t beg1 = a->beg1, beg2 = a->beg2, height = a->end1 - beg1 + 1, width = a->end2 - beg2 + 1;
    uint32_t start1, start2; char strand1, strand2;
    if (!(s1->revCompFlags & LZB_RCF_REV)) { start1 = beg1 - 1 - w1.offset + w1.startLoc; strand1 = '+'; }
    else { start1 = beg1 - 1 - w1.offset + w1.trueLen + 2 - (w1.startLoc + w1.len); strand1 = '-'; }
    if (!(s2->revCompFlags & LZB_RCF_REV)) { start2 = beg2 - 1 - w2.offset + w2.startLoc; strand2 = '+'; }
    else { start2 = beg2 - 1 - w2.offset + w2.trueLen + 2 - (w2.startLoc + w2.len); strand2 = '-'; }
    int len1 = (int)(strlen(name1) + strlen(suff1)), len2 = (int)(strlen(name2) + strlen(suff2));
    int nameW = len1 >= len2 ? len1 : len2;
    int startW = digits_of(start1, start2), endW = digits_of(height, width), lenW = digits_of(w1.trueLen, w2.trueLen);
    fprintf(f, "a score=%d\n", a->s);
    for (int row = 0; row < 2; row++) {
        if (row == 0) fprintf(f, "s %s%s%*s%*u %*u %c %*u ", name1, suff1, nameW + 1 - len1, " ", startW, start1 - 1, endW, height, strand1, lenW, w1.trueLen);
        else fprintf(f, "s %s%s%*s%*u %*u %c %*u ", name2, suff2, nameW + 1 - len2, " ", startW, start2 - 1, endW, width, strand2, lenW, w2.trueLen);
        align_text_row(f, row, s1, s2, a);
    }
    fputc('\n', f);
}

/* ---- --format=axt, print_axt_align axt.c:96-255 (unpartitioned sequences).  The block number runs over the
 * whole output (axtAlignmentNumber).  The comment header carries the same parameters as the reference's. ---- */
void lzb_axt_header(FILE* f, const char* prog, const char* args, const lzb_scoreset* ss, const char* K, const char* L, int32_t X, int32_t Y) {
    static const char acgt[4] = { 'A', 'C', 'G', 'T' };
    fprintf(f, "# %s %s\n#\n# hsp_threshold      = %s\n# gapped_threshold   = %s\n# x_drop             = %d\n# y_drop             = %d\n"
               "# gap_open_penalty   = %d\n# gap_extend_penalty = %d\n", prog, args, K, L, X, Y, ss->gapOpen, ss->gapExtend);
    fprintf(f, "#        A    C    G    T\n");
    for (int r = 0; r < 4; r++) {
        fprintf(f, "#   %c", acgt[r]);
        for (int c = 0; c < 4; c++) fprintf(f, " %4d", ss->sub[(uint32_t)acgt[r] * 256 + acgt[c]]);
        fprintf(f, "\n");
    }
}

void lzb_axt_align(FILE* f, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel* a, uint64_t* number) {
    lzb_seqview w1, w2; lzb_seq_view(s1, a->beg1 - 1, &w1); lzb_seq_view(s2, a->beg2 - 1, &w2);
    const char* name1 = w1.name ? w1.name : "seq1"; const char* name2 = w2.name ? w2.name : "seq2";
    uint32_t height = a->end1 - a->beg1 + 1, width = a->end2 - a->beg2 + 1;
    uint32_t start1 = a->beg1 - 1 - w1.offset + w1.startLoc, start2; char strand2;
    if (!(s2->revCompFlags & LZB_RCF_REV)) { start2 = a->beg2 - 1 - w2.offset + w2.startLoc; strand2 = '+'; }
    else { start2 = a->beg2 - 1 - w2.offset + w2.trueLen + 2 - (w2.startLoc + w2.len); strand2 = '-'; }
    fprintf(f, "%llu %s %u %u %s %u %u %c %d\n", (unsigned long long)(*number)++, name1, start1, start1 + height - 1,
            name2, start2, start2 + width - 1, strand2, a->s);
    align_text_row(f, 0, s1, s2, a);
    align_text_row(f, 1, s1, s2, a);
    fputc('\n', f);
}

/* ---- --format=gfa (gfa.c:95-330, unpartitioned sequences): d/z job lines, s/h per strand pair, one `a` line per
 * gap-free block (an `A` line in front of a gapped alignment's blocks) ---- */
void lzb_gfa_job_header(FILE* f, const char* prog, const char* name1, const char* name2, const char* seedPattern, int withTrans, uint32_t step) {
    fprintf(f, "d %s %s %s\n", prog, name1, name2);
    fprintf(f, "z seed=%s%s\n", seedPattern, withTrans == 0 ? "" : withTrans == 1 ? " w/transition" : " w/2 transitions");   /* print_options lastz.c:10441 */
    fprintf(f, "z step=%u\n", step);
}

void lzb_gfa_strand_header(FILE* f, const lzb_seq* s1, const lzb_seq* s2) {
    if (s1->npart || s2->npart) lzb_die("gfa format can't handle multi-sequences");
    static const char* shortSuffix[4] = { "", "~", "~-", "-" };
    static const char* longSuffix[4] = { "", "~", "~ (reverse complement)", " (reverse complement)" };
    fprintf(f, "s \"%s%s\" %u %u %d %u \"%s%s\" %u %u %d %u\n",
            s1->filename, shortSuffix[s1->revCompFlags & 3], s1->startLoc, s1->startLoc + s1->len - 1, (s1->revCompFlags & LZB_RCF_REV) ? 1 : 0, s1->contig,
            s2->filename, shortSuffix[s2->revCompFlags & 3], s2->startLoc, s2->startLoc + s2->len - 1, (s2->revCompFlags & LZB_RCF_REV) ? 1 : 0, s2->contig);
    fprintf(f, "h \"%s%s\" \"%s%s\"\n", s1->header ? s1->header : "(no header)", longSuffix[s1->revCompFlags & 3],
            s2->header ? s2->header : "(no header)", longSuffix[s2->revCompFlags & 3]);
}

static void gfa_block(FILE* f, const lzb_seq* s1, uint32_t pos1, const lzb_seq* s2, uint32_t pos2, uint32_t length, int32_t s) {
    int pct = length ? pct_identical(s1->v + pos1, s2->v + pos2, length) : 0;
    fprintf(f, "a %u%s/%u%s %u %d %d ; diag %lld\n", pos1 + 1, (s1->revCompFlags & LZB_RCF_REV) ? "-" : "+",
            pos2 + 1, (s2->revCompFlags & LZB_RCF_REV) ? "-" : "+", length, s, pct, (long long)pos1 - (long long)pos2);
}

void lzb_gfa_match(FILE* f, const lzb_seq* s1, const lzb_seq* s2, const lzb_segment* g) {
    gfa_block(f, s1, g->pos1, s2, g->pos2, g->length, g->s);
}

void lzb_gfa_align(FILE* f, const lzb_seq* s1, const lzb_seq* s2, const lzb_alignel* a, const lzb_scoreset* ss) {
    uint32_t beg1 = a->beg1, beg2 = a->beg2, height = a->end1 - beg1 + 1, width = a->end2 - beg2 + 1;
    const lzb_editscript* sc = a->script;
    for (int pass = 0; pass < 2; pass++) {              /* pass 0: the alignment's score from its blocks and gaps; pass 1: the blocks */
        int32_t total = 0; uint32_t k = 0;
        for (uint32_t i = 0, j = 0; i < height || j < width;) {
            uint32_t run = 0, pi = i, pj = j;
            while (k < sc->len && (sc->op[k] & 3) == LZB_OP_SUB) { run += sc->op[k] >> 2; k++; }
            int32_t bs = 0;
            for (uint32_t x = 0; x < run; x++) bs += ss->sub[(uint32_t)s1->v[beg1 - 1 + pi + x] * 256 + s2->v[beg2 - 1 + pj + x]];
            i += run; j += run; total += bs;
            if (pass == 1) gfa_block(f, s1, beg1 - 1 + pi, s2, beg2 - 1 + pj, run, bs);
            if (i < height || j < width) {
                if (k >= sc->len) break;
                uint32_t op = sc->op[k] & 3, rpt = sc->op[k] >> 2; k++;
                if (op == LZB_OP_INS) j += rpt; else if (op == LZB_OP_DEL) i += rpt;
                if (rpt > 0) total -= ss->gapOpen + (int32_t)rpt * ss->gapExtend;
            }
        }
        if (pass == 0) fprintf(f, "A %u%s/%u%s %u/%u %d\n", beg1, (s1->revCompFlags & LZB_RCF_REV) ? "-" : "+",
                               beg2, (s2->revCompFlags & LZB_RCF_REV) ? "-" : "+", height, width, total);
    }
}
