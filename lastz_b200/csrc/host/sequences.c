/*
 * sequences.c -- minimal sequence ingest for lastz_b200: FASTA and 2bit, with the bracketed
 * actions the hot-path tests need ([a..b], [a,b], [a#len], [unmask]) and 2bit contig selection
 * (file.2bit/name).  Reference: sequences.c:2167 (FASTA), :3677 (2bit), :6700ff (actions),
 * :7511 (rev_comp_sequence).  seq->v keeps the bytes exactly as the reference does: case is
 * preserved (lowercase = soft-masked), N stays N, v[len] = 0.
 */
#include <ctype.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "lzb_host.h"

void lzb_die(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    fprintf(stderr, "FAILURE: "); vfprintf(stderr, fmt, ap); fprintf(stderr, "\n");
    va_end(ap);
    exit(EXIT_FAILURE);
}

struct lzb_seqfile {
    char* filename; char* contigName;
    FILE* f;
    int is2bit, bigEndian, isNib, isFastq;
    int nameParse; char* nickname;     /* [nameparse=...], [nickname=...] */
    uint32_t start, end;        /* 1-based inclusive limits, 0 = none */
    int unmask;
    uint32_t contig;            /* sequences delivered so far */
    uint32_t lastIx;            /* 2bit: index of the sequence delivered last */
    /* 2bit index */
    uint32_t n2; char** names; uint32_t* offsets;
    int pendingCh;
    /* [subset=<file>]: the names of the sequences to deliver (sequences.c "contigs of interest") */
    char** subset; uint32_t nsubset, subsetNext;
    /* [nmask=<file>] / [xmask=<file>] / [softmask=<file>]: intervals to overwrite (mask_sequence sequences.c:6973) */
    char* maskFile[4]; int maskChar[4]; int nmasks;
    int multi;                  /* [multi]: deliver every (selected) sequence of the file as one partitioned sequence */
};

enum { NAME_CORE = 0, NAME_DARKSPACE, NAME_ALNUM, NAME_FULL };      /* how a header becomes a name, see short_header_as */
static char* dupstr(const char* s) { char* d = malloc(strlen(s) + 1); strcpy(d, s); return d; }

static uint32_t rd4(lzb_seqfile* sf) {
    unsigned char b[4];
    if (fread(b, 1, 4, sf->f) != 4) lzb_die("premature end of file in %s", sf->filename);
    if (sf->bigEndian) return ((uint32_t)b[0] << 24) | (b[1] << 16) | (b[2] << 8) | b[3];
    return ((uint32_t)b[3] << 24) | (b[2] << 16) | (b[1] << 8) | b[0];
}

static void parse_actions(lzb_seqfile* sf, char* act) {
    /* comma/bracket separated actions; "a,b" is an interval when both parts are numbers */
    for (char* tok = act; *tok;) {
        char* e = tok; int depth = 0;
        while (*e && !(*e == ']' && depth == 0)) e++;
        char save = *e; *e = 0;
        /* split on commas unless the whole thing is "num,num" */
        unsigned a, b; char x;
        if (sscanf(tok, "%u..%u%c", &a, &b, &x) == 2 || sscanf(tok, "%u,%u%c", &a, &b, &x) == 2) { sf->start = a; sf->end = b; }
        else if (sscanf(tok, "%u#%u%c", &a, &b, &x) == 2) { sf->start = a; sf->end = a + b - 1; }
        else {
            for (char* p = strtok(tok, ","); p; p = strtok(NULL, ",")) {
                if (!strcmp(p, "unmask")) sf->unmask = 1;
                else if (!strcmp(p, "nameparse=full") || !strcmp(p, "fullname") || !strcmp(p, "fullnames")) sf->nameParse = NAME_FULL;   /* sequences.c:8273 */
                else if (!strcmp(p, "nameparse=alphanum") || !strcmp(p, "nameparse=alnum")) sf->nameParse = NAME_ALNUM;
                else if (!strcmp(p, "nameparse=darkspace")) sf->nameParse = NAME_DARKSPACE;
                else if (!strncmp(p, "nickname=", 9)) sf->nickname = dupstr(p + 9);
                else if (!strcmp(p, "multi") || !strcmp(p, "multiple")) sf->multi = 1;
                else if (p[0] == '@') goto subset_file;           /* @<file> is the short spelling of subset=<file> */
                else if (!strncmp(p, "nmask=", 6) || !strncmp(p, "xmask=", 6) || !strncmp(p, "softmask=", 9)) {
                    if (sf->nmasks == 4) lzb_die("too many masking actions for %s", sf->filename);
                    sf->maskChar[sf->nmasks] = p[0] == 'n' ? 'N' : p[0] == 'x' ? 'X' : -1;
                    sf->maskFile[sf->nmasks++] = dupstr(strchr(p, '=') + 1);
                }
                else if (!strncmp(p, "subset=", 7)) {
                subset_file:;
                    const char* nfName = p[0] == '@' ? p + 1 : p + 7;
                    FILE* nf = fopen(nfName, "rt");
                    if (!nf) lzb_die("fopen_or_die failed to open \"%s\" for \"rt\"", nfName);
                    char line[1024];
                    while (fgets(line, sizeof line, nf)) {
                        char* w = line; while (*w == ' ' || *w == '\t') w++;
                        size_t n = strcspn(w, " \t\r\n");
                        if (n == 0 || *w == '#') continue;
                        w[n] = 0;
                        sf->subset = realloc(sf->subset, (sf->nsubset + 1) * sizeof(char*));
                        sf->subset[sf->nsubset++] = dupstr(w);
                    }
                    fclose(nf);
                    if (sf->nsubset == 0) lzb_die("contigs-of-interest file is empty: %s", nfName);      /* sequences.c:1204 */
                }
                else if (sscanf(p, "%u..%u%c", &a, &b, &x) == 2) { sf->start = a; sf->end = b; }
                else if (sscanf(p, "%u#%u%c", &a, &b, &x) == 2) { sf->start = a; sf->end = a + b - 1; }
                else lzb_die("sequence action \"%s\" is not supported by lastz_b200 (file %s)", p, sf->filename);
            }
        }
        if (!save) break;
        tok = e + 1;
        while (*tok == '[') tok++;
    }
}

lzb_seqfile* lzb_seqfile_open(const char* spec) {
    lzb_seqfile* sf = calloc(1, sizeof *sf);
    char* s = dupstr(spec);
    char* br = strchr(s, '[');
    if (br) { *br = 0; parse_actions(sf, br + 1); }
    /* file.2bit/contig */
    char* tb = strstr(s, ".2bit/");
    if (tb) { sf->contigName = dupstr(tb + 6); tb[5] = 0; }
    sf->filename = s;
    if (!strcmp(s, "(stdin)")) { sf->f = stdin; int c = fgetc(stdin); if (c != EOF) ungetc(c, stdin); sf->isFastq = c == '@'; return sf; }    /* the query may come from stdin, as FASTA (lastz.c:8762ff) */
    sf->f = fopen(s, "rb");
    if (!sf->f) lzb_die("fopen_or_die failed to open \"%s\" for \"rb\"", s);
    unsigned char magic[4];
    size_t got = fread(magic, 1, 4, sf->f);
    uint32_t be = got == 4 ? (((uint32_t)magic[0] << 24) | (magic[1] << 16) | (magic[2] << 8) | magic[3]) : 0;
    uint32_t le = got == 4 ? (((uint32_t)magic[3] << 24) | (magic[2] << 16) | (magic[1] << 8) | magic[0]) : 0;
    if (be == 0x1A412743u || le == 0x1A412743u) {
        sf->is2bit = 1; sf->bigEndian = (be == 0x1A412743u);
        rd4(sf);                               /* version */
        sf->n2 = rd4(sf); rd4(sf);             /* count, reserved */
        sf->names = calloc(sf->n2, sizeof(char*)); sf->offsets = calloc(sf->n2, 4);
        for (uint32_t i = 0; i < sf->n2; i++) {
            int nl = fgetc(sf->f);
            if (nl == EOF) lzb_die("premature end of file in %s", s);
            sf->names[i] = calloc((size_t)nl + 1, 1);
            if (fread(sf->names[i], 1, (size_t)nl, sf->f) != (size_t)nl) lzb_die("bad 2bit index in %s", s);
            sf->offsets[i] = rd4(sf);
        }
    } else if (be == 0x6BE93D3Au || le == 0x6BE93D3Au) {   /* nib: nibMagicBig / nibMagicLittle sequences.c:635 */
        sf->isNib = 1; sf->bigEndian = (be == 0x6BE93D3Au);
    } else {
        rewind(sf->f);
        sf->isFastq = got > 0 && magic[0] == '@';           /* the first character decides (sequences.c:9106) */
    }
    return sf;
}

void lzb_seq_view(const lzb_seq* s, uint32_t pos0, lzb_seqview* o) {
    if (s->npart == 0) {
        o->name = (s->shortHeader && s->shortHeader[0]) ? s->shortHeader : NULL;
        o->offset = 0; o->startLoc = s->startLoc; o->len = s->len; o->trueLen = s->trueLen; o->contig = s->contig;
        return;
    }
    uint32_t lo = 0, hi = s->npart;                      /* lookup_partition sequences.c: last partition with sepBefore < pos0+1 */
    while (hi - lo > 1) { uint32_t mid = (lo + hi) / 2; if (s->part[mid].sepBefore <= pos0) lo = mid; else hi = mid; }
    const lzb_partition* p = &s->part[lo];
    o->name = p->shortHeader; o->offset = p->sepBefore + 1; o->startLoc = p->startLoc; o->len = p->sepAfter - o->offset; o->trueLen = p->trueLen; o->contig = p->contig;
}

void lzb_seqfile_close(lzb_seqfile* sf) {
    if (!sf) return;
    if (sf->f) fclose(sf->f);
    for (uint32_t i = 0; i < sf->n2; i++) free(sf->names[i]);
    free(sf->names); free(sf->offsets); free(sf->filename); free(sf->contigName); free(sf);
}

void lzb_seq_free(lzb_seq* s) {
    for (uint32_t k = 0; k < s->npart; k++) { free(s->part[k].header); free(s->part[k].shortHeader); }
    free(s->part);
    free(s->v); free(s->vq); free(s->filename); free(s->header); free(s->shortHeader);
    memset(s, 0, sizeof *s);
}

/* shorten_header sequences.c:5913-6020: the name a sequence goes by in the output.  '>' and blanks are skipped, then
 * "reverse complement of " and "positions <x> of "; the name runs to the first blank, '|' or ':' (default), to the first
 * blank ([nameparse=darkspace]) or over letters, digits and '_' ([nameparse=alphanum]); the file suffixes .nib .2bit
 * .hsx .fasta .fa are dropped except in the alphanumeric mode.  [nameparse=full] keeps the whole header line. */
static char* short_header_as(const char* h, int how) {
    if (how == NAME_FULL) return dupstr(h);              /* the header as it stands (seq->header, '>' included for FASTA) */
    if (*h == '>') h++;
    while (*h == ' ' || *h == '\t') h++;
    if (!strncmp(h, "reverse complement of ", 22)) { h += 22; while (*h == ' ' || *h == '\t') h++; }
    if (!strncmp(h, "positions ", 10)) {
        const char* g = h + 10;
        while (*g == ' ' || *g == '\t') g++;
        while (*g && *g != ' ' && *g != '\t') g++;
        while (*g == ' ' || *g == '\t') g++;
        if (!strncmp(g, "of ", 3)) { h = g + 3; while (*h == ' ' || *h == '\t') h++; }
    }
    size_t n;
    if (how == NAME_ALNUM) n = strspn(h, "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789_");
    else {
        n = strcspn(h, how == NAME_DARKSPACE ? " \t" : " \t|:");
        static const char* const suffix[5] = { ".nib", ".2bit", ".hsx", ".fasta", ".fa" };
        for (int k = 0; k < 5; k++) { size_t sl = strlen(suffix[k]); if (n > sl && !strncmp(h + n - sl, suffix[k], sl)) n -= sl; }
    }
    char* d = malloc(n + 1); memcpy(d, h, n); d[n] = 0; return d;
}
static char* name_for(const lzb_seqfile* sf, const char* h) { return sf->nickname ? dupstr(sf->nickname) : short_header_as(h, sf->nameParse); }

static void apply_limits(lzb_seqfile* sf, lzb_seq* out, uint8_t* all, uint32_t total) {
    if (total == 0 && !sf->start && !sf->end) {          /* an empty record is reported and carried as a sequence of length 0 (sequences.c:2398) */
        out->len = 0; out->startLoc = 1; out->trueLen = 0;
        out->v = malloc(1); out->v[0] = 0;
        if (out->vq) { free(out->vq); out->vq = malloc(1); out->vq[0] = 0; }
        free(all);
        return;
    }
    uint32_t a = sf->start ? sf->start : 1, b = sf->end ? sf->end : total;
    if (a > total) lzb_die("beyond end of sequence in %s (start limit %u, length %u)", sf->filename, a, total);
    if (b > total) lzb_die("beyond end of sequence in %s (end limit %u, length %u)", sf->filename, b, total);
    if (b < a) lzb_die("bad sequence interval in %s (%u..%u)", sf->filename, a, b);
    out->len = b - a + 1; out->startLoc = a; out->trueLen = total;
    out->v = malloc((size_t)out->len + 1);
    memcpy(out->v, all + a - 1, out->len); out->v[out->len] = 0;
    if (out->vq) { uint8_t* allq = out->vq; out->vq = malloc((size_t)out->len + 1); memcpy(out->vq, allq + a - 1, out->len); out->vq[out->len] = 0; free(allq); }
    if (sf->unmask) for (uint32_t i = 0; i < out->len; i++) out->v[i] = (uint8_t)toupper(out->v[i]);
    free(all);
    for (int m = 0; m < sf->nmasks; m++) {              /* mask_sequence: "begin end" lines, 1-based inclusive, full-sequence coordinates */
        FILE* mf = fopen(sf->maskFile[m], "rt");
        if (!mf) lzb_die("fopen_or_die failed to open \"%s\" for \"rt\"", sf->maskFile[m]);
        char line[512]; int lineNum = 0;
        while (fgets(line, sizeof line, mf)) {
            lineNum++;
            char* w = strchr(line, '#'); if (w) *w = 0;
            unsigned long long b, e; char extra;
            int items = sscanf(line, "%llu %llu%c", &b, &e, &extra);
            if (items <= 0) continue;
            if (items == 3 && isspace((unsigned char)extra)) items = 2;
            if (items != 2) lzb_die("bad interval (in %s, line %d): \"%s\"", sf->maskFile[m], lineNum, line);
            if (e < out->startLoc) continue;
            if (b < out->startLoc) b = out->startLoc;
            b -= out->startLoc; e -= out->startLoc - 1;      /* zero-based, half open */
            if (b >= out->len) continue;
            if (e >= out->len) e = out->len;
            for (; b < e; b++) {
                if (sf->maskChar[m] >= 0) out->v[b] = (uint8_t)sf->maskChar[m];
                else if (out->v[b] >= 'A' && out->v[b] <= 'Z') out->v[b] = (uint8_t)(out->v[b] + 'a' - 'A');
            }
        }
        fclose(mf);
    }
}

static int next_fasta(lzb_seqfile* sf, lzb_seq* out) {
    int ch;
    do ch = fgetc(sf->f); while (ch != EOF && isspace(ch));
    if (ch == EOF) return 0;
    size_t hcap = 256, hl = 0; char* hdr = malloc(hcap);
    if (ch == '>') {
        hdr[hl++] = '>';
        while ((ch = fgetc(sf->f)) != EOF && ch != '\n' && ch != '\r') {
            if (hl + 2 > hcap) { hcap *= 2; hdr = realloc(hdr, hcap); }
            hdr[hl++] = (char)ch;
        }
        hdr[hl] = 0;
    } else { ungetc(ch, sf->f); hdr[0] = 0; }
    size_t cap = 1 << 20, n = 0; uint8_t* v = malloc(cap);
    int prev = '\n';
    while ((ch = fgetc(sf->f)) != EOF) {
        if (prev == '\n' && ch == '>') { ungetc(ch, sf->f); break; }
        if (ch == '\n' || ch == '\r') { prev = '\n'; continue; }
        if (isspace(ch) || (ch >= '0' && ch <= '9')) { prev = ch; continue; }    /* char_to_fasta_type sequences.c:580-598: blanks and digits are skipped */
        if (!strchr("ACGTNXacgtnx", ch)) {
            /* the ambiguity codes BDHKMRSVWY need --ambiguous=iupac, which this front end does not have; anything else is never legal (:2476-2485) */
            char what[40];
            if (ch >= 'A' && ch <= 'Z') snprintf(what, sizeof what, "uppercase %c", ch);
            else if (ch >= 'a' && ch <= 'z') snprintf(what, sizeof what, "lowercase %c", ch);
            else snprintf(what, sizeof what, "ascii %02X", ch);
            if (hdr[0]) lzb_die("bad fasta character in %s, %s (%s)\nremove or replace non-ACGTN characters or consider using --ambiguous=iupac", sf->filename, hdr, what);
            lzb_die("bad fasta character in %s (%s)\nremove or replace non-ACGTN characters or consider using --ambiguous=iupac", sf->filename, what);
        }
        if (n + 1 > cap) { cap *= 2; v = realloc(v, cap); }
        v[n++] = (uint8_t)ch; prev = ch;
    }
    if (n > 0x7FFFFFFFu) lzb_die("sequence length %zu exceeds maximum", n);
    if (n == 0) { if (hdr[0]) fprintf(stderr, "WARNING. %s contains an empty sequence:\n%s\n", sf->filename, hdr); else fprintf(stderr, "WARNING. %s contains an empty sequence\n", sf->filename); }
    apply_limits(sf, out, v, (uint32_t)n);
    out->header = hdr; out->shortHeader = name_for(sf, hdr);
    return 1;
}

/* load_fastq_sequence / parse_fastq sequences.c:2540-2900: four-line records -- @header, bases, +[header], qualities.
 * The qualities ride along in seq->vq for the SAM writer. */
static int next_fastq(lzb_seqfile* sf, lzb_seq* out) {
    int ch;
    do ch = fgetc(sf->f); while (ch != EOF && isspace(ch));
    if (ch == EOF) return 0;
    if (ch != '@') lzb_die("bad fastq header character in %s (expected \"@\" but read \"%c\")", sf->filename, ch);
    size_t hcap = 256, hl = 0; char* hdr = malloc(hcap);
    while ((ch = fgetc(sf->f)) != EOF && ch != '\n' && ch != '\r') {
        if (hl + 2 > hcap) { hcap *= 2; hdr = realloc(hdr, hcap); }
        hdr[hl++] = (char)ch;
    }
    hdr[hl] = 0;
    if (ch == '\r') { ch = fgetc(sf->f); if (ch != '\n') ungetc(ch, sf->f); }
    size_t cap = 1 << 12, n = 0; uint8_t* v = malloc(cap);
    while ((ch = fgetc(sf->f)) != EOF && ch != '\n' && ch != '\r') {
        if (!isalpha(ch)) lzb_die("bad fastq nucleotide character in %s, %s (ascii %02X)", sf->filename, hdr, ch);
        if (n + 1 > cap) { cap *= 2; v = realloc(v, cap); }
        v[n++] = (uint8_t)ch;
    }
    if (ch == '\r') { ch = fgetc(sf->f); if (ch != '\n') ungetc(ch, sf->f); }
    ch = fgetc(sf->f);
    if (ch == EOF) lzb_die("premature end of fastq file %s", sf->filename);
    if (ch != '+') lzb_die("bad fastq separator character in %s, %s (expected \"+\" but read \"%c\")", sf->filename, hdr, ch);
    size_t k = 0; int same = 1;
    while ((ch = fgetc(sf->f)) != EOF && ch != '\n' && ch != '\r') { if (k >= hl || hdr[k] != ch) same = 0; k++; }
    if (k != 0 && !(same && k == hl)) lzb_die("fastq mismatch between sequence and quality headers in %s (%s)", sf->filename, hdr);
    if (ch == '\r') { ch = fgetc(sf->f); if (ch != '\n') ungetc(ch, sf->f); }
    size_t q = 0; uint8_t* quals = malloc(n + 1);
    while ((ch = fgetc(sf->f)) != EOF && ch != '\n' && ch != '\r') {
        if (ch < '!' || ch > '~') lzb_die("bad fastq quality character in %s, %s (ascii %02X)", sf->filename, hdr, ch);
        if (q < n) quals[q] = (uint8_t)ch;
        q++;
    }
    if (q != n) lzb_die("fastq quality length (%zu) differs from sequence length (%zu) in %s, %s", q, n, sf->filename, hdr);
    out->vq = quals;
    apply_limits(sf, out, v, (uint32_t)n);
    out->header = hdr; out->shortHeader = name_for(sf, hdr);
    return 1;
}

static int next_2bit(lzb_seqfile* sf, lzb_seq* out) {
    uint32_t ix;
    if (sf->contigName) {
        if (sf->contig > 0) return 0;
        for (ix = 0; ix < sf->n2; ix++) if (!strcmp(sf->names[ix], sf->contigName)) break;
        if (ix == sf->n2) lzb_die("2bit file %s doesn't contain %s", sf->filename, sf->contigName);
    } else if (sf->subset) {
        if (sf->subsetNext >= sf->nsubset) return 0;
        const char* want = sf->subset[sf->subsetNext++];
        for (ix = 0; ix < sf->n2; ix++) if (!strcmp(sf->names[ix], want)) break;
        if (ix == sf->n2) lzb_die("2bit file %s doesn't contain %s", sf->filename, want);
    } else {
        ix = sf->contig;
        if (ix >= sf->n2) return 0;
    }
    sf->lastIx = ix;
    if (fseek(sf->f, (long)sf->offsets[ix], SEEK_SET) != 0) lzb_die("bad 2bit index in %s (can't seek to %u for %s)", sf->filename, sf->offsets[ix], sf->names[ix]);
    uint32_t dna = rd4(sf);
    uint32_t nb = rd4(sf);
    if (nb > dna) lzb_die("bad 2bit block count in %s, %s", sf->filename, sf->names[ix]);
    uint32_t* ns = malloc(((size_t)nb * 2 + 1) * 4);
    for (uint32_t i = 0; i < nb; i++) ns[i] = rd4(sf);
    for (uint32_t i = 0; i < nb; i++) ns[nb + i] = rd4(sf);
    uint32_t mb = rd4(sf);
    if (mb > dna) lzb_die("bad 2bit block count in %s, %s", sf->filename, sf->names[ix]);
    uint32_t* ms = malloc(((size_t)mb * 2 + 1) * 4);
    for (uint32_t i = 0; i < mb; i++) ms[i] = rd4(sf);
    for (uint32_t i = 0; i < mb; i++) ms[mb + i] = rd4(sf);
    rd4(sf);
    /* the N and mask blocks must lie inside the sequence: the tables come from the file */
    for (uint32_t k = 0; k < nb; k++) if ((uint64_t)ns[k] + ns[nb + k] > dna) lzb_die("bad 2bit N block in %s, %s (%u+%u > %u)", sf->filename, sf->names[ix], ns[k], ns[nb + k], dna);
    for (uint32_t k = 0; k < mb; k++) if ((uint64_t)ms[k] + ms[mb + k] > dna) lzb_die("bad 2bit mask block in %s, %s (%u+%u > %u)", sf->filename, sf->names[ix], ms[k], ms[mb + k], dna);
    size_t pb = ((size_t)dna + 3) / 4;
    uint8_t* packed = malloc(pb + 1);
    if (fread(packed, 1, pb, sf->f) != pb) lzb_die("premature end of file in %s", sf->filename);
    uint8_t* v = malloc((size_t)dna + 1);
    static const char code[4] = { 'T', 'C', 'A', 'G' };
    for (uint32_t i = 0; i < dna; i++) v[i] = (uint8_t)code[(packed[i >> 2] >> (6 - 2 * (i & 3))) & 3];
    for (uint32_t k = 0; k < nb; k++) for (uint32_t i = 0; i < ns[nb + k]; i++) v[ns[k] + i] = 'N';
    for (uint32_t k = 0; k < mb; k++) for (uint32_t i = 0; i < ms[mb + k]; i++) v[ms[k] + i] = (uint8_t)tolower(v[ms[k] + i]);
    free(packed); free(ns); free(ms);
    apply_limits(sf, out, v, dna);
    out->header = dupstr(sf->names[ix]); out->shortHeader = name_for(sf, sf->names[ix]);
    return 1;
}

/* load_nib_sequence sequences.c:3418-3560: length, then two bases per byte, high nybble first;
 * nybbles 0..4 = T C A G N, bit 3 = soft mask (lower case), everything else X */
static int next_nib(lzb_seqfile* sf, lzb_seq* out) {
    if (sf->contig > 0) return 0;
    uint32_t length = rd4(sf);
    if (length == 0 || length == 0xFFFFFFFFu) lzb_die("bad nib length in %s (%08X)", sf->filename, length);
    size_t nb = ((size_t)length + 1) / 2;
    uint8_t* packed = malloc(nb + 1);
    if (fread(packed, 1, nb, sf->f) != nb) lzb_die("premature end of file in %s", sf->filename);
    static const char code[17] = "TCAGNXXXtcagnxxx";
    uint8_t* v = malloc((size_t)length + 1);
    for (uint32_t i = 0; i < length; i++) v[i] = (uint8_t)code[(i & 1) ? (packed[i >> 1] & 15) : (packed[i >> 1] >> 4)];
    free(packed);
    apply_limits(sf, out, v, length);
    char hdr[1200];
    snprintf(hdr, sizeof hdr, "%s:%u-%u", sf->filename, out->startLoc, out->startLoc + out->len - 1);
    out->header = dupstr(hdr); out->shortHeader = name_for(sf, hdr);
    return 1;
}

static int next_single(lzb_seqfile* sf, lzb_seq* out) {
    memset(out, 0, sizeof *out);
    for (;;) {
        if (!sf->is2bit && sf->subset && sf->subsetNext >= sf->nsubset) return 0;      /* every wanted sequence has been delivered */
        int ok = sf->is2bit ? next_2bit(sf, out) : sf->isNib ? next_nib(sf, out) : sf->isFastq ? next_fastq(sf, out) : next_fasta(sf, out);
        if (!ok) {
            if (!sf->is2bit && sf->subset && sf->subsetNext < sf->nsubset)
                lzb_die("%s does not contain (or contains out of order)\n         the sequence \"%s\"", sf->filename, sf->subset[sf->subsetNext]);
            return 0;
        }
        sf->contig++;
        if (sf->is2bit || !sf->subset) break;
        /* FASTA: the names file is followed in order -- skip ahead to the next wanted name (find_next_fasta_coi sequences.c:5160ff) */
        if (sf->subsetNext < sf->nsubset && !strcmp(sf->subset[sf->subsetNext], out->shortHeader)) { sf->subsetNext++; break; }
        lzb_seq_free(out); memset(out, 0, sizeof *out);
    }
    out->contig = sf->is2bit && (sf->contigName || sf->subset) ? sf->lastIx + 1 : sf->contig;   /* ordinal within the file (sequences.c:3677ff) */
    out->filename = dupstr(sf->filename);
    out->revCompFlags = LZB_RCF_FORWARD;
    return 1;
}

/* rev_comp_sequence sequences.c:7511-7559 with nuc_to_complement dna_utilities.c:96-114 */
void lzb_seq_revcomp(lzb_seq* s) {
    static uint8_t comp[256]; static int init = 0;
    if (!init) {
        for (int i = 0; i < 256; i++) comp[i] = (uint8_t)i;
        const char* a = "ACGTRYMKBDHVNSW", *b = "TGCAYRKMVHDBNSW";
        for (int i = 0; a[i]; i++) { comp[(int)a[i]] = (uint8_t)b[i]; comp[tolower(a[i])] = (uint8_t)tolower(b[i]); }
        init = 1;
    }
    /* a partitioned sequence is reverse-complemented partition by partition, in place (sequences.c:7540-7552) */
    for (uint32_t k = 0; k < (s->npart ? s->npart : 1u); k++) {
        uint8_t* v = s->npart ? s->v + s->part[k].sepBefore + 1 : s->v;
        uint32_t n = s->npart ? s->part[k].sepAfter - (s->part[k].sepBefore + 1) : s->len;
        for (uint32_t i = 0, j = n ? n - 1 : 0; n && i <= j; i++, j--) {
            uint8_t x = comp[v[i]], y = comp[v[j]];
            v[i] = y; v[j] = x;
            if (j == 0) break;
        }
        if (s->vq) {                                         /* qualities are reversed, not complemented (sequences.c:7538) */
            uint8_t* q = s->npart ? s->vq + s->part[k].sepBefore + 1 : s->vq;
            for (uint32_t i = 0, j = n; i + 1 < j; i++) { j--; uint8_t x = q[i]; q[i] = q[j]; q[j] = x; }
        }
    }
    s->revCompFlags ^= LZB_RCF_REVCOMP;
}

/* load_sequence with doPartitioning (sequences.c:1840-1960): all remaining sequences, NUL-separated */
int lzb_seqfile_next(lzb_seqfile* sf, lzb_seq* out) {
    if (!sf->multi) return next_single(sf, out);
    lzb_seq one;
    if (!next_single(sf, &one)) return 0;
    memset(out, 0, sizeof *out);
    size_t cap = (size_t)one.len + 1024, n = 0;
    uint8_t* v = malloc(cap + 2);
    uint8_t* vq = one.vq ? malloc(cap + 2) : NULL;
    if (vq) vq[n] = 0;
    v[n++] = 0;
    do {
        if (n + one.len + 2 > cap) { cap = (n + one.len + 2) * 2; v = realloc(v, cap + 2); if (vq) vq = realloc(vq, cap + 2); }
        out->part = realloc(out->part, (out->npart + 1) * sizeof(lzb_partition));
        lzb_partition* p = &out->part[out->npart++];
        p->sepBefore = (uint32_t)(n - 1);
        memcpy(v + n, one.v, one.len);
        if (vq) { if (!one.vq) lzb_die("%s mixes sequences with and without qualities", sf->filename); memcpy(vq + n, one.vq, one.len); vq[n + one.len] = 0; }
        n += one.len;
        p->sepAfter = (uint32_t)n; v[n++] = 0;
        p->contig = one.contig; p->startLoc = one.startLoc; p->trueLen = one.trueLen;
        p->header = one.header; p->shortHeader = one.shortHeader; one.header = one.shortHeader = NULL;
        if (!out->filename) { out->filename = one.filename; one.filename = NULL; }
        lzb_seq_free(&one);
        if (n > 0x7FFFFFF0u) lzb_die("sequences in %s are too long to be combined with [multi]", sf->filename);
    } while (next_single(sf, &one));
    out->v = v; out->vq = vq; out->len = (uint32_t)(n - 1);            /* the last NUL is the terminator: v[len] == 0 */
    out->startLoc = 1; out->trueLen = out->len; out->contig = 1; out->revCompFlags = LZB_RCF_FORWARD;
    out->header = dupstr(""); out->shortHeader = dupstr("");
    return 1;
}
