/*
 * lastz_main.c -- the `lastz` command line (subset) over the lastz_b200 C-ABI.
 *
 * Mirrors the driver in the reference's lastz.c: option parsing (parse_options_loop :5357ff),
 * derived defaults (:9313-9339), the query/strand loop (main :1453-1760), start_one_strand
 * (:3006) and finish_one_strand (:3262).  All seed-and-extend work is done by whichever library
 * implementing include/lastz_b200.h this binary is linked with: liblastz_b200.so (CUDA, the
 * product, `lastz_b200`) or oracle/liblzb_oracle.so (CPU restatement, test tool `lastz_oracle`).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include "lzb_host.h"

typedef struct {
    const char* targetSpec; const char* querySpec;
    const char* seedPattern; int withTrans, haveTrans;
    uint32_t step; int haveStep;
    int whichStrand;               /* 0 plus, 1 both, -1 minus */
    int gfExtend, gfMismatches, gapped, entropy, chain, selfCompare, inhibitTrivial, allBounds, trimToPeak;
    int recoverSeeds;              /* --recoverseeds: process_for_recoverable_hit + merge_segments (lastz.c:5712-5720, :2791, :2811) */
    int twins, twinMinGap, twinMaxGap, seedQueue;    /* --twins=<min>..<max>, --seedqueue= (lastz.c:5671-5710, :9826-9850) */
    double queryDepth; int depthWarn, depthKeep; uint64_t maxPairedBases;   /* --querydepth=[keep:|nowarn:|keep,nowarn:|discard:]<depth> (lastz.c:6063-6105) */
    uint32_t hspLimit; int hspLimitWarn, hspLimitKeep;   /* --queryhsplimit=[keep:|nowarn:|keep,nowarn:|warn:]<n> (lastz.c:5988-6040, :3139-3151) */
    int gpus;                      /* --gpus=<n>: the query cut into n intervals, one process and one device each (an addition) */
    int32_t K, L, X, Y, O, E; int haveK, haveL, haveX, haveY, haveO, haveE;
    int adaptive; double adaptFraction; uint32_t adaptBases;   /* K=top<N>% ('P') or K=top<bases> ('C'), string_to_score_thresh dna_utilities.c:2248 */
    uint32_t tracebackBytes;
    int hashBits;
    const char* scoresFile; const char* segmentsFile; const char* outputFile;
    int format;                    /* 0 lav, 1 segments, 2 general, 3 general-, 4 maf-, 5 axt, 6 gfa, 7 cigar, 8 sam */
    int samSoft, samEqx, samHeader, blastHeader;
    int formatIsSegments, haveGappedOption;   /* --format=segments stops at HSPs unless a gapped option says otherwise (lastz.c:8940, :9053) */
    int dotScore; const char* dotplotFile; int dotplotFileScore;   /* --format=rdotplot[+score] (format 9), --rdotplot[+score]=<file> */
    lzb_filters filters;           /* --filter=identity:.. and friends */
    int unitScores; int32_t unitMatch, unitMismatch;   /* --match=<reward>[,<penalty>] lastz.c:6138 */
    lzb_fieldlist* fields;         /* columns of --format=general[-][:<names>] / mapping[-] */
    int device, showStats, speculation, mafHeader;
    uint32_t wordCountLimit; float wordCountKeep;        /* --maxwordcount=<limit>[%] (lastz.c:6509-6545) */
    int anyOrNone;                                           /* --anyornone: hspImmediate + searchLimit 1 (lastz.c:5962) */
    int nIsAmbiguous; int32_t ambiMatch, ambiMismatch;     /* --ambiguous=n[,[<match>,]<penalty>] lastz.c:5767-5852 */
    int chainDiag, chainAnti;
    char args[4096];
} options;

static int starts(const char* a, const char* p) { return strncmp(a, p, strlen(p)) == 0; }

static long unitized_thousands(const char* s) {   /* string_to_unitized_int, units of 1000 */
    char* e; double v = strtod(s, &e);
    if (*e == 'K' || *e == 'k') { v *= 1000; e++; } else if (*e == 'M' || *e == 'm') { v *= 1000 * 1000; e++; } else if (*e == 'G' || *e == 'g') { v *= 1000.0 * 1000 * 1000; e++; }
    if (e == s || *e) lzb_die("\"%s\" is not an integer", s);
    return (long)v;
}

static long unitized(const char* s) {       /* string_to_unitized_int, units of 1024 */
    char* e; double v = strtod(s, &e);
    if (*e == 'K' || *e == 'k') v *= 1024; else if (*e == 'M' || *e == 'm') v *= 1024 * 1024;
    else if (*e == 'G' || *e == 'g') v *= 1024.0 * 1024 * 1024;
    return (long)v;
}

/* <min>[..<max>] in percent, each with an optional % sign (lastz.c:6680-6725); ..<max> and <min>.. are allowed */
static void percent_range(const char* a, const char* v, float* lo, float* hi) {
    float mn = 0.0f, mx = 100.0f; char text[64]; char* e;
    if (!strcmp(v, "..") || strlen(v) >= sizeof text) lzb_die("Can't understand \"%s\"", a);
    strcpy(text, v);
    char* dots = strstr(text, ".."); char* upper = NULL;
    if (dots) { *dots = 0; upper = dots + 2; }
    if (text[0]) { mn = strtof(text, &e); if (e == text) lzb_die("Can't understand \"%s\"", a); if (*e == '%') e++; if (*e && !(!dots && !strcmp(e, "."))) lzb_die("Can't understand \"%s\"", a); }
    if (upper && upper[0]) { mx = strtof(upper, &e); if (e == upper) lzb_die("Can't understand \"%s\"", a); if (*e == '%') e++; if (*e) lzb_die("Can't understand \"%s\"", a); }
    if (mn < 0 || mx > 100 || mn > mx) lzb_die("Can't understand \"%s\"", a);
    *lo = mn / 100.0; *hi = mx / 100.0;
}

static void parse_options(options* o, int argc, char** argv) {
    memset(o, 0, sizeof *o);
    lzb_filters_init(&o->filters);
    o->withTrans = 1; o->step = 1; o->whichStrand = 1; o->gfExtend = LZB_GFEX_XDROP; o->gapped = 1;
    o->entropy = 1; o->trimToPeak = 1; o->tracebackBytes = 80u * 1024 * 1024; o->hashBits = 16;
    o->speculation = 256;
    char* wordSeed = NULL;
    /* shortcuts that stand for a string of options (expanders[] lastz.c:559-577, current versions); the command line
     * recorded for the headers keeps the shortcut itself */
    static const struct { const char* name; const char* expansion; } shortcuts[] = {
        { "--yasra98", "T=2 Z=20 --match=1,6 O=8 E=1 Y=20 K=22 L=30 --identity=98..100 --ambiguous=n --noytrim" },
        { "--yasra95", "T=2 Z=20 --match=1,5 O=8 E=1 Y=20 K=22 L=30 --identity=95..100 --ambiguous=n --noytrim" },
        { "--yasra90", "T=2 Z=20 --match=1,5 O=6 E=1 Y=20 K=22 L=30 --identity=90..100 --ambiguous=n --noytrim" },
        { "--yasra85", "T=2 --match=1,2 O=4 E=1 Y=20 K=22 L=30 --identity=85..100 --ambiguous=n --noytrim" },
        { "--yasra75", "T=2 --match=1,1 O=3 E=1 Y=20 K=22 L=30 --identity=75..100 --ambiguous=n --noytrim" },
        { "--yasra95short", "T=2 --match=1,7 O=6 E=1 Y=14 K=10 L=14 --identity=95..100 --ambiguous=n --noytrim" },
        { "--yasra85short", "T=2 --match=1,3 O=4 E=1 Y=14 K=11 L=14 --identity=85..100 --ambiguous=n --noytrim" },
    };
    const char** words = malloc(((size_t)argc + 1) * 16 * sizeof *words); char* silent = calloc(((size_t)argc + 1) * 16, 1); int nwords = 0;
    for (int i = 1; i < argc; i++) {
        int sc = -1;
        for (size_t k = 0; k < sizeof shortcuts / sizeof shortcuts[0]; k++) if (!strcmp(argv[i], shortcuts[k].name)) sc = (int)k;
        if (sc < 0) { words[nwords++] = argv[i]; continue; }
        if (strlen(o->args) + strlen(argv[i]) + 2 < sizeof o->args) { strcat(o->args, argv[i]); strcat(o->args, " "); }
        char* copy = strdup(shortcuts[sc].expansion);
        for (char* tok = strtok(copy, " "); tok; tok = strtok(NULL, " ")) { silent[nwords] = 1; words[nwords++] = tok; }
    }
    for (int i = 0; i < nwords; i++) {
        const char* a = words[i]; const char* v = strchr(a, '='); v = v ? v + 1 : "";
        if (a[0] != '-' && !(strlen(a) > 1 && a[1] == '=' && strchr("CTWKLXYOEZQ", a[0]))) {
            if (!o->targetSpec) o->targetSpec = a; else if (!o->querySpec) o->querySpec = a;
            else lzb_die("Can't understand \"%s\"", a);
            continue;
        }
        if (!strcmp(a, "--version")) { printf("lastz_b200 (the seed-and-extend path of lastz 1.04.58 on sm_100a; back end: %s)\n", lzb_backend()); exit(0); }
        if (!strcmp(a, "--help") || starts(a, "--help=")) {
            printf("usage: lastz_b200 target[actions] [query[actions]] [options]\n"
                   "  sequences  FASTA, FASTQ, 2bit (file.2bit/contig), nib; query from stdin when omitted\n"
                   "  actions    [a..b] [a#len] [unmask] [multi] [subset=<file>] [@<file>] [nmask=|xmask=|softmask=<file>]\n"
                   "             [nameparse=darkspace|alphanum|full] [fullname] [nickname=<name>]\n"
                   "  seeding    --seed=<pattern>|12of19|14of22|match<N>  W=<N>  T=0..4  --[no]transition[=2]  --step=<N>  --strand=both|plus|minus  --self\n"
                   "  extension  --[no]gfextend  --xdrop=  --hspthresh=<score>|top<N>%%|top<bases>  --exact=<N>  --mismatch=<M>,<N>  --[no]entropy\n"
                   "             --[no]gapped  --ydrop=  --gappedthresh=  --gap=<open>,<extend>  --scores=<file>  --match=<reward>[,<penalty>]\n"
                   "             --ambiguous=n[,..]  --noytrim  --allgappedbounds  --notrivial  --allocate:traceback=<bytes>  --[no]chain[=<diag>,<anti>]\n"
                   "             --segments=<file>  --anyornone  --justhits  --querydepth=[keep:]<depth>  --yasra98|95|90|85|75|95short|85short\n"
                   "  filters    --identity=  --coverage=  --continuity=  --matchcount=  --filter=nmismatch:0..<n>|ngap:0..<n>|cgap:0..<n>\n"
                   "  output     --format=lav|axt|maf[-]|gfa|segments|cigar|general[-][:<fields>]|mapping[-]|sam[-]|softsam[-]|paf[:wfmash]|blastn[-]|rdotplot[+score]\n"
                   "             --rdotplot[+score]=<file>  --output=<file>\n"
                   "  additions  --device=<n>  --gpus=<n>  --diaghash=<bits>  --speculation=<n>  --stats\n"
                   "Option meanings are those of lastz 1.04.58 (INTEGRATION.md section 6 lists what is and is not built).\n");
            exit(0);
        }
        /* (the echo of the command line in headers: --device=<k> is left out -- it names hardware, not the computation, and an
         * interval process of --gpus=<n> must print what a run on that subrange prints) */
        if (!silent[i] && !starts(a, "--device=") && strlen(o->args) + strlen(a) + 2 < sizeof o->args) { strcat(o->args, a); strcat(o->args, " "); }
        if (!strcmp(a, "T=0")) { o->withTrans = 0; o->haveTrans = 1; }
        else if (!strcmp(a, "T=1")) { o->seedPattern = LZB_SEED_12OF19; o->withTrans = 1; }
        else if (!strcmp(a, "T=2")) { o->seedPattern = LZB_SEED_12OF19; o->withTrans = 0; }
        else if (!strcmp(a, "T=3")) { o->seedPattern = LZB_SEED_14OF22; o->withTrans = 1; }
        else if (!strcmp(a, "T=4")) { o->seedPattern = LZB_SEED_14OF22; o->withTrans = 0; }
        else if (starts(a, "--word=")) ;                         /* lastz.c:5665: max index bits; only a table-layout matter (overweight
                                                                    seeds are "resolved", seed_search.c:878) -- this index holds 28 bits anyway */
        else if (!strcmp(a, "--anyornone") || !strcmp(a, "--stopafterone")) o->anyOrNone = 1;
        else if (starts(a, "--queryhsplimit=") || starts(a, "--queryhsplimit+=")) {   /* HSPs allowed per query, both strands together (lastz.c:5988-6049) */
            const char* d = v; o->hspLimitWarn = 1; o->hspLimitKeep = starts(a, "--queryhsplimit+=");
            if (starts(a, "--queryhsplimit=keep,nowarn:")) { o->hspLimitWarn = 0; o->hspLimitKeep = 1; d = v + 12; }
            else if (starts(a, "--queryhsplimit+=nowarn:")) { o->hspLimitWarn = 0; d = v + 7; }
            else if (starts(a, "--queryhsplimit+=warn:")) d = v + 5;
            else if (starts(a, "--queryhsplimit=keep:")) o->hspLimitKeep = 1;       /* (the number is read from the '=', as there: "keep:5" is not an integer) */
            else if (starts(a, "--queryhsplimit=nowarn:")) { o->hspLimitWarn = 0; d = v + 7; }
            else if (starts(a, "--queryhsplimit=warn:")) d = v + 5;
            char* end = NULL; double x = strtod(d, &end);
            if (end == d) lzb_die("\"%s\" is not an integer", d);
            if (*end == 'K' || *end == 'k') { x *= 1000; end++; } else if (*end == 'M' || *end == 'm') { x *= 1000 * 1000; end++; }
            if (*end) lzb_die("\"%s\" is not an integer", d);
            if (x <= 0) lzb_die("--queryhsplimit must be positive");
            o->hspLimit = (uint32_t)x;
        }
        else if (starts(a, "--querydepth=")) {                     /* depth in units of the query length; K/M suffixes as in string_to_unitized_double */
            const char* d = v; o->depthWarn = 1; o->depthKeep = 0;
            if (starts(v, "nowarn:")) { o->depthWarn = 0; d = v + 7; }
            else if (starts(v, "keep,nowarn:")) { o->depthWarn = 0; o->depthKeep = 1; d = v + 12; }
            else if (starts(v, "keep:")) { o->depthKeep = 1; d = v + 5; }
            else if (starts(v, "discard:")) d = v + 8;
            char* end = NULL; double x = strtod(d, &end);
            if (end == d) lzb_die("\"%s\" is not a number", d);
            if (*end == 'K' || *end == 'k') { x *= 1000; end++; } else if (*end == 'M' || *end == 'm') { x *= 1000 * 1000; end++; } else if (*end == 'G' || *end == 'g') { x *= 1000.0 * 1000 * 1000; end++; }
            if (*end) lzb_die("\"%s\" is not a number", d);
            o->queryDepth = x < 0 ? 0 : x;
        }
        else if (!strcmp(a, "--notwins")) o->twins = 0;
        else if (starts(a, "--twins=")) {                        /* <min>..<max>, the historical <min>:<max>, or <max> alone */
            const char* sep = strstr(v, ".."); int w = 2;
            if (!sep) { sep = strchr(v, ':'); w = 1; }
            o->twins = 1;
            if (sep) { o->twinMinGap = atoi(v); o->twinMaxGap = atoi(sep + w); }      /* (atoi stops at the separator) */
            else { o->twinMinGap = 0; o->twinMaxGap = atoi(v); }
        }
        else if (starts(a, "--seedqueue=")) o->seedQueue = atoi(v);
        else if (!strcmp(a, "--recoverseeds") || !strcmp(a, "--recoverhits")) o->recoverSeeds = 1;
        else if (!strcmp(a, "--norecoverseeds") || !strcmp(a, "--norecoverhits")) o->recoverSeeds = 0;
        else if (!strcmp(a, "--justhits") || !strcmp(a, "--hitsonly")) { o->gfExtend = LZB_GFEX_NONE; o->gapped = 0; }   /* lastz.c:5875 */
        else if (starts(a, "W=") || starts(a, "--seed=match")) {
            int w = atoi(starts(a, "--seed=match") ? a + 12 : v);
            if (w < 1 || w > 15) lzb_die("%d is not a valid word length", w);
            wordSeed = malloc((size_t)w + 1); memset(wordSeed, '1', (size_t)w); wordSeed[w] = 0;
            o->seedPattern = wordSeed;
            if (!o->haveTrans) { o->withTrans = 0; o->haveTrans = 1; }
        }
        else if (!strcmp(a, "--seed=12of19")) o->seedPattern = LZB_SEED_12OF19;
        else if (!strcmp(a, "--seed=14of22")) o->seedPattern = LZB_SEED_14OF22;
        else if (starts(a, "--seed=")) o->seedPattern = v;
        else if (!strcmp(a, "--notransition") || !strcmp(a, "--notrans")) { o->withTrans = 0; o->haveTrans = 1; }
        else if (!strcmp(a, "--transition")) { o->withTrans = 1; o->haveTrans = 1; }
        else if (starts(a, "--transition=")) { o->withTrans = atoi(v); o->haveTrans = 1; }
        else if (starts(a, "--step=") || starts(a, "Z=")) { o->step = (uint32_t)atoi(v); o->haveStep = 1; }
        else if (starts(a, "--maxwordcount=")) {
            if (strchr(v, ',')) lzb_die("--maxwordcount's max interval (the chasm) is not supported by lastz_b200");
            const size_t n = strlen(v);
            if (n > 0 && v[n - 1] == '%') {
                const double pct = atof(v) / 100.0;
                if (pct <= 0) lzb_die("--maxwordcount cannot be zero");
                if (pct == 1) lzb_die("--maxwordcount cannot be 100%%");
                if (pct > 1) lzb_die("--maxwordcount cannot be more than 100%%");
                o->wordCountKeep = (float)pct; o->wordCountLimit = 0;
            } else {
                const int lim = atoi(v);
                if (lim < 1) lzb_die("--maxwordcount must be at least 1");
                o->wordCountLimit = (uint32_t)lim; o->wordCountKeep = 0;
            }
        }
        else if (!strcmp(a, "--strand=both")) o->whichStrand = 1;
        else if (!strcmp(a, "--strand=plus") || !strcmp(a, "--plus")) o->whichStrand = 0;
        else if (!strcmp(a, "--strand=minus")) o->whichStrand = -1;
        else if (!strcmp(a, "--self")) { o->selfCompare = 1; o->inhibitTrivial = 1; }
        else if (!strcmp(a, "--notrivial")) o->inhibitTrivial = 1;
        else if (starts(a, "--exact=")) {
            if (o->haveK && o->gfExtend == LZB_GFEX_XDROP) lzb_die("can't use %s with --hspthreshold", a);          /* lastz.c:6333-6341 */
            if (o->haveX && o->gfExtend == LZB_GFEX_XDROP) lzb_die("can't use %s with --xdrop", a);
            if (o->haveK && o->gfExtend == LZB_GFEX_MISMATCH) lzb_die("can't use %s with --%dmismatch", a, o->gfMismatches);
            o->gfExtend = LZB_GFEX_EXACT; o->K = atoi(v); o->haveK = 1;
            if (o->K <= 0) lzb_die("%s is not a valid exact match threshold", v);
        }      /* lastz.c:6330-6350 */
        else if (starts(a, "--mismatch=") || (a[0] == '-' && a[1] == '-' && a[2] >= '0' && a[2] <= '9' && strstr(a, "mismatch="))) {
            int M = 0, N = 0;                                    /* --mismatch=M,N or --<M>mismatch=N, lastz.c:6355-6390 */
            if (o->haveK && o->gfExtend == LZB_GFEX_XDROP) lzb_die("can't use %s with --hspthreshold", a);
            if (o->haveX && o->gfExtend == LZB_GFEX_XDROP) lzb_die("can't use %s with --xdrop", a);
            if (o->haveK && o->gfExtend == LZB_GFEX_EXACT) lzb_die("can't use %s with --exact", a);
            if (starts(a, "--mismatch=")) { if (sscanf(v, "%d,%d", &M, &N) != 2) lzb_die("--mismatch requires two values (count and length)"); }
            else { M = atoi(a + 2); N = atoi(v); }
            if (M == 0) o->gfExtend = LZB_GFEX_EXACT;
            else {
                if (M < 1 || M > LZB_GFEX_MISMATCH_MAX) lzb_die("%d is out of range for N-mismatch (valid range is 1..%d)", M, LZB_GFEX_MISMATCH_MAX);
                if (N < M) lzb_die("%d is not a valid exact %dmismatch threshold", N, M);
                o->gfExtend = LZB_GFEX_MISMATCH; o->gfMismatches = M;
            }
            o->K = N; o->haveK = 1;
        }
        else if (!strcmp(a, "--nogfextend")) o->gfExtend = LZB_GFEX_NONE;
        else if (!strcmp(a, "--gfextend")) o->gfExtend = LZB_GFEX_XDROP;
        else if (!strcmp(a, "--nogapped") || !strcmp(a, "--ungapped")) { o->gapped = 0; o->haveGappedOption = 1; }
        else if (!strcmp(a, "--gapped")) { o->gapped = 1; o->haveGappedOption = 1; }
        else if (!strcmp(a, "C=0")) { o->chain = 0; o->gapped = 1; o->haveGappedOption = 1; }
        else if (!strcmp(a, "C=1")) { o->chain = 1; o->gapped = 0; o->haveGappedOption = 1; }
        else if (!strcmp(a, "C=2")) { o->chain = 1; o->gapped = 1; o->haveGappedOption = 1; }
        else if (!strcmp(a, "C=3")) { o->chain = 0; o->gapped = 0; o->haveGappedOption = 1; }
        else if (!strcmp(a, "--chain")) o->chain = 1;
        else if (starts(a, "--chain=")) { o->chain = 1; if (sscanf(v, "%d,%d", &o->chainDiag, &o->chainAnti) != 2) lzb_die("can't understand %s", a); }
        else if (!strcmp(a, "--nochain")) o->chain = 0;
        else if (!strcmp(a, "--ambiguous=n") || !strcmp(a, "--ambiguousn") || !strcmp(a, "--ambig=n")) o->nIsAmbiguous = 1;
        else if (starts(a, "--ambiguous=n,") || starts(a, "--ambig=n,")) {
            const char* p1 = strchr(a, ',') + 1; const char* p2 = strchr(p1, ',');
            o->nIsAmbiguous = 1;
            if (p2) { o->ambiMatch = atoi(p1); o->ambiMismatch = atoi(p2 + 1); } else { o->ambiMatch = 0; o->ambiMismatch = atoi(p1); }
            if (o->ambiMismatch < 0) lzb_die("penalty for --ambiguous=n must be non-negative");
        }
        else if (starts(a, "--identity=") || starts(a, "--filter=identity:")) percent_range(a, starts(a, "--filter=") ? strchr(a, ':') + 1 : v, &o->filters.minIdentity, &o->filters.maxIdentity);
        else if (starts(a, "--coverage=") || starts(a, "--filter=coverage:")) percent_range(a, starts(a, "--filter=") ? strchr(a, ':') + 1 : v, &o->filters.minCoverage, &o->filters.maxCoverage);
        else if (starts(a, "--continuity=") || starts(a, "--filter=continuity:")) percent_range(a, starts(a, "--filter=") ? strchr(a, ':') + 1 : v, &o->filters.minContinuity, &o->filters.maxContinuity);
        else if (starts(a, "--matchcount=") || starts(a, "--filter=nmatch:")) {                       /* lastz.c:6850-6872 */
            const char* n = starts(a, "--filter=") ? strchr(a, ':') + 1 : v;
            if (n[0] && n[strlen(n) - 1] == '%') lzb_die("lastz_b200 does not implement a match count relative to the sequence length (%s)", a);
            long c = unitized_thousands(n);
            if (c <= 0) lzb_die("--filter=nmatch must be positive");
            o->filters.minMatchCount = (uint32_t)c;
        }
        else if (starts(a, "--filter=nmismatch:..") || starts(a, "--filter=nmismatch:0..")) {
            long c = unitized_thousands(strstr(a, "..") + 2); if (c < 0) lzb_die("--filter=nmismatch can't be negative");
            o->filters.maxMismatchCount = (int32_t)c;
        }
        else if (starts(a, "--filter=ngap:..") || starts(a, "--filter=ngap:0..")) { o->filters.maxSeparateGaps = atoi(strstr(a, "..") + 2); if (o->filters.maxSeparateGaps < 0) lzb_die("--filter=ngap can't be negative"); }
        else if (starts(a, "--filter=cgap:..") || starts(a, "--filter=cgap:0..")) { o->filters.maxGapColumns = atoi(strstr(a, "..") + 2); if (o->filters.maxGapColumns < 0) lzb_die("--filter=cgap can't be negative"); }
        else if (!strcmp(a, "--noentropy")) o->entropy = 0;
        else if (!strcmp(a, "--entropy")) o->entropy = 1;
        else if (!strcmp(a, "--allgappedbounds")) o->allBounds = 1;
        else if (!strcmp(a, "--noytrim")) o->trimToPeak = 0;
        else if (starts(a, "--hspthresh=top") || starts(a, "K=top")) {
            const char* n = v + 3; size_t ln = strlen(n); char* e;
            if (ln > 0 && n[ln - 1] == '%') { o->adaptive = 'P'; o->adaptFraction = strtod(n, &e) / 100.0; if (e != n + ln - 1 || o->adaptFraction < 0) lzb_die("\"%s\" is not a valid percentage", n); }
            else {                                               /* string_to_unitized_int by thousands, utilities.c */
                double c = strtod(n, &e);
                if (*e == 'K' || *e == 'k') { c *= 1000; e++; } else if (*e == 'M' || *e == 'm') { c *= 1000 * 1000; e++; } else if (*e == 'G' || *e == 'g') { c *= 1000.0 * 1000 * 1000; e++; }
                if (e == n || *e || c < 0 || c > 4294967295.0) lzb_die("\"%s\" is not a valid base count", n);
                o->adaptive = 'C'; o->adaptBases = (uint32_t)c;
            }
            o->haveK = 1;
        }
        else if (starts(a, "--hspthresh=") || starts(a, "K=")) {
            if (o->haveK && o->gfExtend == LZB_GFEX_EXACT) lzb_die("can't use %s with --exact", a);               /* lastz.c:6312-6318 */
            if (o->haveK && o->gfExtend == LZB_GFEX_MISMATCH) lzb_die("can't use %s with --%dmismatch", a, o->gfMismatches);
            o->K = atoi(v); o->haveK = 1; o->adaptive = 0;
        }
        else if (starts(a, "--gappedthresh=top") || starts(a, "L=top")) lzb_die("lastz_b200 does not implement an adaptive gapped threshold (%s)", a);
        else if (starts(a, "--gappedthresh=") || starts(a, "L=")) { o->L = atoi(v); o->haveL = 1; }
        else if (starts(a, "--xdrop=") || starts(a, "X=")) {
            if (o->haveK && o->gfExtend == LZB_GFEX_EXACT) lzb_die("can't use %s with --exact", a);               /* lastz.c:6271-6277 */
            if (o->haveK && o->gfExtend == LZB_GFEX_MISMATCH) lzb_die("can't use %s with --%dmismatch", a, o->gfMismatches);
            o->X = atoi(v); o->haveX = 1; o->gfExtend = LZB_GFEX_XDROP;
        }   /* an x-drop value selects x-drop extension, whatever came before (lastz.c:6278) */
        else if (starts(a, "--ydrop=") || starts(a, "Y=")) { o->Y = atoi(v); o->haveY = 1; }
        else if (starts(a, "O=")) { o->O = atoi(v); o->haveO = 1; }
        else if (starts(a, "E=")) { o->E = atoi(v); o->haveE = 1; }
        else if (starts(a, "--gap=")) { if (sscanf(v, "%d,%d", &o->O, &o->E) != 2) lzb_die("can't understand %s", a); o->haveO = o->haveE = 1; }
        else if (starts(a, "--scores=") || starts(a, "Q=")) o->scoresFile = v;
        else if (starts(a, "--match=")) {
            const char* comma = strchr(v, ',');
            o->unitScores = 1; o->unitMatch = atoi(v); o->unitMismatch = comma ? -atoi(comma + 1) : -o->unitMatch;
            if (o->unitMatch <= 0) lzb_die("%s is not a valid match score", v);
            if (o->unitMismatch >= 0) lzb_die("%s is not a valid mismatch penalty", comma + 1);
        }
        else if (starts(a, "--segments=") || starts(a, "--anchors=")) o->segmentsFile = v;   /* --anchors: the older spelling, lastz.c:5855 */
        else if (starts(a, "--allocate:traceback=") || starts(a, "--traceback=")) o->tracebackBytes = (uint32_t)unitized(v);
        else if (starts(a, "--output=")) o->outputFile = v;
        else if (!strcmp(a, "--format=lav") || !strcmp(a, "--lav")) o->format = 0;
        else if (!strcmp(a, "--maf-")) o->format = 4;
        else if (!strcmp(a, "--format=axt") || !strcmp(a, "--axt")) o->format = 5;
        else if (!strcmp(a, "--format=maf") || !strcmp(a, "--maf")) { o->format = 4; o->mafHeader = 1; }
        else if (!strcmp(a, "--format=gfa") || !strcmp(a, "--gfa")) o->format = 6;
        else if (!strcmp(a, "--format=segments")) { o->format = 2; o->formatIsSegments = 1; o->fields = lzb_fieldlist_parse("name1,start1,end1,name2,start2,end2,strand2,score"); }   /* genpafSegmentKeys, lastz.c:7267 */
        else if (!strcmp(a, "--format=general") || !strcmp(a, "--format=gen")) { o->format = 2; o->fields = lzb_fieldlist_standard(); }   /* default fields, genpaf.h:117 */
        else if (!strcmp(a, "--format=general-") || !strcmp(a, "--format=gen-")) { o->format = 3; o->fields = lzb_fieldlist_standard(); } /* ... without the header line */
        else if (starts(a, "--format=general:") || starts(a, "--format=gen:")) { o->format = 2; o->fields = lzb_fieldlist_parse(strchr(a, ':') + 1); }   /* lastz.c:7319 */
        else if (starts(a, "--format=general-:") || starts(a, "--format=gen-:")) { o->format = 3; o->fields = lzb_fieldlist_parse(strchr(a, ':') + 1); }
        else if (!strcmp(a, "--format=mapping")) { o->format = 2; o->fields = lzb_fieldlist_mapping(); }                                 /* lastz.c:7347 */
        else if (!strcmp(a, "--format=mapping-")) { o->format = 3; o->fields = lzb_fieldlist_mapping(); }
        else if (!strcmp(a, "--format=paf") || !strcmp(a, "--format=paf:minimap2")) { o->format = 3; o->fields = lzb_fieldlist_paf(0); }    /* lastz.c:7384 */
        else if (!strcmp(a, "--format=paf:wfmash")) { o->format = 3; o->fields = lzb_fieldlist_paf(1); }
        else if (!strcmp(a, "--format=blastn")) { o->format = 3; o->fields = lzb_fieldlist_blastn(); o->blastHeader = 1; }              /* lastz.c:7365 */
        else if (!strcmp(a, "--format=blastn-")) { o->format = 3; o->fields = lzb_fieldlist_blastn(); }
        else if (!strcmp(a, "--format=rdotplot")) { o->format = 9; o->dotScore = 0; }                                                   /* lastz.c:7405 */
        else if (!strcmp(a, "--format=rdotplot+score")) { o->format = 9; o->dotScore = 1; }
        else if (starts(a, "--rdotplot=")) { if (o->dotplotFile) lzb_die("duplicated or conflicting option \"%s\"", a); o->dotplotFile = v; o->dotplotFileScore = 0; }   /* lastz.c:7489 */
        else if (starts(a, "--rdotplot+score=")) { if (o->dotplotFile) lzb_die("duplicated or conflicting option \"%s\"", a); o->dotplotFile = v; o->dotplotFileScore = 1; }
        else if (!strcmp(a, "--format=cigar") || !strcmp(a, "--cigar")) o->format = 7;
        else if (starts(a, "--format=sam") || starts(a, "--format=softsam") || starts(a, "--sam") || starts(a, "--softsam")) {           /* lastz.c:7170-7248 */
            const char* n = starts(a, "--format=") ? a + 9 : a + 2;
            o->samSoft = starts(n, "soft"); n += o->samSoft ? 7 : 3;
            o->samEqx = starts(n, "+eqx"); if (o->samEqx) n += 4;
            o->samHeader = n[0] != '-';
            if (strcmp(n, "") && strcmp(n, "-")) lzb_die("lastz_b200 does not implement option \"%s\" (seed-and-extend hot path only)", a);
            o->format = 8;
        }                                                  /* lastz.c:7250 */
        else if (!strcmp(a, "--format=maf-")) o->format = 4;             /* MAF blocks, no parameter header */
        /* lastz_b200 additions */
        else if (starts(a, "--device=")) o->device = atoi(v);
        else if (starts(a, "--gpus=")) { o->gpus = atoi(v); if (o->gpus < 1 || o->gpus > 64) lzb_die("--gpus wants a device count from 1 to 64"); }
        else if (starts(a, "--diaghash=")) o->hashBits = atoi(v);
        else if (starts(a, "--speculation=")) o->speculation = atoi(v);
        else if (!strcmp(a, "--stats")) o->showStats = 1;
        else if (!strcmp(a, "--progress") || starts(a, "--progress=") || starts(a, "--progress+") || starts(a, "--verbosity=") || !strcmp(a, "--noruntime")) ;   /* diagnostics on stderr only (lastz.c:7700ff): accepted, nothing to report */
        else lzb_die("lastz_b200 does not implement option \"%s\" (seed-and-extend hot path only)", a);
    }
    if (!o->targetSpec) lzb_die("You must specify a target file");
    if (o->selfCompare && !o->querySpec) o->querySpec = o->targetSpec;
    if (!o->querySpec) o->querySpec = "(stdin)";                    /* lastz.c:8762: no query file => read it from stdin */
    if (o->formatIsSegments && !o->haveGappedOption) o->gapped = 0;
    if (o->formatIsSegments && o->gapped) lzb_die("can't used --writesegments with --gapped");
    if (o->anyOrNone && lzb_filters_active(&o->filters)) lzb_die("lastz_b200 does not combine --anyornone with --filter options yet");
    if (o->adaptive) {
        if (o->gfExtend != LZB_GFEX_XDROP) lzb_die("an adaptive HSP threshold requires --gfextend");   /* the other extensions assume a score, seed_search.c:3003 */
        if (o->anyOrNone) lzb_die("can't use --anyornone with adaptive hsp score threshold");         /* lastz.c:8898 */
        if (o->segmentsFile) lzb_die("lastz_b200 does not combine --segments with an adaptive HSP threshold");
        if (o->selfCompare) lzb_die("lastz_b200 does not combine --self with an adaptive HSP threshold (the reference stops with an internal error there)");
    }
}

/* read_segment_table segment.c:456: rows name1 start1 end1 name2 start2 end2 strand [score] */
static uint64_t read_segments(const char* path, const lzb_seq* t, const lzb_seq* q, lzb_segment** out) {
    FILE* f = fopen(path, "rt");
    if (!f) lzb_die("fopen_or_die failed to open \"%s\" for \"rt\"", path);
    uint64_t n = 0, cap = 0; lzb_segment* g = NULL; char line[1024];
    char want = (q->revCompFlags & LZB_RCF_REV) ? '-' : '+';
    while (fgets(line, sizeof line, f)) {
        if (line[0] == '#') continue;
        char n1[256], n2[256], st; unsigned s1, e1, s2, e2; int sc = 0;
        int items = sscanf(line, "%255s %u %u %255s %u %u %c %d", n1, &s1, &e1, n2, &s2, &e2, &st, &sc);
        if (items < 7) continue;
        if (strcmp(n1, t->shortHeader) || strcmp(n2, q->shortHeader) || st != want) continue;
        if (n == cap) { cap = cap * 2 + 256; g = realloc(g, cap * sizeof *g); }
        memset(&g[n], 0, sizeof *g);
        g[n].pos1 = s1 - t->startLoc; g[n].pos2 = s2 - q->startLoc; g[n].length = e1 + 1 - s1;
        g[n].s = sc; g[n].id = q->revCompFlags; g[n].scoreCov = g[n].length;
        n++;
    }
    fclose(f);
    *out = g; return n;
}

/* --gpus=<n>: the reference treats a query subrange q[a..b] as an independent unit of work, so the query is
 * cut into n intervals and each goes to its own process and device (--device=k), with the target replicated.  The
 * children are this same program run as `target query[a..b] <the other options> --device=k`; their outputs are
 * gathered in interval order, so the result is, byte for byte, what n runs on those subranges print one after the
 * other -- each with its own header lines, and an alignment that would cross a cut ends at it, as in the reference run
 * on the same subranges.  Only for a query file that holds one sequence and carries no actions of its own. */
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>
#include <fcntl.h>
static int run_sharded(int argc, char** argv, const options* o) {
    if (!strcmp(o->querySpec, "(stdin)") || o->selfCompare) lzb_die("--gpus needs a query file (not stdin, not --self)");
    if (strchr(o->querySpec, '[')) lzb_die("--gpus cuts the query itself: the query can't carry actions (\"%s\")", o->querySpec);
    lzb_seqfile* qf = lzb_seqfile_open(o->querySpec);
    lzb_seq q; memset(&q, 0, sizeof q);
    if (!lzb_seqfile_next(qf, &q)) lzb_die("%s contains no sequence", o->querySpec);
    { lzb_seq extra; memset(&extra, 0, sizeof extra); if (lzb_seqfile_next(qf, &extra)) lzb_die("--gpus needs a query file with one sequence (%s holds more)", o->querySpec); }
    const uint64_t len = q.len;
    lzb_seqfile_close(qf);
    const int n = o->gpus;
    if (len < (uint64_t)n) lzb_die("the query (%llu bases) is shorter than the number of intervals (%d)", (unsigned long long)len, n);
    pid_t* pid = calloc((size_t)n, sizeof *pid); char (*tmp)[64] = calloc((size_t)n, 64);
    for (int k = 0; k < n; k++) {
        const uint64_t lo = (uint64_t)k * len / (uint64_t)n, hi = (uint64_t)(k + 1) * len / (uint64_t)n;     /* [lo, hi), 0-based */
        snprintf(tmp[k], 64, "/tmp/lastz_b200_shard_XXXXXX");
        const int fd = mkstemp(tmp[k]);
        if (fd < 0) lzb_die("can't create a temporary file for interval %d", k);
        char** av = calloc((size_t)argc + 3, sizeof *av); int ac = 0;
        char* qspec = malloc(strlen(o->querySpec) + 64), * dev = malloc(32);
        snprintf(qspec, strlen(o->querySpec) + 64, "%s[%llu..%llu]", o->querySpec, (unsigned long long)lo + 1, (unsigned long long)hi);
        snprintf(dev, 32, "--device=%d", o->device + k);
        av[ac++] = argv[0];
        for (int i = 1; i < argc; i++) {
            if (!strncmp(argv[i], "--gpus=", 7) || !strncmp(argv[i], "--device=", 9) || !strncmp(argv[i], "--output=", 9)) continue;
            av[ac++] = argv[i] == o->querySpec ? qspec : argv[i];
        }
        av[ac++] = dev; av[ac] = NULL;
        fflush(NULL);
        pid[k] = fork();
        if (pid[k] < 0) lzb_die("can't start the process for interval %d", k);
        if (pid[k] == 0) { dup2(fd, 1); close(fd); execv("/proc/self/exe", av); fprintf(stderr, "FAILURE: can't re-run %s\n", argv[0]); _exit(127); }
        close(fd);
    }
    int bad = 0;
    for (int k = 0; k < n; k++) { int st = 0; if (waitpid(pid[k], &st, 0) < 0 || !WIFEXITED(st) || WEXITSTATUS(st) != 0) bad = 1; }
    FILE* out = o->outputFile ? fopen(o->outputFile, "w") : stdout;
    if (!out) lzb_die("fopen_or_die failed to open \"%s\" for \"w\"", o->outputFile);
    for (int k = 0; k < n && !bad; k++) {                        /* the gather: interval order */
        FILE* f = fopen(tmp[k], "r"); char buf[1 << 16]; size_t got;
        if (!f) lzb_die("lost the output of interval %d", k);
        while ((got = fread(buf, 1, sizeof buf, f)) > 0) fwrite(buf, 1, got, out);
        fclose(f);
    }
    for (int k = 0; k < n; k++) unlink(tmp[k]);
    if (bad) lzb_die("one of the %d interval processes failed (its message is above)", n);
    if (out != stdout) fclose(out);
    return 0;
}

int main(int argc, char** argv) {
    options o; parse_options(&o, argc, argv);
    if (o.gpus > 1) return run_sharded(argc, argv, &o);
    /* scoring + derived defaults, lastz.c:9127-9339 */
    static lzb_scoreset ss;
    if (o.scoresFile && o.unitScores) lzb_die("can't use --scores and --match together");
    if (o.scoresFile) lzb_scores_read_file(&ss, o.scoresFile);
    else if (o.unitScores) {                                     /* unit scores and what is derived from them, lastz.c:9168-9236, dna_utilities.c:158-162 */
        const int32_t R = o.unitMatch, P = -o.unitMismatch;
        int32_t unit[4][4];
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) unit[r][c] = r == c ? R : -P;
        if (!o.haveO) { o.O = (int32_t)ceil(3.25 * P); o.haveO = 1; }
        if (!o.haveE) { o.E = (int32_t)ceil(0.24375 * P); o.haveE = 1; }
        if (!o.haveK) { o.K = (int32_t)ceil(30.0 * R); o.haveK = 1; }
        if (!o.haveL && o.gfExtend == LZB_GFEX_EXACT) { o.L = (int32_t)ceil(30.0 * R); o.haveL = 1; }
        if (!o.haveX) { o.X = (int32_t)ceil(10.0 * sqrt((double)P)); o.haveX = 1; }
        if (!o.haveY) { o.Y = 2 * o.X; o.haveY = 1; }
        lzb_scores_from_template(&ss, unit, (int32_t)(-10.0 * P), (int32_t)(-1.0 * P), o.O, o.E);
    }
    else lzb_scores_default(&ss);
    if (o.haveO) ss.gapOpen = o.O;
    if (o.haveE) ss.gapExtend = o.E;
    if (o.nIsAmbiguous) {                                        /* ambiguate_n dna_utilities.c:1544 on both matrices (lastz.c:9424-9429) */
        int32_t* mats[2] = { ss.sub, ss.masked };
        for (int m = 0; m < 2; m++) {
            int32_t* sub = mats[m];
            const char* ns = "Nn";
            for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) sub[(uint32_t)ns[i] * 256 + ns[j]] = o.ambiMatch;
            for (const char* c = "ACGT"; *c; c++) for (int lc = 0; lc < 2; lc++) {
                int ch = lc ? *c + 32 : *c;
                for (int i = 0; i < 2; i++) { sub[(uint32_t)ch * 256 + ns[i]] = -o.ambiMismatch; sub[(uint32_t)ns[i] * 256 + ch] = -o.ambiMismatch; }
            }
        }
    }
    if (!o.haveK) o.K = ss.hspThresholdSet ? ss.hspThreshold : 3000;
    if (o.gfExtend == LZB_GFEX_NONE && !o.haveK) o.K = 0;          /* lastz.c:8984-9010 */
    if (!o.haveX) o.X = ss.xDropSet ? ss.xDrop : 10 * ss.sub['A' * 256 + 'A'];
    if (!o.haveY) o.Y = ss.yDropSet ? ss.yDrop : ss.gapOpen + 300 * ss.gapExtend;
    if (!o.haveL) o.L = ss.gappedThresholdSet ? ss.gappedThreshold : (o.gfExtend == LZB_GFEX_XDROP ? o.K : 3000);
    if (!o.haveStep && ss.stepSet) o.step = ss.step;
    lzb_seed seed; lzb_seed_parse(&seed, o.seedPattern ? o.seedPattern : LZB_SEED_12OF19, o.withTrans);

    FILE* out = stdout;
    if (o.outputFile) { out = fopen(o.outputFile, "wt"); if (!out) lzb_die("can't open %s", o.outputFile); }

    lzb_seqfile* tf = lzb_seqfile_open(o.targetSpec);
    lzb_seq target;
    if (!lzb_seqfile_next(tf, &target)) lzb_die("%s contains no sequence", o.targetSpec);
    { lzb_seq extra; if (lzb_seqfile_next(tf, &extra)) lzb_die("%s contains more than one sequence, consider using the \"multiple\" action", o.targetSpec); }

    lzb_ctx* ctx = lzb_open(o.device);
    if (!ctx) lzb_die("%s", lzb_last_error());
    if (lzb_set_scoring(ctx, ss.sub, ss.masked, ss.gapOpen, ss.gapExtend)) lzb_die("%s", lzb_last_error());
    lzb_target* T = lzb_target_build(ctx, target.v, target.len, 0, 0, lzb_upper_nuc_to_bits, &seed, o.step);
    if (!T) lzb_die("%s", lzb_last_error());
    if (o.wordCountKeep > 0) {
        /* find_position_table_limit pos_table.c:2000-2034: the smallest count such that the words occurring at most that
         * often hold the wanted share of all positions (counts taken in increasing order; single-precision product as there) */
        const uint64_t nw = 1ull << seed.weight;
        uint32_t* counts = malloc(nw * sizeof(uint32_t));
        const int64_t npos = lzb_target_export_index(T, counts, NULL);
        if (npos < 0) lzb_die("%s", lzb_last_error());
        uint32_t maxc = 0;
        for (uint64_t w = 0; w < nw; w++) if (counts[w] > maxc) maxc = counts[w];
        uint64_t* occ = calloc((size_t)maxc + 2, sizeof(uint64_t));
        for (uint64_t w = 0; w < nw; w++) occ[counts[w]]++;
        uint32_t numPositions = (uint32_t)npos;
        uint32_t minToKeep = (uint32_t)ceil(numPositions * o.wordCountKeep);
        uint32_t limit = 0;
        for (uint32_t c = 1; c <= maxc; c++) {
            if (!occ[c]) continue;
            const uint32_t held = (uint32_t)(c * occ[c]);
            if (held >= minToKeep) { limit = c; break; }
            minToKeep -= held;
        }
        free(occ); free(counts);
        o.wordCountLimit = limit;
    }
    if (o.wordCountLimit > 0 && lzb_target_limit(T, o.wordCountLimit)) lzb_die("%s", lzb_last_error());

    if (o.hspLimit && (o.adaptive || o.selfCompare || o.anyOrNone || o.segmentsFile)) lzb_die("lastz_b200 does not combine --queryhsplimit with an adaptive threshold, --self, --anyornone or --segments");
    if (o.twins) {                  /* lastz.c:9826-9836 */
        if (o.twinMinGap <= -seed.length) lzb_die("minGap for twins (%d) must be greater than negative of seed length (%d)", o.twinMinGap, -seed.length);
        if (o.twinMaxGap < o.twinMinGap) lzb_die("maxGap for twins (%d) can't be less than min gap (%d)", o.twinMaxGap, o.twinMinGap);
        if (o.gfExtend == LZB_GFEX_EXACT || o.gfExtend == LZB_GFEX_MISMATCH) lzb_die("lastz_b200 implements --twins with x-drop extension or --nogfextend only");
        if (o.selfCompare) lzb_die("lastz_b200 does not implement --twins together with --self");
        if (o.adaptive) lzb_die("lastz_b200 does not implement --twins together with an adaptive --hspthresh");
        if (o.seedQueue < 0) lzb_die("--seedqueue can't be negative");
    }
    if (o.recoverSeeds) {           /* built: x-drop extension or none, a fixed threshold, no --self mirroring of the merged table */
        if (o.gfExtend == LZB_GFEX_EXACT || o.gfExtend == LZB_GFEX_MISMATCH) lzb_die("lastz_b200 implements --recoverseeds with x-drop extension or --nogfextend only");
        if (o.selfCompare) lzb_die("lastz_b200 does not implement --recoverseeds together with --self");
        if (o.adaptive) lzb_die("lastz_b200 does not implement --recoverseeds together with an adaptive --hspthresh");
    }

    /* thresholds as the headers print them (score_thresh_to_string dna_utilities.c:2290): an adaptive one as top<bases>,
     * the percentage already resolved against the target length; an unset L copies K */
    char textK[32], textL[32];
    const uint32_t adaptLimit = o.adaptive == 'P' ? (uint32_t)(o.adaptFraction * target.len + 0.5) : o.adaptBases;
    if (o.adaptive) snprintf(textK, sizeof textK, "top%u", adaptLimit); else snprintf(textK, sizeof textK, "%d", o.K);
    if (o.adaptive && !o.haveL) snprintf(textL, sizeof textL, "top%u", adaptLimit); else snprintf(textL, sizeof textL, "%d", o.L);
    /* job headers name the FILES: no actions, no 2bit contig (seq->filename, lav.c:62, gfa.c:108) */
    char n1[1024], n2[1024];
    { char* cut;
      snprintf(n1, sizeof n1, "%s", o.targetSpec); snprintf(n2, sizeof n2, "%s", o.querySpec);
      if ((cut = strchr(n1, '['))) *cut = 0;
      if ((cut = strchr(n2, '['))) *cut = 0;
      if ((cut = strstr(n1, ".2bit/"))) cut[5] = 0;
      if ((cut = strstr(n2, ".2bit/"))) cut[5] = 0; }
    if (o.format == 0) lzb_lav_job_header(out, "lastz.v1.04.58", n1, n2, o.args, &ss, textK, textL);
    else if (o.format == 2) lzb_fieldlist_header(out, o.fields);
    else if (o.format == 8 && o.samHeader) lzb_sam_header(out, &target);
    else if (o.format == 6) lzb_gfa_job_header(out, "lastz.v1.04.58", n1, n2, o.seedPattern ? o.seedPattern : LZB_SEED_12OF19, seed.withTrans, o.step);
    else if (o.format == 5) lzb_axt_header(out, "lastz.v1.04.58", o.args, &ss, textK, textL, o.X, o.Y);
    else if (o.format == 4 && o.mafHeader) {                     /* maf.c:96-130: version line + the same parameter comments */
        fprintf(out, "##maf version=1 scoring=lastz.v1.04.58\n");
        lzb_axt_header(out, "lastz.v1.04.58", o.args, &ss, textK, textL, o.X, o.Y);
    }
    uint64_t axtNumber = 0, rowNumber = 0;
    lzb_rdotplot dotMain, dotSide; memset(&dotMain, 0, sizeof dotMain); memset(&dotSide, 0, sizeof dotSide);
    FILE* dotOut = NULL;
    if (o.dotplotFile) { dotOut = fopen(o.dotplotFile, "wt"); if (!dotOut) lzb_die("fopen_or_die failed to open \"%s\" for \"wt\"", o.dotplotFile); }

    lzb_seed_stats sst; lzb_gapped_stats gst;
    uint64_t totHits = 0, totCells = 0, totHsps = 0; double seedSec = 0, gapSec = 0; int queriesOverLimit = 0;
    double ks[12] = {0}, gk = 0; uint64_t gExt = 0, gSpec = 0, gRedo = 0, gLaunch = 0, gTrunc = 0;
    lzb_seqfile* qf = lzb_seqfile_open(o.querySpec);
    lzb_seq query;
    while (lzb_seqfile_next(qf, &query)) {
        if (query.len == 0) { lzb_seq_free(&query); continue; }
        if ((query.npart || target.npart) && (o.format == 0 || o.format == 6)) lzb_die("%s format can't handle multi-sequences", o.format == 0 ? "lav" : "gfa");
        if (target.npart) {
            /* a partitioned target: hits, extensions and DP sweeps stop at its NULs like they do in a partitioned query.
             * A query that equals one target partition gets its trivial self-alignment from the library
             * (identical_partition_of_sequence gapped_extend.c:2034, :1185-1230).  Not built: what --notrivial does with
             * partitions (:1131-1141).  Such runs stop here. */
            if (o.selfCompare || o.segmentsFile || o.anyOrNone || o.adaptive || o.inhibitTrivial)
                lzb_die("lastz_b200 does not combine a [multi] target with --self, --notrivial, --segments, --anyornone or an adaptive threshold yet");
            /* both partitioned and identical: one trivial self-alignment per partition (gapped_extend.c:1232-1290), not built */
            if (o.gapped && o.whichStrand >= 0 && query.npart && query.len == target.len && !strncasecmp((const char*)query.v + 1, (const char*)target.v + 1, query.len - 1))
                lzb_die("the [multi] query is identical to the [multi] target; lastz_b200 does not build the trivial self-alignments of partitions yet");
        }
        if (query.npart && (o.selfCompare || o.segmentsFile || o.anyOrNone))
            lzb_die("lastz_b200 does not combine a [multi] query with --self, --segments or --anyornone yet");
        if (query.npart && o.adaptive) lzb_die("lastz_b200 does not combine a [multi] query with an adaptive HSP threshold yet");
        int reported = 0;                                        /* --anyornone: alignments reported for this query */
        if (o.hspLimit) { dotMain.limited = 1; dotMain.blocksLeft = o.hspLimit; }   /* rdotplot prints block by block (deGapifyOutput, lastz.c:7410): the cap counts blocks */
        uint32_t printedForQuery = 0;                            /* --queryhsplimit also caps what is printed per query: HSPs, or (gapped) strand lists (output.c:556-559, :744-747) */
        /* Order of work for one query (main, lastz.c:1566-1700): each strand is searched and finished in turn -- unless the
         * HSP threshold is adaptive: then both strands are searched into ONE table (collectHspsFromBoth :1426), the table
         * is split by strand (split_anchors :1678), and the - strand is finished BEFORE the + strand (:1680-1700).  A limit
         * on the HSPs of a query (--queryhsplimit) also searches both strands first, into separate tables
         * (collectHspsSeparately :1431), so that a query over the limit can be dropped as a whole. */
        struct { int strand, search, finish; } steps[4]; int nsteps = 0;
        const int doPlus = o.whichStrand >= 0, doMinus = o.whichStrand != 0;
        if (!o.adaptive && !o.hspLimit) {
            if (doPlus)  { steps[nsteps].strand = 0; steps[nsteps].search = 1; steps[nsteps++].finish = 1; }
            if (doMinus) { steps[nsteps].strand = 1; steps[nsteps].search = 1; steps[nsteps++].finish = 1; }
        } else {
            if (doPlus)  { steps[nsteps].strand = 0; steps[nsteps].search = 1; steps[nsteps++].finish = 0; }
            if (doMinus) { steps[nsteps].strand = 1; steps[nsteps].search = 1; steps[nsteps++].finish = 0; }
            if (doMinus) { steps[nsteps].strand = 1; steps[nsteps].search = 0; steps[nsteps++].finish = 1; }
            if (doPlus)  { steps[nsteps].strand = 0; steps[nsteps].search = 0; steps[nsteps++].finish = 1; }
        }
        lzb_query* strandQ[2] = { NULL, NULL }; lzb_segment* strandSegs[2] = { NULL, NULL }; uint64_t strandN[2] = { 0, 0 };
        int strandId[2] = { 0, 0 }, orientation = 0, tableSplit = 0, abortQuery = 0, overLimit = 0; int32_t lowAnchorScore = 0;
        lzb_hsptable table, others;                              /* anchors / secondaryAnchors of the adaptive flow */
        if (o.adaptive) {                                        /* resolve_score_thresh dna_utilities.c:2220, limit_segment_table lastz.c:1399 */
            lzb_hsptable_init(&table, adaptLimit);
            lzb_hsptable_init(&others, 0);
        }
        for (int step = 0; step < nsteps; step++) {
            const int pass = steps[step].strand;
            if (o.anyOrNone && reported) continue;               /* the search limit of 1 is per query, both strands */
            if (pass != orientation) { lzb_seq_revcomp(&query); orientation = pass; }
            if (steps[step].search) {
                strandQ[pass] = lzb_query_load(ctx, query.v, query.len);
                if (!strandQ[pass]) lzb_die("%s", lzb_last_error());
                strandId[pass] = query.revCompFlags;
            }
            lzb_query* Q = strandQ[pass];
            lzb_segment* segs = NULL; uint64_t nsegs = 0;
            if (!steps[step].search) ;
            else if (o.segmentsFile) nsegs = read_segments(o.segmentsFile, &target, &query, &segs);
            else {
                lzb_seed_params sp; memset(&sp, 0, sizeof sp);
                sp.gfExtend = o.gfExtend; sp.gfMismatches = o.gfMismatches; sp.xDrop = o.X; sp.hspThreshold = o.K; sp.entropy = o.entropy;
                /* adaptive: every extension is an HSP candidate (seed_search.c:2907 rejects on score only for a fixed
                 * threshold) and the entropy adjustment depends on the table as it stands (below) */
                if (o.adaptive) { sp.hspThreshold = -600000000; sp.entropy = 0; }
                sp.hashBits = o.hashBits; sp.selfCompare = o.selfCompare;
                sp.sameStrand = o.selfCompare && query.revCompFlags == target.revCompFlags;
                sp.strandId = query.revCompFlags;
                sp.plainHits = (o.gfExtend == LZB_GFEX_NONE && !o.gapped);
                sp.recoverSeeds = o.recoverSeeds && !sp.plainHits;        /* the plain processor wins (lastz.c:2789-2792) */
                if (o.twins) {                                            /* the twin processor goes before the others (lastz.c:2787, :9835) */
                    sp.twinMinSpan = 2 * seed.length + o.twinMinGap; sp.twinMaxSpan = 2 * seed.length + o.twinMaxGap;
                    sp.seedQueueSize = o.seedQueue; sp.plainHits = 0; sp.recoverSeeds = 0;
                }
                if (o.hspLimit) {                                         /* start_one_strand lastz.c:3067-3072: what the other strand found counts */
                    const uint64_t prev = pass == 1 && doPlus ? strandN[0] : 0;
                    sp.searchLimit = prev == 0 ? o.hspLimit : prev < o.hspLimit ? o.hspLimit - (uint32_t)prev : 1;
                }
                if (lzb_seed_hit_search(ctx, T, Q, &seed, lzb_upper_nuc_to_bits, &sp, &segs, &nsegs, &sst))
                    lzb_die("%s", lzb_last_error());
                totHits += sst.rawSeedHits; totHsps += sst.hsps; seedSec += sst.seconds;
                for (int z = 0; z < 12; z++) ks[z] += sst.kernelSeconds[z];
            }
            /* --self mirrors every HSP across the main diagonal when the run stops at HSPs (report_hsps /
             * collect_hsps lastz.c:3858-3880, :4050-4075; with gapped extension the mirroring moves to the
             * alignments, lastz.c:9055-9061) */
            if (o.adaptive && steps[step].search) {
                /* collect_hsps -> add_segment in discovery order.  An extension's score is entropy-adjusted only if it is
                 * positive and could make the table as it stands (seed_search.c:2856-2863); its --self mirror follows it
                 * with the same score (lastz.c:4050-4075). */
                const int mirror = o.selfCompare && !o.gapped, same = query.revCompFlags == target.revCompFlags;
                for (uint64_t k = 0; k < nsegs; k++) {
                    lzb_segment g = segs[k];
                    if (o.entropy && g.s > 0 && table.len > 0 && g.s >= table.lowScore)
                        g.s = (int32_t)((double)g.s * lzb_hsp_entropy(target.v + g.pos1, query.v + g.pos2, g.length));
                    lzb_hsptable_add(&table, &g);
                    if (!mirror) continue;
                    uint32_t len = g.length, e1 = g.pos1 + len, e2 = g.pos2 + len, s1, s2;
                    if (same) { s1 = e1; s2 = e2; }
                    else { s1 = target.len - e1 + len; s2 = query.len - e2 + len; if (s2 == e1 && s1 == e2) continue; }
                    g.pos1 = s2 - len; g.pos2 = s1 - len; g.hspId = 0;
                    lzb_hsptable_add(&table, &g);
                }
                lzb_free(segs); segs = NULL; nsegs = 0;
            }
            if (o.hspLimit && steps[step].search) {                 /* :3139-3151 */
                const uint64_t prev = pass == 1 && doPlus ? strandN[0] : 0;
                if (nsegs + prev > o.hspLimit) {
                    overLimit = 1;
                    if (!o.hspLimitKeep) {                            /* the query is dropped as a whole: nothing of either strand is reported */
                        lzb_free(segs); lzb_free(strandSegs[0]); strandSegs[0] = NULL;
                        if (pass == 1 && doPlus) lzb_query_free(strandQ[0]);
                        lzb_query_free(Q);
                        abortQuery = 1; break;
                    }
                }
                strandSegs[pass] = segs; strandN[pass] = nsegs; segs = NULL; nsegs = 0;
            }
            if (!steps[step].finish) continue;
            if (o.hspLimit) { segs = strandSegs[pass]; nsegs = strandN[pass]; strandSegs[pass] = NULL; }
            if (o.adaptive) {
                if (!tableSplit) {
                    tableSplit = 1;
                    if (doPlus && doMinus) {                     /* - strand stays in the table, + strand moves to the second one */
                        lzb_hsptable_split(&table, strandId[1], &others);
                        strandSegs[1] = table.seg; strandN[1] = table.len; strandSegs[0] = others.seg; strandN[0] = others.len;
                    } else { strandSegs[pass] = table.seg; strandN[pass] = table.len; free(others.seg); }
                    /* the gapped threshold follows the lowest HSP score kept, over both tables; an EMPTY table counts as
                     * the worst possible score (finish_one_strand lastz.c:3278-3285, segment.c:127,209,1372) */
                    lowAnchorScore = table.lowScore < others.lowScore ? table.lowScore : others.lowScore;
                }
                segs = strandSegs[pass]; nsegs = strandN[pass]; strandSegs[pass] = NULL;
            }
            /* anchors from a file, and HSPs of a processor that may report overlapping ones, are merged per diagonal before
             * anything else looks at them (lastz.c:757, :2811, :3296; merge_segments segment.c:1527) */
            if (o.segmentsFile || o.twins || (o.recoverSeeds && !(o.gfExtend == LZB_GFEX_NONE && !o.gapped))) lzb_merge_segments(segs, &nsegs);
            if (!o.adaptive && o.selfCompare && !o.gapped && !o.segmentsFile && nsegs) {
                lzb_segment* both = malloc(2 * nsegs * sizeof *both); uint64_t m = 0;
                int same = query.revCompFlags == target.revCompFlags;
                for (uint64_t k = 0; k < nsegs; k++) {
                    lzb_segment g = segs[k]; both[m++] = g;
                    uint32_t len = g.length, e1 = g.pos1 + len, e2 = g.pos2 + len, s1, s2;
                    if (same) { s1 = e1; s2 = e2; }
                    else { s1 = target.len - e1 + len; s2 = query.len - e2 + len; if (s2 == e1 && s1 == e2) continue; }
                    g.pos1 = s2 - len; g.pos2 = s1 - len; g.hspId = 0;
                    both[m++] = g;
                }
                lzb_free(segs); segs = both; nsegs = m;
            }
            /* only the x-drop extension leaves real scores in the table (seed_search.c:2953); the other modes
             * are scored here when chaining or the gapped stage needs them (lastz.c:3336-3340, score_segments segment.c:1262) */
            if (!o.gapped && lzb_filters_active(&o.filters)) {   /* filter_segments_by_* lastz.c:3312-3332: identity, coverage, match counts */
                uint64_t kept = 0;
                for (uint64_t k = 0; k < nsegs; k++) {
                    lzb_editscript es = { 1, 1, LZB_OP_SUB, { LZB_OP_SUB | (segs[k].length << 2) } };
                    lzb_alignel al; memset(&al, 0, sizeof al);
                    al.beg1 = segs[k].pos1 + 1; al.end1 = segs[k].pos1 + segs[k].length; al.beg2 = segs[k].pos2 + 1; al.end2 = segs[k].pos2 + segs[k].length; al.script = &es;
                    if (!lzb_filters_reject(&o.filters, &target, &query, &al, 1)) segs[kept++] = segs[k];
                }
                nsegs = kept;
            }
            /* adaptive, both strands: the + strand's HSPs were MOVED to the second table, which never saw an extension and
             * so does not count as scored (haveScores: seed_search.c:2953 marks the table being searched into, segment.c:207
             * clears it on the emptied one) -- they are scored again here, which undoes their entropy adjustment */
            const int movedTable = o.adaptive && doPlus && doMinus && pass == 0;
            if (!o.segmentsFile && (o.gfExtend != LZB_GFEX_XDROP || movedTable) && (o.chain || o.gapped))
                for (uint64_t k = 0; k < nsegs; k++) {
                    int32_t sc = 0;
                    for (uint32_t j = 0; j < segs[k].length; j++) sc += ss.masked[(uint32_t)target.v[segs[k].pos1 + j] * 256 + query.v[segs[k].pos2 + j]];
                    segs[k].s = sc;
                }
            if (o.chain)                                         /* try_reduce_to_chain lastz.c:3349, chainScale = 100 (:511) */
                lzb_reduce_to_chains(&target, &query, segs, &nsegs, o.chainDiag, o.chainAnti, 100, ss.sub['A' * 256 + 'A']);
            if (o.anyOrNone && nsegs) {
                /* gappily_extend_hsps (gapped_extend.c:5279; a17): every HSP, in discovery order, is reduced to its
                 * peak and extended on its own, unconstrained by other alignments; the first one that reaches the
                 * gapped threshold is reported and the search stops (searchLimit 1).  Without gapped extension
                 * the first HSP is reported (report_filtered_hsps lastz.c:3905).  The anchors handed over are
                 * single-element tables, so the library's anchor loop has nothing to bound or to skip. */
                lzb_segment first; int have = 0; lzb_alignel* one = NULL;
                for (uint64_t k = 0; k < nsegs && !have; k++) {
                    first = segs[k];
                    if (!o.gapped) { have = 1; break; }
                    lzb_gapped_params gp; memset(&gp, 0, sizeof gp);
                    gp.yDrop = o.Y; gp.trimToPeak = o.trimToPeak; gp.scoreThreshold = o.L; gp.tracebackBytes = o.tracebackBytes; gp.speculation = 1;
                    lzb_segment anchor = first;
                    if (lzb_reduce_to_points(ctx, T, Q, &anchor, 1)) lzb_die("%s", lzb_last_error());
                    if (lzb_gapped_extend(ctx, T, Q, target.v, query.v, &anchor, 1, &gp, &one, &gst)) lzb_die("%s", lzb_last_error());
                    totCells += gst.dpCells; gapSec += gst.seconds;
                    if (one) have = 1;
                }
                lzb_free(segs);
                if (!have) { segs = NULL; nsegs = 0; }
                else if (!o.gapped) { segs = malloc(sizeof *segs); segs[0] = first; nsegs = 1; reported = 1; }
                else {                                            /* print through the normal path below: one alignment, no anchors */
                    reported = 1; segs = NULL; nsegs = 0;
                    int hd = 0;
                    for (lzb_alignel* a = one; a; a = a->next) {
                        if (o.blastHeader && !hd) { lzb_blastn_header(out, "lastz.v1.04.58", o.args, n1, &query); hd = 1; }
                        if (o.format == 0) { if (!hd) { lzb_lav_strand_header(out, &target, &query); hd = 1; } lzb_lav_align(out, &target, &query, a); }
                        else if (o.format == 4) lzb_maf_align(out, &target, &query, a);
                        else if (o.format == 5) lzb_axt_align(out, &target, &query, a, &axtNumber);
                        else if (o.format == 6) { if (!hd) { lzb_gfa_strand_header(out, &target, &query); hd = 1; } lzb_gfa_align(out, &target, &query, a, &ss); }
                        else if (o.format == 7) lzb_cigar_align(out, &target, &query, a);
                    else if (o.format == 9) lzb_rdotplot_align(out, &dotMain, &target, &query, a, &ss, o.dotScore);
                    else if (o.format == 8) lzb_sam_align(out, &target, &query, a, o.samEqx, o.samSoft);
                        else if (o.format >= 2) lzb_fieldlist_align(out, o.fields, &target, &query, a, &rowNumber);
                    }
                    lzb_free_align_list(one);
                    lzb_query_free(Q);
                    continue;
                }
            }
            int headerDone = 0;
            if (!o.gapped) {
                for (uint64_t k = 0; k < nsegs; k++) {
                    if (o.hspLimit && o.format != 9) { if (printedForQuery >= o.hspLimit) break; printedForQuery++; }
                    if (dotOut) lzb_rdotplot_match(dotOut, &dotSide, &target, &query, &segs[k], &ss, o.dotplotFileScore);
                    if (o.blastHeader && !headerDone) { lzb_blastn_header(out, "lastz.v1.04.58", o.args, n1, &query); headerDone = 1; }   /* per query and strand, before its first row */
                    if (o.format == 0) {
                        if (!headerDone) { lzb_lav_strand_header(out, &target, &query); headerDone = 1; }
                        lzb_lav_match(out, &target, &query, &segs[k]);
                    } else if (o.format == 4 || o.format == 5) { /* an HSP is an alignment with one run of substitutions */
                        lzb_editscript es = { 1, 1, LZB_OP_SUB, { LZB_OP_SUB | (segs[k].length << 2) } };
                        lzb_alignel al; memset(&al, 0, sizeof al);
                        al.beg1 = segs[k].pos1 + 1; al.end1 = segs[k].pos1 + segs[k].length; al.beg2 = segs[k].pos2 + 1; al.end2 = segs[k].pos2 + segs[k].length;
                        al.s = segs[k].s; al.script = &es;
                        if (o.format == 4) lzb_maf_align(out, &target, &query, &al); else lzb_axt_align(out, &target, &query, &al, &axtNumber);
                    } else if (o.format == 6) {
                        if (!headerDone) { lzb_gfa_strand_header(out, &target, &query); headerDone = 1; }
                        lzb_gfa_match(out, &target, &query, &segs[k]);
                    }
                    else if (o.format == 7) lzb_cigar_match(out, &target, &query, &segs[k]);
                    else if (o.format == 9) lzb_rdotplot_match(out, &dotMain, &target, &query, &segs[k], &ss, o.dotScore);
                    else if (o.format == 8) lzb_sam_match(out, &target, &query, &segs[k], o.samEqx, o.samSoft);
                    else if (o.format >= 2) lzb_fieldlist_match(out, o.fields, &target, &query, &segs[k], &rowNumber);
                }
            } else {
                lzb_gapped_params gp; memset(&gp, 0, sizeof gp);
                gp.yDrop = o.Y; gp.trimToPeak = o.trimToPeak; gp.scoreThreshold = o.L; gp.allBounds = o.allBounds;
                if (o.adaptive && !o.haveL) gp.scoreThreshold = lowAnchorScore;   /* lastz.c:3405-3410 */
                gp.inhibitTrivial = o.inhibitTrivial; gp.tracebackBytes = o.tracebackBytes;
                gp.identityCheck = query.revCompFlags == target.revCompFlags && !query.npart;   /* identical_sequences :1147, identical_partition_of_sequence :1131 */
                gp.speculation = o.speculation;
                if (o.queryDepth > 0) { gp.maxPairedBases = (uint64_t)ceil(o.queryDepth * query.len); gp.overlyPairedKeep = o.depthKeep; }   /* lastz.c:3414-3417 */
                if (lzb_reduce_to_points(ctx, T, Q, segs, nsegs)) lzb_die("%s", lzb_last_error());
                lzb_alignel* list = NULL;
                if (lzb_gapped_extend(ctx, T, Q, target.v, query.v, segs, nsegs, &gp, &list, &gst))
                    lzb_die("%s", lzb_last_error());
                if (gst.overlyPaired && o.depthWarn) {                /* warn_for_paired_bases_limit gapped_extend.c:5725 */
                    static int firstReport = 1;
                    char num[40], com[56]; snprintf(num, sizeof num, "%llu", (unsigned long long)gp.maxPairedBases);
                    size_t nl = strlen(num), w = 0;                    /* commatize utilities.c:1233 */
                    for (size_t z = 0; z < nl; z++) { com[w++] = num[z]; if ((nl - 1 - z) % 3 == 0 && z + 1 < nl) com[w++] = ','; }
                    com[w] = 0;
                    fprintf(stderr, "WARNING. Query %s (%c strand) contains more than %s paired bases.\n", query.npart ? "seq2" : query.shortHeader,
                            (query.revCompFlags & LZB_RCF_REV) ? '-' : '+', com);
                    if (firstReport) fprintf(stderr, o.depthKeep ? "Any gapped alignments already found for this query/strand are reported but the\nquery/strand is not processed further.\n"
                                                                 : "All gapped alignments for this query/strand are discarded and the query/strand\nis not processed further.\n");
                    firstReport = 0;
                }
                totCells += gst.dpCells; gapSec += gst.seconds; gk += gst.kernelSeconds[0];
                gExt += gst.anchorsExtended; gSpec += gst.speculated; gRedo += gst.redone; gLaunch += gst.launches; gTrunc += gst.truncated;
                if (lzb_filters_active(&o.filters)) {             /* filter_aligns_by_* lastz.c:3430-3462 */
                    lzb_alignel* head = NULL; lzb_alignel** tail = &head;
                    for (lzb_alignel* a = list; a;) {
                        lzb_alignel* next = a->next; a->next = NULL;
                        if (lzb_filters_reject(&o.filters, &target, &query, a, 0)) lzb_free_align_list(a);
                        else { *tail = a; tail = &a->next; }
                        a = next;
                    }
                    list = head;
                }
                if (o.selfCompare && list)                        /* mirrorGapped, lastz.c:3494-3498 */
                    list = lzb_mirror_alignments(list, &target, &query, &ss);
                if (o.hspLimit && list && o.format != 9) { if (printedForQuery >= o.hspLimit) { lzb_free_align_list(list); list = NULL; } else printedForQuery++; }
                for (lzb_alignel* a = list; a; a = a->next) {
                    if (dotOut) lzb_rdotplot_align(dotOut, &dotSide, &target, &query, a, &ss, o.dotplotFileScore);
                    if (o.blastHeader && !headerDone) { lzb_blastn_header(out, "lastz.v1.04.58", o.args, n1, &query); headerDone = 1; }
                    if (o.format == 0) {
                        if (!headerDone) { lzb_lav_strand_header(out, &target, &query); headerDone = 1; }
                        lzb_lav_align(out, &target, &query, a);
                    } else if (o.format == 4) lzb_maf_align(out, &target, &query, a);
                    else if (o.format == 5) lzb_axt_align(out, &target, &query, a, &axtNumber);
                    else if (o.format == 6) {
                        if (!headerDone) { lzb_gfa_strand_header(out, &target, &query); headerDone = 1; }
                        lzb_gfa_align(out, &target, &query, a, &ss);
                    }
                    else if (o.format == 7) lzb_cigar_align(out, &target, &query, a);
                    else if (o.format == 9) lzb_rdotplot_align(out, &dotMain, &target, &query, a, &ss, o.dotScore);
                    else if (o.format == 8) lzb_sam_align(out, &target, &query, a, o.samEqx, o.samSoft);
                    else if (o.format >= 2) lzb_fieldlist_align(out, o.fields, &target, &query, a, &rowNumber);
                }
                lzb_free_align_list(list);
            }
            lzb_free(segs);
            lzb_query_free(Q);
        }
        if (overLimit) queriesOverLimit++;
        (void)abortQuery;
        lzb_seq_free(&query);
    }
    if (queriesOverLimit && o.hspLimitWarn)                      /* lastz.c:1777-1793 */
        fprintf(stderr, queriesOverLimit == 1 ? "1 query exceeded the HSP limit\n" : "%d queries exceeded the HSP limit\n", queriesOverLimit);
    if (o.format == 0) lzb_lav_footer(out);
    if (o.showStats)
        fprintf(stderr, "backend=%s raw_seed_hits=%llu hsps=%llu dp_cells=%llu seed_seconds=%.6f gapped_seconds=%.6f\n",
                lzb_backend(), (unsigned long long)totHits, (unsigned long long)totHsps,
                (unsigned long long)totCells, seedSec, gapSec);
    if (o.showStats) {
        fprintf(stderr, "seed kernels (s): words=%.4f count=%.4f slots=%.4f scan=%.4f expand=%.4f sort=%.4f bounds=%.4f extend=%.4f right=%.4f replay=%.4f left=%.4f\n",
                ks[0], ks[1], ks[2], ks[3], ks[4], ks[5], ks[6], ks[7], ks[8], ks[9], ks[10]);
        fprintf(stderr, "gapped: extended=%llu speculated=%llu redone=%llu truncated=%llu launches=%llu dp_kernel_seconds=%.4f\n",
                (unsigned long long)gExt, (unsigned long long)gSpec, (unsigned long long)gRedo,
                (unsigned long long)gTrunc, (unsigned long long)gLaunch, gk);
    }
    lzb_seqfile_close(qf); lzb_seqfile_close(tf);
    lzb_target_free(T); lzb_close(ctx); lzb_seq_free(&target);
    if (out != stdout) fclose(out);
    if (dotOut) fclose(dotOut);
    return 0;
}
