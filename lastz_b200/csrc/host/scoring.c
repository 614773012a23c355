/*
 * scoring.c -- scoring sets for lastz_b200: HOXD70 defaults, the masked copy used by x-drop
 * extension, and the scoring-file grammar.  Reference: dna_utilities.c:137-148 (HOXD70),
 * :215-313 (new_dna_score_set), :497-558 (masked_score_set), :581-628 (file grammar).
 */
#include <ctype.h>
#include <stdlib.h>
#include <string.h>
#include "lzb_host.h"

#define VERY_BAD_SCORE (-107374182)   /* dna_utilities.h:139, (score)(worstPossibleScore/20) */

#define __ (-1)
const int8_t lzb_upper_nuc_to_bits[256] = {
    __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__, __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__,
    __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__, __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__,
    __, 0,__, 1,__,__,__, 2,__,__,__,__,__,__,__,__, __,__,__,__, 3,__,__,__,__,__,__,__,__,__,__,__,
    __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__, __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__,
    __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__, __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__,
    __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__, __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__,
    __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__, __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__,
    __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__, __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__ };
const int8_t lzb_nuc_to_bits[256] = {
    __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__, __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__,
    __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__, __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__,
    __, 0,__, 1,__,__,__, 2,__,__,__,__,__,__,__,__, __,__,__,__, 3,__,__,__,__,__,__,__,__,__,__,__,
    __, 0,__, 1,__,__,__, 2,__,__,__,__,__,__,__,__, __,__,__,__, 3,__,__,__,__,__,__,__,__,__,__,__,
    __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__, __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__,
    __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__, __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__,
    __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__, __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__,
    __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__, __,__,__,__,__,__,__,__,__,__,__,__,__,__,__,__ };
#undef __

static int32_t hoxd70[4][4] = {
    {  91, -114,  -31, -123 },
    {-114,  100, -125,  -31 },
    { -31, -125,  100, -114 },
    {-123,  -31, -114,   91 } };

#define S(ss, r, c) (ss)->sub[(r) * 256 + (c)]

void lzb_scores_from_template(lzb_scoreset* ss, int32_t t[4][4], int32_t bad, int32_t fill,
                              int32_t gapOpen, int32_t gapExtend) {
    const char* nuc = "ACGT";
    memset(ss, 0, sizeof *ss);
    ss->gapOpen = gapOpen; ss->gapExtend = gapExtend;
    for (int c = 0; c < 256; c++) S(ss, 0, c) = VERY_BAD_SCORE;
    for (int r = 1; r < 256; r++) { S(ss, r, 0) = VERY_BAD_SCORE; for (int c = 1; c < 256; c++) S(ss, r, c) = fill; }
    for (int c = 0; c < 256; c++) S(ss, 'X', c) = S(ss, 'x', c) = S(ss, c, 'X') = S(ss, c, 'x') = bad;
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) {
        int R = nuc[r], C = nuc[c], rl = tolower(R), cl = tolower(C);
        S(ss, R, C) = S(ss, R, cl) = S(ss, rl, C) = S(ss, rl, cl) = t[r][c];
    }
    lzb_scores_mask(ss);
}

void lzb_scores_default(lzb_scoreset* ss) { lzb_scores_from_template(ss, hoxd70, -1000, -100, 400, 30); }

/* masked_score_set for a DNA score set whose row/column characters are ACGTacgt */
void lzb_scores_mask(lzb_scoreset* ss) {
    memcpy(ss->masked, ss->sub, sizeof ss->sub);
    int32_t bad = S(ss, 'A', 'X');
    const char* low = "acgt";
    for (int k = 0; k < 4; k++) for (int c = 1; c < 256; c++) ss->masked[low[k] * 256 + c] = bad;
    for (int c = 1; c < 256; c++) ss->masked['N' * 256 + c] = ss->masked['n' * 256 + c] = ss->masked['X' * 256 + c] = bad;
    for (int k = 0; k < 4; k++) for (int r = 1; r < 256; r++) ss->masked[r * 256 + low[k]] = bad;
    for (int r = 1; r < 256; r++) ss->masked[r * 256 + 'N'] = ss->masked[r * 256 + 'n'] = ss->masked[r * 256 + 'X'] = bad;
}

/* read_score_set dna_utilities.c:657ff, DNA (ACGT) matrices only */
void lzb_scores_read_file(lzb_scoreset* ss, const char* path) {
    FILE* f = fopen(path, "rt");
    if (!f) lzb_die("fopen_or_die failed to open \"%s\" for \"rt\"", path);
    int32_t bad = -1000, fill = -100, go = 400, ge = 30, t[4][4];
    int goSet = 0, geSet = 0, rows = 0, haveCols = 0, colIx[4] = {0, 1, 2, 3};
    lzb_scoreset ex; memset(&ex, 0, sizeof ex);
    char line[1024];
    while (fgets(line, sizeof line, f)) {
        char* h = strchr(line, '#'); if (h) *h = 0;
        char* p = line; while (isspace((unsigned char)*p)) p++;
        if (!*p) continue;
        char* eq = strchr(p, '=');
        if (eq && !haveCols) {
            *eq = 0; char* val = eq + 1; while (isspace((unsigned char)*val)) val++;
            char name[64]; sscanf(p, "%63s", name);
            char* colon = strrchr(val, ':'); if (colon) val = colon + 1;
            long v = strtol(val, NULL, 10);
            if (!strcmp(name, "bad_score")) bad = (int32_t)v;
            else if (!strcmp(name, "fill_score")) fill = (int32_t)v;
            else if (!strcmp(name, "gap_open_penalty")) { go = (int32_t)v; goSet = 1; }
            else if (!strcmp(name, "gap_extend_penalty")) { ge = (int32_t)v; geSet = 1; }
            else if (!strcmp(name, "hsp_threshold") || !strcmp(name, "hsp_thresh")) { ex.hspThreshold = (int32_t)v; ex.hspThresholdSet = 1; }
            else if (!strcmp(name, "gapped_threshold") || !strcmp(name, "gapped_thresh")) { ex.gappedThreshold = (int32_t)v; ex.gappedThresholdSet = 1; }
            else if (!strcmp(name, "x_drop")) { ex.xDrop = (int32_t)v; ex.xDropSet = 1; }
            else if (!strcmp(name, "y_drop")) { ex.yDrop = (int32_t)v; ex.yDropSet = 1; }
            else if (!strcmp(name, "step")) { ex.step = (uint32_t)v; ex.stepSet = 1; }
            else lzb_die("unknown setting \"%s\" in score file %s", name, path);
            continue;
        }
        if (!haveCols) {
            char lab[4][8];
            if (sscanf(p, "%7s %7s %7s %7s", lab[0], lab[1], lab[2], lab[3]) != 4) lzb_die("bad column labels in score file %s", path);
            for (int k = 0; k < 4; k++) {
                const char* w = strchr("ACGT", toupper((unsigned char)lab[k][0]));
                if (!w || lab[k][1]) lzb_die("lastz_b200 supports only A,C,G,T score files (%s)", path);
                colIx[k] = (int)(w - "ACGT");
            }
            haveCols = 1; continue;
        }
        if (rows >= 4) lzb_die("too many rows in score file %s", path);
        int r = colIx[rows];
        if (isalpha((unsigned char)*p)) {
            const char* w = strchr("ACGT", toupper((unsigned char)*p));
            if (!w) lzb_die("lastz_b200 supports only A,C,G,T score files (%s)", path);
            r = (int)(w - "ACGT"); p++;
        }
        long v[4];
        if (sscanf(p, "%ld %ld %ld %ld", &v[0], &v[1], &v[2], &v[3]) != 4) lzb_die("bad matrix row in score file %s", path);
        for (int k = 0; k < 4; k++) t[r][colIx[k]] = (int32_t)v[k];
        rows++;
    }
    fclose(f);
    if (rows != 4) lzb_die("score file %s has %d matrix rows (need 4)", path, rows);
    lzb_scores_from_template(ss, t, bad, fill, go, ge);
    ss->gapOpenSet = goSet; ss->gapExtendSet = geSet;
    ss->hspThresholdSet = ex.hspThresholdSet; ss->hspThreshold = ex.hspThreshold;
    ss->gappedThresholdSet = ex.gappedThresholdSet; ss->gappedThreshold = ex.gappedThreshold;
    ss->xDropSet = ex.xDropSet; ss->xDrop = ex.xDrop; ss->yDropSet = ex.yDropSet; ss->yDrop = ex.yDrop;
    ss->stepSet = ex.stepSet; ss->step = ex.step;
}
