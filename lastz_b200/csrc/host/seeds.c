/*
 * seeds.c -- spaced-seed parsing and packing recipe.  Reference: seeds.c:321-632
 * (parse_one_seed), :1399-1418 (best_shift).  The packing (which unpacked bit lands on which index
 * bit) must be the reference's own greedy recipe: transition variants are probed in the order of
 * transFlips[] (seed_search.c:528-533) -- rightmost seed position first (seeds.c:165,603-613) -- which decides the
 * order in which seed hits are discovered.
 */
#include <string.h>
#include "lzb_host.h"

static int popcount64(uint64_t x) { return __builtin_popcountll(x); }

/* the shift that brings the most not-yet-covered index bits into place; first best wins */
static int greedy_shift(uint32_t uncovered, uint64_t remaining) {
    int best = -1, bestShift = -1;
    for (int sh = 0; remaining != 0; remaining >>= 1, sh++) {
        int cover = popcount64(remaining & uncovered);
        if (cover > best) { best = cover; bestShift = sh; }
    }
    return bestShift;
}

void lzb_seed_parse(lzb_seed* out, const char* pattern, int withTrans) {
    const char* s = pattern; const char* e = pattern + strlen(pattern);
    while (s < e && (*s == '0' || *s == 'X' || *s == 'x')) s++;
    if (s == e) lzb_die("seed string is empty!");
    while (e[-1] == '0' || e[-1] == 'X' || e[-1] == 'x') e--;
    uint64_t bits = 0, flips = 0; int length = 0, weight = 0, allStrict = 1, anyMatch = 0;
    for (const char* p = s; p < e; p++) {
        switch (*p) {
            case '1': bits = (bits << 2) + 3; flips = (flips << 2) + 2; weight += 2; length++; anyMatch = 1; break;
            case '0': case 'X': case 'x': bits <<= 2; flips <<= 2; length++; break;
            /* a transition position matches on the purine/pyrimidine bit alone (seeds.c:471-481).  The
             * reference packs all-T ("half-weight") seeds from a 1-bit-per-base word (seeds.c:404, pos_table.c:479);
             * here they go through the same 2-bit word with 1-bit masks: another index layout, the same hits
             * in the same order (half-weight seeds have no transition variants, seeds.c:540) */
            case 'T': case 't': bits = (bits << 2) + 1; flips <<= 2; weight += 1; length++; allStrict = 0; break;
            default: lzb_die("seed string %s contains illegal character %c", pattern, *p);
        }
    }
    if (!allStrict && !anyMatch) flips = 0;             /* type 'H' */
    if (weight == 0) lzb_die("seed string (%s) cannot have zero weight.", pattern);
    if (length > 31) lzb_die("seed string (%s) cannot have length exceeding 31 (it's %d).", pattern, length);
    if (weight > 28) lzb_die("seed (%s) needs %d index bits; lastz_b200 does not implement overweight seeds (max 28)", pattern, weight);
    memset(out, 0, sizeof *out);
    out->length = length; out->weight = weight; out->withTrans = withTrans;
    uint32_t wbits = (uint32_t)((1ull << weight) - 1);
    uint32_t covered = (uint32_t)bits & wbits;
    uint64_t rem = bits - covered;
    out->shift[0] = 0; out->mask[0] = covered; out->numParts = 1;
    while (covered != wbits) {
        int sh = greedy_shift(~covered & wbits, rem);
        uint32_t m = (uint32_t)(rem >> sh) & ~covered & wbits;
        covered += m; rem -= (uint64_t)m << sh;
        if (out->numParts >= LZB_MAX_SEED_PARTS) lzb_die("seed (%s) needs too many shift/mask parts", pattern);
        out->shift[out->numParts] = sh; out->mask[out->numParts] = m; out->numParts++;
    }
    /* single-bit transition flips in UNPACKED bit order, rightmost seed position first, each one carried through the
     * packing (seeds.c:603-613: the reference builds with maintainFlippedBitOrder defined, seeds.c:165) */
    while (flips) {
        uint64_t low = flips & (~flips + 1);
        flips -= low;
        uint32_t packedBit = 0;
        for (int i = 0; i < out->numParts; i++) packedBit |= (uint32_t)(low >> out->shift[i]) & out->mask[i];
        out->transFlips[out->numFlips++] = packedBit;
    }
}
