// test_xdrop.cpp -- the warp-cooperative bucket replay (lastz_b200/csrc/cuda/xdrop_warp.cuh) on the
// host lane emulator against a plain sequential restatement of process_for_simple_hit +
// xdrop_extend_seed_hit (seed_search.c:1056-1192, :2528-2959).  TEST INFRASTRUCTURE.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
typedef uint8_t u8; typedef uint32_t u32; typedef uint64_t u64; typedef int32_t s32; typedef int64_t s64;
#define LZB_WARP_EMU 1
#include "wemu.h"
#include "../../lastz_b200/csrc/cuda/xdrop_warp.cuh"

static u64 rng_state = 88172645463325252ull;
static u64 rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

// classes: 0..3 = ACGT, 4 = masked/N (-1000), 5 = NUL (veryBadScore)
static s32 g_lut[256];
static const s32 HOX[4][4] = { { 91, -114, -31, -123 }, { -114, 100, -125, -31 }, { -31, -125, 100, -114 }, { -123, -31, -114, 91 } };
static s32 pair_score(u32 a, u32 b) {
    if (a == 5 || b == 5) return -107374182;
    if (a == 4 || b == 4) return -1000;
    return HOX[a][b];
}

struct ref_out { std::vector<cand_rec> cand; u32 E; u64 nExt, nBp; };

static void count_matches(const u8* a1, const u8* a2, cand_rec& r) {
    r.cA = r.cC = r.cG = r.cT = 0;
    for (u32 i = 0; i < r.length; i++) {
        u8 x = a1[r.pos1 + i];
        if (x == a2[r.pos2 + i]) { r.cA += x == 'A'; r.cC += x == 'C'; r.cG += x == 'G'; r.cT += x == 'T'; }
    }
}

static void reference(const std::vector<u8>& c1, const std::vector<u8>& c2, const std::vector<u8>& a1, const std::vector<u8>& a2,
                      u32 len1, u32 len2, u32 L, s32 xDrop, s32 K, int entropy, const std::vector<u64>& hits, u32 E0, ref_out& out) {
    u32 E = E0; out.nExt = out.nBp = 0;
    for (u64 rec : hits) {
        u32 pos1 = (u32)rec, pos2 = (u32)(rec >> 32);
        s64 diag = (s64)pos1 - (s64)pos2;
        if (E > pos2 - L) continue;
        s64 blk = (s64)E + diag; u32 stop = blk > 0 ? (u32)blk : 0;
        u32 a = pos1, b = pos2, leftStart = pos1; s32 run = 0, leftScore = 0;
        while (a > stop && run >= leftScore - xDrop) { --a; --b; run += pair_score(c1[a], c2[b]); if (run > leftScore) { leftStart = a; leftScore = run; } }
        u32 leftScanned = a;
        s64 lim = (s64)len2 + diag; u32 rstop = ((s64)len1 <= lim) ? len1 : (u32)lim;
        a = pos1; b = pos2; u32 rightStop = pos1; s32 rightScore = 0; run = 0;
        while (a < rstop && run >= rightScore - xDrop) { run += pair_score(c1[a], c2[b]); a++; b++; if (run > rightScore) { rightStop = a; rightScore = run; } }
        out.nExt++; out.nBp += a - leftScanned;
        u32 extent = (u32)((s64)a - diag);
        if (extent > E) E = extent;
        s32 sim = leftScore + rightScore;
        if (sim < K) continue;
        cand_rec r; r.hit1 = pos1; r.hit2 = pos2; r.pos1 = leftStart; r.pos2 = (u32)((s64)leftStart - diag);
        r.length = rightStop - leftStart; r.score = sim; r.cA = r.cC = r.cG = r.cT = 0;
        if (entropy && sim <= 3 * K) count_matches(a1.data(), a2.data(), r);
        out.cand.push_back(r);
    }
    out.E = E;
}

static bool cand_less(const cand_rec& a, const cand_rec& b) { return a.hit2 != b.hit2 ? a.hit2 < b.hit2 : a.hit1 < b.hit1; }

static int one_case(int caseNo, u32 len1, u32 len2, u32 nhits, double homolog, double mut, double junk, s32 xDrop, s32 K, int entropy, u32 E0) {
    const u32 L = 19;
    // sequences: seq2 = mutated copy of seq1 shifted by `shift`, so that one diagonal is homologous
    std::vector<u8> c1(len1 + 64, 5), c2(len2 + 64, 5), a1(len1 + 64, 0), a2(len2 + 64, 0);
    const char* ACGT = "ACGT";
    for (u32 i = 0; i < len1; i++) { c1[i] = rnd() & 3; if ((rnd() % 1000) < junk * 1000) c1[i] = 4; }
    s64 shift = (s64)(rnd() % 64) - 32;
    // homology comes in blocks (100..2100 columns) separated by unrelated stretches (30..330) when homolog < 1,
    // so that a bucket sees many separate HSPs; homolog == 1 keeps one uninterrupted diagonal
    u32 blockLeft = 0; bool inBlock = true;
    for (u32 j = 0; j < len2; j++) {
        s64 i = (s64)j + shift;
        if (homolog < 1.0 && blockLeft-- == 0) { inBlock = (rnd() % 1000) < homolog * 1000 ? !inBlock : inBlock; inBlock = !inBlock; blockLeft = inBlock ? 100 + rnd() % 2000 : 30 + rnd() % 300; }
        u8 v = (i >= 0 && i < (s64)len1 && inBlock) ? c1[i] : (u8)(rnd() & 3);
        if (v < 4 && (rnd() % 1000) < mut * 1000) v = (v + 1 + rnd() % 3) & 3;
        if ((rnd() % 1000) < junk * 1000) v = 4;
        c2[j] = v;
    }
    for (u32 i = 0; i < len1; i++) a1[i] = c1[i] < 4 ? ACGT[c1[i]] : (c1[i] == 4 ? 'n' : 0);
    for (u32 j = 0; j < len2; j++) a2[j] = c2[j] < 4 ? ACGT[c2[j]] : (c2[j] == 4 ? 'N' : 0);
    // hits of one bucket in discovery order: pos2 ascending; most on the homologous diagonal
    std::vector<u64> hits;
    for (u32 k = 0; k < nhits; k++) {
        u32 pos2 = L + (u32)(rnd() % (len2 - L + 1)), pos1;
        if (rnd() % 100 < 70) { s64 p = (s64)pos2 + shift; if (p < (s64)L || p > (s64)len1) continue; pos1 = (u32)p; }
        else pos1 = L + (u32)(rnd() % (len1 - L + 1));
        hits.push_back(((u64)pos2 << 32) | pos1);
    }
    std::stable_sort(hits.begin(), hits.end(), [](u64 a, u64 b) { return (a >> 32) < (b >> 32); });
    ref_out want; reference(c1, c2, a1, a2, len1, len2, L, xDrop, K, entropy, hits, E0, want);
    // the warp
    std::vector<cand_rec> got(hits.size() + 1); unsigned long long ncand = 0;
    xd_env e; e.cls1 = c1.data(); e.cls2 = c2.data(); e.asc1 = a1.data(); e.asc2 = a2.data(); e.lut = g_lut;
    e.len1 = len1; e.len2 = len2; e.L = L; e.xDrop = xDrop; e.K = K; e.entropy = entropy;
    e.cand = got.data(); e.candCap = (u32)got.size(); e.ncand = &ncand;
    u32 Eout[32]; unsigned long long nExt[32], nBp[32];
    wemu_run([&](int lane) {
        u32 E = E0; unsigned long long x = 0, b = 0;
        xd_bucket(e, (u32)lane, hits.data(), 0, (u32)hits.size(), E, x, b);
        Eout[lane] = E; nExt[lane] = x; nBp[lane] = b;
    });
    u64 tx = 0, tb = 0; for (int l = 0; l < 32; l++) { tx += nExt[l]; tb += nBp[l]; }
    int bad = 0;
    for (int l = 0; l < 32; l++) if (Eout[l] != want.E) { bad++; break; }
    if (tx != want.nExt || tb != want.nBp) bad++;
    got.resize(ncand);
    std::sort(got.begin(), got.end(), cand_less); std::sort(want.cand.begin(), want.cand.end(), cand_less);
    if (got.size() != want.cand.size()) bad++;
    else for (size_t i = 0; i < got.size(); i++) if (memcmp(&got[i], &want.cand[i], sizeof(cand_rec))) { bad++;
        fprintf(stderr, "  cand %zu: got hit=(%u,%u) pos=(%u,%u) len=%u s=%d cnt=%u,%u,%u,%u  want hit=(%u,%u) pos=(%u,%u) len=%u s=%d cnt=%u,%u,%u,%u\n", i,
                got[i].hit1, got[i].hit2, got[i].pos1, got[i].pos2, got[i].length, got[i].score, got[i].cA, got[i].cC, got[i].cG, got[i].cT,
                want.cand[i].hit1, want.cand[i].hit2, want.cand[i].pos1, want.cand[i].pos2, want.cand[i].length, want.cand[i].score,
                want.cand[i].cA, want.cand[i].cC, want.cand[i].cG, want.cand[i].cT); break; }
    printf("case %2d: len=%u/%u hits=%zu homolog=%.2f mut=%.3f junk=%.3f X=%d K=%d ent=%d E0=%u -> ext=%llu bp=%llu cand=%zu E=%u  %s\n", caseNo, len1, len2,
           hits.size(), homolog, mut, junk, xDrop, K, entropy, E0, (unsigned long long)want.nExt, (unsigned long long)want.nBp, want.cand.size(), want.E, bad ? "MISMATCH" : "ok");
    if (bad) fprintf(stderr, "  got ext=%llu bp=%llu cand=%zu E=%u\n", (unsigned long long)tx, (unsigned long long)tb, got.size(), Eout[0]);
    return bad;
}

int main(int argc, char** argv) {
    for (u32 a = 0; a < 6; a++) for (u32 b = 0; b < 6; b++) g_lut[a * XD_LUT_STRIDE + b] = pair_score(a, b);
    int reps = argc > 1 ? atoi(argv[1]) : 1, bad = 0, n = 0;
    for (int r = 0; r < reps; r++) {
        bad += one_case(n++, 3000, 3000, 400, 1.0, 0.05, 0.0, 910, 3000, 1, 0);       // the default channel: HSPs of a few hundred columns
        bad += one_case(n++, 5000, 4800, 600, 1.0, 0.01, 0.0, 910, 3000, 1, 0);       // long HSPs: several 256-column trips
        bad += one_case(n++, 6000, 6000, 500, 1.0, 0.0, 0.0, 910, 3000, 1, 0);        // identical: scans run to the sequence ends
        bad += one_case(n++, 2000, 2500, 800, 0.0, 0.0, 0.0, 910, 3000, 1, 0);        // random only
        bad += one_case(n++, 3000, 3000, 500, 0.9, 0.03, 0.01, 910, 3000, 1, 0);      // masked bases and partial homology
        bad += one_case(n++, 3000, 3000, 500, 1.0, 0.04, 0.0, 400, 2000, 0, 0);       // other x-drop, no entropy
        bad += one_case(n++, 3000, 3000, 500, 1.0, 0.04, 0.0, 910, 0, 1, 700);        // K = 0 (every extension is a candidate), bucket already advanced
        bad += one_case(n++, 300, 280, 200, 1.0, 0.02, 0.0, 910, 1000, 1, 0);         // tiny sequences: the near-start byte path
        bad += one_case(n++, 4000, 4000, 33, 1.0, 0.002, 0.0, 5000, 3000, 1, 0);      // huge x-drop, one batch and a bit
        bad += one_case(n++, 200000, 200000, 20000, 0.5, 0.05, 0.0, 910, 3000, 1, 0);  // many separate HSPs in one bucket (the heavy-bucket case)
        bad += one_case(n++, 200000, 190000, 20000, 0.5, 0.02, 0.002, 910, 3000, 1, 0);
        bad += one_case(n++, 100000, 100000, 30000, 0.3, 0.10, 0.0, 910, 1500, 1, 0);  // short, frequent HSPs
        bad += one_case(n++, 100000, 100000, 30000, 0.5, 0.08, 0.01, 300, 1000, 0, 0);
        bad += one_case(n++, 400000, 400000, 3000, 0.0, 0.0, 0.0, 910, 3000, 1, 0);    // sparse random hits: nearly all are extended
    }
    printf("%d cases, %d mismatching, %llu collectives emulated\n", n, bad, wemu_collectives);
    return bad ? 1 : 0;
}
