// test_ydrop_mw.cpp -- the Y-drop kernels (k_ydrop_mw<8,4> of ydrop_mw.cuh, the kernel that is 99 % of the bench
// step; its fallbacks k_ydrop_warp<16> and the shared-memory k_ydrop<256>) compiled for the host block emulator and checked against the
// ORACLE library through the C-ABI: for one anchor without neighbours, the two one-sided DPs of the kernel
// must give the oracle's alignment -- score, end points and edit script, column by column.
// TEST INFRASTRUCTURE: this is the one place outside tests/*.py that links liblzb_oracle.so.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/lastz_b200.h"
#include "../../lastz_b200/csrc/cuda/lzb_types.h"
#include "cuda_emu.h"
#define LZB_DYNAMIC_SHARED(name_) static unsigned char name_[200 * 1024] __attribute__((aligned(16)))   /* k_ydrop's ring: one block runs at a time */
#include "../../lastz_b200/csrc/cuda/ydrop_common.cuh"
#include "../../lastz_b200/csrc/cuda/ydrop_smem.cuh"
#include "../../lastz_b200/csrc/cuda/ydrop_warp.cuh"
#include "../../lastz_b200/csrc/cuda/ydrop_mw.cuh"

static u64 rng_state = 0x2545F4914F6CDD1Dull;
static u64 rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

static const s32 HOX[4][4] = { { 91, -114, -31, -123 }, { -114, 100, -125, -31 }, { -31, -125, 100, -114 }, { -123, -31, -114, 91 } };

// the score set of lzb_scores_default (HOXD70; NUL row/column veryBadScore; others -100), as 256x256 + its class reduction
static std::vector<int32_t> g_sub(65536), g_msub(65536);
static lzb_scoring_dev g_sc;
static void build_scoring() {
    const char* acgt = "ACGT";
    for (int a = 0; a < 256; a++) for (int b = 0; b < 256; b++) {
        s32 v = -100;
        if (a == 0 || b == 0) v = -107374182;
        else {
            const char* pa = strchr(acgt, a & ~32), *pb = strchr(acgt, b & ~32);
            if (pa && pb && *pa && *pb) v = HOX[pa - acgt][pb - acgt];
        }
        g_sub[a * 256 + b] = v; g_msub[a * 256 + b] = v;
    }
    memset(&g_sc, 0, sizeof g_sc);
    int rep[LZB_MAX_CLASSES], nc = 0;
    for (int b = 0; b < 256; b++) {
        int found = -1;
        for (int k = 0; k < nc && found < 0; k++) { bool same = true; for (int x = 0; x < 256 && same; x++) same = g_sub[b * 256 + x] == g_sub[rep[k] * 256 + x] && g_sub[x * 256 + b] == g_sub[x * 256 + rep[k]]; if (same) found = k; }
        if (found < 0) { rep[nc] = b; found = nc++; }
        g_sc.cls[b] = (u8)found;
    }
    g_sc.numClasses = nc;
    for (int i = 0; i < nc; i++) for (int j = 0; j < nc; j++) g_sc.subC[i * LZB_MAX_CLASSES + j] = g_sc.msubC[i * LZB_MAX_CLASSES + j] = g_sub[rep[i] * 256 + rep[j]];
    g_sc.gapOpen = 400; g_sc.gapExtend = 30;
}

// run-length ops -> one character per alignment column
static void expand(std::string& out, const u32* ops, u32 n, bool reversedOrder) {
    if (!reversedOrder) for (u32 k = 0; k < n; k++) out.append(ops[k] >> 2, "?IDS"[ops[k] & 3]);
    else for (u32 k = n; k-- > 0;) out.append(ops[k] >> 2, "?IDS"[ops[k] & 3]);
}

static int one_case(int caseNo, u32 len, double sub, double indel, u32 tbBytes, s32 yDrop, int trim, int oneWarp = 0, int partitioned = 0) {
    // query = target through a substitution/indel channel, so that one long alignment exists
    std::string t, q; const char* acgt = "ACGT";
    for (u32 i = 0; i < len; i++) t.push_back(acgt[rnd() & 3]);
    for (u32 i = 0; i < len; i++) {
        double r = (rnd() % 100000) / 100000.0;
        if (r < sub) q.push_back(acgt[rnd() & 3]);
        else if (r < sub + indel / 2) continue;
        else if (r < sub + indel) { q.push_back(t[i]); q.push_back(acgt[rnd() & 3]); }
        else q.push_back(t[i]);
    }
    if (partitioned) { q[q.size() / 3] = 0; q[2 * q.size() / 3] = 0; }
    if (partitioned == 2) { t[t.size() / 3 + 40] = 0; t[2 * t.size() / 3 - 55] = 0; }   // ... and a [multi] target, cut elsewhere   // a [multi] query: the sweeps must stop at the NULs around the anchor (gapped_extend.c:1357-1372)
    const u32 len1 = (u32)t.size(), len2 = (u32)q.size();
    // anchor somewhere in the middle, on the homologous diagonal as far as we can tell: use the oracle's own seed stage
    lzb_ctx* oc = lzb_open(0);
    lzb_set_scoring(oc, g_sub.data(), g_msub.data(), 400, 30);
    int8_t ctb[256]; memset(ctb, -1, 256); ctb['A'] = 0; ctb['C'] = 1; ctb['G'] = 2; ctb['T'] = 3;
    lzb_seed seed; memset(&seed, 0, sizeof seed);                     // 12 consecutive matches, no transitions
    seed.length = 12; seed.weight = 24; seed.numParts = 1; seed.shift[0] = 0; seed.mask[0] = 0xFFFFFF;
    lzb_target* T = lzb_target_build(oc, (const uint8_t*)t.data(), len1, 0, 0, ctb, &seed, 1);
    lzb_query* Q = lzb_query_load(oc, (const uint8_t*)q.data(), len2);
    lzb_seed_params sp; memset(&sp, 0, sizeof sp); sp.gfExtend = LZB_GFEX_XDROP; sp.xDrop = 910; sp.hspThreshold = 3000; sp.entropy = 1; sp.hashBits = 16;
    lzb_segment* segs = NULL; uint64_t nsegs = 0;
    if (lzb_seed_hit_search(oc, T, Q, &seed, ctb, &sp, &segs, &nsegs, NULL) || nsegs == 0) { printf("case %d: no HSP, skipped\n", caseNo); return 0; }
    lzb_segment best = segs[0]; for (uint64_t k = 1; k < nsegs; k++) if (segs[k].s > best.s) best = segs[k];
    lzb_reduce_to_points(oc, T, Q, &best, 1);
    const u32 a1 = best.pos1, a2 = best.pos2;
    lzb_gapped_params gp; memset(&gp, 0, sizeof gp);
    gp.yDrop = yDrop; gp.trimToPeak = trim; gp.scoreThreshold = -2000000000; gp.tracebackBytes = tbBytes; gp.speculation = 1;
    lzb_alignel* want = NULL; lzb_segment anchor = best;
    if (lzb_gapped_extend(oc, T, Q, (const uint8_t*)t.data(), (const uint8_t*)q.data(), &anchor, 1, &gp, &want, NULL)) { fprintf(stderr, "oracle failed: %s\n", lzb_last_error()); return 1; }
    // the kernel, both sides, on the emulator
    std::vector<u8> c1(len1 + 64, g_sc.cls[0]), c2(len2 + 64, g_sc.cls[0]);
    for (u32 i = 0; i < len1; i++) c1[i] = g_sc.cls[(u8)t[i]];
    for (u32 i = 0; i < len2; i++) c2[i] = g_sc.cls[(u8)q[i]];
    const u32 tbLen = 1 + (tbBytes - 8);
    u32 lo2 = a2, hi2 = a2;                                            // the anchor's partition, as gapped.cu's launch() cuts it
    while (lo2 > 0 && q[lo2 - 1] != 0) lo2--;
    while (hi2 < len2 && q[hi2] != 0) hi2++;
    u32 lo1 = a1, hi1 = a1;
    while (lo1 > 0 && t[lo1 - 1] != 0) lo1--;
    while (hi1 < len1 && t[hi1] != 0) hi1++;
    dp_job jobs[2]; memset(jobs, 0, sizeof jobs);
    std::vector<u8> tb[2]; std::vector<u32> tbRow[2], ops[2]; std::vector<int> act[2];
    for (int side = 0; side < 2; side++) {
        dp_job& J = jobs[side]; const int rev = side == 0;
        J.reversed = rev; J.a1 = a1; J.a2 = a2;
        J.M = rev ? a1 + 1 - lo1 : hi1 - (a1 + 1); J.N = rev ? a2 + 1 - lo2 : hi2 - (a2 + 1);
        J.L0 = 0; J.R0 = (s32)(J.N + 1); J.leftSeg = { -1, -1 }; J.rightSeg = { -1, -1 }; J.listv = NULL; J.alignList = 0; J.al = NULL; J.resume = -1; J.token = 7;
        tb[side].resize((size_t)tbBytes + 64); tbRow[side].resize(len1 + len2 + 16); ops[side].resize(2 * (len1 + len2) + 16); act[side].resize(5 * 16);
        J.tb = tb[side].data(); J.tbLen = tbLen; J.tbRow = tbRow[side].data(); J.tbRowCap = (u32)tbRow[side].size();
        J.ops = ops[side].data(); J.opsCap = (u32)ops[side].size(); J.act = act[side].data(); J.actCap = 16;
    }
    launch_list ll; memset(&ll, 0, sizeof ll); ll.ix[0] = 0; ll.ix[1] = 1;
    if (oneWarp == 2) emu_launch(2, 256, [&]() { k_ydrop<256>(jobs, ll, (const dseg*)NULL, c1.data(), c2.data(), len1, len2, &g_sc, yDrop, trim, 4096u); });   // the shared-memory fallback, 4096-column ring
    else if (oneWarp) emu_launch(2, 32, [&]() { k_ydrop_warp<16>(jobs, ll, (const dseg*)NULL, c1.data(), c2.data(), len1, len2, &g_sc, yDrop, trim); });   // the one-warp kernel (512-column window)
    else emu_launch(2, 128, [&]() { k_ydrop_mw<8, 4>(jobs, ll, (const dseg*)NULL, c1.data(), c2.data(), len1, len2, &g_sc, yDrop, trim); });
    int bad = 0;
    /* the same two sweeps stopped by a row limit and continued from their last checkpoint must end exactly alike */
    if (oneWarp != 2) {
        dp_job pj[2]; memcpy(pj, jobs, sizeof pj);
        std::vector<u8> tb2[2]; std::vector<u32> tbRow2[2], ops2[2], ck[2]; std::vector<int> act2[2];
        const u32 every = 64;
        for (int side = 0; side < 2; side++) {
            dp_job& J = pj[side];
            tb2[side].resize((size_t)tbBytes + 64); tbRow2[side].resize(tbRow[side].size()); ops2[side].resize(ops[side].size()); act2[side].resize(5 * 16);
            ck[side].resize((size_t)256 * CK_RECORD_WORDS);
            J.tb = tb2[side].data(); J.tbRow = tbRow2[side].data(); J.ops = ops2[side].data(); J.act = act2[side].data();
            J.ckpt = ck[side].data(); J.ckptCap = 256; J.ckptEvery = every; J.resume = -1; J.done = 0; J.token = 8;
            J.rowLimit = jobs[side].rows > 3 * every ? jobs[side].rows * 2 / 3 : 0;
        }
        auto run = [&]() {
            if (oneWarp) emu_launch(2, 32, [&]() { k_ydrop_warp<16>(pj, ll, (const dseg*)NULL, c1.data(), c2.data(), len1, len2, &g_sc, yDrop, trim); });
            else emu_launch(2, 128, [&]() { k_ydrop_mw<8, 4>(pj, ll, (const dseg*)NULL, c1.data(), c2.data(), len1, len2, &g_sc, yDrop, trim); });
        };
        run();
        int paused = 0;
        for (int side = 0; side < 2; side++) if (pj[side].status == DP_PAUSED) {
            paused++;
            if (pj[side].ckptCount == 0) { fprintf(stderr, "  side %d paused without a checkpoint\n", side); bad++; }
            pj[side].resume = (int)pj[side].ckptCount - 1; pj[side].rowLimit = 0; pj[side].done = 0; pj[side].token = 9;
        } else pj[side].token = 0;
        if (paused) {
            launch_list l2; memset(&l2, 0, sizeof l2); int n2 = 0;
            for (int side = 0; side < 2; side++) if (pj[side].token == 9) l2.ix[n2++] = (u16)side;
            if (oneWarp) emu_launch(n2, 32, [&]() { k_ydrop_warp<16>(pj, l2, (const dseg*)NULL, c1.data(), c2.data(), len1, len2, &g_sc, yDrop, trim); });
            else emu_launch(n2, 128, [&]() { k_ydrop_mw<8, 4>(pj, l2, (const dseg*)NULL, c1.data(), c2.data(), len1, len2, &g_sc, yDrop, trim); });
        }
        for (int side = 0; side < 2; side++) {
            const dp_job& a = jobs[side]; const dp_job& b = pj[side];
            if (a.status != b.status || a.score != b.score || a.end1 != b.end1 || a.end2 != b.end2 || a.rows != b.rows || a.cells != b.cells || a.nops != b.nops ||
                memcmp(ops[side].data(), ops2[side].data(), (size_t)a.nops * 4)) {
                fprintf(stderr, "  side %d: continued from a checkpoint: status %d/%d score %d/%d end (%u,%u)/(%u,%u) rows %u/%u cells %llu/%llu nops %u/%u\n", side, a.status, b.status, a.score, b.score,
                        a.end1, a.end2, b.end1, b.end2, a.rows, b.rows, a.cells, b.cells, a.nops, b.nops);
                bad++;
            }
        }
        printf("        pause/continue: %d of 2 sweeps stopped at two thirds and continued from their last checkpoint\n", paused);
    }
    for (int side = 0; side < 2; side++) if (jobs[side].done != 7 || jobs[side].opsOverflow) { fprintf(stderr, "  side %d: done=%u overflow=%d\n", side, jobs[side].done, jobs[side].opsOverflow); bad++; }
    for (int side = 0; side < 2; side++) if (jobs[side].status != DP_OK && jobs[side].status != DP_TRUNCATED) { fprintf(stderr, "  side %d: kernel status %d\n", side, jobs[side].status); bad++; }
    // assemble like ydrop_align (gapped_extend.c:2529-2560): left script in emission order, right script reversed
    std::string gotCols; expand(gotCols, ops[0].data(), jobs[0].nops, false); expand(gotCols, ops[1].data(), jobs[1].nops, true);
    const u32 start1 = a1 + 1 - jobs[0].end1, start2 = a2 + 1 - jobs[0].end2, stop1 = a1 + jobs[1].end1, stop2 = a2 + jobs[1].end2;
    const s32 score = jobs[0].score + jobs[1].score;
    bool lopped = gotCols.empty() || gotCols.front() != 'S' || gotCols.back() != 'S';     // lop_initial/final_indels: skip those
    if (!want) { if (!gotCols.empty() && !lopped) { fprintf(stderr, "  oracle found nothing, kernel did\n"); bad++; } }
    else if (!lopped) {
        std::string wantCols; expand(wantCols, want->script->op, want->script->len, false);
        if (want->s != score || want->beg1 != start1 + 1 || want->beg2 != start2 + 1 || want->end1 != stop1 + 1 || want->end2 != stop2 + 1 || wantCols != gotCols) {
            fprintf(stderr, "  oracle: s=%d (%u,%u)-(%u,%u) %zu columns; kernel: s=%d (%u,%u)-(%u,%u) %zu columns\n", want->s, want->beg1, want->beg2, want->end1, want->end2,
                    wantCols.size(), score, start1 + 1, start2 + 1, stop1 + 1, stop2 + 1, gotCols.size());
            bad++;
        }
    }
    printf("case %2d%s: %u x %u bp, sub=%.2f indel=%.3f traceback=%u yDrop=%d trim=%d: score %d, %zu columns, rows %u+%u, cells %llu, status %d/%d%s  %s\n", caseNo, oneWarp == 2 ? " (shared-memory kernel)" : oneWarp ? " (one-warp kernel)" : "", len1, len2, sub, indel,
           tbBytes, yDrop, trim, score, gotCols.size(), jobs[0].rows, jobs[1].rows, jobs[0].cells + jobs[1].cells, jobs[0].status, jobs[1].status, lopped ? " (lopped, not compared)" : "", bad ? "MISMATCH" : "ok");
    lzb_free_align_list(want); lzb_free(segs); lzb_query_free(Q); lzb_target_free(T); lzb_close(oc);
    return bad;
}

int main() {
    build_scoring();
    int bad = 0, n = 0;
    bad += one_case(n++, 3000, 0.04, 0.010, 80u << 20, 9400, 1);         // the bench channel
    bad += one_case(n++, 6000, 0.08, 0.020, 80u << 20, 9400, 1);
    bad += one_case(n++, 4000, 0.04, 0.010, 300000, 9400, 1);            // traceback runs out: truncated, same stopping row
    bad += one_case(n++, 3000, 0.15, 0.030, 80u << 20, 9400, 1);         // ends early inside the sequences
    bad += one_case(n++, 3000, 0.04, 0.010, 80u << 20, 3000, 1);         // narrow band
    bad += one_case(n++, 3000, 0.05, 0.010, 80u << 20, 9400, 0);         // --noytrim: boundary scores
    bad += one_case(n++, 1500, 0.30, 0.050, 80u << 20, 9400, 1);         // mostly noise
    bad += one_case(n++, 2500, 0.04, 0.010, 80u << 20, 6000, 1, 1);      // k_ydrop_warp<16>
    bad += one_case(n++, 2500, 0.06, 0.015, 200000, 6000, 0, 1);         // ... truncated, --noytrim
    bad += one_case(n++, 2000, 0.04, 0.010, 80u << 20, 9400, 1, 2);      // k_ydrop<256> (shared-memory ring)
    bad += one_case(n++, 3000, 0.04, 0.010, 80u << 20, 400, 1);          // a Y-drop of a few mismatches (read-mapping settings): bands of a handful of columns
    bad += one_case(n++, 3000, 0.02, 0.004, 80u << 20, 14, 0);           // ... smaller than one mismatch, --noytrim
    bad += one_case(n++, 2000, 0.04, 0.010, 80u << 20, 60, 1, 1);        // ... on the one-warp kernel
    bad += one_case(n++, 3000, 0.04, 0.010, 80u << 20, 9400, 1, 0, 1);   // partitioned query: both sweeps end at a NUL
    bad += one_case(n++, 3000, 0.04, 0.010, 80u << 20, 9400, 0, 0, 1);   // ... --noytrim: the partition edge is a sequence edge
    bad += one_case(n++, 3000, 0.04, 0.010, 80u << 20, 9400, 1, 0, 2);   // target and query both partitioned
    bad += one_case(n++, 3000, 0.04, 0.010, 80u << 20, 9400, 0, 0, 2);
    bad += one_case(n++, 2000, 0.07, 0.020, 150000, 9400, 0, 2);         // ... truncated, --noytrim
    printf("%d cases, %d mismatching, %llu collectives emulated\n", n, bad, emu_collectives);
    return bad ? 1 : 0;
}
