// test_gapped_sched.cpp -- the gapped stage's anchor loop (lastz_b200/csrc/cuda/gapped_sched.hpp: speculative sweeps,
// validation against later commits, resume from checkpoints) run on the HOST: the scheduler is the product's source,
// the sweeps are the product's kernels (k_ydrop_mw<8,4> with its fallbacks) on the block emulator, and job completion
// is delivered in a shuffled order after a random number of polls so that harvest order differs from launch order.
// The alignment list must equal the ORACLE's gapped_extend for the same anchors -- end points, scores, edit scripts op
// for op, and the counters the reference keeps (anchors extended, DP cells, truncations).
// TEST INFRASTRUCTURE: links liblzb_oracle.so (the checker); nothing here is part of the product library.
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/lastz_b200.h"
#include "../../lastz_b200/csrc/cuda/lzb_types.h"
#include "cuda_emu.h"
#define LZB_DYNAMIC_SHARED(name_) static unsigned char name_[200 * 1024] __attribute__((aligned(16)))
#include "../../lastz_b200/csrc/cuda/ydrop_common.cuh"
#include "../../lastz_b200/csrc/cuda/ydrop_smem.cuh"
#include "../../lastz_b200/csrc/cuda/ydrop_warp.cuh"
#include "../../lastz_b200/csrc/cuda/ydrop_mw.cuh"
#include "../../lastz_b200/csrc/cuda/gapped_sched.hpp"

static u64 rng_state = 0x2545F4914F6CDD1Dull;
static u64 rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static const s32 HOX[4][4] = { { 91, -114, -31, -123 }, { -114, 100, -125, -31 }, { -31, -125, 100, -114 }, { -123, -31, -114, 91 } };
static std::vector<int32_t> g_sub(65536), g_msub(65536);
static lzb_scoring_dev g_sc;
static void build_scoring() {
    const char* acgt = "ACGT";
    for (int a = 0; a < 256; a++) for (int b = 0; b < 256; b++) {
        s32 v = -100;
        if (a == 0 || b == 0) v = -107374182;
        else { const char* pa = strchr(acgt, a & ~32), *pb = strchr(acgt, b & ~32); if (pa && pb && *pa && *pb) v = HOX[pa - acgt][pb - acgt]; }
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

static char g_err[512];
static int emu_fail(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); fprintf(stderr, "  scheduler: %s\n", g_err); return -1; }

struct host_lane { std::vector<u8> tb[2]; std::vector<u32> tbRow[2], ops[2], ckpt[2]; std::vector<int> act[2], list[2]; };
struct emu_backend {
    u32 tbBytes, tbLen, every; s32 yDrop; int trim, shuffle;
    const u8* c1; const u8* c2; u32 len1, len2;
    std::vector<host_lane> L; std::vector<dp_job> jobs;
    std::vector<dseg> dsegs; std::vector<dalign> daligns;
    std::vector<u32> pendingToken; std::vector<int> pendingPolls;      // a finished job becomes visible after that many polls
    u64 launches = 0, jobsRun = 0;
    const char* error() { return "emulator"; }
    u32 ckpt_every() { return every; }
    u32 ring(int mode) { return mode == 2 ? 4096u : 8192u; }
    int lanes(int want) {                                              // called again when the scheduler wants more
        const int old = (int)L.size();
        if (want <= old) return old;
        L.resize(want); jobs.resize(2 * (size_t)want); pendingToken.resize(2 * (size_t)want, 0); pendingPolls.resize(2 * (size_t)want, 0);
        for (int z = old; z < want; z++) for (int s = 0; s < 2; s++) {
            host_lane& ln = L[z];
            ln.tb[s].resize((size_t)tbBytes + 64); ln.tbRow[s].resize(4096); ln.ops[s].resize(64); ln.act[s].resize(5 * 2);
            ln.ckpt[s].resize((size_t)64 * CK_RECORD_WORDS);
            dp_job& J = jobs[2 * z + s]; memset(&J, 0, sizeof J);
            fill(z, s);
        }
        return want;
    }
    void fill(int z, int s) {
        host_lane& ln = L[z]; dp_job& J = jobs[2 * z + s];
        J.tb = ln.tb[s].data(); J.tbLen = tbLen; J.tbRow = ln.tbRow[s].data(); J.tbRowCap = (u32)ln.tbRow[s].size();
        J.ops = ln.ops[s].data(); J.opsCap = (u32)ln.ops[s].size(); J.act = ln.act[s].data(); J.actCap = (u32)(ln.act[s].size() / 5);
        J.ckpt = ln.ckpt[s].data(); J.ckptCap = 64; J.ckptEvery = every; J.resume = -1;
    }
    dp_job* job(int z, int s) { return &jobs[2 * z + s]; }
    const u32* ops(int z, int s) { return L[z].ops[s].data(); }
    int* list(int z, int s, size_t n) { if (L[z].list[s].size() < n) L[z].list[s].resize(n * 2); return L[z].list[s].data(); }
    int tables(const dseg* s, size_t s0, size_t s1, const dalign* a, size_t a0, size_t a1) {
        dsegs.insert(dsegs.end(), s + s0, s + s1); daligns.insert(daligns.end(), a + a0, a + a1);
        return 0;
    }
    const dseg* segs() { return dsegs.data(); }
    const dalign* aligns() { return daligns.data(); }
    int launch(int mode, const u16* ix, int n) {
        launch_list ll; memset(&ll, 0, sizeof ll);
        for (int k = 0; k < n; k++) { ll.ix[k] = ix[k]; jobs[ix[k]].al = daligns.data(); }      // the tables are vectors: they may have moved
        launches++; jobsRun += (u64)n;
        dp_job* jv = jobs.data(); const dseg* sg = dsegs.data();
        // completion is withheld: remember each job's token, clear it for the kernel's duration
        std::vector<u32> tok(n);
        for (int k = 0; k < n; k++) tok[k] = jobs[ix[k]].token;
        if (mode == 0) emu_launch(n, 128, [&]() { k_ydrop_mw<8, 4>(jv, ll, sg, c1, c2, len1, len2, &g_sc, yDrop, trim); });
        else if (mode == 1) emu_launch(n, 32, [&]() { k_ydrop_warp<16>(jv, ll, sg, c1, c2, len1, len2, &g_sc, yDrop, trim); });
        else emu_launch(n, 256, [&]() { k_ydrop<256>(jv, ll, sg, c1, c2, len1, len2, &g_sc, yDrop, trim, ring(mode)); });
        for (int k = 0; k < n; k++) {
            dp_job& J = jobs[ix[k]];
            if (J.done != tok[k]) { fprintf(stderr, "  job %d did not signal completion\n", ix[k]); return -1; }
            if (shuffle) { J.done = 0; pendingToken[ix[k]] = tok[k]; pendingPolls[ix[k]] = 1 + (int)(rnd() % 7); }
        }
        return 0;
    }
    bool poll() {
        for (size_t k = 0; k < jobs.size(); k++) if (pendingToken[k] && --pendingPolls[k] <= 0) { jobs[k].done = pendingToken[k]; pendingToken[k] = 0; }
        return true;
    }
    int grow(int z, int s, int what) {
        host_lane& ln = L[z];
        if (what == DP_TBROW) ln.tbRow[s].resize(ln.tbRow[s].size() * 4);
        else if (what == DP_ACT) ln.act[s].resize(ln.act[s].size() * 4);
        else if (what == DP_OPS) ln.ops[s].resize(ln.ops[s].size() * 4);
        fill(z, s);
        return 0;
    }
};

static void expand(std::string& out, const u32* ops, u32 n) { for (u32 k = 0; k < n; k++) out.append(ops[k] >> 2, "?IDS"[ops[k] & 3]); }

// homologous pair with `blocks` rearranged pieces: enough structure for neighbours on both sides, above and below
static int one_case(int caseNo, u32 len, double sub, double indel, u32 tbBytes, s32 yDrop, int trim, int W, u32 every, int shuffle, int dupes, const char* slack = "-1", const char* mode = "1", u64 maxPaired = 0, int keepPaired = 0) {
    setenv("LZB_DP_MODE", mode, 1);         // 1: the one-warp kernel (the default), 0: the four-warp kernel, 2: the shared-memory kernel (no checkpoints)
    setenv("LZB_GAP_SLACK", slack, 1);     // -1: every anchor is started as soon as a lane is free (most speculation); else the margin in rows
    std::string t, q; const char* acgt = "ACGT";
    for (u32 i = 0; i < len; i++) t.push_back(acgt[rnd() & 3]);
    auto channel = [&](const std::string& src, std::string& dst) {
        for (size_t i = 0; i < src.size(); i++) {
            double r = (rnd() % 100000) / 100000.0;
            if (r < sub) dst.push_back(acgt[rnd() & 3]);
            else if (r < sub + indel / 2) continue;
            else if (r < sub + indel) { dst.push_back(src[i]); dst.push_back(acgt[rnd() & 3]); }
            else dst.push_back(src[i]);
        }
    };
    channel(t, q);
    // repeats: copies of target pieces appended to the query, so that alignments on other diagonals cross anchor rows
    for (int d = 0; d < dupes; d++) { const u32 a = (u32)(rnd() % (len / 2)), l = len / 6 + (u32)(rnd() % (len / 6)); std::string piece; channel(t.substr(a, l), piece); q += piece; }
    const u32 len1 = (u32)t.size(), len2 = (u32)q.size();
    lzb_ctx* oc = lzb_open(0);
    lzb_set_scoring(oc, g_sub.data(), g_msub.data(), 400, 30);
    int8_t ctb[256]; memset(ctb, -1, 256); ctb['A'] = 0; ctb['C'] = 1; ctb['G'] = 2; ctb['T'] = 3;
    lzb_seed seed; memset(&seed, 0, sizeof seed);
    seed.length = 12; seed.weight = 24; seed.numParts = 1; seed.shift[0] = 0; seed.mask[0] = 0xFFFFFF;
    lzb_target* T = lzb_target_build(oc, (const uint8_t*)t.data(), len1, 0, 0, ctb, &seed, 1);
    lzb_query* Q = lzb_query_load(oc, (const uint8_t*)q.data(), len2);
    lzb_seed_params sp; memset(&sp, 0, sizeof sp); sp.gfExtend = LZB_GFEX_XDROP; sp.xDrop = 910; sp.hspThreshold = 3000; sp.entropy = 1; sp.hashBits = 16;
    lzb_segment* segs = NULL; uint64_t nsegs = 0;
    if (lzb_seed_hit_search(oc, T, Q, &seed, ctb, &sp, &segs, &nsegs, NULL) || nsegs == 0) { printf("case %d: no HSP, skipped\n", caseNo); return 0; }
    lzb_reduce_to_points(oc, T, Q, segs, nsegs);
    std::vector<lzb_segment> a1(segs, segs + nsegs), a2(segs, segs + nsegs);
    lzb_gapped_params gp; memset(&gp, 0, sizeof gp);
    gp.yDrop = yDrop; gp.trimToPeak = trim; gp.scoreThreshold = 3000; gp.tracebackBytes = tbBytes; gp.speculation = W;
    gp.maxPairedBases = maxPaired; gp.overlyPairedKeep = keepPaired;       // --querydepth: nothing is extended once the kept alignments pair more bases than this
    lzb_alignel* want = NULL; lzb_gapped_stats so; memset(&so, 0, sizeof so);
    if (lzb_gapped_extend(oc, T, Q, (const uint8_t*)t.data(), (const uint8_t*)q.data(), a1.data(), nsegs, &gp, &want, &so)) { fprintf(stderr, "oracle failed: %s\n", lzb_last_error()); return 1; }
    // the scheduler over the emulated kernels
    std::vector<u8> c1(len1 + 64, g_sc.cls[0]), c2(len2 + 64, g_sc.cls[0]);
    for (u32 i = 0; i < len1; i++) c1[i] = g_sc.cls[(u8)t[i]];
    for (u32 i = 0; i < len2; i++) c2[i] = g_sc.cls[(u8)q[i]];
    emu_backend B; B.tbBytes = tbBytes; B.tbLen = 1 + (tbBytes - 8); B.every = every; B.yDrop = yDrop; B.trim = trim; B.shuffle = shuffle;
    B.c1 = c1.data(); B.c2 = c2.data(); B.len1 = len1; B.len2 = len2;
    std::string tz = t, qz = q; tz.push_back(0); qz.push_back(0);
    gx_input in; in.h_seq1 = (const u8*)tz.data(); in.h_seq2 = (const u8*)qz.data(); in.len1 = len1; in.len2 = len2; in.hostSub = g_sub.data(); in.gapOpen = 400; in.gapExtend = 30;
    lzb_alignel* got = NULL; lzb_gapped_stats sg; memset(&sg, 0, sizeof sg);
    int bad = 0;
    if (gx_run(B, in, a2.data(), nsegs, &gp, &got, &sg, emu_fail)) bad++;
    u32 nw = 0, ng = 0;
    lzb_alignel* w = want; lzb_alignel* g = got;
    for (; w && g; w = w->next, g = g->next, nw++, ng++) {
        std::string cw, cg; expand(cw, w->script->op, w->script->len); expand(cg, g->script->op, g->script->len);
        if (w->s != g->s || w->beg1 != g->beg1 || w->beg2 != g->beg2 || w->end1 != g->end1 || w->end2 != g->end2 || cw != cg) {
            if (bad < 5) fprintf(stderr, "  alignment %u: oracle s=%d (%u,%u)-(%u,%u) %zu columns; scheduler s=%d (%u,%u)-(%u,%u) %zu columns\n", nw, w->s, w->beg1, w->beg2, w->end1, w->end2, cw.size(),
                                 g->s, g->beg1, g->beg2, g->end1, g->end2, cg.size());
            bad++;
        }
    }
    for (; w; w = w->next) nw++;
    for (; g; g = g->next) ng++;
    if (nw != ng) { fprintf(stderr, "  oracle %u alignments, scheduler %u\n", nw, ng); bad++; }
    if (so.overlyPaired != sg.overlyPaired || (maxPaired && !so.overlyPaired)) { fprintf(stderr, "  paired-bases limit: oracle %llu, scheduler %llu (a case with a limit must reach it)\n", (unsigned long long)so.overlyPaired, (unsigned long long)sg.overlyPaired); bad++; }
    if (so.anchorsExtended != sg.anchorsExtended || so.dpCells != sg.dpCells || so.truncated != sg.truncated) {
        fprintf(stderr, "  counters: oracle extended=%llu cells=%llu truncated=%llu; scheduler %llu %llu %llu\n", (unsigned long long)so.anchorsExtended, (unsigned long long)so.dpCells,
                (unsigned long long)so.truncated, (unsigned long long)sg.anchorsExtended, (unsigned long long)sg.dpCells, (unsigned long long)sg.truncated);
        bad++;
    }
    printf("case %2d: %u x %u bp, %llu anchors, traceback=%u yDrop=%d trim=%d lanes=%d ckpt=%u shuffle=%d kernel=%s slack=%s: %u alignments, extended %llu, speculated %llu, redone %llu, truncated %llu, launches %llu (%llu jobs)  %s\n",
           caseNo, len1, len2, (unsigned long long)nsegs, tbBytes, yDrop, trim, W, every, shuffle, mode, slack, ng, (unsigned long long)sg.anchorsExtended, (unsigned long long)sg.speculated,
           (unsigned long long)sg.redone, (unsigned long long)sg.truncated, (unsigned long long)B.launches, (unsigned long long)B.jobsRun, bad ? "MISMATCH" : "ok");
    fflush(stdout);
    lzb_free_align_list(want); lzb_free_align_list(got); lzb_free(segs); lzb_query_free(Q); lzb_target_free(T); lzb_close(oc);
    return bad;
}

int main(int argc, char** argv) {
    build_scoring();
    const int big = argc > 1 ? atoi(argv[1]) : 0;
    int bad = 0, n = 0;
    bad += one_case(n++, 6000, 0.04, 0.010, 60000, 9400, 1, 1, 64, 0, 0);       // one lane: the sequential order, tiled by truncation
    bad += one_case(n++, 6000, 0.04, 0.010, 60000, 9400, 1, 16, 64, 1, 0);      // 16 lanes: every sweep speculative, resumed from checkpoints
    bad += one_case(n++, 6000, 0.05, 0.012, 50000, 9400, 0, 8, 32, 1, 2, "-1", "0");   // repeats: neighbours across anchor rows, --noytrim; four-warp kernel
    bad += one_case(n++, 5000, 0.04, 0.010, 80000, 6000, 1, 3, 96, 1, 1);       // fewer lanes than anchors worth starting
    bad += one_case(n++, 4000, 0.04, 0.010, 60000, 9400, 1, 8, 64, 1, 1, "-1", "2");   // the shared-memory kernel: no checkpoints, a touched sweep restarts
    bad += one_case(n++, 6000, 0.04, 0.010, 60000, 9400, 1, 16, 64, 1, 1, "20"); // the product's waiting rule (anchors near an expected reach wait for the commit)
    bad += one_case(n++, 6000, 0.04, 0.010, 60000, 9400, 1, 16, 64, 1, 0, "-1", "1", 1500, 1);   // --querydepth=keep: the limit is passed after a few commits, the rest is dropped
    bad += one_case(n++, 6000, 0.05, 0.012, 50000, 9400, 1, 8, 32, 1, 2, "20", "0", 2500, 0);    // --querydepth: everything is discarded
    bad += one_case(n++, 14000, 0.04, 0.010, 200000, 9400, 1, 16, 32, 1, 1, "10", "0"); // sweeps long enough to be stopped short of an earlier anchor's alignment and continued
    if (big) {
        bad += one_case(n++, 20000, 0.04, 0.010, 100000, 9400, 1, 32, 64, 1, 3);
        bad += one_case(n++, 12000, 0.06, 0.015, 70000, 9400, 1, 24, 128, 1, 4);
        bad += one_case(n++, 12000, 0.04, 0.010, 70000, 9400, 0, 5, 32, 0, 2);
        bad += one_case(n++, 20000, 0.04, 0.010, 100000, 9400, 1, 32, 64, 1, 3, "30");
    }
    printf("%d cases, %d mismatching\n", n, bad);
    return bad ? 1 : 0;
}
