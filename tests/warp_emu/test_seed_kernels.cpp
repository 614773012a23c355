// test_seed_kernels.cpp -- every kernel of the seed stage (lastz_b200/csrc/cuda/{classify_kernel,index_kernels,
// seed_kernels,peaks_kernel}.cuh) compiled for the host block emulator and driven by a host loop that
// mirrors lzb_seed_hit_search (seed_search.cu), checked against the ORACLE library through the C-ABI:
// index contents, raw hit counts, HSP tables (x-drop through k_extend2 and the first k_extend, --exact and
// --mismatch through k_extend_alt, the three-kernel split, raw hits) and anchor peaks.  TEST INFRASTRUCTURE (links liblzb_oracle.so).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>
#include "../../include/lastz_b200.h"
#include "../../lastz_b200/csrc/cuda/lzb_types.h"
#include "cuda_emu.h"

// stand-in for the one cub primitive the kernels use (k_count_hits): result valid in thread 0, like cub's
namespace cub {
template <class T, int N> struct BlockReduce {
    struct TempStorage { T v[N]; };
    TempStorage& t;
    BlockReduce(TempStorage& t_) : t(t_) {}
    T Sum(T x) {
        t.v[threadIdx.x] = x; __syncthreads();
        T s = 0; if (threadIdx.x == 0) for (unsigned i = 0; i < blockDim.x; i++) s += t.v[i];
        __syncthreads(); return s;
    }
};
}
#include "../../lastz_b200/csrc/cuda/classify_kernel.cuh"
#include "../../lastz_b200/csrc/cuda/index_kernels.cuh"
#include "../../lastz_b200/csrc/cuda/seed_kernels.cuh"
#include "../../lastz_b200/csrc/cuda/xdrop_split.cuh"
#include "../../lastz_b200/csrc/cuda/peaks_kernel.cuh"

static u64 rng_state = 0x9E3779B97F4A7C15ull;
static u64 rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static const s32 HOX[4][4] = { { 91, -114, -31, -123 }, { -114, 100, -125, -31 }, { -31, -125, 100, -114 }, { -123, -31, -114, 91 } };

// HOXD70 score set + masked copy (lzb_scores_default / lzb_scores_mask) and their class reduction (lzb_set_scoring)
static std::vector<int32_t> g_sub(65536), g_msub(65536);
static lzb_scoring_dev g_sc;
static void build_scoring() {
    const char* acgt = "ACGT";
    for (int a = 0; a < 256; a++) for (int b = 0; b < 256; b++) {
        s32 v = -100;
        if (a == 0 || b == 0) v = -107374182;
        else if (a == 'X' || a == 'x' || b == 'X' || b == 'x') v = -1000;
        else { const char* pa = strchr(acgt, a & ~32), *pb = strchr(acgt, b & ~32); if (pa && pb && *pa && *pb && isalpha(a) && isalpha(b)) v = HOX[pa - acgt][pb - acgt]; }
        g_sub[a * 256 + b] = v;
    }
    g_msub = g_sub;                                                    // masked_score_set dna_utilities.c:530-555
    for (int a = 1; a < 256; a++) for (int b = 1; b < 256; b++)
        if ((a >= 'a' && a <= 'z') || (b >= 'a' && b <= 'z') || a == 'N' || b == 'N' || a == 'X' || b == 'X') g_msub[a * 256 + b] = -1000;
    memset(&g_sc, 0, sizeof g_sc);
    int rep[LZB_MAX_CLASSES], nc = 0;
    for (int bi = 0; bi < 260; bi++) {                                 // A C G T first, as context.cu does
        const int b = bi < 4 ? "ACGT"[bi] : bi - 4;
        if (bi >= 4 && (b == 'A' || b == 'C' || b == 'G' || b == 'T')) continue;
        int found = -1;
        for (int k = 0; k < nc && found < 0; k++) {
            bool same = true; const int r = rep[k];
            for (int x = 0; x < 256 && same; x++) same = g_sub[b * 256 + x] == g_sub[r * 256 + x] && g_sub[x * 256 + b] == g_sub[x * 256 + r] &&
                                                         g_msub[b * 256 + x] == g_msub[r * 256 + x] && g_msub[x * 256 + b] == g_msub[x * 256 + r];
            if (same) found = k;
        }
        if (found < 0) { rep[nc] = b; found = nc++; }
        g_sc.cls[b] = (u8)found;
    }
    g_sc.numClasses = nc;
    for (int i = 0; i < nc; i++) for (int j = 0; j < nc; j++) { g_sc.subC[i * LZB_MAX_CLASSES + j] = g_sub[rep[i] * 256 + rep[j]]; g_sc.msubC[i * LZB_MAX_CLASSES + j] = g_msub[rep[i] * 256 + rep[j]]; }
    g_sc.gapOpen = 400; g_sc.gapExtend = 30;
}

// seeds.c's greedy shift/mask recipe (csrc/host/seeds.c), enough for the two seeds used here
static void parse_seed(lzb_seed* out, const char* pattern, int withTrans) {
    u64 bits = 0, flips = 0; int length = 0, weight = 0;
    for (const char* p = pattern; *p; p++) { if (*p == '1') { bits = (bits << 2) + 3; flips = (flips << 2) + 2; weight += 2; } else { bits <<= 2; flips <<= 2; } length++; }
    memset(out, 0, sizeof *out); out->length = length; out->weight = weight; out->withTrans = withTrans;
    u32 wbits = (u32)((1ull << weight) - 1), covered = (u32)bits & wbits; u64 rem = bits - covered;
    out->shift[0] = 0; out->mask[0] = covered; out->numParts = 1;
    while (covered != wbits) {
        int best = -1, bestShift = -1; u64 r = rem;
        for (int sh = 0; r != 0; r >>= 1, sh++) { int cover = __builtin_popcountll(r & (~covered & wbits)); if (cover > best) { best = cover; bestShift = sh; } }
        u32 m = (u32)(rem >> bestShift) & ~covered & wbits;
        covered += m; rem -= (u64)m << bestShift;
        out->shift[out->numParts] = bestShift; out->mask[out->numParts] = m; out->numParts++;
    }
    while (flips) {                                                    // rightmost seed position first, seeds.c:603-613
        u64 low = flips & (~flips + 1); flips -= low; u32 bit = 0;
        for (int i = 0; i < out->numParts; i++) bit |= (u32)(low >> out->shift[i]) & out->mask[i];
        out->transFlips[out->numFlips++] = bit;
    }
}
static u32 host_word_at(const u8* v, u32 endPos, const lzb_seed* sd, const int8_t* ctb) {
    u64 w = 0; for (int j = 0; j < sd->length; j++) w = (w << 2) | (u64)(ctb[v[endPos - sd->length + j]] & 3);
    u32 p = 0; for (int i = 0; i < sd->numParts; i++) p |= (u32)(w >> sd->shift[i]) & sd->mask[i];
    return p;
}

struct padded { std::vector<u8> asc, cls; };
static void upload(const std::string& s, padded& o) {                  // lzb_upload_classes + k_classify
    size_t pad = ((s.size() + 1 + 15) / 16) * 16 + 16;
    o.asc.assign(pad, 0); memcpy(o.asc.data(), s.data(), s.size()); o.cls.assign(pad, 0xEE);
    emu_launch(2, 256, [&]() { k_classify((const uint4*)o.asc.data(), (uint4*)o.cls.data(), pad / 16, &g_sc); });
}

static int g_bad = 0;
#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, "  FAILED %s: ", #cond); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); g_bad++; } } while (0)

struct mode { const char* name; int gfExtend, mismatches, K, plain, entropy, hashBits, useFirstKernel, recover, twinMin, twinMax, chunkBlocks, searchLimit; };

static void one_pair(int caseNo, u32 len, const char* pattern, int withTrans, u32 step, const std::vector<mode>& modes, u32 partitionEvery = 0) {
    std::string t, q; const char* acgt = "ACGT";
    for (u32 i = 0; i < len; i++) t.push_back(acgt[rnd() & 3]);
    for (u32 i = 0; i < len; i++) { u64 r = rnd() % 1000; if (r < 50) q.push_back(acgt[rnd() & 3]); else if (r < 55) continue; else if (r < 60) { q.push_back(t[i]); q.push_back(acgt[rnd() & 3]); } else q.push_back(t[i]); }
    for (u32 i = len / 3; i < len / 3 + 400 && i < q.size(); i++) q[i] = (char)tolower(q[i]);        // a soft-masked stretch
    for (u32 i = len / 2; i < len / 2 + 60 && i < t.size(); i++) t[i] = 'N';                          // and a run of N
    if (partitionEvery) for (u32 i = partitionEvery; i + 1 < q.size(); i += partitionEvery + (u32)(rnd() % 97)) q[i] = 0;   // a [multi] query: NUL between partitions (sequences.h:188)
    const u32 len1 = (u32)t.size(), len2 = (u32)q.size();
    lzb_seed seed; parse_seed(&seed, pattern, withTrans);
    int8_t ctb[256]; memset(ctb, -1, 256); ctb['A'] = 0; ctb['C'] = 1; ctb['G'] = 2; ctb['T'] = 3;
    lzb_ctx* oc = lzb_open(0); lzb_set_scoring(oc, g_sub.data(), g_msub.data(), 400, 30);
    lzb_target* T = lzb_target_build(oc, (const uint8_t*)t.data(), len1, 0, 0, ctb, &seed, step);
    lzb_query* Q = lzb_query_load(oc, (const uint8_t*)q.data(), len2);
    padded P1, P2; upload(t, P1); upload(q, P2);
    for (u32 i = 0; i < len1; i++) if (P1.cls[i] != g_sc.cls[(u8)t[i]]) { CHECK(false, "k_classify target byte %u", i); break; }
    CHECK(P1.cls[len1] == g_sc.cls[0], "class of the terminator");

    // ---- K1: index (lzb_target_build in index.cu) ----
    const u64 nw = 1ull << seed.weight;
    const u32 L = (u32)seed.length, pmin = ((L + step - 1) / step) * step, pmaxAll = (len1 / step) * step;
    std::vector<u32> off(nw + 2, 0), pos;
    if (pmaxAll >= pmin && len1 >= L) {
        const u64 nent = (pmaxAll - pmin) / step + 1;
        std::vector<u32> keys(nent), vals(nent), hist(nw + 1, 0);
        seed_dev sd; sd.length = seed.length; sd.numParts = seed.numParts; for (int i = 0; i < seed.numParts; i++) { sd.shift[i] = seed.shift[i]; sd.mask[i] = seed.mask[i]; }
        ctb_dev cd; memcpy(cd.v, ctb, 256);
        emu_launch(3, 128, [&]() { k_index_words(P1.asc.data(), 0, pmaxAll, step, nent, sd, cd, seed.weight, keys.data(), vals.data(), hist.data()); });
        std::vector<u32> order(nent); for (u64 i = 0; i < nent; i++) order[i] = (u32)i;
        std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return keys[a] < keys[b]; });   // cub stable radix sort on the word bits
        for (u64 w = 0; w < nw; w++) off[w + 1] = off[w] + hist[w];
        pos.resize(off[nw]);
        for (u64 i = 0; i < nent && keys[order[i]] < nw; i++) pos[i] = vals[order[i]];
    }
    std::vector<u32> wantCounts(nw), wantPos; int64_t np = lzb_target_export_index(T, wantCounts.data(), NULL);
    wantPos.resize(np > 0 ? np : 1); lzb_target_export_index(T, wantCounts.data(), wantPos.data());
    CHECK((int64_t)pos.size() == np, "index size %zu vs %lld", pos.size(), (long long)np);
    bool sameIdx = (int64_t)pos.size() == np;
    for (u64 w = 0; w < nw && sameIdx; w++) sameIdx = off[w + 1] - off[w] == wantCounts[w];
    if (sameIdx && np > 0) sameIdx = memcmp(pos.data(), wantPos.data(), (size_t)np * 4) == 0;
    CHECK(sameIdx, "index contents (counts per word, positions descending)");
    pos.resize(pos.size() + 4);

    for (const mode& M : modes) {
        // ---- oracle ----
        lzb_seed_params sp; memset(&sp, 0, sizeof sp);
        sp.gfExtend = M.gfExtend; sp.gfMismatches = M.mismatches; sp.xDrop = 910; sp.hspThreshold = M.K; sp.entropy = M.entropy; sp.hashBits = M.hashBits; sp.plainHits = M.plain; sp.recoverSeeds = M.recover; sp.twinMinSpan = M.twinMin; sp.twinMaxSpan = M.twinMax; sp.searchLimit = (uint32_t)M.searchLimit;
        lzb_segment* want = NULL; uint64_t nwant = 0; lzb_seed_stats wst;
        if (lzb_seed_hit_search(oc, T, Q, &seed, ctb, &sp, &want, &nwant, &wst)) { CHECK(false, "oracle: %s", lzb_last_error()); continue; }
        // ---- the kernels, orchestrated like lzb_seed_hit_search (one chunk) ----
        std::vector<u32> flips; flips.push_back(0);
        if (seed.withTrans == 1) for (int f = 0; f < seed.numFlips; f++) flips.push_back(seed.transFlips[f]);
        else if (seed.withTrans >= 2) for (int f = 0; f < seed.numFlips; f++) { flips.push_back(seed.transFlips[f]); for (int g = f + 1; g < seed.numFlips; g++) flips.push_back(seed.transFlips[f] ^ seed.transFlips[g]); }
        sp_dev P; memset(&P, 0, sizeof P);
        P.qstart = 0; P.qend = len2; P.len1 = len1; P.len2 = len2; P.L = seed.length; P.V = (int)flips.size(); P.hashBits = M.hashBits;
        P.xDrop = 910; P.K = M.K; P.gfExtend = M.gfExtend; P.plain = M.plain; P.entropy = M.entropy;
        seed_dev sd; sd.length = seed.length; sd.numParts = seed.numParts; for (int i = 0; i < seed.numParts; i++) { sd.shift[i] = seed.shift[i]; sd.mask[i] = seed.mask[i]; }
        ctb_dev2 cd; memcpy(cd.v, ctb, 256);
        const u32 n = len2, nblk = (n + POS_PER_BLOCK - 1) / POS_PER_BLOCK, nbuckets = 1u << M.hashBits;
        std::vector<u32> qword(n + 4); search_counters cnt; memset(&cnt, 0, sizeof cnt);
        std::vector<unsigned long long> blkcnt(nblk);
        emu_launch(2, 256, [&]() { k_query_words(P2.asc.data(), P, sd, cd, qword.data(), &cnt); });
        emu_launch(nblk, 256, [&]() { k_count_hits(qword.data(), off.data(), pos.data(), flips.data(), P, blkcnt.data()); });
        u64 totalHits = 0; for (auto b : blkcnt) totalHits += b;
        if (!M.searchLimit) CHECK(totalHits == wst.rawSeedHits, "%s: raw hits %llu vs %llu", M.name, (unsigned long long)totalHits, (unsigned long long)wst.rawSeedHits);
        if (!M.searchLimit) CHECK(cnt.words == wst.wordsInQuery, "%s: words %llu vs %llu", M.name, cnt.words, (unsigned long long)wst.wordsInQuery);
        const u32 nslots = n * (u32)P.V;
        std::vector<u32> slotcnt(nslots + 2), slotoff(nslots + 2, 0);
        emu_launch(2, 256, [&]() { k_slot_count(qword.data(), off.data(), pos.data(), flips.data(), P, 0, nslots, slotcnt.data()); });
        for (u32 s = 0; s < nslots; s++) slotoff[s + 1] = slotoff[s] + slotcnt[s];
        CHECK(slotoff[nslots] == totalHits, "%s: slot counts", M.name);
        const u32 nh = (u32)totalHits;
        std::vector<u32> keys(nh + 1), bstart(nbuckets + 2); std::vector<u64> vals(nh + 1);
        emu_launch(2, 256, [&]() { k_expand(qword.data(), off.data(), pos.data(), flips.data(), P, 0, nslots, slotoff.data(), keys.data(), vals.data()); });
        std::vector<u32> ord(nh); for (u32 i = 0; i < nh; i++) ord[i] = i;
        std::stable_sort(ord.begin(), ord.end(), [&](u32 a, u32 b) { return keys[a] < keys[b]; });        // cub stable radix sort on the bucket bits
        std::vector<u32> keysB(nh + 1); std::vector<u64> valsB(nh + 1);
        for (u32 i = 0; i < nh; i++) { keysB[i] = keys[ord[i]]; valsB[i] = vals[ord[i]]; }
        emu_launch(1, 256, [&]() { k_bucket_bounds(keysB.data(), nh, nbuckets, bstart.data()); });
        std::vector<u32> diagEnd(nbuckets, 0); std::vector<cand_rec> cand(nh + 16); const u32 candCap = (u32)cand.size();
        if (M.twinMin > 0) {
            /* --twins: k_extend_twin over SEVERAL chunks of the query (chunkBlocks x 1024 positions each), the way the host loop
             * cuts a long query: the bucket extents and the carried queue entries persist from chunk to chunk */
            std::vector<unsigned long long> prefix(nblk + 1, 0);
            for (u32 b = 0; b < nblk; b++) prefix[b + 1] = prefix[b] + blkcnt[b];
            const u32 carryCap = TWIN_CARRY_CAP(M.twinMax); std::vector<twin_ent> carry((size_t)nbuckets * carryCap); std::vector<u32> ncarry(nbuckets, 0);
            const u32 cb = (u32)M.chunkBlocks;
            for (u32 b0 = 0; b0 < nblk; b0 += cb) {
                const u32 k0 = b0 * POS_PER_BLOCK, k1 = std::min<u32>((b0 + cb) * POS_PER_BLOCK, n), ns = (k1 - k0) * (u32)P.V;
                std::vector<u32> sc2(ns + 2), so2(ns + 2, 0);
                emu_launch(2, 256, [&]() { k_slot_count(qword.data(), off.data(), pos.data(), flips.data(), P, k0, ns, sc2.data()); });
                for (u32 s2 = 0; s2 < ns; s2++) so2[s2 + 1] = so2[s2] + sc2[s2];
                const u32 nh2 = so2[ns];
                if (!nh2) continue;
                std::vector<u32> ka(nh2 + 1), kb(nh2 + 1), bs(nbuckets + 2); std::vector<u64> va(nh2 + 1), vb(nh2 + 1);
                emu_launch(2, 256, [&]() { k_expand(qword.data(), off.data(), pos.data(), flips.data(), P, k0, ns, so2.data(), ka.data(), va.data()); });
                std::vector<u32> od(nh2); for (u32 i = 0; i < nh2; i++) od[i] = i;
                std::stable_sort(od.begin(), od.end(), [&](u32 a, u32 b) { return ka[a] < ka[b]; });
                for (u32 i = 0; i < nh2; i++) { kb[i] = ka[od[i]]; vb[i] = va[od[i]]; }
                emu_launch(1, 256, [&]() { k_bucket_bounds(kb.data(), nh2, nbuckets, bs.data()); });
                std::vector<twin_ent> ent(nh2 + 1);
                emu_launch(2, 128, [&]() { k_extend_twin(vb.data(), bs.data(), nbuckets, P1.cls.data(), P2.cls.data(), P1.asc.data(), P2.asc.data(), &g_sc, P, (u32)M.twinMin, (u32)M.twinMax,
                                                         prefix.data(), nblk, 256u * 1024u, diagEnd.data(), ent.data(), carry.data(), carryCap, ncarry.data(), cand.data(), candCap, &cnt); });
            }
            CHECK(cnt.overflow == 0, "%s: the twin replay flagged %llu entries", M.name, cnt.overflow);
        } else if (M.recover) {                                          /* --recoverseeds: k_extend_recover, the bucket's actual diagonal beside its extent */
            std::vector<s32> diagActual(nbuckets, 0);
            emu_launch(2, 128, [&]() { k_extend_recover(valsB.data(), bstart.data(), nbuckets, P1.cls.data(), P2.cls.data(), P1.asc.data(), P2.asc.data(), &g_sc, P, diagEnd.data(), diagActual.data(), cand.data(), candCap, &cnt); });
        } else if (M.gfExtend == LZB_GFEX_EXACT || M.gfExtend == LZB_GFEX_MISMATCH)
            emu_launch(2, 128, [&]() { k_extend_alt(valsB.data(), bstart.data(), nbuckets, P1.asc.data(), P2.asc.data(), P, M.gfExtend == LZB_GFEX_EXACT ? 0 : M.mismatches, diagEnd.data(), cand.data(), candCap, &cnt); });
        else if (M.gfExtend == LZB_GFEX_XDROP && !M.plain && !M.useFirstKernel) {
            std::vector<u32> bcnt(nbuckets), bid(nbuckets); u32 next = 0;
            emu_launch(1, 256, [&]() { k_bucket_sizes(bstart.data(), nbuckets, bcnt.data(), bid.data()); });
            std::stable_sort(bid.begin(), bid.end(), [&](u32 a, u32 b) { return bcnt[a] > bcnt[b]; });   // largest bucket first
            emu_launch(2, 256, [&]() { k_extend2(valsB.data(), bstart.data(), bid.data(), nbuckets, P1.cls.data(), P2.cls.data(), P1.asc.data(), P2.asc.data(), &g_sc, P, diagEnd.data(), cand.data(), candCap, &cnt, &next); });
        } else if (M.useFirstKernel == 2) {                      /* the three-kernel extension (xdrop_split.cuh, LZB_SPLIT_EXTEND=1) */
            std::vector<right_rec> right(nh + 1); std::vector<live_rec> live(nh + 1); unsigned long long nlive = 0;
            emu_launch(2, 256, [&]() { k_right(valsB.data(), nh, P1.cls.data(), P2.cls.data(), &g_sc, P, diagEnd.data(), right.data()); });
            emu_launch(2, 256, [&]() { k_replay(valsB.data(), bstart.data(), nbuckets, right.data(), P1.cls.data(), P2.cls.data(), &g_sc, P, diagEnd.data(), live.data(), &nlive); });
            emu_launch(2, 256, [&]() { k_left(live.data(), &nlive, P1.cls.data(), P2.cls.data(), P1.asc.data(), P2.asc.data(), &g_sc, P, cand.data(), candCap, &cnt); });
        } else
            emu_launch(2, 256, [&]() { k_extend<true>(valsB.data(), bstart.data(), nbuckets, P1.cls.data(), P2.cls.data(), P1.asc.data(), P2.asc.data(), &g_sc, P, diagEnd.data(), cand.data(), candCap, &cnt); });
        // ---- candidates -> HSP table (the host tail of lzb_seed_hit_search) ----
        cand.resize(cnt.ncand);
        struct keyed { u32 hit2, variant, hit1, ix; }; std::vector<keyed> order(cand.size());
        for (size_t i = 0; i < cand.size(); i++) {
            u32 x = host_word_at((const u8*)q.data(), cand[i].hit2, &seed, ctb) ^ host_word_at((const u8*)t.data(), cand[i].hit1, &seed, ctb), v = 0;
            for (; v < flips.size(); v++) if (flips[v] == x) break;
            order[i] = { cand[i].hit2, v, cand[i].hit1, (u32)i };
        }
        std::sort(order.begin(), order.end(), [](const keyed& a, const keyed& b) { return a.hit2 != b.hit2 ? a.hit2 < b.hit2 : a.variant != b.variant ? a.variant < b.variant : a.hit1 > b.hit1; });
        std::vector<lzb_segment> got; u32 limitAt = 0xFFFFFFFFu;
        for (auto& o : order) {
            const cand_rec& r = cand[o.ix]; s32 sim = r.score;
            if (!M.plain && M.gfExtend == LZB_GFEX_XDROP && M.entropy && sim >= M.K && sim <= 3 * M.K) {
                double qf = 1.0;
                if (r.cA + r.cC + r.cG + r.cT >= 20) {
                    double ln = (double)(int)r.length, pA = (int)r.cA / ln, pC = (int)r.cC / ln, pG = (int)r.cG / ln, pT = (int)r.cT / ln;
                    qf = -(pA * (r.cA ? log(pA) : 0.0) + pC * (r.cC ? log(pC) : 0.0) + pG * (r.cG ? log(pG) : 0.0) + pT * (r.cT ? log(pT) : 0.0)) / log(4.0);
                }
                sim = (s32)((double)sim * qf);
                if (sim < M.K) continue;
            }
            if (limitAt != 0xFFFFFFFFu && r.hit2 != limitAt) break;          /* searchLimit: the scan ended with that query position (the host tail of lzb_seed_hit_search does the same) */
            lzb_segment g; memset(&g, 0, sizeof g); g.pos1 = r.pos1; g.pos2 = r.pos2; g.length = r.length; g.s = sim; got.push_back(g);
            if (M.searchLimit > 0 && M.twinMin <= 0 && got.size() > (size_t)M.searchLimit && limitAt == 0xFFFFFFFFu) limitAt = r.hit2;
        }
        CHECK(got.size() == nwant, "%s: %zu HSPs vs %llu", M.name, got.size(), (unsigned long long)nwant);
        bool same = got.size() == nwant;
        for (size_t i = 0; i < got.size() && same; i++) same = got[i].pos1 == want[i].pos1 && got[i].pos2 == want[i].pos2 && got[i].length == want[i].length && got[i].s == want[i].s;
        CHECK(same, "%s: HSP table (coordinates, scores, order)", M.name);
        if (!M.plain && M.gfExtend != LZB_GFEX_NONE && !M.searchLimit) CHECK(cnt.extensions == wst.extensions, "%s: extensions %llu vs %llu", M.name, cnt.extensions, (unsigned long long)wst.extensions);
        if (!M.plain && M.gfExtend == LZB_GFEX_XDROP && !M.searchLimit) CHECK(cnt.bpExtended == wst.bpExtended, "%s: bp extended", M.name);
        // ---- K4: anchor peaks ----
        if (M.gfExtend == LZB_GFEX_XDROP && !M.plain && !got.empty()) {
            std::vector<lzb_segment> pk = got, wpk(want, want + nwant);
            emu_launch(1, 128, [&]() { k_peaks(pk.data(), (u64)pk.size(), P1.cls.data(), P2.cls.data(), &g_sc); });
            lzb_reduce_to_points(oc, T, Q, wpk.data(), wpk.size());
            bool sp2 = true; for (size_t i = 0; i < pk.size() && sp2; i++) sp2 = pk[i].pos1 == wpk[i].pos1 && pk[i].pos2 == wpk[i].pos2 && pk[i].length == 0;
            CHECK(sp2, "%s: anchor peaks", M.name);
        }
        printf("case %d %-28s %u x %u bp seed=%s T=%d step=%u: %llu raw hits, %llu extensions, %zu HSPs  %s\n", caseNo, M.name, len1, len2, pattern, withTrans, step,
               (unsigned long long)totalHits, cnt.extensions, got.size(), g_bad ? "MISMATCH" : "ok");
        cnt.extensions = cnt.bpExtended = cnt.ncand = 0;
        lzb_free(want);
    }
    lzb_query_free(Q); lzb_target_free(T); lzb_close(oc);
}

int main() {
    build_scoring();
    std::vector<mode> a = {
        { "x-drop (k_extend2)",          LZB_GFEX_XDROP, 0, 3000, 0, 1, 16, 0 },
        { "x-drop, 2^6 buckets",         LZB_GFEX_XDROP, 0, 2000, 0, 0, 6, 0 },
        { "x-drop (first k_extend)",     LZB_GFEX_XDROP, 0, 3000, 0, 1, 16, 1 },
        { "x-drop (right/replay/left)",  LZB_GFEX_XDROP, 0, 3000, 0, 1, 16, 2 },
        { "--nogfextend (diag filter)",  LZB_GFEX_NONE, 0, 0, 0, 0, 16, 1 },
        { "--mismatch=2,40",             LZB_GFEX_MISMATCH, 2, 40, 0, 0, 16, 0 },
        { "--recoverseeds, 2^6 buckets", LZB_GFEX_XDROP, 0, 2000, 0, 1, 6, 0, 1 },
        { "--recoverseeds",              LZB_GFEX_XDROP, 0, 3000, 0, 1, 16, 0, 1 },
        { "--recoverseeds --nogfextend, 2^8", LZB_GFEX_NONE, 0, 0, 0, 0, 8, 0, 1 },
        { "--twins=0..40, 2 chunks",     LZB_GFEX_XDROP, 0, 3000, 0, 1, 16, 0, 0, 38, 78, 10 },
        { "--twins=-5..30, 2^6, 20 chunks", LZB_GFEX_XDROP, 0, 2000, 0, 1, 6, 0, 0, 33, 68, 1 },
        { "--twins=10..200 --nogfextend, 2^8", LZB_GFEX_NONE, 0, 0, 0, 0, 8, 0, 0, 48, 238, 3 },
        { "searchLimit 40 (x-drop)",     LZB_GFEX_XDROP, 0, 3000, 0, 1, 16, 0, 0, 0, 0, 0, 40 },
        { "searchLimit 7 (--nogfextend, 2^8)", LZB_GFEX_NONE, 0, 0, 0, 0, 8, 1, 0, 0, 0, 0, 7 },
        { "searchLimit 100 (raw hits)",  LZB_GFEX_NONE, 0, 0, 1, 0, 16, 1, 0, 0, 0, 0, 100 },
    };
    one_pair(0, 20000, "1110100110010101111", 1, 1, a);
    std::vector<mode> b = {
        { "x-drop (k_extend2)",          LZB_GFEX_XDROP, 0, 2500, 0, 1, 16, 0 },
        { "raw hits (plain)",            LZB_GFEX_NONE, 0, 0, 1, 0, 16, 1 },
        { "--exact=30",                  LZB_GFEX_EXACT, 0, 30, 0, 0, 16, 0 },
        { "--mismatch=1,25, 2^8 buckets", LZB_GFEX_MISMATCH, 1, 25, 0, 0, 8, 0 },
        { "--twins=0..30 (match12, step 3), 5 chunks", LZB_GFEX_XDROP, 0, 2500, 0, 1, 16, 0, 0, 24, 54, 3 },
    };
    one_pair(1, 15000, "111111111111", 0, 3, b);
    std::vector<mode> c = {
        { "[multi] query, x-drop",       LZB_GFEX_XDROP, 0, 2200, 0, 1, 16, 0 },
        { "every extension kept (K=top%)", LZB_GFEX_XDROP, 0, -600000000, 0, 0, 16, 0 },   // what an adaptive threshold asks of the library
        { "[multi] query, --exact=25",   LZB_GFEX_EXACT, 0, 25, 0, 0, 16, 0 },
        { "[multi] query, raw hits",     LZB_GFEX_NONE, 0, 0, 1, 0, 16, 1 },
        { "[multi] query, --recoverseeds, 2^7", LZB_GFEX_XDROP, 0, 2200, 0, 1, 7, 0, 1 },
        { "[multi] query, --twins=0..60, 2^7, 12 chunks", LZB_GFEX_XDROP, 0, 2200, 0, 1, 7, 0, 0, 38, 98, 1 },
    };
    one_pair(2, 12000, "1110100110010101111", 1, 1, c, 311);
    printf("%d checks failed, %llu collectives emulated\n", g_bad, emu_collectives);
    return g_bad ? 1 : 0;
}
