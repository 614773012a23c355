/*
 * seed_search.cu -- K2/K3: seed-hit enumeration, diagonal-hash filtering and x-drop extension.
 *
 * Replaces seed_hit_search / private_hit_search / find_table_matches (seed_search.c:322, :464,
 * :810), the default hit processor process_for_simple_hit (:1056) with the diag-hash protocol of
 * diag_hash.h:61-101, xdrop_extend_seed_hit (:2528) and the collect_hsps reporter
 * (lastz.c:3991).
 *
 * The reference handles hits one at a time in discovery order (query position up, seed variant,
 * target position down) and threads every hit through diagEnd[(pos1-pos2) & 0xFFFF].  That state
 * is the ONLY coupling between hits, and it couples only hits in the same bucket.  So:
 *
 *   k_query_words   pack the seed word at every query position                    (hot loop B)
 *   k_count_hits    hits per 1024 query positions -> host cuts the query into chunks that fit
 *   per chunk:
 *     k_slot_count  hits per (position, variant) slot; cub exclusive scan
 *     k_expand      CSR lists -> (pos1,pos2) records in discovery order, key = bucket  (loop C)
 *     cub radix sort on the bucket bits, STABLE => each bucket keeps discovery order
 *     k_bucket_bounds
 *     k_extend      one warp per bucket: 32 hits at a time; right x-drop scans in parallel (they
 *                   do not depend on the bucket state), a 32-step register walk that replays the
 *                   diagEnd test/update exactly, then left scans for the surviving hits (loop D)
 *
 * diagEnd persists in HBM across chunks.  HSP candidates come back with the integer match counts
 * the entropy factor needs; the double/libm part of entropy (dna_utilities.c:2926-2936) and the
 * final ordering are O(#HSPs) host work in lzb_seed_hit_search.
 */
#include <cub/cub.cuh>
#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <vector>
#include "lzb_cuda.h"

#include "seed_kernels.cuh"

/* ------------------------------------------------------------------------------------------ */

#include "xdrop_split.cuh"

static inline u32 host_pack(const lzb_seed* sd, u64 w) {
    u32 p = 0;
    for (int i = 0; i < sd->numParts; i++) p |= (u32)(w >> sd->shift[i]) & sd->mask[i];
    return p;
}
static u32 host_word_at(const u8* v, u32 endPos, const lzb_seed* sd, const int8_t* ctb) {
    u64 w = 0;
    for (int j = 0; j < sd->length; j++) w = (w << 2) | (u64)(ctb[v[endPos - sd->length + j]] & 3);
    return host_pack(sd, w);
}

struct ev_pair { cudaEvent_t a, b; int which; };

/* Device scratch of the seed stage, kept in the context across calls: at 50 Mbp x 50 Mbp the hit buffers are
 * 3.5 GB and cudaMalloc + cudaFree of them cost 0.1-0.7 s of host time per call (LZB_SEED_TRACE) against
 * 0.34 s of device work.  A buffer only ever grows. */
#define SEED_SCRATCH_SLOTS 24
struct seed_scratch { void* p[SEED_SCRATCH_SLOTS]; size_t cap[SEED_SCRATCH_SLOTS]; };
static int scratch_get(lzb_ctx* c, int slot, size_t bytes, void** out) {
    if (!c->seedScratch) c->seedScratch = calloc(1, sizeof(seed_scratch));
    seed_scratch* sc = (seed_scratch*)c->seedScratch;
    if (bytes == 0) bytes = 16;
    if (sc->cap[slot] < bytes) {
        if (sc->p[slot]) { cudaStreamSynchronize(c->stream); cudaFree(sc->p[slot]); sc->p[slot] = NULL; sc->cap[slot] = 0; }
        CUDA_TRY(cudaMalloc(&sc->p[slot], bytes));
        sc->cap[slot] = bytes;
    }
    *out = sc->p[slot];
    return 0;
}
#define SCRATCH(slot_, var_, size_) do { if (scratch_get(c, (slot_), (size_), (void**)&var_)) return -1; } while (0)
void lzb_seed_scratch_free(lzb_ctx* c) {
    seed_scratch* sc = (seed_scratch*)c->seedScratch;
    if (!sc) return;
    for (int i = 0; i < SEED_SCRATCH_SLOTS; i++) cudaFree(sc->p[i]);
    free(sc); c->seedScratch = NULL;
}

extern "C" int lzb_seed_hit_search(lzb_ctx* c, lzb_target* t, lzb_query* q, const lzb_seed* seed,
                                   const int8_t ctb[256], const lzb_seed_params* prm,
                                   lzb_segment** segs, uint64_t* nsegs, lzb_seed_stats* stats) {
    cudaSetDevice(c->device);
    if (!c->haveScoring) return lzb_fail("lzb_set_scoring has not been called");
    u32 qstart = prm->start, qend = prm->end ? prm->end : q->len;
    if (qend <= qstart) return lzb_fail("in seed_hit_search(), interval is void (%u-%u)", qstart, qend);
    if (qend > q->len) return lzb_fail("in seed_hit_search(), interval end is bad (%u>%u)", qend, q->len);
    if (seed->length < 2) return lzb_fail("seed length must be at least two (yours is %d)", seed->length);
    if (prm->gfExtend == LZB_GFEX_MISMATCH && (prm->gfMismatches < 1 || prm->gfMismatches > LZB_GFEX_MISMATCH_MAX))
        return lzb_fail("%d is out of range for N-mismatch (valid range is 1..%d)", prm->gfMismatches, LZB_GFEX_MISMATCH_MAX);
    const bool twin = prm->twinMinSpan > 0;                          /* process_for_twin_hit goes before every other processor (lastz.c:2787) */
    if (twin && prm->gfExtend != LZB_GFEX_XDROP && prm->gfExtend != LZB_GFEX_NONE) return lzb_fail("twins are built for x-drop extension and --nogfextend only");
    if (twin && prm->twinMaxSpan < prm->twinMinSpan) return lzb_fail("maxGap for twins can't be less than min gap");
    const u32 twinQueue = prm->seedQueueSize > 0 ? (u32)prm->seedQueueSize : 256u * 1024u;          /* defaultSeedHitQueueSize diag_hash.h:112 */
    const bool recover = prm->recoverSeeds && !prm->plainHits && !twin;   /* process_for_recoverable_hit; the plain processor wins (lastz.c:2789-2792) */
    if (recover && prm->gfExtend != LZB_GFEX_XDROP && prm->gfExtend != LZB_GFEX_NONE)
        return lzb_fail("recoverSeeds is built for x-drop extension and --nogfextend only");
    if (seed->weight != t->wordBits || seed->length != t->seedLength)
        return lzb_fail("the seed does not match the one the target index was built with");
    int hashBits = prm->hashBits ? prm->hashBits : 16;
    if (hashBits < 4 || hashBits > 26) return lzb_fail("diag hash size 2^%d is not supported", hashBits);
    *segs = NULL; *nsegs = 0;
    if (stats) memset(stats, 0, sizeof *stats);
    cudaStream_t st = c->stream;
    /* LZB_SEED_TRACE=1: host wall clock of the call's phases on stderr (what the CUDA-event time does not see) */
    const bool wtrace = getenv("LZB_SEED_TRACE") != NULL;
    const auto w0 = std::chrono::steady_clock::now(); double wt[8] = {0}; int wk = 0;
#define WMARK() do { if (wtrace && wk < 8) wt[wk++] = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count(); } while (0)

    /* seed variants in probing order (seed_search.c:522-549) */
    std::vector<u32> flips; flips.push_back(0);
    if (seed->withTrans == 1) for (int f = 0; f < seed->numFlips; f++) flips.push_back(seed->transFlips[f]);
    else if (seed->withTrans >= 2)
        for (int f = 0; f < seed->numFlips; f++) {
            flips.push_back(seed->transFlips[f]);
            for (int g = f + 1; g < seed->numFlips; g++) flips.push_back(seed->transFlips[f] ^ seed->transFlips[g]);
        }
    if (flips.size() > MAX_VARIANTS) return lzb_fail("too many seed variants (%zu)", flips.size());
    sp_dev P; memset(&P, 0, sizeof P);
    P.qstart = qstart; P.qend = qend; P.len1 = t->len; P.len2 = q->len; P.L = seed->length; P.V = (int)flips.size();
    P.hashBits = hashBits; P.selfCompare = prm->selfCompare; P.sameStrand = prm->sameStrand;
    P.xDrop = prm->xDrop; P.K = prm->hspThreshold; P.gfExtend = prm->gfExtend; P.plain = prm->plainHits && !twin; P.entropy = prm->entropy;
    seed_dev sd; seed_to_dev(&sd, seed);
    ctb_dev2 cd; memcpy(cd.v, ctb, 256);

    u32 n = qend - qstart;
    u32 nblk = (n + POS_PER_BLOCK - 1) / POS_PER_BLOCK;
    u32 nbuckets = 1u << hashBits;
    u32 *d_qword = NULL, *d_flips = NULL, *d_E = NULL, *d_bstart = NULL;
    unsigned long long* d_blkcnt = NULL; search_counters* d_cnt = NULL;
    std::vector<ev_pair> evs;
    cudaEvent_t evBegin, evMid, evMid2, evEnd;       /* device time = [begin,mid] + [mid2,end]; planning + cudaMalloc sit between */
    CUDA_TRY(cudaEventCreate(&evBegin)); CUDA_TRY(cudaEventCreate(&evEnd));
    CUDA_TRY(cudaEventCreate(&evMid)); CUDA_TRY(cudaEventCreate(&evMid2));
    SCRATCH(0, d_qword, (size_t)n * 4 + 16);
    SCRATCH(1, d_flips, flips.size() * 4);
    SCRATCH(2, d_blkcnt, (size_t)nblk * 8);
    SCRATCH(3, d_cnt, sizeof(search_counters));
    SCRATCH(4, d_E, (size_t)nbuckets * 4);
    SCRATCH(5, d_bstart, ((size_t)nbuckets + 2) * 4);
    CUDA_TRY(cudaMemcpyAsync(d_flips, flips.data(), flips.size() * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(d_cnt, 0, sizeof(search_counters), st));
    CUDA_TRY(cudaMemsetAsync(d_E, 0, (size_t)nbuckets * 4, st));        /* empty_diag_hash diag_hash.c:125 */
    s32* d_A = NULL;                                                     /* diagActual[] (diag_hash.h:70), recoverable processor only */
    if (recover) { SCRATCH(22, d_A, (size_t)nbuckets * 4); CUDA_TRY(cudaMemsetAsync(d_A, 0, (size_t)nbuckets * 4, st)); }
    CUDA_TRY(cudaEventRecord(evBegin, st));
    int grid = c->smCount * 8;
#define TIMED(which_, launch_)                                                          \
    do { ev_pair e_; e_.which = (which_); cudaEventCreate(&e_.a); cudaEventCreate(&e_.b); \
         cudaEventRecord(e_.a, st); launch_; cudaEventRecord(e_.b, st); evs.push_back(e_); \
         c->launches++; } while (0)
    TIMED(0, (k_query_words<<<grid, 256, 0, st>>>(q->d_seq, P, sd, cd, d_qword, d_cnt)));
    TIMED(1, (k_count_hits<<<nblk, 256, 0, st>>>(d_qword, t->d_off, t->d_pos, d_flips, P, d_blkcnt)));
    CUDA_TRY(cudaGetLastError());
    std::vector<unsigned long long> blkcnt(nblk);
    CUDA_TRY(cudaMemcpyAsync(blkcnt.data(), d_blkcnt, (size_t)nblk * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(evMid, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    WMARK();                                            /* [0] small allocations + words + hit counts + sync */
    u64 totalHits = 0, maxBlk = 0;
    for (u32 b = 0; b < nblk; b++) { totalHits += blkcnt[b]; if (blkcnt[b] > maxBlk) maxBlk = blkcnt[b]; }
    /* twins: the reference's seed hit queue holds the last twinQueue entries of ALL buckets.  The bucket replay is exact
     * only while no hit within maxSpan columns can have been pushed out of it (the reference prints "seed hit queue
     * shortfall" when that happens): refuse inputs whose hit density over any window that wide reaches the queue size.
     * Blocks consulted long after they were made are checked in the kernel against the same prefix sums. */
    unsigned long long* d_blkprefix = NULL; twin_ent* d_ent = NULL; twin_ent* d_carry = NULL; u32* d_ncarry = NULL;
    if (twin) {
        std::vector<unsigned long long> prefix(nblk + 1, 0);
        for (u32 b = 0; b < nblk; b++) prefix[b + 1] = prefix[b] + blkcnt[b];
        const u32 wblocks = ((u32)prm->twinMaxSpan + (u32)seed->length + POS_PER_BLOCK - 1) / POS_PER_BLOCK + 1;
        for (u32 b = 0; b < nblk; b++) {
            const u32 e = std::min<u64>((u64)b + wblocks, nblk);
            if (prefix[e] - prefix[b] + 2 >= twinQueue)
                return lzb_fail("seed hit queue shortfall: %llu seed hits within %u query positions reach the queue size %u (--seedqueue); "
                                "the twin processor's result would depend on what the queue has forgotten", prefix[e] - prefix[b], wblocks * POS_PER_BLOCK, twinQueue);
        }
        SCRATCH(22, d_blkprefix, ((size_t)nblk + 1) * 8);
        CUDA_TRY(cudaMemcpyAsync(d_blkprefix, prefix.data(), ((size_t)nblk + 1) * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));                          /* (prefix is a local) */
        SCRATCH(23, d_carry, (size_t)nbuckets * TWIN_CARRY_CAP(prm->twinMaxSpan) * sizeof(twin_ent));
        SCRATCH(13, d_ncarry, (size_t)nbuckets * 4);
        CUDA_TRY(cudaMemsetAsync(d_ncarry, 0, (size_t)nbuckets * 4, st));
    }

    const char* capEnv = getenv("LZB_HIT_CAP");
    u64 hitCap = capEnv ? strtoull(capEnv, 0, 10) : (1ull << 27);
    if (hitCap > 0xFFFFFFF0ull) hitCap = 0xFFFFFFF0ull;
    if (maxBlk > hitCap) hitCap = maxBlk;
    if (hitCap > 0xFFFFFFF0ull) return lzb_fail("%llu seed hits within %d query positions; use --step or mask repeats", (unsigned long long)maxBlk, POS_PER_BLOCK);
    if (totalHits < hitCap) hitCap = totalHits ? totalHits : 1;
    u64 slotCap = 1ull << 26;
    { u64 allSlots = (u64)nblk * POS_PER_BLOCK * P.V; if (allSlots < slotCap) slotCap = allSlots; }
    if (slotCap < (u64)POS_PER_BLOCK * P.V) slotCap = (u64)POS_PER_BLOCK * P.V;
    /* (the twin and recoverable processors extend again and again along a homologous diagonal -- the merge comes later -- so
     * their candidate count is bounded by the hits, not by the number of distinct HSPs) */
    u32 candCap = prm->plainHits || prm->gfExtend != LZB_GFEX_XDROP || twin || recover ? (u32)std::min<u64>(totalHits + 1, 1ull << 26) : (1u << 22);
    if (c->candCapWanted > candCap) candCap = (u32)std::min<u64>(c->candCapWanted, totalHits + 1);      /* an earlier call ran out of room */

    u32 *d_slotcnt = NULL, *d_slotoff = NULL, *keysA = NULL, *keysB = NULL; u64 *valsA = NULL, *valsB = NULL;
    cand_rec* d_cand = NULL; void* d_tmp = NULL; size_t tmpBytes = 0, tmpScan = 0;
    SCRATCH(6, d_slotcnt, (slotCap + 2) * 4); SCRATCH(7, d_slotoff, (slotCap + 2) * 4);
    SCRATCH(8, keysA, hitCap * 4); SCRATCH(9, keysB, hitCap * 4);
    SCRATCH(10, valsA, hitCap * 8); SCRATCH(11, valsB, hitCap * 8);
    SCRATCH(12, d_cand, (size_t)candCap * sizeof(cand_rec));
    /* the three-kernel extension (xdrop_split.cuh) measured SLOWER than the fused kernel (0.70 s vs 0.62 s
     * at 50 Mbp x 50 Mbp: the extra passes over the hit records cost more than the lockstep idling they
     * remove), so it is opt-in: LZB_SPLIT_EXTEND=1.  Kept because it bounds the work on long repeats. */
    const bool splitExtend = prm->gfExtend == LZB_GFEX_XDROP && !prm->plainHits && !recover && !twin && c->sc.numClasses <= 16 &&
                             getenv("LZB_SPLIT_EXTEND") && atoi(getenv("LZB_SPLIT_EXTEND"));
    /* the warp-cooperative kernel (xdrop_warp.cuh) is the default x-drop path; LZB_EXTEND_V1=1 keeps the first one */
    const bool coopExtend = prm->gfExtend == LZB_GFEX_XDROP && !prm->plainHits && !recover && !twin && c->sc.numClasses <= XD_LUT_MAX_CLASSES && !splitExtend &&
                            !(getenv("LZB_EXTEND_V1") && atoi(getenv("LZB_EXTEND_V1")));
    u32 *d_bcnt = NULL, *d_bcnt2 = NULL, *d_bid = NULL, *d_border = NULL, *d_next = NULL; size_t tmpOrder = 0;
    /* persistent CTAs of k_extend2 per SM: four fill the register file; a caller that runs this stage beside another
     * context's Y-drop sweeps can ask for fewer so that those keep their issue slots (lzb_seed_params.extendCtasPerSm; LZB_EXTEND_CTAS_PER_SM for experiments) */
    int extCtas = 4;
    { const char* e = getenv("LZB_EXTEND_CTAS_PER_SM"); if (e) { int v = atoi(e); if (v >= 1 && v <= 4) extCtas = v; } }
    if (prm->extendCtasPerSm >= 1 && prm->extendCtasPerSm <= 4) extCtas = prm->extendCtasPerSm;
    if (coopExtend) {
        SCRATCH(13, d_bcnt, (size_t)nbuckets * 4); SCRATCH(14, d_bcnt2, (size_t)nbuckets * 4);
        SCRATCH(15, d_bid, (size_t)nbuckets * 4); SCRATCH(16, d_border, (size_t)nbuckets * 4);
        SCRATCH(17, d_next, 4);
        CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(NULL, tmpOrder, d_bcnt, d_bcnt2, d_bid, d_border, nbuckets, 0, 32, st));
    }
    right_rec* d_right = NULL; live_rec* d_live = NULL; unsigned long long* d_nlive = NULL;
    if (splitExtend) {
        SCRATCH(18, d_right, hitCap * sizeof(right_rec));
        SCRATCH(19, d_live, hitCap * sizeof(live_rec));
        SCRATCH(20, d_nlive, 8);
    }
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(NULL, tmpBytes, keysA, keysB, valsA, valsB, hitCap, 0, hashBits, st));
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(NULL, tmpScan, d_slotcnt, d_slotoff, slotCap + 1, st));
    if (tmpScan > tmpBytes) tmpBytes = tmpScan;
    if (tmpOrder > tmpBytes) tmpBytes = tmpOrder;
    SCRATCH(21, d_tmp, tmpBytes);
    if (twin) SCRATCH(18, d_ent, hitCap * sizeof(twin_ent));

    /* chunk loop */
    WMARK();                                            /* [1] hit/slot/candidate buffers allocated */
    CUDA_TRY(cudaEventRecord(evMid2, st));
    u64 chunks = 0;
    for (u32 b0 = 0; b0 < nblk;) {
        u64 h = 0; u32 b1 = b0;
        while (b1 < nblk && h + blkcnt[b1] <= hitCap && (u64)(b1 + 1 - b0) * POS_PER_BLOCK * P.V <= slotCap) { h += blkcnt[b1]; b1++; }
        if (b1 == b0) return lzb_fail("internal error: seed-hit chunk planner made no progress");
        u32 k0 = b0 * POS_PER_BLOCK, k1 = std::min<u64>((u64)b1 * POS_PER_BLOCK, n);
        if (h > 0) {
            u32 nslots = (k1 - k0) * (u32)P.V, nh = (u32)h;
            size_t tb = tmpBytes;
            TIMED(2, (k_slot_count<<<grid, 256, 0, st>>>(d_qword, t->d_off, t->d_pos, d_flips, P, k0, nslots, d_slotcnt)));
            TIMED(3, cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_slotcnt, d_slotoff, nslots + 1, st));
            TIMED(4, (k_expand<<<grid, 256, 0, st>>>(d_qword, t->d_off, t->d_pos, d_flips, P, k0, nslots, d_slotoff, keysA, valsA)));
            tb = tmpBytes;
            TIMED(5, cub::DeviceRadixSort::SortPairs(d_tmp, tb, keysA, keysB, valsA, valsB, nh, 0, hashBits, st));
            TIMED(6, (k_bucket_bounds<<<(nbuckets + 256) / 256, 256, 0, st>>>(keysB, nh, nbuckets, d_bstart)));
            if (twin) {
                TIMED(7, (k_extend_twin<<<(nbuckets + 127) / 128, 128, 0, st>>>(valsB, d_bstart, nbuckets, t->d_cls, q->d_cls, t->d_seq, q->d_seq, c->d_sc, P,
                                                                                  (u32)prm->twinMinSpan, (u32)prm->twinMaxSpan, d_blkprefix, nblk, twinQueue,
                                                                                  d_E, d_ent, d_carry, TWIN_CARRY_CAP(prm->twinMaxSpan), d_ncarry, d_cand, candCap, d_cnt)));
            } else if (recover) {
                TIMED(7, (k_extend_recover<<<(nbuckets + 127) / 128, 128, 0, st>>>(valsB, d_bstart, nbuckets, t->d_cls, q->d_cls, t->d_seq, q->d_seq,
                                                                                     c->d_sc, P, d_E, d_A, d_cand, candCap, d_cnt)));
            } else if (splitExtend) {
                CUDA_TRY(cudaMemsetAsync(d_nlive, 0, 8, st));
                TIMED(8, (k_right<<<c->smCount * 8, 256, 0, st>>>(valsB, nh, t->d_cls, q->d_cls, c->d_sc, P, d_E, d_right)));
                TIMED(9, (k_replay<<<grid, 256, 0, st>>>(valsB, d_bstart, nbuckets, d_right, t->d_cls, q->d_cls, c->d_sc, P, d_E, d_live, d_nlive)));
                TIMED(10, (k_left<<<c->smCount * 8, 256, 0, st>>>(d_live, d_nlive, t->d_cls, q->d_cls, t->d_seq, q->d_seq, c->d_sc, P, d_cand, candCap, d_cnt)));
            } else if (prm->gfExtend == LZB_GFEX_EXACT || prm->gfExtend == LZB_GFEX_MISMATCH) {
                TIMED(7, (k_extend_alt<<<(nbuckets + 127) / 128, 128, 0, st>>>(valsB, d_bstart, nbuckets, t->d_seq, q->d_seq, P,
                                                                                 prm->gfExtend == LZB_GFEX_EXACT ? 0 : prm->gfMismatches,
                                                                                 d_E, d_cand, candCap, d_cnt)));
            } else if (coopExtend) {
                int bits = 1; while (bits < 32 && (nh >> bits)) bits++;          /* bucket sizes are <= nh */
                size_t tbo = tmpBytes;
                CUDA_TRY(cudaMemsetAsync(d_next, 0, 4, st));
                TIMED(11, (k_bucket_sizes<<<(nbuckets + 255) / 256, 256, 0, st>>>(d_bstart, nbuckets, d_bcnt, d_bid)));
                TIMED(11, cub::DeviceRadixSort::SortPairsDescending(d_tmp, tbo, d_bcnt, d_bcnt2, d_bid, d_border, nbuckets, 0, bits, st));
                TIMED(7, (k_extend2<<<c->smCount * extCtas, 256, 0, st>>>(valsB, d_bstart, d_border, nbuckets, t->d_cls, q->d_cls, t->d_seq, q->d_seq,
                                                                     c->d_sc, P, d_E, d_cand, candCap, d_cnt, d_next)));
            } else if (c->sc.numClasses <= 16)
                TIMED(7, (k_extend<true><<<grid, 256, 0, st>>>(valsB, d_bstart, nbuckets, t->d_cls, q->d_cls, t->d_seq, q->d_seq,
                                                               c->d_sc, P, d_E, d_cand, candCap, d_cnt)));
            else
                TIMED(7, (k_extend<false><<<grid, 256, 0, st>>>(valsB, d_bstart, nbuckets, t->d_cls, q->d_cls, t->d_seq, q->d_seq,
                                                                c->d_sc, P, d_E, d_cand, candCap, d_cnt)));
            CUDA_TRY(cudaGetLastError());
            chunks++;
        }
        b0 = b1;
    }
    CUDA_TRY(cudaEventRecord(evEnd, st));
    WMARK();                                            /* [2] chunk loop enqueued */
    search_counters hc;
    CUDA_TRY(cudaMemcpyAsync(&hc, d_cnt, sizeof hc, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    WMARK();                                            /* [3] device work finished */
    if (twin && hc.overflow) {
        lzb_fail("the twin processor's seed hit queue cannot be replayed exactly on this input (%llu times: more than %u live entries in one hash "
                 "bucket, or an extension's block consulted after %u or more newer seed hits); try a larger --seedqueue", hc.overflow, TWIN_CARRY_CAP(prm->twinMaxSpan), twinQueue);
        goto cleanup_fail;
    }
    if (hc.ncand > candCap) {
        /* the candidate buffer was too small (the kernels count every candidate and store only what fits): the exact count is
         * known now, so the call is made again with room for all of them -- once; the context remembers the size */
        if (hc.ncand <= (1ull << 28) && c->candCapWanted < hc.ncand) {
            for (auto& e : evs) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
            c->candCapWanted = (unsigned)std::min<u64>(hc.ncand + hc.ncand / 8 + 1024, 1ull << 28);
            return lzb_seed_hit_search(c, t, q, seed, ctb, prm, segs, nsegs, stats);
        }
        lzb_fail("%llu HSP candidates exceed the %u-entry result buffer; raise the threshold (--hspthresh)", hc.ncand, candCap);
        goto cleanup_fail;
    }
    {
        std::vector<cand_rec> cand(hc.ncand);
        if (hc.ncand) CUDA_TRY(cudaMemcpy(cand.data(), d_cand, (size_t)hc.ncand * sizeof(cand_rec), cudaMemcpyDeviceToHost));
        /* discovery order = (query position, variant index, target position descending).  Candidates
         * that share a query position are rare, so: sort (hit2, index) as plain 64-bit keys, then order
         * each run of equal hit2 by the seed variant that produced the hit (recomputed from the bytes) */
        struct keyed { u32 variant; u32 hit1; u32 ix; };
        std::vector<u64> order(cand.size());
        for (size_t i = 0; i < cand.size(); i++) order[i] = ((u64)cand[i].hit2 << 32) | (u32)i;
        std::sort(order.begin(), order.end());
        for (size_t i = 0; i < order.size();) {
            size_t j = i + 1;
            while (j < order.size() && (order[j] >> 32) == (order[i] >> 32)) j++;
            if (j - i > 1) {
                std::vector<keyed> run(j - i);
                for (size_t k = i; k < j; k++) {
                    const cand_rec& r = cand[(u32)order[k]];
                    u32 x = host_word_at(q->h_seq, r.hit2, seed, ctb) ^ host_word_at(t->h_seq, r.hit1, seed, ctb), v = 0;
                    for (; v < flips.size(); v++) if (flips[v] == x) break;
                    run[k - i] = { v, r.hit1, (u32)order[k] };
                }
                std::sort(run.begin(), run.end(), [](const keyed& a, const keyed& b) {
                    if (a.variant != b.variant) return a.variant < b.variant;
                    return a.hit1 > b.hit1;
                });
                for (size_t k = i; k < j; k++) order[k] = (order[k] & 0xFFFFFFFF00000000ull) | run[k - i].ix;
            }
            i = j;
        }
        lzb_segment* out = (lzb_segment*)malloc((cand.size() + 1) * sizeof(lzb_segment));
        u64 m = 0; u32 limitAt = 0xFFFFFFFFu;
        for (size_t i = 0; i < order.size(); i++) {
            const cand_rec& r = cand[(u32)order[i]];
            s32 sim = r.score;
            if (!prm->plainHits && prm->gfExtend == LZB_GFEX_XDROP && prm->entropy &&
                sim >= prm->hspThreshold && sim <= 3 * prm->hspThreshold) {
                /* entropy() dna_utilities.c:2918-2936 on the device's integer counts */
                double qf = 1.0;
                if (r.cA + r.cC + r.cG + r.cT >= 20) {
                    double len = (double)(int)r.length;
                    double pA = (double)(int)r.cA / len, pC = (double)(int)r.cC / len, pG = (double)(int)r.cG / len, pT = (double)(int)r.cT / len;
                    double qA = r.cA ? log(pA) : 0.0, qC = r.cC ? log(pC) : 0.0, qG = r.cG ? log(pG) : 0.0, qT = r.cT ? log(pT) : 0.0;
                    qf = -(pA * qA + pC * qC + pG * qG + pT * qT) / log(4.0);
                }
                sim = (s32)((double)sim * qf);
                if (sim < prm->hspThreshold) continue;
            }
            /* seed_hit_search's searchLimit (seed_search.c:551, :1190): the scan ends with the query position whose hits took
             * the number of reported HSPs past the limit; later positions were never looked at */
            if (limitAt != 0xFFFFFFFFu && r.hit2 != limitAt) break;
            lzb_segment* g = &out[m++];
            memset(g, 0, sizeof *g);
            g->pos1 = r.pos1; g->pos2 = r.pos2; g->length = r.length; g->s = sim; g->id = prm->strandId; g->scoreCov = r.length;
            if (prm->searchLimit > 0 && !twin && m > prm->searchLimit && limitAt == 0xFFFFFFFFu) limitAt = r.hit2;   /* (the twin processor never counts: no searchToGo-- in :1814-2046) */
        }
        *segs = out; *nsegs = m;
        WMARK();                                        /* [4] candidates copied back, entropy, ordering */
        if (stats) {
            stats->wordsInQuery = hc.words; stats->rawSeedHits = totalHits; stats->extensions = hc.extensions;
            stats->bpExtended = hc.bpExtended; stats->hsps = m;
            float ms = 0, ms2 = 0; cudaEventElapsedTime(&ms, evBegin, evMid); cudaEventElapsedTime(&ms2, evMid2, evEnd);
            stats->seconds = (ms + ms2) / 1e3;
            for (auto& e : evs) { float x = 0; cudaEventElapsedTime(&x, e.a, e.b); stats->kernelSeconds[e.which] += x / 1e3; stats->kernelLaunches[e.which]++; }
        }
    }
    for (auto& e : evs) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    cudaEventDestroy(evBegin); cudaEventDestroy(evEnd); cudaEventDestroy(evMid); cudaEventDestroy(evMid2);
    WMARK();                                            /* [5] buffers freed */
    if (wtrace) fprintf(stderr, "[seed trace] plan=%.4f alloc=%.4f enqueue=%.4f device_wait=%.4f post=%.4f free=%.4f total=%.4f s (chunks=%llu hits=%llu cand=%llu)\n",
                        wt[0], wt[1] - wt[0], wt[2] - wt[1], wt[3] - wt[2], wt[4] - wt[3], wt[5] - wt[4], wt[5],
                        (unsigned long long)chunks, (unsigned long long)totalHits, (unsigned long long)hc.ncand);
    return 0;
cleanup_fail:
    for (auto& e : evs) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    return -1;
}
