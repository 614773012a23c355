/*
 * gapped.cu -- K4/K5: anchor peaks, Y-drop banded gapped DP with traceback, and the
 * best-score-first anchor loop.
 *
 * Replaces reduce_to_points / segment_peak (gapped_extend.c:463, :515), gapped_extend (:1012),
 * ydrop_align (:2459), ydrop_one_sided_align (:3388), update_LR_bounds (:4588),
 * update_active_segs (:4885) and format_alignment (:5153).
 *
 * DP kernel (k_ydrop): one 256-thread CTA per one-sided alignment.  A sweep row lives in shared
 * memory (ring-indexed by column); the threads own contiguous column chunks.
 * The reference visits a row left to right with three loop-carried values: the insertion score
 * I, the running bestScore (which moves the prune threshold WITHIN the row) and the band edges.
 * Here a row is three short passes joined by block-wide scans:
 *   1. max-plus scan of the insertion chain I (affine maps x -> max(a, x - e), reset at cells
 *      masked by earlier alignments),
 *   2. cell values + traceback links, then an exclusive prefix-max of the cells that may raise
 *      bestScore (diagonal winners) => the exact threshold each cell saw in the reference,
 *   3. pruning against that threshold, band edges, new best / end cell.
 * Pruned cells are kept out of the result exactly as in the reference: a pruned cell's score is
 * below the threshold, the threshold never decreases, so anything derived from it stays below
 * every later threshold and can neither survive nor be traced back through.
 * One traceback byte per visited cell goes to HBM (the path's only unavoidable traffic); the
 * same warp then walks it back 32 diagonal steps at a time and emits run-length edit ops.
 *
 * Anchor loop (host, C++): identical order and bookkeeping to the reference.  Because each
 * alignment constrains later ones, anchors are extended SPECULATIVELY on a pool of streams
 * ("lanes") against the alignments committed so far, and committed strictly in score order; a speculative result is
 * used only if no alignment committed after its launch overlaps the rows its DP examined,
 * otherwise it is recomputed -- so the output equals the sequential algorithm's.
 */
#include <algorithm>
#include <chrono>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <vector>
#include "lzb_cuda.h"

#include "ydrop_common.cuh"

#include "ydrop_smem.cuh"
#include "ydrop_warp.cuh"
#include "ydrop_mw.cuh"

#include "peaks_kernel.cuh"

extern "C" int lzb_reduce_to_points(lzb_ctx* c, lzb_target* t, lzb_query* q, lzb_segment* anchors, uint64_t n) {
    cudaSetDevice(c->device);
    if (n == 0) return 0;
    lzb_segment* d = NULL;
    CUDA_TRY(cudaMalloc(&d, n * sizeof(lzb_segment)));
    CUDA_TRY(cudaMemcpyAsync(d, anchors, n * sizeof(lzb_segment), cudaMemcpyHostToDevice, c->stream));
    int blocks = (int)((n + 127) / 128); if (blocks > c->smCount * 16) blocks = c->smCount * 16;
    k_peaks<<<blocks, 128, 0, c->stream>>>(d, n, t->d_cls, q->d_cls, c->d_sc);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(anchors, d, n * sizeof(lzb_segment), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    cudaFree(d);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * host side of gapped_extend
 * ---------------------------------------------------------------------------------------- */

struct hseg { int type; u32 b1, b2, e1, e2; };
static const segref NOSEG = { -1, -1 };

struct galn {                          /* galign gapped_extend.c:214-245 */
    u32 pos1, pos2, end1, end2; u64 hspId;
    std::vector<hseg> segs;
    segref left1, right1, left2, right2;
    lzb_alignel* align;
    int next, prev;
    int devIx;                         /* index in the device alignment table once committed */
};

#define MAX_RPT ((1u << 30) - 1)
static lzb_editscript* es_new(u32 cap) {
    if (cap < 16) cap = 16;
    lzb_editscript* s = (lzb_editscript*)calloc(1, sizeof(lzb_editscript) + (size_t)(cap - 1) * 4);
    s->size = cap; return s;
}
/* edit_script_add edit_script.c:261 applied to an already run-length-encoded op */
static void es_add(lzb_editscript** ps, u32 op, u32 rpt) {
    lzb_editscript* s = *ps;
    if (s->len > 0 && (s->tailOp & 3) == op) {
        u32 tr = s->op[s->len - 1] >> 2;
        if ((u64)tr + rpt <= MAX_RPT) { s->op[s->len - 1] += rpt << 2; return; }
        s->op[s->len - 1] = op | (MAX_RPT << 2); rpt = tr + rpt - MAX_RPT;
    }
    if (s->len + 2 > s->size) {
        u32 nsz = s->size * 2 + 16;
        s = (lzb_editscript*)realloc(s, sizeof(lzb_editscript) + (size_t)(nsz - 1) * 4); s->size = nsz; *ps = s;
    }
    while (rpt > MAX_RPT) { s->op[s->len++] = op | (MAX_RPT << 2); rpt -= MAX_RPT; }
    s->op[s->len++] = op | (rpt << 2); s->tailOp = op;
}

struct dp_result { s32 score; u32 end1, end2, rows; int status; unsigned long long cells; std::vector<u32> ops; };

struct gx {                             /* state of one lzb_gapped_extend call */
    lzb_ctx* c; lzb_target* t; lzb_query* q;
    const lzb_gapped_params* P;
    std::vector<galn> al; int obi, oed;
    std::vector<int> committed;         /* host alignment indices in commit order = device table order */
    std::vector<dseg> hsegs;            /* device segment table (host mirror) */
    std::vector<dalign> haligns;
    lzb_gapped_stats st;
};

/* msp_left_right gapped_extend.c:3953-4040 */
static bool anchor_neighbours(gx& G, galn& m, int* coverer = NULL) {
    u32 pos1 = m.pos1, pos2 = m.pos2, right = 0xFFFFFFFFu, left = 0xFFFFFFFFu;
    segref R = NOSEG, Lf = NOSEG;
    for (int o = G.obi; o >= 0 && G.al[o].pos1 <= pos1; o = G.al[o].next) {
        galn& x = G.al[o];
        if (x.end1 < pos1) continue;
        /* first segment whose e1 >= pos1 (e1 never decreases along an alignment) */
        int ns = (int)x.segs.size(), k = 0, hi2 = ns;
        while (k < hi2) { int mid = (k + hi2) >> 1; if (x.segs[mid].e1 < pos1) k = mid + 1; else hi2 = mid; }
        if (k == ns) continue;
        hseg& bp = x.segs[k]; s32 d;
        if (bp.type == SEG_DIAG) d = (s32)(bp.b2 - pos2) + (s32)(pos1 - bp.b1); else d = (s32)(bp.b2 - pos2);
        if (d == 0) { if (coverer) *coverer = o; return false; }
        if (d > 0 && (u32)d < right) { right = (u32)d; R.al = o; R.sg = k; }
        else if (d < 0 && (u32)-d < left) { left = (u32)-d; Lf.al = o; Lf.sg = k; }
    }
    m.right1 = m.right2 = R; m.left1 = m.left2 = Lf;
    return true;
}

/* align_left_right gapped_extend.c:4078-4175 */
static void alignment_neighbours(gx& G, galn& m) {
    u32 pos1 = m.pos1, pos2 = m.pos2, end1 = m.end1, end2 = m.end2;
    u32 rB = 0xFFFFFFFFu, rT = rB, lB = rB, lT = rB;
    segref RB = NOSEG, RT = NOSEG, LB = NOSEG, LT = NOSEG;
    for (int o = G.obi; o >= 0; o = G.al[o].next) {
        galn& x = G.al[o];
        if (x.pos1 > end1 || x.end1 < pos1) continue;
        int k = 0, ns = (int)x.segs.size();
        while (k < ns && !(x.segs[k].type != SEG_HORZ && x.segs[k].e1 >= pos1)) k++;
        if (k < ns && x.segs[k].b1 <= pos1) {
            hseg& bp = x.segs[k]; s32 d;
            if (bp.type == SEG_DIAG) d = (s32)(bp.b2 - pos2) + (s32)(pos1 - bp.b1); else d = (s32)(bp.b2 - pos2);
            if (d > 0 && (u32)d < rB) { rB = (u32)d; RB.al = o; RB.sg = k; }
            else if (d < 0 && (u32)-d < lB) { lB = (u32)-d; LB.al = o; LB.sg = k; }
        }
        while (k < ns && !(x.segs[k].type != SEG_HORZ && x.segs[k].e1 >= end1)) k++;
        if (k < ns) {
            hseg& bp = x.segs[k]; s32 d;
            if (bp.type == SEG_DIAG) d = (s32)(bp.b2 - end2) + (s32)(end1 - bp.b1); else d = (s32)(bp.b2 - end2);
            if (d > 0 && (u32)d < rT) { rT = (u32)d; RT.al = o; RT.sg = k; }
            else if (d < 0 && (u32)-d < lT) { lT = (u32)-d; LT.al = o; LT.sg = k; }
        }
    }
    m.right1 = RB; m.right2 = RT; m.left1 = LB; m.left2 = LT;
}

/* insert_align gapped_extend.c:4210-4240 */
static void list_insert(gx& G, int mi) {
    galn& m = G.al[mi];
    int qq = -1, p = G.obi;
    while (p >= 0 && G.al[p].pos1 < m.pos1) { qq = p; p = G.al[p].next; }
    if (qq >= 0) { G.al[qq].next = mi; m.next = p; } else { m.next = G.obi; G.obi = mi; }
    qq = -1; p = G.oed;
    while (p >= 0 && G.al[p].end1 > m.end1) { qq = p; p = G.al[p].prev; }
    if (qq >= 0) { G.al[qq].prev = mi; m.prev = p; } else { m.prev = G.oed; G.oed = mi; }
}

/* save_seg gapped_extend.c:5220-5262 */
static void add_diag(galn& m, u32 b1, u32 b2, u32 e1, u32 e2) {
    if (!m.segs.empty()) {
        hseg& last = m.segs.back();
        hseg g; g.type = (b1 == last.e1 + 1) ? SEG_HORZ : SEG_VERT;
        g.b1 = last.e1 + 1; g.b2 = last.e2 + 1; g.e1 = b1 - 1; g.e2 = b2 - 1;
        m.segs.push_back(g);
    }
    hseg d = { SEG_DIAG, b1, b2, e1, e2 };
    m.segs.push_back(d);
}

/* score_alignment gapped_extend.c:5631-5690 (host bytes; O(alignment length), only after lopping) */
static s32 rescore(gx& G, u32 p1, u32 p2, lzb_editscript* s) {
    const s32* sub = G.c->hostSub; s32 sim = 0;
    const u8* s1 = G.t->h_seq; const u8* s2 = G.q->h_seq;
    for (u32 k = 0; k < s->len; k++) {
        u32 rpt = s->op[k] >> 2, op = s->op[k] & 3;
        if (!rpt) continue;
        if (op == LZB_OP_SUB) { for (u32 j = 0; j < rpt; j++) sim += sub[(u32)s1[p1 + j] * 256 + s2[p2 + j]]; p1 += rpt; p2 += rpt; }
        else if (op == LZB_OP_INS) { sim -= G.c->sc.gapOpen + (s32)rpt * G.c->sc.gapExtend; p2 += rpt; }
        else { sim -= G.c->sc.gapOpen + (s32)rpt * G.c->sc.gapExtend; p1 += rpt; }
    }
    return sim;
}

static segref dev_ref(gx& G, segref r) { segref o = NOSEG; if (r.al >= 0) { o.al = G.al[r.al].devIx; o.sg = r.sg; } return o; }

/* one speculation lane: a stream, the two one-sided DPs of one anchor, their buffers */
struct gx_lane {
    cudaStream_t stream; cudaEvent_t evA, evB; double launchedAt;
    dp_job* h_jobs;                      /* pinned, 2 entries */
    dp_job* d_jobs;
    u32* dbg[2];
    u8* tb[2]; u32 tbBytes; u32* tbRow[2]; u32 tbRowCap[2]; u32* ops[2]; u32 opsCap[2]; int* act[2]; u32 actCap[2];
    /* state */
    bool busy; u64 anchor; size_t snapshot; segref left1, right1; u32 ring;
    int mode;                            /* 0 four-warp register kernel (1024-column window), 1 one-warp kernel (512); 2/3 shared-memory kernel, ring 4096/8192 */
    u64 estLo, estHi;                    /* seq1 rows this extension is expected to examine */
};

struct gx_cache {                        /* lives in the context: lanes are expensive to allocate */
    std::vector<gx_lane> lanes; u32 tbBytes;
    dseg* d_segs; size_t segsCap, segsUploaded;
    /* uploads go through mapped pinned staging + a copy KERNEL: an H2D copy on a stream that shares a
     * hardware queue with a running DP kernel waits for that kernel (engine switch inside one channel) */
    u32* h_stage; u32* d_stage; size_t stageWords; cudaStream_t upStream;
    std::vector<char*> epochChunks; size_t epochUsed;        /* bump pool for the per-epoch alignment tables */
};
#define GX_STAGE_WORDS (16u << 20)                          /* 64 MB */
#define GX_EPOCH_CHUNK ((size_t)16 << 20)

__global__ void k_upload_words(u32* __restrict__ dst, const u32* __restrict__ src, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

/* host words -> device through the staging buffer; returns with the data in place */
static int gx_upload(lzb_ctx* c, gx_cache* gc, void* dst, const void* src, size_t bytes) {
    const u32* w = (const u32*)src; u32* d = (u32*)dst; size_t n = (bytes + 3) / 4;
    while (n) {
        size_t k = n < gc->stageWords ? n : gc->stageWords;
        memcpy(gc->h_stage, w, k * 4);
        int blocks = (int)((k + 255) / 256); if (blocks > 64) blocks = 64;
        k_upload_words<<<blocks, 256, 0, gc->upStream>>>(d, gc->d_stage, k);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(gc->upStream));
        w += k; d += k; n -= k;
    }
    return 0;
}

static void free_lane(gx_lane& ln) {
    cudaStreamDestroy(ln.stream); cudaEventDestroy(ln.evA); cudaEventDestroy(ln.evB);
    cudaFreeHost(ln.h_jobs);
    for (int s = 0; s < 2; s++) { cudaFree(ln.dbg[s]); cudaFree(ln.tb[s]); cudaFree(ln.tbRow[s]); cudaFree(ln.ops[s]); cudaFree(ln.act[s]); }
}

void lzb_gapped_cache_free(lzb_ctx* c) {
    gx_cache* gc = (gx_cache*)c->gappedCache;
    if (!gc) return;
    for (auto& ln : gc->lanes) free_lane(ln);
    cudaFree(gc->d_segs); cudaFreeHost(gc->h_stage);
    if (gc->upStream != c->stream) cudaStreamDestroy(gc->upStream);
    for (char* ch : gc->epochChunks) cudaFree(ch);
    delete gc; c->gappedCache = NULL;
}

static int make_lane(gx_lane& ln, u32 tbBytes, u32 tbLen) {
    memset(&ln, 0, sizeof ln);
    CUDA_TRY(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&ln.evA)); CUDA_TRY(cudaEventCreate(&ln.evB));
    /* job descriptors live in mapped pinned memory: the kernel reads its job and writes its result
     * there, so a lane's stream carries nothing but kernels (no copy that another stream sharing the
     * hardware queue could get stuck behind) */
    CUDA_TRY(cudaHostAlloc(&ln.h_jobs, 2 * sizeof(dp_job), cudaHostAllocMapped));
    CUDA_TRY(cudaHostGetDevicePointer((void**)&ln.d_jobs, ln.h_jobs, 0));
    ln.tbBytes = tbBytes;
    for (int s = 0; s < 2; s++) {
        CUDA_TRY(cudaMalloc(&ln.tb[s], (size_t)tbBytes + 64));
        ln.tbRowCap[s] = tbLen / 24 + 4096; CUDA_TRY(cudaMalloc(&ln.tbRow[s], (size_t)ln.tbRowCap[s] * 4));
        ln.opsCap[s] = 1u << 20; CUDA_TRY(cudaMalloc(&ln.ops[s], (size_t)ln.opsCap[s] * 4));
        ln.actCap[s] = 256; CUDA_TRY(cudaMalloc(&ln.act[s], (size_t)ln.actCap[s] * 5 * 4));
    }
    return 0;
}

struct spec_result {                    /* a finished (possibly speculative) two-sided extension */
    bool have; size_t snapshot;         /* committed.size() when it was launched */
    segref left1, right1;               /* anchor neighbours it was computed with */
    dp_result L, R;
};

extern "C" int lzb_gapped_extend(lzb_ctx* c, lzb_target* t, lzb_query* q, const uint8_t* h1, const uint8_t* h2,
                                 lzb_segment* anchors, uint64_t n, const lzb_gapped_params* P,
                                 lzb_alignel** list, lzb_gapped_stats* stats) {
    cudaSetDevice(c->device);
    if (!c->haveScoring) return lzb_fail("lzb_set_scoring has not been called");
    if (P->tracebackBytes < 8) return lzb_fail("in new_traceback(), size can't be %u", P->tracebackBytes);
    if (c->sc.gapOpen < 0) return lzb_fail("lastz_b200's Y-drop kernel requires a non-negative gap open penalty (got %d)", c->sc.gapOpen);
    /* [multi] sequences (NUL-separated partitions, sequences.h:188-191): a sweep ends at the NULs around its anchor
     * (gapped_extend.c:1357-1372); the limits go into the job's M and N.  Alignments of different partitions cannot
     * meet, so one pass over all anchors gives what the reference's per-partition batches give (gapped_extend.c:1058).
     * The trivial self-alignment of identical PARTITIONS (:1185-1290) is not built: callers keep such pairs away. */
    std::vector<u32> tSeparators, qSeparators;                   /* positions of the NULs inside each sequence, ascending */
    for (const u8* z = (const u8*)memchr(t->h_seq, 0, t->len); z; z = (const u8*)memchr(z + 1, 0, t->len - (size_t)(z + 1 - t->h_seq)))
        tSeparators.push_back((u32)(z - t->h_seq));
    for (const u8* z = (const u8*)memchr(q->h_seq, 0, q->len); z; z = (const u8*)memchr(z + 1, 0, q->len - (size_t)(z + 1 - q->h_seq)))
        qSeparators.push_back((u32)(z - q->h_seq));
    auto wall0 = std::chrono::steady_clock::now();
    u64 launches0 = c->launches;
    *list = NULL;
    gx G; G.c = c; G.t = t; G.q = q; G.P = P; G.obi = G.oed = -1;
    memset(&G.st, 0, sizeof G.st); G.st.anchors = n;
    const u32 tbLen = 1 + (P->tracebackBytes - 8);                    /* new_traceback :2272-2290 */
    const u32 len1 = t->len, len2 = q->len;

    /* qSegmentsByDecreasingScore segment.c:1748-1771 (a total order, so any sort gives the same result) */
    std::sort(anchors, anchors + n, [](const lzb_segment& a, const lzb_segment& b) {
        if (a.s != b.s) return a.s > b.s;
        if (a.length != b.length) return a.length < b.length;
        if (a.pos2 != b.pos2) return a.pos2 < b.pos2;
        if (a.pos1 != b.pos1) return a.pos1 < b.pos1;
        return a.id < b.id;
    });
    G.al.resize(n + 1);
    for (u64 i = 0; i <= n; i++) { galn& m = G.al[i]; m.align = NULL; m.next = m.prev = -1; m.devIx = -1; m.left1 = m.right1 = m.left2 = m.right2 = NOSEG; m.pos1 = m.pos2 = m.end1 = m.end2 = 0; m.hspId = 0; }
    for (u64 i = 0; i < n; i++) { G.al[i].pos1 = anchors[i].pos1; G.al[i].pos2 = anchors[i].pos2; G.al[i].hspId = anchors[i].hspId; }

    /* identical_sequences gapped_extend.c:1886-1930 -> trivial self alignment :1113-1151 */
    if (P->identityCheck && len1 == len2) {
        bool same = true; s32 s = 0; const s32* sub = c->hostSub;
        for (u32 i = 0; i < len1 && same; i++) {
            u8 a = t->h_seq[i], b = q->h_seq[i];
            if (a >= 'a' && a <= 'z') a -= 32;
            if (b >= 'a' && b <= 'z') b -= 32;
            if (a != b) { same = false; break; }
            s32 v = sub[(u32)a * 256 + b];
            if (s == 0x7FFFFFFF) ; else if (v <= 0 || s < 0x7FFFFFFF - v) s += v; else s = 0x7FFFFFFF;
        }
        if (same) {
            galn& m = G.al[n];
            m.pos1 = m.pos2 = 0; m.end1 = m.end2 = len1 - 1;
            add_diag(m, 0, 0, m.end1, m.end2);
            list_insert(G, (int)n);
            m.devIx = (int)G.committed.size(); G.committed.push_back((int)n);
            lzb_alignel* a = (lzb_alignel*)calloc(1, sizeof *a);
            a->script = es_new(4); es_add(&a->script, LZB_OP_SUB, len1);
            a->beg1 = a->beg2 = 1; a->end1 = a->end2 = len1; a->seq1 = h1; a->seq2 = h2;
            a->s = s < P->scoreThreshold ? P->scoreThreshold : s; a->isTrivial = 1; m.align = a;
        }
    }
    /* identical_partition_of_sequence gapped_extend.c:2034-2120 -> :1185-1230: an unpartitioned query that equals
     * one partition of a [multi] target (the first such partition) gets the trivial alignment of that partition */
    if (P->identityCheck && !G.al[n].align && !tSeparators.empty() && qSeparators.empty()) {
        for (size_t k = 0; k < tSeparators.size(); k++) {        /* partition k lies between separator k and the next one (or the end) */
            const u32 before = tSeparators[k], after = k + 1 < tSeparators.size() ? tSeparators[k + 1] : len1;
            if (after - (before + 1) != len2) continue;
            bool same = true; s32 s = 0; const s32* sub = c->hostSub;
            for (u32 i = 0; i < len2; i++) {
                u8 a = t->h_seq[before + 1 + i], b = q->h_seq[i];
                if (a >= 'a' && a <= 'z') a -= 32;
                if (b >= 'a' && b <= 'z') b -= 32;
                if (a != b) { same = false; break; }
                s32 v = sub[(u32)a * 256 + b];
                if (s == 0x7FFFFFFF) ; else if (v <= 0 || s < 0x7FFFFFFF - v) s += v; else s = 0x7FFFFFFF;
            }
            if (!same) continue;
            galn& m = G.al[n];
            m.pos1 = before + 1; m.pos2 = 0; m.end1 = after - 1; m.end2 = len2 - 1;
            add_diag(m, m.pos1, m.pos2, m.end1, m.end2);
            list_insert(G, (int)n);
            m.devIx = (int)G.committed.size(); G.committed.push_back((int)n);
            lzb_alignel* a = (lzb_alignel*)calloc(1, sizeof *a);
            a->script = es_new(4); es_add(&a->script, LZB_OP_SUB, len2);
            a->beg1 = before + 2; a->beg2 = 1; a->end1 = after; a->end2 = len2; a->seq1 = h1; a->seq2 = h2;
            a->s = s < P->scoreThreshold ? P->scoreThreshold : s; a->isTrivial = 1; m.align = a;
            break;
        }
    }

    /* ---- speculation lanes (cached in the context across calls) ---- */
    const bool strict = P->speculation < 0;                    /* internal: exact-order rerun after a scheduling violation */
    int W = abs(P->speculation); if (W < 1) W = 1; if (W > 192) W = 192;
    const char* wenv = getenv("LZB_SPECULATION"); if (wenv) { W = atoi(wenv); if (W < 1) W = 1; if (W > 192) W = 192; }
    if ((u64)W > n) W = n ? (int)n : 1;
    u32 ring0 = 4096;                                       /* sweep-row ring, columns */
    const char* cenv = getenv("LZB_RING"); if (cenv) ring0 = (u32)atoi(cenv);
    gx_cache* gc = (gx_cache*)c->gappedCache;
    if (gc && gc->tbBytes != P->tracebackBytes) { lzb_gapped_cache_free(c); gc = NULL; }
    if (!gc) {
        gc = new gx_cache(); gc->tbBytes = P->tracebackBytes; gc->d_segs = NULL; gc->segsCap = 0; gc->segsUploaded = 0;
        gc->h_stage = gc->d_stage = NULL; gc->stageWords = GX_STAGE_WORDS; gc->epochUsed = 0;
        c->gappedCache = gc;
        gc->upStream = c->stream;
        CUDA_TRY(cudaHostAlloc(&gc->h_stage, (size_t)gc->stageWords * 4, cudaHostAllocMapped));
        CUDA_TRY(cudaHostGetDevicePointer((void**)&gc->d_stage, gc->h_stage, 0));
        gc->segsCap = 16u << 20;                            /* 320 MB up front: regrowing has to drain every lane */
        CUDA_TRY(cudaMalloc(&gc->d_segs, gc->segsCap * sizeof(dseg)));
    }
    gc->epochUsed = 0;
    while ((int)gc->lanes.size() < W) {
        gx_lane ln;
        if (make_lane(ln, P->tracebackBytes, tbLen)) return -1;
        gc->lanes.push_back(ln);
    }
    for (auto& ln : gc->lanes) { ln.busy = false; }
    gc->segsUploaded = 0;
    CUDA_TRY(cudaFuncSetAttribute(k_ydrop<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(k_ydrop<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    int dpThreads = 256;
    /* start with the register-resident kernel (1024-column window), fall back to the shared-memory kernel (4096, 8192);
     * LZB_DP_MODE forces the starting mode */
    int firstMode = 0;
    { const char* e = getenv("LZB_DP_MODE"); if (e) { int mdv = atoi(e); if (mdv >= 0 && mdv <= 3) firstMode = mdv; } }
    if (cenv) firstMode = 2;                                   /* an explicit ring size means the shared-memory kernel */
    { const char* e = getenv("LZB_DP_THREADS"); if (e && atoi(e) == 128) dpThreads = 128; }

    const bool trace = getenv("LZB_GAP_TRACE") != NULL;
    const char* dbgPath = getenv("LZB_DP_DEBUG");           /* file that receives every finished DP's per-row record */
    const u32 DBG_ROWS = 1u << 20;
    const bool prof = getenv("LZB_GAP_PROFILE") != NULL;     /* host-side breakdown of the scheduler on stderr */
    double pfSweep = 0, pfWait = 0, pfHarvest = 0, pfPush = 0, pfLaneBusy = 0, pfNbr = 0; u64 pfSweeps = 0, pfExamined = 0, pfNbrCalls = 0;
    u64 headAnchor = 0;
    auto now = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count(); };
    /* an event recorded behind a DP kernel blocks whatever else shares that hardware queue until the kernel ends;
     * LZB_LANE_EVENTS=0 times the DP launches by the host clock instead */
    const bool laneEvents = !(getenv("LZB_LANE_EVENTS") && !atoi(getenv("LZB_LANE_EVENTS")));
    std::vector<spec_result> spec(n);
    for (auto& s : spec) s.have = false;
    std::vector<char> inflight(n, 0);
    const u64 reach = tbLen / 300 + 1000;                    /* rows a DP is expected to cover, before any has finished */
    bool tablesDirtyInit = true; (void)tablesDirtyInit;

    /* append newly committed alignments to the device segment table (append-only, so running
     * kernels are undisturbed); the alignment table itself is snapshotted per launch */
    dalign* curAligns = NULL; std::vector<dalign*> alignEpochs;
    auto push_segments = [&]() -> int {
        size_t haveA = G.haligns.size();
        for (size_t k = haveA; k < G.committed.size(); k++) {
            galn& m = G.al[G.committed[k]];
            dalign d; memset(&d, 0, sizeof d);
            d.segBegin = (int)G.hsegs.size(); d.segCount = (int)m.segs.size();
            for (auto& s : m.segs) { dseg x = { s.b1, s.b2, s.e1, s.e2, s.type }; G.hsegs.push_back(x); }
            G.haligns.push_back(d);
        }
        if (G.hsegs.size() > gc->segsCap) {
            /* running kernels hold the old pointer: drain them first */
            for (auto& ln : gc->lanes) if (ln.busy) CUDA_TRY(cudaStreamSynchronize(ln.stream));
            cudaFree(gc->d_segs); gc->segsCap = G.hsegs.size() * 2 + (16u << 20);
            CUDA_TRY(cudaMalloc(&gc->d_segs, gc->segsCap * sizeof(dseg)));
            gc->segsUploaded = 0;
        }
        if (G.hsegs.size() > gc->segsUploaded) {
            if (gx_upload(c, gc, gc->d_segs + gc->segsUploaded, G.hsegs.data() + gc->segsUploaded,
                          (G.hsegs.size() - gc->segsUploaded) * sizeof(dseg))) return -1;
            gc->segsUploaded = G.hsegs.size();
        }
        for (size_t k = 0; k < G.committed.size(); k++) {
            galn& m = G.al[G.committed[k]]; dalign& d = G.haligns[k];
            d.pos1 = m.pos1; d.end1 = m.end1;
            d.left1 = dev_ref(G, m.left1); d.right1 = dev_ref(G, m.right1); d.left2 = dev_ref(G, m.left2); d.right2 = dev_ref(G, m.right2);
            d.next = m.next >= 0 ? G.al[m.next].devIx : -1; d.prev = m.prev >= 0 ? G.al[m.prev].devIx : -1;
        }
        /* a fresh, immutable copy of the alignment table per commit epoch: kernels already running
         * keep reading the copy they were launched with; the copies are tiny and freed at the end */
        curAligns = NULL;
        if (!G.haligns.empty()) {
            const size_t need = ((G.haligns.size() * sizeof(dalign) + 255) / 256) * 256;
            if (need > GX_EPOCH_CHUNK) {                       /* enormous table: its own allocation, freed at the end of the call */
                CUDA_TRY(cudaMalloc(&curAligns, need)); alignEpochs.push_back(curAligns);
            } else {
                const size_t chunk = gc->epochUsed / GX_EPOCH_CHUNK, off = gc->epochUsed % GX_EPOCH_CHUNK;
                size_t at = gc->epochUsed;
                if (off + need > GX_EPOCH_CHUNK) at = (chunk + 1) * GX_EPOCH_CHUNK;   /* does not fit the rest of this chunk */
                while (gc->epochChunks.size() <= at / GX_EPOCH_CHUNK) { char* ch = NULL; CUDA_TRY(cudaMalloc(&ch, GX_EPOCH_CHUNK)); gc->epochChunks.push_back(ch); }
                curAligns = (dalign*)(gc->epochChunks[at / GX_EPOCH_CHUNK] + at % GX_EPOCH_CHUNK);
                gc->epochUsed = at + need;
            }
            if (gx_upload(c, gc, curAligns, G.haligns.data(), G.haligns.size() * sizeof(dalign))) return -1;
        }
        return 0;
    };
    bool tablesDirty = true;

    auto launch = [&](gx_lane& ln, int onlySide) -> int {
        galn& m = G.al[ln.anchor];
        const segref mLeft = ln.left1, mRight = ln.right1;   /* the neighbours this anchor was started with */
        if (tablesDirty) { const double p0 = prof ? now() : 0; if (push_segments()) return -1; tablesDirty = false; if (prof) pfPush += now() - p0; }
        /* get_above_below :4043-4060 */
        int below = G.oed; while (below >= 0 && !(G.al[below].end1 < m.pos1)) below = G.al[below].prev;
        int above = G.obi; while (above >= 0 && !(G.al[above].pos1 > m.pos1)) above = G.al[above].next;
        for (int side = 0; side < 2; side++) {
            dp_job& J = ln.h_jobs[side];
            memset(&J, 0, sizeof J);
            if (onlySide >= 0 && side != onlySide) { J.skip = 1; continue; }     /* kernel returns at once */
            int rev = side == 0;
            J.reversed = rev; J.a1 = m.pos1; J.a2 = m.pos2;
            u32 low1 = 0, high1 = len1, low2 = 0, high2 = len2;  /* the anchor's partition in each sequence: first base, one past the last */
            if (!tSeparators.empty()) {
                auto after = std::upper_bound(tSeparators.begin(), tSeparators.end(), m.pos1);
                if (after != tSeparators.end()) high1 = *after;
                if (after != tSeparators.begin()) low1 = *(after - 1) + 1;
            }
            if (!qSeparators.empty()) {
                auto after = std::upper_bound(qSeparators.begin(), qSeparators.end(), m.pos2);
                if (after != qSeparators.end()) high2 = *after;
                if (after != qSeparators.begin()) low2 = *(after - 1) + 1;
            }
            J.M = rev ? m.pos1 + 1 - low1 : high1 - (m.pos1 + 1); J.N = rev ? m.pos2 + 1 - low2 : high2 - (m.pos2 + 1);
            /* initial L/R, gapped_extend.c:3500-3543 */
            s32 L = 0, R = (s32)(J.N + 1);
            if (mLeft.al >= 0) { hseg& s = G.al[mLeft.al].segs[mLeft.sg]; L = (s32)(s.b2 - m.pos2); if (s.type == SEG_DIAG) L -= (s32)(s.b1 - m.pos1); }
            if (mRight.al >= 0) { hseg& s = G.al[mRight.al].segs[mRight.sg]; R = (s32)(s.b2 - m.pos2); if (s.type == SEG_DIAG) R -= (s32)(s.b1 - m.pos1); }
            if (rev) {
                if (mLeft.al < 0 && mRight.al >= 0) { L = -R + 1; R = (s32)(J.N + 1); }
                else if (mLeft.al >= 0 && mRight.al < 0) { R = -L - 1; L = 0; }
                else if (mLeft.al >= 0 && mRight.al >= 0) { s32 tt = -L - 1; L = -R + 1; R = tt; }
            }
            J.L0 = L; J.R0 = R;
            J.leftSeg = dev_ref(G, mLeft); J.rightSeg = dev_ref(G, mRight);
            int lst = rev ? below : above;
            J.alignList = lst >= 0 ? G.al[lst].devIx : -1;
            J.al = curAligns;
            J.tb = ln.tb[side]; J.tbLen = tbLen; J.tbRow = ln.tbRow[side]; J.tbRowCap = ln.tbRowCap[side];
            J.ops = ln.ops[side]; J.opsCap = ln.opsCap[side]; J.act = ln.act[side]; J.actCap = ln.actCap[side];
            if (dbgPath) {
                if (!ln.dbg[side]) CUDA_TRY(cudaMalloc(&ln.dbg[side], (size_t)DBG_ROWS * 16));
                J.dbg = ln.dbg[side]; J.dbgCap = DBG_ROWS;
            }
        }
        if (laneEvents) CUDA_TRY(cudaEventRecord(ln.evA, ln.stream)); else ln.launchedAt = now();
        /* other shapes of the register kernel were measured and lost on the 50 Mbp pair (DP-kernel wait 1.81 s):
         * 8 warps x 4 columns 2.13 s, 4 warps x 6 columns with escalation 3.05 s (gpurun round 12) */
        if (ln.mode == 0)
            k_ydrop_mw<8, 4><<<2, 128, 0, ln.stream>>>(ln.d_jobs, gc->d_segs, t->d_cls, q->d_cls, len1, len2, c->d_sc, P->yDrop, P->trimToPeak);
        else if (ln.mode == 1)
            k_ydrop_warp<16><<<2, 32, 0, ln.stream>>>(ln.d_jobs, gc->d_segs, t->d_cls, q->d_cls, len1, len2, c->d_sc, P->yDrop, P->trimToPeak);
        else {
            ln.ring = ln.mode == 2 ? ring0 : ring0 * 2;
            size_t smem = (size_t)ln.ring * 17 + LZB_MAX_CLASSES * LZB_MAX_CLASSES * 4 + 1024;
            if (dpThreads == 128)
                k_ydrop<128><<<2, 128, smem, ln.stream>>>(ln.d_jobs, gc->d_segs, t->d_cls, q->d_cls, len1, len2,
                                                         c->d_sc, P->yDrop, P->trimToPeak, ln.ring);
            else
                k_ydrop<256><<<2, 256, smem, ln.stream>>>(ln.d_jobs, gc->d_segs, t->d_cls, q->d_cls, len1, len2,
                                                         c->d_sc, P->yDrop, P->trimToPeak, ln.ring);
        }
        c->launches++;
        CUDA_TRY(cudaGetLastError());
        if (laneEvents) CUDA_TRY(cudaEventRecord(ln.evB, ln.stream));
        return 0;
    };

    /* rows an extension from anchor y will probably examine: +-reach, cut short by committed
     * alignments that end/start on a nearby diagonal (its DP stops at their masked cells).  Only a
     * scheduling hint: correctness rests on the validation at commit time. */
    auto est_region = [&](galn& y, u64 rch) -> std::pair<u64, u64> {
        s64 dy = (s64)y.pos1 - (s64)y.pos2;
        u64 lo = y.pos1 > rch ? y.pos1 - rch : 0, hi = (u64)y.pos1 + rch;
        for (int ci : G.committed) {
            galn& x = G.al[ci];
            s64 dEnd = (s64)x.end1 - (s64)x.end2, dBeg = (s64)x.pos1 - (s64)x.pos2;
            if (x.end1 < y.pos1 && llabs(dEnd - dy) < 3000) { u64 b = x.end1 > 600 ? x.end1 - 600 : 0; if (b > lo) lo = b; }
            if (x.pos1 > y.pos1 && llabs(dBeg - dy) < 3000) { u64 b = (u64)x.pos1 + 600; if (b < hi) hi = b; }
        }
        return std::make_pair(lo, hi);
    };

    auto start_anchor = [&](gx_lane& ln, u64 ai) -> int {
        galn& m = G.al[ai];
        ln.estLo = 0; ln.estHi = ~0ull;                          /* the caller fills in its estimate */
        ln.busy = true; ln.anchor = ai; ln.snapshot = G.committed.size(); ln.left1 = m.left1; ln.right1 = m.right1; ln.ring = ring0; ln.mode = firstMode;
        inflight[ai] = 1;
        if (trace) fprintf(stderr, "[gx %.4f] launch a=%llu pos1=%u est=[%llu,%llu] head=%llu committed=%zu\n", now(), (unsigned long long)ai, m.pos1, (unsigned long long)ln.estLo, (unsigned long long)ln.estHi, (unsigned long long)headAnchor, G.committed.size());
        return launch(ln, -1);
    };

    /* a lane's stream has drained: collect the result, or rerun a side that outgrew a buffer */
    auto harvest = [&](gx_lane& ln) -> int {
        float ms = 0;
        if (laneEvents) cudaEventElapsedTime(&ms, ln.evA, ln.evB); else ms = (float)((now() - ln.launchedAt) * 1e3);   /* host clock: launch to harvest */
        G.st.kernelSeconds[0] += ms / 1e3;
        spec_result& sr = spec[ln.anchor];
        int redo = -2;                                       /* -2 none, -1 both, 0/1 one side */
        for (int side = 0; side < 2; side++) {
            dp_job& J = ln.h_jobs[side];
            if (J.skip) continue;                            /* side finished in an earlier pass */
            bool again = false;
            if (J.status == DP_RING) {
                if (ln.mode >= 3) return lzb_fail("Y-drop band wider than %u columns; lower --ydrop", ln.ring);
                again = true;
            } else if (J.status == DP_TBROW) {
                cudaFree(ln.tbRow[side]); ln.tbRowCap[side] = ln.tbRowCap[side] * 4 < tbLen ? ln.tbRowCap[side] * 4 : tbLen + 8;
                CUDA_TRY(cudaMalloc(&ln.tbRow[side], (size_t)ln.tbRowCap[side] * 4)); again = true;
            } else if (J.status == DP_OPS) {
                cudaFree(ln.ops[side]); ln.opsCap[side] *= 4; CUDA_TRY(cudaMalloc(&ln.ops[side], (size_t)ln.opsCap[side] * 4)); again = true;
            } else if (J.status == DP_ACT) {
                cudaFree(ln.act[side]); ln.actCap[side] *= 4; CUDA_TRY(cudaMalloc(&ln.act[side], (size_t)ln.actCap[side] * 5 * 4)); again = true;
            }
            if (again) { redo = (redo == -2) ? side : -1; continue; }
            dp_result& r = side ? sr.R : sr.L;
            r.score = J.score; r.end1 = J.end1; r.end2 = J.end2; r.rows = J.rows; r.status = J.status; r.cells = J.cells;
            r.ops.resize(J.nops);
            if (J.nops) CUDA_TRY(cudaMemcpyAsync(r.ops.data(), ln.ops[side], (size_t)J.nops * 4, cudaMemcpyDeviceToHost, ln.stream));
            G.st.dpCellsComputed += J.cells;
            if (dbgPath && J.dbg) {
                u32 nr = J.rows < DBG_ROWS ? J.rows : DBG_ROWS;
                std::vector<u32> rows((size_t)nr * 4);
                CUDA_TRY(cudaMemcpy(rows.data(), J.dbg, (size_t)nr * 16, cudaMemcpyDeviceToHost));
                FILE* df = fopen(dbgPath, "ab");
                if (df) {
                    u32 hdr[8] = { 0x44504447u, (u32)ln.anchor, (u32)side, J.rows, (u32)J.status, (u32)J.cells, (u32)ln.mode, nr };
                    fwrite(hdr, 4, 8, df); fwrite(rows.data(), 16, nr, df); fclose(df);
                }
            }
        }
        CUDA_TRY(cudaStreamSynchronize(ln.stream));
        if (redo != -2) {
            if (trace) fprintf(stderr, "[gx %.4f] rerun a=%llu side=%d statusL=%d statusR=%d rowsL=%u rowsR=%u kernel_ms=%.1f mode=%d\n", now(), (unsigned long long)ln.anchor, redo,
                               ln.h_jobs[0].status, ln.h_jobs[1].status, ln.h_jobs[0].rows, ln.h_jobs[1].rows, ms, ln.mode);
            bool ringGrow = false;
            for (int side = 0; side < 2; side++) if (ln.h_jobs[side].status == DP_RING) ringGrow = true;
            if (ringGrow) ln.mode = ln.mode < 2 ? 2 : ln.mode + 1;
            /* rerun with the neighbours it was started with; the (possibly newer) alignment table is a
             * superset, and validation still uses the ORIGINAL snapshot, so any difference is caught */
            return launch(ln, redo);
        }
        sr.have = true; sr.snapshot = ln.snapshot; sr.left1 = ln.left1; sr.right1 = ln.right1;
        if (trace) fprintf(stderr, "[gx %.4f] done a=%llu rowsL=%u rowsR=%u kernel_ms=%.1f mode=%d\n", now(), (unsigned long long)ln.anchor, sr.L.rows, sr.R.rows, ms, ln.mode);
        ln.busy = false; inflight[ln.anchor] = 0;
        return 0;
    };

    /* ---- the anchor loop, gapped_extend.c:1300-1470 ----
     * Sequential semantics: anchors are extended best score first and every kept alignment
     * constrains the later ones.  Two alignments whose DPs examined disjoint seq-1 row ranges cannot
     * see each other (bounds, masks and the skip rule are all row-local), so they may be committed in
     * either order.  The loop therefore sweeps the not-yet-final anchors in score order carrying the
     * row ranges of everything EARLIER that is still unresolved (in flight, waiting, or finished but
     * itself waiting): a finished extension commits as soon as its exact rows clear all of them; an
     * anchor may start as soon as its estimated rows do.  Estimates only schedule.  At every commit
     * the exact ranges are checked against alignments that were committed ahead of their turn; if
     * an estimate was too small and two such ranges do overlap, the whole call is redone in strict
     * order (`strict`), so the result always equals the sequential algorithm's. */
    std::vector<u8> fin(n, 0);                               /* 1 = skipped / committed / dropped */
    /* the anchor POINTS (commit overwrites al[i].pos1/pos2 with the alignment's start) and the anchors by
     * seq-1 position: when an alignment is committed the anchors lying on it are retired there and then
     * (msp_left_right's d == 0 test, gapped_extend.c:4008), instead of re-testing thousands of anchors
     * against every alignment on every sweep -- that was 1.2-2.2 s of host time per 50 Mbp strand */
    std::vector<u32> apos1(n), apos2(n);
    for (u64 i = 0; i < n; i++) { apos1[i] = G.al[i].pos1; apos2[i] = G.al[i].pos2; }
    std::vector<u32> byPos(n);
    for (u64 i = 0; i < n; i++) byPos[i] = (u32)i;
    std::sort(byPos.begin(), byPos.end(), [&](u32 a, u32 b) { return apos1[a] != apos1[b] ? apos1[a] < apos1[b] : a < b; });
    std::vector<int> blockedBy(n, -1);                      /* the running earlier extension a waiting anchor clashed with ... */
    std::vector<u32> blockedAt(n, 0);                       /* ... and the number of commits at that time */
    std::vector<u64> dpLo(n + 1, 0), dpHi(n + 1, 0);        /* rows examined by a committed anchor's DPs */
    u64 hd = 0; bool violation = false;
    u64 maxRows = 0;
    /* retire the anchors that lie on alignment `ai` (index into G.al; n = the trivial self alignment) */
    auto retire_covered = [&](int ai) {
        galn& x = G.al[ai];
        const int ns = (int)x.segs.size();
        if (ns == 0) return;
        size_t lo = std::lower_bound(byPos.begin(), byPos.end(), x.pos1, [&](u32 a, u32 v) { return apos1[a] < v; }) - byPos.begin();
        for (size_t z = lo; z < byPos.size() && apos1[byPos[z]] <= x.end1; z++) {
            const u32 j = byPos[z];
            if (fin[j] || (int)j == ai) continue;
            const u32 pos1 = apos1[j], pos2 = apos2[j];
            int k = 0, hi2 = ns;                             /* first segment whose e1 >= pos1 */
            while (k < hi2) { int mid = (k + hi2) >> 1; if (x.segs[mid].e1 < pos1) k = mid + 1; else hi2 = mid; }
            if (k == ns) continue;
            const hseg& bp = x.segs[k]; s32 d;
            if (bp.type == SEG_DIAG) d = (s32)(bp.b2 - pos2) + (s32)(pos1 - bp.b1); else d = (s32)(bp.b2 - pos2);
            if (d != 0) continue;
            /* on an alignment committed EARLIER in the order: skipped for good (:1335); on one committed
             * ahead of its turn: the estimates were wrong */
            if (ai != (int)n && (u64)ai > j) { violation = true; return; }
            fin[j] = 1; spec[j].have = false; spec[j].L.ops.clear(); spec[j].R.ops.clear();
        }
    };
    if (G.obi == (int)n) retire_covered((int)n);

    /* commit anchor i from its finished, validated result */
    auto commit_anchor = [&](u64 i) {
        galn& m = G.al[i]; spec_result& sr = spec[i];
        G.st.anchorsExtended++;
        /* the reference's counters see exactly the DPs whose results are used (gapped_extend.c:3593,3776) */
        G.st.dpCells += sr.L.cells + sr.R.cells; G.st.dpRows += (u64)sr.L.rows + sr.R.rows;
        G.st.truncated += (sr.L.status == DP_TRUNCATED) + (sr.R.status == DP_TRUNCATED);
        if (trace) fprintf(stderr, "[gx %.4f] commit a=%llu pos1=%u hd=%llu\n", now(), (unsigned long long)i, m.pos1, (unsigned long long)hd);
        u32 a1 = m.pos1, a2 = m.pos2;
        dpLo[i] = (u64)a1 + 1 >= (u64)sr.L.rows + 2 ? (u64)a1 + 1 - sr.L.rows - 2 : 0; dpHi[i] = (u64)a1 + sr.R.rows + 2;
        u32 start1 = a1 + 1 - sr.L.end1, start2 = a2 + 1 - sr.L.end2, stop1 = a1 + sr.R.end1, stop2 = a2 + sr.R.end2;
        lzb_editscript* sl = es_new((u32)(sr.L.ops.size() + sr.R.ops.size() + 4));
        /* left script: ops in emission order; right script: emitted far-end first, so reversed (:2529-2551) */
        for (size_t k = 0; k < sr.L.ops.size(); k++) es_add(&sl, sr.L.ops[k] & 3, sr.L.ops[k] >> 2);
        for (size_t k = sr.R.ops.size(); k-- > 0;) es_add(&sl, sr.R.ops[k] & 3, sr.R.ops[k] >> 2);
        if (sl->len > 0 && sr.R.ops.size() > 0) sl->tailOp = sr.R.ops.back() & 3;
        s32 score = sr.L.score + sr.R.score;
        if (sl->len != 0) {
            if ((sl->op[0] & 3) != LZB_OP_SUB) {             /* lop_initial_indels :2589 */
                u32 p1 = start1, p2 = start2, k = 0;
                for (; k < sl->len; k++) { u32 op = sl->op[k] & 3, rpt = sl->op[k] >> 2; if (op == LZB_OP_SUB) break; if (op == LZB_OP_INS) p2 += rpt; else p1 += rpt; }
                if (k == sl->len) score = (s32)(-0x7FFFFFFF - 1);
                else { start1 = p1; start2 = p2; sl->len -= k; memmove(sl->op, sl->op + k, (size_t)sl->len * 4); score = rescore(G, start1, start2, sl); }
            }
            if (score != (s32)(-0x7FFFFFFF - 1) && (sl->op[sl->len - 1] & 3) != LZB_OP_SUB) {   /* lop_final_indels :2640 */
                u32 p1 = stop1, p2 = stop2, k = sl->len;
                while (k > 0) { k--; u32 op = sl->op[k] & 3, rpt = sl->op[k] >> 2; if (op == LZB_OP_SUB) { k++; break; } if (op == LZB_OP_INS) p2 -= rpt; else p1 -= rpt; }
                if (k == 0) score = (s32)(-0x7FFFFFFF - 1);
                else { stop1 = p1; stop2 = p2; sl->len = k; score = rescore(G, start1, start2, sl); }
            }
        }
        /* format_alignment :5153 */
        u32 beg1 = start1 + 1, end1 = stop1 + 1, beg2 = start2 + 1, end2 = stop2 + 1;
        u32 height = end1 - beg1 + 1, width = end2 - beg2 + 1, k = 0;
        m.segs.clear();
        for (u32 ii = 0, jj = 0; ii < height || jj < width;) {
            u32 si = ii, sj = jj, run = 0;
            while (k < sl->len && (sl->op[k] & 3) == LZB_OP_SUB) { run += sl->op[k] >> 2; k++; }
            ii += run; jj += run;
            add_diag(m, beg1 + si - 1, beg2 + sj - 1, beg1 + ii - 2, beg2 + jj - 2);
            if (ii < height || jj < width) {
                if (k < sl->len) { u32 op = sl->op[k] & 3, rpt = sl->op[k] >> 2; if (op == LZB_OP_INS) jj += rpt; else if (op == LZB_OP_DEL) ii += rpt; k++; }
                else break;
            }
        }
        lzb_alignel* a = (lzb_alignel*)calloc(1, sizeof *a);
        a->script = sl; a->beg1 = beg1; a->beg2 = beg2; a->end1 = end1; a->end2 = end2;
        a->seq1 = h1; a->seq2 = h2; a->s = score; a->hspId = m.hspId;
        m.align = a; m.pos1 = start1; m.pos2 = start2; m.end1 = stop1; m.end2 = stop2;
        sr.have = false; sr.L.ops.clear(); sr.L.ops.shrink_to_fit(); sr.R.ops.clear(); sr.R.ops.shrink_to_fit();
        fin[i] = 1;
        if (m.segs.empty()) return;
        if (!P->allBounds && a->s < P->scoreThreshold) { free(a->script); free(a); m.align = NULL; m.segs.clear(); return; }
        alignment_neighbours(G, m);
        list_insert(G, (int)i);
        m.devIx = (int)G.committed.size(); G.committed.push_back((int)i);
        tablesDirty = true;
        retire_covered((int)i);
    };

    /* an earlier, still unresolved anchor: [lo,hi] = rows its extension covers (estimated while in
     * flight, exact once finished) -- nothing that overlaps it may START; [clo,chi] = rows that it, or an
     * anchor waiting on it, could still come to cover -- nothing that overlaps it may COMMIT */
    struct pend { u64 lo, hi, clo, chi; int owner; };
    auto widen = [](u64 v, u64 by, bool down) -> u64 { return down ? (v > by ? v - by : 0) : v + by; };
    while (hd < n && !violation) {
        while (hd < n && fin[hd]) hd++;
        if (hd >= n) break;
        headAnchor = hd;
        /* ---- one sweep in score order over the not-yet-final anchors ---- */
        std::vector<pend> unresolved;                        /* earlier anchors whose outcome is still open */
        int freeLanes = 0;
        for (int z = 0; z < W; z++) if (!gc->lanes[z].busy) freeLanes++;
        const u64 est = std::max<u64>(reach, maxRows + maxRows / 4);
        bool progressed = false; u64 examined = 0; int starved = 0;
        const double sw0 = prof ? now() : 0; pfSweeps++;
        for (u64 j = hd; j < n && examined < 8192; j++) {
            if (fin[j]) continue;
            examined++; pfExamined++;
            galn& y = G.al[j];
            const s64 dy = (s64)y.pos1 - (s64)y.pos2;
            int coverer = -1;
            /* neighbours (msp_left_right) are needed only to validate a finished extension or to start one */
            auto fresh_neighbours = [&]() -> bool {
                const double n0 = prof ? now() : 0;
                const bool open = anchor_neighbours(G, y, &coverer);
                if (prof) { pfNbr += now() - n0; pfNbrCalls++; }
                if (!open) {
                    if (coverer >= 0 && (u64)coverer > j && coverer != (int)n) { violation = true; return false; }
                    fin[j] = 1; spec[j].have = false; spec[j].L.ops.clear(); spec[j].R.ops.clear(); progressed = true;
                    return false;
                }
                return true;
            };
            if (spec[j].have && !fresh_neighbours()) { if (violation) break; continue; }
            if (spec[j].have) {
                spec_result& sr = spec[j];
                const u64 lo = (u64)y.pos1 + 1 >= (u64)sr.L.rows + 2 ? (u64)y.pos1 + 1 - sr.L.rows - 2 : 0;
                const u64 hi = (u64)y.pos1 + sr.R.rows + 2;
                bool usable = true;
                /* nothing committed since its launch may touch the rows its DPs examined (:1335-1389 inputs) */
                for (size_t k = sr.snapshot; k < G.committed.size() && usable; k++) {
                    galn& x = G.al[G.committed[k]];
                    if (!((u64)x.end1 < lo || (u64)x.pos1 > hi)) usable = false;
                }
                if (usable && (sr.left1.al != y.left1.al || sr.left1.sg != y.left1.sg || sr.right1.al != y.right1.al || sr.right1.sg != y.right1.sg)) usable = false;
                if (!usable) {
                    G.st.redone++; sr.have = false;
                    if (trace) fprintf(stderr, "[gx %.4f] invalid a=%llu rows=[%llu,%llu]\n", now(), (unsigned long long)j, (unsigned long long)lo, (unsigned long long)hi);
                } else {
                    bool clear = true;
                    for (auto& u : unresolved) if (!(hi < u.clo || lo > u.chi)) { clear = false; break; }
                    if (strict && !unresolved.empty()) clear = false;
                    if (clear) {
                        /* alignments committed ahead of their turn must be out of reach of this one, and vice versa */
                        for (int ci : G.committed) if ((u64)ci > j && ci != (int)n && !(dpHi[ci] < lo || dpLo[ci] > hi)) { violation = true; break; }
                        if (violation) break;
                        commit_anchor(j); progressed = true;
                        if (violation) break;
                        continue;
                    }
                    unresolved.push_back(pend{ lo, hi, widen(lo, 2 * est, true), widen(hi, 2 * est, false), (int)j });
                    continue;
                }
            }
            if (inflight[j]) {
                gx_lane* ln = NULL;
                for (int z = 0; z < W; z++) if (gc->lanes[z].busy && gc->lanes[z].anchor == j) ln = &gc->lanes[z];
                const u64 lo = ln ? ln->estLo : 0, hi = ln ? ln->estHi : ~0ull;
                unresolved.push_back(pend{ lo, hi, widen(lo, 2 * est, true), widen(hi, 2 * est, false), (int)j });
                continue;
            }
            /* not started: may start if its estimated rows clear every earlier extension */
            if (blockedBy[j] >= 0 && inflight[blockedBy[j]] && !spec[blockedBy[j]].have && blockedAt[j] == (u32)G.committed.size()) {
                /* same running extension (its estimated rows are fixed at launch), same committed set
                 * (so this anchor's estimate cannot have shrunk): still clashing */
                continue;
            }
            std::pair<u64, u64> rg = est_region(y, est);
            bool clash = false;
            for (auto& u : unresolved) if (!(rg.second < u.lo || rg.first > u.hi)) { clash = true; blockedBy[j] = u.owner; blockedAt[j] = (u32)G.committed.size(); break; }
            if (strict && !unresolved.empty()) clash = true;
            if (clash) continue;                             /* waits on that extension; its commit window already covers this anchor */
            blockedBy[j] = -1;
            if (!fresh_neighbours()) { if (violation) break; continue; }
            if (freeLanes == 0) {
                /* could start but no lane is free: hold later commits off its rows, and stop looking */
                unresolved.push_back(pend{ 1, 0, rg.first, rg.second, (int)j });
                if (++starved >= 8) break;
                continue;
            }
            gx_lane* fl = NULL;
            for (int z = 0; z < W; z++) if (!gc->lanes[z].busy) { fl = &gc->lanes[z]; break; }
            if (start_anchor(*fl, j)) return -1;
            fl->estLo = rg.first; fl->estHi = rg.second;
            freeLanes--; if (j != hd) G.st.speculated++;
            unresolved.push_back(pend{ rg.first, rg.second, widen(rg.first, 2 * est, true), widen(rg.second, 2 * est, false), (int)j });
        }
        if (prof) pfSweep += now() - sw0;
        if (violation) break;
        if (progressed) continue;                            /* commits/skips may have unblocked more */
        /* ---- nothing more to decide: wait for a lane ---- */
        bool any = false; int nbusy = 0;
        for (int z = 0; z < W; z++) if (gc->lanes[z].busy) { any = true; nbusy++; }
        if (!any) return lzb_fail("internal error: gapped scheduler stalled at anchor %llu", (unsigned long long)hd);
        bool got = false;
        const double w0 = prof ? now() : 0; double hv = 0;
        while (!got) {
            for (int z = 0; z < W; z++) {
                gx_lane& ln = gc->lanes[z];
                if (!ln.busy) continue;
                cudaError_t e = cudaStreamQuery(ln.stream);
                if (e == cudaSuccess) {
                    const u64 a = ln.anchor; const double h0 = prof ? now() : 0;
                    if (harvest(ln)) return -1;
                    if (prof) hv += now() - h0;
                    if (!ln.busy) { got = true; maxRows = std::max<u64>(maxRows, std::max<u64>(spec[a].L.rows, spec[a].R.rows)); }
                }
                else if (e != cudaErrorNotReady) return lzb_fail("Y-drop kernel failed: %s", cudaGetErrorString(e));
            }
            if (!got) std::this_thread::sleep_for(std::chrono::microseconds(20));
        }
        if (prof) { const double dtw = now() - w0; pfWait += dtw - hv; pfHarvest += hv; pfLaneBusy += dtw * nbusy; }
    }
    if (prof)
        fprintf(stderr, "[gx profile] W=%d wall=%.3f sweeps=%llu examined=%llu sweep_s=%.3f (neighbour calls %llu, %.3f s) push_s=%.3f wait_s=%.3f harvest_s=%.3f "
                        "avg_busy_lanes_while_waiting=%.1f kernel_lane_s=%.3f extended=%llu redone=%llu rows=%llu committed=%zu\n",
                W, now(), (unsigned long long)pfSweeps, (unsigned long long)pfExamined, pfSweep, (unsigned long long)pfNbrCalls, pfNbr, pfPush, pfWait, pfHarvest,
                pfWait + pfHarvest > 0 ? pfLaneBusy / (pfWait + pfHarvest) : 0.0, G.st.kernelSeconds[0], (unsigned long long)G.st.anchorsExtended,
                (unsigned long long)G.st.redone, (unsigned long long)G.st.dpRows, G.committed.size());
    /* abandon speculative work that was never needed */
    for (auto& ln : gc->lanes) if (ln.busy) { cudaStreamSynchronize(ln.stream); ln.busy = false; }
    for (dalign* d : alignEpochs) cudaFree(d);
    alignEpochs.clear();
    if (violation) {
        /* an estimate was too small: redo everything in strict order (exact by construction) */
        for (int o = G.obi; o >= 0; o = G.al[o].next) { galn& m = G.al[o]; if (m.align) { free(m.align->script); free(m.align); } }
        if (trace) fprintf(stderr, "[gx %.4f] out-of-order commit violated an estimate; strict rerun\n", now());
        if (strict) return lzb_fail("internal error: ordering violation in strict mode");
        lzb_gapped_params P2 = *P; P2.speculation = -W;      /* negative => strict */
        lzb_gapped_stats st2;
        int rc = lzb_gapped_extend(c, t, q, h1, h2, anchors, n, &P2, list, &st2);
        if (stats) { *stats = st2; stats->redone += G.st.redone + 1000000; }
        return rc;
    }
    lzb_alignel* head = NULL, *last = NULL;
    for (int o = G.obi; o >= 0; o = G.al[o].next) {
        galn& m = G.al[o];
        bool drop = m.align->s < P->scoreThreshold || (P->inhibitTrivial && m.align->isTrivial);
        if (drop) { free(m.align->script); free(m.align); }
        else { if (!head) head = last = m.align; else { last->next = m.align; last = m.align; } }
    }
    *list = head;
    G.st.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
    G.st.launches = c->launches - launches0;
    if (stats) *stats = G.st;
    return 0;
}

extern "C" void lzb_free_align_list(lzb_alignel* a) {
    while (a) { lzb_alignel* nx = a->next; free(a->script); free(a); a = nx; }
}
