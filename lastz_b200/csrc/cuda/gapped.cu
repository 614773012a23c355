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
    if (n > c->peaksCap) {                                   /* scratch kept in the context: cudaMalloc/cudaFree per call synchronise the device */
        cudaFree(c->peaksBuf); c->peaksBuf = NULL; c->peaksCap = 0;
        const size_t cap = n + n / 2 + 1024;
        CUDA_TRY(cudaMalloc(&c->peaksBuf, cap * sizeof(lzb_segment))); c->peaksCap = cap;
    }
    lzb_segment* d = (lzb_segment*)c->peaksBuf;
    CUDA_TRY(cudaMemcpyAsync(d, anchors, n * sizeof(lzb_segment), cudaMemcpyHostToDevice, c->stream));
    int blocks = (int)((n + 127) / 128); if (blocks > c->smCount * 16) blocks = c->smCount * 16;
    k_peaks<<<blocks, 128, 0, c->stream>>>(d, n, t->d_cls, q->d_cls, c->d_sc);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(anchors, d, n * sizeof(lzb_segment), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * host side of gapped_extend: the scheduler (gapped_sched.hpp) over the CUDA backend below
 * ---------------------------------------------------------------------------------------- */
#include "gapped_sched.hpp"

#define GX_MAX_LANES (LZB_LAUNCH_MAX / 2)
#define GX_STREAMS 120                                       /* a launch goes to a stream with nothing in flight; the device runs at most 128 grids at once */
#define GX_STAGE_WORDS (4u << 20)                           /* 16 MB */
#define GX_CKPT_CAP 1024u

struct cuda_lane {
    u8* tb[2]; u32* tbRow[2]; u32 tbRowCap[2];
    u32* ops[2]; u32* d_ops[2]; u32 opsCap[2];              /* mapped pinned: the host reads a finished job's ops in place */
    int* act[2]; u32 actCap[2];
    u32* ckpt[2];
    int* list[2]; int* d_list[2]; size_t listCap[2];        /* mapped pinned */
    bool ownTbRow[2], ownOps[2], ownAct[2];                  /* regrown on its own (else part of the chunk's slab) */
};
#define GX_CHUNK 16                                          /* lanes made at a time: a few big allocations instead of hundreds */
struct lane_chunk { u8* tb; u32* tbRow; int* act; u32* ckpt; u32* ops; };

struct gx_cache {                        /* lives in the context: lanes are expensive to allocate */
    lzb_ctx* c;
    std::vector<cuda_lane> lanes; std::vector<lane_chunk> chunks; u32 tbBytes, tbLen, ckptEvery;
    dp_job* h_jobs; dp_job* d_jobs;      /* mapped pinned, 2 per lane: a kernel reads its job and writes its result there,
                                            so the launch streams carry nothing but kernels */
    dseg* d_segs; size_t segsCap; dalign* d_aligns; size_t alignsCap;
    /* uploads go through mapped pinned staging + a copy KERNEL: an H2D copy on a stream that shares a
     * hardware queue with a running DP kernel waits for that kernel (engine switch inside one channel) */
    u32* h_stage; u32* d_stage; cudaStream_t upStream;
    cudaStream_t streams[GX_STREAMS]; int nStreams;
    size_t warpSmem;                                         /* dynamic shared memory asked for per one-warp CTA (caps CTAs per SM) */
    std::vector<std::pair<u16, u32> > inflight[GX_STREAMS];   /* (job, token) of the last launch on each stream */
    const u8* cls1; const u8* cls2; u32 len1, len2; s32 yDrop; int trim;   /* of the current call */
    u32 launched[2 * GX_MAX_LANES];      /* token of the last launch of each job (0: never launched) */
    u64 polls; char err[256];
};

__global__ void k_upload_words(u32* __restrict__ dst, const u32* __restrict__ src, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

/* host words -> device through the staging buffer; returns with the data in place */
static int gx_upload(gx_cache* gc, void* dst, const void* src, size_t bytes) {
    lzb_ctx* c = gc->c;
    const u32* w = (const u32*)src; u32* d = (u32*)dst; size_t n = (bytes + 3) / 4;
    while (n) {
        size_t k = n < GX_STAGE_WORDS ? n : GX_STAGE_WORDS;
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

static void free_lane(cuda_lane& ln) {
    for (int s = 0; s < 2; s++) {
        if (ln.ownTbRow[s]) cudaFree(ln.tbRow[s]);
        if (ln.ownOps[s]) cudaFreeHost(ln.ops[s]);
        if (ln.ownAct[s]) cudaFree(ln.act[s]);
        if (ln.list[s]) cudaFreeHost(ln.list[s]);
    }
}
static void free_chunk(lane_chunk& ch) { cudaFree(ch.tb); cudaFree(ch.tbRow); cudaFree(ch.act); cudaFree(ch.ckpt); cudaFreeHost(ch.ops); }

void lzb_gapped_cache_free(lzb_ctx* c) {
    gx_cache* gc = (gx_cache*)c->gappedCache;
    if (!gc) return;
    for (auto& ln : gc->lanes) free_lane(ln);
    for (auto& ch : gc->chunks) free_chunk(ch);
    cudaFree(gc->d_segs); cudaFree(gc->d_aligns); cudaFreeHost(gc->h_stage); cudaFreeHost(gc->h_jobs);
    for (int k = 0; k < gc->nStreams; k++) cudaStreamDestroy(gc->streams[k]);
    cudaStreamDestroy(gc->upStream);
    delete gc; c->gappedCache = NULL;
}

#define GX_OPS_CAP (1u << 16)
#define GX_ACT_CAP 256u
/* GX_CHUNK more lanes out of five slabs */
static int make_chunk(gx_cache* gc) {
    lane_chunk ch; memset(&ch, 0, sizeof ch);
    const size_t tbStride = (((size_t)gc->tbBytes + 64) + 255) & ~(size_t)255, nside = 2 * GX_CHUNK;
    const u32 tbRowCap = gc->tbLen / 128 + 4096;
    const size_t ckWords = (size_t)GX_CKPT_CAP * CK_RECORD_WORDS;
    u32* d_ops = NULL;
    if (cudaMalloc(&ch.tb, tbStride * nside) != cudaSuccess || cudaMalloc(&ch.tbRow, (size_t)tbRowCap * 4 * nside) != cudaSuccess ||
        cudaMalloc(&ch.act, (size_t)GX_ACT_CAP * 5 * 4 * nside) != cudaSuccess || cudaMalloc(&ch.ckpt, ckWords * 4 * nside) != cudaSuccess ||
        cudaHostAlloc(&ch.ops, (size_t)GX_OPS_CAP * 4 * nside, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&d_ops, ch.ops, 0) != cudaSuccess) { free_chunk(ch); cudaGetLastError(); return -1; }
    gc->chunks.push_back(ch);
    for (int k = 0; k < GX_CHUNK; k++) {
        cuda_lane ln; memset(&ln, 0, sizeof ln);
        for (int s = 0; s < 2; s++) {
            const size_t e = (size_t)2 * k + s;
            ln.tb[s] = ch.tb + tbStride * e;
            ln.tbRow[s] = ch.tbRow + (size_t)tbRowCap * e; ln.tbRowCap[s] = tbRowCap;
            ln.ops[s] = ch.ops + (size_t)GX_OPS_CAP * e; ln.d_ops[s] = d_ops + (size_t)GX_OPS_CAP * e; ln.opsCap[s] = GX_OPS_CAP;
            ln.act[s] = ch.act + (size_t)GX_ACT_CAP * 5 * e; ln.actCap[s] = GX_ACT_CAP;
            ln.ckpt[s] = ch.ckpt + ckWords * e;
        }
        gc->lanes.push_back(ln);
    }
    return 0;
}

struct cuda_backend {
    gx_cache* gc;
    const char* error() { return gc->err; }
    u32 ckpt_every() { return gc->ckptEvery; }
    u32 ring(int mode) { return mode == 2 ? 4096u : 8192u; }
    int lanes(int want) {                                    /* called again when the scheduler wants more */
        if (want > GX_MAX_LANES) want = GX_MAX_LANES;
        const size_t laneBytes = 2 * ((size_t)gc->tbBytes + (size_t)(gc->tbLen / 128 + 4096) * 4 + (size_t)GX_CKPT_CAP * CK_RECORD_WORDS * 4 + 8192);
        while ((int)gc->lanes.size() < want) {
            size_t freeB = 0, totalB = 0;
            if (cudaMemGetInfo(&freeB, &totalB) != cudaSuccess) break;
            /* keep an eighth of the device for the seed stage and the caller; one chunk is always made */
            if (!gc->lanes.empty() && freeB < totalB / 8 + laneBytes * GX_CHUNK) {
                if (getenv("LZB_GAP_PROFILE")) fprintf(stderr, "[gx lanes] %zu lanes, %.1f GB of %.1f GB free: no more\n", gc->lanes.size(), freeB / 1e9, totalB / 1e9);
                break;
            }
            const int first = (int)gc->lanes.size();
            if (make_chunk(gc)) {
                if (gc->lanes.empty()) { lzb_fail("no device memory for the traceback of one Y-drop lane"); return -1; }
                if (getenv("LZB_GAP_PROFILE")) fprintf(stderr, "[gx lanes] %zu lanes, %.1f GB free: the next chunk could not be allocated\n", gc->lanes.size(), freeB / 1e9);
                break;
            }
            for (int z = first; z < (int)gc->lanes.size(); z++) { fill_job(z, 0); fill_job(z, 1); }
        }
        return (int)std::min<size_t>(gc->lanes.size(), (size_t)want);
    }
    void fill_job(int z, int s) {
        cuda_lane& ln = gc->lanes[z]; dp_job& J = gc->h_jobs[2 * z + s];
        memset(&J, 0, sizeof J);
        J.tb = ln.tb[s]; J.tbLen = gc->tbLen; J.tbRow = ln.tbRow[s]; J.tbRowCap = ln.tbRowCap[s];
        J.ops = ln.d_ops[s]; J.opsCap = ln.opsCap[s]; J.act = ln.act[s]; J.actCap = ln.actCap[s];
        J.ckpt = ln.ckpt[s]; J.ckptCap = GX_CKPT_CAP; J.ckptEvery = gc->ckptEvery; J.resume = -1;
    }
    dp_job* job(int z, int s) { return &gc->h_jobs[2 * z + s]; }
    const u32* ops(int z, int s) { return gc->lanes[z].ops[s]; }
    int* list(int z, int s, size_t n) {
        cuda_lane& ln = gc->lanes[z];
        if (n > ln.listCap[s]) {
            if (ln.list[s]) cudaFreeHost(ln.list[s]);
            ln.list[s] = NULL; ln.listCap[s] = 0;
            const size_t cap = std::max<size_t>(1024, n * 2);
            if (cudaHostAlloc(&ln.list[s], cap * sizeof(int), cudaHostAllocMapped) != cudaSuccess) { lzb_fail("cudaHostAlloc of an alignment list failed"); return NULL; }
            if (cudaHostGetDevicePointer((void**)&ln.d_list[s], ln.list[s], 0) != cudaSuccess) { lzb_fail("cudaHostGetDevicePointer failed"); return NULL; }
            ln.listCap[s] = cap;
        }
        return ln.list[s];
    }
    /* every running job must have finished before a table moves */
    int drain() {
        for (size_t k = 0; k < 2 * gc->lanes.size(); k++) {
            dp_job& J = gc->h_jobs[k];
            const auto t0 = std::chrono::steady_clock::now();
            while (gc->launched[k] != 0 && J.done != gc->launched[k]) {
                if (!poll()) return lzb_fail("Y-drop kernel failed: %s", gc->err);
                if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 120.0) return lzb_fail("a Y-drop sweep did not come back within 120 s");
            }
        }
        return 0;
    }
    int tables(const dseg* s, size_t s0, size_t s1, const dalign* a, size_t a0, size_t a1) {
        if (s1 > gc->segsCap) {
            if (drain()) return -1;
            cudaFree(gc->d_segs); gc->segsCap = s1 * 2 + (4u << 20);
            CUDA_TRY(cudaMalloc(&gc->d_segs, gc->segsCap * sizeof(dseg)));
            s0 = 0;
        }
        if (a1 > gc->alignsCap) {
            if (drain()) return -1;
            cudaFree(gc->d_aligns); gc->alignsCap = a1 * 2 + (1u << 18);
            CUDA_TRY(cudaMalloc(&gc->d_aligns, gc->alignsCap * sizeof(dalign)));
            a0 = 0;
        }
        if (s1 > s0 && gx_upload(gc, gc->d_segs + s0, s + s0, (s1 - s0) * sizeof(dseg))) return -1;
        if (a1 > a0 && gx_upload(gc, gc->d_aligns + a0, a + a0, (a1 - a0) * sizeof(dalign))) return -1;
        return 0;
    }
    const dseg* segs() { return gc->d_segs; }
    const dalign* aligns() { return gc->d_aligns; }
    int launch(int mode, const u16* ix, int n) {
        lzb_ctx* c = gc->c;
        launch_list ll; memset(&ll, 0, sizeof ll);
        for (int k = 0; k < n; k++) {
            ll.ix[k] = ix[k];
            /* the device-side addresses of the job's list */
            dp_job& J = gc->h_jobs[ix[k]]; cuda_lane& ln = gc->lanes[ix[k] >> 1];
            if (J.listv == ln.list[ix[k] & 1]) J.listv = ln.d_list[ix[k] & 1];
        }
        std::atomic_thread_fence(std::memory_order_seq_cst);
        /* A stream runs its launches one after the other, and a launch lasts as long as its longest sweep: a short resume
         * queued behind a batch of fresh sweeps would wait for all of them.  So every launch takes a stream whose last
         * launch has finished (all its jobs have signalled), a new one if there is none. */
        int pick = -1;
        while (pick < 0) {
            for (int k = 0; k < gc->nStreams && pick < 0; k++) {
                bool idle = true;
                for (auto& jt : gc->inflight[k]) if (gc->h_jobs[jt.first].done != jt.second && gc->launched[jt.first] == jt.second) { idle = false; break; }
                if (idle) pick = k;
            }
            if (pick < 0 && gc->nStreams < GX_STREAMS) {
                CUDA_TRY(cudaStreamCreateWithFlags(&gc->streams[gc->nStreams], cudaStreamNonBlocking));
                pick = gc->nStreams++;
            }
            if (pick < 0 && !poll()) return lzb_fail("Y-drop kernel failed: %s", gc->err);
        }
        cudaStream_t st = gc->streams[pick];
        gc->inflight[pick].clear();
        for (int k = 0; k < n; k++) gc->inflight[pick].push_back(std::make_pair(ix[k], gc->h_jobs[ix[k]].token));
        if (mode == 0)
            k_ydrop_mw<8, 4><<<n, 128, 0, st>>>(gc->d_jobs, ll, gc->d_segs, gc->cls1, gc->cls2, gc->len1, gc->len2, c->d_sc, gc->yDrop, gc->trim);
        else if (mode == 1)
            /* 47 KB of (unused) dynamic shared memory per one-warp CTA: at most four of them fit an SM, one per scheduler --
             * left to itself the block scheduler stacks dozens of 32-thread CTAs on some SMs while others idle, and the
             * sweeps there run at a fraction of their speed (measured: 1.8 us a row alone, 3.4 median, 9 worst) */
            k_ydrop_warp<16><<<n, 32, gc->warpSmem, st>>>(gc->d_jobs, ll, gc->d_segs, gc->cls1, gc->cls2, gc->len1, gc->len2, c->d_sc, gc->yDrop, gc->trim);
        else {
            const u32 rg = ring(mode);
            size_t smem = (size_t)rg * 17 + LZB_MAX_CLASSES * LZB_MAX_CLASSES * 4 + 1024;
            k_ydrop<256><<<n, 256, smem, st>>>(gc->d_jobs, ll, gc->d_segs, gc->cls1, gc->cls2, gc->len1, gc->len2, c->d_sc, gc->yDrop, gc->trim, rg);
        }
        c->launches++;
        CUDA_TRY(cudaGetLastError());
        for (int k = 0; k < n; k++) gc->launched[ix[k]] = gc->h_jobs[ix[k]].token;      /* only now is there something to wait for */
        return 0;
    }
    bool poll() {
        gc->polls++;
        if ((gc->polls & 1023) == 0) {
            for (int k = 0; k < gc->nStreams; k++) {
                cudaError_t e = cudaStreamQuery(gc->streams[k]);
                if (e != cudaSuccess && e != cudaErrorNotReady) { snprintf(gc->err, sizeof gc->err, "%s", cudaGetErrorString(e)); return false; }
            }
        }
        std::this_thread::sleep_for(std::chrono::microseconds(15));
        return true;
    }
    int grow(int z, int s, int what) {
        cuda_lane& ln = gc->lanes[z];
        if (what == DP_TBROW) {
            if (ln.ownTbRow[s]) cudaFree(ln.tbRow[s]);
            ln.ownTbRow[s] = true; ln.tbRowCap[s] = ln.tbRowCap[s] < gc->tbLen / 4 ? ln.tbRowCap[s] * 4 : gc->tbLen + 8;
            CUDA_TRY(cudaMalloc(&ln.tbRow[s], (size_t)ln.tbRowCap[s] * 4));
        } else if (what == DP_ACT) {
            if (ln.ownAct[s]) cudaFree(ln.act[s]);
            ln.ownAct[s] = true; ln.actCap[s] *= 4; CUDA_TRY(cudaMalloc(&ln.act[s], (size_t)ln.actCap[s] * 5 * 4));
        } else if (what == DP_OPS) {
            if (ln.ownOps[s]) cudaFreeHost(ln.ops[s]);
            ln.ownOps[s] = true; ln.opsCap[s] *= 4;
            CUDA_TRY(cudaHostAlloc(&ln.ops[s], (size_t)ln.opsCap[s] * 4, cudaHostAllocMapped));
            CUDA_TRY(cudaHostGetDevicePointer((void**)&ln.d_ops[s], ln.ops[s], 0));
        }
        dp_job& J = gc->h_jobs[2 * z + s];
        J.tbRow = ln.tbRow[s]; J.tbRowCap = ln.tbRowCap[s]; J.ops = ln.d_ops[s]; J.opsCap = ln.opsCap[s]; J.act = ln.act[s]; J.actCap = ln.actCap[s];
        return 0;
    }
};

extern "C" int lzb_gapped_extend(lzb_ctx* c, lzb_target* t, lzb_query* q, const uint8_t* h1, const uint8_t* h2,
                                 lzb_segment* anchors, uint64_t n, const lzb_gapped_params* P,
                                 lzb_alignel** list, lzb_gapped_stats* stats) {
    cudaSetDevice(c->device);
    if (!c->haveScoring) return lzb_fail("lzb_set_scoring has not been called");
    if (P->tracebackBytes < 8) return lzb_fail("in new_traceback(), size can't be %u", P->tracebackBytes);
    if (c->sc.gapOpen < 0) return lzb_fail("lastz_b200's Y-drop kernel requires a non-negative gap open penalty (got %d)", c->sc.gapOpen);
    const u64 launches0 = c->launches;
    gx_cache* gc = (gx_cache*)c->gappedCache;
    if (gc && gc->tbBytes != P->tracebackBytes) { lzb_gapped_cache_free(c); gc = NULL; }
    if (!gc) {
        gc = new gx_cache();                               /* value-initialised: every scalar member starts at zero */
        gc->c = c; gc->tbBytes = P->tracebackBytes; gc->tbLen = 1 + (P->tracebackBytes - 8);
        gc->ckptEvery = 256;
        { const char* e = getenv("LZB_CKPT_EVERY"); if (e) { int v = atoi(e); if (v >= 32 && v % 32 == 0) gc->ckptEvery = (u32)v; } }
        c->gappedCache = gc;
        CUDA_TRY(cudaStreamCreateWithFlags(&gc->upStream, cudaStreamNonBlocking));
        CUDA_TRY(cudaHostAlloc(&gc->h_stage, (size_t)GX_STAGE_WORDS * 4, cudaHostAllocMapped));
        CUDA_TRY(cudaHostGetDevicePointer((void**)&gc->d_stage, gc->h_stage, 0));
        CUDA_TRY(cudaHostAlloc(&gc->h_jobs, (size_t)2 * GX_MAX_LANES * sizeof(dp_job), cudaHostAllocMapped));
        CUDA_TRY(cudaHostGetDevicePointer((void**)&gc->d_jobs, gc->h_jobs, 0));
        memset(gc->h_jobs, 0, (size_t)2 * GX_MAX_LANES * sizeof(dp_job));
        gc->segsCap = 4u << 20; CUDA_TRY(cudaMalloc(&gc->d_segs, gc->segsCap * sizeof(dseg)));
        gc->alignsCap = 1u << 18; CUDA_TRY(cudaMalloc(&gc->d_aligns, gc->alignsCap * sizeof(dalign)));
        CUDA_TRY(cudaFuncSetAttribute(k_ydrop<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        CUDA_TRY(cudaFuncSetAttribute(k_ydrop_warp<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 47 * 1024));
        gc->warpSmem = 47 * 1024;
        { const char* e = getenv("LZB_WARP_SMEM_KB"); if (e) { int v = atoi(e); if (v >= 0 && v <= 47) gc->warpSmem = (size_t)v * 1024; } }
    }
    gc->cls1 = t->d_cls; gc->cls2 = q->d_cls; gc->len1 = t->len; gc->len2 = q->len; gc->yDrop = P->yDrop; gc->trim = P->trimToPeak;
    cuda_backend B; B.gc = gc;
    gx_input in; in.h_seq1 = t->h_seq; in.h_seq2 = q->h_seq; in.len1 = t->len; in.len2 = q->len;
    in.hostSub = c->hostSub; in.gapOpen = c->sc.gapOpen; in.gapExtend = c->sc.gapExtend;
    int rc = gx_run(B, in, anchors, n, P, list, stats, lzb_fail);
    if (rc == 0) {
        /* the alignments point at the CALLER's bytes (may be NULL), the scheduler worked on the context's copies */
        for (lzb_alignel* a = *list; a; a = a->next) { a->seq1 = h1; a->seq2 = h2; }
        if (stats) stats->launches = c->launches - launches0;
    } else B.drain();
    return rc;
}

extern "C" void lzb_free_align_list(lzb_alignel* a) {
    while (a) { lzb_alignel* nx = a->next; free(a->script); free(a); a = nx; }
}
