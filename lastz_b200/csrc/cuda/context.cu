/*
 * context.cu -- device context, scoring classes, sequence residency.
 * Part of liblastz_b200.so (sm_100a only; there is no CPU path: lzb_open fails without a GPU).
 */
#include <chrono>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "lzb_cuda.h"

static thread_local char g_err[1024];

int lzb_fail(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
    return -1;
}

extern "C" const char* lzb_last_error(void) { return g_err; }
extern "C" const char* lzb_backend(void) { return "cuda-sm_100a"; }
extern "C" void lzb_free(void* p) { free(p); }
extern "C" uint64_t lzb_launch_count(lzb_ctx* c) { return c ? c->launches : 0; }

extern "C" lzb_ctx* lzb_open(int device) {
    /* The gapped stage keeps up to 32 streams busy; with the default 8 hardware queues unrelated
     * streams share a queue and a kernel waits behind another stream's pending D2H copy.  Must be
     * set before the CUDA context exists (a host process that created it earlier sets it itself). */
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        lzb_fail("lastz_b200 needs an sm_100 GPU and found none (%s); there is no CPU fallback",
                 e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
        return NULL;
    }
    if (device < 0 || device >= n) { lzb_fail("cuda device %d does not exist (%d present)", device, n); return NULL; }
    cudaDeviceProp prop;
    CUDA_TRYP(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        lzb_fail("device %d is sm_%d%d; liblastz_b200 carries sm_100a code only", device, prop.major, prop.minor);
        return NULL;
    }
    CUDA_TRYP(cudaSetDevice(device));
    lzb_ctx* c = (lzb_ctx*)calloc(1, sizeof *c);
    c->device = device; c->smCount = prop.multiProcessorCount;
    CUDA_TRYP(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRYP(cudaMalloc(&c->d_sc, sizeof(lzb_scoring_dev)));
    return c;
}

extern "C" void lzb_close(lzb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    lzb_gapped_cache_free(c);
    lzb_seed_scratch_free(c);
    cudaFree(c->peaksBuf);
    for (int i = 0; i < 4; i++) { cudaFree(c->qpool[i].seq); cudaFree(c->qpool[i].cls); }
    cudaStreamDestroy(c->stream);
    cudaFree(c->d_sc);
    free(c->hostSub); free(c->hostMsub);
    free(c);
}

/* group bytes whose rows and columns agree in both matrices */
extern "C" int lzb_set_scoring(lzb_ctx* c, const int32_t* sub, const int32_t* msub, int32_t go, int32_t ge) {
    cudaSetDevice(c->device);
    free(c->hostSub); free(c->hostMsub);
    c->hostSub = (s32*)malloc(65536 * 4); c->hostMsub = (s32*)malloc(65536 * 4);
    memcpy(c->hostSub, sub, 65536 * 4); memcpy(c->hostMsub, msub, 65536 * 4);
    lzb_scoring_dev* sc = &c->sc;
    memset(sc, 0, sizeof *sc);
    int rep[LZB_MAX_CLASSES]; int nc = 0;
    /* A, C, G, T are numbered first (classes 0..3 for any DNA score set): k_extend2's pair table relies
     * on it for conflict-free shared-memory banks, nothing relies on it for correctness */
    for (int bi = 0; bi < 256 + 4; bi++) {
        const int b = bi < 4 ? "ACGT"[bi] : bi - 4;
        if (bi >= 4 && (b == 'A' || b == 'C' || b == 'G' || b == 'T')) continue;
        int found = -1;
        for (int k = 0; k < nc && found < 0; k++) {
            int r = rep[k]; bool same = true;
            for (int x = 0; x < 256 && same; x++)
                same = sub[b * 256 + x] == sub[r * 256 + x] && sub[x * 256 + b] == sub[x * 256 + r] &&
                       msub[b * 256 + x] == msub[r * 256 + x] && msub[x * 256 + b] == msub[x * 256 + r];
            if (same) found = k;
        }
        if (found < 0) {
            if (nc == LZB_MAX_CLASSES)
                return lzb_fail("the scoring set distinguishes more than %d byte classes; lastz_b200 supports DNA score sets only", LZB_MAX_CLASSES);
            rep[nc] = b; found = nc++;
        }
        sc->cls[b] = (u8)found;
    }
    sc->numClasses = nc;
    for (int i = 0; i < nc; i++) for (int j = 0; j < nc; j++) {
        sc->subC[i * LZB_MAX_CLASSES + j] = sub[rep[i] * 256 + rep[j]];
        sc->msubC[i * LZB_MAX_CLASSES + j] = msub[rep[i] * 256 + rep[j]];
    }
    sc->gapOpen = go; sc->gapExtend = ge;
    CUDA_TRY(cudaMemcpyAsync(c->d_sc, sc, sizeof *sc, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->haveScoring = true;
    return 0;
}

#include "classify_kernel.cuh"

/* copies len bytes (+ NUL + zero pad to 16) to the device and derives the class-code copy */
int lzb_upload_classes(lzb_ctx* c, const u8* h_seq, u32 len, u8** d_seq, u8** d_cls, size_t* cap) {
    if (!c->haveScoring) return lzb_fail("lzb_set_scoring has not been called");
    size_t padded = (((size_t)len + 1 + 15) / 16) * 16 + 16;
    int best = -1;                                       /* smallest pooled pair that is large enough */
    for (int i = 0; i < 4; i++) if (c->qpool[i].seq && c->qpool[i].cap >= padded && (best < 0 || c->qpool[i].cap < c->qpool[best].cap)) best = i;
    if (best >= 0) { *d_seq = c->qpool[best].seq; *d_cls = c->qpool[best].cls; *cap = c->qpool[best].cap; c->qpool[best].seq = c->qpool[best].cls = NULL; c->qpool[best].cap = 0; }
    else {
        CUDA_TRY(cudaMalloc(d_seq, padded));
        CUDA_TRY(cudaMalloc(d_cls, padded));
        *cap = padded;
    }
    CUDA_TRY(cudaMemsetAsync(*d_seq, 0, padded, c->stream));
    CUDA_TRY(cudaMemcpyAsync(*d_seq, h_seq, len, cudaMemcpyHostToDevice, c->stream));
    size_t n16 = padded / 16;
    int blocks = (int)((n16 + 255) / 256); if (blocks > c->smCount * 8) blocks = c->smCount * 8;
    k_classify<<<blocks, 256, 0, c->stream>>>((const uint4*)*d_seq, (uint4*)*d_cls, n16, c->d_sc);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" lzb_query* lzb_query_load(lzb_ctx* c, const uint8_t* seq2, uint32_t len2) {
    if (len2 > 0x7FFFFFFFu) { lzb_fail("sequence length %u exceeds maximum (positions are 31-bit like the reference's default build; lastz_32 widths are not built)", len2); return NULL; }
    cudaSetDevice(c->device);
    const bool wtrace = getenv("LZB_SEED_TRACE") != NULL;
    const auto w0 = std::chrono::steady_clock::now();
    lzb_query* q = (lzb_query*)calloc(1, sizeof *q);
    q->ctx = c; q->len = len2;
    q->h_seq = (u8*)malloc((size_t)len2 + 1); memcpy(q->h_seq, seq2, len2); q->h_seq[len2] = 0;
    const double w1 = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
    if (lzb_upload_classes(c, q->h_seq, len2, &q->d_seq, &q->d_cls, &q->cap)) { free(q->h_seq); free(q); return NULL; }
    if (wtrace) fprintf(stderr, "[query load] host copy=%.4f upload+classify enqueue=%.4f s (%u bp)\n", w1,
                        std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count() - w1, len2);
    return q;
}

extern "C" void lzb_query_free(lzb_query* q) {
    if (!q) return;
    cudaSetDevice(q->ctx->device);
    lzb_ctx* c = q->ctx;
    cudaStreamSynchronize(c->stream);
    int slot = -1;
    for (int i = 0; i < 4 && slot < 0; i++) if (!c->qpool[i].seq) slot = i;
    if (slot < 0) {                                      /* pool full: evict the smallest if this pair is larger */
        int sm = 0; for (int i = 1; i < 4; i++) if (c->qpool[i].cap < c->qpool[sm].cap) sm = i;
        if (c->qpool[sm].cap < q->cap) { cudaFree(c->qpool[sm].seq); cudaFree(c->qpool[sm].cls); slot = sm; }
    }
    if (slot >= 0) { c->qpool[slot].seq = q->d_seq; c->qpool[slot].cls = q->d_cls; c->qpool[slot].cap = q->cap; }
    else { cudaFree(q->d_seq); cudaFree(q->d_cls); }
    free(q->h_seq); free(q);
}
