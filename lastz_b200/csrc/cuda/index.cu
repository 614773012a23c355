/*
 * index.cu -- K1: the target's seed position table, built and kept in HBM.
 *
 * Replaces build_seed_position_table / record_seed_positions / add_word
 * (pos_table.c:144, :396, :1326).  The reference keeps last[word] -> prev[] chains that
 * find_table_matches walks from the largest position down (seed_search.c:832).  Here the table is
 * CSR: off[2^wordBits+1] and pos[], each word's positions contiguous and in DEcreasing order, so
 * a probe is one coalesced read of the list instead of a pointer chase.
 *
 * Build = one pass that emits (word, position) for every all-valid window ending on a step
 * multiple, visiting positions from the top of the sequence down, then a STABLE radix sort on
 * the word bits only (cub) -- stability preserves the descending position order inside a word --
 * and an exclusive scan of the word histogram for the offsets.
 */
#include <cub/cub.cuh>
#include <stdlib.h>
#include <string.h>
#include "lzb_cuda.h"

#include "index_kernels.cuh"

extern "C" lzb_target* lzb_target_build(lzb_ctx* c, const uint8_t* seq1, uint32_t len1, uint32_t start,
                                        uint32_t end, const int8_t ctb[256], const lzb_seed* seed, uint32_t step) {
    cudaSetDevice(c->device);
    if (step < 1) { lzb_fail("in build_seed_position_table(), step can't be %u", step); return NULL; }
    if (len1 > 0x7FFFFFFFu) { lzb_fail("sequence length %u exceeds maximum (positions are 31-bit like the reference's default build; lastz_32 widths are not built)", len1); return NULL; }
    if (end == 0) end = len1;
    if (end <= start || end > len1) { lzb_fail("in build_seed_position_table(), interval is bad (%u..%u of %u)", start, end, len1); return NULL; }
    if (seed->weight > 28) { lzb_fail("new_position_table can't support >28 seed bits (%d requested)", seed->weight); return NULL; }
    lzb_target* t = (lzb_target*)calloc(1, sizeof *t);
    t->ctx = c; t->len = len1; t->start = start; t->end = end; t->step = step;
    t->wordBits = seed->weight; t->seedLength = seed->length;
    t->h_seq = (u8*)malloc((size_t)len1 + 1); memcpy(t->h_seq, seq1, len1); t->h_seq[len1] = 0;
    size_t capUnused = 0;
    if (lzb_upload_classes(c, t->h_seq, len1, &t->d_seq, &t->d_cls, &capUnused)) return NULL;

    u64 nw = 1ull << seed->weight;
    CUDA_TRYP(cudaMalloc(&t->d_off, (nw + 2) * 4));
    CUDA_TRYP(cudaMemsetAsync(t->d_off, 0, (nw + 2) * 4, c->stream));
    u32 L = (u32)seed->length;
    u64 nent = 0; u32 pmax = 0;
    if (len1 >= L && end - start >= L) {
        pmax = end - end % step;
        u32 pmin = start + L;                       /* smallest admissible end position */
        if (pmax >= pmin) nent = (u64)(pmax - pmin) / step + 1;
    }
    if (nent == 0) {
        CUDA_TRYP(cudaMalloc(&t->d_pos, 16));
        t->npos = 0;
        CUDA_TRYP(cudaStreamSynchronize(c->stream));
        return t;
    }
    u32 *keysA, *keysB, *valsA, *valsB, *hist;
    CUDA_TRYP(cudaMalloc(&keysA, nent * 4)); CUDA_TRYP(cudaMalloc(&keysB, nent * 4));
    CUDA_TRYP(cudaMalloc(&valsA, nent * 4)); CUDA_TRYP(cudaMalloc(&valsB, nent * 4));
    CUDA_TRYP(cudaMalloc(&hist, (nw + 2) * 4));
    CUDA_TRYP(cudaMemsetAsync(hist, 0, (nw + 2) * 4, c->stream));
    seed_dev sd; seed_to_dev(&sd, seed);
    ctb_dev cd; memcpy(cd.v, ctb, 256);
    int blocks = (int)((nent + 255) / 256); if (blocks > c->smCount * 16) blocks = c->smCount * 16;
    k_index_words<<<blocks, 256, 0, c->stream>>>(t->d_seq, start, pmax, step, nent, sd, cd, seed->weight,
                                                keysA, valsA, hist);
    c->launches++;
    CUDA_TRYP(cudaGetLastError());
    /* stable LSD radix sort on the word bits (+ the sentinel bit) */
    void* tmp = NULL; size_t tmpBytes = 0;
    CUDA_TRYP(cub::DeviceRadixSort::SortPairs(NULL, tmpBytes, keysA, keysB, valsA, valsB, nent, 0, seed->weight + 1, c->stream));
    CUDA_TRYP(cudaMalloc(&tmp, tmpBytes));
    CUDA_TRYP(cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, keysA, keysB, valsA, valsB, nent, 0, seed->weight + 1, c->stream));
    cudaFree(tmp); tmp = NULL; tmpBytes = 0;
    CUDA_TRYP(cub::DeviceScan::ExclusiveSum(NULL, tmpBytes, hist, t->d_off, nw + 1, c->stream));
    CUDA_TRYP(cudaMalloc(&tmp, tmpBytes));
    CUDA_TRYP(cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, hist, t->d_off, nw + 1, c->stream));
    u32 total = 0;
    CUDA_TRYP(cudaMemcpyAsync(&total, t->d_off + nw, 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRYP(cudaStreamSynchronize(c->stream));
    t->npos = total;
    CUDA_TRYP(cudaMalloc(&t->d_pos, (size_t)(total ? total : 1) * 4 + 16));
    CUDA_TRYP(cudaMemcpyAsync(t->d_pos, valsB, (size_t)total * 4, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRYP(cudaStreamSynchronize(c->stream));
    cudaFree(tmp); cudaFree(keysA); cudaFree(keysB); cudaFree(valsA); cudaFree(valsB); cudaFree(hist);
    return t;
}

extern "C" int lzb_target_limit(lzb_target* t, uint32_t limit) {
    lzb_ctx* c = t->ctx;
    cudaSetDevice(c->device);
    if (t->npos == 0) return 0;
    const u64 nw = 1ull << t->wordBits;
    u32 *cnt = NULL, *newOff = NULL, *newPos = NULL; void* tmp = NULL; size_t tmpBytes = 0;
    CUDA_TRY(cudaMalloc(&cnt, (nw + 2) * 4)); CUDA_TRY(cudaMalloc(&newOff, (nw + 2) * 4));
    int blocks = (int)((nw + 255) / 256); if (blocks > c->smCount * 16) blocks = c->smCount * 16;
    k_limit_counts<<<blocks, 256, 0, c->stream>>>(t->d_off, cnt, nw, limit);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(NULL, tmpBytes, cnt, newOff, nw + 1, c->stream));
    CUDA_TRY(cudaMalloc(&tmp, tmpBytes));
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, cnt, newOff, nw + 1, c->stream));
    u32 total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, newOff + nw, 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMalloc(&newPos, (size_t)(total ? total : 1) * 4 + 16));
    k_limit_compact<<<c->smCount * 16, 256, 0, c->stream>>>(t->d_off, newOff, t->d_pos, newPos, nw);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    cudaFree(t->d_off); cudaFree(t->d_pos); cudaFree(cnt); cudaFree(tmp);
    t->d_off = newOff; t->d_pos = newPos; t->npos = total;
    return 0;
}

extern "C" void lzb_target_free(lzb_target* t) {
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    cudaFree(t->d_seq); cudaFree(t->d_cls); cudaFree(t->d_off); cudaFree(t->d_pos);
    free(t->h_seq); free(t);
}

extern "C" int64_t lzb_target_export_index(lzb_target* t, uint32_t* counts, uint32_t* positions) {
    cudaSetDevice(t->ctx->device);
    u64 nw = 1ull << t->wordBits;
    u32* off = (u32*)malloc((nw + 1) * 4);
    if (cudaMemcpy(off, t->d_off, (nw + 1) * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { free(off); return lzb_fail("index export failed"); }
    for (u64 k = 0; k < nw; k++) counts[k] = off[k + 1] - off[k];
    free(off);
    if (positions && t->npos)
        if (cudaMemcpy(positions, t->d_pos, t->npos * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return lzb_fail("index export failed");
    return (int64_t)t->npos;
}
