/*
 * index_kernels.cuh -- K1 k_index_words: (word, position) pairs of the target, positions descending (index.cu).
 * Device code only (also compiled for the host block emulator, tests/warp_emu/cuda_emu.h).
 */
#ifndef LZB_INDEX_KERNELS_CUH
#define LZB_INDEX_KERNELS_CUH

struct ctb_dev { int8_t v[256]; };

/* entry j (0-based) is the window ending at pos = pmax - j*step; key = packed word, or the
 * sentinel 1<<wordBits when the window holds an invalid character */
__global__ void k_index_words(const u8* __restrict__ seq, u32 start, u32 pmax, u32 step, u64 nent,
                              seed_dev sd, ctb_dev ctb, int wordBits,
                              u32* __restrict__ keys, u32* __restrict__ vals, u32* __restrict__ hist) {
    for (u64 j = blockIdx.x * (u64)blockDim.x + threadIdx.x; j < nent; j += (u64)gridDim.x * blockDim.x) {
        u32 pos = pmax - (u32)(j * step);
        u64 w = 0; bool ok = true;
        u32 first = pos - (u32)sd.length;
        (void)start;
        for (int k = 0; k < sd.length; k++) {
            int b = ctb.v[seq[first + k]];
            ok = ok && (b >= 0);
            w = (w << 2) | (u64)(b & 3);
        }
        u32 word = 0;
        for (int p = 0; p < sd.numParts; p++) word |= (u32)(w >> sd.shift[p]) & sd.mask[p];
        u32 key = ok ? word : (1u << wordBits);
        keys[j] = key; vals[j] = pos;
        if (ok) atomicAdd(&hist[word], 1u);
    }
}


/* ---- limit_position_table (pos_table.c:1763): words with more than `limit` positions lose their list ---- */
__global__ void k_limit_counts(const u32* __restrict__ off, u32* __restrict__ cnt, u64 nw, u32 limit) {
    for (u64 w = blockIdx.x * (u64)blockDim.x + threadIdx.x; w <= nw; w += (u64)gridDim.x * blockDim.x) {
        const u32 c = w < nw ? off[w + 1] - off[w] : 0u;
        cnt[w] = c > limit ? 0u : c;
    }
}
/* one warp per kept word copies its list to the word's new place */
__global__ void k_limit_compact(const u32* __restrict__ off, const u32* __restrict__ newOff, const u32* __restrict__ pos,
                                u32* __restrict__ newPos, u64 nw) {
    const u32 lane = threadIdx.x & 31u;
    for (u64 w = (blockIdx.x * (u64)blockDim.x + threadIdx.x) >> 5; w < nw; w += ((u64)gridDim.x * blockDim.x) >> 5) {
        const u32 n = newOff[w + 1] - newOff[w];
        const u32 a = off[w], b = newOff[w];
        for (u32 i = lane; i < n; i += 32u) newPos[b + i] = pos[a + i];
    }
}
#endif
