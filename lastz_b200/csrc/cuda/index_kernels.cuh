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

#endif
