/*
 * peaks_kernel.cuh -- K4 k_peaks: segment_peak gapped_extend.c:515-559 (gapped.cu).
 * Device code only (also compiled for the host block emulator, tests/warp_emu/cuda_emu.h).
 */
#ifndef LZB_PEAKS_KERNEL_CUH
#define LZB_PEAKS_KERNEL_CUH

/* ---- K4: segment_peak gapped_extend.c:515-559, one thread per HSP ---- */
__global__ void k_peaks(lzb_segment* __restrict__ seg, u64 n, const u8* __restrict__ cls1,
                        const u8* __restrict__ cls2, const lzb_scoring_dev* __restrict__ sc) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        lzb_segment g = seg[i];
        u32 peak;
        if (g.length <= 31) peak = g.length / 2;
        else {
            const u8* s1 = cls1 + g.pos1; const u8* s2 = cls2 + g.pos2;
            s32 sum = 0;
            for (u32 k = 0; k < 31; k++) sum += sc->subC[s1[k] * LZB_MAX_CLASSES + s2[k]];
            s32 bestv = sum; peak = 15;
            for (u32 k = 31; k < g.length; k++) {
                sum -= sc->subC[s1[k - 31] * LZB_MAX_CLASSES + s2[k - 31]];
                sum += sc->subC[s1[k] * LZB_MAX_CLASSES + s2[k]];
                if (sum > bestv) { bestv = sum; peak = k - 15; }
            }
        }
        g.pos1 += peak; g.pos2 += peak; g.length = 0;
        seg[i] = g;
    }
}

#endif
