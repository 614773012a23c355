/*
 * classify_kernel.cuh -- K0 k_classify: ASCII bytes -> byte-class codes (context.cu).
 * Device code only (also compiled for the host block emulator, tests/warp_emu/cuda_emu.h).
 */
#ifndef LZB_CLASSIFY_KERNEL_CUH
#define LZB_CLASSIFY_KERNEL_CUH

/* K0: ASCII -> class codes, 16 bytes per thread (128-bit loads/stores) */
__global__ void k_classify(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n16,
                           const lzb_scoring_dev* __restrict__ sc) {
    __shared__ u8 lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = sc->cls[i];
    __syncthreads();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = in[i]; u32 w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int k = 0; k < 4; k++) {
            u32 x = w[k];
            w[k] = (u32)lut[x & 255] | ((u32)lut[(x >> 8) & 255] << 8) | ((u32)lut[(x >> 16) & 255] << 16) | ((u32)lut[x >> 24] << 24);
        }
        out[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

#endif
