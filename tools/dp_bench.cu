/*
 * dp_bench.cu -- microbenchmark of the Y-drop sweep kernels (measurement tool, not product code).
 *
 * One synthetic homologous pair (the SURVEY 8d channel), one anchor in the middle, the forward and the reverse
 * one-sided DP of that anchor run by each kernel shape; prints microseconds per DP row when the two DPs have
 * the GPU to themselves and when `copies` identical pairs run side by side (what the anchor loop does).
 *   dp_bench [len=1000000] [tracebackMiB=80] [copies=296]
 * Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/dp_bench tools/dp_bench.cu
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../include/lastz_b200.h"
#include "../lastz_b200/csrc/cuda/lzb_types.h"
#include "../lastz_b200/csrc/cuda/ydrop_common.cuh"
#include "../lastz_b200/csrc/cuda/ydrop_warp.cuh"
#include "../lastz_b200/csrc/cuda/ydrop_mw.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

static uint64_t st;
static uint64_t next64(void) {
    st += 0x9E3779B97F4A7C15ull; uint64_t z = st;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static const s32 HOX[4][4] = { { 91, -114, -31, -123 }, { -114, 100, -125, -31 }, { -31, -125, 100, -114 }, { -123, -31, -114, 91 } };

struct shape { const char* name; int threads; void (*launch)(int grid, dp_job* jobs, const u8* c1, const u8* c2, u32 l1, u32 l2, const lzb_scoring_dev* sc); };
static launch_list g_ll;
#define MW(K, NW, MB) { "mw<" #K "," #NW "," #MB ">", 32 * NW, [](int g, dp_job* j, const u8* c1, const u8* c2, u32 l1, u32 l2, const lzb_scoring_dev* sc) { k_ydrop_mw<K, NW, MB><<<g, 32 * NW>>>(j, g_ll, (const dseg*)NULL, c1, c2, l1, l2, sc, 9400, 1); } }
#define WP(K) { "warp<" #K ">", 32, [](int g, dp_job* j, const u8* c1, const u8* c2, u32 l1, u32 l2, const lzb_scoring_dev* sc) { k_ydrop_warp<K><<<g, 32>>>(j, g_ll, (const dseg*)NULL, c1, c2, l1, l2, sc, 9400, 1); } }
static shape shapes[] = { MW(8, 4, 1), MW(8, 4, 4), WP(16) };

int main(int argc, char** argv) {
    const u32 L = argc > 1 ? (u32)atol(argv[1]) : 1000000u;
    const u32 tbMiB = argc > 2 ? (u32)atol(argv[2]) : 80u;
    const int copies = argc > 3 ? atoi(argv[3]) : 296;
    const int ckpt = argc > 4 ? atoi(argv[4]) : 1;          /* checkpoints every 256 rows, as the product takes them */
    for (int k = 0; k < LZB_LAUNCH_MAX; k++) g_ll.ix[k] = (u16)k;
    std::string t(L, 'A'), q; std::vector<u32> qposOf(L);
    st = 20260925; for (u32 i = 0; i < L; i++) t[i] = "ACGT"[next64() >> 62];
    st = 20260926;
    for (u32 i = 0; i < L; i++) {
        char c = t[i]; int code = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3;
        uint64_t r = next64(); uint32_t u = (uint32_t)(r & 0xFFFFFF);
        qposOf[i] = (u32)q.size();
        if (u < 671089) q.push_back("ACGT"[(code + 1 + (int)((r >> 24) % 3)) & 3]);
        else if (u < 754975) ;
        else if (u < 838861) { q.push_back(c); q.push_back("ACGT"[(r >> 24) & 3]); }
        else q.push_back(c);
    }
    const u32 len1 = L, len2 = (u32)q.size();
    /* anchor: first position past the middle followed by 24 identical bases on both sides */
    u32 a1 = L / 2, a2 = 0;
    for (;; a1++) { a2 = qposOf[a1]; bool ok = true; for (int k = -12; k < 12 && ok; k++) ok = t[a1 + k] == q[a2 + k]; if (ok) break; }
    /* scoring classes: ACGT + everything else */
    lzb_scoring_dev sc; memset(&sc, 0, sizeof sc);
    sc.numClasses = 6; sc.gapOpen = 400; sc.gapExtend = 30;
    for (int b = 0; b < 256; b++) sc.cls[b] = 5;
    sc.cls[0] = 0; const char* acgt = "ACGT";
    for (int k = 0; k < 4; k++) { sc.cls[(u8)acgt[k]] = (u8)(1 + k); sc.cls[(u8)(acgt[k] | 32)] = (u8)(1 + k); }
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) {
        s32 v = -100; if (i == 0 || j == 0) v = -107374182; else if (i <= 4 && j <= 4) v = HOX[i - 1][j - 1];
        sc.subC[i * LZB_MAX_CLASSES + j] = sc.msubC[i * LZB_MAX_CLASSES + j] = v;
    }
    std::vector<u8> c1(len1 + 64, 0), c2(len2 + 64, 0);
    for (u32 i = 0; i < len1; i++) c1[i] = sc.cls[(u8)t[i]];
    for (u32 i = 0; i < len2; i++) c2[i] = sc.cls[(u8)q[i]];
    u8 *d1, *d2; lzb_scoring_dev* dsc;
    CK(cudaMalloc(&d1, c1.size())); CK(cudaMalloc(&d2, c2.size())); CK(cudaMalloc(&dsc, sizeof sc));
    CK(cudaMemcpy(d1, c1.data(), c1.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(d2, c2.data(), c2.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dsc, &sc, sizeof sc, cudaMemcpyHostToDevice));
    const u32 tbBytes = tbMiB << 20, tbLen = 1 + (tbBytes - 8);
    const int njobs = 2 * copies;
    std::vector<dp_job> hj(njobs);
    const u32 tbRowCap = tbLen / 24 + 4096;
    for (int k = 0; k < njobs; k++) {
        dp_job& J = hj[k]; memset(&J, 0, sizeof J);
        const int rev = (k & 1) == 0;
        J.reversed = rev; J.a1 = a1; J.a2 = a2;
        J.M = rev ? a1 + 1 : len1 - (a1 + 1); J.N = rev ? a2 + 1 : len2 - (a2 + 1);
        J.L0 = 0; J.R0 = (s32)(J.N + 1); J.leftSeg = { -1, -1 }; J.rightSeg = { -1, -1 }; J.listv = NULL; J.alignList = 0; J.al = NULL; J.resume = -1; J.token = 1;
        CK(cudaMalloc(&J.tb, (size_t)tbBytes + 64)); J.tbLen = tbLen;
        CK(cudaMalloc(&J.tbRow, (size_t)tbRowCap * 4)); J.tbRowCap = tbRowCap;
        J.opsCap = 1u << 18; CK(cudaMalloc(&J.ops, (size_t)J.opsCap * 4));
        J.actCap = 16; CK(cudaMalloc(&J.act, 16 * 5 * 4));
        if (ckpt) { CK(cudaMalloc(&J.ckpt, (size_t)1024 * CK_RECORD_WORDS * 4)); J.ckptCap = 1024; J.ckptEvery = 256; }
    }
    dp_job* dj; CK(cudaMalloc(&dj, njobs * sizeof(dp_job)));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    printf("pair %u x %u bp, anchor (%u,%u), traceback %u MiB, copies %d, checkpoints %s\n", len1, len2, a1, a2, tbMiB, copies, ckpt ? "every 256 rows" : "off");
    printf("%-12s %8s | %10s %10s %8s %8s %6s | %10s %10s\n", "kernel", "threads", "rows(rev)", "rows(fwd)", "cells/row", "ms", "status", "us/row x1", "us/row xN");
    for (int gN : { 148, 332, 592 }) if (gN <= njobs)
    for (auto& s : shapes) {
        double usrow[2] = { 0, 0 }; u32 rows[2] = { 0, 0 }; double ms1 = 0; int stt[2] = { 0, 0 }; double cpr = 0;
        for (int pass = 0; pass < 2; pass++) {
            const int g = pass == 0 ? 2 : gN;
            CK(cudaMemcpy(dj, hj.data(), njobs * sizeof(dp_job), cudaMemcpyHostToDevice));
            CK(cudaEventRecord(e0));
            s.launch(g, dj, d1, d2, len1, len2, dsc);
            CK(cudaEventRecord(e1));
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%-12s launch failed: %s\n", s.name, cudaGetErrorString(e)); break; }
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            std::vector<dp_job> out(njobs); CK(cudaMemcpy(out.data(), dj, njobs * sizeof(dp_job), cudaMemcpyDeviceToHost));
            u32 mr = out[0].rows > out[1].rows ? out[0].rows : out[1].rows;
            usrow[pass] = mr ? ms * 1e3 / mr : 0;
            if (pass == 0) { rows[0] = out[0].rows; rows[1] = out[1].rows; ms1 = ms; stt[0] = out[0].status; stt[1] = out[1].status; cpr = (double)(out[0].cells + out[1].cells) / (out[0].rows + out[1].rows + 1e-9); }
        }
        printf("%-12s %4d CTAs | %10u %10u %8.1f %8.2f %3d/%-3d | %10.3f %10.3f\n", s.name, gN, rows[0], rows[1], cpr, ms1, stt[0], stt[1], usrow[0], usrow[1]);
        fflush(stdout);
    }
    /* two grids on two streams, the way the anchor loop launches them: 86 sweeps, then 398 more */
    {
        cudaStream_t sa, sb; CK(cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking));
        for (int dyn : { 0, 47 * 1024 }) {
            CK(cudaFuncSetAttribute(k_ydrop_warp<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 47 * 1024));
            CK(cudaMemcpy(dj, hj.data(), njobs * sizeof(dp_job), cudaMemcpyHostToDevice));
            launch_list la = g_ll, lb = g_ll;
            for (int k = 0; k < LZB_LAUNCH_MAX - 86; k++) lb.ix[k] = (u16)(86 + k);
            cudaEvent_t ea0, ea1, eb1; CK(cudaEventCreate(&ea0)); CK(cudaEventCreate(&ea1)); CK(cudaEventCreate(&eb1));
            CK(cudaEventRecord(ea0, sa));
            k_ydrop_warp<16><<<86, 32, dyn, sa>>>(dj, la, (const dseg*)NULL, d1, d2, len1, len2, dsc, 9400, 1);
            CK(cudaEventRecord(ea1, sa));
            CK(cudaStreamWaitEvent(sb, ea0, 0));
            k_ydrop_warp<16><<<398, 32, dyn, sb>>>(dj, lb, (const dseg*)NULL, d1, d2, len1, len2, dsc, 9400, 1);
            CK(cudaEventRecord(eb1, sb));
            CK(cudaDeviceSynchronize());
            float ma, mb; CK(cudaEventElapsedTime(&ma, ea0, ea1)); CK(cudaEventElapsedTime(&mb, ea0, eb1));
            printf("two grids (86 + 398 one-warp sweeps, %d KB dynamic smem): first done after %.1f ms, second after %.1f ms\n", dyn / 1024, ma, mb);
        }
    }
    return 0;
}
