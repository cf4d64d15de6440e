// fast3_phases.cu -- phase cycle accounting of mas_fast3_kernel (drop-in, skewed-lane DP) on one shape.
// Links against art_tts_b200/lib/libmas_sm100.so and calls the internal launcher with MasArgs::stats set.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o profiles/microbench/fast3_phases \
//        profiles/microbench/fast3_phases.cu -Lart_tts_b200/lib -lmas_sm100 -Xlinker -rpath=art_tts_b200/lib
//   ./fast3_phases B T_x T_y [ragged]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../../art_tts_b200/csrc/mas_internal.h"

using namespace mas;

int main(int argc, char **argv)
{
    const int B = argc > 1 ? atoi(argv[1]) : 16, T_x = argc > 2 ? atoi(argv[2]) : 190, T_y = argc > 3 ? atoi(argv[3]) : 872;
    const bool ragged = argc > 4 && atoi(argv[4]);
    const bool nopath = argc > 5 && atoi(argv[5]);   // durations only: no dense output, no zero fill
    std::vector<int32_t> tx(B, T_x), ty(B, std::min(T_y, 870));
    srand(1);
    if (ragged)
        for (int b = 1; b < B; ++b) {
            tx[b] = 60 + rand() % (T_x - 59);
            ty[b] = std::min(870, 4 * tx[b] + rand() % 100);
        }
    const size_t n = (size_t)B * T_x * T_y;
    std::vector<float> v(n);
    for (size_t i = 0; i < n; ++i) v[i] = -(50.0f + 100.0f * (rand() / (float)RAND_MAX));
    float *dv, *dp;
    int32_t *dtx, *dty, *ddur;
    long long *dstats;
    char *flush;
    cudaMalloc(&dv, n * 4);
    cudaMalloc(&dp, n * 4);
    cudaMalloc(&dtx, B * 4);
    cudaMalloc(&dty, B * 4);
    cudaMalloc(&ddur, (size_t)B * T_x * 4);
    cudaMalloc(&dstats, (size_t)B * 16 * 8);
    cudaMalloc(&flush, 256 << 20);
    cudaMemcpy(dv, v.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dtx, tx.data(), B * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dty, ty.data(), B * 4, cudaMemcpyHostToDevice);
    MasArgs a{};
    if (!fast3_layout(T_x, T_y, &a.lay)) { printf("shape not eligible\n"); return 1; }
    a.value = dv; a.t_x = dtx; a.t_y = dty; a.path = nopath ? nullptr : dp; a.durations = ddur;
    a.B = B; a.T_x = T_x; a.T_y = T_y; a.path_esize = 4; a.one = 0x3f800000ull;
    a.load_mode = (T_y % 4 == 0) ? 2 : 1; a.skewed = 1; a.dp_warps = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 4; ++rep) {
        const bool st = rep >= 2;
        a.stats = st ? dstats : nullptr;
        cudaMemset(dstats, 0, (size_t)B * 16 * 8);
        cudaMemset(flush, rep, 256 << 20);
        cudaEventRecord(e0);
        cudaError_t e = launch_fast3(a, MAS_F32, 0);
        cudaEventRecord(e1);
        cudaError_t e2 = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("rep %d stats=%d: %.4f ms (%s %s) smem %zu\n", rep, (int)st, ms, cudaGetErrorString(e), cudaGetErrorString(e2), a.lay.total);
    }
    {   // the lock-step kernel (mas_fast_kernel) on the same inputs, for comparison
        MasArgs o = a;
        o.stats = nullptr; o.skewed = 0;
        choose_plan(T_x, T_y, 0, &o.lay);
        for (int rep = 0; rep < 3; ++rep) {
            cudaMemset(flush, rep, 256 << 20);
            cudaEventRecord(e0);
            cudaError_t e = launch_fast(o, MAS_F32, 0);
            cudaEventRecord(e1);
            cudaError_t e2 = cudaDeviceSynchronize();
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("lock-step kernel rep %d: %.4f ms (%s %s) smem %zu\n", rep, ms, cudaGetErrorString(e), cudaGetErrorString(e2), o.lay.total);
        }
    }
    std::vector<long long> s((size_t)B * 16);
    cudaMemcpy(s.data(), dstats, s.size() * 8, cudaMemcpyDeviceToHost);
    const char *names[12] = {"forward", " starved", "backtrack", "t_y", "staging loop", " wait empty", "bulk wait", "ones+dur", "whole CTA", "stg issue", "stg arrive", "stg zero"};
    for (int b = 0; b < std::min(B, 4); ++b) {
        printf("utt %d:", b);
        for (int i = 0; i < 12; ++i) printf(" %s=%lld", names[i], s[b * 16 + i]);
        printf("\n");
    }
    double mean[12] = {0};
    for (int b = 0; b < B; ++b) for (int i = 0; i < 12; ++i) mean[i] += (double)s[b * 16 + i] / B;
    printf("mean:");
    for (int i = 0; i < 12; ++i) printf(" %s=%.0f", names[i], mean[i]);
    printf("\ncycles per frame (forward - starved) / t_y = %.1f ; starved per frame %.1f\n", (mean[0] - mean[1]) / mean[3], mean[1] / mean[3]);
    return 0;
}
