// Latency of the instruction chains the DP warps execute (one warp per SM, nothing else resident).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o dp_chain dp_chain.cu && ./dp_chain
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void k(float *out, long long *cyc, float a0, float b0)
{
    float a = a0 + threadIdx.x, b = b0, c = a0 * 0.5f;
    long long t0, t1;
    // 1: FMNMX -> FADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = fmaxf(a, b) + c;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // 2: FSETP -> FSEL -> FADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { float m = (b > a) ? b : a; a = m + c; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // 3: SHFL.UP chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = __shfl_up_sync(0xffffffffu, a, 1) + c;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // 4: FADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a + c;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // 5: the P = 2 recurrence, 2 tokens per lane + shuffle (as tdp_group)
    float V0 = a, V1 = b, left = c;
    unsigned acc0 = 0, acc1 = 0;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) {
        if (V0 > V1) acc1 |= 1u << (i & 7);
        float n1 = fmaxf(V0, V1) + c;
        float nx = __shfl_up_sync(0xffffffffu, n1, 1);
        if (left > V0) acc0 |= 1u << (i & 7);
        float n0 = fmaxf(left, V0) + c;
        V0 = n0; V1 = n1;
        left = threadIdx.x == 0 ? b0 : nx;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // 6: the same with the work of the fused kernel: prior adds (2 FADD per cell) as filler
    float q[8], d0[8], d1[8];
    for (int s = 0; s < 8; ++s) { q[s] = a0 * s; d0[s] = b0 * s + threadIdx.x; d1[s] = c * s; }
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N / 8; ++i) {
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const float v0 = (q[s] + d0[s]) + c, v1 = (q[s] + d1[s]) + b;
            if (V0 > V1) acc1 |= 1u << s;
            float n1 = fmaxf(V0, V1) + v1;
            float nx = __shfl_up_sync(0xffffffffu, n1, 1);
            if (left > V0) acc0 |= 1u << s;
            float n0 = fmaxf(left, V0) + v0;
            V0 = n0; V1 = n1;
            left = threadIdx.x == 0 ? b0 : nx;
            d0[s] = n0; d1[s] = n1;   // keeps the adds inside the loop
        }
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // 7: four independent FADD chains (issue rate of one warp)
    float e0 = a, e1 = b, e2 = c, e3 = a0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { e0 += c; e1 += c; e2 += c; e3 += c; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = t1 - t0;
    out[threadIdx.x] = a + V0 + V1 + (float)(acc0 + acc1) + e0 + e1 + e2 + e3;
}
int main()
{
    float *out; long long *cyc, h[7];
    cudaMalloc(&out, 128); cudaMalloc(&cyc, 56);
    for (int r = 0; r < 2; ++r) k<<<1, 32>>>(out, cyc, 1.0f, -3.0f);
    cudaMemcpy(h, cyc, 56, cudaMemcpyDeviceToHost);
    const char *n[7] = {"FMNMX+FADD", "FSETP+FSEL+FADD", "SHFL.UP+FADD", "FADD", "P=2 step (per frame)", "P=2 step + prior adds", "4 independent FADDs"};
    for (int i = 0; i < 7; ++i) printf("%-24s %.2f cycles/iteration\n", n[i], (double)h[i] / N);
    return 0;
}
