// Micro-benchmark of the fused kernel's inner product loop (one "pass" = 32 tokens x 32 frames x F):
// which formulation of the 4x8 register tile issues fastest on sm_100a?
//   v0  scalar FFMA, r-outer / k-inner (what prior_pass does)
//   v1  scalar FFMA, k-outer / r-inner
//   v2  packed fma.rn.f32x2 (two frames per instruction)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o fma_pass fma_pass.cu && ./fma_pass
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int F = 80, XR = 192, TY = 32;

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds128(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void ffma2(float &d0, float &d1, float a, float b0, float b1)
{
    // (d0,d1) = (a,a) * (b0,b1) + (d0,d1)
    asm volatile(
        "{\n\t.reg .b64 ra, rb, rc;\n\t"
        "mov.b64 ra, {%2, %2};\n\t"
        "mov.b64 rb, {%3, %4};\n\t"
        "mov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rc;\n\t}"
        : "+f"(d0), "+f"(d1)
        : "f"(a), "f"(b0), "f"(b1));
}

template <int V>
__device__ __forceinline__ float pass(const float *mu_s, const float *ys, int p, int lane)
{
    const int xg = lane >> 2, yg = lane & 3;
    uint32_t ma = s32(mu_s + 32 * p + 4 * xg), ya = s32(ys + 8 * yg);
    float acc[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[r][k] = 0.f;
#pragma unroll 4
    for (int f = 0; f < F; ++f) {
        const float4 m = lds128(ma), y0 = lds128(ya), y1 = lds128(ya + 16);
        ma += 4 * XR;
        ya += 4 * TY;
        const float mr[4] = {m.x, m.y, m.z, m.w};
        const float yk[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
        if (V == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[r][k] = __fmaf_rn(mr[r], yk[k], acc[r][k]);
        } else if (V == 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[r][k] = __fmaf_rn(mr[r], yk[k], acc[r][k]);
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 8; k += 2) ffma2(acc[r][k], acc[r][k + 1], mr[r], yk[k], yk[k + 1]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k) s += acc[r][k];
    return s;
}

// v3: packed FMA + explicit software pipelining (operands of feature f+1 are loaded before the
// FMAs of feature f are issued)
__device__ __forceinline__ float pass_v3(const float *mu_s, const float *ys, int p, int lane)
{
    const int xg = lane >> 2, yg = lane & 3;
    uint32_t ma = s32(mu_s + 32 * p + 4 * xg), ya = s32(ys + 8 * yg);
    float acc[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[r][k] = 0.f;
    float4 m = lds128(ma), y0 = lds128(ya), y1 = lds128(ya + 16);
#pragma unroll 4
    for (int f = 0; f < F; ++f) {
        ma += 4 * XR;
        ya += 4 * TY;
        const float4 mn = lds128(ma), y0n = lds128(ya), y1n = lds128(ya + 16);  // (reads one row past the end: harmless here)
        const float mr[4] = {m.x, m.y, m.z, m.w};
        const float yk[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 8; k += 2) ffma2(acc[r][k], acc[r][k + 1], mr[r], yk[k], yk[k + 1]);
        m = mn; y0 = y0n; y1 = y1n;
    }
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k) s += acc[r][k];
    return s;
}

// v4: 8 tokens x 8 frames per thread (warp = 64 tokens x 32 frames), packed FMA
__device__ __forceinline__ float pass_v4(const float *mu_s, const float *ys, int p, int lane)
{
    const int xg = lane >> 2, yg = lane & 3;
    uint32_t ma = s32(mu_s + 64 * (p % 3) + 8 * xg), ya = s32(ys + 8 * yg);
    float acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[r][k] = 0.f;
#pragma unroll 2
    for (int f = 0; f < F; ++f) {
        const float4 m0 = lds128(ma), m1 = lds128(ma + 16), y0 = lds128(ya), y1 = lds128(ya + 16);
        ma += 4 * XR;
        ya += 4 * TY;
        const float mr[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
        const float yk[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int k = 0; k < 8; k += 2) ffma2(acc[r][k], acc[r][k + 1], mr[r], yk[k], yk[k + 1]);
    }
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k) s += acc[r][k];
    return s;
}

template <int V>
__global__ void __launch_bounds__(512) bench(float *out, int iters, int nwarps)
{
    extern __shared__ __align__(16) float sm[];
    float *mu_s = sm, *ys = sm + F * XR;
    for (int i = threadIdx.x; i < F * XR + F * TY; i += blockDim.x) sm[i] = (float)((i * 2654435761u) >> 20) * 1e-4f;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= nwarps) return;
    float s = 0.f;
    for (int it = 0; it < iters; ++it) {
        if (V == 3) s += pass_v3(mu_s, ys, (warp + it) % 6, lane);
        else if (V == 4) s += pass_v4(mu_s, ys, (warp + it) % 6, lane);
        else s += pass<V>(mu_s, ys, (warp + it) % 6, lane);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V>
void run(const char *name, int nwarps)
{
    float *out;
    cudaMalloc(&out, 148 * 512 * 4);
    const int iters = 200;
    const size_t smem = (F * XR + F * TY) * 4 + 2048;   // slack: v3 prefetches one row past the end
    cudaFuncSetAttribute(bench<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    bench<V><<<148, 512, smem>>>(out, 10, nwarps);
    cudaEventRecord(e0);
    bench<V><<<148, 512, smem>>>(out, iters, nwarps);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = 148.0 * nwarps * iters * 32.0 * 32 * F * (V == 4 ? 2 : 1);   // FMAs
    const double cyc = ms * 1e-3 * 1.965e9;
    printf("%-28s warps=%2d  %.3f ms  %.1f FMA/clk/SM  (%.1f%% of 128)  %.2f TFLOP/s  err=%s\n", name, nwarps, ms,
           fma / 148 / cyc, 100 * fma / 148 / cyc / 128, 2 * fma / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main()
{
    for (int nw : {8, 12, 16}) {
        run<0>("v0 scalar r-outer", nw);
        run<1>("v1 scalar k-outer", nw);
        run<2>("v2 fma.rn.f32x2", nw);
        run<3>("v3 f32x2 + sw pipeline", nw);
        if (nw <= 8) run<4>("v4 f32x2 8x8 tile", nw);
    }
    return 0;
}
