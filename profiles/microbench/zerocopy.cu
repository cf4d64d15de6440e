// Can SMs pull a batch out of pinned host memory (zero-copy over PCIe) as fast as the copy engine pushes it?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o zerocopy zerocopy.cu && ./zerocopy
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__global__ void pull(const float4 *__restrict__ src, float4 *__restrict__ dst, size_t n16, int unroll)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        float4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride),
               d = __ldcs(src + i + 3 * stride);
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < n16; i += stride) dst[i] = __ldcs(src + i);
}
// rows of `w` floats out of rows of `pitch` floats (trimmed 2-D pull), one warp per row piece
__global__ void pull_rows(const float *__restrict__ src, float *__restrict__ dst, int rows, int pitch, int w)
{
    const int warps = gridDim.x * (blockDim.x >> 5), wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    for (int r = wid; r < rows; r += warps) {
        const float4 *s = reinterpret_cast<const float4 *>(src + (size_t)r * pitch);
        float4 *d = reinterpret_cast<float4 *>(dst + (size_t)r * pitch);
        for (int i = lane; i < w / 4; i += 32) d[i] = __ldcs(s + i);
    }
}
int main()
{
    const size_t bytes = 256u << 20;
    float *h, *d;
    cudaHostAlloc(&h, bytes, cudaHostAllocMapped);
    cudaMalloc(&d, bytes);
    for (size_t i = 0; i < bytes / 4; i += 1024) h[i] = (float)i;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("copy engine, flat 256 MB: %.3f ms  %.1f GB/s\n", ms, bytes / ms / 1e6);
    for (int grid : {8, 16, 32, 64, 148, 296}) for (int threads : {256, 1024}) {
        pull<<<grid, threads>>>((const float4 *)h, (float4 *)d, bytes / 16, 4);
        cudaEventRecord(e0);
        pull<<<grid, threads>>>((const float4 *)h, (float4 *)d, bytes / 16, 4);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("SM pull  grid %3d x %4d: %.3f ms  %.1f GB/s\n", grid, threads, ms, bytes / ms / 1e6);
    }
    // y-like rows: 3488-byte pitch, 2300 bytes used
    const int pitch = 872, w = 576, rows = (int)(bytes / 4 / pitch);
    for (int grid : {32, 148}) {
        pull_rows<<<grid, 512>>>(h, d, rows, pitch, w);
        cudaEventRecord(e0);
        pull_rows<<<grid, 512>>>(h, d, rows, pitch, w);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("SM pull rows %d of %d floats, grid %3d: %.3f ms  %.1f GB/s (useful bytes)\n", w, pitch, grid, ms,
               (double)rows * w * 4 / ms / 1e6);
    }
    cudaEventRecord(e0);
    cudaMemcpy2DAsync(d, pitch * 4, h, pitch * 4, w * 4, rows, cudaMemcpyHostToDevice);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("copy engine 2D rows %d of %d floats: %.3f ms  %.1f GB/s (useful bytes)\n", w, pitch, ms, (double)rows * w * 4 / ms / 1e6);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
