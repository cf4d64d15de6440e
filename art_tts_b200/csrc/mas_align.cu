// mas_align.cu -- the consumers that sit directly behind MAS in GradTTS/ArtTTS.compute_loss
// (SURVEY.md 8f), driven by the compact outputs of the MAS kernels (durations [B,T_x], frame
// index [B,T_y]) instead of the dense [B,T_x,T_y] path:
//
//   duration targets + duration loss      tts.py:503-506, model/utils.py:46-48
//   out_size crop of y / attn             tts.py:509-549
//   mu_y = attn^T @ mu_x^T (one-hot GEMM) tts.py:552-555   -> gather through the frame index
//   prior_loss                            tts.py:562-563
//   their backward passes w.r.t. mu_x / logw (autograd does them through the dense GEMM)
//
// A "segment" is the window of frames [off[b], off[b]+len[b]) of utterance b that lands in
// output columns [0, len[b]) of a [.., T_out] tensor; columns >= len[b] are zero.  Without a
// crop off == 0, len == y_lengths and T_out == T_y.
//
// All of this is HBM-bound elementwise / gather work: coalesced along the frame axis, one
// pass over the data, deterministic two-stage reductions (no atomics), fp32 arithmetic in the
// reference's operation order.
#include <algorithm>

#include "mas_common.cuh"
#include "mas_internal.h"

namespace mas {

namespace {

constexpr float kLog2Pi = 1.8378770664093453f;  // math.log(2 * math.pi) rounded to fp32
constexpr int kAlignThreads = 128;

__device__ __forceinline__ void segment_of(const int32_t *offset, const int32_t *seg_len, int b,
                                           int T_y, int T_out, int &off, int &len)
{
    off = offset ? offset[b] : 0;
    off = min(max(off, 0), T_y);
    len = seg_len ? seg_len[b] : T_out;
    len = min(min(max(len, 0), T_out), T_y - off);
}

// sum over the block, result valid in thread 0 (fixed tree: deterministic)
__device__ __forceinline__ float block_sum(float v, float *red)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(kFull, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float s = 0.0f;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    __syncthreads();
    return s;
}

__device__ __forceinline__ double block_sum_f64(double v, double *red)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(kFull, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    __syncthreads();
    return s;
}

}  // namespace

// ------------------------------------------------------------------------------------
// durations -> per-frame token index (the compact form of the path); -1 on padding.
// The fused kernel emits it directly; this serves callers of the drop-in maximum_path.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) frame_index_kernel(const int32_t *__restrict__ dur,
                                                          const int32_t *__restrict__ t_x,
                                                          const int32_t *__restrict__ t_y,
                                                          int32_t *__restrict__ fidx, int T_x, int T_y)
{
    extern __shared__ __align__(16) unsigned char smem[];
    int *start = reinterpret_cast<int *>(smem);  // [T_x + 1]
    const int b = blockIdx.x, tid = threadIdx.x;
    const int tx = t_x ? min(max(t_x[b], 0), T_x) : T_x;
    const int ty = t_y ? min(max(t_y[b], 0), T_y) : T_y;
    if (tid < 32) {
        const int32_t *d = dur + (int64_t)b * T_x;
        int carry = 0;
        for (int x0 = 0; x0 < T_x; x0 += 32) {
            const int x = x0 + tid;
            int v = (x < tx) ? max(d[x], 0) : 0;
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(kFull, v, o);
                if (tid >= o) v += n;
            }
            if (x < T_x) start[x + 1] = min(carry + v, ty);
            carry += __shfl_sync(kFull, v, 31);
        }
        if (tid == 0) start[0] = 0;
    }
    __syncthreads();
    int32_t *out = fidx + (int64_t)b * T_y;
    for (int y = tid; y < T_y; y += 256) out[y] = -1;
    __syncthreads();
    for (int x = tid; x < tx; x += 256)
        for (int y = start[x]; y < start[x + 1]; ++y) out[y] = x;
}

// ------------------------------------------------------------------------------------
// duration targets and loss: logw_ = log(1e-8 + dur) * x_mask; sum((logw - logw_)^2) / sum(len)
// Grid-stride over B*T_x elements, per-block partial sums (fp64) -> one finalize block: deterministic.
// ------------------------------------------------------------------------------------
constexpr int kDurBlocks = 148, kDurThreads = 256;

__global__ void __launch_bounds__(kDurThreads) duration_loss_kernel(const float *__restrict__ logw,
                                                                    const int32_t *__restrict__ dur,
                                                                    const int32_t *__restrict__ t_x,
                                                                    float *__restrict__ logw_target,
                                                                    float *__restrict__ grad_unit,
                                                                    double *__restrict__ partials, int B, int T_x)
{
    __shared__ double red[kDurThreads / 32];
    __shared__ float s_den;
    const int tid = threadIdx.x;
    double lsum = 0.0;   // every block recomputes sum(x_lengths): B integers
    for (int b = tid; b < B; b += kDurThreads) lsum += (double)min(max(t_x[b], 0), T_x);
    const double total_len = block_sum_f64(lsum, red);
    if (tid == 0) s_den = (float)total_len;  // torch.sum(lengths) used as an fp32 divisor
    __syncthreads();
    const float denom = s_den;
    double acc = 0.0;
    const int64_t n = (int64_t)B * T_x;
    for (int64_t i = (int64_t)blockIdx.x * kDurThreads + tid; i < n; i += (int64_t)gridDim.x * kDurThreads) {
        const int b = (int)(i / T_x), x = (int)(i - (int64_t)b * T_x);
        const float m = (x < t_x[b]) ? 1.0f : 0.0f;
        const float target = logf(1e-8f + (float)dur[i]) * m;  // tts.py:503-505
        if (logw_target) logw_target[i] = target;
        float d = 0.0f;
        if (logw) {
            d = logw[i] - target;
            acc += (double)(d * d);  // utils.py:47, squared in fp32 like the reference
        }
        if (grad_unit) grad_unit[i] = (2.0f * d) / denom;
    }
    const double s = block_sum_f64(acc, red);
    if (tid == 0 && partials) {
        partials[blockIdx.x] = s;
        if (blockIdx.x == 0) partials[gridDim.x] = (double)denom;
    }
}

__global__ void __launch_bounds__(32) duration_loss_finalize_kernel(const double *__restrict__ partials, int n,
                                                                    float *__restrict__ loss)
{
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) s += partials[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(kFull, s, o);
    if (threadIdx.x == 0) loss[0] = (float)s / (float)partials[n];
}

// ------------------------------------------------------------------------------------
// crop: y_seg[b,f,j] = y[b,f,off+j] (j < len) else 0;  path_seg[b,x,j] = (fidx[b,off+j] == x)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAlignThreads) crop_rows_kernel(const float *__restrict__ src,
                                                                  const int32_t *__restrict__ offset,
                                                                  const int32_t *__restrict__ seg_len,
                                                                  float *__restrict__ dst, int R,
                                                                  int T_y, int T_out)
{
    // grid (ceil(T_out/128), ceil(R/8), B): 8 rows of one utterance per block
    const int b = blockIdx.z, j = blockIdx.x * kAlignThreads + threadIdx.x;
    int off, len;
    segment_of(offset, seg_len, b, T_y, T_out, off, len);
    if (j >= T_out) return;
    const int r1 = min(R, (int)(blockIdx.y + 1) * 8);
    for (int r = blockIdx.y * 8; r < r1; ++r) {
        const int64_t row = (int64_t)b * R + r;
        dst[row * T_out + j] = (j < len) ? __ldg(src + row * T_y + off + j) : 0.0f;
    }
}

__global__ void __launch_bounds__(kAlignThreads) path_segment_kernel(const int32_t *__restrict__ fidx,
                                                                     const int32_t *__restrict__ offset,
                                                                     const int32_t *__restrict__ seg_len,
                                                                     void *__restrict__ path, int esize,
                                                                     unsigned long long one, int T_x,
                                                                     int T_y, int T_out)
{
    const int b = blockIdx.z, j = blockIdx.x * kAlignThreads + threadIdx.x;
    int off, len;
    segment_of(offset, seg_len, b, T_y, T_out, off, len);
    if (j >= T_out) return;
    const int xi = (j < len) ? fidx[(int64_t)b * T_y + off + j] : -1;
    const int x1 = min(T_x, (int)(blockIdx.y + 1) * 8);
    for (int x = blockIdx.y * 8; x < x1; ++x) {
        const int64_t e = ((int64_t)b * T_x + x) * T_out + j;
        const unsigned long long v = (x == xi) ? one : 0ull;
        switch (esize) {
        case 1: static_cast<uint8_t *>(path)[e] = (uint8_t)v; break;
        case 2: static_cast<uint16_t *>(path)[e] = (uint16_t)v; break;
        case 4: static_cast<uint32_t *>(path)[e] = (uint32_t)v; break;
        default: static_cast<uint64_t *>(path)[e] = v; break;
        }
    }
}

// ------------------------------------------------------------------------------------
// mu_y[b,f,j] = mu_x[b,f,fidx[b,off+j]]  (== attn^T @ mu_x^T for a one-hot attn, exactly:
// 1.0*m plus zeros), optionally with the prior-loss partial sums of the same elements.
// grid (ceil(T_out/128), ceil(F/fchunk), B); thread = one output column, loops its features.
// ------------------------------------------------------------------------------------
template <bool kLoss>
__global__ void __launch_bounds__(kAlignThreads) align_gather_kernel(
    const float *__restrict__ mu_x, const int32_t *__restrict__ fidx,
    const int32_t *__restrict__ offset, const int32_t *__restrict__ seg_len,
    const float *__restrict__ y_seg, float *__restrict__ mu_y, float *__restrict__ partials, int F,
    int T_x, int T_y, int T_out, int fchunk)
{
    __shared__ float red[kAlignThreads / 32];
    const int b = blockIdx.z, j = blockIdx.x * kAlignThreads + threadIdx.x;
    int off, len;
    segment_of(offset, seg_len, b, T_y, T_out, off, len);
    const bool in_seg = j < len;
    int x = in_seg ? fidx[(int64_t)b * T_y + off + j] : -1;
    if (x >= T_x) x = -1;
    const int f0 = blockIdx.y * fchunk, f1 = min(F, f0 + fchunk);
    float acc = 0.0f;
    if (j < T_out) {
        for (int f = f0; f < f1; ++f) {
            const int64_t row = (int64_t)b * F + f;
            const float m = (x >= 0) ? __ldg(mu_x + row * T_x + x) : 0.0f;
            if (mu_y) mu_y[row * T_out + j] = m;
            if (kLoss && in_seg) {
                const float d = __ldg(y_seg + row * T_out + j) - m;
                acc += 0.5f * (d * d + kLog2Pi);  // tts.py:562, elementwise in fp32
            }
        }
    }
    if (kLoss) {
        const float s = block_sum(acc, red);
        if (threadIdx.x == 0)
            partials[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
    }
}

// second stage: partial sums -> loss[0] = sum / (sum_b len_b * F), loss[1] = that normaliser
__global__ void __launch_bounds__(256) prior_loss_finalize_kernel(const float *__restrict__ partials,
                                                                  int64_t n,
                                                                  const int32_t *__restrict__ offset,
                                                                  const int32_t *__restrict__ seg_len,
                                                                  float *__restrict__ loss, int B, int F,
                                                                  int T_y, int T_out)
{
    __shared__ double red[8];
    double s = 0.0, l = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) s += (double)partials[i];
    for (int b = threadIdx.x; b < B; b += 256) {
        int off, len;
        segment_of(offset, seg_len, b, T_y, T_out, off, len);
        l += (double)len;
    }
    s = block_sum_f64(s, red);
    l = block_sum_f64(l, red);
    if (threadIdx.x == 0) {
        const float norm = (float)l * (float)F;  // torch.sum(y_mask) * n_feats
        loss[0] = (float)s / norm;
        loss[1] = norm;
    }
}

// ------------------------------------------------------------------------------------
// backward of the gather (+ of prior_loss): the frames of a token are contiguous, so
//   grad_mu_x[b,f,x] = sum_{j in seg(x)} ( g_mu_y[b,f,j] + c * (mu_x[b,f,x] - y_seg[b,f,j]) )
// with c = g_loss / (sum(y_mask) * F) is a plain segmented sum: no atomics, deterministic.
// grid (ceil(F/fchunk), B).  Each feature row of the gradient is read ONCE, coalesced, into shared
// memory (already combined: g - c*y); the token threads then add up their runs from there and
// write grad_mu_x coalesced along x.  Requires a non-decreasing frame index (what MAS produces).
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAlignThreads) align_gather_bwd_kernel(
    const float *__restrict__ g_mu_y, const float *__restrict__ y_seg, const float *__restrict__ mu_x,
    const float *__restrict__ g_loss, const float *__restrict__ loss_norm,
    const int32_t *__restrict__ fidx, const int32_t *__restrict__ offset,
    const int32_t *__restrict__ seg_len, float *__restrict__ g_mu_x, int F, int T_x, int T_y,
    int T_out, int fchunk)
{
    extern __shared__ __align__(16) unsigned char smem[];
    int *s_start = reinterpret_cast<int *>(smem);  // [T_x] first output column of the token, -1 = none
    int *s_end = s_start + T_x;                    // [T_x] one past its last column
    float *rowbuf = reinterpret_cast<float *>(s_end + T_x);   // [2][T_out] combined gradient rows
    const int b = blockIdx.y, tid = threadIdx.x;
    int off, len;
    segment_of(offset, seg_len, b, T_y, T_out, off, len);
    for (int x = tid; x < T_x; x += kAlignThreads) {
        s_start[x] = -1;
        s_end[x] = -1;
    }
    __syncthreads();
    const int32_t *ib = fidx + (int64_t)b * T_y + off;
    for (int j = tid; j < len; j += kAlignThreads) {
        const int x = ib[j];
        const int xp = (j > 0) ? ib[j - 1] : -2;
        if (x != xp) {
            if (x >= 0 && x < T_x) s_start[x] = j;
            if (xp >= 0 && xp < T_x) s_end[xp] = j;
        }
        if (j == len - 1 && x >= 0 && x < T_x) s_end[x] = len;
    }
    const bool with_loss = (y_seg != nullptr) && (g_loss != nullptr);
    const float c = with_loss ? g_loss[0] / loss_norm[0] : 0.0f;
    const int f0 = blockIdx.x * fchunk, f1 = min(F, f0 + fchunk);
    auto load_row = [&](int f, float *dst) {
        const int64_t row = (int64_t)b * F + f;
        for (int j = tid; j < len; j += kAlignThreads) {
            float t = g_mu_y ? __ldg(g_mu_y + row * T_out + j) : 0.0f;
            if (with_loss) t -= c * __ldg(y_seg + row * T_out + j);
            dst[j] = t;
        }
    };
    if (f0 < f1) load_row(f0, rowbuf);
    __syncthreads();
    for (int f = f0; f < f1; ++f) {
        float *cur = rowbuf + ((f - f0) & 1) * T_out, *nxt = rowbuf + (((f - f0) & 1) ^ 1) * T_out;
        if (f + 1 < f1) load_row(f + 1, nxt);          // next row in flight while this one is reduced
        const int64_t row = (int64_t)b * F + f;
        for (int x = tid; x < T_x; x += kAlignThreads) {
            const int s = s_start[x], e = s_end[x];
            float sum = 0.0f;
            if (s >= 0) {
                for (int j = s; j < e; ++j) sum += cur[j];
                if (with_loss) sum += c * __ldg(mu_x + row * T_x + x) * (float)(e - s);
            }
            g_mu_x[row * T_x + x] = sum;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------
// launches
// ------------------------------------------------------------------------------------
static int feature_chunk(int B, int F, int col_blocks)
{
    // enough blocks to fill 148 SMs a few times over, but at least 8 features per thread so the
    // frame index is amortised
    int fchunk = F;
    while (fchunk > 8 && (int64_t)B * col_blocks * ((F + fchunk - 1) / fchunk) < 148 * 8) fchunk = (fchunk + 1) / 2;
    return fchunk;
}

cudaError_t launch_frame_index(const int32_t *dur, const int32_t *t_x, const int32_t *t_y,
                               int32_t *fidx, int B, int T_x, int T_y, cudaStream_t st)
{
    const size_t smem = (size_t)(T_x + 1) * sizeof(int);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(frame_index_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    frame_index_kernel<<<B, 256, smem, st>>>(dur, t_x, t_y, fidx, T_x, T_y);
    count_launch();
    return cudaGetLastError();
}

size_t duration_loss_scratch_bytes() { return (size_t)(kDurBlocks + 1) * sizeof(double); }

cudaError_t launch_duration_loss(const float *logw, const int32_t *dur, const int32_t *t_x,
                                 float *logw_target, float *grad_unit, float *loss, double *partials,
                                 int B, int T_x, cudaStream_t st)
{
    const int64_t n = (int64_t)B * T_x;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(kDurBlocks, (n + kDurThreads - 1) / kDurThreads));
    duration_loss_kernel<<<blocks, kDurThreads, 0, st>>>(logw, dur, t_x, logw_target, grad_unit,
                                                         loss ? partials : nullptr, B, T_x);
    count_launch();
    if (loss) {
        duration_loss_finalize_kernel<<<1, 32, 0, st>>>(partials, blocks, loss);
        count_launch();
    }
    return cudaGetLastError();
}

cudaError_t launch_crop_rows(const float *src, const int32_t *offset, const int32_t *seg_len,
                             float *dst, int B, int R, int T_y, int T_out, cudaStream_t st)
{
    dim3 grid((T_out + kAlignThreads - 1) / kAlignThreads, (R + 7) / 8, B);
    crop_rows_kernel<<<grid, kAlignThreads, 0, st>>>(src, offset, seg_len, dst, R, T_y, T_out);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_path_segment(const int32_t *fidx, const int32_t *offset, const int32_t *seg_len,
                                void *path, int esize, unsigned long long one, int B, int T_x,
                                int T_y, int T_out, cudaStream_t st)
{
    dim3 grid((T_out + kAlignThreads - 1) / kAlignThreads, (T_x + 7) / 8, B);
    path_segment_kernel<<<grid, kAlignThreads, 0, st>>>(fidx, offset, seg_len, path, esize, one, T_x,
                                                        T_y, T_out);
    count_launch();
    return cudaGetLastError();
}

size_t align_partials(int B, int F, int T_out)
{
    const int cb = (T_out + kAlignThreads - 1) / kAlignThreads;
    const int fchunk = feature_chunk(B, F, cb);
    return (size_t)B * cb * ((F + fchunk - 1) / fchunk);
}

cudaError_t launch_align_gather(const float *mu_x, const int32_t *fidx, const int32_t *offset,
                                const int32_t *seg_len, const float *y_seg, float *mu_y,
                                float *loss, float *partials, int B, int F, int T_x, int T_y,
                                int T_out, cudaStream_t st)
{
    const int cb = (T_out + kAlignThreads - 1) / kAlignThreads;
    const int fchunk = feature_chunk(B, F, cb);
    dim3 grid(cb, (F + fchunk - 1) / fchunk, B);
    if (loss) {
        align_gather_kernel<true><<<grid, kAlignThreads, 0, st>>>(mu_x, fidx, offset, seg_len, y_seg,
                                                                  mu_y, partials, F, T_x, T_y, T_out,
                                                                  fchunk);
        prior_loss_finalize_kernel<<<1, 256, 0, st>>>(partials, (int64_t)grid.x * grid.y * grid.z,
                                                      offset, seg_len, loss, B, F, T_y, T_out);
        count_launch(2);
    } else {
        align_gather_kernel<false><<<grid, kAlignThreads, 0, st>>>(mu_x, fidx, offset, seg_len,
                                                                   nullptr, mu_y, nullptr, F, T_x,
                                                                   T_y, T_out, fchunk);
        count_launch();
    }
    return cudaGetLastError();
}

cudaError_t launch_align_gather_bwd(const float *g_mu_y, const float *y_seg, const float *mu_x,
                                    const float *g_loss, const float *loss_norm,
                                    const int32_t *fidx, const int32_t *offset,
                                    const int32_t *seg_len, float *g_mu_x, int B, int F, int T_x,
                                    int T_y, int T_out, cudaStream_t st)
{
    const size_t smem = (size_t)T_x * 2 * sizeof(int) + (size_t)T_out * 2 * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(align_gather_bwd_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int fchunk = F;
    while (fchunk > 4 && (int64_t)B * ((F + fchunk - 1) / fchunk) < 148 * 8) fchunk = (fchunk + 1) / 2;
    dim3 grid((F + fchunk - 1) / fchunk, B);
    align_gather_bwd_kernel<<<grid, kAlignThreads, smem, st>>>(g_mu_y, y_seg, mu_x, g_loss, loss_norm,
                                                               fidx, offset, seg_len, g_mu_x, F, T_x,
                                                               T_y, T_out, fchunk);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mas
