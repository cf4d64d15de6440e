// mas_host.cu -- host-buffer entry point of the fused path: the batch still lives in (pinned)
// host memory, as it does when the reference's training loop hands y to compute_loss
// (train_v2.py:203 -> tts.py:466 relocate_input).  The reference moves the whole padded batch;
// here only the part of each utterance MAS can touch crosses PCIe, chunk by chunk, and the
// kernel of chunk k runs while chunk k+1 is still in flight:
//
//   copy stream : lengths | mu_x[c0] y[c0] | mu_x[c1] y[c1] | ...          (cudaMemcpy2DAsync,
//   main stream :                          | kernel(c0)     | kernel(c1) ... | D2H durations
//
// Batches are length-bucketed, so a chunk's rows are trimmed to the
// chunk's longest utterance (rounded up to the kernels' 32-frame tile).  Padding beyond that
// is never read by the kernels (mas_prior.cu loads whole tiles below t_y only).
#include <algorithm>
#include <map>
#include <mutex>
#include <utility>

#include "mas_internal.h"

namespace mas {

namespace {

constexpr int kMaxChunks = 64;

// one internal copy stream + events per (device, caller stream), created on first use: calls on
// different streams (or from different threads on different streams) never share a copy stream or
// its events, so the H2D copies of one call cannot overtake the kernels of another
struct HostState {
    cudaStream_t copy = nullptr;
    cudaEvent_t ready[kMaxChunks] = {};
    cudaEvent_t entry = nullptr;
};
std::mutex g_mu;
std::map<std::pair<int, cudaStream_t>, HostState *> g_state;

HostState *host_state(int dev, cudaStream_t main)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_state.find({dev, main});
    if (it != g_state.end()) return it->second;
    HostState *s = new HostState;
    bool ok = cudaStreamCreateWithFlags(&s->copy, cudaStreamNonBlocking) == cudaSuccess;
    for (auto &e : s->ready) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&s->entry, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        delete s;
        return nullptr;
    }
    g_state[{dev, main}] = s;
    return s;
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

}  // namespace mas

using namespace mas;

extern "C" int mas_from_prior_host_peer_f32(const float *mu_x_host, const float *y_host,
                                            const int32_t *t_x_host, const int32_t *t_y_host,
                                            float *mu_x_dev, float *y_dev, int32_t *t_x_dev,
                                            int32_t *t_y_dev, void *path, int path_dtype,
                                            int32_t *durations, int32_t *frame_idx, float *score,
                                            int32_t *durations_host, float *score_host, int B, int F,
                                            int T_x, int T_y, void *workspace, size_t workspace_bytes,
                                            int chunk, int flags, void *stream,
                                            uint64_t *h2d_bytes_out, const mas_peer_gather *peer)
{
    if (!mu_x_host || !y_host || !t_x_host || !t_y_host || !mu_x_dev || !y_dev || !t_x_dev || !t_y_dev)
        return MAS_ERR_NULL;
    if (durations_host && !durations) return MAS_ERR_NULL;
    if (score_host && !score) return MAS_ERR_NULL;
    if (B < 0 || F < 1 || T_x < 1 || T_y < 1) return MAS_ERR_SHAPE;
    if (B == 0) return MAS_OK;
    // default chunk: 128 utterances for a chip-filling batch (measured best at B=1024), but at least ~8
    // chunks per call so that a small length-bucketed shard (batch-sharded steps: 128 utterances) is still
    // trimmed chunk by chunk instead of as one block padded to its longest utterance
    if (chunk <= 0) chunk = std::min(128, std::max(16, (B + 7) / 8));
    if ((B + chunk - 1) / chunk > kMaxChunks) chunk = (B + kMaxChunks - 1) / kMaxChunks;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return MAS_ERR_NO_DEVICE;
    cudaStream_t main = static_cast<cudaStream_t>(stream);
    HostState *hs = host_state(dev, main);
    if (!hs) return MAS_ERR_NO_DEVICE;
    cudaStream_t copy = hs->copy;
    const bool no_trim = (flags & MAS_FLAG_HOST_NO_TRIM) != 0;
    cudaError_t e;
#define MAS_TRY(x)                  \
    do {                            \
        e = (x);                    \
        if (e != cudaSuccess) return (int)e; \
    } while (0)

    // the staging buffers may still be read by work already queued on `main`
    MAS_TRY(cudaEventRecord(hs->entry, main));
    MAS_TRY(cudaStreamWaitEvent(copy, hs->entry, 0));
    MAS_TRY(cudaMemcpyAsync(t_x_dev, t_x_host, (size_t)B * 4, cudaMemcpyHostToDevice, copy));
    MAS_TRY(cudaMemcpyAsync(t_y_dev, t_y_host, (size_t)B * 4, cudaMemcpyHostToDevice, copy));
    uint64_t moved = (uint64_t)B * 8;
    int ci = 0;
    for (int b0 = 0; b0 < B; b0 += chunk, ++ci) {
        const int nb = std::min(chunk, B - b0);
        int mx = 0, my = 0;
        for (int b = b0; b < b0 + nb; ++b) {
            mx = std::max(mx, std::min(std::max(t_x_host[b], 0), T_x));
            my = std::max(my, std::min(std::max(t_y_host[b], 0), T_y));
        }
        const int wx = no_trim ? T_x : std::min(T_x, round_up(std::max(mx, 1), 32));
        const int wy = no_trim ? T_y : std::min(T_y, round_up(std::max(my, 1), 32));
        const size_t rows = (size_t)nb * F;
        const size_t ox = (size_t)b0 * F * T_x, oy = (size_t)b0 * F * T_y;
        if (wx == T_x)
            MAS_TRY(cudaMemcpyAsync(mu_x_dev + ox, mu_x_host + ox, rows * T_x * 4, cudaMemcpyHostToDevice, copy));
        else
            MAS_TRY(cudaMemcpy2DAsync(mu_x_dev + ox, (size_t)T_x * 4, mu_x_host + ox, (size_t)T_x * 4,
                                      (size_t)wx * 4, rows, cudaMemcpyHostToDevice, copy));
        if (wy == T_y)
            MAS_TRY(cudaMemcpyAsync(y_dev + oy, y_host + oy, rows * T_y * 4, cudaMemcpyHostToDevice, copy));
        else
            MAS_TRY(cudaMemcpy2DAsync(y_dev + oy, (size_t)T_y * 4, y_host + oy, (size_t)T_y * 4,
                                      (size_t)wy * 4, rows, cudaMemcpyHostToDevice, copy));
        moved += (uint64_t)rows * (wx + wy) * 4;
        MAS_TRY(cudaEventRecord(hs->ready[ci], copy));
        MAS_TRY(cudaStreamWaitEvent(main, hs->ready[ci], 0));
        const size_t esz = (size_t)element_size(path_dtype);
        const int rc = from_prior_impl(
            mu_x_dev + ox, nullptr, y_dev + oy, t_x_dev + b0, t_y_dev + b0,
            path ? static_cast<char *>(path) + (size_t)b0 * T_x * T_y * esz : nullptr, path_dtype,
            durations ? durations + (size_t)b0 * T_x : nullptr,
            frame_idx ? frame_idx + (size_t)b0 * T_y : nullptr, score ? score + b0 : nullptr, nullptr,
            nb, F, T_x, T_y, workspace, workspace_bytes, flags & ~MAS_FLAG_HOST_NO_TRIM, main, peer,
            b0);   // b0: peer rows of this chunk
        if (rc != MAS_OK) return rc;
    }
    if (durations_host)
        MAS_TRY(cudaMemcpyAsync(durations_host, durations, (size_t)B * T_x * 4, cudaMemcpyDeviceToHost, main));
    if (score_host)
        MAS_TRY(cudaMemcpyAsync(score_host, score, (size_t)B * 4, cudaMemcpyDeviceToHost, main));
#undef MAS_TRY
    if (h2d_bytes_out) *h2d_bytes_out = moved;
    return MAS_OK;
}

extern "C" int mas_from_prior_host_f32(const float *mu_x_host, const float *y_host,
                                       const int32_t *t_x_host, const int32_t *t_y_host,
                                       float *mu_x_dev, float *y_dev, int32_t *t_x_dev,
                                       int32_t *t_y_dev, void *path, int path_dtype,
                                       int32_t *durations, int32_t *frame_idx, float *score,
                                       int32_t *durations_host, float *score_host, int B, int F,
                                       int T_x, int T_y, void *workspace, size_t workspace_bytes,
                                       int chunk, int flags, void *stream,
                                       uint64_t *h2d_bytes_out)
{
    return mas_from_prior_host_peer_f32(mu_x_host, y_host, t_x_host, t_y_host, mu_x_dev, y_dev, t_x_dev,
                                        t_y_dev, path, path_dtype, durations, frame_idx, score,
                                        durations_host, score_host, B, F, T_x, T_y, workspace,
                                        workspace_bytes, chunk, flags, stream, h2d_bytes_out, nullptr);
}
