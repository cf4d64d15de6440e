// mas_internal.h -- host-side declarations shared by the .cu files of libmas_sm100.so.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/mas_b200.h"

namespace mas {

constexpr int kMaxFastTx = 512;            // single-warp DP: 16 tokens per lane
constexpr int kSmemBudget = 227 * 1024;    // opt-in dynamic shared memory per CTA on sm_100
constexpr int kFastThreads = 160;          // 1 DP warp + 4 staging warps
constexpr int kFastZeroBytes = 2048;       // zeroed shared buffer behind the bulk (TMA) zero fill of the fast kernel
constexpr int kFast2Threads = 192;         // 2 DP warps + 4 staging warps (long utterances)
constexpr int kGeneralThreads = 256;

enum Plan { kPlanFastSmemBits = 0, kPlanFastSpillBits = 1, kPlanGeneral = 2 };

constexpr int kTmaBoxRows = 16;            // rows of one TMA box of the drop-in kernel ([16 rows][32 frames] = 2 KB)

// layout-compatible stand-in for CUtensorMap (cuda.h: 64-byte aligned, 16 x 64-bit opaque words), so that
// the kernels' headers do not need the driver API header
struct alignas(64) TensorMap {
    unsigned long long opaque[16];
};

struct FastLayout {       // dynamic shared memory carve-up of the fast kernels
    int xrows;            // 32 * ceil(T_x / 32)
    int srows;            // rows of one ring stage: xrows (+ kTmaBoxRows of slack for the last TMA box of a tile)
    int nch;              // ceil(T_y / 32) direction words per token
    int nstages;
    int bits_in_smem;
    size_t off_stages, off_bits, off_first, off_dur, off_bars, total;
};

struct MasArgs {
    const void *value;
    const float *cell_mask;
    const int32_t *t_x;
    const int32_t *t_y;
    void *path;
    int32_t *durations;
    float *score;
    int32_t *frame_idx;
    uint32_t *bits_ws;     // workspace: [B][nch][xrows] (fast, spilled) or [B][32*nch][xw] (general)
    int B, T_x, T_y;
    int path_esize;
    int dp_warps;          // 1, or 2 for long utterances (mas_fast2_kernel: two DP warps split the tokens)
    int skewed;            // 1: mas_fast3_kernel (skewed-lane recurrence, ring of mas_dp3.cuh)
    long long *stats;      // optional [B][16] phase cycle counters (profiles/microbench/fast3_phases.cu), else NULL
    int load_mode;         // fast kernel staging: 0 = LDG/STS (any dtype, cell mask), 1 = cp.async 4 B,
                           // 2 = cp.async 16 B (fp32, rows 16-byte aligned), 3 = TMA tensor boxes (same
                           // conditions; mas_fast_kernel only; rows in natural token order),
                           // 4 = cp.async 8 B (fp32, T_y even: rows 8-byte aligned)
    unsigned long long one;
    FastLayout lay;
};

int element_size(int dtype);
unsigned long long one_pattern(int dtype);
Plan choose_plan(int T_x, int T_y, int flags, FastLayout *lay, size_t extra_smem = 0,
                 int max_stages = 3, int row_align = 32, bool one_wave = false);

cudaError_t launch_lengths_from_mask(const void *mask, int mask_dtype, int B, int T_x, int T_y,
                                     int64_t sb, int64_t sx, int64_t sy, int32_t *t_x, int32_t *t_y,
                                     cudaStream_t st);
cudaError_t launch_seq_lengths(const void *xm, int xdt, int64_t xsb, int64_t xst, const void *ym, int ydt,
                               int64_t ysb, int64_t yst, int B, int T_x, int T_y, int32_t *t_x, int32_t *t_y,
                               cudaStream_t st);
cudaError_t launch_fast(const MasArgs &a, int value_dtype, cudaStream_t st, const TensorMap *tmap = nullptr);
cudaError_t launch_general(const MasArgs &a, int value_dtype, cudaStream_t st);
// mas_fast3.cu: drop-in kernel on the skewed-lane recurrence (T_x <= 256)
bool fast3_layout(int T_x, int T_y, FastLayout *lay);
cudaError_t launch_fast3(const MasArgs &a, int value_dtype, cudaStream_t st);
cudaError_t launch_generate_path(const void *dur, int dur_dtype, const int32_t *t_x,
                                 const int32_t *t_y, void *path, int esize, unsigned long long one,
                                 int B, int T_x, int T_y, cudaStream_t st);

struct PriorArgs {
    const float *mu_x;
    const float *y;
    const int32_t *t_x;
    const int32_t *t_y;
    void *path;
    int32_t *durations;
    int32_t *frame_idx;
    float *score;
    uint32_t *bits_ws;
    int bits_slots;        // direction-bit buffers in shared memory: 2, 1, or 0 (spilled to the workspace)
    int fma_per_smsp;      // FMA warps on each of schedulers 0-2 (1..4)
    int extra_fma;         // FMA warps sharing the DP warp's scheduler (0..2), see mas_prior.cu
    long long *stats;      // optional [grid][16] cycle counters (MAS_PRIOR_STATS=1), else NULL
    int B, F, T_x, T_y;
    int path_esize;
    unsigned long long one;
    FastLayout lay;
};
cudaError_t launch_from_prior(const PriorArgs &a, cudaStream_t st);
size_t prior_extra_smem(int F, int T_x, int T_y, bool second_bits);
cudaError_t launch_log_prior(const float *mu_x, const float *y, float *lp, int B, int F, int T_x,
                             int T_y, cudaStream_t st);

// mas_align.cu: consumers of the alignment (durations / frame index)
cudaError_t launch_frame_index(const int32_t *dur, const int32_t *t_x, const int32_t *t_y,
                               int32_t *fidx, int B, int T_x, int T_y, cudaStream_t st);
size_t duration_loss_scratch_bytes();
cudaError_t launch_duration_loss(const float *logw, const int32_t *dur, const int32_t *t_x,
                                 float *logw_target, float *grad_unit, float *loss, double *partials,
                                 int B, int T_x, cudaStream_t st);
cudaError_t launch_crop_rows(const float *src, const int32_t *offset, const int32_t *seg_len,
                             float *dst, int B, int R, int T_y, int T_out, cudaStream_t st);
cudaError_t launch_path_segment(const int32_t *fidx, const int32_t *offset, const int32_t *seg_len,
                                void *path, int esize, unsigned long long one, int B, int T_x,
                                int T_y, int T_out, cudaStream_t st);
size_t align_partials(int B, int F, int T_out);
cudaError_t launch_align_gather(const float *mu_x, const int32_t *fidx, const int32_t *offset,
                                const int32_t *seg_len, const float *y_seg, float *mu_y,
                                float *loss, float *partials, int B, int F, int T_x, int T_y,
                                int T_out, cudaStream_t st);
cudaError_t launch_align_gather_bwd(const float *g_mu_y, const float *y_seg, const float *mu_x,
                                    const float *g_loss, const float *loss_norm,
                                    const int32_t *fidx, const int32_t *offset,
                                    const int32_t *seg_len, float *g_mu_x, int B, int F, int T_x,
                                    int T_y, int T_out, cudaStream_t st);

// mas_prior_tc.cu: tensor-core variant of the fused kernel
struct TcLayout {
    int ok;                // 0: shape not eligible (use the CUDA-core kernel)
    int Fp;                // F rounded up to the MMA K step (8)
    int xrows, nch, nstages, bits_slots;
    int nb;                // accumulator buffers in TMEM
    int col_ahi, col_alo, col_d;   // TMEM column map
    int cluster;           // CTAs per utterance: 1, or a thread-block cluster of 2 / 4 splitting the token axis (T_x > 256)
    int xs;                // tokens per CTA (cluster mode: 256 / 128; else xrows)
    int xr_tot;            // cluster mode: row stride of the direction words in the workspace (T_x rounded up to 64)
    size_t off_stages, off_slabs, off_staging, off_bits, off_ysq, off_musq, off_first, off_dur, off_zero, off_bars, off_cl, total;
};
constexpr int kMaxPeers = 16;
struct PriorTcArgs {
    const float *mu_x;
    const float *y;
    const int32_t *t_x;
    const int32_t *t_y;
    void *path;
    int32_t *durations;
    int32_t *frame_idx;
    float *score;
    long long *stats;      // optional [grid][32] cycle counters (MAS_PRIOR_STATS=1), else NULL
    float *lp_out;         // optional parity tap [B,T_x,T_y]: the prior exactly as the DP consumed it
    uint32_t *bits_ws;     // cluster mode: direction words [B][nch][xr_tot] in the caller's workspace
    int32_t *peer[kMaxPeers];      // mas_peer_gather: every rank's durations buffer, or npeer == 0
    int32_t *peer_fi[kMaxPeers];   // and (optionally, else NULL) every rank's frame-index buffer
    int npeer;
    long long peer_row0, peer_stride, peer_fi_stride;
    int mma_stagger;       // MMA issue order: 1 = staggered half-chains of consecutive tiles (MAS_FLAG_STAGGER_MMA, opt-in), 0 = tile pairs (default)
    int path_zeroed;       // MAS_FLAG_PATH_ZEROED: `path` is all zeros already, only the 1-cells are written
    int utt_per_cta;       // MAS_FLAG_UTT_PER_CTA(k): at least k utterances per persistent CTA (0/1: one CTA per SM)
    int B, F, T_x, T_y;
    int path_esize;
    unsigned long long one;
    TcLayout lay;
};
TcLayout tc_layout(int F, int T_x, int T_y, int cluster = 4);
cudaError_t launch_from_prior_tc(const PriorTcArgs &a, cudaStream_t st);

int from_prior_impl(const float *mu_x, const float *logs, const float *y, const int32_t *t_x,
                    const int32_t *t_y, void *path, int path_dtype, int32_t *durations,
                    int32_t *frame_idx, float *score, float *log_prior_out, int B, int F, int T_x,
                    int T_y, void *workspace, size_t workspace_bytes, int flags, cudaStream_t st,
                    const mas_peer_gather *peer, long long row_extra);
int sm_count();
int sm_reserve();   // SMs the persistent kernels leave free (mas_set_sm_reserve)
void count_launch(int n = 1);

}  // namespace mas
