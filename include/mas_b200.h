/*
 * mas_b200.h -- C ABI of libmas_sm100.so: Monotonic Alignment Search for B200 (sm_100a).
 *
 * This is the drop-in boundary for ONE path of antoinelii/art-tts (paths relative to the
 * reference tree):
 *
 *   log_prior(mu_x, y) -> monotonic_align.maximum_path(value, mask) -> path -> durations
 *
 * Every entry point below names the reference interface it replaces.  Conventions:
 *   - all pointers are DEVICE pointers unless the name says `host`; the library never
 *     allocates, frees or keeps them; inputs are never modified (the reference clobbers
 *     its private host copy of `value`, core.pyx:30);
 *   - tensors are dense, C-contiguous, innermost axis = frames (T_y), exactly the
 *     memoryview layout of core.pyx:40 (`float[:,:,::1]`);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on
 *     it, nothing synchronises the host, every call is CUDA-graph capturable;
 *   - return value: 0 = ok, <0 = MAS_ERR_* (argument errors, detected on the host before
 *     any launch), >0 = cudaError_t of the failed launch.  Nothing throws across the ABI.
 */
#ifndef MAS_B200_H
#define MAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAS_ABI_VERSION 2

/* element types accepted for value / mask / path tensors */
enum {
    MAS_F32 = 0,
    MAS_F16 = 1,
    MAS_BF16 = 2,
    MAS_F64 = 3,
    MAS_I32 = 4,
    MAS_U8 = 5, /* also torch.bool */
    MAS_I64 = 6
};

enum {
    MAS_OK = 0,
    MAS_ERR_NULL = -1,      /* a required pointer is NULL                          */
    MAS_ERR_SHAPE = -2,     /* B, T_x, T_y or F out of range                       */
    MAS_ERR_DTYPE = -3,     /* unsupported element type for this argument          */
    MAS_ERR_WORKSPACE = -4, /* workspace missing or smaller than *_workspace_bytes */
    MAS_ERR_ALIGN = -5,     /* pointer not aligned to its element size             */
    MAS_ERR_NO_DEVICE = -6, /* no sm_100 device / driver available                 */
    MAS_ERR_PEER = -7       /* a mas_peer_gather was passed but this call would run an engine that does
                               not write peer memory (mas_peer_durations_supported), or the call does
                               not fit the peers' buffers                                           */
};

/* flags for mas_maximum_path / mas_from_prior_f32 */
enum {
    MAS_FLAG_NONE = 0,
    MAS_FLAG_FORCE_GENERAL = 1, /* use the size-agnostic kernel even when the fast one fits */
    MAS_FLAG_NO_ASYNC = 2,      /* fast kernel: stage tiles with LDG/STS instead of cp.async */
    MAS_FLAG_SPILL_BITS = 4,    /* keep the direction bits in the workspace even if they fit  */
    MAS_FLAG_HOST_NO_TRIM = 8,  /* mas_from_prior_host_f32: copy whole padded rows            */
    MAS_FLAG_NO_TENSOR = 16,    /* mas_from_prior_f32: fp32 FMA prior on CUDA cores instead of  */
                                /* the 3xTF32 tensor-core prior                                 */
    MAS_FLAG_FORCE_TENSOR = 32, /* tensor-core prior whenever the shape allows it (default: F >= 32) */
    MAS_FLAG_LOCKSTEP_DP = 64,  /* mas_maximum_path: the lock-step recurrence kernels (the default)      */
    MAS_FLAG_SKEWED_DP = 128    /* mas_maximum_path: the skewed-lane recurrence kernel (mas_fast3.cu,    */
                                /* T_x <= 256; same results, A/B measurements and tests)                 */
};
#define MAS_FLAG_TMA (1 << 16) /* mas_maximum_path: stage tiles with TMA tensor loads (cp.async.bulk.tensor.2d) instead of
                                  per-thread cp.async; fp32 value, rows 16-byte aligned, T_x <= 256; same results */
/* mas_maximum_path: keep ONE DP warp per utterance also when the batch fits the SMs in one wave (B <= SM count),
 * where the default splits the token axis of each utterance over two DP warps from 64 tokens on (latency of one
 * utterance instead of throughput).  Same results; A/B measurements and tests. */
#define MAS_FLAG_ONE_DP_WARP (1 << 18)
/* mas_from_prior_f32 (tensor-core engine; the others ignore it and clear the path themselves): `path` already holds
 * zeros in every element (the caller cleared it, or re-uses a buffer whose 1-cells it reset) -- the kernel only
 * writes the 1-cells.  Measured (profiles/zero_overlap.py): the kernel's own clearing hides in the slack of
 * its pipeline, so this saves nothing at B = 1024; it exists for callers that own pre-cleared buffers anyway. */
#define MAS_FLAG_PATH_ZEROED (1 << 19)
/* mas_from_prior_f32 (tensor-core engine): issue the MMAs as staggered half-chains of consecutive tiles instead of
 * tile pairs (same chains, same results; A/B measurements and tests). */
#define MAS_FLAG_STAGGER_MMA (1 << 20)
/* mas_from_prior_f32, 256 < T_x <= 512 (one thread-block cluster per utterance, the token axis split over its CTAs,
 * the recurrence crossing CTAs through distributed shared memory): clusters of 2 CTAs x 256 tokens instead of the
 * default 4 CTAs x 128 tokens.  Same results. */
#define MAS_FLAG_CLUSTER2 (1 << 17)
/* mas_from_prior_f32 (tensor-core engine): run at least k (1..255) utterances per persistent CTA, i.e. at
 * most ceil(B / k) CTAs.  For callers that keep several launches in flight (batch-sharded steps on several
 * streams): throughput instead of the latency of one call.  Bits 8..15 of `flags`. */
#define MAS_FLAG_UTT_PER_CTA(k) (((k) & 0xff) << 8)

int mas_abi_version(void);
const char *mas_strerror(int code);

/*
 * Lengths from a mask -- replaces monotonic_align/__init__.py:18-21
 *     t_x_max = mask.sum(1)[:, 0];  t_y_max = mask.sum(2)[:, 0]
 * Reads only column 0 and row 0 of each utterance's mask (element strides given in
 * elements, so views such as attn_mask.squeeze(1) need no copy).  Sums are truncated to
 * int32 like `.astype(np.int32)`.
 */
int mas_lengths_from_mask(const void *mask, int mask_dtype, int B, int T_x, int T_y,
                          int64_t stride_b, int64_t stride_x, int64_t stride_y,
                          int32_t *t_x_out, int32_t *t_y_out, void *stream);

/*
 * The same for the fused entry, whose caller holds the two sequence masks the reference multiplies
 * into attn_mask (tts.py:477-480): t_x = sum(x_mask[b]), t_y = sum(y_mask[b]) -- what
 * mask.sum(1)[:,0] / mask.sum(2)[:,0] give on their outer product.  Strides in elements
 * (batch, time), so [B,1,T] views need no copy.
 */
int mas_lengths_from_seq_masks(const void *x_mask, int x_dtype, int64_t x_stride_b, int64_t x_stride_t,
                               const void *y_mask, int y_dtype, int64_t y_stride_b, int64_t y_stride_t,
                               int B, int T_x, int T_y, int32_t *t_x_out, int32_t *t_y_out,
                               void *stream);

/*
 * Bytes of scratch mas_maximum_path / mas_from_prior_f32 need for this shape (packed
 * 1-bit-per-cell direction mask when it does not fit in shared memory, plus scheduling
 * state).  Never 0, so one allocation can be reused for every call of the same shape.
 */
size_t mas_workspace_bytes(int B, int T_x, int T_y);

/*
 * The drop-in kernel -- replaces maximum_path_c (core.pyx:38-45) together with the host
 * glue of maximum_path (monotonic_align/__init__.py:13-23).
 *
 *   value      [B,T_x,T_y]  value_dtype in {F32,F16,BF16,F64}; converted to fp32 on load
 *                           (== `.astype(np.float32)`, __init__.py:16)
 *   cell_mask  NULL, or [B,T_x,T_y] fp32 0/1: each value is multiplied by it on load
 *              (`value * mask`, __init__.py:13) -- only needed for masks that are not
 *              rectangular; for the sequence masks the reference builds (tts.py:477-480)
 *              the product is the identity on every cell the algorithm reads.
 *   t_x, t_y   [B] int32 lengths (from mas_lengths_from_mask).  t_x<=0 or t_y<=0: the
 *              utterance's path is all zeros.  t_x>t_y (the reference's degenerate case:
 *              empty band, backtrack over raw values) is reproduced exactly.
 *   path       [B,T_x,T_y]  path_dtype in {F32,F16,BF16,F64,I32,U8}; receives 0/1, padding
 *              included (== the zero-initialised int32 path cast to value.dtype, :17,:23)
 *   durations  NULL or [B,T_x] int32: sum_y path (tts.py:503-505), free by-product
 *   score      NULL or [B] fp32: V[t_x-1,t_y-1], the total alignment log-likelihood
 *
 * Arithmetic is the reference's: fp32 compare-select + one fp32 add per cell,
 * max_neg_val = -1e9f, strict `<` in the backtrack (ties stay on the current token).
 */
int mas_maximum_path(const void *value, int value_dtype, const float *cell_mask,
                     const int32_t *t_x, const int32_t *t_y, void *path, int path_dtype,
                     int32_t *durations, float *score, int B, int T_x, int T_y,
                     void *workspace, size_t workspace_bytes, int flags, void *stream);

/*
 * Fused Gaussian log-prior + MAS -- replaces the block GradTTS.compute_loss runs under
 * torch.no_grad() (tts.py:483-500; identical copies at :200-214, :776-790, :1067-1081)
 * plus the duration sum of tts.py:503-505.  The [T_x,T_y] fp32 score matrix never
 * exists in HBM.
 *
 *   mu_x  [B,F,T_x] fp32, y [B,F,T_y] fp32 (layouts of text_encoder.py:432 / the collate)
 *   logs  must be NULL: the reference has a unit-variance prior only
 *   lp[x,j] = -0.5*sum_f y^2 + sum_f mu*y - 0.5*sum_f mu^2 - 0.5*F*log(2*pi)   (fp32)
 *   path       NULL or [B,T_x,T_y] in path_dtype
 *   durations  NULL or [B,T_x] int32
 *   frame_idx  NULL or [B,T_y] int32: token index of every frame, -1 on padding
 *   score      NULL or [B] fp32
 *   log_prior_out  NULL or [B,T_x,T_y] fp32: parity tap -- the prior written by a separate
 *              kernel with the same operation order as the fused producers (bit-identical)
 */
int mas_from_prior_f32(const float *mu_x, const float *logs, const float *y,
                       const int32_t *t_x, const int32_t *t_y, void *path, int path_dtype,
                       int32_t *durations, int32_t *frame_idx, float *score,
                       float *log_prior_out, int B, int F, int T_x, int T_y,
                       void *workspace, size_t workspace_bytes, int flags, void *stream);

/*
 * Host-buffer variant of mas_from_prior_f32: the batch is still in (pinned) HOST memory, as
 * when the reference's training loop hands y to compute_loss (train_v2.py:203 ->
 * tts.py:466 relocate_input, which moves the whole padded batch).  Copies only what MAS can
 * touch -- each chunk of `chunk` consecutive utterances (0 = min(128, max(16, B/8))) is trimmed to its longest
 * utterance, so length-bucketed batches move ~35 % fewer bytes -- on an internal copy stream,
 * and launches the chunk's kernel on `stream` as soon as its rows have landed, so transfers
 * and kernels overlap.  Optionally copies durations / score back to host at the end.
 *   *_host     host pointers (pinned for asynchronous copies); t_x_host/t_y_host are READ by
 *              the CPU inside the call (chunk extents)
 *   *_dev      device staging of the full padded shapes [B,F,T_x], [B,F,T_y], [B], [B]
 *   h2d_bytes_out  NULL or receives the host->device bytes this call enqueued
 * Everything is enqueued asynchronously; the caller synchronises `stream` before reading the
 * host outputs.  Requires the fused plan (mas_from_prior_plan() == 0).
 */
int mas_from_prior_host_f32(const float *mu_x_host, const float *y_host, const int32_t *t_x_host,
                            const int32_t *t_y_host, float *mu_x_dev, float *y_dev,
                            int32_t *t_x_dev, int32_t *t_y_dev, void *path, int path_dtype,
                            int32_t *durations, int32_t *frame_idx, float *score,
                            int32_t *durations_host, float *score_host, int B, int F, int T_x,
                            int T_y, void *workspace, size_t workspace_bytes, int chunk, int flags,
                            void *stream, uint64_t *h2d_bytes_out);

/*
 * 0 when mas_from_prior_f32 runs fused for this shape; 1 when mu_x (F*T_x floats) does not
 * fit in shared memory next to the tile ring, in which case the prior is written once to
 * `log_prior_out` (then REQUIRED, used as scratch) and the drop-in kernel consumes it.
 */
int mas_from_prior_plan(int B, int F, int T_x, int T_y, int flags);

/*
 * Durations -> path -- replaces generate_path (src/model/utils.py:26-43 and
 * src/model_ms/utils.py:20-37) for rectangular masks:
 *   path[b,x,y] = (cum[x-1] <= y < cum[x]) && x < t_x[b] && y < t_y[b],  cum = cumsum(dur)
 * durations: [B,T_x], dur_dtype MAS_I32 or MAS_F32 (the reference passes the fp32 tensor
 * ceil(w)*length_scale, tts.py:132,146; fp32 sums run left to right like torch.cumsum on
 * the CPU and `y < cum` is evaluated as y < ceil(cum)).  t_x / t_y may be NULL (= full).
 * Also used to rebuild dense paths from all-gathered durations.
 */
int mas_generate_path(const void *durations, int dur_dtype, const int32_t *t_x,
                      const int32_t *t_y, void *path, int path_dtype, int B, int T_x, int T_y,
                      void *stream);

/* ------------------------------------------------------------------------------------------
 * Consumers of the alignment: what GradTTS.compute_loss does with `attn` right
 * after MAS, computed from the compact outputs (durations, frame index) instead of the dense
 * path.  A "segment" is the window of frames [offset[b], offset[b]+seg_len[b]) of utterance b,
 * written to output columns [0, seg_len[b]) of a [.., T_out] tensor; columns beyond are zero.
 * offset == NULL means 0, seg_len == NULL means T_out (both clamped to the tensor).
 * ------------------------------------------------------------------------------------------ */

/* durations [B,T_x] int32 -> frame_idx [B,T_y] int32 (token of every frame, -1 on padding):
 * the compact form of the path for callers of the drop-in mas_maximum_path. */
int mas_frame_index(const int32_t *durations, const int32_t *t_x, const int32_t *t_y,
                    int32_t *frame_idx, int B, int T_x, int T_y, void *stream);

/*
 * Duration targets and loss -- replaces tts.py:503-506 + duration_loss (model/utils.py:46-48):
 *     logw_ = log(1e-8 + sum_y attn) * x_mask;  loss = sum((logw - logw_)^2) / sum(x_lengths)
 *   logw         NULL or [B,T_x] fp32 (the duration predictor's output)
 *   logw_target  NULL or [B,T_x] fp32: receives logw_
 *   grad_unit    NULL or [B,T_x] fp32: d loss / d logw = 2 (logw - logw_) / sum(x_lengths)
 *   loss         NULL or [1] fp32; needs `workspace` of >= mas_align_workspace_bytes(B, 1, 1)
 *                bytes (8-byte aligned): per-block partial sums, reduced in a fixed order
 */
int mas_duration_loss_f32(const float *logw, const int32_t *durations, const int32_t *t_x,
                          float *logw_target, float *grad_unit, float *loss, int B, int T_x,
                          void *workspace, size_t workspace_bytes, void *stream);

/* out_size crop of a [B,R,T_y] fp32 tensor (y: R = n_feats) -- replaces the per-item slicing
 * loop of tts.py:524-544: dst[b,r,j] = src[b,r,offset[b]+j] for j < seg_len[b], else 0. */
int mas_crop_f32(const float *src, const int32_t *offset, const int32_t *seg_len, float *dst,
                 int B, int R, int T_y, int T_out, void *stream);

/* attn_cut of tts.py:524-544 straight from the frame index:
 * path[b,x,j] = (frame_idx[b,offset[b]+j] == x) for j < seg_len[b], else 0, in path_dtype. */
int mas_path_segment(const int32_t *frame_idx, const int32_t *offset, const int32_t *seg_len,
                     void *path, int path_dtype, int B, int T_x, int T_y, int T_out, void *stream);

/* scratch bytes of mas_align_gather_f32 (prior_loss != NULL) and mas_duration_loss_f32 (loss != NULL); never 0 */
size_t mas_align_workspace_bytes(int B, int F, int T_out);

/*
 * mu_y = (attn^T @ mu_x^T)^T -- replaces the one-hot GEMM of tts.py:552-555 by a gather:
 *     mu_y[b,f,j] = mu_x[b,f,frame_idx[b,offset[b]+j]]   (j < seg_len[b], else 0)
 * (bit-identical to the GEMM: 1.0*mu plus exact zeros), and optionally tts.py:562-563
 *     prior_loss[0] = sum(0.5 * ((y - mu_y)^2 + log(2 pi)) * y_mask) / (sum(y_mask) * F)
 *     prior_loss[1] = sum(y_mask) * F        (kept for the backward pass)
 *   y_seg  [B,F,T_out] fp32, required iff prior_loss != NULL (the cropped y)
 *   mu_y   NULL or [B,F,T_out] fp32
 */
int mas_align_gather_f32(const float *mu_x, const int32_t *frame_idx, const int32_t *offset,
                         const int32_t *seg_len, const float *y_seg, float *mu_y,
                         float *prior_loss, int B, int F, int T_x, int T_y, int T_out,
                         void *workspace, size_t workspace_bytes, void *stream);

/*
 * Backward of the above w.r.t. mu_x (autograd's attn @ grad_mu_y^T, tts.py:552-563):
 *     grad_mu_x[b,f,x] = sum_{j: frame_idx[b,offset+j]==x} ( grad_mu_y[b,f,j]
 *                          + grad_loss[0] / loss_norm[0] * (mu_x[b,f,x] - y_seg[b,f,j]) )
 * A token's frames are contiguous, so this is a segmented sum: deterministic, no atomics.
 * grad_mu_y may be NULL (no decoder gradient); the loss term needs y_seg, mu_x, grad_loss and
 * loss_norm (= prior_loss[1]) all non-NULL and is skipped otherwise.
 * frame_idx must be non-decreasing over each segment (always true for MAS output).
 */
int mas_align_gather_bwd_f32(const float *grad_mu_y, const float *y_seg, const float *mu_x,
                             const float *grad_loss, const float *loss_norm,
                             const int32_t *frame_idx, const int32_t *offset,
                             const int32_t *seg_len, float *grad_mu_x, int B, int F, int T_x,
                             int T_y, int T_out, void *stream);

/*
 * Which kernel a shape dispatches to (for tests, bench and DESIGN.md):
 * 0 = fast single-warp-DP kernel with bits in shared memory, 1 = fast kernel with the
 * direction mask spilled to the workspace, 2 = general kernel.
 */
int mas_plan(int B, int T_x, int T_y, int flags);

/*
 * The fused kernels are PERSISTENT (one CTA per SM for the whole batch).  A kernel that should run
 * concurrently -- the NCCL all-gather of the previous step's durations in data-parallel training
 *  -- finds no free SM until they finish.  n_sms > 0 makes them leave that many SMs
 * idle (default 0, or the MAS_RESERVE_SMS environment variable).  Process-wide setting.
 */
int mas_set_sm_reserve(int n_sms);

/*
 * Data-parallel training (train_v1_1_dist.py:249): the only exchange of the path is the
 * all-gather of the int32 durations (optionally the frame index).  With peer-mapped buffers (NVLink peer
 * memory: CUDA IPC / VMM handles, e.g. torch.distributed._symmetric_memory) the fused tensor-core kernel
 * does it itself: besides `durations`, its backtrack warp stores every utterance's row into row
 * `row0 + b` of EACH of the n_peers buffers (this rank's own among them), so that no collective kernel
 * runs at all; afterwards the ranks only need a barrier.
 *
 * The description travels WITH THE CALL (no process-wide state): any number of streams, devices and
 * buffer sets may be in flight, and a caller alternates two buffer sets (step parity) so that a fast rank's
 * next step never overwrites rows a slower rank is still reading.
 *   durations_ptrs[i]  base of rank i's [rows_total, row_stride] int32 buffer as mapped in this process
 *   row0               first row this call writes (rank * B for a batch-sharded step)
 *   rows               rows this call may write, counted from row0: B of the call (of the whole
 *                      host-buffer call for mas_from_prior_host_peer_f32); more -> MAS_ERR_PEER
 *   row_stride         elements between rows, >= T_x (one buffer serves every padded T_x up to it)
 *   frame_idx_ptrs     NULL, or [n_peers] bases of [rows_total, frame_idx_stride] int32 buffers that
 *                      receive the token index of every frame (-1 on padding), frame_idx_stride >= T_y
 * Only the tensor-core engine writes peer memory (mas_peer_durations_supported); every other plan
 * refuses the call with MAS_ERR_PEER rather than leave the buffers unwritten.  n_peers == 0 or a NULL
 * description = plain mas_from_prior_f32.  No reference equivalent (the reference runs one independent
 * MAS per rank and never exchanges durations).
 */
typedef struct mas_peer_gather {
    int n_peers;
    const uint64_t *durations_ptrs;
    int64_t row0, rows, row_stride;
    const uint64_t *frame_idx_ptrs;
    int64_t frame_idx_stride;
} mas_peer_gather;

int mas_from_prior_peer_f32(const float *mu_x, const float *logs, const float *y,
                            const int32_t *t_x, const int32_t *t_y, void *path, int path_dtype,
                            int32_t *durations, int32_t *frame_idx, float *score,
                            float *log_prior_out, int B, int F, int T_x, int T_y, void *workspace,
                            size_t workspace_bytes, int flags, void *stream,
                            const mas_peer_gather *peer);
int mas_from_prior_host_peer_f32(const float *mu_x_host, const float *y_host, const int32_t *t_x_host,
                                 const int32_t *t_y_host, float *mu_x_dev, float *y_dev,
                                 int32_t *t_x_dev, int32_t *t_y_dev, void *path, int path_dtype,
                                 int32_t *durations, int32_t *frame_idx, float *score,
                                 int32_t *durations_host, float *score_host, int B, int F, int T_x,
                                 int T_y, void *workspace, size_t workspace_bytes, int chunk,
                                 int flags, void *stream, uint64_t *h2d_bytes_out,
                                 const mas_peer_gather *peer);
int mas_peer_durations_supported(int B, int F, int T_x, int T_y, int flags);

/* Kernel launches enqueued by this library in this process (bench.py's gpu_launches). */
uint64_t mas_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MAS_B200_H */
