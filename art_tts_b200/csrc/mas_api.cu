// mas_api.cu -- the C ABI of libmas_sm100.so (see include/mas_b200.h for the contract and the
// reference interfaces each entry point replaces).  Host-side only: argument validation,
// plan selection, launches.  No allocation, no host synchronisation, no CPU fallback.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <cstdlib>
#include <initializer_list>

#include <cuda.h>   // CUtensorMap types / enums only: the encoder is fetched with cudaGetDriverEntryPoint, libcuda is not linked

#include "mas_internal.h"

namespace mas {

static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_sm_reserve{-1};
int sm_reserve()
{
    int v = g_sm_reserve.load(std::memory_order_relaxed);
    if (v < 0) {
        const char *e = std::getenv("MAS_RESERVE_SMS");
        v = (e && *e) ? std::atoi(e) : 0;
        if (v < 0) v = 0;
        g_sm_reserve.store(v, std::memory_order_relaxed);
    }
    return v;
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int element_size(int dtype)
{
    switch (dtype) {
    case MAS_F32: return 4;
    case MAS_F16: return 2;
    case MAS_BF16: return 2;
    case MAS_F64: return 8;
    case MAS_I32: return 4;
    case MAS_U8: return 1;
    case MAS_I64: return 8;
    default: return 0;
    }
}

unsigned long long one_pattern(int dtype)
{
    switch (dtype) {
    case MAS_F32: return 0x3f800000ull;
    case MAS_F16: return 0x3c00ull;
    case MAS_BF16: return 0x3f80ull;
    case MAS_F64: return 0x3ff0000000000000ull;
    default: return 1ull;
    }
}

static int env_int(const char *name, int dflt)
{
    const char *s = std::getenv(name);
    return (s && *s) ? std::atoi(s) : dflt;
}

// Tuning knobs (A/B measurements, profiling): the environment is read ONCE, at the first call into
// the library, so that plan queries (mas_plan, mas_from_prior_plan, mas_peer_durations_supported)
// and the launches that follow can never disagree.
struct Tuning {
    int prior_tc, prior_tc_min_f, stages, dp2_min_tx, dp2_min_tx_one_wave, prior_spill, prior_stats, fma_per_smsp, extra_fma, fast3, fast_tma, tc_cluster, tc_stagger;
};
static const Tuning &tuning()
{
    static const Tuning t = [] {
        Tuning v;
        v.prior_tc = env_int("MAS_PRIOR_TC", 1);
        v.prior_tc_min_f = env_int("MAS_PRIOR_TC_MIN_F", 32);
        v.stages = env_int("MAS_STAGES", 0);
        v.dp2_min_tx = env_int("MAS_DP2_MIN_TX", 256);
        v.dp2_min_tx_one_wave = env_int("MAS_DP2_MIN_TX_ONE_WAVE", 64);
        v.prior_spill = env_int("MAS_PRIOR_SPILL", 0);
        v.prior_stats = env_int("MAS_PRIOR_STATS", 0);
        v.fma_per_smsp = env_int("MAS_PRIOR_FMA_PER_SMSP", 2);
        if (v.fma_per_smsp < 1 || v.fma_per_smsp > 4) v.fma_per_smsp = 2;
        v.tc_stagger = env_int("MAS_TC_STAGGER", 0);   // MMA issue order of the tensor-core kernel: 1 = staggered half-chains (measured slower)
        v.tc_cluster = env_int("MAS_TC_CLUSTER", 4);   // CTAs per utterance of the tensor-core kernel when T_x > 256: 4 (default), 2, 0 = off
        v.fast_tma = env_int("MAS_FAST_TMA", 0);   // TMA tensor-load staging of the drop-in kernel: opt-in (measured: no faster)
        v.fast3 = env_int("MAS_FAST3", 0);   // skewed-lane drop-in kernel: opt-in until its HBM staging beats the lock-step one
        v.extra_fma = env_int("MAS_PRIOR_EXTRA_FMA", 1);
        if (v.extra_fma < 0 || v.extra_fma > 1) v.extra_fma = 1;
        return v;
    }();
    return t;
}

// Shared-memory carve-up of the fast kernels and the plan that follows from it.
// `extra_smem` = bytes the caller needs besides ring + bits (the fused kernel's operands).
Plan choose_plan(int T_x, int T_y, int flags, FastLayout *lay, size_t extra_smem, int max_stages, int row_align,
                 bool one_wave)
{
    FastLayout L{};
    L.xrows = ((T_x + row_align - 1) / row_align) * row_align;
    // single-DP-warp kernel: a stage has kTmaBoxRows of slack behind its rows (the last TMA box of a tile
    // may reach past the band); stage bases stay 1024-byte aligned (SWIZZLE_128B atom)
    // (only when TMA staging is asked for: the 2 KB per stage cost the 190 x 872 shape its third resident CTA)
    const bool tma_slack = row_align == 32 && extra_smem == 0 && (tuning().fast_tma || (flags & MAS_FLAG_TMA));
    L.srows = L.xrows + (tma_slack ? kTmaBoxRows : 0);
    L.nch = (T_y + 31) / 32;
    const size_t stage_bytes = (size_t)L.srows * 128;
    const size_t bits_bytes = (size_t)L.nch * L.xrows * 4;
    const size_t misc = (((size_t)T_x * 8 + 15) & ~(size_t)15) + 128 + kFastZeroBytes + extra_smem;
    Plan plan = kPlanGeneral;
    auto total = [&](int S, bool bits_smem) {
        return (size_t)S * stage_bytes + (bits_smem ? bits_bytes : 0) + misc;
    };
    if (T_x <= kMaxFastTx && !(flags & MAS_FLAG_FORCE_GENERAL)) {
        const bool fits = !(flags & MAS_FLAG_SPILL_BITS) && total(2, true) <= (size_t)kSmemBudget;
        L.bits_in_smem = fits ? 1 : 0;
        plan = fits ? kPlanFastSmemBits : kPlanFastSpillBits;
        if (total(2, fits) > (size_t)kSmemBudget) plan = kPlanGeneral;  // cannot happen for T_x<=512
        // ring depth: deepest ring that does not cost a resident CTA (or leaves >= 3 per SM)
        int S = 2;
        const int forced = tuning().stages;
        auto occ = [&](int s) { return (int)((size_t)(kSmemBudget + 1024) / (total(s, fits) + 1024)); };
        for (int s = 3; s <= max_stages; ++s)
            if (total(s, fits) <= (size_t)kSmemBudget && (one_wave || occ(s) >= occ(2) || occ(s) >= 3)) S = s;
        if (forced >= 2 && forced <= (extra_smem ? max_stages : 6) &&
            total(forced, fits) <= (size_t)kSmemBudget)
            S = forced;
        L.nstages = S;
    }
    if (plan == kPlanGeneral) {
        L.nstages = 0;
        L.bits_in_smem = 0;
    }
    L.off_stages = 0;
    L.off_bits = (size_t)L.nstages * stage_bytes;
    L.off_first = L.off_bits + (L.bits_in_smem ? bits_bytes : 0);
    L.off_dur = L.off_first + (size_t)T_x * 4;
    L.off_bars = (L.off_dur + (size_t)T_x * 4 + 15) & ~(size_t)15;
    L.total = L.off_bars + 128 + kFastZeroBytes + extra_smem;
    if (lay) *lay = L;
    return plan;
}

// Tensor map of `value` viewed as [B*T_x rows, T_y frames] fp32 for the TMA staging of mas_fast_kernel:
// boxes of [kTmaBoxRows rows][32 frames] (128-byte rows, SWIZZLE_128B = the XOR pattern of tile_index()).
// false when the driver entry point is missing or the encoder refuses the shape.
static bool encode_value_map(TensorMap *out, const void *base, int B, int T_x, int T_y)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeFn>(p);
    }();
    if (!fn) return false;
    static_assert(sizeof(TensorMap) == sizeof(CUtensorMap) && alignof(TensorMap) >= alignof(CUtensorMap), "CUtensorMap layout");
    const cuuint64_t dims[2] = {(cuuint64_t)T_y, (cuuint64_t)B * (cuuint64_t)T_x};
    const cuuint64_t strides[1] = {(cuuint64_t)T_y * 4};
    const cuuint32_t box[2] = {32u, (cuuint32_t)kTmaBoxRows};   // 32 frames = one tile = one 128-byte swizzle row
    const cuuint32_t estr[2] = {1, 1};
    return fn(reinterpret_cast<CUtensorMap *>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims,
              strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static size_t bits_workspace_bytes(int B, int T_x, int T_y)
{
    const size_t xrows = (size_t)((T_x + 63) / 64) * 64, nch = (size_t)(T_y + 31) / 32;   // 64: two-DP-warp layout
    return (size_t)B * nch * xrows * 4;
}

}  // namespace mas

using namespace mas;

extern "C" {

int mas_abi_version(void) { return MAS_ABI_VERSION; }

const char *mas_strerror(int code)
{
    switch (code) {
    case MAS_OK: return "ok";
    case MAS_ERR_NULL: return "required pointer is NULL";
    case MAS_ERR_SHAPE: return "shape out of range";
    case MAS_ERR_DTYPE: return "unsupported element type";
    case MAS_ERR_WORKSPACE: return "workspace missing or too small";
    case MAS_ERR_ALIGN: return "pointer not aligned to its element size";
    case MAS_ERR_NO_DEVICE: return "no usable CUDA device";
    case MAS_ERR_PEER: return "a peer gather was requested but this engine does not write peer memory, or the call does not fit the peer buffers";
    default: break;
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "unknown error";
}

uint64_t mas_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

static bool tc_engine_selected(int F, int T_x, int T_y, int flags, mas::TcLayout *lay)
{
    if (!tuning().prior_tc || (flags & (MAS_FLAG_FORCE_GENERAL | MAS_FLAG_SPILL_BITS | MAS_FLAG_NO_TENSOR))) return false;
    if (T_x > 256 && tuning().tc_cluster != 2 && tuning().tc_cluster != 4) return false;
    const mas::TcLayout l = mas::tc_layout(F, T_x, T_y, (flags & MAS_FLAG_CLUSTER2) ? 2 : tuning().tc_cluster);
    if (lay) *lay = l;
    // measured (profiles/config_sweep.py): below ~32 features the FMA work is so small that the
    // CUDA-core kernel is as fast or faster (F=16: 0.157 vs 0.166 ms at B=1024, 160x512)
    return l.ok && (F >= tuning().prior_tc_min_f || (flags & MAS_FLAG_FORCE_TENSOR));
}

int mas_peer_durations_supported(int B, int F, int T_x, int T_y, int flags)
{
    (void)B;
    if (T_x < 1 || T_y < 1 || F < 1) return 0;
    return tc_engine_selected(F, T_x, T_y, flags, nullptr) ? 1 : 0;
}

// validate a per-call peer description against the call it accompanies
static int check_peer(const mas_peer_gather *peer, int B, int T_x, int T_y, long long row_extra)
{
    if (!peer || peer->n_peers == 0) return MAS_OK;
    if (peer->n_peers < 0 || peer->n_peers > mas::kMaxPeers || !peer->durations_ptrs) return MAS_ERR_SHAPE;
    if (peer->row0 < 0 || peer->rows < 1 || peer->row_stride < T_x) return MAS_ERR_PEER;
    if (row_extra + B > peer->rows) return MAS_ERR_PEER;   // would write outside the peers' buffers
    if (peer->frame_idx_ptrs && peer->frame_idx_stride < T_y) return MAS_ERR_PEER;
    for (int i = 0; i < peer->n_peers; ++i) {
        if (!peer->durations_ptrs[i] || peer->durations_ptrs[i] % 4) return MAS_ERR_ALIGN;
        if (peer->frame_idx_ptrs && (!peer->frame_idx_ptrs[i] || peer->frame_idx_ptrs[i] % 4)) return MAS_ERR_ALIGN;
    }
    return MAS_OK;
}

int mas_set_sm_reserve(int n_sms)
{
    if (n_sms < 0 || n_sms > 1024) return MAS_ERR_SHAPE;
    g_sm_reserve.store(n_sms, std::memory_order_relaxed);
    return MAS_OK;
}

size_t mas_workspace_bytes(int B, int T_x, int T_y)
{
    if (B < 0 || T_x < 1 || T_y < 1) return 256;
    return bits_workspace_bytes(B, T_x, T_y) + 256;
}

static Plan plan_fast(int B, int T_x, int T_y, int flags, MasArgs *a);

int mas_plan(int B, int T_x, int T_y, int flags)
{
    (void)B;
    if (T_x < 1 || T_y < 1) return MAS_ERR_SHAPE;
    MasArgs a{};
    return (int)plan_fast(B, T_x, T_y, flags, &a);
}

// Layout + DP-warp count of the drop-in kernel.  Long token axis (config 4: 512 x 4096): the
// frame-sequential recurrence is the whole run time, so two DP warps split the tokens (mas_dp.cuh
// dp_forward2); that needs a third ring stage and 64-row alignment of the tiles.
// A batch that fits the SMs in one wave (B <= SM count: BASELINE configs 1-3) is bound by the latency of ONE
// utterance's recurrence, not by HBM or occupancy: two DP warps from 64 tokens on.
static Plan plan_fast(int B, int T_x, int T_y, int flags, MasArgs *a)
{
    a->skewed = 0;
    if ((tuning().fast3 || (flags & MAS_FLAG_SKEWED_DP)) &&
        !(flags & (MAS_FLAG_FORCE_GENERAL | MAS_FLAG_SPILL_BITS | MAS_FLAG_LOCKSTEP_DP)) &&
        fast3_layout(T_x, T_y, &a->lay)) {
        a->skewed = 1;          // skewed-lane recurrence (mas_fast3.cu), direction bits in shared memory
        a->dp_warps = 1;
        return kPlanFastSmemBits;
    }
    Plan plan = choose_plan(T_x, T_y, flags, &a->lay);
    a->dp_warps = 1;
    const bool one_wave = B >= 1 && B <= sm_count() && !(flags & (MAS_FLAG_ONE_DP_WARP | MAS_FLAG_TMA));
    if (plan != kPlanGeneral && (T_x > tuning().dp2_min_tx || (one_wave && T_x > tuning().dp2_min_tx_one_wave))) {
        FastLayout l2;
        const Plan p2 = choose_plan(T_x, T_y, flags, &l2, 512, 3, 64, one_wave);
        if (p2 != kPlanGeneral && l2.nstages >= 3) {
            a->lay = l2;
            plan = p2;
            a->dp_warps = 2;
        }
    }
    return plan;
}

static bool shape_ok(int B, int T_x, int T_y)
{
    if (B < 0 || T_x < 1 || T_y < 1) return false;
    const double cells = (double)B * T_x * T_y;
    return cells < 9.0e18 / 8.0 && T_x <= (1 << 24) && T_y <= (1 << 24);
}

int mas_lengths_from_mask(const void *mask, int mask_dtype, int B, int T_x, int T_y,
                          int64_t stride_b, int64_t stride_x, int64_t stride_y, int32_t *t_x_out,
                          int32_t *t_y_out, void *stream)
{
    if (!mask || !t_x_out || !t_y_out) return MAS_ERR_NULL;
    if (!shape_ok(B, T_x, T_y)) return MAS_ERR_SHAPE;
    if (element_size(mask_dtype) == 0) return MAS_ERR_DTYPE;
    if (B == 0) return MAS_OK;
    return (int)launch_lengths_from_mask(mask, mask_dtype, B, T_x, T_y, stride_b, stride_x, stride_y,
                                         t_x_out, t_y_out, static_cast<cudaStream_t>(stream));
}

int mas_lengths_from_seq_masks(const void *x_mask, int x_dtype, int64_t x_stride_b, int64_t x_stride_t,
                               const void *y_mask, int y_dtype, int64_t y_stride_b, int64_t y_stride_t,
                               int B, int T_x, int T_y, int32_t *t_x_out, int32_t *t_y_out, void *stream)
{
    if (!x_mask || !y_mask || !t_x_out || !t_y_out) return MAS_ERR_NULL;
    if (!shape_ok(B, T_x, T_y)) return MAS_ERR_SHAPE;
    if (element_size(x_dtype) == 0 || element_size(y_dtype) == 0) return MAS_ERR_DTYPE;
    if (B == 0) return MAS_OK;
    return (int)launch_seq_lengths(x_mask, x_dtype, x_stride_b, x_stride_t, y_mask, y_dtype, y_stride_b,
                                   y_stride_t, B, T_x, T_y, t_x_out, t_y_out,
                                   static_cast<cudaStream_t>(stream));
}

int mas_maximum_path(const void *value, int value_dtype, const float *cell_mask, const int32_t *t_x,
                     const int32_t *t_y, void *path, int path_dtype, int32_t *durations,
                     float *score, int B, int T_x, int T_y, void *workspace, size_t workspace_bytes,
                     int flags, void *stream)
{
    if (!value || !t_x || !t_y) return MAS_ERR_NULL;
    if (!path && !durations && !score) return MAS_ERR_NULL;
    if (!shape_ok(B, T_x, T_y)) return MAS_ERR_SHAPE;
    if (value_dtype != MAS_F32 && value_dtype != MAS_F16 && value_dtype != MAS_BF16 &&
        value_dtype != MAS_F64)
        return MAS_ERR_DTYPE;
    const int esize = element_size(path_dtype);
    if (path && (esize == 0 || path_dtype == MAS_I64)) return MAS_ERR_DTYPE;
    if ((uintptr_t)value % element_size(value_dtype) || (path && (uintptr_t)path % esize) ||
        (uintptr_t)t_x % 4 || (uintptr_t)t_y % 4 || (uintptr_t)durations % 4 ||
        (uintptr_t)score % 4 || (uintptr_t)cell_mask % 4)
        return MAS_ERR_ALIGN;
    if (B == 0) return MAS_OK;

    MasArgs a{};
    const Plan plan = plan_fast(B, T_x, T_y, flags, &a);
    if (plan != kPlanFastSmemBits) {
        if (!workspace || workspace_bytes < mas_workspace_bytes(B, T_x, T_y)) return MAS_ERR_WORKSPACE;
        if ((uintptr_t)workspace % 16) return MAS_ERR_ALIGN;
    }
    if (plan == kPlanGeneral && (size_t)T_x * 16 > (size_t)kSmemBudget) return MAS_ERR_SHAPE;
    a.value = value;
    a.cell_mask = cell_mask;
    a.t_x = t_x;
    a.t_y = t_y;
    a.path = path;
    a.durations = durations;
    a.score = score;
    a.bits_ws = static_cast<uint32_t *>(workspace);
    a.B = B;
    a.T_x = T_x;
    a.T_y = T_y;
    a.path_esize = path ? esize : 4;
    a.one = one_pattern(path_dtype);
    a.load_mode = 0;
    if (value_dtype == MAS_F32 && !cell_mask && !(flags & MAS_FLAG_NO_ASYNC))
        a.load_mode = (T_y % 4 == 0 && (uintptr_t)value % 16 == 0) ? 2
                      : (T_y % 2 == 0 && (uintptr_t)value % 8 == 0 && !a.skewed) ? 4 : 1;
    // TMA tensor loads for the single-DP-warp kernel (same conditions as the 16-byte cp.async path)
    TensorMap tmap{};
    if (a.load_mode == 2 && plan != kPlanGeneral && !a.skewed && a.dp_warps == 1 && a.lay.srows >= a.lay.xrows + kTmaBoxRows &&
        (tuning().fast_tma || (flags & MAS_FLAG_TMA)) && encode_value_map(&tmap, value, B, T_x, T_y))
        a.load_mode = 3;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const cudaError_t e = (plan == kPlanGeneral) ? launch_general(a, value_dtype, st)
                          : a.skewed           ? launch_fast3(a, value_dtype, st)
                                               : launch_fast(a, value_dtype, st, a.load_mode == 3 ? &tmap : nullptr);
    return (int)e;
}

int mas_from_prior_plan(int B, int F, int T_x, int T_y, int flags)
{
    (void)B;
    if (T_x < 1 || T_y < 1 || F < 1) return MAS_ERR_SHAPE;
    if (tc_engine_selected(F, T_x, T_y, flags, nullptr)) return 0;   // tensor-core engine (cluster of CTAs when T_x > 256)
    FastLayout lay;
    const Plan plan = choose_plan(T_x, T_y, flags, &lay, prior_extra_smem(F, T_x, T_y, false), 3);
    return plan == kPlanGeneral ? 1 : 0;
}

}  // extern "C"

namespace mas {

// mas_from_prior_f32 / mas_from_prior_peer_f32 / one chunk of the host-buffer entries.  `row_extra` =
// first utterance of this launch within the caller's batch (chunked host entry): added to peer->row0.
int from_prior_impl(const float *mu_x, const float *logs, const float *y, const int32_t *t_x,
                    const int32_t *t_y, void *path, int path_dtype, int32_t *durations,
                    int32_t *frame_idx, float *score, float *log_prior_out, int B, int F, int T_x,
                    int T_y, void *workspace, size_t workspace_bytes, int flags, cudaStream_t st,
                    const mas_peer_gather *peer, long long row_extra)
{
    if (!mu_x || !y || !t_x || !t_y) return MAS_ERR_NULL;
    if (!path && !durations && !frame_idx && !score) return MAS_ERR_NULL;
    if (logs) return MAS_ERR_DTYPE;  // unit-variance prior only (tts.py:483-495 has no logs term)
    if (!shape_ok(B, T_x, T_y) || F < 1 || F > 4096) return MAS_ERR_SHAPE;
    const int esize = element_size(path_dtype);
    if (path && (esize == 0 || path_dtype == MAS_I64)) return MAS_ERR_DTYPE;
    if ((uintptr_t)mu_x % 4 || (uintptr_t)y % 4 || (path && (uintptr_t)path % esize) ||
        (uintptr_t)t_x % 4 || (uintptr_t)t_y % 4 || (uintptr_t)durations % 4 ||
        (uintptr_t)frame_idx % 4 || (uintptr_t)score % 4 || (uintptr_t)log_prior_out % 4)
        return MAS_ERR_ALIGN;
    const int prc = check_peer(peer, B, T_x, T_y, row_extra);
    if (prc != MAS_OK) return prc;
    const bool want_peer = peer && peer->n_peers > 0;
    if (B == 0) return MAS_OK;

    // tensor-core prior (mas_prior_tc.cu) for every shape it covers; MAS_PRIOR_TC=0 forces the
    // CUDA-core kernel (A/B measurements, and the shapes beyond 256 tokens / 96 features)
    PriorTcArgs t{};
    if (tc_engine_selected(F, T_x, T_y, flags, &t.lay)) {
        t.mu_x = mu_x;
        t.y = y;
        t.t_x = t_x;
        t.t_y = t_y;
        t.path = path;
        t.durations = durations;
        t.frame_idx = frame_idx;
        t.score = score;
        t.lp_out = log_prior_out;
        t.bits_ws = nullptr;
        if (t.lay.cluster > 1) {   // direction words of the cluster kernel live in the workspace
            if (!workspace || workspace_bytes < mas_workspace_bytes(B, T_x, T_y)) return MAS_ERR_WORKSPACE;
            if ((uintptr_t)workspace % 16) return MAS_ERR_ALIGN;
            t.bits_ws = static_cast<uint32_t *>(workspace);
        }
        t.stats = nullptr;
        if (tuning().prior_stats) {  // profiling aid: counters at the tail of an over-sized workspace
            const size_t need = mas_workspace_bytes(B, T_x, T_y), sbytes = (size_t)1024 * 32 * 8;
            if (workspace && workspace_bytes >= need + sbytes + 16)
                t.stats = reinterpret_cast<long long *>(static_cast<char *>(workspace) +
                                                        ((workspace_bytes - sbytes) & ~(size_t)15));
        }
        t.B = B;
        t.F = F;
        t.T_x = T_x;
        t.T_y = T_y;
        t.path_esize = path ? esize : 4;
        t.one = one_pattern(path_dtype);
        t.utt_per_cta = (flags >> 8) & 0xff;
        t.path_zeroed = (flags & MAS_FLAG_PATH_ZEROED) ? 1 : 0;
        t.mma_stagger = (tuning().tc_stagger || (flags & MAS_FLAG_STAGGER_MMA)) ? 1 : 0;
        t.npeer = 0;
        if (want_peer) {   // fused all-gather of the durations (and frame index) over peer memory
            t.npeer = peer->n_peers;
            for (int i = 0; i < peer->n_peers; ++i) {
                t.peer[i] = reinterpret_cast<int32_t *>(peer->durations_ptrs[i]);
                t.peer_fi[i] = peer->frame_idx_ptrs ? reinterpret_cast<int32_t *>(peer->frame_idx_ptrs[i]) : nullptr;
            }
            t.peer_row0 = peer->row0 + row_extra;
            t.peer_stride = peer->row_stride;
            t.peer_fi_stride = peer->frame_idx_ptrs ? peer->frame_idx_stride : 0;
        }
        return (int)launch_from_prior_tc(t, st);
    }

    // fail loudly rather than leave the peers' buffers unwritten
    if (want_peer) return MAS_ERR_PEER;
    PriorArgs a{};
    // prefer two direction-bit buffers in shared memory (backtrack of utterance k overlaps the
    // forward pass of k+1); else one; else the bits spill to the workspace
    Plan plan = choose_plan(T_x, T_y, flags, &a.lay, prior_extra_smem(F, T_x, T_y, true), 3);
    a.bits_slots = 2;
    if (tuning().prior_spill) {  // tuning knob: bits in L2/HBM, shared memory spent on a deeper ring
        plan = choose_plan(T_x, T_y, flags | MAS_FLAG_SPILL_BITS, &a.lay,
                           prior_extra_smem(F, T_x, T_y, false), 6);
        a.bits_slots = 0;
    } else if (plan != kPlanFastSmemBits) {
        plan = choose_plan(T_x, T_y, flags, &a.lay, prior_extra_smem(F, T_x, T_y, false), 3);
        a.bits_slots = (plan == kPlanFastSmemBits) ? 1 : 0;
    }
    const bool fused = plan != kPlanGeneral;
    if (!workspace || workspace_bytes < mas_workspace_bytes(B, T_x, T_y)) return MAS_ERR_WORKSPACE;
    if ((uintptr_t)workspace % 16) return MAS_ERR_ALIGN;
    if (!fused && !log_prior_out) return MAS_ERR_WORKSPACE;  // unfused plan needs the lp buffer
    if (log_prior_out) {
        const cudaError_t e = launch_log_prior(mu_x, y, log_prior_out, B, F, T_x, T_y, st);
        if (e != cudaSuccess) return (int)e;
    }
    if (!fused) {
        // operands do not fit next to the tile ring: prior to HBM once, then the drop-in kernel
        MasArgs m{};
        const Plan p2 = plan_fast(B, T_x, T_y, flags, &m);
        if (p2 == kPlanGeneral && (size_t)T_x * 16 > (size_t)kSmemBudget) return MAS_ERR_SHAPE;
        m.value = log_prior_out;
        m.t_x = t_x;
        m.t_y = t_y;
        m.path = path;
        m.durations = durations;
        m.score = score;
        m.frame_idx = frame_idx;
        m.bits_ws = static_cast<uint32_t *>(workspace);
        m.B = B;
        m.T_x = T_x;
        m.T_y = T_y;
        m.path_esize = path ? esize : 4;
        m.one = one_pattern(path_dtype);
        m.load_mode = (T_y % 4 == 0 && (uintptr_t)log_prior_out % 16 == 0) ? 2
                      : (T_y % 2 == 0 && (uintptr_t)log_prior_out % 8 == 0 && !m.skewed) ? 4 : 1;
        return (int)((p2 == kPlanGeneral) ? launch_general(m, MAS_F32, st)
                     : m.skewed           ? launch_fast3(m, MAS_F32, st)
                                          : launch_fast(m, MAS_F32, st));
    }
    a.mu_x = mu_x;
    a.y = y;
    a.t_x = t_x;
    a.t_y = t_y;
    a.path = path;
    a.durations = durations;
    a.frame_idx = frame_idx;
    a.score = score;
    a.bits_ws = static_cast<uint32_t *>(workspace);
    a.stats = nullptr;
    if (tuning().prior_stats) {  // profiling aid: counters at the tail of an over-sized workspace
        const size_t need = mas_workspace_bytes(B, T_x, T_y), sbytes = (size_t)1024 * 16 * 8;
        if (workspace_bytes >= need + sbytes + 16)
            a.stats = reinterpret_cast<long long *>(static_cast<char *>(workspace) +
                                                    ((workspace_bytes - sbytes) & ~(size_t)15));
    }
    a.fma_per_smsp = tuning().fma_per_smsp;
    a.extra_fma = tuning().extra_fma;
    a.B = B;
    a.F = F;
    a.T_x = T_x;
    a.T_y = T_y;
    a.path_esize = path ? esize : 4;
    a.one = one_pattern(path_dtype);
    return (int)launch_from_prior(a, st);
}

}  // namespace mas

extern "C" {

int mas_from_prior_f32(const float *mu_x, const float *logs, const float *y, const int32_t *t_x,
                       const int32_t *t_y, void *path, int path_dtype, int32_t *durations,
                       int32_t *frame_idx, float *score, float *log_prior_out, int B, int F, int T_x,
                       int T_y, void *workspace, size_t workspace_bytes, int flags, void *stream)
{
    return from_prior_impl(mu_x, logs, y, t_x, t_y, path, path_dtype, durations, frame_idx, score,
                           log_prior_out, B, F, T_x, T_y, workspace, workspace_bytes, flags,
                           static_cast<cudaStream_t>(stream), nullptr, 0);
}

int mas_from_prior_peer_f32(const float *mu_x, const float *logs, const float *y, const int32_t *t_x,
                            const int32_t *t_y, void *path, int path_dtype, int32_t *durations,
                            int32_t *frame_idx, float *score, float *log_prior_out, int B, int F,
                            int T_x, int T_y, void *workspace, size_t workspace_bytes, int flags,
                            void *stream, const mas_peer_gather *peer)
{
    return from_prior_impl(mu_x, logs, y, t_x, t_y, path, path_dtype, durations, frame_idx, score,
                           log_prior_out, B, F, T_x, T_y, workspace, workspace_bytes, flags,
                           static_cast<cudaStream_t>(stream), peer, 0);
}

int mas_generate_path(const void *durations, int dur_dtype, const int32_t *t_x, const int32_t *t_y,
                      void *path, int path_dtype, int B, int T_x, int T_y, void *stream)
{
    if (!durations || !path) return MAS_ERR_NULL;
    if (!shape_ok(B, T_x, T_y)) return MAS_ERR_SHAPE;
    if (dur_dtype != MAS_I32 && dur_dtype != MAS_F32) return MAS_ERR_DTYPE;
    const int esize = element_size(path_dtype);
    if (esize == 0 || path_dtype == MAS_I64) return MAS_ERR_DTYPE;
    if ((uintptr_t)durations % 4 || (uintptr_t)path % esize || (uintptr_t)t_x % 4 ||
        (uintptr_t)t_y % 4)
        return MAS_ERR_ALIGN;
    if ((size_t)(T_x + 1) * 4 > (size_t)kSmemBudget) return MAS_ERR_SHAPE;
    if (B == 0) return MAS_OK;
    return (int)launch_generate_path(durations, dur_dtype, t_x, t_y, path, esize,
                                     one_pattern(path_dtype), B, T_x, T_y,
                                     static_cast<cudaStream_t>(stream));
}

static bool misaligned4(std::initializer_list<const void *> ps)
{
    for (const void *p : ps)
        if ((uintptr_t)p % 4) return true;
    return false;
}

int mas_frame_index(const int32_t *durations, const int32_t *t_x, const int32_t *t_y,
                    int32_t *frame_idx, int B, int T_x, int T_y, void *stream)
{
    if (!durations || !frame_idx) return MAS_ERR_NULL;
    if (!shape_ok(B, T_x, T_y) || (size_t)(T_x + 1) * 4 > (size_t)kSmemBudget) return MAS_ERR_SHAPE;
    if (misaligned4({durations, t_x, t_y, frame_idx})) return MAS_ERR_ALIGN;
    if (B == 0) return MAS_OK;
    return (int)launch_frame_index(durations, t_x, t_y, frame_idx, B, T_x, T_y,
                                   static_cast<cudaStream_t>(stream));
}

int mas_duration_loss_f32(const float *logw, const int32_t *durations, const int32_t *t_x,
                          float *logw_target, float *grad_unit, float *loss, int B, int T_x,
                          void *workspace, size_t workspace_bytes, void *stream)
{
    if (!durations || !t_x) return MAS_ERR_NULL;
    if (!logw_target && !grad_unit && !loss) return MAS_ERR_NULL;
    if ((loss || grad_unit) && !logw) return MAS_ERR_NULL;
    if (B < 0 || T_x < 1) return MAS_ERR_SHAPE;
    if (misaligned4({logw, durations, t_x, logw_target, grad_unit, loss})) return MAS_ERR_ALIGN;
    if (loss && (!workspace || workspace_bytes < duration_loss_scratch_bytes())) return MAS_ERR_WORKSPACE;
    if ((uintptr_t)workspace % 8) return MAS_ERR_ALIGN;
    if (B == 0) return MAS_OK;
    return (int)launch_duration_loss(logw, durations, t_x, logw_target, grad_unit, loss,
                                     static_cast<double *>(workspace), B, T_x,
                                     static_cast<cudaStream_t>(stream));
}

static bool segment_shape_ok(int B, int R, int T_y, int T_out)
{
    return B >= 0 && B <= 65535 && R >= 1 && T_y >= 1 && T_out >= 1 && R <= (1 << 24) &&
           T_y <= (1 << 24) && T_out <= (1 << 24) && (R + 7) / 8 <= 65535;
}

int mas_crop_f32(const float *src, const int32_t *offset, const int32_t *seg_len, float *dst, int B,
                 int R, int T_y, int T_out, void *stream)
{
    if (!src || !dst) return MAS_ERR_NULL;
    if (!segment_shape_ok(B, R, T_y, T_out)) return MAS_ERR_SHAPE;
    if (misaligned4({src, offset, seg_len, dst})) return MAS_ERR_ALIGN;
    if (B == 0) return MAS_OK;
    return (int)launch_crop_rows(src, offset, seg_len, dst, B, R, T_y, T_out,
                                 static_cast<cudaStream_t>(stream));
}

int mas_path_segment(const int32_t *frame_idx, const int32_t *offset, const int32_t *seg_len,
                     void *path, int path_dtype, int B, int T_x, int T_y, int T_out, void *stream)
{
    if (!frame_idx || !path) return MAS_ERR_NULL;
    if (!segment_shape_ok(B, T_x, T_y, T_out)) return MAS_ERR_SHAPE;
    const int esize = element_size(path_dtype);
    if (esize == 0 || path_dtype == MAS_I64) return MAS_ERR_DTYPE;
    if (misaligned4({frame_idx, offset, seg_len}) || (uintptr_t)path % esize) return MAS_ERR_ALIGN;
    if (B == 0) return MAS_OK;
    return (int)launch_path_segment(frame_idx, offset, seg_len, path, esize, one_pattern(path_dtype),
                                    B, T_x, T_y, T_out, static_cast<cudaStream_t>(stream));
}

size_t mas_align_workspace_bytes(int B, int F, int T_out)
{
    if (B < 1 || F < 1 || T_out < 1) return 2048;
    return std::max(align_partials(B, F, T_out) * sizeof(float), duration_loss_scratch_bytes()) + 256;
}

int mas_align_gather_f32(const float *mu_x, const int32_t *frame_idx, const int32_t *offset,
                         const int32_t *seg_len, const float *y_seg, float *mu_y, float *prior_loss,
                         int B, int F, int T_x, int T_y, int T_out, void *workspace,
                         size_t workspace_bytes, void *stream)
{
    if (!mu_x || !frame_idx) return MAS_ERR_NULL;
    if (!mu_y && !prior_loss) return MAS_ERR_NULL;
    if (prior_loss && !y_seg) return MAS_ERR_NULL;
    if (!segment_shape_ok(B, F, T_y, T_out) || T_x < 1 || T_x > (1 << 24)) return MAS_ERR_SHAPE;
    if (misaligned4({mu_x, frame_idx, offset, seg_len, y_seg, mu_y, prior_loss, workspace}))
        return MAS_ERR_ALIGN;
    if (prior_loss && (!workspace || workspace_bytes < mas_align_workspace_bytes(B, F, T_out)))
        return MAS_ERR_WORKSPACE;
    if (B == 0) return MAS_OK;
    return (int)launch_align_gather(mu_x, frame_idx, offset, seg_len, y_seg, mu_y, prior_loss,
                                    static_cast<float *>(workspace), B, F, T_x, T_y, T_out,
                                    static_cast<cudaStream_t>(stream));
}

int mas_align_gather_bwd_f32(const float *grad_mu_y, const float *y_seg, const float *mu_x,
                             const float *grad_loss, const float *loss_norm,
                             const int32_t *frame_idx, const int32_t *offset, const int32_t *seg_len,
                             float *grad_mu_x, int B, int F, int T_x, int T_y, int T_out,
                             void *stream)
{
    if (!frame_idx || !grad_mu_x) return MAS_ERR_NULL;
    if (!segment_shape_ok(B, F, T_y, T_out) || T_x < 1 ||
        (size_t)T_x * 8 + (size_t)T_out * 8 > (size_t)kSmemBudget)
        return MAS_ERR_SHAPE;
    if (misaligned4({grad_mu_y, y_seg, mu_x, grad_loss, loss_norm, frame_idx, offset, seg_len,
                     grad_mu_x}))
        return MAS_ERR_ALIGN;
    const bool with_loss = y_seg && mu_x && grad_loss && loss_norm;
    if (B == 0) return MAS_OK;
    return (int)launch_align_gather_bwd(grad_mu_y, with_loss ? y_seg : nullptr, mu_x,
                                        with_loss ? grad_loss : nullptr, loss_norm, frame_idx, offset,
                                        seg_len, grad_mu_x, B, F, T_x, T_y, T_out,
                                        static_cast<cudaStream_t>(stream));
}

}  // extern "C"
