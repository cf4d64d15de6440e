// mas_fast3.cu -- drop-in maximum_path kernel on the skewed-lane recurrence (mas_dp3.cuh), T_x <= 256.
//
// Replaces maximum_path_c (src/model/monotonic_align/core.pyx:38-45) + the host glue of maximum_path
// (monotonic_align/__init__.py:13-23), like mas_fast_kernel (mas_kernels.cu), with the same roles --
// one CTA per utterance, four staging warps stream the band of every 32-frame tile from HBM and clear
// the dense output, one DP warp runs the recurrence, the backtrack runs on direction bits in shared
// memory -- but:
//   * the DP warp is the skewed-lane one: no shuffle and no shared-memory load on the dependency
//     chain (mas_dp3.cuh; 73 -> ~25-40 cycles per frame), which is what the latency-bound shapes
//     (BASELINE configs 1-3: fewer utterances than SMs) are made of;
//   * tiles land in per-row circular buffers four tiles deep (two in use by the skewed lanes, two of
//     prefetch), rows in natural token order;
//   * the dense output is cleared with bulk (TMA) stores from a zeroed shared buffer, one instruction
//     per 2 KB instead of one 16-byte store per thread.
#include "mas_dp3.cuh"
#include "mas_internal.h"

namespace mas {

namespace {

constexpr int kF3Threads = 160;   // warps 0-3: staging, warp 4: DP
constexpr int kF3Helpers = 4;
constexpr int kF3ZeroBytes = 2048;

// Tile t -> ring rows, then ONE arrival of this thread on the tile's `full` barrier.  The band of the tile
// (max(0, t_x + y - t_y) <= x <= min(t_x - 1, y)) is copied from HBM; every cell ABOVE the diagonal
// (x > frame) is stored as 0.0 -- that is what keeps those cells at exactly -1e9 in the unguarded
// recurrence (mas_dp3.cuh) -- so tiles that still reach the diagonal (32 t < t_x) also write the rows
// above their band.
// mode 2: cp.async 16 B (fp32, rows 16-byte aligned), mode 1: cp.async 4 B, mode 0: LDG + convert
// (+ cell mask) + STS.  (One bulk/TMA copy per row was measured and is far slower: the copy engine
// retires one small request per ~26 cycles, profiles/r2_fast3_phases.txt.)
template <typename InT>
__device__ __forceinline__ void stage3(const InT *__restrict__ vb, const float *__restrict__ mb, float *rows,
                                       uint64_t *full, int t, int g0, int tx, int ty, int64_t T_y, int htid,
                                       int mode)
{
    const int lane = htid & 31, hw = htid >> 5;
    const int y0 = t * kTileY;
    const int fl0 = 32 * (g0 + t);               // lifetime frame of the tile's first column
    const int lo = max(0, tx + y0 - ty);
    const int hi = (y0 < tx) ? tx - 1 : min(tx - 1, y0 + kTileY - 1);   // diagonal tiles: zeros above the band
    if (sizeof(InT) == 4 && mode == 2) {
        const float *v32 = reinterpret_cast<const float *>(vb);
        const int c = lane & 7, rq = lane >> 3;
        const int f0 = y0 + 4 * c;               // first frame of this lane's 16-byte chunk
        const int left = ty - f0;
        const uint32_t bytes = left >= 4 ? 16u : (left > 0 ? 4u * left : 0u);
        for (int x = lo + 4 * hw + rq; x <= hi; x += 4 * kF3Helpers) {
            float *dst = rows + x * kRing3Pitch + dp3_col(fl0 + 4 * c);
            if (x <= f0) {                       // on / below the diagonal: plain copy
                cp_async16(dst, v32 + (int64_t)x * T_y + (bytes ? f0 : 0), bytes);
            } else if (x > f0 + 3) {             // wholly above: zeros, nothing is read
                cp_async16(dst, v32, 0u);
            } else {                             // the chunk the diagonal crosses: frame by frame
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int f = f0 + e;
                    const bool take = f >= x && f < ty;
                    cp_async4(dst + e, v32 + (int64_t)x * T_y + (take ? f : 0), take ? 4u : 0u);
                }
            }
        }
        cp_async_arrive(full);
    } else {
        const int y = y0 + lane;
        const bool in = y < ty;
        for (int x = lo + hw; x <= hi; x += kF3Helpers) {
            float *dst = rows + x * kRing3Pitch + dp3_col(fl0 + lane);
            const bool take = in && y >= x;
            if (sizeof(InT) == 4 && mode == 1) {
                cp_async4(dst, reinterpret_cast<const float *>(vb) + (int64_t)x * T_y + (take ? y : 0), take ? 4u : 0u);
            } else {
                float v = 0.0f;
                if (take) {
                    const int64_t e = (int64_t)x * T_y + y;
                    v = load_as_f32(vb + e);
                    if (mb) v *= __ldg(mb + e);
                }
                *dst = v;
            }
        }
        if (sizeof(InT) == 4 && mode == 1) cp_async_arrive(full);
        else mbar_arrive(full);
    }
}

template <typename InT, int XPLMAX>
__global__ void __launch_bounds__(kF3Threads) mas_fast3_kernel(const MasArgs a)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const FastLayout &L = a.lay;
    float *rows = reinterpret_cast<float *>(smem + L.off_stages);
    uint32_t *bits = reinterpret_cast<uint32_t *>(smem + L.off_bits);
    int *first = reinterpret_cast<int *>(smem + L.off_first);
    int *dur = reinterpret_cast<int *>(smem + L.off_dur);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.off_bars);
    uint32_t *zbuf = reinterpret_cast<uint32_t *>(smem + L.off_bars + 128);

    const int b = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T_x = a.T_x;
    const int64_t T_y = a.T_y;
    const int tx = min(max(a.t_x[b], 0), T_x);
    const int ty = min(max(a.t_y[b], 0), a.T_y);
    const bool degenerate = tx > ty && ty >= 1;
    const bool active = tx >= 1 && ty >= 1 && !degenerate;
    const int ntiles = active ? (ty + kTileY - 1) / kTileY : 0;

    Ring3 ring;
    ring.rows = rows;
    ring.full = bars;
    ring.empty = bars + kRing3Stages;
    if (tid == 0) {
        for (int s = 0; s < kRing3Stages; ++s) {
            mbar_init(&ring.full[s], kF3Helpers * 32);   // every staging thread arrives once per tile
            mbar_init(&ring.empty[s], 1);
        }
        mbar_fence_init();
    }
    for (int i = tid; i < kF3ZeroBytes / 4; i += kF3Threads) zbuf[i] = 0u;
    fence_proxy_async_smem();
    __syncthreads();

    const InT *vb = static_cast<const InT *>(a.value) + (int64_t)b * T_x * T_y;
    const float *mb = a.cell_mask ? a.cell_mask + (int64_t)b * T_x * T_y : nullptr;
    char *pb = a.path ? static_cast<char *>(a.path) + (int64_t)b * T_x * T_y * a.path_esize : nullptr;
    const int64_t pbytes = a.path ? (int64_t)T_x * T_y * a.path_esize : 0;

    long long *so = a.stats ? a.stats + (size_t)b * 16 : nullptr;
    const long long t_begin = so ? clock64() : 0;
    if (warp == kF3Helpers) {
        // ---------------- DP warp: forward recurrence + backtrack ----------------
        for (int x = lane; x < T_x; x += 32) dur[x] = 0;
        __syncwarp();
        float score = 0.0f;
        if (active) {
            long long wacc = 0;
            score = dp3_forward_dispatch<XPLMAX>(ring, bits, L.xrows, tx, ty, lane, 0, so ? &wacc : nullptr);
            __syncwarp();
            const long long t_fwd = so ? clock64() : 0;
            if (lane == 0) backtrack_nat(bits, L.xrows, tx, ty, first);
            __syncwarp();
            for (int x = lane; x < tx; x += 32) dur[x] = ((x == tx - 1) ? ty : first[x + 1]) - first[x];
            if (so && lane == 0) {
                so[0] = t_fwd - t_begin;      // forward pass (incl. waits)
                so[1] = wacc;                 // ... of which starved of tiles
                so[2] = clock64() - t_fwd;    // backtrack
                so[3] = ty;
            }
        } else if (degenerate) {
            if (lane == 0) {
                auto val = [&](int x, int y) {
                    const int64_t e = (int64_t)x * T_y + y;
                    float v = load_as_f32(vb + e);
                    if (mb) v *= mb[e];
                    return v;
                };
                backtrack_degenerate(val, tx, ty, first, dur);
                score = val(tx - 1, ty - 1);  // the reference leaves value[t_x-1,t_y-1] untouched
            }
            score = __shfl_sync(kFull, score, 0);
        }
        if (lane == 0 && a.score) a.score[b] = score;
    } else {
        // ---------------- staging warps: HBM -> ring, and the zero fill of the output ----------------
        const bool zbulk = bulk_zero_ok(pb, pbytes);
        const int nparts = max(ntiles, 1);
        long long w_empty = 0, w_issue = 0, w_arrive = 0, w_zero = 0;
        for (int t = 0; t < ntiles; ++t) {
            const int s = t & 3;
            const long long tw = so ? clock64() : 0;
            if (t >= kRing3Stages) mbar_wait(&ring.empty[s], (uint32_t)((t >> 2) - 1) & 1u);
            const long long t1 = so ? clock64() : 0;
            if (so) w_empty += t1 - tw;
            stage3<InT>(vb, mb, rows, &ring.full[s], t, 0, tx, ty, T_y, tid, a.load_mode);
            const long long t2 = so ? clock64() : 0;
            const long long t3 = t2;
            if (zbulk) zero_fill_bulk_part(pb, pbytes, t, nparts, zbuf, kF3ZeroBytes, tid, kF3Helpers * 32);
            else zero_fill_part(pb, pbytes, t, nparts, tid, kF3Helpers * 32);
            if (so) {
                w_issue += t2 - t1;
                w_arrive += t3 - t2;
                w_zero += clock64() - t3;
            }
        }
        if (ntiles == 0) {
            if (zbulk) zero_fill_bulk_part(pb, pbytes, 0, 1, zbuf, kF3ZeroBytes, tid, kF3Helpers * 32);
            else zero_fill_part(pb, pbytes, 0, 1, tid, kF3Helpers * 32);
        }
        const long long tz = so ? clock64() : 0;
        if (zbulk) {   // the zeros must be in memory before anybody writes a 1-cell
            bulk_commit();
            bulk_wait_all();
        }
        if (so && tid == 0) {
            so[4] = tz - t_begin;          // staging loop (incl. waits for free stages)
            so[5] = w_empty;               // ... of which waiting for the DP warp
            so[6] = clock64() - tz;        // waiting for the bulk stores
            so[9] = w_issue;               // issuing the copies of the tiles
            so[10] = w_arrive;             // cp.async -> mbarrier arrive
            so[11] = w_zero;               // issuing the zero fill
        }
    }
    __syncthreads();
    const long long t_tail = so ? clock64() : 0;
    write_path_ones(pb, a.durations ? a.durations + (int64_t)b * T_x : nullptr, first, dur, T_x, T_y,
                    a.path_esize, a.one, tid, kF3Threads);
    write_frame_idx(a.frame_idx ? a.frame_idx + (int64_t)b * T_y : nullptr, first, dur, T_x, ty, a.T_y, tid,
                    kF3Threads);
    if (so && tid == 0) {
        so[7] = clock64() - t_tail;        // 1-cells, durations, frame index
        so[8] = clock64() - t_begin;       // whole CTA
    }
}

template <typename InT>
cudaError_t launch3_typed(const MasArgs &a, cudaStream_t st)
{
    const int xplmax = (a.T_x + 31) / 32;
    void (*k)(const MasArgs) = nullptr;
    if constexpr (sizeof(InT) == 4) {   // fp32: register budget sized to the batch's T_x bucket
        if (xplmax <= 2) k = mas_fast3_kernel<InT, 2>;
        else if (xplmax <= 4) k = mas_fast3_kernel<InT, 4>;
        else if (xplmax <= 6) k = mas_fast3_kernel<InT, 6>;
        else k = mas_fast3_kernel<InT, 8>;
    } else {
        k = mas_fast3_kernel<InT, 8>;
    }
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.lay.total);
    if (e != cudaSuccess) return e;
    k<<<a.B, kF3Threads, a.lay.total, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

// shared-memory carve-up; ok == false when the shape does not qualify (T_x > 256, or bits + ring do not fit)
bool fast3_layout(int T_x, int T_y, FastLayout *lay)
{
    if (T_x > 256) return false;
    FastLayout L{};
    L.xrows = (T_x + 31) / 32 * 32;
    L.nch = (T_y + 31) / 32;
    L.nstages = kRing3Stages;
    L.bits_in_smem = 1;
    const size_t ring = (size_t)L.xrows * kRing3Pitch * 4;
    const size_t bits = (size_t)L.nch * L.xrows * 4;
    L.off_stages = 0;
    L.off_bits = ring;
    L.off_first = L.off_bits + bits;
    L.off_dur = L.off_first + (size_t)T_x * 4;
    L.off_bars = (L.off_dur + (size_t)T_x * 4 + 15) & ~(size_t)15;
    L.total = L.off_bars + 128 + kF3ZeroBytes;
    if (L.total > (size_t)kSmemBudget) return false;
    *lay = L;
    return true;
}

cudaError_t launch_fast3(const MasArgs &a, int value_dtype, cudaStream_t st)
{
    switch (value_dtype) {
    case MAS_F32: return launch3_typed<float>(a, st);
    case MAS_F16: return launch3_typed<__half>(a, st);
    case MAS_BF16: return launch3_typed<__nv_bfloat16>(a, st);
    case MAS_F64: return launch3_typed<double>(a, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace mas
