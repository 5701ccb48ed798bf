// Signal-mapping refinement on the GPU ("next" row 4): the banded dynamic programme of
// SigMapRefiner.refine_sig_map (reference src/remora/refine_signal_map.py:474-499, 783-840;
// core src/remora/refine_signal_map_core.pyx:118-473) for a whole batch of reads in one launch.
//
//   refine_normalise_kernel   (dacs - shift) / scale -> float32, in the precision numpy uses for the
//                             DAC dtype (reference data_chunks.py:191-197, refine_signal_map.py:482)
//   refine_dp_kernel          persistent CTAs, one warp per read taken from an atomic work queue
//                             (longest reads first); per-warp rows in shared memory, traceback in HBM.
// The per-read algorithm and why it is organised the way it is: rb200_refine_core.cuh.
#include "rb200_internal.cuh"
#include "rb200_refine_core.cuh"

namespace rb200 {

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kMinNearCap = 64, kMaxNearCap = 1024;  // shared-memory row capacity (samples) per warp
constexpr size_t kSmemBudget = 220 * 1024;
constexpr int kWarpExtraWords = 4 + 32;  // slot (padded) + speculation buffer of the traceback

// Default shared-memory row capacity for a batch whose widest band is max_w (callers that know the
// distribution of band widths pass a smaller one: bases with wider bands use global scratch rows).
inline int default_near_cap(int max_w) { return max_w <= 256 ? 256 : (max_w <= 512 ? 512 : kMaxNearCap); }
inline size_t warp_words(int cap) { return (size_t)refine::kRowsPerWarp * cap + kWarpExtraWords; }
inline size_t cta_smem_bytes(int cap) { return (size_t)kWarpsPerCta * warp_words(cap) * sizeof(float); }
inline int ctas_per_sm(int cap) {
    const int by_smem = (int)(kSmemBudget / (cta_smem_bytes(cap) + 1024));
    // 126 registers x 256 threads: two CTAs (16 warps) per SM.  Measured on B200: register-capped
    // variants with 24 / 32 resident warps are no faster (80 registers) or much slower (64 registers:
    // the chain's shared-memory addresses get rematerialised every step).
    return by_smem > 2 ? 2 : (by_smem < 1 ? 1 : by_smem);
}

enum { DACS_I16 = 0, DACS_F32 = 1, DACS_F64 = 2, DACS_F32_AS_F64 = 3 };

__device__ __forceinline__ float normalise_one(const void *dacs, int dtype, int64_t i, double shift,
                                               double scale) {
    if (dtype == DACS_I16) {  // int16 - float64 -> float64 in numpy, one rounding to float32
        const double v = (double)reinterpret_cast<const int16_t *>(dacs)[i];
        return (float)((v - shift) / scale);
    } else if (dtype == DACS_F64) {
        const double v = reinterpret_cast<const double *>(dacs)[i];
        return (float)((v - shift) / scale);
    }
    if (dtype == DACS_F32_AS_F64) {
        // float32 samples with numpy float64 scalars (shift/scale after rough re-scaling): numpy
        // promotes the expression to float64
        const double v = (double)reinterpret_cast<const float *>(dacs)[i];
        return (float)((v - shift) / scale);
    }
    // float32 samples with python-float scalars stay float32 in numpy (two float32 roundings)
    const float v = reinterpret_cast<const float *>(dacs)[i];
    return __fdiv_rn(__fsub_rn(v, (float)shift), (float)scale);
}

// grid (ceil(max_len / 1024), n_reads) x 256, 4 samples per thread
__global__ void __launch_bounds__(256)
refine_normalise_kernel(const void *__restrict__ dacs, int dtype, const int64_t *__restrict__ sig_off,
                        const double *__restrict__ shift, const double *__restrict__ scale,
                        float *__restrict__ out) {
    const int r = blockIdx.y;
    const int64_t lo = sig_off[r], n = sig_off[r + 1] - lo;
    const double sh = shift[r], sc = scale[r];
    for (int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x; i < n && i < ((int64_t)blockIdx.x + 1) * 1024;
         i += 256)
        out[lo + i] = normalise_one(dacs, dtype, lo + i, sh, sc);
}

struct WarpCtx {
    int lane;
    static constexpr int nl = 32;
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ int scan_max(int v) const {
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v = max(v, o);
        }
        return v;
    }
    __device__ __forceinline__ int bcast_last(int v) const { return __shfl_sync(0xffffffffu, v, 31); }
};

struct RefineArgs {
    const float *sig;          // normalised signal of all reads, read r at sig_off[r]
    const int64_t *sig_off;    // [n_reads + 1]
    const float *levels;       // expected level per base, read r at seq_off[r]
    const int32_t *band_st;    // band start (signal coordinate relative to the read) per base
    const int32_t *band_en;    // band end per base
    const int64_t *seq_off;    // [n_reads + 1]
    const int64_t *tb_off;     // [n_reads + 1] offsets into the traceback workspace
    const int32_t *order;      // [n_reads] processing order (longest first) or NULL
    int n_reads;
    int n_pen;
    int algo;
    float pen[refine::kMaxPen];
    int32_t *tb;               // traceback workspace
    int32_t *path;             // read r at seq_off[r] + r, seq_len + 1 entries
    float *score;              // [n_reads] final forward score (all_scores[-1])
    int32_t *status;           // [n_reads]
    int32_t *counter;          // work queue head (zeroed before the launch)
    float *wide_scratch;       // kRowsPerWarp * wide_w words per warp of the grid, or NULL
    int wide_w;                // capacity of the global scratch rows (widest band, multiple of 4)
    int cap;                   // capacity of the shared-memory rows (multiple of 4)
};

__global__ void __launch_bounds__(kWarpsPerCta * 32, 2) refine_dp_kernel(const RefineArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5;
    WarpCtx ctx{(int)(threadIdx.x & 31)};
    float *sbase = smem + (size_t)warp * ((size_t)refine::kRowsPerWarp * a.cap + kWarpExtraWords);
    const refine::Rows near = refine::carve_rows(sbase, (size_t)a.cap);
    int32_t *slot = reinterpret_cast<int32_t *>(sbase + (size_t)refine::kRowsPerWarp * a.cap);
    int32_t *spec = slot + 4;
    refine::Rows far = {};
    if (a.wide_scratch)
        far = refine::carve_rows(a.wide_scratch + ((size_t)blockIdx.x * kWarpsPerCta + warp) *
                                                      ((size_t)refine::kRowsPerWarp * a.wide_w),
                                 (size_t)a.wide_w);
    for (;;) {
        int q = 0;
        if (ctx.lane == 0) q = atomicAdd(a.counter, 1);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= a.n_reads) break;
        const int r = a.order ? a.order[q] : q;
        const int64_t so = a.seq_off[r];
        const int n_bases = (int)(a.seq_off[r + 1] - so);
        refine::refine_read_warp(ctx, a.sig + a.sig_off[r], a.levels + so, a.band_st + so, a.band_en + so,
                                 n_bases, a.pen, a.n_pen, a.algo, a.tb + a.tb_off[r], a.path + so + r,
                                 a.score + r, a.status + r, near, a.cap, far, slot, spec);
        __syncwarp();
    }
}

}  // namespace

int launch_refine_normalise(const void *dacs, int dtype, const int64_t *sig_off, const double *shift,
                            const double *scale, int n_reads, int64_t max_len, float *out,
                            cudaStream_t stream) {
    if (n_reads == 0 || max_len == 0) return RB200_OK;
    dim3 grid((unsigned)((max_len + 1023) / 1024), (unsigned)n_reads);
    refine_normalise_kernel<<<grid, 256, 0, stream>>>(dacs, dtype, sig_off, shift, scale, out);
    RB200_CUDA_TRY(cudaGetLastError());
    return RB200_OK;
}

static int resolve_near_cap(int near_cap, int max_band_width) {
    if (near_cap <= 0) near_cap = default_near_cap(max_band_width);
    if (near_cap < kMinNearCap) near_cap = kMinNearCap;
    if (near_cap > kMaxNearCap) near_cap = kMaxNearCap;
    return (near_cap + 3) & ~3;
}

size_t refine_wide_scratch_bytes(int sm_count, int near_cap, int max_band_width) {
    near_cap = resolve_near_cap(near_cap, max_band_width);
    if (max_band_width <= near_cap) return 0;
    const size_t warps = (size_t)sm_count * ctas_per_sm(near_cap) * kWarpsPerCta;
    // + 64 bytes: the chain's look-ahead load reads one 16-byte group past the last row of the last warp
    return warps * (size_t)refine::kRowsPerWarp * ((max_band_width + 3) & ~3) * sizeof(float) + 64;
}

int launch_refine_dp(const float *sig, const int64_t *sig_off, const float *levels, const int32_t *band_st,
                     const int32_t *band_en, const int64_t *seq_off, const int64_t *tb_off,
                     const int32_t *order, int n_reads, const float *pen, int n_pen, int algo, int near_cap,
                     int max_band_width, int32_t *tb, int32_t *path, float *score, int32_t *status,
                     int32_t *counter, float *wide_scratch, int sm_count, cudaStream_t stream) {
    if (n_reads == 0) return RB200_OK;
    RefineArgs a;
    a.sig = sig;
    a.sig_off = sig_off;
    a.levels = levels;
    a.band_st = band_st;
    a.band_en = band_en;
    a.seq_off = seq_off;
    a.tb_off = tb_off;
    a.order = order;
    a.n_reads = n_reads;
    a.n_pen = n_pen;
    a.algo = algo;
    for (int i = 0; i < refine::kMaxPen; ++i) a.pen[i] = (pen && i < n_pen) ? pen[i] : 0.0f;
    a.tb = tb;
    a.path = path;
    a.score = score;
    a.status = status;
    a.counter = counter;
    a.cap = resolve_near_cap(near_cap, max_band_width);
    a.wide_scratch = max_band_width > a.cap ? wide_scratch : nullptr;
    a.wide_w = (max_band_width + 3) & ~3;  // rows are read in 16-byte groups
    const size_t smem = cta_smem_bytes(a.cap);
    RB200_CUDA_TRY(cudaFuncSetAttribute(refine_dp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    RB200_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(int32_t), stream));
    // persistent CTAs of 8 warps, as many per SM as the rows allow; never more warps than reads
    int ctas = sm_count * ctas_per_sm(a.cap);
    const int need = (n_reads + kWarpsPerCta - 1) / kWarpsPerCta;
    if (ctas > need) ctas = need;
    refine_dp_kernel<<<ctas, kWarpsPerCta * 32, smem, stream>>>(a);
    RB200_CUDA_TRY(cudaGetLastError());
    return RB200_OK;
}

}  // namespace rb200
