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
constexpr int kRowsPerWarp = 6;     // row_a, row_b, unp, utb, bs, mvs
constexpr int kMaxSmemCap = 1024;   // widest band served from shared memory (24 KB of rows per warp)
constexpr size_t kSmemBudget = 220 * 1024;

// Shared-memory row capacity for a batch whose widest band is max_w: the smallest of 256/512/1024
// that holds it (a smaller capacity lets more warps share an SM); reads with still wider bands
// (long stalls) take their rows from a global scratch area instead.
inline int pick_cap(int max_w) { return max_w <= 256 ? 256 : (max_w <= 512 ? 512 : kMaxSmemCap); }
inline size_t warp_words(int cap) { return (size_t)kRowsPerWarp * cap + 4; }  // + slot (padded)
inline size_t cta_smem_bytes(int cap) { return (size_t)kWarpsPerCta * warp_words(cap) * sizeof(float); }
inline int ctas_per_sm(int cap) {
    const int by_smem = (int)(kSmemBudget / (cta_smem_bytes(cap) + 1024));
    return by_smem > 4 ? 4 : (by_smem < 1 ? 1 : by_smem);  // 48 registers x 256 threads: <= 5 CTAs
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
};

struct RefineArgs {
    const float *sig;          // normalised signal of all reads, read r at sig_off[r]
    const int64_t *sig_off;    // [n_reads + 1]
    const float *levels;       // expected level per base, read r at seq_off[r]
    const int32_t *band_st;    // band start (signal coordinate relative to the read) per base
    const int32_t *band_en;    // band end per base
    const int64_t *seq_off;    // [n_reads + 1]
    const int64_t *tb_off;     // [n_reads + 1] offsets into the traceback workspace
    const int32_t *max_w;      // [n_reads] widest band of the read
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
    float *wide_scratch;       // kRowsPerWarp * wide_w + 4 words per warp of the grid, or NULL
    int wide_w;                // band capacity of the global scratch rows
    int cap;                   // band capacity of the shared-memory rows
};

__global__ void __launch_bounds__(kWarpsPerCta * 32) refine_dp_kernel(const RefineArgs a) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5;
    WarpCtx ctx{(int)(threadIdx.x & 31)};
    float *sbase = smem + (size_t)warp * ((size_t)kRowsPerWarp * a.cap + 4);
    float *gbase = a.wide_scratch
                       ? a.wide_scratch + ((size_t)blockIdx.x * kWarpsPerCta + warp) *
                                              ((size_t)kRowsPerWarp * a.wide_w + 4)
                       : nullptr;
    for (;;) {
        int q = 0;
        if (ctx.lane == 0) q = atomicAdd(a.counter, 1);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= a.n_reads) break;
        const int r = a.order ? a.order[q] : q;
        const int64_t so = a.seq_off[r];
        const int n_bases = (int)(a.seq_off[r + 1] - so);
        const int w = a.max_w[r];
        float *base = sbase;
        int cap = a.cap;
        if (w > a.cap) {  // rare: a stall wider than the shared-memory rows
            base = gbase;
            cap = a.wide_w;
        }
        refine::refine_read_warp(ctx, a.sig + a.sig_off[r], a.levels + so, a.band_st + so, a.band_en + so,
                                 n_bases, a.pen, a.n_pen, a.algo, a.tb + a.tb_off[r], a.path + so + r,
                                 a.score + r, a.status + r, base, base + cap, base + 2 * (size_t)cap,
                                 reinterpret_cast<int32_t *>(base + 3 * (size_t)cap), base + 4 * (size_t)cap,
                                 base + 5 * (size_t)cap, reinterpret_cast<int32_t *>(base + 6 * (size_t)cap));
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

size_t refine_wide_scratch_bytes(int sm_count, int max_band_width) {
    if (max_band_width <= kMaxSmemCap) return 0;
    const size_t warps = (size_t)sm_count * ctas_per_sm(kMaxSmemCap) * kWarpsPerCta;
    return warps * warp_words((max_band_width + 3) & ~3) * sizeof(float);
}

int launch_refine_dp(const float *sig, const int64_t *sig_off, const float *levels, const int32_t *band_st,
                     const int32_t *band_en, const int64_t *seq_off, const int64_t *tb_off,
                     const int32_t *max_w, const int32_t *order, int n_reads, const float *pen, int n_pen,
                     int algo, int max_band_width, int32_t *tb, int32_t *path, float *score,
                     int32_t *status, int32_t *counter, float *wide_scratch, int sm_count,
                     cudaStream_t stream) {
    if (n_reads == 0) return RB200_OK;
    RefineArgs a;
    a.sig = sig;
    a.sig_off = sig_off;
    a.levels = levels;
    a.band_st = band_st;
    a.band_en = band_en;
    a.seq_off = seq_off;
    a.tb_off = tb_off;
    a.max_w = max_w;
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
    a.cap = pick_cap(max_band_width);
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
