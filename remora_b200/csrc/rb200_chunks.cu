// On-GPU chunk extraction ("next" row 1): from one read's raw pieces
// (DAC samples, shift/scale, sequence-to-signal map, integer sequence) and its focus bases, build
// the reference's compact chunk arrays directly in device memory.  Restates, per chunk and in
// parallel, RemoraRead.sig (src/remora/data_chunks.py:191-197), iter_chunks (:425-466) and
// extract_chunk (:331-423); the integer outputs are bit-identical to that code, the float32 signal
// too because the normalisation is evaluated in the same precision numpy uses for the DAC dtype.
#include "rb200_internal.cuh"

namespace rb200 {

__device__ __forceinline__ int upper_bound_i32(const int32_t *a, int n, int v) {  // first i: a[i] > v
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] <= v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int lower_bound_i32(const int32_t *a, int n, int v) {  // first i: a[i] >= v
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// one thread per focus base: chunk centre, overlapping base range
__global__ void chunk_plan_kernel(const int32_t *__restrict__ ssm, int n_map, int sig_len,
                                  const int32_t *__restrict__ focus, int n, int c0, int c1,
                                  int base_start_justify, int offset, int32_t *__restrict__ focus_adj,
                                  int32_t *__restrict__ focus_sig, int32_t *__restrict__ seq_start,
                                  int32_t *__restrict__ seq_len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // add offset and keep inside the read (data_chunks.py:443-446)
    int fb = focus[i] + offset;
    fb = max(min(fb, n_map - 2), 0);
    const int idx = base_start_justify ? ssm[fb] : (ssm[fb] + ssm[fb + 1]) / 2;  // non-negative: // == /
    const int s0 = max(idx - c0, 0), s1 = min(idx + c1, sig_len);
    const int st = upper_bound_i32(ssm, n_map, s0) - 1;  // searchsorted(side="right") - 1
    const int en = lower_bound_i32(ssm, n_map, s1);      // searchsorted(side="left")
    focus_adj[i] = fb;
    focus_sig[i] = idx;
    seq_start[i] = st;
    seq_len[i] = en - st;
}

// dacs dtype codes
enum { DACS_I16 = 0, DACS_F32 = 1, DACS_F64 = 2 };

__device__ __forceinline__ float normalise(const void *dacs, int dtype, int i, double shift, double scale) {
    if (dtype == DACS_I16) {  // int16 - float64 -> float64 in numpy
        const double v = (double)reinterpret_cast<const int16_t *>(dacs)[i];
        return (float)((v - shift) / scale);
    } else if (dtype == DACS_F64) {
        const double v = reinterpret_cast<const double *>(dacs)[i];
        return (float)((v - shift) / scale);
    }
    // float32 array with python-float scalars stays float32 in numpy (two float32 roundings)
    const float v = reinterpret_cast<const float *>(dacs)[i];
    return __fdiv_rn(__fsub_rn(v, (float)shift), (float)scale);
}

// one CTA per chunk
__global__ void __launch_bounds__(128)
chunk_fill_kernel(const void *__restrict__ dacs, int dtype, int sig_len, double shift, double scale,
                  const int32_t *__restrict__ ssm, int n_map, const int8_t *__restrict__ int_seq,
                  int n_bases, const int32_t *__restrict__ focus_sig,
                  const int32_t *__restrict__ seq_start, const int32_t *__restrict__ seq_len, int n,
                  int c0, int c1, int kb, int ka, int lmax, float *__restrict__ signal,
                  int8_t *__restrict__ sequence, int16_t *__restrict__ mapping,
                  int16_t *__restrict__ lens) {
    const int c = blockIdx.x;
    if (c >= n) return;
    const int T = c0 + c1;
    const int raw_start = focus_sig[c] - c0;
    const int pad_left = max(-raw_start, 0);
    const int s0 = max(raw_start, 0);
    const int st = seq_start[c], L = seq_len[c];
    // signal, zero padded where the chunk sticks out of the read (data_chunks.py:346-361)
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const int src = raw_start + t;
        signal[(size_t)c * T + t] =
            (src >= 0 && src < sig_len) ? normalise(dacs, dtype, src, shift, scale) : 0.0f;
    }
    // mapping relative to the chunk, ends pinned to the chunk boundaries (:376-382); zero past seq_len
    for (int i = threadIdx.x; i <= lmax; i += blockDim.x) {
        int v = 0;
        if (i <= L) {
            v = ssm[min(st + i, n_map - 1)] - (s0 - pad_left);
            if (i == 0) v = 0;
            if (i == L) v = T;
        }
        mapping[(size_t)c * (lmax + 1) + i] = (int16_t)v;
    }
    // sequence with k-mer context, -1 beyond the read ends and past the chunk's bases (:385-409)
    const int width = lmax + kb + ka;
    for (int i = threadIdx.x; i < width; i += blockDim.x) {
        const int si = st - kb + i;
        const bool in = si >= 0 && si < n_bases && i < L + kb + ka;
        sequence[(size_t)c * width + i] = in ? int_seq[si] : (int8_t)-1;
    }
    if (threadIdx.x == 0) lens[c] = (int16_t)L;
}

int launch_chunk_plan(const int32_t *ssm, int n_map, int sig_len, const int32_t *focus, int n, int c0,
                      int c1, int bsj, int offset, int32_t *focus_adj, int32_t *focus_sig,
                      int32_t *seq_start, int32_t *seq_len, cudaStream_t stream) {
    if (n == 0) return RB200_OK;
    chunk_plan_kernel<<<(n + 127) / 128, 128, 0, stream>>>(ssm, n_map, sig_len, focus, n, c0, c1, bsj,
                                                          offset, focus_adj, focus_sig, seq_start,
                                                          seq_len);
    RB200_CUDA_TRY(cudaGetLastError());
    return RB200_OK;
}

int launch_chunk_fill(const void *dacs, int dtype, int sig_len, double shift, double scale,
                      const int32_t *ssm, int n_map, const int8_t *int_seq, int n_bases,
                      const int32_t *focus_sig, const int32_t *seq_start, const int32_t *seq_len, int n,
                      int c0, int c1, int kb, int ka, int lmax, float *signal, int8_t *sequence,
                      int16_t *mapping, int16_t *lens, cudaStream_t stream) {
    if (n == 0) return RB200_OK;
    chunk_fill_kernel<<<n, 128, 0, stream>>>(dacs, dtype, sig_len, shift, scale, ssm, n_map, int_seq,
                                            n_bases, focus_sig, seq_start, seq_len, n, c0, c1, kb, ka,
                                            lmax, signal, sequence, mapping, lens);
    RB200_CUDA_TRY(cudaGetLastError());
    return RB200_OK;
}

}  // namespace rb200
