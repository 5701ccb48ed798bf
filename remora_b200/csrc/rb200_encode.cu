// Dense k-mer one-hot encoder for sm_100a.
//
// Replaces the reference's Cython op compute_encoded_kmer_batch
// (src/remora/encoded_kmers.pyx:13-45).  Same scatter semantics as the reference loop nest:
// for chunk c, k-mer offset p, base index s < seq_len: b = seqs[c, s+p]; if b != -1 the row
// 4p+b is set to 1.0f over [map[s], map[s+1]).  The [4k x T] tile of one chunk is built in shared
// memory (zero fill + scatter) and leaves the SM as ONE bulk asynchronous copy (TMA,
// cp.async.bulk shared->global), double buffered so that the store of chunk i overlaps the fill of
// chunk i+1.  HBM traffic = the 16*k*T output bytes per chunk + ~100 B of compact input: the
// kernel is a pure HBM-write stream (roofline: DESIGN.md "encode_dense").
#include "rb200_internal.cuh"

namespace rb200 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void scatter_chunk(float *tile, const int8_t *__restrict__ seq,
                                              const int16_t *__restrict__ map, int seq_len,
                                              int kmer_len, int T, int t_lo, int t_hi, int pitch) {
    // one (p, s) pair per thread iteration; writes are confined to [t_lo, t_hi)
    const int n_pairs = kmer_len * seq_len;
    for (int idx = threadIdx.x; idx < n_pairs; idx += blockDim.x) {
        const int p = idx / seq_len;
        const int s = idx - p * seq_len;
        const int b = seq[s + p];
        if (b < 0 || b > 3) continue;  // -1 = N / beyond read end (pyx:39-40)
        int st = map[s], en = map[s + 1];
        st = max(st, t_lo);
        en = min(min(en, T), t_hi);
        float *row = tile + (4 * p + b) * pitch - t_lo;
        for (int t = st; t < en; ++t) row[t] = 1.0f;
    }
}

// Whole-chunk tiles, TMA bulk store.  Dynamic smem: n_buf * rows * T floats.
__global__ void __launch_bounds__(256)
encode_dense_tma_kernel(const int8_t *__restrict__ seqs, int seq_width,
                        const int16_t *__restrict__ maps, int map_width,
                        const int16_t *__restrict__ lens, int n_chunks, int kmer_len, int T,
                        float *__restrict__ out, int n_buf) {
    extern __shared__ __align__(128) float enc_smem[];
    const int rows = 4 * kmer_len;
    const int tile_floats = rows * T;
    const uint32_t tile_bytes = static_cast<uint32_t>(tile_floats) * 4u;
    int buf = 0;
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        float *tile = enc_smem + (size_t)buf * tile_floats;
        // the bulk store that last read this buffer must have finished reading it
        if (threadIdx.x == 0) {
            if (n_buf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
        float4 *t4 = reinterpret_cast<float4 *>(tile);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = threadIdx.x; i < tile_floats / 4; i += blockDim.x) t4[i] = z;
        __syncthreads();
        int seq_len = lens[c];
        seq_len = max(0, min(seq_len, min(map_width - 1, seq_width - kmer_len + 1)));
        scatter_chunk(tile, seqs + (size_t)c * seq_width, maps + (size_t)c * map_width, seq_len,
                      kmer_len, T, 0, T, T);
        // make the generic-proxy smem writes visible to the async (TMA) proxy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            float *dst = out + (size_t)c * tile_floats;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                         "r"(smem_u32(tile)), "r"(tile_bytes)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (n_buf == 2) buf ^= 1;
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Fallback for tiles that do not fit shared memory or an unaligned output: time-tiled, plain stores.
__global__ void __launch_bounds__(256)
encode_dense_tiled_kernel(const int8_t *__restrict__ seqs, int seq_width,
                          const int16_t *__restrict__ maps, int map_width,
                          const int16_t *__restrict__ lens, int n_chunks, int kmer_len, int T,
                          float *__restrict__ out, int tile_t) {
    extern __shared__ __align__(128) float enc_smem[];
    const int rows = 4 * kmer_len;
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        int seq_len = lens[c];
        seq_len = max(0, min(seq_len, min(map_width - 1, seq_width - kmer_len + 1)));
        for (int t_lo = 0; t_lo < T; t_lo += tile_t) {
            const int t_hi = min(T, t_lo + tile_t);
            const int w = t_hi - t_lo;
            __syncthreads();
            for (int i = threadIdx.x; i < rows * tile_t; i += blockDim.x) enc_smem[i] = 0.f;
            __syncthreads();
            scatter_chunk(enc_smem, seqs + (size_t)c * seq_width, maps + (size_t)c * map_width,
                          seq_len, kmer_len, T, t_lo, t_hi, tile_t);
            __syncthreads();
            float *dst = out + (size_t)c * rows * T;
            for (int i = threadIdx.x; i < rows * w; i += blockDim.x) {
                const int r = i / w, t = i - r * w;
                dst[(size_t)r * T + t_lo + t] = enc_smem[r * tile_t + t];
            }
        }
    }
}

int launch_encode_dense(const int8_t *seqs, int seq_width, const int16_t *maps, int map_width,
                        const int16_t *lens, int n_chunks, int kmer_len, int T, float *out,
                        int sm_count, cudaStream_t stream, uint64_t *launches) {
    if (n_chunks == 0) return RB200_OK;
    const int rows = 4 * kmer_len;
    const size_t tile_bytes = (size_t)rows * T * sizeof(float);
    const bool aligned = (reinterpret_cast<uintptr_t>(out) % 16) == 0;
    const size_t smem_cap = 200 * 1024;
    if (aligned && tile_bytes <= smem_cap) {
        // two tiles when they fit in ~100 KB (two CTAs per SM stay resident), else one
        const int n_buf = (2 * tile_bytes <= 100 * 1024) ? 2 : 1;
        const size_t smem = n_buf * tile_bytes;
        RB200_CUDA_TRY(cudaFuncSetAttribute(encode_dense_tma_kernel,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem_cap));
        // latency-bound per chunk (zero fill, scatter, two barriers): keep many small CTAs resident
        int ctas_per_sm = (int)(smem_cap / (smem + 1024));
        ctas_per_sm = max(1, min(ctas_per_sm, 8));
        const int grid = min(n_chunks, sm_count * ctas_per_sm);
        encode_dense_tma_kernel<<<grid, 128, smem, stream>>>(seqs, seq_width, maps, map_width, lens,
                                                             n_chunks, kmer_len, T, out, n_buf);
    } else {
        int tile_t = (int)(smem_cap / 2 / (rows * sizeof(float)));
        tile_t = max(32, min(tile_t, T));
        const size_t smem = (size_t)rows * tile_t * sizeof(float);
        RB200_CUDA_TRY(cudaFuncSetAttribute(encode_dense_tiled_kernel,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem_cap));
        const int grid = min(n_chunks, sm_count * 2);
        encode_dense_tiled_kernel<<<grid, 256, smem, stream>>>(seqs, seq_width, maps, map_width,
                                                               lens, n_chunks, kmer_len, T, out,
                                                               tile_t);
    }
    if (launches) ++*launches;
    RB200_CUDA_TRY(cudaGetLastError());
    return RB200_OK;
}

}  // namespace rb200
