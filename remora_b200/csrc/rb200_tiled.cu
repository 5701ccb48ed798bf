// Register-tiled layer kernels (RB200_IMPL_TILED): the fast path for every model the fused
// ConvLSTM_w_ref/64 kernels do not cover, first of all Conv_w_ref (models/Conv_w_ref.py:44-62).
//
//  * conv_tile_kernel<KW,S,NT>: Conv1d + folded BatchNorm + swish as a shared-memory tiled direct
//    convolution on packed FFMA2.  A CTA owns `nch` chunks and one tile of `ct` <= 64 output
//    channels; the input [nch][c_in][t_in] is staged once in shared memory, the weights
//    (re-laid out at create time as [co tile][ci][KW][ct]) stream through a double-buffered
//    cp.async ring in slabs of input channels.  A thread holds 8 output channels x NT output
//    positions in registers: per (ci, tap) it reads two broadcast LDS.128 of weights and re-uses a
//    register window of the input row across the KW taps, i.e. 4*NT FFMA2 per 2 LDS.
//    Lanes of a warp are (chunk, position tile) pairs, warps are channel groups, so weight reads
//    are warp-uniform and input reads are contiguous across lanes.
//  * seq1_gather_kernel: the first sequence convolution straight from the compact arrays.  The
//    one-hot input makes the convolution a sum of weight columns:
//    stage A sums the k columns of every (base index, tap) once, stage B adds KW of those sums
//    per output position.  No dense one-hot tensor is written or read.
//  * fc_warp_kernel: Linear over the flattened features, one warp per chunk.
//
// Activations stay channel-first float32 [B][C][T] like the reference tensors, so the layer-kernel
// path (rb200_layers.cu) and this one are interchangeable layer by layer.
#include "rb200_internal.cuh"

#include <cstdlib>

namespace rb200 {

namespace {

constexpr int TILE_THREADS = 256;
constexpr int TILE_WARPS = TILE_THREADS / 32;
constexpr int SLAB_FLOATS = 4096;  // weight floats per ring stage (16 KB)
constexpr size_t SMEM_CAP = 200 * 1024;

__device__ __forceinline__ float2 fma2(float2 w, float x, float2 acc) {
    return __ffma2_rn(w, make_float2(x, x), acc);
}

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
// programmatic dependent launch: the next kernel's prologue may overlap this kernel's tail
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, int block, size_t smem,
                       cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    static const bool no_pdl = getenv("RB200_NO_PDL") != nullptr;  // measurement aid
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // A grid that fits the GPU in one wave is launched plainly: started early, its CTAs would be
    // packed onto whichever SMs the previous kernel has left free instead of one per SM
    // (measured: Conv_w_ref at batch 1024, 270 us with the attribute on every layer vs 227 us).
    cfg.numAttrs = (pdl && !no_pdl) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ void cp_async4(void *dst_smem, const void *src) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct TileArgs {
    const float *x;
    int64_t x_bstride;
    const float *wt;    // [n co tiles][c_in][KW][ct]
    const float *bias;  // [c_out]
    float *y;
    int64_t y_bstride;
    int B, c_in, t_in, c_out, t_out;
    int nch;          // chunks per CTA
    int xp;           // shared-memory row pitch of the input (floats, multiple of 4)
    int chunk_pitch;  // shared-memory pitch between chunks (floats, multiple of 4)
    int ct;           // output channels per CTA (8, 16, 32 or 64)
    int ci_slab;      // input channels per weight ring stage
    int xs_floats;    // input tile + overrun tail
};

template <int KW, int S, int NT>
__global__ void __launch_bounds__(TILE_THREADS, 2) conv_tile_kernel(const TileArgs a) {
    constexpr int V = (NT * S) % 4 == 0 ? 4 : ((NT * S) % 2 == 0 ? 2 : 1);  // window vector width
    constexpr int W = (NT - 1) * S + KW;                                    // input window per thread
    constexpr int WP = (W + V - 1) / V * V;
    extern __shared__ float4 tile_smem4[];
    float *xs = reinterpret_cast<float *>(tile_smem4);
    float *wsm = xs + a.xs_floats;
    const int slab_floats = a.ci_slab * KW * a.ct;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nthr = TILE_THREADS;
    const int chunk0 = blockIdx.x * a.nch;
    const int nvalid = min(a.nch, a.B - chunk0);
    const int co_base = blockIdx.y * a.ct;
    const float *wt = a.wt + (size_t)blockIdx.y * a.c_in * KW * a.ct;
    const int n_slabs = (a.c_in + a.ci_slab - 1) / a.ci_slab;

    auto fetch_slab = [&](int sl) {
        const int ci0 = sl * a.ci_slab;
        const int n4 = min(a.ci_slab, a.c_in - ci0) * KW * a.ct / 4;
        const float4 *src = reinterpret_cast<const float4 *>(wt + (size_t)ci0 * KW * a.ct);
        float4 *dst = reinterpret_cast<float4 *>(wsm + (sl & 1) * slab_floats);
        for (int i = tid; i < n4; i += nthr) cp_async16(dst + i, src + i);
        cp_async_commit();
    };
    fetch_slab(0);  // weights do not depend on the previous kernel
    pdl_wait();     // the input does

    // stage the input of this CTA's chunks: asynchronous copies, all in flight at once (a chunk's
    // [c_in][t_in] block is contiguous in global memory; shared-memory rows are padded to xp)
    {
        const int per_chunk = a.c_in * a.t_in;
        const bool vec = (a.t_in & 3) == 0 && (a.x_bstride & 3) == 0 &&
                         (reinterpret_cast<uintptr_t>(a.x) & 15) == 0;
        if (vec) {
            const int t4n = a.t_in >> 2, pc4 = per_chunk >> 2;
            for (int q = tid; q < nvalid * pc4; q += nthr) {
                const int c = q / pc4, r = q - c * pc4;
                const int ci = r / t4n, t4 = r - ci * t4n;
                cp_async16(xs + c * a.chunk_pitch + ci * a.xp + 4 * t4,
                           a.x + (size_t)(chunk0 + c) * a.x_bstride + 4 * (size_t)r);
            }
        } else {
            for (int e = tid; e < nvalid * per_chunk; e += nthr) {
                const int c = e / per_chunk, r = e - c * per_chunk;
                const int ci = r / a.t_in, t = r - ci * a.t_in;
                cp_async4(xs + c * a.chunk_pitch + ci * a.xp + t,
                          a.x + (size_t)(chunk0 + c) * a.x_bstride + r);
            }
        }
        cp_async_commit();
    }

    // work item of this thread: 8 output channels (warp-uniform) x NT positions of one chunk
    const int n_cg = a.ct >> 3;
    const int cg = warp % n_cg, lb = warp / n_cg;
    const int ntile = (a.t_out + NT - 1) / NT;
    const int item = lb * 32 + lane;
    const bool active = item < a.nch * ntile;
    const int chunk = active ? item / ntile : 0;
    const int t0 = active ? (item - chunk * ntile) * NT : 0;
    const float *xbase = xs + chunk * a.chunk_pitch + t0 * S;

    float2 acc[4][NT];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int t = 0; t < NT; ++t) acc[q][t] = make_float2(0.f, 0.f);

    cp_async_wait_all();
    __syncthreads();
    pdl_launch_dependents();  // the next layer may set up (and park at its pdl_wait) from here on

    for (int sl = 0; sl < n_slabs; ++sl) {
        if (sl + 1 < n_slabs) fetch_slab(sl + 1);  // the other stage was last read in iteration sl-1
        const int ci0 = sl * a.ci_slab;
        const int nci = min(a.ci_slab, a.c_in - ci0);
        const float *wb = wsm + (sl & 1) * slab_floats + cg * 8;
        const float *xr = xbase + ci0 * a.xp;
#pragma unroll 1
        for (int cil = 0; cil < nci; ++cil, xr += a.xp, wb += KW * a.ct) {
            float xw[WP];
            if constexpr (V == 4) {
#pragma unroll
                for (int i = 0; i < WP / 4; ++i) {
                    const float4 v = *reinterpret_cast<const float4 *>(xr + 4 * i);
                    xw[4 * i] = v.x, xw[4 * i + 1] = v.y, xw[4 * i + 2] = v.z, xw[4 * i + 3] = v.w;
                }
            } else if constexpr (V == 2) {
#pragma unroll
                for (int i = 0; i < WP / 2; ++i) {
                    const float2 v = *reinterpret_cast<const float2 *>(xr + 2 * i);
                    xw[2 * i] = v.x, xw[2 * i + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int i = 0; i < WP; ++i) xw[i] = xr[i];
            }
#pragma unroll
            for (int j = 0; j < KW; ++j) {
                const float4 wa = *reinterpret_cast<const float4 *>(wb + j * a.ct);
                const float4 wc = *reinterpret_cast<const float4 *>(wb + j * a.ct + 4);
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const float xv = xw[t * S + j];
                    acc[0][t] = fma2(make_float2(wa.x, wa.y), xv, acc[0][t]);
                    acc[1][t] = fma2(make_float2(wa.z, wa.w), xv, acc[1][t]);
                    acc[2][t] = fma2(make_float2(wc.x, wc.y), xv, acc[2][t]);
                    acc[3][t] = fma2(make_float2(wc.z, wc.w), xv, acc[3][t]);
                }
            }
        }
        cp_async_wait_all();
        __syncthreads();
    }

    if (active && chunk < nvalid) {
        const int co = co_base + cg * 8;
        float *yb = a.y + (size_t)(chunk0 + chunk) * a.y_bstride + (size_t)co * a.t_out + t0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float b0 = __ldg(a.bias + co + 2 * q), b1 = __ldg(a.bias + co + 2 * q + 1);
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                if (t0 + t < a.t_out) {
                    yb[(size_t)(2 * q) * a.t_out + t] = swishf_fast(acc[q][t].x + b0);
                    yb[(size_t)(2 * q + 1) * a.t_out + t] = swishf_fast(acc[q][t].y + b1);
                }
            }
        }
    }
}

// ---- first sequence convolution from the compact arrays ------------------------------------------
// wg: [KW][wj_pitch] with rows [4k + 1][c_out] (wg[j][(4p+b)*c_out + o] = w[o][4p+b][j], row 4k = 0).
// Same input semantics as the dense encoder (rb200_encode.cu, encoded_kmers.pyx:13-45) for
// non-decreasing mappings: position t carries base index s iff map[s] <= t < map[s+1], s < seq_len;
// bases outside 0..3 contribute nothing.
// One warp per chunk (no CTA-wide barriers in the chunk loop): per-warp scratch R / idx / sq.
__global__ void __launch_bounds__(256)
seq1_gather_kernel(const int8_t *__restrict__ seqs, int seq_width, const int16_t *__restrict__ maps,
                   int map_width, const int16_t *__restrict__ lens, const float *__restrict__ wg,
                   int wj_pitch, const float *__restrict__ bias, float *__restrict__ y,
                   int64_t y_bstride, int B, int T, int kmer_len, int kw, int stride, int c_out,
                   int t_out, int lmax, int warp_floats) {
    extern __shared__ float4 gather_smem4[];
    float *wsm = reinterpret_cast<float *>(gather_smem4);  // kw * wj_pitch
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    float *R = wsm + kw * wj_pitch + (size_t)warp * warp_floats;  // lmax * kw * c_out (+ zero row)
    int *idx = reinterpret_cast<int *>(R + ((size_t)lmax * kw + 1) * c_out);  // T
    int8_t *sq = reinterpret_cast<int8_t *>(idx + T);                   // seq_width
    const int c4 = c_out >> 2;
    {
        const float4 *src = reinterpret_cast<const float4 *>(wg);
        float4 *dst = reinterpret_cast<float4 *>(wsm);
        for (int i = tid; i < kw * wj_pitch / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    pdl_wait();  // y may still be read by the previous forward's kernels
    __syncthreads();
    const float4 *w4 = reinterpret_cast<const float4 *>(wsm);
    float4 *R4 = reinterpret_cast<float4 *>(R);
    const int wj4 = wj_pitch >> 2;
    const int zero_row = 4 * kmer_len;   // wg rows [0, 4k) are weights, row 4k is zero
    const int r_zero = lmax * kw * c4;   // float4 index of the zero row behind R
    for (int i = lane; i < c4; i += 32) R4[r_zero + i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = blockIdx.x * nwarp + warp; c < B; c += gridDim.x * nwarp) {
        int seq_len = lens[c];
        seq_len = max(0, min(seq_len, min(map_width - 1, seq_width - kmer_len + 1)));
        seq_len = min(seq_len, lmax);
        const int16_t *map = maps + (size_t)c * map_width;
        __syncwarp();  // readers of the previous chunk are done
        for (int t = lane; t < T; t += 32) idx[t] = -1;
        for (int i = lane; i < seq_width; i += 32) sq[i] = seqs[(size_t)c * seq_width + i];
        __syncwarp();
        for (int s = lane; s < seq_len; s += 32) {
            const int st = max((int)map[s], 0), en = min((int)map[s + 1], T);
            for (int t = st; t < en; ++t) idx[t] = s;
        }
        // stage A: R[s][j][:] = sum over k-mer offsets p of the weight column of base seq[s+p]
        for (int it = lane; it < seq_len * kw * c4; it += 32) {
            const int o4 = it % c4;
            const int sj = it / c4;
            const int j = sj % kw, s = sj / kw;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 *wj = w4 + j * wj4 + o4;
            const int8_t *sp = sq + s;
#pragma unroll 3
            for (int p = 0; p < kmer_len; ++p) {
                const int b = sp[p];
                // N / padding (-1): the all-zero row 4k (branch-free, loads stay independent)
                const int row = (unsigned)b <= 3u ? 4 * p + b : zero_row;
                const float4 w = wj[row * c4];
                acc.x += w.x, acc.y += w.y, acc.z += w.z, acc.w += w.w;
            }
            R4[it] = acc;
        }
        __syncwarp();
        // stage B: y[o][t] = swish(bias[o] + sum_j R[idx[t*stride+j]][j][o]); lanes run along t
        float *yb = y + (size_t)c * y_bstride;
        for (int it = lane; it < c4 * t_out; it += 32) {
            const int o4 = it / t_out, t = it - o4 * t_out;
            float4 acc = make_float4(__ldg(bias + 4 * o4), __ldg(bias + 4 * o4 + 1),
                                     __ldg(bias + 4 * o4 + 2), __ldg(bias + 4 * o4 + 3));
            const int *ip = idx + t * stride;
#pragma unroll 4
            for (int j = 0; j < kw; ++j) {
                const int s = ip[j];
                // positions no base covers (-1): the all-zero row behind R
                const float4 r = R4[s >= 0 ? (s * kw + j) * c4 + o4 : r_zero + o4];
                acc.x += r.x, acc.y += r.y, acc.z += r.z, acc.w += r.w;
            }
            float *yo = yb + (size_t)(4 * o4) * t_out + t;
            yo[0] = swishf_fast(acc.x);
            yo[t_out] = swishf_fast(acc.y);
            yo[2 * (size_t)t_out] = swishf_fast(acc.z);
            yo[3 * (size_t)t_out] = swishf_fast(acc.w);
        }
    }
}

// logits[b][o] = fc_b[o] + sum_k fc_w[o][k] * x[b*bstride + k]; one warp per chunk
__global__ void __launch_bounds__(256)
fc_warp_kernel(const float *__restrict__ x, int64_t bstride, const float *__restrict__ w,
               const float *__restrict__ bias, float *__restrict__ logits, int B, int fc_in,
               int num_out) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    pdl_wait();
    if (b >= B) return;
    const float *xb = x + (size_t)b * bstride;
    for (int o = 0; o < num_out; ++o) {
        const float *wo = w + (size_t)o * fc_in;
        float acc = 0.f;
        for (int k = lane; k < fc_in; k += 32) acc = fmaf(__ldg(wo + k), xb[k], acc);
#pragma unroll
        for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) logits[(size_t)b * num_out + o] = acc + __ldg(bias + o);
    }
}

// ---- host side -------------------------------------------------------------------------------------

inline int pick_ct(int c_out) {
    for (int ct : {64, 32, 16, 8})
        if (c_out % ct == 0) return ct;
    return 0;
}

struct TilePlan {
    int NT = 0, nch = 0, xp = 0, chunk_pitch = 0, ci_slab = 0, xs_floats = 0, threads = 0;
    size_t smem = 0;
};

// Pick positions per thread (NT) and chunks per CTA (nch) for one layer and batch.
TilePlan plan_tile(const rb200_conv_desc &c, int ct, int B, int t_in, int t_out, int sm_count) {
    TilePlan best;
    double best_cost = 0;
    const int lanes = 32 * (TILE_WARPS / (ct / 8));
    const int xp = (t_in + 3) & ~3;
    const int ci_slab = max(1, min(c.c_in, SLAB_FLOATS / (c.kw * ct)));
    const size_t w_bytes = (size_t)2 * ci_slab * c.kw * ct * 4;
    const int threads = TILE_THREADS;
    static const int nts_s1[] = {4, 5, 6, 8}, nts_sn[] = {4, 6};
    const int *nts = c.stride == 1 ? nts_s1 : nts_sn;
    const int n_nts = c.stride == 1 ? 4 : 2;
    for (int k = 0; k < n_nts; ++k) {
        const int NT = nts[k];
        const int ntile = (t_out + NT - 1) / NT;
        if (ntile > lanes) continue;
        int skew = ((ntile * NT * c.stride - c.c_in * xp) % 32 + 32) % 32;
        skew &= ~3;
        const int chunk_pitch = c.c_in * xp + skew;
        for (int nch = 1; nch <= lanes / ntile && nch <= B; ++nch) {
            // tail of 64 floats: the windows of unused positions overrun the last row
            const int xs_floats = nch * chunk_pitch + 64;
            const size_t smem = (size_t)xs_floats * 4 + w_bytes;
            if (smem > SMEM_CAP) break;
            const int resident = smem * 2 + 2048 <= 227 * 1024 ? 2 : 1;
            const long ctas = ((long)B + nch - 1) / nch * (c.c_out / ct);
            // the busiest SM runs per_sm CTAs, each for a time ~ NT (lanes are (chunk, position
            // tile) pairs, filled or not); two co-resident CTAs overlap a little
            const long per_sm = (ctas + sm_count - 1) / sm_count;
            const double cost = (double)per_sm * (NT + 0.75) / (resident == 2 && per_sm >= 2 ? 1.25 : 1.0);
            if (best.NT == 0 || cost < best_cost - 1e-9 ||
                (cost < best_cost + 1e-9 && smem < best.smem)) {
                best_cost = cost;
                best.NT = NT, best.nch = nch, best.xp = xp, best.chunk_pitch = chunk_pitch;
                best.ci_slab = ci_slab, best.smem = smem, best.xs_floats = xs_floats;
                best.threads = threads;
            }
        }
    }
    return best;
}

template <int KW, int S, int NT>
int launch_tile(const TileArgs &a, dim3 grid, int threads, size_t smem, int sm_count,
                cudaStream_t stream) {
    RB200_CUDA_TRY(cudaFuncSetAttribute(conv_tile_kernel<KW, S, NT>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_CAP));
    const bool pdl = (long)grid.x * grid.y > 2L * sm_count;
    RB200_CUDA_TRY(launch_pdl(pdl, conv_tile_kernel<KW, S, NT>, grid, threads, smem, stream, a));
    return RB200_OK;
}

template <int KW, int S>
int launch_tile_nt(int NT, const TileArgs &a, dim3 grid, int threads, size_t smem, int sm_count,
                   cudaStream_t stream) {
    switch (NT) {
        case 4: return launch_tile<KW, S, 4>(a, grid, threads, smem, sm_count, stream);
        case 6: return launch_tile<KW, S, 6>(a, grid, threads, smem, sm_count, stream);
        case 5:
            if constexpr (S == 1) return launch_tile<KW, S, 5>(a, grid, threads, smem, sm_count, stream);
            break;
        case 8:
            if constexpr (S == 1) return launch_tile<KW, S, 8>(a, grid, threads, smem, sm_count, stream);
            break;
    }
    return RB200_ERR_UNSUPPORTED;
}

}  // namespace

struct TiledWeights {
    float *dev = nullptr;
    // float offsets into dev, -1 when the layer has no tiled form
    int64_t sig[RB200_MAX_CONVS], seq[RB200_MAX_CONVS], mrg[RB200_MAX_CONVS];
    int64_t gather = -1;
    int gather_pitch = 0;
};

static bool tile_shape_ok(const rb200_conv_desc &c) {
    const bool kws = (c.kw == 5 && c.stride == 1) || (c.kw == 11 && c.stride == 1) ||
                     (c.kw == 9 && c.stride == 3) || (c.kw == 13 && c.stride == 3) ||
                     (c.kw == 3 && c.stride == 2);
    return kws && pick_ct(c.c_out) != 0;
}

int tiled_create(rb200_model *m, const float *blob) {
    const rb200_model_desc &d = m->desc;
    TiledWeights *tw = new TiledWeights();
    std::vector<float> host;
    auto add_conv = [&](const rb200_conv_desc &c) -> int64_t {
        if (!tile_shape_ok(c)) return -1;
        const int ct = pick_ct(c.c_out);
        const int64_t off = (int64_t)host.size();
        host.resize(host.size() + (size_t)c.c_out * c.c_in * c.kw);
        float *dst = host.data() + off;
        const float *w = blob + c.w_off;  // [co][ci][kw]
        for (int co = 0; co < c.c_out; ++co)
            for (int ci = 0; ci < c.c_in; ++ci)
                for (int j = 0; j < c.kw; ++j)
                    dst[(((size_t)(co / ct) * c.c_in + ci) * c.kw + j) * ct + co % ct] =
                        w[((size_t)co * c.c_in + ci) * c.kw + j];
        while (host.size() % 4) host.push_back(0.f);  // keep every layer 16-byte aligned
        return off;
    };
    for (int i = 0; i < RB200_MAX_CONVS; ++i) tw->sig[i] = tw->seq[i] = tw->mrg[i] = -1;
    for (int i = 0; i < d.n_sig_conv; ++i) tw->sig[i] = add_conv(d.sig_conv[i]);
    for (int i = 0; i < d.n_seq_conv; ++i) tw->seq[i] = add_conv(d.seq_conv[i]);
    for (int i = 0; i < d.n_merge_conv; ++i) tw->mrg[i] = add_conv(d.merge_conv[i]);
    if (d.n_seq_conv > 0) {
        const rb200_conv_desc &c = d.seq_conv[0];
        if (c.c_in == 4 * d.kmer_len && c.c_out % 4 == 0) {
            // consecutive taps continue the bank pattern of the c_out floats one item reads
            const int rows = (c.c_in + 1) * c.c_out;  // + one all-zero row for N / padding bases
            int pad = ((c.c_out - rows) % 32 + 32) % 32;
            pad &= ~3;
            tw->gather_pitch = rows + pad;
            tw->gather = (int64_t)host.size();
            host.resize(host.size() + (size_t)c.kw * tw->gather_pitch, 0.f);
            float *dst = host.data() + tw->gather;
            const float *w = blob + c.w_off;
            for (int o = 0; o < c.c_out; ++o)
                for (int r = 0; r < c.c_in; ++r)
                    for (int j = 0; j < c.kw; ++j)
                        dst[(size_t)j * tw->gather_pitch + (size_t)r * c.c_out + o] =
                            w[((size_t)o * c.c_in + r) * c.kw + j];
        }
    }
    if (!host.empty()) {
        cudaError_t e = cudaMalloc(&tw->dev, host.size() * sizeof(float));
        if (e == cudaSuccess)
            e = cudaMemcpy(tw->dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            set_error("tiled weight upload failed: %s", cudaGetErrorString(e));
            if (tw->dev) cudaFree(tw->dev);
            delete tw;
            return RB200_ERR_CUDA;
        }
    }
    m->tiled = tw;
    return RB200_OK;
}

void tiled_destroy(rb200_model *m) {
    if (!m->tiled) return;
    if (m->tiled->dev) cudaFree(m->tiled->dev);
    delete m->tiled;
    m->tiled = nullptr;
}

// track: 0 sig, 1 seq, 2 merge.  RB200_ERR_UNSUPPORTED = no tiled form for this layer / shape.
int tiled_conv(rb200_model *m, int track, int layer, const rb200_conv_desc &c, const float *x,
               int64_t x_bstride, int t_in, float *y, int64_t y_bstride, int B,
               cudaStream_t stream) {
    const TiledWeights *tw = m->tiled;
    if (!tw) return RB200_ERR_UNSUPPORTED;
    const int64_t off = track == 0 ? tw->sig[layer] : track == 1 ? tw->seq[layer] : tw->mrg[layer];
    if (off < 0) return RB200_ERR_UNSUPPORTED;
    const int t_out = t_in < c.kw ? 0 : (t_in - c.kw) / c.stride + 1;
    if (t_out <= 0) return RB200_ERR_UNSUPPORTED;
    const int ct = pick_ct(c.c_out);
    const TilePlan p = plan_tile(c, ct, B, t_in, t_out, m->sm_count);
    if (p.NT == 0) return RB200_ERR_UNSUPPORTED;
    TileArgs a;
    a.x = x, a.x_bstride = x_bstride, a.wt = tw->dev + off, a.bias = m->blob_dev + c.b_off;
    a.y = y, a.y_bstride = y_bstride;
    a.B = B, a.c_in = c.c_in, a.t_in = t_in, a.c_out = c.c_out, a.t_out = t_out;
    a.nch = p.nch, a.xp = p.xp, a.chunk_pitch = p.chunk_pitch, a.ct = ct, a.ci_slab = p.ci_slab;
    a.xs_floats = p.xs_floats;
    const dim3 grid((B + p.nch - 1) / p.nch, c.c_out / ct);
    int rc = RB200_ERR_UNSUPPORTED;
    if (c.kw == 5 && c.stride == 1) rc = launch_tile_nt<5, 1>(p.NT, a, grid, p.threads, p.smem, m->sm_count, stream);
    else if (c.kw == 11 && c.stride == 1) rc = launch_tile_nt<11, 1>(p.NT, a, grid, p.threads, p.smem, m->sm_count, stream);
    else if (c.kw == 9 && c.stride == 3) rc = launch_tile_nt<9, 3>(p.NT, a, grid, p.threads, p.smem, m->sm_count, stream);
    else if (c.kw == 13 && c.stride == 3) rc = launch_tile_nt<13, 3>(p.NT, a, grid, p.threads, p.smem, m->sm_count, stream);
    else if (c.kw == 3 && c.stride == 2) rc = launch_tile_nt<3, 2>(p.NT, a, grid, p.threads, p.smem, m->sm_count, stream);
    if (rc == RB200_OK) m->launches++;
    return rc;
}

// per-warp scratch of seq1_gather_kernel in floats (R, idx, sq), 16-byte granular
static int gather_warp_floats(const rb200_model *m, int seq_width, int map_width, int T) {
    const TiledWeights *tw = m->tiled;
    if (!tw || tw->gather < 0) return 0;
    const rb200_conv_desc &c = m->desc.seq_conv[0];
    const int lmax = min(map_width - 1, seq_width - m->desc.kmer_len + 1);
    if (lmax <= 0 || T < c.kw) return 0;
    const size_t bytes = (((size_t)lmax * c.kw + 1) * c.c_out + T) * 4 + (size_t)seq_width;
    return (int)((bytes + 15) / 16 * 4);
}

// long mappings (per-read Lmax far above chunk_len / 5) do not fit: dense encode + conv instead
bool tiled_gather_ok(const rb200_model *m, int seq_width, int map_width, int T) {
    const int wf = gather_warp_floats(m, seq_width, map_width, T);
    if (wf == 0) return false;
    const rb200_conv_desc &c = m->desc.seq_conv[0];
    return ((size_t)c.kw * m->tiled->gather_pitch + (size_t)wf) * 4 <= SMEM_CAP;
}

int tiled_seq1_gather(rb200_model *m, const int8_t *seqs, int seq_width, const int16_t *maps,
                      int map_width, const int16_t *lens, int B, int T, float *y, int64_t y_bstride,
                      cudaStream_t stream) {
    if (!tiled_gather_ok(m, seq_width, map_width, T)) return RB200_ERR_UNSUPPORTED;
    const TiledWeights *tw = m->tiled;
    const rb200_conv_desc &c = m->desc.seq_conv[0];
    const int t_out = (T - c.kw) / c.stride + 1;
    const int lmax = min(map_width - 1, seq_width - m->desc.kmer_len + 1);
    const int wf = gather_warp_floats(m, seq_width, map_width, T);
    const size_t w_bytes = (size_t)c.kw * tw->gather_pitch * 4;
    // as many warps (chunks in flight) per CTA as shared memory allows, at most 8; spread the
    // batch over all SMs before stacking chunks on one warp
    int nwarp = (int)min((size_t)8, (SMEM_CAP - w_bytes) / ((size_t)wf * 4));
    nwarp = max(1, min(nwarp, (B + m->sm_count - 1) / m->sm_count));
    const size_t smem = w_bytes + (size_t)nwarp * wf * 4;
    RB200_CUDA_TRY(cudaFuncSetAttribute(seq1_gather_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_CAP));
    const int resident = (int)max((size_t)1, min((size_t)4, (size_t)(220 * 1024) / (smem + 1024)));
    const int grid = min((B + nwarp - 1) / nwarp, m->sm_count * resident);
    RB200_CUDA_TRY(launch_pdl(false, seq1_gather_kernel, dim3(grid), 32 * nwarp, smem, stream, seqs, seq_width,
                              maps, map_width, lens, (const float *)(tw->dev + tw->gather),
                              tw->gather_pitch, (const float *)(m->blob_dev + c.b_off), y, y_bstride, B,
                              T, m->desc.kmer_len, c.kw, c.stride, c.c_out, t_out, lmax, wf));
    m->launches++;
    return RB200_OK;
}

int tiled_fc(rb200_model *m, const float *x, int64_t bstride, float *logits, int B,
             cudaStream_t stream) {
    const rb200_model_desc &d = m->desc;
    RB200_CUDA_TRY(launch_pdl((B + 7) / 8 > 8 * m->sm_count, fc_warp_kernel, dim3((B + 7) / 8), 256, 0, stream, x, bstride,
                              (const float *)(m->blob_dev + d.fc_w_off),
                              (const float *)(m->blob_dev + d.fc_b_off), logits, B, d.fc_in,
                              d.num_out));
    m->launches++;
    return RB200_OK;
}

}  // namespace rb200
