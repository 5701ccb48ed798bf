// Fused sm_100a kernels for ConvLSTM_w_ref (size 64) on the reference's COMPACT chunk arrays:
// the k-mer one-hot tensor is never materialised.
//
//   K1  front   : move-table expansion + seq_conv1 in gather-add form (one-hot conv == sum of
//                 weight columns), sig_conv1..3, seq_conv2  -> cat [B][T3][128] channel-last
//   K2  merge   : merge_conv1 (implicit GEMM M=64,K=640) + LSTM1 input projection (M=256,K=64)
//                 -> xp [B][TM][256]
//   K3  lstm    : LSTM1 recurrence with W_hh resident in registers, the single needed step of the
//                 time-reversed LSTM2, fc with warp-shuffle reduction -> logits
//
// Reference semantics: models/ConvLSTM_w_ref.py:39-58 with eval BatchNorm folded into the convs.
// All arithmetic is fp32 FMA (packed FFMA2, `fma.rn.f32x2`, two output channels per instruction),
// fp32 accumulate.  Activations are channel-last so that 4 input channels arrive per LDS.128 and the
// two halves of an FFMA2 are two adjacent output channels; weights are staged in shared memory
// k-major by bulk asynchronous copies (TMA, cp.async.bulk + mbarrier), shared by every chunk of the
// CTA and read as warp-wide broadcasts.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "rb200_internal.cuh"

namespace rb200 {

namespace {

constexpr int SIZE = 64;          // channel width this file is specialised for
constexpr int XP = 2 * SIZE + 4;  // cat row pitch in floats (128 channels + 4 pad)
constexpr int MP = SIZE + 4;      // merge-out row pitch in shared memory
constexpr int QP = 20;            // 16-channel row pitch (s2 / q1) in shared memory
constexpr int NR1 = 7;            // K1: output positions per thread
constexpr int NR2 = 6;            // K2: output positions per thread
constexpr int MAX_CL = 8;         // chunks per CTA (lane = tb * CL + chunk)
constexpr int MRW = 8;                        // output channels per warp in the implicit GEMMs
constexpr int NWARPS = SIZE / MRW;            // 8 warps (16 warps x 4 channels measured no faster)
constexpr int THREADS = NWARPS * 32;          // 256
constexpr int XR = 256 / NWARPS / 2;          // x-projection: gate rows per warp per pass (2 passes)
constexpr int SLAB_C = 32;                              // merge conv: input channels per slab
constexpr int SLAB_FLOATS = 5 * SLAB_C * SIZE;          // 10240 floats = 40 KB
constexpr int N_SLABS = 2 * SIZE / SLAB_C;              // 4
constexpr int KW_SIG1 = 5, KW_SIG2 = 5, KW_SIG3 = 9, KW_SEQ1 = 5, KW_SEQ2 = 13, KW_MRG = 5;
constexpr int GROW = 20;  // pitch (floats) of 16-float gather rows: 20 r mod 32 spreads rows over all banks

// ---- PTX helpers: mbarrier + bulk async copy (TMA) --------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared bulk copy, completion counted on `bar`; bytes % 16 == 0, 16 B aligned both ends
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_addr(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
        : "memory");
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor in the stream is still draining; everything before
// pdl_wait() (barrier init, TMA weight loads, register-resident weights) overlaps that tail, and
// nothing produced by the predecessor is touched before it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ float2 ffma2(float2 a, float b, float2 c) {
    // SASS: FFMA2 Rd, Ra.F32x2.HI_LO, Rb.F32, Rc.F32x2.HI_LO  (scalar operand broadcast to both halves)
    return __ffma2_rn(a, make_float2(b, b), c);
}

struct FrontOffsets {  // float offsets inside the K1 weight blob (all multiples of 4)
    int w_sig1, b_sig1, w_sig2, b_sig2, w_sig3, b_sig3, w_seq1, z_seq1, b_seq1, w_seq2, b_seq2, total;
};

__host__ __device__ inline FrontOffsets front_offsets(int kmer_len, bool tc = false) {
    FrontOffsets o;
    int cur = 0;
    auto take = [&](int n) {
        int at = cur;
        cur += (n + 3) & ~3;
        return at;
    };
    o.w_sig1 = take(KW_SIG1 * 4);            // [j][co]
    o.b_sig1 = take(4);
    o.w_sig2 = take(KW_SIG2 * 4 * 16);       // [j][ci][co]
    o.b_sig2 = take(16);
    o.w_sig3 = take(tc ? 0 : KW_SIG3 * 16 * SIZE);    // [j*16+ci][co] (tensor-core K1 streams its own tiles)
    o.b_sig3 = take(SIZE);
    o.w_seq1 = take(KW_SEQ1 * kmer_len * 4 * GROW);  // [j][p][base][co], row pitch GROW
    o.z_seq1 = take(GROW);                           // all-zero row: target of N bases / uncovered samples
    o.b_seq1 = take(16);
    o.w_seq2 = take(tc ? 0 : KW_SEQ2 * 16 * SIZE);    // [j*16+ci][co]
    o.b_seq2 = take(SIZE);
    o.total = cur;
    return o;
}

struct Geometry {
    int T, T1, T2, T3, TM;  // chunk_len, after sig_conv1, sig_conv2, sig_conv3 (= after seq_conv2), merge
    int Q1;                 // after seq_conv1
    int NB1, NB2, CL;       // position blocks per chunk in K1 / K2, max chunks per CTA
    int s2_stride, q1_stride, act_stride;  // per-chunk strides (floats) of the 16-channel tiles
    int cat_stride;                        // per-chunk stride (floats) of cat rows in HBM and smem
    bool ok;
};

__host__ __device__ inline int odd_quads(int floats) {
    // pad so that stride/4 is odd: 8 lanes reading 16 B at this stride hit 8 different bank groups
    int s = (floats + 3) & ~3;
    if (((s >> 2) & 1) == 0) s += 4;
    return s;
}

__host__ __device__ inline Geometry make_geometry(int T) {
    Geometry g;
    g.T = T;
    g.T1 = T - (KW_SIG1 - 1);
    g.T2 = g.T1 - (KW_SIG2 - 1);
    g.T3 = g.T2 >= KW_SIG3 ? (g.T2 - KW_SIG3) / 3 + 1 : 0;
    g.Q1 = T - (KW_SEQ1 - 1);
    const int q2 = g.Q1 >= KW_SEQ2 ? (g.Q1 - KW_SEQ2) / 3 + 1 : 0;
    g.TM = g.T3 - (KW_MRG - 1);
    g.ok = g.T3 > 0 && q2 == g.T3 && g.TM > 0;
    g.NB1 = (g.T3 + NR1 - 1) / NR1;
    g.NB2 = (g.TM + NR2 - 1) / NR2;
    const int nb = g.NB1 > g.NB2 ? g.NB1 : g.NB2;
    g.CL = nb > 0 ? 32 / nb : 0;
    if (g.CL > MAX_CL) g.CL = MAX_CL;
    if (g.CL < 1) g.ok = false;
    g.s2_stride = odd_quads(g.T2 * QP);
    g.q1_stride = odd_quads(g.Q1 * QP);
    g.act_stride = g.s2_stride > g.q1_stride ? g.s2_stride : g.q1_stride;
    g.cat_stride = odd_quads(g.T3 * XP);
    return g;
}

// =================================================================================================
// K1: front
// =================================================================================================
struct K1Smem {  // float offsets in dynamic shared memory
    int weights, sig, s1, act, sidx, seq, map, len, total_bytes;
};

__host__ __device__ inline K1Smem k1_smem(const Geometry &g, int kmer_len, int seq_width,
                                          int map_width) {
    K1Smem s;
    const FrontOffsets fo = front_offsets(kmer_len);
    int cur = 4;  // floats 0..1 hold the mbarrier
    auto take = [&](int n) {
        int at = cur;
        cur += (n + 3) & ~3;
        return at;
    };
    s.weights = take(fo.total);
    s.sig = take(g.CL * g.T);
    s.s1 = take(g.CL * g.T1 * 4);
    s.act = take(g.CL * g.act_stride);
    s.sidx = take((g.CL * g.T + 1) / 2);        // int16 per sample
    s.seq = take((g.CL * seq_width + 3) / 4);   // int8
    s.map = take((g.CL * map_width + 1) / 2);   // int16
    s.len = take(g.CL);
    s.total_bytes = cur * 4;
    return s;
}

// a = hi + lo with hi exactly representable in TF32 (round to nearest) and lo = a - hi (exact in fp32);
// hi*b_hi + lo*b_hi + hi*b_lo on the tensor cores then carries ~2^-21 relative error ("3xTF32").
__device__ __forceinline__ float tf32_hi(float a) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a));
    return __uint_as_float(u);
}
// remainder after the tensor core's own fp32 -> TF32 conversion (truncation of the low 13 bits)
__device__ __forceinline__ float tf32_trunc_lo(float a) {
    return a - __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
}
__device__ __forceinline__ void split_tf32(const float4 &a, float4 &hi, float4 &lo) {
    hi = make_float4(tf32_hi(a.x), tf32_hi(a.y), tf32_hi(a.z), tf32_hi(a.w));
    lo = make_float4(a.x - hi.x, a.y - hi.y, a.z - hi.z, a.w - hi.w);
}

// Implicit-GEMM conv, stride 3, 16 input channels (channel-last, pitch QP) -> 64 output channels.
// Warp w owns output channels [MRW*w, MRW*w+MRW); lane = tb * CL + chunk owns NR1 consecutive steps.
// Taps are processed by residue class rho = j mod 3: taps rho, rho+3, rho+6, ... of output step t
// read input rows 3(t+i)+rho, i.e. a stride-1 window over the decimated rows u = t+i.  One window of
// NR1 + ntaps - 1 rows (LDS.128 each) then serves ntaps * NR1 row uses.
template <int KW, int RHO>
__device__ __forceinline__ void conv16_s3_residue(float2 (&acc)[NR1][MRW / 2], const float *__restrict__ xbase,
                                                  int t0, int T3, const float *__restrict__ ws,
                                                  int m0) {
    constexpr int NT = (KW - RHO + 2) / 3;  // taps in this residue class
    constexpr int NW = NR1 + NT - 1;        // window rows
    const float *xrow[NW];
#pragma unroll
    for (int r = 0; r < NW; ++r) {
        int u = t0 + r;
        if (u > T3 + NT - 2) u = T3 + NT - 2;  // last row any valid output step touches
        xrow[r] = xbase + (3 * u + RHO) * QP;
    }
#pragma unroll 1
    for (int c4 = 0; c4 < 16; c4 += 4) {
        float4 x[NW];
#pragma unroll
        for (int r = 0; r < NW; ++r) x[r] = *reinterpret_cast<const float4 *>(xrow[r] + c4);
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            const int j = RHO + 3 * i;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 *wp =
                    reinterpret_cast<const float4 *>(ws + (j * 16 + c4 + kk) * SIZE + m0);
                float2 w[MRW / 2];
#pragma unroll
                for (int q = 0; q < MRW / 4; ++q) {
                    const float4 wv = wp[q];
                    w[2 * q] = make_float2(wv.x, wv.y);
                    w[2 * q + 1] = make_float2(wv.z, wv.w);
                }
#pragma unroll
                for (int n = 0; n < NR1; ++n) {
                    const float4 xq = x[n + i];
                    const float xv = kk == 0 ? xq.x : kk == 1 ? xq.y : kk == 2 ? xq.z : xq.w;
#pragma unroll
                    for (int p2 = 0; p2 < MRW / 2; ++p2) acc[n][p2] = ffma2(w[p2], xv, acc[n][p2]);
                }
            }
        }
    }
}

template <int KW>
__device__ __forceinline__ void conv16_s3_to_cat(const float *__restrict__ xs, int x_stride,
                                                 const float *__restrict__ ws,
                                                 const float *__restrict__ bias,
                                                 float *__restrict__ cat, int cat_stride,
                                                 int ch_off, int C, int CL, int NB, int T3,
                                                 int tc_rpad) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = warp * MRW;
    int tb = lane / CL, chunk = lane - tb * CL;
    const bool lane_ok = tb < NB && chunk < C;
    if (tb >= NB) tb = NB - 1;
    if (chunk >= C) chunk = C - 1;
    const int t0 = tb * NR1;
    const float *xbase = xs + chunk * x_stride;
    float2 acc[NR1][MRW / 2];
#pragma unroll
    for (int n = 0; n < NR1; ++n)
#pragma unroll
        for (int p = 0; p < MRW / 2; ++p) acc[n][p] = make_float2(0.f, 0.f);
    conv16_s3_residue<KW, 0>(acc, xbase, t0, T3, ws, m0);
    conv16_s3_residue<KW, 1>(acc, xbase, t0, T3, ws, m0);
    conv16_s3_residue<KW, 2>(acc, xbase, t0, T3, ws, m0);
    if (!lane_ok) return;
    float b[MRW];
#pragma unroll
    for (int i = 0; i < MRW; ++i) b[i] = bias[m0 + i];
#pragma unroll
    for (int n = 0; n < NR1; ++n) {
        const int t = t0 + n;
        if (t < T3) {
            float4 o[MRW / 4];
#pragma unroll
            for (int q = 0; q < MRW / 4; ++q)
                o[q] = make_float4(swishf_fast(acc[n][2 * q].x + b[4 * q]),
                                   swishf_fast(acc[n][2 * q].y + b[4 * q + 1]),
                                   swishf_fast(acc[n][2 * q + 1].x + b[4 * q + 2]),
                                   swishf_fast(acc[n][2 * q + 1].y + b[4 * q + 3]));
            if (tc_rpad == 0) {
                float4 *dst = reinterpret_cast<float4 *>(cat + (size_t)chunk * cat_stride + t * XP +
                                                         ch_off + m0);
#pragma unroll
                for (int q = 0; q < MRW / 4; ++q) dst[q] = o[q];
            } else {
                // tensor-core image of this CTA: [channel block][row][32 ch, 128B-swizzled], plain fp32
                // (the tensor core truncates it to TF32 = "hi"; K2-TC derives the "lo" tile on chip)
                const int row = chunk * T3 + t;
#pragma unroll
                for (int q = 0; q < MRW / 4; ++q) {
                    const int ch = ch_off + m0 + 4 * q;
                    const int off = ((ch >> 5) * tc_rpad + row) * 32 + ((((ch & 31) >> 2) ^ (row & 7)) << 2);
                    *reinterpret_cast<float4 *>(cat + off) = o[q];
                }
            }
        }
    }
}

// seq_conv1 on the (virtual) one-hot input = gather-add of weight columns:
// q1[t][o] = swish(b[o] + sum_{j<5} sum_{p<k} W[o][4p + base(t+j, p)][j]),  base(t, p) = seq[sidx[t] + p].
// -1 bases and uncovered samples contribute nothing (encoded_kmers.pyx:39-40): they are redirected to
// an all-zero table row so that the 5*k gathers of one output step are branch-free and independent.
// One thread per (chunk, t): 16 channels as 8 packed float2 sums.  KT = compile-time kmer_len (0 = runtime).
template <int KT>
__device__ __forceinline__ void seq1_gather(const float *__restrict__ w, const float *__restrict__ b,
                                            int zero_off, const int8_t *__restrict__ seq_s,
                                            const int16_t *__restrict__ sidx_s,
                                            float *__restrict__ act_s, int seq_width, int C, int T,
                                            int Q1, int q1_stride, int kmer_rt, int qp = QP) {
    const int K = KT > 0 ? KT : kmer_rt;
    for (int i = threadIdx.x; i < C * Q1; i += THREADS) {
        const int c = i / Q1, t = i - c * Q1;
        float2 a[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) a[o] = make_float2(b[2 * o], b[2 * o + 1]);
        const int8_t *sq = seq_s + c * seq_width;
#pragma unroll
        for (int j = 0; j < KW_SEQ1; ++j) {
            const int s = sidx_s[c * T + t + j];
            const int8_t *sp = sq + (s < 0 ? 0 : s);
            const int joff = j * K * 4 * GROW;
#pragma unroll
            for (int p = 0; p < (KT > 0 ? KT : 16); ++p) {
                if (KT == 0 && p >= K) break;
                const int base = sp[p];
                const bool ok = s >= 0 && base >= 0 && base <= 3;
                const int off = ok ? joff + (p * 4 + base) * GROW : zero_off;
                const float4 *wv = reinterpret_cast<const float4 *>(w + off);
                const float4 v0 = wv[0], v1 = wv[1], v2 = wv[2], v3 = wv[3];
                a[0] = __fadd2_rn(a[0], make_float2(v0.x, v0.y));
                a[1] = __fadd2_rn(a[1], make_float2(v0.z, v0.w));
                a[2] = __fadd2_rn(a[2], make_float2(v1.x, v1.y));
                a[3] = __fadd2_rn(a[3], make_float2(v1.z, v1.w));
                a[4] = __fadd2_rn(a[4], make_float2(v2.x, v2.y));
                a[5] = __fadd2_rn(a[5], make_float2(v2.z, v2.w));
                a[6] = __fadd2_rn(a[6], make_float2(v3.x, v3.y));
                a[7] = __fadd2_rn(a[7], make_float2(v3.z, v3.w));
            }
        }
        float4 *dst = reinterpret_cast<float4 *>(act_s + c * q1_stride + t * qp);
#pragma unroll
        for (int o = 0; o < 4; ++o)
            dst[o] = make_float4(swishf_fast(a[2 * o].x), swishf_fast(a[2 * o].y),
                                 swishf_fast(a[2 * o + 1].x), swishf_fast(a[2 * o + 1].y));
    }
}

// Two-stage form of the seq_conv1 gather: every sample covered by the same base shares its k-mer, so
// first sum the k weight columns once per (base, tap) - Gs[c][s][j][16] - and then add five of those
// rows per output step.  Cuts the shared-memory gather traffic ~9x versus gathering 5*k columns per
// output step (the direct form is bandwidth-bound on shared memory: 45 x 64 B per step).
template <int KT>
__device__ __forceinline__ void seq1_gather_two_stage(
    const float *__restrict__ w, const float *__restrict__ b, int zero_off,
    const int8_t *__restrict__ seq_s, const int16_t *__restrict__ sidx_s, const int *__restrict__ len_s,
    float *__restrict__ gs, float *__restrict__ act_s, int seq_width, int LM, int C, int T, int Q1,
    int q1_stride, int kmer_rt, int qp) {
    const int K = KT > 0 ? KT : kmer_rt;
    for (int i = threadIdx.x; i < C * LM * KW_SEQ1; i += THREADS) {
        const int j = i % KW_SEQ1;
        const int cs = i / KW_SEQ1;
        const int c = cs / LM, sb = cs - c * LM;
        if (sb >= len_s[c]) continue;
        float2 a[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) a[o] = make_float2(0.f, 0.f);
        const int8_t *sp = seq_s + c * seq_width + sb;
        const int joff = j * K * 4 * GROW;
#pragma unroll
        for (int p = 0; p < (KT > 0 ? KT : 16); ++p) {
            if (KT == 0 && p >= K) break;
            const int base = sp[p];
            const int off = (base >= 0 && base <= 3) ? joff + (p * 4 + base) * GROW : zero_off;
            const float4 *wv = reinterpret_cast<const float4 *>(w + off);
            const float4 v0 = wv[0], v1 = wv[1], v2 = wv[2], v3 = wv[3];
            a[0] = __fadd2_rn(a[0], make_float2(v0.x, v0.y));
            a[1] = __fadd2_rn(a[1], make_float2(v0.z, v0.w));
            a[2] = __fadd2_rn(a[2], make_float2(v1.x, v1.y));
            a[3] = __fadd2_rn(a[3], make_float2(v1.z, v1.w));
            a[4] = __fadd2_rn(a[4], make_float2(v2.x, v2.y));
            a[5] = __fadd2_rn(a[5], make_float2(v2.z, v2.w));
            a[6] = __fadd2_rn(a[6], make_float2(v3.x, v3.y));
            a[7] = __fadd2_rn(a[7], make_float2(v3.z, v3.w));
        }
        float4 *dst = reinterpret_cast<float4 *>(gs + (size_t)i * GROW);
#pragma unroll
        for (int o = 0; o < 4; ++o)
            dst[o] = make_float4(a[2 * o].x, a[2 * o].y, a[2 * o + 1].x, a[2 * o + 1].y);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * Q1; i += THREADS) {
        const int c = i / Q1, t = i - c * Q1;
        float2 a[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) a[o] = make_float2(b[2 * o], b[2 * o + 1]);
#pragma unroll
        for (int j = 0; j < KW_SEQ1; ++j) {
            const int sb = sidx_s[c * T + t + j];
            if (sb < 0) continue;  // sample not covered by any base: no one-hot entries
            const float4 *gv = reinterpret_cast<const float4 *>(gs + ((size_t)(c * LM + sb) * KW_SEQ1 + j) * GROW);
            const float4 v0 = gv[0], v1 = gv[1], v2 = gv[2], v3 = gv[3];
            a[0] = __fadd2_rn(a[0], make_float2(v0.x, v0.y));
            a[1] = __fadd2_rn(a[1], make_float2(v0.z, v0.w));
            a[2] = __fadd2_rn(a[2], make_float2(v1.x, v1.y));
            a[3] = __fadd2_rn(a[3], make_float2(v1.z, v1.w));
            a[4] = __fadd2_rn(a[4], make_float2(v2.x, v2.y));
            a[5] = __fadd2_rn(a[5], make_float2(v2.z, v2.w));
            a[6] = __fadd2_rn(a[6], make_float2(v3.x, v3.y));
            a[7] = __fadd2_rn(a[7], make_float2(v3.z, v3.w));
        }
        float4 *dst = reinterpret_cast<float4 *>(act_s + c * q1_stride + t * qp);
#pragma unroll
        for (int o = 0; o < 4; ++o)
            dst[o] = make_float4(swishf_fast(a[2 * o].x), swishf_fast(a[2 * o].y),
                                 swishf_fast(a[2 * o + 1].x), swishf_fast(a[2 * o + 1].y));
    }
}

// Phases shared by both K1 variants.  k1_stage_inputs: stage the compact inputs, move-table expansion,
// wait for the front weight blob.  k1_sig12: sig_conv1, sig_conv2 (output s2 channel-last with row
// pitch `qp`, chunk stride `s2_stride`).
__device__ __forceinline__ void k1_stage_inputs(
    const float *__restrict__ sigs, const int8_t *__restrict__ seqs, int seq_width,
    const int16_t *__restrict__ maps, int map_width, const int16_t *__restrict__ lens, int chunk0, int C,
    int T, int kmer_len, const Geometry &g, const FrontOffsets &fo, uint64_t *bar, const float *wsm,
    float *sig_s, float *s1_s, float *act_s, int16_t *sidx_s, int8_t *seq_s, int16_t *map_s, int *len_s,
    int s2_stride, int qp, bool dense = false) {
    const int tid = threadIdx.x;
    if (dense) {  // dense one-hot interface: only the signal is staged, seq_conv1 ran as its own kernel
        for (int i = tid; i < C * T; i += THREADS) sig_s[i] = sigs[(size_t)chunk0 * T + i];
        mbar_wait(bar, 0);
        __syncthreads();
        return;
    }
    // ---- stage the compact inputs of this CTA's chunks ------------------------------------------
    for (int i = tid; i < C * T; i += THREADS) {
        sig_s[i] = sigs[(size_t)chunk0 * T + i];
        sidx_s[i] = -1;
    }
    for (int i = tid; i < C; i += THREADS) {
        int L = lens[chunk0 + i];
        L = max(0, min(L, min(map_width - 1, seq_width - kmer_len + 1)));
        len_s[i] = L;
    }
    __syncthreads();
    for (int i = tid; i < C * seq_width; i += THREADS) {
        const int c = i / seq_width, s = i - c * seq_width;
        // padding past seq_len + kmer_len - 1 is uninitialised in the reference's arrays: never read
        seq_s[i] = s < len_s[c] + kmer_len - 1 ? seqs[(size_t)(chunk0 + c) * seq_width + s] : (int8_t)-1;
    }
    for (int i = tid; i < C * map_width; i += THREADS) {
        const int c = i / map_width, s = i - c * map_width;
        map_s[i] = s <= len_s[c] ? maps[(size_t)(chunk0 + c) * map_width + s] : (int16_t)0;
    }
    __syncthreads();
    // ---- move-table expansion: sidx[t] = index of the base whose dwell covers sample t ----------
    for (int i = tid; i < C * (map_width - 1); i += THREADS) {
        const int c = i / (map_width - 1), s = i - c * (map_width - 1);
        if (s < len_s[c]) {
            const int st = max((int)map_s[c * map_width + s], 0);
            const int en = min((int)map_s[c * map_width + s + 1], T);
            for (int t = st; t < en; ++t) sidx_s[c * T + t] = (int16_t)s;
        }
    }
    mbar_wait(bar, 0);  // weights have landed
    __syncthreads();
}

__device__ __forceinline__ void k1_sig12(int C, int T, const Geometry &g, const FrontOffsets &fo,
                                         const float *wsm, const float *sig_s, float *s1_s, float *act_s,
                                         int s2_stride, int qp) {
    const int tid = threadIdx.x;
    // ---- sig_conv1 (1 -> 4, k5) -------------------------------------------------------------------
    {
        const float *w = wsm + fo.w_sig1;  // [j][co]
        const float *b = wsm + fo.b_sig1;
        for (int i = tid; i < C * g.T1; i += THREADS) {
            const int c = i / g.T1, t = i - c * g.T1;
            const float *x = sig_s + c * T + t;
            float4 a = *reinterpret_cast<const float4 *>(b);
#pragma unroll
            for (int j = 0; j < KW_SIG1; ++j) {
                const float xv = x[j];
                const float4 wv = *reinterpret_cast<const float4 *>(w + 4 * j);
                a.x = fmaf(wv.x, xv, a.x);
                a.y = fmaf(wv.y, xv, a.y);
                a.z = fmaf(wv.z, xv, a.z);
                a.w = fmaf(wv.w, xv, a.w);
            }
            a.x = swishf_fast(a.x);
            a.y = swishf_fast(a.y);
            a.z = swishf_fast(a.z);
            a.w = swishf_fast(a.w);
            *reinterpret_cast<float4 *>(s1_s + (size_t)i * 4) = a;
        }
    }
    __syncthreads();
    // ---- sig_conv2 (4 -> 16, k5): one thread per (chunk, t, 8-channel half) ----------------------
    {
        const float *w = wsm + fo.w_sig2;  // [j][ci][co]
        const float *b = wsm + fo.b_sig2;
        for (int i = tid; i < C * g.T2 * 2; i += THREADS) {
            const int half = i & 1;
            const int ct = i >> 1;
            const int c = ct / g.T2, t = ct - c * g.T2;
            float acc[8];
#pragma unroll
            for (int o = 0; o < 8; ++o) acc[o] = b[half * 8 + o];
#pragma unroll
            for (int j = 0; j < KW_SIG2; ++j) {
                const float4 xv = *reinterpret_cast<const float4 *>(s1_s + (size_t)(c * g.T1 + t + j) * 4);
                const float xs4[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) {
                    const float4 *wp = reinterpret_cast<const float4 *>(w + (j * 4 + ci) * 16 + half * 8);
                    const float4 wa = wp[0], wb = wp[1];
                    acc[0] = fmaf(wa.x, xs4[ci], acc[0]);
                    acc[1] = fmaf(wa.y, xs4[ci], acc[1]);
                    acc[2] = fmaf(wa.z, xs4[ci], acc[2]);
                    acc[3] = fmaf(wa.w, xs4[ci], acc[3]);
                    acc[4] = fmaf(wb.x, xs4[ci], acc[4]);
                    acc[5] = fmaf(wb.y, xs4[ci], acc[5]);
                    acc[6] = fmaf(wb.z, xs4[ci], acc[6]);
                    acc[7] = fmaf(wb.w, xs4[ci], acc[7]);
                }
            }
            float4 *dst = reinterpret_cast<float4 *>(act_s + c * s2_stride + t * qp + half * 8);
            dst[0] = make_float4(swishf_fast(acc[0]), swishf_fast(acc[1]), swishf_fast(acc[2]), swishf_fast(acc[3]));
            dst[1] = make_float4(swishf_fast(acc[4]), swishf_fast(acc[5]), swishf_fast(acc[6]), swishf_fast(acc[7]));
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void k1_stage_and_sig12(
    const float *__restrict__ sigs, const int8_t *__restrict__ seqs, int seq_width,
    const int16_t *__restrict__ maps, int map_width, const int16_t *__restrict__ lens, int chunk0, int C,
    int T, int kmer_len, const Geometry &g, const FrontOffsets &fo, uint64_t *bar, const float *wsm,
    float *sig_s, float *s1_s, float *act_s, int16_t *sidx_s, int8_t *seq_s, int16_t *map_s, int *len_s,
    int s2_stride, int qp) {
    k1_stage_inputs(sigs, seqs, seq_width, maps, map_width, lens, chunk0, C, T, kmer_len, g, fo, bar, wsm,
                    sig_s, s1_s, act_s, sidx_s, seq_s, map_s, len_s, s2_stride, qp);
    k1_sig12(C, T, g, fo, wsm, sig_s, s1_s, act_s, s2_stride, qp);
}

__global__ void __launch_bounds__(THREADS, 1)
k1_front_kernel(const float *__restrict__ sigs, const int8_t *__restrict__ seqs, int seq_width,
                const int16_t *__restrict__ maps, int map_width, const int16_t *__restrict__ lens,
                const float *__restrict__ wfront, float *__restrict__ cat, int B, int CPB, int T,
                int kmer_len, int tc_rpad) {
    extern __shared__ __align__(128) float sm[];
    const Geometry g = make_geometry(T);
    const K1Smem lay = k1_smem(g, kmer_len, seq_width, map_width);
    const FrontOffsets fo = front_offsets(kmer_len);
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm);
    float *wsm = sm + lay.weights;
    float *sig_s = sm + lay.sig;
    float *s1_s = sm + lay.s1;
    float *act_s = sm + lay.act;
    int16_t *sidx_s = reinterpret_cast<int16_t *>(sm + lay.sidx);
    int8_t *seq_s = reinterpret_cast<int8_t *>(sm + lay.seq);
    int16_t *map_s = reinterpret_cast<int16_t *>(sm + lay.map);
    int *len_s = reinterpret_cast<int *>(sm + lay.len);

    const int tid = threadIdx.x;
    const int chunk0 = blockIdx.x * CPB;
    const int C = min(CPB, B - chunk0);
    const int CL = g.CL;

    pdl_launch_dependents();  // K2 may begin its weight prefetch as soon as SMs free up
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)fo.total * 4u;
        mbar_expect_tx(bar, bytes);
        bulk_g2s(wsm, wfront, bytes, bar);
    }
    pdl_wait();  // inputs (and the cat buffer we overwrite) belong to earlier work in the stream
    k1_stage_and_sig12(sigs, seqs, seq_width, maps, map_width, lens, chunk0, C, T, kmer_len, g, fo, bar, wsm,
                       sig_s, s1_s, act_s, sidx_s, seq_s, map_s, len_s, g.s2_stride, QP);
    // classic layout: chunk-major rows; tensor-core layout: one pre-swizzled hi/lo image per CTA
    float *cat_cta = tc_rpad ? cat + (size_t)blockIdx.x * (4 * tc_rpad * 32)
                             : cat + (size_t)chunk0 * g.cat_stride;
    // ---- sig_conv3 (16 -> 64, k9, stride 3) -> cat[:, :, 0:64] -----------------------------------
    conv16_s3_to_cat<KW_SIG3>(act_s, g.s2_stride, wsm + fo.w_sig3, wsm + fo.b_sig3, cat_cta,
                              g.cat_stride, 0, C, CL, g.NB1, g.T3, tc_rpad);
    __syncthreads();
    // ---- seq_conv1 on the (virtual) one-hot input = gather-add of weight columns -----------------
    if (kmer_len == 9)
        seq1_gather<9>(wsm + fo.w_seq1, wsm + fo.b_seq1, fo.z_seq1 - fo.w_seq1, seq_s, sidx_s, act_s,
                       seq_width, C, T, g.Q1, g.q1_stride, kmer_len);
    else
        seq1_gather<0>(wsm + fo.w_seq1, wsm + fo.b_seq1, fo.z_seq1 - fo.w_seq1, seq_s, sidx_s, act_s,
                       seq_width, C, T, g.Q1, g.q1_stride, kmer_len);
    __syncthreads();
    // ---- seq_conv2 (16 -> 64, k13, stride 3) -> cat[:, :, 64:128] --------------------------------
    conv16_s3_to_cat<KW_SEQ2>(act_s, g.q1_stride, wsm + fo.w_seq2, wsm + fo.b_seq2, cat_cta,
                              g.cat_stride, SIZE, C, CL, g.NB1, g.T3, tc_rpad);
}

// =================================================================================================
// K0 (dense interface only): seq_conv1 + BN + swish on a materialised one-hot tensor
// =================================================================================================
// model(sigs, enc_kmers) - the reference's own call form - hands over float32 [B][4k][T].  This
// kernel is the honest dense convolution (any float input, not only one-hot) producing the same
// q1 [B][T-4][16] channel-last activations the gather produces on the compact path, so that the rest
// of the fused pipeline (K1-TC from sig/q1, K2-TC, K3) is shared.  One CTA per chunk: the chunk's
// [4k x T] tile arrives by one TMA bulk copy, weights [row][tap][16] by another; one thread per output
// step accumulates 16 channels as 8 FFMA2 pairs.
__global__ void __launch_bounds__(128)
k0_dense_seq1_kernel(const float *__restrict__ enc, const float *__restrict__ w, const float *__restrict__ bias,
                     float *__restrict__ q1, int B, int T, int rows) {
    extern __shared__ __align__(128) float sm0[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm0);
    float *x_s = sm0 + 4;                         // [rows][T]
    float *w_s = x_s + ((rows * T + 3) & ~3);     // [rows][5][16]
    const int c = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t xb = (uint32_t)rows * T * 4u, wb = (uint32_t)rows * KW_SEQ1 * 16 * 4u;
        mbar_expect_tx(bar, xb + wb);
        bulk_g2s(x_s, enc + (size_t)c * rows * T, xb, bar);
        bulk_g2s(w_s, w, wb, bar);
    }
    mbar_wait(bar, 0);
    const int Q1 = T - (KW_SEQ1 - 1);
    for (int t = tid; t < Q1; t += blockDim.x) {
        float2 a[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) a[o] = make_float2(bias[2 * o], bias[2 * o + 1]);
        for (int r = 0; r < rows; ++r) {
            const float *xr = x_s + r * T + t;
#pragma unroll
            for (int j = 0; j < KW_SEQ1; ++j) {
                const float xv = xr[j];
                const float4 *wp = reinterpret_cast<const float4 *>(w_s + (r * KW_SEQ1 + j) * 16);
                const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
                a[0] = ffma2(make_float2(w0.x, w0.y), xv, a[0]);
                a[1] = ffma2(make_float2(w0.z, w0.w), xv, a[1]);
                a[2] = ffma2(make_float2(w1.x, w1.y), xv, a[2]);
                a[3] = ffma2(make_float2(w1.z, w1.w), xv, a[3]);
                a[4] = ffma2(make_float2(w2.x, w2.y), xv, a[4]);
                a[5] = ffma2(make_float2(w2.z, w2.w), xv, a[5]);
                a[6] = ffma2(make_float2(w3.x, w3.y), xv, a[6]);
                a[7] = ffma2(make_float2(w3.z, w3.w), xv, a[7]);
            }
        }
        float4 *dst = reinterpret_cast<float4 *>(q1 + ((size_t)c * Q1 + t) * 16);
#pragma unroll
        for (int o = 0; o < 4; ++o)
            dst[o] = make_float4(swishf_fast(a[2 * o].x), swishf_fast(a[2 * o].y),
                                 swishf_fast(a[2 * o + 1].x), swishf_fast(a[2 * o + 1].y));
    }
}

// =================================================================================================
// K2: merge_conv1 + LSTM1 input projection
// =================================================================================================
struct K2Smem {
    int xs, ws, total_bytes;
};
__host__ __device__ inline K2Smem k2_smem(const Geometry &g) {
    K2Smem s;
    int cur = 8;  // 4 mbarriers
    s.xs = cur;
    int x_floats = g.CL * g.cat_stride;
    const int ms_floats = 192 * MP;  // x-projection input rows (<= 32 lanes * NR2 positions)
    if (x_floats < ms_floats) x_floats = ms_floats;
    cur += (x_floats + 31) & ~31;
    s.ws = cur;
    cur += 2 * SLAB_FLOATS;
    s.total_bytes = cur * 4;
    return s;
}

__global__ void __launch_bounds__(THREADS, 1)
k2_merge_kernel(const float *__restrict__ cat, const float *__restrict__ wslabs,
                const float *__restrict__ bmerge, const float *__restrict__ wih1T,
                const float *__restrict__ b1, float *__restrict__ xp, int B, int CPB, int T) {
    extern __shared__ __align__(128) float sm[];
    const Geometry g = make_geometry(T);
    const K2Smem lay = k2_smem(g);
    uint64_t *bar_x = reinterpret_cast<uint64_t *>(sm);
    uint64_t *bar_w = bar_x + 1;  // [2]
    float *xs = sm + lay.xs;
    float *ws = sm + lay.ws;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunk0 = blockIdx.x * CPB;
    const int C = min(CPB, B - chunk0);
    const int CL = g.CL;

    pdl_launch_dependents();
    if (tid == 0) {
        mbar_init(bar_x, 1);
        mbar_init(&bar_w[0], 1);
        mbar_init(&bar_w[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_expect_tx(&bar_w[s], SLAB_FLOATS * 4u);
            bulk_g2s(ws + s * SLAB_FLOATS, wslabs + (size_t)s * SLAB_FLOATS, SLAB_FLOATS * 4u,
                     &bar_w[s]);
        }
    }
    pdl_wait();  // cat is produced by K1
    if (tid == 0) {
        const uint32_t xbytes = (uint32_t)C * g.cat_stride * 4u;
        mbar_expect_tx(bar_x, xbytes);
        bulk_g2s(xs, cat + (size_t)chunk0 * g.cat_stride, xbytes, bar_x);
    }
    // ---- merge conv: warp w -> output channels [MRW*w, MRW*w+MRW); lane = tb*CL + chunk -> NR2 steps
    const int m0 = warp * MRW;
    int tb = lane / CL, chunk = lane - tb * CL;
    const bool lane_ok = tb < g.NB2 && chunk < C;
    if (tb >= g.NB2) tb = g.NB2 - 1;
    if (chunk >= C) chunk = C - 1;
    const int t0 = tb * NR2;
    // window rows t0 .. t0+NR2+3 (clamped into the chunk)
    const float *xrow[NR2 + KW_MRG - 1];
#pragma unroll
    for (int r = 0; r < NR2 + KW_MRG - 1; ++r) {
        int t = t0 + r;
        if (t > g.T3 - 1) t = g.T3 - 1;
        xrow[r] = xs + chunk * g.cat_stride + t * XP;
    }
    float2 acc[NR2][MRW / 2];
#pragma unroll
    for (int n = 0; n < NR2; ++n)
#pragma unroll
        for (int p = 0; p < MRW / 2; ++p) acc[n][p] = make_float2(0.f, 0.f);

    mbar_wait(bar_x, 0);
#pragma unroll 1
    for (int s = 0; s < N_SLABS; ++s) {
        const float *wbuf = ws + (s & 1) * SLAB_FLOATS;
        mbar_wait(&bar_w[s & 1], (s >> 1) & 1);
#pragma unroll 1
        for (int c4 = 0; c4 < SLAB_C; c4 += 4) {
            float4 x[NR2 + KW_MRG - 1];
#pragma unroll
            for (int r = 0; r < NR2 + KW_MRG - 1; ++r)
                x[r] = *reinterpret_cast<const float4 *>(xrow[r] + s * SLAB_C + c4);
#pragma unroll
            for (int j = 0; j < KW_MRG; ++j) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const float4 *wp = reinterpret_cast<const float4 *>(
                        wbuf + (j * SLAB_C + c4 + kk) * SIZE + m0);
                    float2 w[MRW / 2];
#pragma unroll
                    for (int q = 0; q < MRW / 4; ++q) {
                        const float4 wv = wp[q];
                        w[2 * q] = make_float2(wv.x, wv.y);
                        w[2 * q + 1] = make_float2(wv.z, wv.w);
                    }
#pragma unroll
                    for (int n = 0; n < NR2; ++n) {
                        const float4 xq = x[n + j];
                        const float xv = kk == 0 ? xq.x : kk == 1 ? xq.y : kk == 2 ? xq.z : xq.w;
#pragma unroll
                        for (int p2 = 0; p2 < MRW / 2; ++p2) acc[n][p2] = ffma2(w[p2], xv, acc[n][p2]);
                    }
                }
            }
        }
        __syncthreads();  // every warp is done with this slab buffer
        if (tid == 0) {
            if (s + 2 < N_SLABS) {
                mbar_expect_tx(&bar_w[s & 1], SLAB_FLOATS * 4u);
                bulk_g2s(ws + (s & 1) * SLAB_FLOATS, wslabs + (size_t)(s + 2) * SLAB_FLOATS,
                         SLAB_FLOATS * 4u, &bar_w[s & 1]);
            } else {
                // W_ih1^T half (32 k-rows x 256) into the buffer that just became free
                const int h = s - (N_SLABS - 2);
                mbar_expect_tx(&bar_w[s & 1], 32 * 256 * 4u);
                bulk_g2s(ws + (s & 1) * SLAB_FLOATS, wih1T + (size_t)h * 32 * 256, 32 * 256 * 4u,
                         &bar_w[s & 1]);
            }
        }
    }
    // ---- merge epilogue: bias + swish -> ms[(chunk*TM + t)][MP] (aliases xs; all warps synced) -----
    float *ms = xs;
    if (lane_ok) {
        float b[MRW];
#pragma unroll
        for (int i = 0; i < MRW; ++i) b[i] = bmerge[m0 + i];
#pragma unroll
        for (int n = 0; n < NR2; ++n) {
            const int t = t0 + n;
            if (t < g.TM) {
                float4 *dst = reinterpret_cast<float4 *>(ms + (chunk * g.TM + t) * MP + m0);
#pragma unroll
                for (int q = 0; q < MRW / 4; ++q)
                    dst[q] = make_float4(swishf_fast(acc[n][2 * q].x + b[4 * q]),
                                         swishf_fast(acc[n][2 * q].y + b[4 * q + 1]),
                                         swishf_fast(acc[n][2 * q + 1].x + b[4 * q + 2]),
                                         swishf_fast(acc[n][2 * q + 1].y + b[4 * q + 3]));
            }
        }
    }
    __syncthreads();
    // ---- LSTM1 input projection: xp[pos][r] = b1[r] + sum_k W_ih1[r][k] * m[pos][k] ---------------
    // warp w, pass p -> gate rows [128 p + XR w, +XR); lane -> positions lane + 32 n
    const int npos = C * g.TM;
    const float *mrow[NR2];
#pragma unroll
    for (int n = 0; n < NR2; ++n) {
        int pos = lane + 32 * n;
        if (pos > npos - 1) pos = npos - 1;
        mrow[n] = ms + pos * MP;
    }
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int r0 = pass * 128 + warp * XR;
        float2 pa[NR2][XR / 2];
#pragma unroll
        for (int n = 0; n < NR2; ++n)
#pragma unroll
            for (int p = 0; p < XR / 2; ++p) pa[n][p] = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            // half h of W_ih1^T lives in buffer ((N_SLABS-2+h) & 1); its load was the
            // (N_SLABS-2+h)/2 + 1 -th use of that barrier -> parity below
            const int sidx_buf = (N_SLABS - 2 + h) & 1;
            const int use = ((N_SLABS - 2 + h) >> 1) + 1;
            if (pass == 0) mbar_wait(&bar_w[sidx_buf], use & 1);
            const float *wbuf = ws + sidx_buf * SLAB_FLOATS;
#pragma unroll 1
            for (int c4 = 0; c4 < 32; c4 += 4) {
                float4 x[NR2];
#pragma unroll
                for (int n = 0; n < NR2; ++n)
                    x[n] = *reinterpret_cast<const float4 *>(mrow[n] + h * 32 + c4);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const float4 *wp = reinterpret_cast<const float4 *>(wbuf + (c4 + kk) * 256 + r0);
                    float2 w[XR / 2];
#pragma unroll
                    for (int q = 0; q < XR / 4; ++q) {
                        const float4 wv = wp[q];
                        w[2 * q] = make_float2(wv.x, wv.y);
                        w[2 * q + 1] = make_float2(wv.z, wv.w);
                    }
#pragma unroll
                    for (int n = 0; n < NR2; ++n) {
                        const float xv = kk == 0 ? x[n].x : kk == 1 ? x[n].y : kk == 2 ? x[n].z : x[n].w;
#pragma unroll
                        for (int p = 0; p < XR / 2; ++p) pa[n][p] = ffma2(w[p], xv, pa[n][p]);
                    }
                }
            }
        }
        float bb[XR];
#pragma unroll
        for (int i = 0; i < XR; ++i) bb[i] = b1[r0 + i];
#pragma unroll
        for (int n = 0; n < NR2; ++n) {
            const int pos = lane + 32 * n;
            if (pos < npos) {
                float4 *dst = reinterpret_cast<float4 *>(xp + ((size_t)chunk0 * g.TM + pos) * 256 + r0);
#pragma unroll
                for (int q = 0; q < XR / 4; ++q)
                    dst[q] = make_float4(pa[n][2 * q].x + bb[4 * q], pa[n][2 * q].y + bb[4 * q + 1],
                                         pa[n][2 * q + 1].x + bb[4 * q + 2],
                                         pa[n][2 * q + 1].y + bb[4 * q + 3]);
            }
        }
    }
}

// =================================================================================================
// K2-TC: merge_conv1 + LSTM1 input projection on the 5th-generation tensor cores (tcgen05)
// =================================================================================================
// Both are GEMMs, so they run as tcgen05.mma kind::tf32 with the accumulators in TMEM.  A single TF32
// pass (10-bit mantissa) cannot meet the 1e-4 fp32 parity gate, so every product is issued as the
// 3xTF32 split  a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo  (fp32 accumulate in TMEM, ~2^-21 relative).
//   * positions are the MMA's M dimension (128 rows = (chunk, t) pairs), output channels its N;
//   * the convolution's taps need no im2col: tap j is the same 128B-swizzled K-major activation tile
//     with the descriptor start address advanced by j rows (j*128 B) - the swizzle is a function of
//     the absolute shared-memory address, so shifted descriptors stay consistent (verified on B200,
//     scripts/microbench/umma_test.cu);
//   * K1 writes the activations already split into hi/lo and pre-swizzled, one image per CTA, so a
//     plain bulk TMA copy lands them MMA-ready; weights are pre-split/pre-swizzled on the host;
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM allocator,
//     warps 2..5 = epilogue (TMEM -> registers -> bias/swish/split -> shared memory or HBM).
namespace tc {
constexpr int WST_BYTES = 2 * 64 * 128;      // merge weights of one (channel block, tap): hi + lo
constexpr int N_WST = 4;                     // weight ring depth (16 KB stages)
constexpr int XST_BYTES = 2 * 128 * 128;     // W_ih tile (128 gate rows x 32 k): hi + lo
constexpr int N_XST = 2;
constexpr int MT_BYTES = 128 * 128;          // one [128 rows][32 k] tile
constexpr int THREADS_TC = 320;              // producer + MMA + 8 epilogue/converter warps
constexpr int EPI_THREADS = 256;
constexpr int TMEM_COLS = 512;

struct Geo {
    int R;        // activation rows of the CTA = chunks * T3
    int rpad;     // rows per K-block tile in the image (multiple of 8, >= R + 8)
    int n_mt;     // 1 or 2 M-tiles of 128 rows
    int base1;    // first row of the second M-tile (= R - 128, overlaps the first)
};
__host__ __device__ inline int rpad_for(int cl, int T3) { return ((cl * T3 + 7) & ~7) + 8; }

struct Smem {
    int bars, a_ring, w_ring, total;
    int a_stage_bytes;
};
__host__ __device__ inline Smem smem_layout(int rpad) {
    Smem l;
    l.bars = 0;
    l.a_ring = 1024;
    l.a_stage_bytes = 2 * rpad * 128;  // hi + lo tile of one channel block
    const int merge_bytes = 2 * l.a_stage_bytes + N_WST * WST_BYTES;
    l.w_ring = l.a_ring + 2 * l.a_stage_bytes;
    const int xproj_bytes = 2 * 2 * 2 * MT_BYTES + N_XST * XST_BYTES;  // m tiles [mt][kb][hi|lo] + W_ih ring
    l.total = 1024 + (merge_bytes > xproj_bytes ? merge_bytes : xproj_bytes);
    return l;
}

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    // K-major, SWIZZLE_128B, 8-row x 128 B atoms stacked every 1024 B, descriptor version 1 (sm_100)
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    // D = f32 (bits 4-5 = 1), A = B = tf32 (format 2 at bits 7-9 / 10-12), K-major both, N>>3, M>>4
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 3xTF32: hi*hi + lo*hi + hi*lo over one 32-wide K block (4 MMAs of K = 8 each).  The tensor core adds
// into the TMEM accumulator with truncation, so a long chain picks up a systematic bias
// (scripts/microbench/umma_acc_test.cu: 5.7e-6 relative after 240 accumulate steps); the two small
// correction products therefore go to their own accumulator and the main chain is kept short.
__device__ __forceinline__ void mma3_kblock(uint32_t d_main, uint32_t d_small, uint32_t a_hi, uint32_t a_lo,
                                            uint32_t b_hi, uint32_t b_lo, uint32_t idesc, bool first_main,
                                            bool first_small) {
#pragma unroll
    for (int k8 = 0; k8 < 4; ++k8) {
        const uint32_t o = k8 * 32;
        mma_tf32(d_main, desc_sw128(a_hi + o), desc_sw128(b_hi + o), idesc, (first_main && k8 == 0) ? 0u : 1u);
        mma_tf32(d_small, desc_sw128(a_lo + o), desc_sw128(b_hi + o), idesc, (first_small && k8 == 0) ? 0u : 1u);
        mma_tf32(d_small, desc_sw128(a_hi + o), desc_sw128(b_lo + o), idesc, 1u);
    }
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
        "%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
          "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
          "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
          "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

struct Bars {
    uint64_t a_full[2], a_conv[2], a_empty[2], w_full[N_WST], w_empty[N_WST], x_full[N_XST], x_empty[N_XST];
    uint64_t d_full, m_ready, d2_full[2];
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(THREADS_TC, 1)
k2tc_kernel(const float *__restrict__ cat_img, const float *__restrict__ wm_tc,
            const float *__restrict__ bmerge, const float *__restrict__ wih_tc,
            const float *__restrict__ b1, float *__restrict__ xp, int B, int CPB, int T3, int TM,
            int rpad, long long *__restrict__ stamps) {
    extern __shared__ __align__(1024) uint8_t smt[];
    // optional phase timestamps of CTA 0 (profiling aid; null in production)
#define TC_STAMP(i) do { if (stamps && blockIdx.x == 0) stamps[i] = clock64(); } while (0)
    const Smem lay = smem_layout(rpad);
    Bars *bars = reinterpret_cast<Bars *>(smt);
    uint8_t *a_ring = smt + lay.a_ring;
    uint8_t *w_ring = smt + lay.w_ring;
    uint8_t *m_tiles = smt + lay.a_ring;                       // xproj phase: aliases the A ring
    uint8_t *x_ring = m_tiles + 2 * 2 * 2 * MT_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunk0 = blockIdx.x * CPB;
    const int C = min(CPB, B - chunk0);
    const int R = C * T3;
    const int n_mt = R > 128 ? 2 : 1;
    const int mt_base[2] = {0, R > 128 ? R - 128 : 0};
    const int tile_bytes = rpad * 128;

    pdl_launch_dependents();
    if (tid == 0) TC_STAMP(0);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->a_full[i], 1);
            mbar_init(&bars->a_conv[i], EPI_THREADS);
            mbar_init(&bars->a_empty[i], 1);
            mbar_init(&bars->d2_full[i], 1);
        }
        for (int i = 0; i < N_WST; ++i) {
            mbar_init(&bars->w_full[i], 1);
            mbar_init(&bars->w_empty[i], 1);
        }
        for (int i = 0; i < N_XST; ++i) {
            mbar_init(&bars->x_full[i], 1);
            mbar_init(&bars->x_empty[i], 1);
        }
        mbar_init(&bars->d_full, 1);
        mbar_init(&bars->m_ready, EPI_THREADS);
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_addr(&bars->tmem_base)),
                     "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = bars->tmem_base;

    if (warp == 0) {
        // ===================================== TMA producer ========================================
        if (lane == 0) {
            const float *img = cat_img + (size_t)blockIdx.x * (4 * rpad * 32);
            const uint32_t a_bytes = (uint32_t)(((R + 7) & ~7) * 128);  // rows that exist in HBM
            int wcount = 0;
            // merge weights do not depend on K1: the first ring fill may run ahead of pdl_wait()
            for (int cb = 0; cb < 4; ++cb) {
                if (cb == 0) {
                    // prefetch the first weight stages, then wait for K1's activations
                    for (int tap = 0; tap < N_WST; ++tap) {
                        mbar_expect_tx(&bars->w_full[tap], WST_BYTES);
                        bulk_g2s(w_ring + tap * WST_BYTES, wm_tc + (size_t)(cb * 5 + tap) * (WST_BYTES / 4),
                                 WST_BYTES, &bars->w_full[tap]);
                    }
                    pdl_wait();
                    TC_STAMP(1);
                }
                const int as = cb & 1;
                if (cb >= 2) mbar_wait(&bars->a_empty[as], ((cb >> 1) - 1) & 1);
                mbar_expect_tx(&bars->a_full[as], a_bytes);
                bulk_g2s(a_ring + as * lay.a_stage_bytes, img + (size_t)cb * rpad * 32, a_bytes,
                         &bars->a_full[as]);
                for (int tap = 0; tap < 5; ++tap, ++wcount) {
                    if (wcount < N_WST) continue;  // already issued above
                    const int ws = wcount % N_WST;
                    mbar_wait(&bars->w_empty[ws], ((wcount / N_WST) - 1) & 1);
                    mbar_expect_tx(&bars->w_full[ws], WST_BYTES);
                    bulk_g2s(w_ring + ws * WST_BYTES, wm_tc + (size_t)wcount * (WST_BYTES / 4), WST_BYTES,
                             &bars->w_full[ws]);
                }
            }
            // x-projection weights: the ring aliases the merge rings, so wait until the merge MMAs
            // have drained them (d_full) before the first fill
            mbar_wait(&bars->d_full, 0);
            for (int i = 0; i < 4; ++i) {  // i = nh * 2 + kb
                const int xs = i % N_XST;
                if (i >= N_XST) mbar_wait(&bars->x_empty[xs], ((i / N_XST) - 1) & 1);
                mbar_expect_tx(&bars->x_full[xs], XST_BYTES);
                bulk_g2s(x_ring + xs * XST_BYTES, wih_tc + (size_t)i * (XST_BYTES / 4), XST_BYTES,
                         &bars->x_full[xs]);
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer ==========================================
        if (lane == 0) {
            constexpr uint32_t idesc_m = idesc_tf32(128, 64);
            constexpr uint32_t idesc_m128 = idesc_tf32(128, 128);
            constexpr uint32_t idesc_x = idesc_tf32(128, 128);
            int wcount = 0;
            for (int cb = 0; cb < 4; ++cb) {
                const int as = cb & 1;
                mbar_wait(&bars->a_conv[as], (cb >> 1) & 1);  // fp32 tile landed and its lo tile is written
                if (cb == 0) TC_STAMP(2);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = smem_addr(a_ring + as * lay.a_stage_bytes);
                const uint32_t a_lo = a_hi + tile_bytes;
                for (int tap = 0; tap < 5; ++tap, ++wcount) {
                    const int ws = wcount % N_WST;
                    mbar_wait(&bars->w_full[ws], (wcount / N_WST) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t b_hi = smem_addr(w_ring + ws * WST_BYTES);
                    const uint32_t b_lo = b_hi + 64 * 128;
                    for (int mt = 0; mt < n_mt; ++mt) {
                        // TMEM columns of M tile mt: chain 0 (channel blocks 0,1) = [main0 | corrections],
                        // chain 1 (blocks 2,3) = [main1 | hi*lo of chain 1], 64 columns each.  The weight
                        // stage holds the 64-row hi tile and the 64-row lo tile back to back, so ONE
                        // N=128 MMA with the activation hi tile yields a_hi*b_hi and a_hi*b_lo side by
                        // side; the remaining a_lo*b_hi product (N=64) always adds into chain 0's
                        // correction columns.  The activation tile is read from shared memory twice per
                        // K step instead of three times, and the main chains stay short (40 steps).
                        const uint32_t row_off = (uint32_t)(mt_base[mt] + tap) * 128u;
                        const uint32_t d0 = tmem + mt * 256;
                        const uint32_t dc = d0 + (cb >= 2 ? 128 : 0);
                        const uint32_t b_t = smem_addr(w_ring + ws * WST_BYTES);  // [hi (64 rows); lo (64 rows)]
                        const bool first_chain = (cb & 1) == 0 && tap == 0;
#pragma unroll
                        for (int k8 = 0; k8 < 4; ++k8) {
                            const uint32_t o = k8 * 32;
                            mma_tf32(dc, desc_sw128(a_hi + row_off + o), desc_sw128(b_t + o), idesc_m128,
                                     (first_chain && k8 == 0) ? 0u : 1u);
                            mma_tf32(d0 + 64, desc_sw128(a_lo + row_off + o), desc_sw128(b_t + o), idesc_m, 1u);
                        }
                    }
                    umma_commit(&bars->w_empty[ws]);
                }
                umma_commit(&bars->a_empty[as]);
            }
            umma_commit(&bars->d_full);
            TC_STAMP(3);
            // ---- x-projection: D2[nh][mt] (128 x 128) = m[mt] (128 x 64) . W_ih[nh]^T ----------------
            mbar_wait(&bars->m_ready, 0);  // epilogue warps wrote the hi/lo m tiles
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            TC_STAMP(5);
            for (int i = 0; i < 4; ++i) {
                const int nh = i >> 1, kb = i & 1, xs = i % N_XST;
                mbar_wait(&bars->x_full[xs], (i / N_XST) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b_hi = smem_addr(x_ring + xs * XST_BYTES);
                const uint32_t b_lo = b_hi + 128 * 128;
                for (int mt = 0; mt < n_mt; ++mt) {
                    const uint32_t a_hi = smem_addr(m_tiles + ((mt * 2 + kb) * 2 + 0) * MT_BYTES);
                    const uint32_t a_lo = a_hi + MT_BYTES;
                    const uint32_t d2 = tmem + (nh * 2 + mt) * 128;  // 24 accumulate steps: one chain
                    mma3_kblock(d2, d2, a_hi, a_lo, b_hi, b_lo, idesc_x, kb == 0, false);
                }
                umma_commit(&bars->x_empty[xs]);
                if (kb == 1) umma_commit(&bars->d2_full[nh]);
            }
            TC_STAMP(6);
        }
    } else {
        // ===================================== epilogue warps ======================================
        const int q = warp & 3;            // TMEM lane quarter this warp may access
        const int colh = (warp - 2) >> 2;  // two warps share a quarter: each takes half of the columns
        const int row_in_tile = q * 32 + lane;
        const int et = tid - 64;           // 0..255
        // ---- converter: lo = a - trunc_tf32(a) for every activation tile the producer lands -------
        {
            const int n4 = ((R + 7) & ~7) * 8;  // float4s per tile
            for (int cb = 0; cb < 4; ++cb) {
                const int as = cb & 1;
                mbar_wait(&bars->a_full[as], (cb >> 1) & 1);
                const float4 *src = reinterpret_cast<const float4 *>(a_ring + as * lay.a_stage_bytes);
                float4 *dst = reinterpret_cast<float4 *>(a_ring + as * lay.a_stage_bytes + tile_bytes);
                for (int i = et; i < n4; i += EPI_THREADS) {
                    const float4 a = src[i];
                    dst[i] = make_float4(tf32_trunc_lo(a.x), tf32_trunc_lo(a.y), tf32_trunc_lo(a.z),
                                         tf32_trunc_lo(a.w));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&bars->a_conv[as]);
            }
        }
        // ---- epilogue 1: merge accumulators -> bias + swish -> hi/lo m tiles (MMA-ready) ----------
        mbar_wait(&bars->d_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 64) TC_STAMP(4);
        for (int mt = 0; mt < n_mt; ++mt) {
            {
                const int half = colh;  // 32 channels = one K block of the projection
                float v[32];
                {
                    float v1[32];
                    const uint32_t t0 = tmem + ((uint32_t)(q * 32) << 16) + mt * 256 + half * 32;
                    tmem_ld32(t0, v);         // main chain 0
                    tmem_ld32(t0 + 128, v1);  // main chain 1
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += v1[i];
                    tmem_ld32(t0 + 64, v1);   // corrections (a_lo*b_hi of both chains + a_hi*b_lo of chain 0)
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += v1[i];
                    tmem_ld32(t0 + 192, v1);  // a_hi*b_lo of chain 1
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += v1[i];
                }
                float *t_hi = reinterpret_cast<float *>(m_tiles + ((mt * 2 + half) * 2 + 0) * MT_BYTES);
                float *t_lo = reinterpret_cast<float *>(m_tiles + ((mt * 2 + half) * 2 + 1) * MT_BYTES);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    float4 o;
                    o.x = swishf_fast(v[4 * c4 + 0] + bmerge[half * 32 + 4 * c4 + 0]);
                    o.y = swishf_fast(v[4 * c4 + 1] + bmerge[half * 32 + 4 * c4 + 1]);
                    o.z = swishf_fast(v[4 * c4 + 2] + bmerge[half * 32 + 4 * c4 + 2]);
                    o.w = swishf_fast(v[4 * c4 + 3] + bmerge[half * 32 + 4 * c4 + 3]);
                    float4 hi, lo;
                    split_tf32(o, hi, lo);
                    const int off = row_in_tile * 32 + ((c4 ^ (row_in_tile & 7)) << 2);
                    *reinterpret_cast<float4 *>(t_hi + off) = hi;
                    *reinterpret_cast<float4 *>(t_lo + off) = lo;
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // st.shared -> tensor-core reads
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(&bars->m_ready);
        // ---- epilogue 2: projection accumulators + b1 -> xp[pos][256] in HBM ----------------------
        for (int nh = 0; nh < 2; ++nh) {
            mbar_wait(&bars->d2_full[nh], 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tid == 64) TC_STAMP(7 + nh);
            for (int mt = 0; mt < n_mt; ++mt) {
                const int row = mt_base[mt] + row_in_tile;
                const int chunk = row / T3, t = row - chunk * T3;
                // rows of the second tile that the first one already covered are skipped
                const bool ok = row < R && t < TM && (mt == 0 || row >= 128);
                float *dst = xp + ((size_t)(chunk0 + chunk) * TM + t) * 256 + nh * 128;
#pragma unroll 1
                for (int cq = 2 * colh; cq < 2 * colh + 2; ++cq) {
                    float v[32];
                    tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (nh * 2 + mt) * 128 + cq * 32, v);
                    if (ok) {
#pragma unroll
                        for (int c4 = 0; c4 < 8; ++c4) {
                            const float4 bb = *reinterpret_cast<const float4 *>(b1 + nh * 128 + cq * 32 + 4 * c4);
                            *reinterpret_cast<float4 *>(dst + cq * 32 + 4 * c4) =
                                make_float4(v[4 * c4] + bb.x, v[4 * c4 + 1] + bb.y, v[4 * c4 + 2] + bb.z,
                                            v[4 * c4 + 3] + bb.w);
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) TC_STAMP(9);
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
#undef TC_STAMP
}

// =================================================================================================
// K1-TC: front kernel with sig_conv3 / seq_conv2 on tcgen05 (3xTF32), feeding K2-TC's image format
// =================================================================================================
// The two stride-3 convolutions are GEMMs with M = output steps (same two 128-row tiles as K2-TC),
// N = 64 and K = taps x 16 channels.  A stride-3 window cannot be expressed as a descriptor shift, so
// the im2col tile of one 32-wide K block (two taps x 16 channels = two adjacent 64-byte activation
// rows, contiguous because the activation pitch is exactly 16 floats here) is copied into a
// 128B-swizzled tile by all threads, together with its TF32 remainder tile; one elected thread then
// issues that K block's MMAs, which run asynchronously while the CTA builds the next tile (two tile
// stages, freed by tcgen05.commit -> mbarrier).  Weight tiles (hi/lo, pre-swizzled on the host) stream
// through a 2-stage TMA ring driven by the same thread.  The sig_conv3 MMAs overlap the seq_conv1
// gather phase; both accumulators are drained at the end into the fp32 image K2-TC consumes.
constexpr int K1TC_NKB_SIG = (KW_SIG3 * 16 + 31) / 32;  // 5
constexpr int K1TC_NKB_SEQ = (KW_SEQ2 * 16 + 31) / 32;  // 7
constexpr int K1TC_MAX_CPB = 7;

struct K1Bars {
    uint64_t wfront, a_full[2], a_empty[2], w_full[3], w_empty[3], d_done;
    uint32_t tmem_base;
};
struct K1TcSmem {
    int a_stages, w_ring, weights, sig, s1, act, sidx, seq, map, len, total;  // byte offsets
    int stage_bytes, act_stride;                                             // act_stride in floats
};
__host__ __device__ inline K1TcSmem k1tc_smem(const Geometry &g, int cl, int kmer_len, int seq_width,
                                              int map_width, int rpad, int nw = 2) {
    K1TcSmem l;
    const FrontOffsets fo = front_offsets(kmer_len, true);
    int cur = 1024;  // barriers
    l.a_stages = cur;
    l.stage_bytes = 2 * rpad * 128;
    cur += 2 * l.stage_bytes;
    l.w_ring = cur;
    cur += nw * WST_BYTES;  // weight ring: 3 stages when they fit (hides the TMA latency), else 2
    auto take = [&](int bytes) {
        int at = cur;
        cur += (bytes + 15) & ~15;
        return at;
    };
    l.weights = take(fo.total * 4);
    l.sig = take(cl * g.T * 4);
    l.s1 = take(cl * g.T1 * 16);
    l.act_stride = (g.Q1 > g.T2 ? g.Q1 : g.T2) * 16;
    l.act = take(cl * l.act_stride * 4);
    l.sidx = take(cl * g.T * 2);
    l.seq = take(cl * seq_width);
    l.map = take(cl * map_width * 2);
    l.len = take(cl * 4);
    l.total = cur;
    return l;
}

// one stride-3 convolution.  Warp 0: streams the weight tiles (2-stage TMA ring) and issues the MMAs;
// warps 1..7: build the im2col tiles (2 stages).  Producer/consumer hand-off through mbarriers only:
// a_full (builders -> MMA warp), a_empty / w_empty (tcgen05.commit -> builders / TMA).
constexpr int BUILDERS = THREADS - 32;
template <int KW>
__device__ __forceinline__ void k1tc_conv(const float *__restrict__ act, int act_stride, int T3, int R, int rpad,
                                          uint8_t *a_stages, int stage_bytes, uint8_t *w_ring,
                                          const float *__restrict__ w_tc, K1Bars *bars, uint32_t tmem_d,
                                          int n_mt, int base1, int &kcount, int total_kb, int nw) {
    constexpr int NKB = (KW * 16 + 31) / 32;
    constexpr uint32_t idesc = idesc_tf32(128, 64);
    constexpr uint32_t idesc128 = idesc_tf32(128, 128);
    const int k0 = kcount;
    kcount += NKB;
    if (threadIdx.x >= 32) {
        // ---------------------------------- tile builders -------------------------------------------
        // item i = (row r, 16-byte chunk q): source and destination offsets are fixed per conv except
        // for +32 floats of source per K block, so they are computed once (no divisions in the K loop)
        const int bt = threadIdx.x - 32;
        constexpr int ITEMS = (256 * 8 + BUILDERS - 1) / BUILDERS;  // R <= 256 rows
        int src[ITEMS], dst[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const int i = bt + k * BUILDERS;
            const int r = i >> 3, q = i & 7;
            const int chunk = r / T3, tp = r - chunk * T3;
            // q >= 4 is the second tap of the K block; bit 30 marks it so the tail K block can zero it
            src[k] = i < R * 8 ? (chunk * act_stride + (3 * tp + (q >> 2)) * 16 + (q & 3) * 4) | ((q >> 2) << 30)
                               : -1;
            dst[k] = r * 32 + ((q ^ (r & 7)) << 2);
        }
        for (int kbk = 0; kbk < NKB; ++kbk) {
            const int kc = k0 + kbk, st = kc & 1;
            if (kc >= 2) mbar_wait(&bars->a_empty[st], ((kc >> 1) - 1) & 1);
            float *t_hi = reinterpret_cast<float *>(a_stages + st * stage_bytes);
            float *t_lo = t_hi + rpad * 32;
            const bool tap_hi_ok = 2 * kbk + 1 < KW;
            const float *a_k = act + kbk * 32;
            float4 v[ITEMS];
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (src[k] >= 0 && (tap_hi_ok || !(src[k] >> 30)))
                    v[k] = *reinterpret_cast<const float4 *>(a_k + (src[k] & 0x3FFFFFFF));
            }
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                if (src[k] < 0) continue;
                *reinterpret_cast<float4 *>(t_hi + dst[k]) = v[k];
                *reinterpret_cast<float4 *>(t_lo + dst[k]) = make_float4(
                    tf32_trunc_lo(v[k].x), tf32_trunc_lo(v[k].y), tf32_trunc_lo(v[k].z), tf32_trunc_lo(v[k].w));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&bars->a_full[st]);
        }
    } else if (threadIdx.x == 0) {
        // ---------------------------------- TMA + MMA issuer ----------------------------------------
        for (int kbk = 0; kbk < NKB; ++kbk) {
            const int kc = k0 + kbk, st = kc & 1, ws = kc % nw;
            mbar_wait(&bars->w_full[ws], (kc / nw) & 1);
            mbar_wait(&bars->a_full[st], (kc >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = smem_addr(a_stages + st * stage_bytes), a_lo = a_hi + rpad * 128;
            const uint32_t b_hi = smem_addr(w_ring + ws * WST_BYTES);  // {hi (64 rows); lo (64 rows)}
            for (int mt = 0; mt < n_mt; ++mt) {
                // TMEM columns [main | corrections] of this M tile: one N=128 MMA with the weight stage's
                // {hi; lo} tile pair gives a_hi*b_hi and a_hi*b_lo side by side, the N=64 MMA adds
                // a_lo*b_hi into the correction columns (2 reads of the activation tile per K step, not 3)
                const uint32_t ro = (uint32_t)(mt ? base1 : 0) * 128u;
                const uint32_t d = tmem_d + mt * 128;
#pragma unroll
                for (int k8 = 0; k8 < 4; ++k8) {
                    const uint32_t o = k8 * 32;
                    mma_tf32(d, desc_sw128(a_hi + ro + o), desc_sw128(b_hi + o), idesc128,
                             (kbk == 0 && k8 == 0) ? 0u : 1u);
                    mma_tf32(d + 64, desc_sw128(a_lo + ro + o), desc_sw128(b_hi + o), idesc, 1u);
                }
            }
            umma_commit(&bars->a_empty[st]);
            umma_commit(&bars->w_empty[ws]);
            // keep the weight ring nw-1 K blocks ahead: K block kc+nw-1 goes into the stage that K block
            // kc-1 used (its MMAs were issued one iteration ago)
            const int nxt = kc + nw - 1;
            if (nxt < total_kb) {
                const int ns = nxt % nw;
                if (nxt >= nw) mbar_wait(&bars->w_empty[ns], ((nxt / nw) - 1) & 1);
                mbar_expect_tx(&bars->w_full[ns], WST_BYTES);
                bulk_g2s(w_ring + ns * WST_BYTES, w_tc + (size_t)nxt * (WST_BYTES / 4), WST_BYTES,
                         &bars->w_full[ns]);
            }
        }
    }
}

__global__ void __launch_bounds__(THREADS, 1)
k1tc_front_kernel(const float *__restrict__ sigs, const int8_t *__restrict__ seqs, int seq_width,
                  const int16_t *__restrict__ maps, int map_width, const int16_t *__restrict__ lens,
                  const float *__restrict__ wfront, const float *__restrict__ w_tc,
                  float *__restrict__ cat_img, int B, int CPB, int T, int kmer_len, int rpad, int CLs,
                  int nw, const float *__restrict__ q1_in, long long *__restrict__ stamps) {
    extern __shared__ __align__(1024) uint8_t smk[];
#define K1_STAMP(i) do { if (stamps && blockIdx.x == 1 && threadIdx.x == 0) stamps[i] = clock64(); } while (0)
    const Geometry g = make_geometry(T);
    const K1TcSmem lay = k1tc_smem(g, CLs, kmer_len, seq_width, map_width, rpad, nw);
    const FrontOffsets fo = front_offsets(kmer_len, true);
    K1Bars *bars = reinterpret_cast<K1Bars *>(smk);
    uint8_t *a_stages = smk + lay.a_stages;
    uint8_t *w_ring = smk + lay.w_ring;
    float *wsm = reinterpret_cast<float *>(smk + lay.weights);
    float *sig_s = reinterpret_cast<float *>(smk + lay.sig);
    float *s1_s = reinterpret_cast<float *>(smk + lay.s1);
    float *act_s = reinterpret_cast<float *>(smk + lay.act);
    int16_t *sidx_s = reinterpret_cast<int16_t *>(smk + lay.sidx);
    int8_t *seq_s = reinterpret_cast<int8_t *>(smk + lay.seq);
    int16_t *map_s = reinterpret_cast<int16_t *>(smk + lay.map);
    int *len_s = reinterpret_cast<int *>(smk + lay.len);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunk0 = blockIdx.x * CPB;
    const int C = min(CPB, B - chunk0);
    const int R = C * g.T3;
    const int n_mt = R > 128 ? 2 : 1;
    const int base1 = R > 128 ? R - 128 : 0;
    constexpr int TOTAL_KB = K1TC_NKB_SIG + K1TC_NKB_SEQ;

    pdl_launch_dependents();
    if (tid == 0) {
        mbar_init(&bars->wfront, 1);
        mbar_init(&bars->d_done, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->a_full[i], BUILDERS);
            mbar_init(&bars->a_empty[i], 1);
        }
        for (int i = 0; i < 3; ++i) {
            mbar_init(&bars->w_full[i], 1);
            mbar_init(&bars->w_empty[i], 1);
        }
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_addr(&bars->tmem_base)),
                     "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = bars->tmem_base;
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)fo.total * 4u;
        mbar_expect_tx(&bars->wfront, bytes);
        bulk_g2s(wsm, wfront, bytes, &bars->wfront);
        for (int i = 0; i < nw - 1; ++i) {  // weights of the first nw-1 K blocks
            mbar_expect_tx(&bars->w_full[i], WST_BYTES);
            bulk_g2s(w_ring + i * WST_BYTES, w_tc + (size_t)i * (WST_BYTES / 4), WST_BYTES, &bars->w_full[i]);
        }
    }
    K1_STAMP(0);
    pdl_wait();  // inputs (and the image we overwrite) belong to earlier work in the stream
    K1_STAMP(1);
    k1_stage_inputs(sigs, seqs, seq_width, maps, map_width, lens, chunk0, C, T, kmer_len, g, fo, &bars->wfront,
                    wsm, sig_s, s1_s, act_s, sidx_s, seq_s, map_s, len_s, lay.act_stride, 16, q1_in != nullptr);
    // ---- sequence track first: seq_conv1 (two-stage gather; the per-base sums live in the still unused
    // tile stages) -> q1 in act_s -> seq_conv2 on the tensor core, TMEM columns [256, 512) -----------------
    if (q1_in != nullptr) {
        // dense interface: q1 was computed by k0_dense_seq1_kernel; same [chunk][Q1][16] layout as act_s
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)C * g.Q1 * 16 * 4u;
            mbar_expect_tx(&bars->wfront, bytes);
            bulk_g2s(act_s, q1_in + (size_t)chunk0 * g.Q1 * 16, bytes, &bars->wfront);
        }
        mbar_wait(&bars->wfront, 1);
    } else {
        float *gs = reinterpret_cast<float *>(a_stages);
        const int LM = map_width - 1;
        if (kmer_len == 9)
            seq1_gather_two_stage<9>(wsm + fo.w_seq1, wsm + fo.b_seq1, fo.z_seq1 - fo.w_seq1, seq_s, sidx_s,
                                     len_s, gs, act_s, seq_width, LM, C, T, g.Q1, lay.act_stride, kmer_len, 16);
        else
            seq1_gather_two_stage<0>(wsm + fo.w_seq1, wsm + fo.b_seq1, fo.z_seq1 - fo.w_seq1, seq_s, sidx_s,
                                     len_s, gs, act_s, seq_width, LM, C, T, g.Q1, lay.act_stride, kmer_len, 16);
    }
    __syncthreads();
    K1_STAMP(2);
    int kcount = 0;
    k1tc_conv<KW_SEQ2>(act_s, lay.act_stride, g.T3, R, rpad, a_stages, lay.stage_bytes, w_ring, w_tc, bars,
                       tmem + 256, n_mt, base1, kcount, TOTAL_KB, nw);
    K1_STAMP(3);
    __syncthreads();  // every q1 row has been copied into tiles: act_s may be overwritten
    // ---- signal track: sig_conv1, sig_conv2 (run while the tensor core finishes seq_conv2), then
    // sig_conv3 on the tensor core, TMEM columns [0, 256) ------------------------------------------------
    k1_sig12(C, T, g, fo, wsm, sig_s, s1_s, act_s, lay.act_stride, 16);
    K1_STAMP(4);
    k1tc_conv<KW_SIG3>(act_s, lay.act_stride, g.T3, R, rpad, a_stages, lay.stage_bytes, w_ring, w_tc, bars,
                       tmem, n_mt, base1, kcount, TOTAL_KB, nw);
    if (tid == 0) umma_commit(&bars->d_done);
    K1_STAMP(5);
    // ---- epilogue: TMEM -> bias + swish -> fp32 128B-swizzled image [channel block][row][32] -------
    mbar_wait(&bars->d_done, 0);
    K1_STAMP(6);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        const int q = warp & 3, gsel = warp >> 2;  // TMEM lane quarter; 0 = sig_conv3, 1 = seq_conv2
        const float *bias = wsm + (gsel ? fo.b_seq2 : fo.b_sig3);
        float *img = cat_img + (size_t)blockIdx.x * (4 * rpad * 32);
        for (int mt = 0; mt < n_mt; ++mt) {
            const int row = (mt ? base1 : 0) + q * 32 + lane;
            const bool ok = row < R && (mt == 0 || row >= 128);
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                float v[32], v2[32];
                const uint32_t t0 = tmem + ((uint32_t)(q * 32) << 16) + gsel * 256 + mt * 128 + half * 32;
                tmem_ld32(t0, v);
                tmem_ld32(t0 + 64, v2);
                if (ok) {
                    float *dst = img + ((size_t)(gsel * 2 + half) * rpad + row) * 32;
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const float4 bb = *reinterpret_cast<const float4 *>(bias + half * 32 + 4 * c4);
                        float4 o;
                        o.x = swishf_fast(v[4 * c4 + 0] + v2[4 * c4 + 0] + bb.x);
                        o.y = swishf_fast(v[4 * c4 + 1] + v2[4 * c4 + 1] + bb.y);
                        o.z = swishf_fast(v[4 * c4 + 2] + v2[4 * c4 + 2] + bb.z);
                        o.w = swishf_fast(v[4 * c4 + 3] + v2[4 * c4 + 3] + bb.w);
                        *reinterpret_cast<float4 *>(dst + ((c4 ^ (row & 7)) << 2)) = o;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    K1_STAMP(7);
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
#undef K1_STAMP
}
}  // namespace tc

// =================================================================================================
// K3: LSTM1 recurrence (W_hh in registers) + single-step LSTM2 + fc
// =================================================================================================
// Per step and chunk the recurrence is a 256x64 mat-vec.  Thread (rb, kg) keeps the 4x16 block
// W_hh[4*rb .. 4*rb+3][16*kg .. 16*kg+15] in 64 registers for the whole kernel; h lives in shared
// memory as h[k][chunk] so one LDS.128 feeds 4 chunks and the FFMA2 halves are two chunks.  The four
// k-groups of a row block sit in adjacent lanes: their partial sums are combined with a two-round
// transposing butterfly (xor 2, xor 1) that leaves lane kg with the finished row 4*rb + kg.
//
// The CTA's (up to) 8 chunks are processed as two half-batches A (chunks 0..3) and B (4..7) that
// are software-pipelined against each other: between two barriers every thread runs the mat-vec of
// one half AND the gate non-linearities of the other half, so the MUFU/latency chain of the cell
// update hides behind the other half's FFMA2 stream instead of idling the SM:
//     mv(A,0) | bar | mv(B,0)+cell(A,0) | bar | mv(A,1)+cell(B,0) | bar | mv(B,1)+cell(A,1) | ...
constexpr int C3MAX = 8;
constexpr int HC = 4;              // chunks per half-batch
constexpr int HG = 16 * HC + 4;    // floats per k-group of h (16 k x 4 chunks + 4 pad)

__device__ __forceinline__ float2 shfl_xor2(float2 v, int mask) {
    return make_float2(__shfl_xor_sync(0xffffffffu, v.x, mask), __shfl_xor_sync(0xffffffffu, v.y, mask));
}
__device__ __forceinline__ float2 sel2(bool take_a, float2 a, float2 b) {
    return make_float2(take_a ? a.x : b.x, take_a ? a.y : b.y);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }

// gate pre-activations of one half-batch: g[c][r] = xp[c][r] + sum_k W_hh[r][k] h[k][c]
__device__ __forceinline__ void lstm_matvec_half(const float (&w)[4][16], const float *__restrict__ hk,
                                                 const float (&xin)[HC], float *__restrict__ g_half,
                                                 int r, bool hi2, bool hi1) {
    float2 a[4][2];  // [row][chunk pair]
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i][0] = a[i][1] = make_float2(0.f, 0.f);
#pragma unroll
    for (int kl = 0; kl < 16; ++kl) {
        const float4 hv = *reinterpret_cast<const float4 *>(hk + kl * HC);
        const float2 h01 = make_float2(hv.x, hv.y), h23 = make_float2(hv.z, hv.w);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[i][0] = ffma2(h01, w[i][kl], a[i][0]);
            a[i][1] = ffma2(h23, w[i][kl], a[i][1]);
        }
    }
    float2 rA[2][2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const float2 send = sel2(hi2, a[j][p], a[2 + j][p]);
            const float2 keep = sel2(hi2, a[2 + j][p], a[j][p]);
            rA[j][p] = add2(keep, shfl_xor2(send, 2));
        }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const float2 send = sel2(hi1, rA[0][p], rA[1][p]);
        const float2 keep = sel2(hi1, rA[1][p], rA[0][p]);
        const float2 g2 = add2(keep, shfl_xor2(send, 1));
        g_half[(2 * p) * 256 + r] = g2.x + xin[2 * p];
        g_half[(2 * p + 1) * 256 + r] = g2.y + xin[2 * p + 1];
    }
}

// cell update of (unit u, chunk q) of one half-batch; writes h into the half's h[k][chunk] tile
__device__ __forceinline__ float lstm_cell_half(const float *__restrict__ g_half, float *__restrict__ h_half,
                                                int u, int q, int hu, float &cst) {
    const float *g = g_half + q * 256;
    const float ig = sigmoidf_fast(g[u]), fg = sigmoidf_fast(g[64 + u]);
    const float gg = tanhf_fast(g[128 + u]), og = sigmoidf_fast(g[192 + u]);
    cst = fg * cst + ig * gg;
    const float h = og * tanhf_fast(cst);
    h_half[hu + q] = h;
    return h;
}

// named-barrier helpers (producer: arrive, consumer: sync; `count` threads take part in total)
__device__ __forceinline__ void nbar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void nbar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
constexpr int K3_MV_THREADS = 256;   // warps 0..7 : mat-vec (W_hh in registers)
constexpr int K3_CELL_THREADS = 128; // warps 8..11: gate non-linearities / cell update
constexpr int K3_THREADS = K3_MV_THREADS + K3_CELL_THREADS;
enum { BAR_GA = 1, BAR_HA = 2, BAR_GB = 3, BAR_HB = 4 };

// Warp-specialised schedule (A = chunks 0..3, B = chunks 4..7 of the CTA):
//   mat-vec warps : mv(A,0) | mv(B,t) , wait h_A(t) , mv(A,t+1) , wait h_B(t) ...
//   cell warps    :           wait g_A(t) , cell(A,t) -> h_A , wait g_B(t) , cell(B,t) -> h_B ...
// so the MUFU-latency chain of one half's cell update runs under the other half's FFMA2 stream.
__global__ void __launch_bounds__(K3_THREADS, 1)
k3_lstm_kernel(const float *__restrict__ xp, const float4 *__restrict__ whh4,
               const float *__restrict__ wih2T, const float *__restrict__ b2,
               const float *__restrict__ fcw, const float *__restrict__ fcb,
               float *__restrict__ logits, int B, int CPB, int TM, int num_out) {
    extern __shared__ __align__(128) float sm3[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm3);
    float *w2_s = sm3 + 4;                    // W_ih2^T [64][256], TMA bulk copy
    float *h_s = w2_s + SIZE * 256;           // 2 halves x 4 k-groups x HG
    float *g_s = h_s + 2 * 4 * HG;            // 2 halves x [HC][256] gate pre-activations
    float *y_s = g_s + 2 * HC * 256;          // [C3MAX][SIZE]
    const int tid = threadIdx.x, lane = tid & 31;
    const int chunk0 = blockIdx.x * CPB;
    const int C = min(CPB, B - chunk0);
    float *hA = h_s, *hB = h_s + 4 * HG;
    float *gA = g_s, *gB = g_s + HC * 256;

    pdl_launch_dependents();  // the next forward's K1 may stage its weights under our tail
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    for (int i = tid; i < 2 * 4 * HG; i += K3_THREADS) h_s[i] = 0.f;
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, SIZE * 256 * 4u);
        bulk_g2s(w2_s, wih2T, SIZE * 256 * 4u, bar);
    }

    if (tid < K3_MV_THREADS) {
        // ================================ mat-vec warps ==========================================
        const int kg = lane & 3;  // k-group: k in [16 kg, 16 kg + 16)
        const int r = tid;        // finished gate row owned after the butterfly: i 0..63, f, g, o
        // w[i][kl] = W_hh[4*(tid>>2) + i][16*kg + kl], host layout [q][tid] float4, q = 4*i + kl/4
        float w[4][16];
#pragma unroll
        for (int q4 = 0; q4 < 16; ++q4) {
            const float4 v = whh4[q4 * 256 + tid];
            w[q4 >> 2][(q4 & 3) * 4 + 0] = v.x;
            w[q4 >> 2][(q4 & 3) * 4 + 1] = v.y;
            w[q4 >> 2][(q4 & 3) * 4 + 2] = v.z;
            w[q4 >> 2][(q4 & 3) * 4 + 3] = v.w;
        }
        const float *hkA = hA + kg * HG, *hkB = hB + kg * HG;
        const bool hi2 = (lane & 2) != 0, hi1 = (lane & 1) != 0;
        pdl_wait();  // xp is produced by K2
        const float *xrow[C3MAX];
#pragma unroll
        for (int c = 0; c < C3MAX; ++c)
            xrow[c] = xp + ((size_t)(chunk0 + (c < C ? c : 0)) * TM) * 256 + r;
        float xa[HC], xb[HC];
#pragma unroll
        for (int c = 0; c < HC; ++c) {
            xa[c] = c < C ? xrow[c][0] : 0.f;
            xb[c] = HC + c < C ? xrow[HC + c][0] : 0.f;
        }
        lstm_matvec_half(w, hkA, xa, gA, r, hi2, hi1);  // mv(A, 0)
        nbar_arrive(BAR_GA, K3_THREADS);
#pragma unroll 1
        for (int t = 0; t < TM; ++t) {
            float xa_n[HC], xb_n[HC];
            const int tn = t + 1 < TM ? t + 1 : t;
#pragma unroll
            for (int c = 0; c < HC; ++c) {
                xa_n[c] = c < C ? xrow[c][(size_t)tn * 256] : 0.f;
                xb_n[c] = HC + c < C ? xrow[HC + c][(size_t)tn * 256] : 0.f;
            }
            if (t > 0) nbar_sync(BAR_HB, K3_THREADS);  // h_B(t-1) written
            lstm_matvec_half(w, hkB, xb, gB, r, hi2, hi1);  // mv(B, t)
            nbar_arrive(BAR_GB, K3_THREADS);
            nbar_sync(BAR_HA, K3_THREADS);  // h_A(t) written
            if (t + 1 < TM) {
                lstm_matvec_half(w, hkA, xa_n, gA, r, hi2, hi1);  // mv(A, t+1)
                nbar_arrive(BAR_GA, K3_THREADS);
            }
#pragma unroll
            for (int c = 0; c < HC; ++c) xb[c] = xb_n[c];
        }
        nbar_sync(BAR_HB, K3_THREADS);  // h_B(TM-1)
    } else {
        // ================================ cell-update warps ======================================
        const int ct = tid - K3_MV_THREADS;  // 0..127: cells ct and ct+128 of each half
        const int u0 = ct & 63, q0 = ct >> 6, q1 = q0 + 2;
        const int hu = (u0 >> 4) * HG + (u0 & 15) * HC;
        float cA0 = 0.f, cA1 = 0.f, cB0 = 0.f, cB1 = 0.f;
        float hA0 = 0.f, hA1 = 0.f, hB0 = 0.f, hB1 = 0.f;
#pragma unroll 1
        for (int t = 0; t < TM; ++t) {
            nbar_sync(BAR_GA, K3_THREADS);
            hA0 = lstm_cell_half(gA, hA, u0, q0, hu, cA0);
            hA1 = lstm_cell_half(gA, hA, u0, q1, hu, cA1);
            nbar_arrive(BAR_HA, K3_THREADS);
            nbar_sync(BAR_GB, K3_THREADS);
            hB0 = lstm_cell_half(gB, hB, u0, q0, hu, cB0);
            hB1 = lstm_cell_half(gB, hB, u0, q1, hu, cB1);
            nbar_arrive(BAR_HB, K3_THREADS);
        }
        // ---- LSTM2 input: x = swish(h1[T-1]) (ConvLSTM_w_ref.py:53) -------------------------------
        // (the mat-vec warps have passed their last BAR_HB sync only after these cells' arrive, and
        //  read h again only after the __syncthreads below)
        hA[hu + q0] = swishf(hA0);
        hA[hu + q1] = swishf(hA1);
        hB[hu + q0] = swishf(hB0);
        hB[hu + q1] = swishf(hB1);
    }
    // ---- LSTM2: only the first step of the reversed pass is consumed (ConvLSTM_w_ref.py:53-54):
    // h0 = c0 = 0  =>  c = sig(i) * tanh(g), h = sig(o) * tanh(c)
    mbar_wait(bar, 0);  // W_ih2^T landed long ago
    __syncthreads();
    if (tid < K3_MV_THREADS) {
        const int r = tid;
        float a2[C3MAX];
        const float bias = b2[r];
#pragma unroll
        for (int c = 0; c < C3MAX; ++c) a2[c] = bias;
#pragma unroll 8
        for (int k = 0; k < SIZE; ++k) {
            const float wv = w2_s[k * 256 + r];
            const int ho = (k >> 4) * HG + (k & 15) * HC;
            const float4 ha = *reinterpret_cast<const float4 *>(hA + ho);
            const float4 hb = *reinterpret_cast<const float4 *>(hB + ho);
            a2[0] = fmaf(wv, ha.x, a2[0]);
            a2[1] = fmaf(wv, ha.y, a2[1]);
            a2[2] = fmaf(wv, ha.z, a2[2]);
            a2[3] = fmaf(wv, ha.w, a2[3]);
            a2[4] = fmaf(wv, hb.x, a2[4]);
            a2[5] = fmaf(wv, hb.y, a2[5]);
            a2[6] = fmaf(wv, hb.z, a2[6]);
            a2[7] = fmaf(wv, hb.w, a2[7]);
        }
#pragma unroll
        for (int c = 0; c < C3MAX; ++c) g_s[c * 256 + r] = a2[c];  // gA rows 0..3, gB rows 4..7
    }
    __syncthreads();
    if (tid < K3_MV_THREADS) {
        const int u = tid & 63, q = tid >> 6;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c = q + HC * i;
            const float *g = g_s + c * 256;
            const float c2 = sigmoidf_acc(g[u]) * tanhf(g[128 + u]);
            const float h2 = sigmoidf_acc(g[192 + u]) * tanhf(c2);
            y_s[c * SIZE + u] = swishf(h2);
        }
    }
    __syncthreads();
    // ---- fc: one warp per chunk, warp-shuffle reduction over the 64 features ---------------------
    {
        const int warp = tid >> 5;
        if (warp < C) {
            for (int o = 0; o < num_out; ++o) {
                float part = fcw[o * SIZE + lane] * y_s[warp * SIZE + lane] +
                             fcw[o * SIZE + lane + 32] * y_s[warp * SIZE + lane + 32];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
                if (lane == 0) logits[(size_t)(chunk0 + warp) * num_out + o] = part + fcb[o];
            }
        }
    }
}
constexpr int K3_SMEM_BYTES = (4 + SIZE * 256 + 2 * 4 * HG + 2 * HC * 256 + C3MAX * SIZE) * 4;

// repack a channel-last strided activation into canonical [B][C][T] for rb200_debug_tensor
__global__ void repack_kernel(const float *__restrict__ src, int64_t chunk_stride, int row_pitch,
                              int ch_off, float *__restrict__ dst, int B, int C, int T) {
    const int64_t total = (int64_t)B * C * T;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(i % T);
        const int c = (int)((i / T) % C);
        const int b = (int)(i / ((int64_t)T * C));
        dst[i] = src[b * chunk_stride + (int64_t)t * row_pitch + ch_off + c];
    }
}

}  // namespace

// =================================================================================================
// host side
// =================================================================================================
struct FusedWeights {
    float *dev = nullptr;  // one allocation holding every re-laid-out tensor
    size_t off_front = 0, off_slabs = 0, off_bmerge = 0, off_wih1T = 0, off_b1 = 0, off_whh4 = 0,
           off_wih2T = 0, off_b2 = 0, off_fcw = 0, off_fcb = 0, off_wm_tc = 0, off_wih_tc = 0, off_front_tc = 0, off_w1_tc = 0, off_wseq1_dense = 0;
    int kmer_len = 0, num_out = 0;
};

bool fused_supported(const rb200_model_desc &d) {
    if (d.arch != RB200_ARCH_CONVLSTM_W_REF || d.size != SIZE) return false;
    if (d.n_sig_conv != 3 || d.n_seq_conv != 2 || d.n_merge_conv != 1 || d.n_lstm != 2) return false;
    auto is = [](const rb200_conv_desc &c, int ci, int co, int kw, int st) {
        return c.c_in == ci && c.c_out == co && c.kw == kw && c.stride == st;
    };
    if (d.kmer_len > 16) return false;
    return is(d.sig_conv[0], 1, 4, KW_SIG1, 1) && is(d.sig_conv[1], 4, 16, KW_SIG2, 1) &&
           is(d.sig_conv[2], 16, SIZE, KW_SIG3, 3) &&
           is(d.seq_conv[0], 4 * d.kmer_len, 16, KW_SEQ1, 1) &&
           is(d.seq_conv[1], 16, SIZE, KW_SEQ2, 3) && is(d.merge_conv[0], 2 * SIZE, SIZE, KW_MRG, 1);
}

int fused_create(rb200_model *m, const float *blob) {
    const rb200_model_desc &d = m->desc;
    const int K = d.kmer_len;
    const FrontOffsets fo = front_offsets(K);
    std::vector<float> host;
    auto reserve = [&](size_t n) {
        size_t at = host.size();
        host.resize(at + ((n + 31) & ~(size_t)31), 0.f);  // 128-byte aligned pieces
        return at;
    };
    FusedWeights *fw = new FusedWeights();
    fw->kmer_len = K;
    fw->num_out = d.num_out;
    // --- K1 blob (FFMA2 variant) and its tensor-core twin (no FFMA2 weights for the stride-3 convs) ---
    auto fill_front = [&](float *f, const FrontOffsets &o, bool tc) {
        const float *w = blob + d.sig_conv[0].w_off;  // [co=4][ci=1][j=5]
        for (int j = 0; j < KW_SIG1; ++j)
            for (int co = 0; co < 4; ++co) f[o.w_sig1 + j * 4 + co] = w[co * KW_SIG1 + j];
        memcpy(f + o.b_sig1, blob + d.sig_conv[0].b_off, 4 * sizeof(float));
        w = blob + d.sig_conv[1].w_off;  // [16][4][5]
        for (int j = 0; j < KW_SIG2; ++j)
            for (int ci = 0; ci < 4; ++ci)
                for (int co = 0; co < 16; ++co)
                    f[o.w_sig2 + (j * 4 + ci) * 16 + co] = w[(co * 4 + ci) * KW_SIG2 + j];
        memcpy(f + o.b_sig2, blob + d.sig_conv[1].b_off, 16 * sizeof(float));
        if (!tc) {
            w = blob + d.sig_conv[2].w_off;  // [64][16][9]
            for (int j = 0; j < KW_SIG3; ++j)
                for (int ci = 0; ci < 16; ++ci)
                    for (int co = 0; co < SIZE; ++co)
                        f[o.w_sig3 + (j * 16 + ci) * SIZE + co] = w[(co * 16 + ci) * KW_SIG3 + j];
        }
        memcpy(f + o.b_sig3, blob + d.sig_conv[2].b_off, SIZE * sizeof(float));
        w = blob + d.seq_conv[0].w_off;  // [16][4K][5], input row = 4p + base
        for (int j = 0; j < KW_SEQ1; ++j)
            for (int row = 0; row < 4 * K; ++row)
                for (int co = 0; co < 16; ++co)
                    f[o.w_seq1 + (j * 4 * K + row) * GROW + co] = w[(co * 4 * K + row) * KW_SEQ1 + j];
        memcpy(f + o.b_seq1, blob + d.seq_conv[0].b_off, 16 * sizeof(float));
        if (!tc) {
            w = blob + d.seq_conv[1].w_off;  // [64][16][13]
            for (int j = 0; j < KW_SEQ2; ++j)
                for (int ci = 0; ci < 16; ++ci)
                    for (int co = 0; co < SIZE; ++co)
                        f[o.w_seq2 + (j * 16 + ci) * SIZE + co] = w[(co * 16 + ci) * KW_SEQ2 + j];
        }
        memcpy(f + o.b_seq2, blob + d.seq_conv[1].b_off, SIZE * sizeof(float));
    };
    fw->off_front = reserve(fo.total);
    fill_front(host.data() + fw->off_front, fo, false);
    const FrontOffsets fo_tc = front_offsets(K, true);
    fw->off_front_tc = reserve(fo_tc.total);
    fill_front(host.data() + fw->off_front_tc, fo_tc, true);
    // --- K2: merge conv slabs [s][j][c_local][m], bias, W_ih1^T [k][r], b1 ---
    fw->off_slabs = reserve((size_t)N_SLABS * SLAB_FLOATS);
    {
        const float *w = blob + d.merge_conv[0].w_off;  // [64][128][5]
        for (int s = 0; s < N_SLABS; ++s)
            for (int j = 0; j < KW_MRG; ++j)
                for (int cl = 0; cl < SLAB_C; ++cl)
                    for (int mo = 0; mo < SIZE; ++mo)
                        host[fw->off_slabs + (size_t)s * SLAB_FLOATS + (j * SLAB_C + cl) * SIZE + mo] =
                            w[(mo * 2 * SIZE + s * SLAB_C + cl) * KW_MRG + j];
    }
    fw->off_bmerge = reserve(SIZE);
    memcpy(host.data() + fw->off_bmerge, blob + d.merge_conv[0].b_off, SIZE * sizeof(float));
    fw->off_wih1T = reserve(SIZE * 256);
    for (int k = 0; k < SIZE; ++k)
        for (int r = 0; r < 256; ++r)
            host[fw->off_wih1T + k * 256 + r] = blob[d.lstm_w_ih_off[0] + r * SIZE + k];
    fw->off_b1 = reserve(256);
    memcpy(host.data() + fw->off_b1, blob + d.lstm_b_off[0], 256 * sizeof(float));
    // --- K3: W_hh1 as float4 [k/4][r], W_ih2^T, b2, fc ---
    fw->off_whh4 = reserve(SIZE * 256);
    for (int q = 0; q < 16; ++q)
        for (int tid = 0; tid < 256; ++tid)
            for (int e = 0; e < 4; ++e) {
                const int i = q >> 2, kl = (q & 3) * 4 + e;
                const int row = 4 * (tid >> 2) + i, k = 16 * (tid & 3) + kl;
                host[fw->off_whh4 + (q * 256 + tid) * 4 + e] = blob[d.lstm_w_hh_off[0] + row * SIZE + k];
            }
    fw->off_wih2T = reserve(SIZE * 256);
    for (int k = 0; k < SIZE; ++k)
        for (int r = 0; r < 256; ++r)
            host[fw->off_wih2T + k * 256 + r] = blob[d.lstm_w_ih_off[1] + r * SIZE + k];
    fw->off_b2 = reserve(256);
    memcpy(host.data() + fw->off_b2, blob + d.lstm_b_off[1], 256 * sizeof(float));
    fw->off_fcw = reserve((size_t)d.num_out * SIZE);
    memcpy(host.data() + fw->off_fcw, blob + d.fc_w_off, (size_t)d.num_out * SIZE * sizeof(float));
    fw->off_fcb = reserve(d.num_out);
    memcpy(host.data() + fw->off_fcb, blob + d.fc_b_off, d.num_out * sizeof(float));

    // --- K2-TC: merge / W_ih weights split into TF32 hi + lo, K-major, 128B-swizzled tiles ---
    auto tf32_split = [](float a, float &hi, float &lo) {
        uint32_t u;
        memcpy(&u, &a, 4);
        u = (u + 0x1000u) & 0xFFFFE000u;  // round to nearest (ties away), low 13 mantissa bits cleared
        memcpy(&hi, &u, 4);
        lo = a - hi;
    };
    auto sw = [](int row, int k) { return row * 32 + ((((k >> 2) ^ (row & 7))) << 2) + (k & 3); };
    fw->off_wm_tc = reserve((size_t)4 * 5 * 2 * 64 * 32);
    {
        const float *w = blob + d.merge_conv[0].w_off;  // [64][128][5]
        for (int cb = 0; cb < 4; ++cb)
            for (int tap = 0; tap < KW_MRG; ++tap) {
                float *st = host.data() + fw->off_wm_tc + (size_t)(cb * 5 + tap) * (2 * 64 * 32);
                for (int mo = 0; mo < SIZE; ++mo)
                    for (int k = 0; k < 32; ++k) {
                        float hi, lo;
                        tf32_split(w[(mo * 2 * SIZE + cb * 32 + k) * KW_MRG + tap], hi, lo);
                        st[sw(mo, k)] = hi;
                        st[64 * 32 + sw(mo, k)] = lo;
                    }
            }
    }
    fw->off_wih_tc = reserve((size_t)4 * 2 * 128 * 32);
    for (int nh = 0; nh < 2; ++nh)
        for (int kb = 0; kb < 2; ++kb) {
            float *st = host.data() + fw->off_wih_tc + (size_t)(nh * 2 + kb) * (2 * 128 * 32);
            for (int row = 0; row < 128; ++row)
                for (int k = 0; k < 32; ++k) {
                    float hi, lo;
                    tf32_split(blob[d.lstm_w_ih_off[0] + (nh * 128 + row) * SIZE + kb * 32 + k], hi, lo);
                    st[sw(row, k)] = hi;
                    st[128 * 32 + sw(row, k)] = lo;
                }
        }

    // K0 (dense interface): seq_conv1 weights as [input row][tap][16]
    fw->off_wseq1_dense = reserve((size_t)4 * K * KW_SEQ1 * 16);
    {
        const float *w = blob + d.seq_conv[0].w_off;  // [16][4K][5]
        for (int row = 0; row < 4 * K; ++row)
            for (int j = 0; j < KW_SEQ1; ++j)
                for (int co = 0; co < 16; ++co)
                    host[fw->off_wseq1_dense + (row * KW_SEQ1 + j) * 16 + co] = w[(co * 4 * K + row) * KW_SEQ1 + j];
    }
    // K1-TC: stride-3 conv weights as [K block][hi|lo][64][32] tiles, k = (tap - 2*kb)*16 + channel
    fw->off_w1_tc = reserve((size_t)(tc::K1TC_NKB_SIG + tc::K1TC_NKB_SEQ) * 2 * 64 * 32);
    for (int conv = 0; conv < 2; ++conv) {
        const int KW = conv ? KW_SEQ2 : KW_SIG3;
        const int nkb = conv ? tc::K1TC_NKB_SEQ : tc::K1TC_NKB_SIG;
        const float *w = blob + (conv ? d.seq_conv[1].w_off : d.sig_conv[2].w_off);  // [64][16][KW]
        for (int kb = 0; kb < nkb; ++kb) {
            // stream order = execution order: seq_conv2's K blocks first, then sig_conv3's
            float *st = host.data() + fw->off_w1_tc +
                        (size_t)((conv ? 0 : tc::K1TC_NKB_SEQ) + kb) * (2 * 64 * 32);
            for (int mo = 0; mo < SIZE; ++mo)
                for (int k = 0; k < 32; ++k) {
                    const int tap = 2 * kb + (k >> 4), c = k & 15;
                    float hi = 0.f, lo = 0.f;
                    if (tap < KW) tf32_split(w[(mo * 16 + c) * KW + tap], hi, lo);
                    st[sw(mo, k)] = hi;
                    st[64 * 32 + sw(mo, k)] = lo;
                }
        }
    }

    cudaError_t e = cudaMalloc(&fw->dev, host.size() * sizeof(float));
    if (e == cudaSuccess)
        e = cudaMemcpy(fw->dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error("fused weight upload failed: %s", cudaGetErrorString(e));
        if (fw->dev) cudaFree(fw->dev);
        delete fw;
        return RB200_ERR_CUDA;
    }
    const int max_smem = 227 * 1024;
    cudaFuncSetAttribute(k1_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(k2_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(k3_lstm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K3_SMEM_BYTES);
    cudaFuncSetAttribute(tc::k2tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(tc::k1tc_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(k0_dense_seq1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    m->fused = fw;
    return RB200_OK;
}

void fused_destroy(rb200_model *m) {
    if (!m->fused) return;
    if (m->fused->dev) cudaFree(m->fused->dev);
    delete m->fused;
    m->fused = nullptr;
}

bool fused_shape_ok(const rb200_model *m, int T, int seq_width, int map_width) {
    const Geometry g = make_geometry(T);
    if (!g.ok) return false;
    if (k1_smem(g, m->desc.kmer_len, seq_width, map_width).total_bytes > 227 * 1024) return false;
    if (k2_smem(g).total_bytes > 227 * 1024) return false;
    return true;
}

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

size_t fused_workspace_bytes(const rb200_model *m, int B, int T) {
    const Geometry g = make_geometry(T);
    return align256((size_t)B * g.cat_stride * 4) + align256((size_t)B * g.TM * 256 * 4) + 1024;
}

// Launch with the programmatic-stream-serialization attribute (PDL): the kernel's prologue may overlap
// the tail of the previous kernel in the stream; the kernel itself orders its reads with pdl_wait().
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem,
                              cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

int fused_dense_seq1(rb200_model *m, const float *enc_dense, float *q1_out, int B, int T, cudaStream_t stream) {
    const FusedWeights *fw = m->fused;
    RB200_REQUIRE(fw != nullptr, "dense seq_conv1 kernel not available for this model");
    const int rows = 4 * fw->kmer_len;
    const size_t smem0 = (4 + ((rows * T + 3) & ~3) + rows * KW_SEQ1 * 16) * sizeof(float);
    RB200_REQUIRE(smem0 <= 227 * 1024, "chunk_len %d too long for the dense fused path", T);
    k0_dense_seq1_kernel<<<B, 128, smem0, stream>>>(enc_dense, fw->dev + fw->off_wseq1_dense,
                                                   fw->dev + fw->off_front_tc + front_offsets(fw->kmer_len, true).b_seq1,
                                                   q1_out, B, T, rows);
    RB200_CUDA_TRY(cudaGetLastError());
    m->launches += 1;
    return RB200_OK;
}

// chunks per CTA: minimise (waves * chunks-per-CTA), prefer the larger CTA on ties
static int pick_cpb(int B, int cmax, int sm_count) {
    int best = 1;
    long best_cost = -1;
    for (int c = 1; c <= cmax; ++c) {
        const long ctas = (B + c - 1) / c;
        const long waves = (ctas + sm_count - 1) / sm_count;
        const long cost = waves * c;
        if (best_cost < 0 || cost <= best_cost) {
            best_cost = cost;
            best = c;
        }
    }
    return best;
}

int fused_forward_compact(rb200_model *m, Workspace &ws, const float *sigs, const int8_t *seqs,
                          int seq_width, const int16_t *maps, int map_width, const int16_t *lens,
                          int B, int T, float *logits, cudaStream_t stream, bool want_tc,
                          const float *enc_dense) {
    const FusedWeights *fw = m->fused;
    if (enc_dense) {  // dense interface: sizes of the (absent) compact arrays only feed smem layouts
        seq_width = fw->kmer_len;
        map_width = 2;
    }
    const Geometry g = make_geometry(T);
    RB200_REQUIRE(g.ok, "chunk_len %d not supported by the fused kernels", T);
    // tensor-core variants (tcgen05 3xTF32) when requested and the CTA's rows fit two M tiles
    int cpb = pick_cpb(B, want_tc && g.CL > tc::K1TC_MAX_CPB ? tc::K1TC_MAX_CPB : g.CL, m->sm_count);
    int tc_rpad = tc::rpad_for(cpb, g.T3);
    bool use_tc = want_tc && cpb * g.T3 <= 256 && tc::smem_layout(tc_rpad).total <= 227 * 1024;
    const int k1_nw =
        tc::k1tc_smem(g, cpb, fw->kmer_len, seq_width, map_width, tc_rpad, 3).total <= 227 * 1024 ? 3 : 2;
    static const bool no_k1tc = getenv("RB200_NO_K1TC") != nullptr;  // measurement aid, read once
    const bool use_tc_k1 =
        use_tc && (enc_dense != nullptr || !no_k1tc) &&
        tc::k1tc_smem(g, cpb, fw->kmer_len, seq_width, map_width, tc_rpad, k1_nw).total <= 227 * 1024 &&
        // the per-base gather sums borrow the (not yet used) tile stages
        (size_t)cpb * (map_width - 1) * KW_SEQ1 * GROW * 4 <= (size_t)4 * tc_rpad * 128;
    if (enc_dense && !use_tc_k1) {
        set_error("dense fused path not available for this shape");
        return RB200_ERR_UNSUPPORTED;
    }
    if (want_tc && !use_tc) {  // fall back to the fp32 kernels with their own best CTA size
        cpb = pick_cpb(B, g.CL, m->sm_count);
        tc_rpad = 0;
    }
    const int grid = (B + cpb - 1) / cpb;
    const size_t cat_bytes = use_tc ? align256((size_t)grid * 4 * tc_rpad * 32 * 4)
                                    : align256((size_t)B * g.cat_stride * 4);
    const size_t q1_bytes = enc_dense ? align256((size_t)B * g.Q1 * 16 * 4) : 0;
    const size_t live_bytes = cat_bytes + align256((size_t)B * g.TM * 256 * 4) + q1_bytes;
    const size_t n_cat = (size_t)B * 128 * g.T3, n_xp = (size_t)B * 256 * g.TM;
    size_t need = live_bytes + 1024;
    if (m->keep_debug) need += align256(n_cat * 4) + align256(n_xp * 4);
    int rc = ws.ensure(need);
    if (rc) return rc;
    float *cat = reinterpret_cast<float *>(ws.base);
    float *xp = reinterpret_cast<float *>(ws.base + cat_bytes);
    float *q1_buf = enc_dense ? reinterpret_cast<float *>(ws.base + cat_bytes + align256((size_t)B * g.TM * 256 * 4))
                              : nullptr;
    const K1Smem l1 = k1_smem(g, fw->kmer_len, seq_width, map_width);
    const K2Smem l2 = k2_smem(g);
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (m->profile) {
        for (int i = 0; i < 4; ++i) {
            RB200_CUDA_TRY(cudaEventCreate(&ev[i]));
            m->prof_events.push_back(ev[i]);
        }
        RB200_CUDA_TRY(cudaEventRecord(ev[0], stream));
    }
    // with profiling events between the kernels PDL cannot overlap them; launch plainly then
    const bool pdl = !m->profile;
    const int k1_tc = use_tc ? tc_rpad : 0;
    if (enc_dense) {
        const int rows = 4 * fw->kmer_len;
        const size_t smem0 = (4 + ((rows * T + 3) & ~3) + rows * KW_SEQ1 * 16) * sizeof(float);
        RB200_REQUIRE(smem0 <= 227 * 1024, "chunk_len %d too long for the dense fused path", T);
        k0_dense_seq1_kernel<<<B, 128, smem0, stream>>>(enc_dense, fw->dev + fw->off_wseq1_dense,
                                                       fw->dev + fw->off_front_tc + front_offsets(fw->kmer_len, true).b_seq1,
                                                       q1_buf, B, T, rows);
        RB200_CUDA_TRY(cudaGetLastError());
        m->launches += 1;
    }
    if (use_tc_k1) {
        const int smem1 = tc::k1tc_smem(g, cpb, fw->kmer_len, seq_width, map_width, tc_rpad, k1_nw).total;
        static long long *k1_stamps = nullptr;  // profiling aid: RB200_TC_STAMPS=1
        static const bool want_k1_stamps = getenv("RB200_TC_STAMPS") != nullptr;
        if (want_k1_stamps && !k1_stamps) cudaMalloc(&k1_stamps, 16 * sizeof(long long));
        if (pdl) {
            RB200_CUDA_TRY(launch_pdl(tc::k1tc_front_kernel, grid, THREADS, (size_t)smem1, stream, sigs,
                                      seqs, seq_width, maps, map_width, lens,
                                      (const float *)(fw->dev + fw->off_front_tc),
                                      (const float *)(fw->dev + fw->off_w1_tc), cat, B, cpb, T,
                                      fw->kmer_len, tc_rpad, cpb, k1_nw, (const float *)q1_buf, k1_stamps));
        } else {
            tc::k1tc_front_kernel<<<grid, THREADS, smem1, stream>>>(
                sigs, seqs, seq_width, maps, map_width, lens, fw->dev + fw->off_front_tc,
                fw->dev + fw->off_w1_tc, cat, B, cpb, T, fw->kmer_len, tc_rpad, cpb, k1_nw, q1_buf, k1_stamps);
        }
        if (want_k1_stamps) {
            long long h[16];
            cudaStreamSynchronize(stream);
            cudaMemcpy(h, k1_stamps, sizeof(h), cudaMemcpyDeviceToHost);
            fprintf(stderr, "[k1tc stamps, cycles] pdl_wait %lld  stage+gather %lld  seq2_loop %lld  sig12 %lld  "
                    "sig3_loop %lld  mma_drain %lld  epilogue %lld\n", h[1] - h[0], h[2] - h[1], h[3] - h[2],
                    h[4] - h[3], h[5] - h[4], h[6] - h[5], h[7] - h[6]);
        }
    } else if (pdl) {
        RB200_CUDA_TRY(launch_pdl(k1_front_kernel, grid, THREADS, l1.total_bytes, stream, sigs, seqs,
                                  seq_width, maps, map_width, lens,
                                  (const float *)(fw->dev + fw->off_front), cat, B, cpb, T,
                                  fw->kmer_len, k1_tc));
    } else {
        k1_front_kernel<<<grid, THREADS, l1.total_bytes, stream>>>(
            sigs, seqs, seq_width, maps, map_width, lens, fw->dev + fw->off_front, cat, B, cpb, T,
            fw->kmer_len, k1_tc);
    }
    RB200_CUDA_TRY(cudaGetLastError());
    if (m->profile) RB200_CUDA_TRY(cudaEventRecord(ev[1], stream));
    if (use_tc) {
        const int smem_tc = tc::smem_layout(tc_rpad).total;
        static long long *stamps_dev = nullptr;  // profiling aid: RB200_TC_STAMPS=1
        static const bool want_stamps = getenv("RB200_TC_STAMPS") != nullptr;
        if (want_stamps && !stamps_dev) cudaMalloc(&stamps_dev, 16 * sizeof(long long));
        if (pdl) {
            RB200_CUDA_TRY(launch_pdl(tc::k2tc_kernel, grid, tc::THREADS_TC, (size_t)smem_tc, stream,
                                      (const float *)cat, (const float *)(fw->dev + fw->off_wm_tc),
                                      (const float *)(fw->dev + fw->off_bmerge),
                                      (const float *)(fw->dev + fw->off_wih_tc),
                                      (const float *)(fw->dev + fw->off_b1), xp, B, cpb, g.T3, g.TM,
                                      tc_rpad, stamps_dev));
        } else {
            tc::k2tc_kernel<<<grid, tc::THREADS_TC, smem_tc, stream>>>(
                cat, fw->dev + fw->off_wm_tc, fw->dev + fw->off_bmerge, fw->dev + fw->off_wih_tc,
                fw->dev + fw->off_b1, xp, B, cpb, g.T3, g.TM, tc_rpad, stamps_dev);
        }
        if (want_stamps) {
            long long h[16];
            cudaStreamSynchronize(stream);
            cudaMemcpy(h, stamps_dev, sizeof(h), cudaMemcpyDeviceToHost);
            fprintf(stderr, "[k2tc stamps, cycles from start] pdl_wait_done %lld  A0_ready %lld  merge_issued %lld  "
                    "d_full %lld  m_ready %lld  xproj_issued %lld  d2_full0 %lld  d2_full1 %lld  end %lld\n",
                    h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0], h[5] - h[0], h[6] - h[0],
                    h[7] - h[0], h[8] - h[0], h[9] - h[0]);
        }
    } else if (pdl) {
        RB200_CUDA_TRY(launch_pdl(k2_merge_kernel, grid, THREADS, l2.total_bytes, stream,
                                  (const float *)cat, (const float *)(fw->dev + fw->off_slabs),
                                  (const float *)(fw->dev + fw->off_bmerge),
                                  (const float *)(fw->dev + fw->off_wih1T),
                                  (const float *)(fw->dev + fw->off_b1), xp, B, cpb, T));
    } else {
        k2_merge_kernel<<<grid, THREADS, l2.total_bytes, stream>>>(
            cat, fw->dev + fw->off_slabs, fw->dev + fw->off_bmerge, fw->dev + fw->off_wih1T,
            fw->dev + fw->off_b1, xp, B, cpb, T);
    }
    RB200_CUDA_TRY(cudaGetLastError());
    if (m->profile) RB200_CUDA_TRY(cudaEventRecord(ev[2], stream));
    const int cpb3 = pick_cpb(B, C3MAX, m->sm_count);
    const int grid3 = (B + cpb3 - 1) / cpb3;
    if (pdl) {
        RB200_CUDA_TRY(launch_pdl(k3_lstm_kernel, grid3, K3_THREADS, (size_t)K3_SMEM_BYTES, stream,
                                  (const float *)xp,
                                  reinterpret_cast<const float4 *>(fw->dev + fw->off_whh4),
                                  (const float *)(fw->dev + fw->off_wih2T),
                                  (const float *)(fw->dev + fw->off_b2),
                                  (const float *)(fw->dev + fw->off_fcw),
                                  (const float *)(fw->dev + fw->off_fcb), logits, B, cpb3, g.TM,
                                  fw->num_out));
    } else {
        k3_lstm_kernel<<<grid3, K3_THREADS, K3_SMEM_BYTES, stream>>>(
            xp, reinterpret_cast<const float4 *>(fw->dev + fw->off_whh4), fw->dev + fw->off_wih2T,
            fw->dev + fw->off_b2, fw->dev + fw->off_fcw, fw->dev + fw->off_fcb, logits, B, cpb3,
            g.TM, fw->num_out);
    }
    RB200_CUDA_TRY(cudaGetLastError());
    if (m->profile) RB200_CUDA_TRY(cudaEventRecord(ev[3], stream));
    m->launches += 3;
    m->last_impl = use_tc ? RB200_IMPL_FUSED_TC : RB200_IMPL_FUSED;
    if (m->keep_debug) {
        // canonical [B][C][T] copies appended behind the live buffers
        m->debug.clear();
        float *dcat = reinterpret_cast<float *>(ws.base + live_bytes);
        float *dxp = reinterpret_cast<float *>(ws.base + live_bytes + align256(n_cat * 4));
        if (!use_tc) {
            repack_kernel<<<256, 256, 0, stream>>>(cat, g.cat_stride, XP, 0, dcat, B, 128, g.T3);
            m->debug.push_back({"cat", dcat, B, 128, g.T3});
        }
        repack_kernel<<<256, 256, 0, stream>>>(xp, (int64_t)g.TM * 256, 256, 0, dxp, B, 256, g.TM);
        RB200_CUDA_TRY(cudaGetLastError());
        m->debug.push_back({"xproj", dxp, B, 256, g.TM});
    }
    return RB200_OK;
}

}  // namespace rb200
