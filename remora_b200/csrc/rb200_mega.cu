// ConvLSTM_w_ref (size 64) as ONE sm_100a kernel per batch: compact chunk arrays in, logits out,
// nothing but 8 B/chunk of logits ever leaves the SM.
//
// Reference semantics: models/ConvLSTM_w_ref.py:39-58 (eval BatchNorm folded into the convolutions),
// k-mer one-hot of src/remora/encoded_kmers.pyx:13-45 fused into seq_conv1 as a gather-add.
//
// Shape of the kernel:
//   * a CTA owns G = 4 chunks; its activations live in shared memory as MMA-ready tiles whose row index is
//     chunk * 32 + t (28 valid output steps + 4 rows of slack per chunk), i.e. exactly one M = 128 tile.
//     113 KB of shared memory, 256 TMEM columns and 256 threads x 128 registers per CTA, so TWO CTAs share
//     an SM: while one walks the 24-step LSTM chain (latency bound, FFMA2) the other one runs its
//     convolutions on the tensor cores - the hardware interleaves the two dependency chains.
//   * every GEMM-shaped layer (sig_conv3, seq_conv2, merge_conv1, LSTM1 input projection) is a
//     tcgen05.mma kind::f16 with fp32 accumulators in TMEM.  Operand tiles are K-major WITHOUT swizzle,
//     stored [16-byte K chunk][row]: an 8-row core matrix is 128 contiguous bytes (SBO = 128 B), K chunks
//     are LBO = rows * 16 B apart, and a descriptor whose start address is advanced by i * 16 B reads the
//     tile shifted down by i rows.  A stride-1 convolution tap is therefore a descriptor shift (no
//     im2col); for the two stride-3 convolutions the producing layer writes its output de-interleaved by
//     t mod 3 into three tiles, which turns tap j = 3 i + r into "tile r shifted by i rows".
//   * fp32 parity (1e-4 on the logits) with half-precision tensor-core operands: MODE 0 splits every
//     operand into fp16 hi + lo (22 significant bits, weights pre-scaled by a power of two per layer so
//     their lo parts stay normal) and issues a*b ~= ah*bh + ah*bl + al*bh.  ah*[bh;bl] is ONE N = 128
//     MMA (the weight tile stacks hi and lo rows), al*bh an N = 64 MMA into the correction columns.
//     Compared with 3xTF32 the operand bytes and the instruction count halve.  MODE 1 is the bf16 variant
//     (BASELINE.json configs[1]): one pass, bf16 operands, fp32 accumulate, fp32 gates.
//   * weights stream from L2 through a 4-stage ring of 8 KB TMA bulk copies (mbarrier full/empty,
//     tcgen05.commit frees a stage); one elected thread issues TMA and MMA.
//   * the recurrence keeps W_hh in registers (thread = the four gates of one hidden unit x 16 k; a transposing
//     shuffle butterfly leaves the four gate pre-activations of one (unit, chunk) cell in one lane, so the
//     cell update never leaves the registers: one barrier per step), reads the input projection from shared
//     memory ([t][chunk][256], drained from TMEM once), then does the single needed step of the reversed
//     LSTM2 and the classifier.
//   * consecutive batches overlap: the kernel triggers its dependents at once (PDL) and only orders its
//     final 8 B/chunk store after the previous grid (griddepcontrol.wait), it has no global scratch.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "rb200_internal.cuh"

namespace rb200 {
namespace mega {

constexpr int SIZE = 64;
constexpr int G = 4;                 // chunks per CTA
constexpr int U = 32;                // tile rows per chunk
constexpr int MROWS = G * U;         // 128 = MMA M
constexpr int RP = MROWS + 8;        // rows stored per K chunk (tap shifts read up to 4 rows past row 127)
constexpr int LBO_A = RP * 16;       // byte distance between K chunks of an activation tile (2176)
constexpr int THREADS = 256;
constexpr int KW_SIG1 = 5, KW_SIG2 = 5, KW_SIG3 = 9, KW_SEQ1 = 5, KW_SEQ2 = 13, KW_MRG = 5;
constexpr int GROW = 20;             // floats per gather-table row (16 + 4 pad)
constexpr int STAGE_BYTES = 8192;    // one weight stage
constexpr int RING = 4;
constexpr int TMEM_COLS = 256;
constexpr int HG = 16 * G + 4;       // floats per k-group of h
constexpr int XCS = 256 + 8;         // floats per (time step, chunk) of the staged input projection: +8 keeps the
                                     // recurrence's (unit, chunk) reads on distinct banks
constexpr int XPS = G * XCS + 4;     // floats per time step (= 4 mod 32: E3's float4 rows hit distinct banks)
constexpr int MAX_T3 = U - (KW_MRG - 1);  // 28
constexpr int MAX_TM = MAX_T3 - (KW_MRG - 1);  // 24

// ---- constants blob (floats), copied to shared memory once per CTA ------------------------------------
constexpr int C_WSIG1 = 0;      // [j][co]            20
constexpr int C_BSIG1 = 20;     //                     4
constexpr int C_WSIG2 = 24;     // [j][ci][co]       320
constexpr int C_BSIG2 = 344;    //                    16
constexpr int C_BSEQ1 = 360;    //                    16
constexpr int C_BSIG3 = 376;    //                    64
constexpr int C_BSEQ2 = 440;    //                    64
constexpr int C_BMRG = 504;     //                    64
constexpr int C_SCALE = 568;    // inverse weight scales: seq2, sig3, merge, xproj, W_hh1, seq1, W_ih2 (+1 pad)
constexpr int C_B1 = 576;       // LSTM1 bias        256
constexpr int CONST_FLOATS = 832;
constexpr int CONST_BYTES = CONST_FLOATS * 4;  // 3328

// ---- shared memory map (bytes) ------------------------------------------------------------------------
constexpr int OFF_BARS = 0;
constexpr int OFF_CONST = 256;
constexpr int OFF_TILES = 3840;                          // h, g, y of the recurrence
constexpr int TILES_BYTES = (4 * HG + G * 256 + G * SIZE) * 4;  // 6208
constexpr int OFF_RING = 10112;
constexpr int OFF_A = OFF_RING + RING * STAGE_BYTES;     // 42880
constexpr int XT_BYTES = 2 * LBO_A;                      // one residue tile, one of hi/lo: 2 K chunks
constexpr int XSET_BYTES = 6 * XT_BYTES;                 // 3 residues x {hi, lo} = 26112
constexpr int CAT_HALF = 16 * LBO_A;                     // 34816: 128 channels, hi (or lo)
constexpr int M_HALF = 8 * LBO_A;                        // 17408
constexpr int A_BYTES = 2 * CAT_HALF;                    // 69632
constexpr int A_GS = 0, GS_CAP = 32768;                  // per-base gather sums
constexpr int A_XS = 0;                                  // signal-track tiles (over the dead gather sums)
constexpr int A_TAB = 32768, TAB_CAP = XSET_BYTES;       // gather table, then the sequence-track tiles
constexpr int A_XQ = 32768;
constexpr int A_S1 = A_TAB + TAB_CAP;                    // 58880: sig_conv1 output [G][T1][4] fp32
constexpr int A_STG = A_S1 + G * 96 * 16;                // 65024: staged compact inputs
constexpr int STG_SIG = 0, STG_SIDX = 1600, STG_SEQ = 2400, STG_MAP = 3040, STG_LEN = 4080;
constexpr int STG_NOBASE = 4096;                         // 16 bytes of -1: the k-mer of a sample no base covers
constexpr int MAX_T = 100, MAX_SEQ_W = 160, MAX_MAP_W = 130;
constexpr int SMEM_BYTES = OFF_A + A_BYTES;              // 112512  (two CTAs per SM: <= 115712)
static_assert(OFF_TILES + TILES_BYTES <= OFF_RING, "tiles overlap the ring");
static_assert(A_STG + STG_NOBASE + 16 <= A_BYTES, "staging does not fit region A");
static_assert(OFF_RING + MAX_TM * XPS * 4 <= SMEM_BYTES, "staged projection does not fit");
static_assert(SMEM_BYTES <= 115712, "two CTAs per SM need <= 113 KB each");

struct Bars {
    uint64_t w_full[RING], w_empty[RING], front, seq_done, conv_done, mrg_done, xp_done;
    uint32_t tmem_base;
};

struct Params {
    const float *sigs;
    const int8_t *seqs;
    const int16_t *maps;
    const int16_t *lens;
    int seq_width, map_width, B, T, kmer_len, num_out;
    const float *consts;     // CONST_FLOATS
    const float *gtab;       // seq_conv1 weights as mma.sync B fragments [tap][k-tile][n-tile][hi, lo][lane][2]
    int gtab_bytes;
    const uint8_t *wstream;  // weight stages in execution order
    const float *q1_in;      // dense interface: seq_conv1 output [B][T-4][16] (K0 kernel), else null
    const float4 *whh4;      // W_hh1 in the register layout of the recurrence
    const float *wih2T, *b2, *fcw, *fcb;
    float *logits;
    float *dbg_cat, *dbg_m, *dbg_xp;  // optional canonical [B][C][T] copies of the intermediates
    int *flags;                       // [0] != 0: an activation left the fp16 range (MODE 0)
    long long *stamps;                // optional phase timestamps of CTA 0 (profiling aid)
    long long *trace;                 // optional per-CTA record {smid, start ns, end ns, cycles} (profiling aid)
    // fused exchange step (multi-GPU): the classifier epilogue stores the logits into every rank's buffer
    float *const *peers;              // n_peers base pointers (peer-mapped over NVLink), or null
    int n_peers;
    long long peer_off;               // float offset of this call's [B][num_out] block in every buffer
    float *mc_base;                   // NVLS multicast alias of the same buffers (one store reaches all), or null
    long long flag_off;               // >= 0: uint32 slot (counted in 4-byte words) incremented once per CTA
    // deferred form of the exchange: one extra CTA (the last of the grid) waits for the PREVIOUS grid and ships
    // that grid's finished block to every peer, so no compute CTA ever waits on a remote store
    int ship_cta;                     // 1: gridDim.x - 1 is the shipping CTA
    int self_rank;
    const float *ship_src;            // local block of an earlier launch on this stream (may be null: nothing yet)
    long long ship_off;               // its float offset inside every buffer
    int ship_n;                       // floats
    long long ship_flag;              // >= 0: uint32 word bumped by one on every rank after the block is out
};

// ---- PTX helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ float2 ffma2(float2 a, float b, float2 c) { return __ffma2_rn(a, make_float2(b, b), c); }

// K-major, no swizzle: start address, LBO = distance between 16-byte K chunks, SBO = 128 B (8 rows x 16 B)
__device__ __forceinline__ uint64_t desc_ns(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) |
           ((uint64_t)1 << 46);
}
// D = f32; A, B = f16 (0) or bf16 (1), both K-major; N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t idesc_h(int M, int N, int bf16) {
    return (1u << 4) | ((uint32_t)bf16 << 7) | ((uint32_t)bf16 << 10) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_h(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// accumulate flag known at compile time: no predicate set-up in the issuing thread's instruction stream
template <bool ACC>
__device__ __forceinline__ void mma_hc(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a), "l"(b), "r"(idesc), "n"(ACC ? 1 : 0)
        : "memory");
}
// descriptor of the same tile `byte_off` further on (the start-address field counts 16-byte units and never
// carries out of its 14 bits for shared-memory addresses)
__device__ __forceinline__ uint64_t desc_at(uint64_t d, uint32_t byte_off) { return d + (uint64_t)(byte_off >> 4); }
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
        "%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// eight consecutive columns of this thread's TMEM lane (issue only: pair with tmem_ld_wait)
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void nbar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ---- fp32 -> operand conversion: 8 consecutive channels = one 16-byte K chunk --------------------------
// MODE 0: hi = rn_f16(a) (saturating), lo = rn_f16(a - hi); MODE 1: bf16(a), no lo tile.
template <int MODE>
__device__ __forceinline__ uint32_t pack2(float a0, float a1) {
    uint32_t r;
    if (MODE == 0)
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a1), "f"(a0));
    else
        asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a1), "f"(a0));
    return r;
}
__device__ __forceinline__ float2 unpack_h2(uint32_t r) {
    __half2 h;
    memcpy(&h, &r, 4);
    return __half22float2(h);
}
template <int MODE>
__device__ __forceinline__ void store_chunk8(uint8_t *tile_hi, uint8_t *tile_lo, int off, const float (&v)[8]) {
    uint32_t hi[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) hi[e] = pack2<MODE>(v[2 * e], v[2 * e + 1]);
    *reinterpret_cast<uint4 *>(tile_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (MODE == 0) {
        uint32_t lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 h = unpack_h2(hi[e]);
            lo[e] = pack2<0>(v[2 * e] - h.x, v[2 * e + 1] - h.y);
        }
        *reinterpret_cast<uint4 *>(tile_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// ---- weight stream geometry ----------------------------------------------------------------------------
// MODE 0: B tiles stack hi rows then lo rows (N = 128; projection: separate hi / lo stages of N = 256).
// MODE 1: N = 64 (projection 256), twice the K extent per 8 KB stage.
template <int MODE>
struct Cfg {
    static constexpr int NB = MODE == 0 ? 128 : 64;                // rows of a conv / merge B tile
    static constexpr int B_LBO = NB * 16;                          // K-chunk distance in those tiles
    static constexpr int TAPS = STAGE_BYTES / (2 * B_LBO);         // stride-3 conv taps per stage: 2 | 4
    static constexpr int NS_SEQ = (KW_SEQ2 + TAPS - 1) / TAPS;     // 7 | 4
    static constexpr int NS_SIG = (KW_SIG3 + TAPS - 1) / TAPS;     // 5 | 3
    static constexpr int MRG_KC = STAGE_BYTES / B_LBO;             // K chunks (8 channels) per merge stage: 4 | 8
    static constexpr int NS_MRG = KW_MRG * (16 / MRG_KC);          // 20 | 10
    static constexpr int NS_XP = MODE == 0 ? 8 : 4;                // projection: k16 x {hi, lo} | k16
    static constexpr int NS_TOTAL = NS_SEQ + NS_SIG + NS_MRG + NS_XP;
};

__device__ __forceinline__ void load_stage(int s, int total, const uint8_t *wstream, uint8_t *ring, Bars *bars) {
    if (s >= total) return;
    const int slot = s & (RING - 1);
    if (s >= RING) mbar_wait(&bars->w_empty[slot], ((s >> 2) - 1) & 1);
    mbar_expect_tx(&bars->w_full[slot], STAGE_BYTES);
    bulk_g2s(ring + slot * STAGE_BYTES, wstream + (size_t)s * STAGE_BYTES, STAGE_BYTES, &bars->w_full[slot]);
}

// one stride-3 convolution (16 input channels): tap j = 3 i + r reads residue tile r shifted by i rows
// the producer side of the weight ring for the phases in which every other warp waits for the tensor core
// anyway: one thread keeps RING stages in flight (a stage is re-filled as soon as the MMAs that read it
// complete), so the MMA-issuing thread never blocks on a free slot
__device__ __forceinline__ void produce_until(int &next, int upto, int total, const uint8_t *wstream, uint8_t *ring,
                                              Bars *bars) {
    const int end = upto < total ? upto : total;
    for (; next < end; ++next) load_stage(next, total, wstream, ring, bars);
}

// The issuing thread's loop is fully unrolled with compile-time stage numbers: ring slot, barrier parity and
// every descriptor offset are immediates, so one MMA costs a couple of integer adds to issue.  (With run-time
// stage counters the descriptor arithmetic of the single issuing thread, ~125 cycles per MMA, was what bounded
// the tensor-core phases - the MMAs themselves run at 64 (N = 128) / 50 (N = 64) cycles,
// scripts/microbench/umma_rate.cu.)
template <int MODE, int KW, int S0, bool SELF_LOAD>
__device__ __forceinline__ void issue_conv3(uint32_t xset, uint32_t d_tmem, const uint8_t *wstream, uint8_t *ring,
                                            Bars *bars, long long *dbg = nullptr) {
    using C = Cfg<MODE>;
    constexpr int NS = (KW + C::TAPS - 1) / C::TAPS;
    constexpr uint32_t id_main = idesc_h(128, C::NB, MODE), id_corr = idesc_h(128, 64, MODE);
    const uint64_t da = desc_ns(xset, LBO_A), db = desc_ns(smem_addr(ring), C::B_LBO);
#pragma unroll
    for (int st = 0; st < NS; ++st) {
        const int s = S0 + st, slot = s & (RING - 1);
        mbar_wait(&bars->w_full[slot], (s >> 2) & 1);
        if (dbg) dbg[st] = clock64();
        tc_fence_after();
#pragma unroll
        for (int tp = 0; tp < C::TAPS; ++tp) {
            const int j = st * C::TAPS + tp;
            if (j < KW) {
                const int r = j % 3, i = j / 3;
                const uint64_t a_hi = desc_at(da, (2 * r) * XT_BYTES + i * 16);
                const uint64_t b = desc_at(db, slot * STAGE_BYTES + tp * 2 * C::B_LBO);
                if (j == 0)
                    mma_hc<false>(d_tmem, a_hi, b, id_main);
                else
                    mma_hc<true>(d_tmem, a_hi, b, id_main);
                if (MODE == 0) mma_hc<true>(d_tmem + 64, desc_at(a_hi, XT_BYTES), b, id_corr);
            }
        }
        umma_commit(&bars->w_empty[slot]);
        if (SELF_LOAD) load_stage(s + RING - 1, C::NS_TOTAL, wstream, ring, bars);
    }
}

// ---- recurrence helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ float2 shfl_xor2(float2 v, int mask) {
    return make_float2(__shfl_xor_sync(0xffffffffu, v.x, mask), __shfl_xor_sync(0xffffffffu, v.y, mask));
}
__device__ __forceinline__ float2 sel2(bool take_a, float2 a, float2 b) {
    return make_float2(take_a ? a.x : b.x, take_a ? a.y : b.y);
}
// =========================================================================================================
template <int MODE>
__global__ void __launch_bounds__(THREADS, 2) mega_kernel(const __grid_constant__ Params p) {
    using CF = Cfg<MODE>;
    extern __shared__ __align__(128) uint8_t sm[];
    Bars *bars = reinterpret_cast<Bars *>(sm + OFF_BARS);
    float *cst = reinterpret_cast<float *>(sm + OFF_CONST);
    float *h_s = reinterpret_cast<float *>(sm + OFF_TILES);
    float *g_s = h_s + 4 * HG;
    float *y_s = g_s + G * 256;
    uint8_t *ring = sm + OFF_RING;
    uint8_t *ra = sm + OFF_A;
    float *xp_s = reinterpret_cast<float *>(sm + OFF_RING);  // recurrence phase: over the ring and region A

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunk0 = blockIdx.x * G;
    const int C = min(G, p.B - chunk0);
    const int T = p.T, T1 = T - (KW_SIG1 - 1), T2 = T1 - (KW_SIG2 - 1), Q1 = T - (KW_SEQ1 - 1);
    const int T3 = (T2 - KW_SIG3) / 3 + 1, TM = T3 - (KW_MRG - 1);
    const int seq_width = p.seq_width, map_width = p.map_width, K = p.kmer_len;

    long long *trace_st = reinterpret_cast<long long *>(sm + OFF_BARS + 120);  // 16 phase stamps of this CTA
#define MG_STAMP(i)                                                              \
    do {                                                                         \
        if (p.stamps && blockIdx.x == 0 && tid == 0) p.stamps[i] = clock64();    \
        if (p.trace && tid == 0) trace_st[i] = clock64();                        \
    } while (0)
    MG_STAMP(0);
    long long trace_t0 = 0, trace_c0 = 0;
    if (p.trace && tid == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace_t0));
        trace_c0 = clock64();
    }
    pdl_launch_dependents();  // the next batch may start as soon as SM resources free up
    if (p.ship_cta && blockIdx.x == gridDim.x - 1) {
        // the exchange step, one step behind the compute: the previous grid's block is complete and visible
        // after griddepcontrol.wait; peer stores leave from this CTA only (NVLink / NVSwitch)
        pdl_wait();
        if (p.ship_src != nullptr) {
            const float4 *src = reinterpret_cast<const float4 *>(p.ship_src);
            const int n4 = p.ship_n >> 2;
            if (p.mc_base != nullptr) {
                for (int i = tid; i < n4; i += THREADS) {
                    const float4 v = __ldcg(src + i);
                    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(
                                     p.mc_base + p.ship_off + 4 * (size_t)i),
                                 "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                                 : "memory");
                }
                for (int i = 4 * n4 + tid; i < p.ship_n; i += THREADS)
                    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p.mc_base + p.ship_off + i),
                                 "f"(__ldcg(p.ship_src + i))
                                 : "memory");
            } else {
                for (int i = tid; i < n4; i += THREADS) {
                    const float4 v = __ldcg(src + i);
                    for (int r = 0; r < p.n_peers; ++r)
                        if (r != p.self_rank) reinterpret_cast<float4 *>(p.peers[r] + p.ship_off)[i] = v;
                }
                for (int i = 4 * n4 + tid; i < p.ship_n; i += THREADS) {
                    const float v = __ldcg(p.ship_src + i);
                    for (int r = 0; r < p.n_peers; ++r)
                        if (r != p.self_rank) p.peers[r][p.ship_off + i] = v;
                }
            }
            if (p.ship_flag >= 0) {
                __threadfence_system();
                __syncthreads();
                if (tid < p.n_peers)
                    atomicAdd_system(reinterpret_cast<unsigned int *>(p.peers[tid]) + p.ship_flag, 1u);
            }
        }
        return;
    }
    // Prologue without a block-wide barrier: thread 0 initialises the mbarriers and starts the TMA loads
    // (constants, seq_conv1 weight fragments, first weight stages) at once; warp 2 allocates tensor memory
    // meanwhile; everybody else goes straight to staging the inputs.  The barrier that closes the staging phase
    // publishes the mbarriers and the TMEM base address.
    if (tid == 0) {
        for (int i = 0; i < RING; ++i) {
            mbar_init(&bars->w_full[i], 1);
            mbar_init(&bars->w_empty[i], 1);
        }
        mbar_init(&bars->front, 1);
        mbar_init(&bars->seq_done, 1);
        mbar_init(&bars->conv_done, 1);
        mbar_init(&bars->mrg_done, 1);
        mbar_init(&bars->xp_done, 1);
        mbar_fence_init();
        const bool dense = p.q1_in != nullptr;  // the materialised one-hot went through K0: no gather here
        mbar_expect_tx(&bars->front, CONST_BYTES + (dense ? 0 : p.gtab_bytes));
        bulk_g2s(cst, p.consts, CONST_BYTES, &bars->front);
        if (!dense) bulk_g2s(ra + A_TAB, p.gtab, p.gtab_bytes, &bars->front);
        const long long tl0 = p.stamps && blockIdx.x == 0 ? clock64() : 0;
        for (int s = 0; s < RING - 1; ++s) load_stage(s, CF::NS_TOTAL, p.wstream, ring, bars);
        if (p.stamps && blockIdx.x == 0) {  // profiling aid: how long the first TMA loads take to land
            mbar_wait(&bars->w_full[0], 0);
            const long long tl1 = clock64();
            mbar_wait(&bars->w_full[RING - 2], 0);
            p.stamps[15] = (tl1 - tl0) | ((clock64() - tl0) << 32);
        }
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_addr(&bars->tmem_base)),
                     "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        tc_fence_before();
    }

    MG_STAMP(1);
    // ---- P0: stage the compact inputs, move-table expansion ------------------------------------------------
    float *sig_s = reinterpret_cast<float *>(ra + A_STG + STG_SIG);
    int16_t *sidx_s = reinterpret_cast<int16_t *>(ra + A_STG + STG_SIDX);
    int8_t *seq_s = reinterpret_cast<int8_t *>(ra + A_STG + STG_SEQ);
    int16_t *map_s = reinterpret_cast<int16_t *>(ra + A_STG + STG_MAP);
    int *len_s = reinterpret_cast<int *>(ra + A_STG + STG_LEN);
    for (int i = tid; i < C * T; i += THREADS) {
        sig_s[i] = p.sigs[(size_t)chunk0 * T + i];
        sidx_s[i] = -1;
    }
    const bool dense = p.q1_in != nullptr;
    if (!dense) {
        // every load of the compact arrays is issued at once (one memory round trip): each element reads its
        // chunk's length itself and masks what lies past it (that padding is uninitialised in the reference's
        // arrays and never used)
        if (tid < C) {
            int L = p.lens[chunk0 + tid];
            len_s[tid] = max(0, min(L, min(map_width - 1, seq_width - K + 1)));
        }
        if (tid >= 32 && tid < 36) reinterpret_cast<uint32_t *>(ra + A_STG + STG_NOBASE)[tid - 32] = 0xFFFFFFFFu;
        for (int i = tid; i < C * seq_width; i += THREADS) {
            const int c = i / seq_width, s = i - c * seq_width;
            const int L = max(0, min((int)p.lens[chunk0 + c], min(map_width - 1, seq_width - K + 1)));
            const int8_t v = p.seqs[(size_t)(chunk0 + c) * seq_width + s];
            seq_s[i] = s < L + K - 1 ? v : (int8_t)-1;
        }
        for (int i = tid; i < C * map_width; i += THREADS) {
            const int c = i / map_width, s = i - c * map_width;
            const int L = max(0, min((int)p.lens[chunk0 + c], min(map_width - 1, seq_width - K + 1)));
            const int16_t v = p.maps[(size_t)(chunk0 + c) * map_width + s];
            map_s[i] = s <= L ? v : (int16_t)0;
        }
        __syncthreads();
        for (int i = tid; i < C * (map_width - 1); i += THREADS) {
            const int c = i / (map_width - 1), s = i - c * (map_width - 1);
            if (s < len_s[c]) {
                const int st = max((int)map_s[c * map_width + s], 0);
                const int en = min((int)map_s[c * map_width + s + 1], T);
                for (int t = st; t < en; ++t) sidx_s[c * T + t] = (int16_t)s;
            }
        }
    }
    if (dense) __syncthreads();  // (the other form has passed a barrier above) the mbarriers are initialised
    mbar_wait(&bars->front, 0);  // constants and the seq_conv1 weight fragments have landed
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    MG_STAMP(2);

    // ---- P1: seq_conv1 on the (virtual) one-hot = gather-add of weight columns -> residue tiles ------------
    const int LM = map_width - 1;
    const bool two_stage = dense || C * LM * KW_SEQ1 * GROW * 4 <= GS_CAP;
    uint8_t *xq = ra + (two_stage ? A_XQ : A_XS);
    if (dense) {
        // dense interface: q1 = swish(seq_conv1(enc_kmers)) was computed by K0 from the materialised one-hot;
        // split it into the residue tiles exactly as the gather would have
        for (int i = tid; i < C * Q1; i += THREADS) {
            const int c = i / Q1, t = i - c * Q1;
            const float4 *src = reinterpret_cast<const float4 *>(p.q1_in + ((size_t)(chunk0 + c) * Q1 + t) * 16);
            const int r = t % 3, u = t / 3;
            uint8_t *t_hi = xq + (2 * r) * XT_BYTES;
            const int off = (c * U + u) * 16;
#pragma unroll
            for (int kc = 0; kc < 2; ++kc) {
                const float4 v0 = __ldg(src + 2 * kc), v1 = __ldg(src + 2 * kc + 1);
                const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                store_chunk8<MODE>(t_hi, t_hi + XT_BYTES, kc * LBO_A + off, v);
            }
        }
    } else {
        // seq_conv1 weights as mma.sync B fragments (fp16 hi + lo, the products with a one-hot are exact); the A
        // fragment of a row is built in registers from ONE sequence byte: lane (g, t4) owns k-mer position
        // 4 kt + t4, its slots {2 t4, 2 t4 + 1} are bases 0/1 and {2 t4 + 8, 2 t4 + 9} bases 2/3, so the fragment is
        // 1.0 at the base that is there and 0 elsewhere (N bases, no base, positions past the k-mer: all zero)
        const uint2 *bt = reinterpret_cast<const uint2 *>(ra + A_TAB);
        const int KT = (K + 3) >> 2;
        const int8_t *stg8 = reinterpret_cast<const int8_t *>(ra + A_STG);
        const int g8 = lane >> 2, t4 = lane & 3;
        const float inv1 = cst[C_SCALE + 5];
        float *gs = reinterpret_cast<float *>(ra + A_GS);
        if (two_stage) {
            // every sample covered by the same base shares its k-mer: the k columns are summed once per (base,
            // tap) - a [chunk x base] x [k-mer one-hot] x [tap x channel] GEMM, work unit = (16 bases, tap)
            const int n_rows = C * LM, n_mt = (n_rows + 15) >> 4;
            for (int unit = warp; unit < n_mt * KW_SEQ1; unit += THREADS / 32) {
                const int mt = unit / KW_SEQ1, j = unit - mt * KW_SEQ1;
                int rb[2], rr[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int R = 16 * mt + g8 + 8 * h, c = R / LM, sb = R - c * LM;
                    rr[h] = R;
                    rb[h] = (R < n_rows && sb < len_s[min(c, G - 1)] ? STG_SEQ + c * seq_width + sb : STG_NOBASE) + t4;
                }
                float acc[2][4];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
                for (int kt = 0; kt < 4; ++kt) {
                    if (kt >= KT) break;
                    uint32_t af[4];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int base = stg8[rb[h] + 4 * kt];
                        const uint32_t one = 0x3C00u << ((base & 1) << 4);  // fp16 1.0 in the low or the high half
                        af[h] = (base >> 1) == 0 ? one : 0u;                // bases 0 / 1
                        af[2 + h] = (base >> 1) == 1 ? one : 0u;            // bases 2 / 3
                    }
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int part = 0; part < 2; ++part) {
                            const uint2 bf = bt[((((j * KT + kt) * 2 + nt) * 2 + part) << 5) + lane];
                            asm volatile(
                                "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
                                "{%8,%9}, {%0,%1,%2,%3};"
                                : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
                                : "r"(af[0]), "r"(af[1]), "r"(af[2]), "r"(af[3]), "r"(bf.x), "r"(bf.y));
                        }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    if (rr[h] < n_rows) {
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt)
                            *reinterpret_cast<float2 *>(gs + ((size_t)rr[h] * KW_SEQ1 + j) * GROW + 8 * nt + 2 * t4) =
                                make_float2(acc[nt][2 * h] * inv1, acc[nt][2 * h + 1] * inv1);
                    }
            }
            __syncthreads();  // sums complete; the table is dead: the sequence tiles may overwrite it
        }
        const float *b = cst + C_BSEQ1;
        if (two_stage) {
            for (int i = tid; i < C * Q1; i += THREADS) {
                const int c = i / Q1, t = i - c * Q1;
                float2 a[8];
#pragma unroll
                for (int o = 0; o < 8; ++o) a[o] = make_float2(b[2 * o], b[2 * o + 1]);
#pragma unroll
                for (int j = 0; j < KW_SEQ1; ++j) {
                    const int sb = sidx_s[c * T + t + j];
                    if (sb < 0) continue;  // sample not covered by any base: no one-hot entries
                    const float4 *gv =
                        reinterpret_cast<const float4 *>(gs + ((size_t)(c * LM + sb) * KW_SEQ1 + j) * GROW);
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
                const int r = t % 3, u = t / 3;
                uint8_t *t_hi = xq + (2 * r) * XT_BYTES;
                const int off = (c * U + u) * 16;
#pragma unroll
                for (int kc = 0; kc < 2; ++kc) {
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        v[2 * e] = swishf_fast(a[4 * kc + e].x);
                        v[2 * e + 1] = swishf_fast(a[4 * kc + e].y);
                    }
                    store_chunk8<MODE>(t_hi, t_hi + XT_BYTES, kc * LBO_A + off, v);
                }
            }
        } else {
            // more bases per chunk than the per-base sums have room for: the whole layer as ONE implicit GEMM over
            // rows = (chunk, output step): 4 x 96 = 24 tiles of 16, three per warp; K = (tap, k-mer position, base)
            int rowbase[3][2][KW_SEQ1];  // byte offset (in the staging area) of the k-mer position t4 of (row, tap)
#pragma unroll
            for (int mt = 0; mt < 3; ++mt) {
                const int mtile = 3 * warp + mt, c = mtile / 6, tb = (mtile - 6 * c) * 16 + g8;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int t = tb + 8 * h;
                    const bool live = c < C && t < Q1;
#pragma unroll
                    for (int j = 0; j < KW_SEQ1; ++j) {
                        const int sb = live ? (int)sidx_s[c * T + t + j] : -1;
                        rowbase[mt][h][j] = (sb < 0 ? STG_NOBASE : STG_SEQ + c * seq_width + sb) + t4;
                    }
                }
            }
            float acc[3][2][4];
#pragma unroll
            for (int mt = 0; mt < 3; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
#pragma unroll
            for (int j = 0; j < KW_SEQ1; ++j) {
#pragma unroll
                for (int kt = 0; kt < 4; ++kt) {  // KT <= 4 (k-mers up to 16 bases): unrolled so that loads run ahead
                    if (kt >= KT) break;
                    uint2 bf[2][2];
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int part = 0; part < 2; ++part)
                            bf[nt][part] = bt[((((j * KT + kt) * 2 + nt) * 2 + part) << 5) + lane];
#pragma unroll
                    for (int mt = 0; mt < 3; ++mt) {
                        uint32_t a[4];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int base = stg8[rowbase[mt][h][j] + 4 * kt];
                            const uint32_t one = 0x3C00u << ((base & 1) << 4);  // fp16 1.0 in the low or the high half
                            a[h] = (base >> 1) == 0 ? one : 0u;                 // bases 0 / 1
                            a[2 + h] = (base >> 1) == 1 ? one : 0u;             // bases 2 / 3
                        }
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                            for (int part = 0; part < 2; ++part)
                                asm volatile(
                                    "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
                                    "{%8,%9}, {%0,%1,%2,%3};"
                                    : "+f"(acc[mt][nt][0]), "+f"(acc[mt][nt][1]), "+f"(acc[mt][nt][2]), "+f"(acc[mt][nt][3])
                                    : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(bf[nt][part].x), "r"(bf[nt][part].y));
                    }
                }
            }
            // bias + swish -> residue tiles (row g8 of a fragment holds c0, c1, row g8 + 8 holds c2, c3; channels
            // 8 nt + 2 t4, + 1)
#pragma unroll
            for (int mt = 0; mt < 3; ++mt) {
                const int mtile = 3 * warp + mt, c = mtile / 6, tb = (mtile - 6 * c) * 16 + g8;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int t = tb + 8 * h;
                    if (t >= Q1) continue;
                    const int r = t % 3, uu = t / 3;
                    uint8_t *t_hi = xq + (2 * r) * XT_BYTES + (c * U + uu) * 16 + 4 * t4;
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        const float o0 = swishf_fast(fmaf(acc[mt][nt][2 * h], inv1, b[8 * nt + 2 * t4]));
                        const float o1 = swishf_fast(fmaf(acc[mt][nt][2 * h + 1], inv1, b[8 * nt + 2 * t4 + 1]));
                        const uint32_t hi = pack2<MODE>(o0, o1);
                        *reinterpret_cast<uint32_t *>(t_hi + nt * LBO_A) = hi;
                        if (MODE == 0) {
                            const float2 hf = unpack_h2(hi);
                            *reinterpret_cast<uint32_t *>(t_hi + XT_BYTES + nt * LBO_A) = pack2<0>(o0 - hf.x, o1 - hf.y);
                        }
                    }
                }
            }
        }
    }
    fence_async_smem();  // generic-proxy tile writes -> tensor-core (async proxy) reads
    __syncthreads();
    MG_STAMP(3);

    // ---- M1 (warp 0) || P2 (warps 2..7): seq_conv2 on the tensor core under sig_conv1 / sig_conv2 ----------
    // warp 1's lane 0 is the TMA producer of the weight ring for the rest of the kernel: it keeps RING stages in
    // flight and re-fills a slot as soon as the MMAs that read it have completed, so the issuing thread only
    // ever waits for data (when it also produced, every stage cost it a completion round trip)
    int p_next = RING - 1;
    if (warp == 0) {
        if (lane == 0) {
            issue_conv3<MODE, KW_SEQ2, 0, false>(smem_addr(xq), tmem, p.wstream, ring, bars,
                                                 p.stamps && blockIdx.x == 0 ? p.stamps + 25 : nullptr);
            umma_commit(&bars->seq_done);
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            produce_until(p_next, CF::NS_SEQ + RING, CF::NS_TOTAL, p.wstream, ring, bars);
            if (p.stamps && blockIdx.x == 0) p.stamps[22] = clock64();
        }
        __syncwarp();
    } else {
        const int wt = tid - 64, NW = THREADS - 64;
        float *s1_s = reinterpret_cast<float *>(ra + A_S1);
        {
            const float *w = cst + C_WSIG1, *bb = cst + C_BSIG1;
            const int npair = (T1 + 1) >> 1;  // two consecutive steps per thread: one round for chunk_len 100
            for (int i = wt; i < C * npair; i += NW) {
                const int c = i / npair, t = (i - c * npair) * 2;
                const float *xp = sig_s + c * T + t;
                float xv[KW_SIG1 + 1];
#pragma unroll
                for (int j = 0; j <= KW_SIG1; ++j) xv[j] = xp[min(j, T - 1 - t)];
                float4 a = *reinterpret_cast<const float4 *>(bb), b2 = a;
#pragma unroll
                for (int j = 0; j < KW_SIG1; ++j) {
                    const float4 wv = *reinterpret_cast<const float4 *>(w + 4 * j);
                    a.x = fmaf(wv.x, xv[j], a.x);
                    a.y = fmaf(wv.y, xv[j], a.y);
                    a.z = fmaf(wv.z, xv[j], a.z);
                    a.w = fmaf(wv.w, xv[j], a.w);
                    b2.x = fmaf(wv.x, xv[j + 1], b2.x);
                    b2.y = fmaf(wv.y, xv[j + 1], b2.y);
                    b2.z = fmaf(wv.z, xv[j + 1], b2.z);
                    b2.w = fmaf(wv.w, xv[j + 1], b2.w);
                }
                float *dst = s1_s + (size_t)(c * T1 + t) * 4;
                *reinterpret_cast<float4 *>(dst) =
                    make_float4(swishf_fast(a.x), swishf_fast(a.y), swishf_fast(a.z), swishf_fast(a.w));
                if (t + 1 < T1)
                    *reinterpret_cast<float4 *>(dst + 4) =
                        make_float4(swishf_fast(b2.x), swishf_fast(b2.y), swishf_fast(b2.z), swishf_fast(b2.w));
            }
        }
        if (p.stamps && blockIdx.x == 0 && tid == 64) p.stamps[21] = clock64();
        nbar_sync(1, NW);
        // the signal tiles reuse the gather sums' space; without gather sums (direct gather) that space holds
        // the sequence tiles, which the tensor core must have finished reading
        if (!two_stage) mbar_wait(&bars->seq_done, 0);
        {
            // register tile: 4 consecutive output steps x 8 channels per thread - each weight pair is loaded once
            // for four steps and the 16 accumulator chains are independent (only six warps, 1.5 per scheduler,
            // run this phase: with one step per thread it was bound by the latency of its own chain).  For
            // chunk_len 100 the 4 x 23 x 2 items fill the 192 threads once.
            const float *w = cst + C_WSIG2, *bb = cst + C_BSIG2;
            uint8_t *xs = ra + A_XS;
            const int nblk = (T2 + 3) >> 2;
            for (int i = wt; i < C * nblk * 2; i += NW) {
                const int half = i & 1, cb = i >> 1;
                const int c = cb / nblk, t0 = (cb - c * nblk) * 4;
                float2 a2[4][4];
#pragma unroll
                for (int st = 0; st < 4; ++st)
#pragma unroll
                    for (int o = 0; o < 4; ++o) a2[st][o] = make_float2(bb[half * 8 + 2 * o], bb[half * 8 + 2 * o + 1]);
                float x[8][4];  // sig_conv1 rows t0 .. t0 + 7 (rows past the end only feed steps that are dropped)
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float4 xv =
                        *reinterpret_cast<const float4 *>(s1_s + (size_t)(c * T1 + min(t0 + r, T1 - 1)) * 4);
                    x[r][0] = xv.x;
                    x[r][1] = xv.y;
                    x[r][2] = xv.z;
                    x[r][3] = xv.w;
                }
#pragma unroll
                for (int j = 0; j < KW_SIG2; ++j) {
#pragma unroll
                    for (int ci = 0; ci < 4; ++ci) {
                        const float4 *wp = reinterpret_cast<const float4 *>(w + (j * 4 + ci) * 16 + half * 8);
                        const float4 wa = wp[0], wb = wp[1];
#pragma unroll
                        for (int st = 0; st < 4; ++st) {
                            a2[st][0] = ffma2(make_float2(wa.x, wa.y), x[st + j][ci], a2[st][0]);
                            a2[st][1] = ffma2(make_float2(wa.z, wa.w), x[st + j][ci], a2[st][1]);
                            a2[st][2] = ffma2(make_float2(wb.x, wb.y), x[st + j][ci], a2[st][2]);
                            a2[st][3] = ffma2(make_float2(wb.z, wb.w), x[st + j][ci], a2[st][3]);
                        }
                    }
                }
#pragma unroll
                for (int st = 0; st < 4; ++st) {
                    const int t = t0 + st;
                    if (t >= T2) continue;
                    float acc[8];
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        acc[2 * o] = swishf_fast(a2[st][o].x);
                        acc[2 * o + 1] = swishf_fast(a2[st][o].y);
                    }
                    const int r = t % 3, u = t / 3;
                    uint8_t *t_hi = xs + (2 * r) * XT_BYTES;
                    store_chunk8<MODE>(t_hi, t_hi + XT_BYTES, half * LBO_A + (c * U + u) * 16, acc);
                }
            }
        }
        if (p.stamps && blockIdx.x == 0 && tid == 64) p.stamps[23] = clock64();
        fence_async_smem();
    }
    __syncthreads();
    MG_STAMP(4);

    // ---- M2: sig_conv3; E1: both accumulators -> bias + swish -> cat tile (hi / lo) -------------------------
    if (tid == 0) {
        issue_conv3<MODE, KW_SIG3, CF::NS_SEQ, false>(smem_addr(ra + A_XS), tmem + 128, p.wstream, ring, bars,
                                                      p.stamps && blockIdx.x == 0 ? p.stamps + 16 : nullptr);
        if (p.stamps && blockIdx.x == 0) p.stamps[16 + 8] = clock64();
        umma_commit(&bars->conv_done);
    } else if (tid == 32) {
        produce_until(p_next, CF::NS_SEQ + CF::NS_SIG + RING, CF::NS_TOTAL, p.wstream, ring, bars);
    }
    __syncwarp();
    mbar_wait(&bars->conv_done, 0);
    tc_fence_after();
    MG_STAMP(5);
    const int q = warp & 3;           // TMEM lane quarter of this warp = chunk q of the CTA
    const int wh = warp >> 2;         // which half of the columns this warp drains
    const int row = q * U + lane;     // tile row = (chunk q, step lane)
    const bool row_ok = q < C;
    bool overflow = false;
    {
        const int trk = wh;  // 0: sig_conv3 (TMEM columns 128..255) -> channels 0..63; 1: seq_conv2 -> 64..127
        const uint32_t tb = tmem + ((uint32_t)(q * 32) << 16) + (trk ? 0 : 128);
        const float *bias = cst + (trk ? C_BSEQ2 : C_BSIG3);
        const float inv = cst[C_SCALE + (trk ? 0 : 1)];
        uint8_t *cat_hi = ra, *cat_lo = ra + CAT_HALF;
        // a loop over the eight 8-channel groups of this warp's 64 channels: the body (two small TMEM loads,
        // eight swish, one K chunk stored) is short enough to stay in the instruction cache after its first
        // pass - unrolled, the epilogue's code was fetched cold once per CTA and that fetch was a third of it
#pragma unroll 1
        for (int g = 0; g < 8; ++g) {
            uint32_t r0[8], r1[8];
            tmem_ld8_issue(tb + 8 * g, r0);
            if (MODE == 0) tmem_ld8_issue(tb + 64 + 8 * g, r1);
            tmem_ld_wait();
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float v = __uint_as_float(r0[e]);
                if (MODE == 0) v += __uint_as_float(r1[e]);
                o[e] = swishf_fast(fmaf(v, inv, bias[8 * g + e]));
                if (MODE == 0 && row_ok && lane < T3 && !(fabsf(o[e]) < 65504.f)) overflow = true;
            }
            const int kc = trk * 8 + g;
            store_chunk8<MODE>(cat_hi, cat_lo, kc * LBO_A + row * 16, o);
            if (p.dbg_cat && row_ok && lane < T3) {
#pragma unroll
                for (int e = 0; e < 8; ++e) p.dbg_cat[((size_t)(chunk0 + q) * 128 + kc * 8 + e) * T3 + lane] = o[e];
            }
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    MG_STAMP(6);
    // ---- M3: merge_conv1 (K = 5 taps x 128 channels), tap = descriptor shift; E2 -> m tiles -----------------
    if (tid == 0) {
        constexpr int S0 = CF::NS_SEQ + CF::NS_SIG;
        constexpr uint32_t id_main = idesc_h(128, CF::NB, MODE), id_corr = idesc_h(128, 64, MODE);
        const uint64_t da = desc_ns(smem_addr(ra), LBO_A), db = desc_ns(smem_addr(ring), CF::B_LBO);
#pragma unroll
        for (int st = 0; st < CF::NS_MRG; ++st) {  // unrolled: slots, parities and descriptor offsets are immediates
            const int s = S0 + st, slot = s & (RING - 1);
            const int kb = st / KW_MRG, tap = st - kb * KW_MRG;
            mbar_wait(&bars->w_full[slot], (s >> 2) & 1);
            tc_fence_after();
#pragma unroll
            for (int k16 = 0; k16 < CF::MRG_KC / 2; ++k16) {
                const uint64_t a = desc_at(da, (kb * CF::MRG_KC + 2 * k16) * LBO_A + tap * 16);
                const uint64_t b = desc_at(db, slot * STAGE_BYTES + k16 * 2 * CF::B_LBO);
                if (st == 0 && k16 == 0)
                    mma_hc<false>(tmem, a, b, id_main);
                else
                    mma_hc<true>(tmem, a, b, id_main);
                if (MODE == 0) mma_hc<true>(tmem + 64, desc_at(a, CAT_HALF), b, id_corr);
            }
            umma_commit(&bars->w_empty[slot]);
        }
        umma_commit(&bars->mrg_done);
    } else if (tid == 32) {
        produce_until(p_next, CF::NS_SEQ + CF::NS_SIG + CF::NS_MRG + RING, CF::NS_TOTAL, p.wstream, ring, bars);
    }
    __syncwarp();
    mbar_wait(&bars->mrg_done, 0);
    tc_fence_after();
    MG_STAMP(7);
    {
        const uint32_t tb = tmem + ((uint32_t)(q * 32) << 16) + 32 * wh;  // channels 32 wh .. 32 wh + 31
        const float *bias = cst + C_BMRG + 32 * wh;
        const float inv = cst[C_SCALE + 2];
        uint8_t *m_hi = ra, *m_lo = ra + M_HALF;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {  // a short loop body, like epilogue 1
            uint32_t r0[8], r1[8];
            tmem_ld8_issue(tb + 8 * j, r0);
            if (MODE == 0) tmem_ld8_issue(tb + 64 + 8 * j, r1);
            tmem_ld_wait();
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float v = __uint_as_float(r0[e]);
                if (MODE == 0) v += __uint_as_float(r1[e]);
                o[e] = swishf_fast(fmaf(v, inv, bias[8 * j + e]));
                if (MODE == 0 && row_ok && lane < TM && !(fabsf(o[e]) < 65504.f)) overflow = true;
            }
            const int kc = 4 * wh + j;
            store_chunk8<MODE>(m_hi, m_lo, kc * LBO_A + row * 16, o);
            if (p.dbg_m && row_ok && lane < TM) {
#pragma unroll
                for (int e = 0; e < 8; ++e) p.dbg_m[((size_t)(chunk0 + q) * SIZE + kc * 8 + e) * TM + lane] = o[e];
            }
        }
    }
    if (MODE == 0 && overflow && p.flags) atomicOr(p.flags, 1);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    MG_STAMP(8);
    // ---- M4: LSTM1 input projection (N = 256 gate rows, K = 64); E3 -> xp_s[t][chunk][256] ------------------
    if (tid == 0) {
        constexpr int S0 = CF::NS_SEQ + CF::NS_SIG + CF::NS_MRG;
        constexpr uint32_t id_x = idesc_h(128, 256, MODE);
        const uint64_t da = desc_ns(smem_addr(ra), LBO_A), db = desc_ns(smem_addr(ring), 4096);
#pragma unroll
        for (int st = 0; st < CF::NS_XP; ++st) {
            const int s = S0 + st, slot = s & (RING - 1);
            mbar_wait(&bars->w_full[slot], (s >> 2) & 1);
            tc_fence_after();
            const uint64_t b = desc_at(db, slot * STAGE_BYTES);
            if (MODE == 0) {
                const int k16 = st >> 1, lo_stage = st & 1;
                const uint64_t m_hi = desc_at(da, 2 * k16 * LBO_A), m_lo = desc_at(m_hi, M_HALF);
                if (!lo_stage) {  // W_hi: m_hi * W_hi, m_lo * W_hi
                    if (st == 0)
                        mma_hc<false>(tmem, m_hi, b, id_x);
                    else
                        mma_hc<true>(tmem, m_hi, b, id_x);
                    mma_hc<true>(tmem, m_lo, b, id_x);
                } else {  // W_lo: m_hi * W_lo
                    mma_hc<true>(tmem, m_hi, b, id_x);
                }
            } else {
                if (st == 0)
                    mma_hc<false>(tmem, desc_at(da, 2 * st * LBO_A), b, id_x);
                else
                    mma_hc<true>(tmem, desc_at(da, 2 * st * LBO_A), b, id_x);
            }
            umma_commit(&bars->w_empty[slot]);
        }
        umma_commit(&bars->xp_done);
    } else if (tid == 32) {
        produce_until(p_next, CF::NS_TOTAL, CF::NS_TOTAL, p.wstream, ring, bars);
    }
    __syncwarp();
    mbar_wait(&bars->xp_done, 0);  // every MMA has completed: ring, tiles and region A are dead
    tc_fence_after();
    MG_STAMP(9);
    {
        const float inv = cst[C_SCALE + 3];
        const float *b1 = cst + C_B1 + 128 * wh;
        float *dst = xp_s + lane * XPS + q * XCS + 128 * wh;
#pragma unroll 1
        for (int cq = 0; cq < 4; ++cq) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 128 * wh + 32 * cq, v);
            if (lane < TM) {
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float4 bb = *reinterpret_cast<const float4 *>(b1 + 32 * cq + 4 * c4);
                    *reinterpret_cast<float4 *>(dst + 32 * cq + 4 * c4) =
                        make_float4(fmaf(v[4 * c4], inv, bb.x), fmaf(v[4 * c4 + 1], inv, bb.y),
                                    fmaf(v[4 * c4 + 2], inv, bb.z), fmaf(v[4 * c4 + 3], inv, bb.w));
                }
                if (p.dbg_xp && row_ok) {
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        p.dbg_xp[((size_t)(chunk0 + q) * 256 + 128 * wh + 32 * cq + e) * TM + lane] =
                            fmaf(v[e], inv, b1[32 * cq + e]);
                }
            }
        }
    }
    for (int i = tid; i < 2 * 8 * 36; i += THREADS) reinterpret_cast<uint32_t *>(g_s)[i] = 0u;  // h(-1) = 0
    tc_fence_before();
    __syncthreads();
    if (warp == 2)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));

    MG_STAMP(10);
    // ---- R: LSTM1 recurrence, W_hh in registers ---------------------------------------------------------------
    // The 256 x 64 mat-vec of every step runs on the tensor cores through the warp-level mma.sync path
    // (HMMA.16816, operands in registers): measured 8.6 cycles per instruction and SM sub-partition, i.e. 7.7x
    // the MAC rate of FFMA2, and the FMA pipe stays free for the co-resident CTA.
    //   M = gate rows: warp w owns hidden units 8w .. 8w+7; its first 16-row tile holds their input and forget
    //       gate rows, the second one their cell and output gate rows
    //   N = 8 columns = 4 chunks x {h_hi, h_lo}: column 2c is the fp16 rounding of chunk c's h, column 2c+1 the
    //       fp16 rounding of the remainder, so that c0 + c1 of an accumulator fragment is W . (h_hi + h_lo)
    //   K = 64 hidden units = 4 k-tiles; W_hh is held as fp16 hi + lo A fragments (64 registers, pre-scaled by a
    //       power of two), both accumulate into fp32: (W_hi + W_lo) . (h_hi + h_lo), all four products
    // Thread (g = lane >> 2, t = lane & 3) ends up with the four gate pre-activations of cell (unit 8w+g,
    // chunk t): the cell update stays in registers, one barrier per step (h operand tile double buffered).
    const int g8 = lane >> 2, t4 = lane & 3, u = 8 * warp + g8;
    uint32_t wa[16][4];  // [(m-tile * 4 + k-tile) * 2 + part][a0..a3]
#pragma unroll
    for (int q4 = 0; q4 < 16; ++q4) {
        const uint4 v = reinterpret_cast<const uint4 *>(p.whh4)[q4 * 256 + tid];
        wa[q4][0] = v.x;
        wa[q4][1] = v.y;
        wa[q4][2] = v.z;
        wa[q4][3] = v.w;
    }
    MG_STAMP(11);
    constexpr int HBP = 36;                                    // words per column of the h operand tile (+4: banks)
    uint32_t *hb_s = reinterpret_cast<uint32_t *>(g_s);        // [2 buffers][8 columns][HBP] fp16 pairs along k
    const float inv_hh = cst[C_SCALE + 4];
    const float *xcell = xp_s + t4 * XCS + u;                  // + step * XPS + 64 * gate
    float cstate = 0.f, hval = 0.f;
#pragma unroll 1
    for (int t = 0; t < TM; ++t) {
        const uint32_t *hb = hb_s + (t & 1) * 8 * HBP + g8 * HBP + t4;   // column g8 of h(t - 1)
        __half *hn = reinterpret_cast<__half *>(hb_s + ((t + 1) & 1) * 8 * HBP);
        const float *xt = xcell + t * XPS;
        const float x0 = xt[0], x1 = xt[64], x2 = xt[128], x3 = xt[192];
        float acc[4][4];  // [m-tile * 2 + part of W][c0..c3]: four independent accumulation chains
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
            const uint32_t b0 = hb[8 * kt], b1 = hb[8 * kt + 4];
#pragma unroll
            for (int mp = 0; mp < 4; ++mp) {
                const uint32_t(&a)[4] = wa[((mp >> 1) * 4 + kt) * 2 + (mp & 1)];
                asm volatile(
                    "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
                    "{%0,%1,%2,%3};"
                    : "+f"(acc[mp][0]), "+f"(acc[mp][1]), "+f"(acc[mp][2]), "+f"(acc[mp][3])
                    : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
            }
        }
        // rows g8 (c0, c1) / g8 + 8 (c2, c3); columns 2 t4 (h_hi) and 2 t4 + 1 (h_lo); W_hi and W_lo parts
        const float pi = (acc[0][0] + acc[0][1]) + (acc[1][0] + acc[1][1]);
        const float pf = (acc[0][2] + acc[0][3]) + (acc[1][2] + acc[1][3]);
        const float pg = (acc[2][0] + acc[2][1]) + (acc[3][0] + acc[3][1]);
        const float po = (acc[2][2] + acc[2][3]) + (acc[3][2] + acc[3][3]);
        const float ig = sigmoidf_fast(fmaf(pi, inv_hh, x0)), fg = sigmoidf_fast(fmaf(pf, inv_hh, x1));
        const float gg = tanhf_fast(fmaf(pg, inv_hh, x2)), og = sigmoidf_fast(fmaf(po, inv_hh, x3));
        cstate = fg * cstate + ig * gg;
        hval = og * tanhf_fast(cstate);
        const __half hh = __float2half_rn(hval);
        hn[(2 * t4) * (2 * HBP) + u] = hh;
        hn[(2 * t4 + 1) * (2 * HBP) + u] = __float2half_rn(hval - __half2float(hh));
        __syncthreads();
    }
    MG_STAMP(12);
    // ---- LSTM2: only the first step of the reversed pass is consumed (ConvLSTM_w_ref.py:53-54) ----------------
    // same shape as a recurrence step (zero initial state: gates = W_ih2 x + b2, the forget gate is unused):
    // W_ih2 arrives as the same A fragments, straight from L2 into the registers W_hh just left (all 16 loads in
    // flight at once), x = swish(h1) goes through the operand tile, the cell stays in registers
#pragma unroll
    for (int q4 = 0; q4 < 16; ++q4) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p.wih2T) + q4 * 256 + tid);
        wa[q4][0] = v.x;
        wa[q4][1] = v.y;
        wa[q4][2] = v.z;
        wa[q4][3] = v.w;
    }
    {
        const float x = swishf(hval);
        const __half xh = __float2half_rn(x);
        __half *hn = reinterpret_cast<__half *>(hb_s);
        hn[(2 * t4) * (2 * HBP) + u] = xh;
        hn[(2 * t4 + 1) * (2 * HBP) + u] = __float2half_rn(x - __half2float(xh));
    }
    const float bi = p.b2[u], bg = p.b2[128 + u], bo = p.b2[192 + u];
    __syncthreads();
    {
        const uint32_t *hb = hb_s + g8 * HBP + t4;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
            const uint32_t b0 = hb[8 * kt], b1 = hb[8 * kt + 4];
#pragma unroll
            for (int mp = 0; mp < 4; ++mp) {
                const uint32_t(&a)[4] = wa[((mp >> 1) * 4 + kt) * 2 + (mp & 1)];
                asm volatile(
                    "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
                    "{%0,%1,%2,%3};"
                    : "+f"(acc[mp][0]), "+f"(acc[mp][1]), "+f"(acc[mp][2]), "+f"(acc[mp][3])
                    : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
            }
        }
        const float inv2 = cst[C_SCALE + 6];
        const float pi = (acc[0][0] + acc[0][1]) + (acc[1][0] + acc[1][1]);
        const float pg = (acc[2][0] + acc[2][1]) + (acc[3][0] + acc[3][1]);
        const float po = (acc[2][2] + acc[2][3]) + (acc[3][2] + acc[3][3]);
        const float c2 = sigmoidf_acc(fmaf(pi, inv2, bi)) * tanhf(fmaf(pg, inv2, bg));
        const float h2 = sigmoidf_acc(fmaf(po, inv2, bo)) * tanhf(c2);
        y_s[t4 * SIZE + u] = swishf(h2);
    }
    __syncthreads();
    MG_STAMP(13);
    pdl_wait();  // order our only global writes after the previous grid in the stream (output buffer reuse)
    MG_STAMP(14);
    if (warp < C) {
        for (int o = 0; o < p.num_out; ++o) {
            float part = p.fcw[o * SIZE + lane] * y_s[warp * SIZE + lane] +
                         p.fcw[o * SIZE + lane + 32] * y_s[warp * SIZE + lane + 32];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
            const float val = part + p.fcb[o];  // every lane holds the sum
            const size_t idx = (size_t)(chunk0 + warp) * p.num_out + o;
            if (p.logits != nullptr && lane == 0) p.logits[idx] = val;
            if (p.n_peers > 0 && !p.ship_cta) {
                // the exchange step of the path, fused into the producing kernel: lane r stores into rank
                // r's buffer through its peer mapping (NVLink / NVSwitch), or one multimem store reaches
                // every rank through the switch
                if (p.mc_base != nullptr) {
                    if (lane == 0)
                        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p.mc_base + p.peer_off + idx),
                                     "f"(val)
                                     : "memory");
                } else if (lane < p.n_peers) {
                    p.peers[lane][p.peer_off + idx] = val;
                }
            }
        }
    }
    if (p.n_peers > 0 && !p.ship_cta && p.flag_off >= 0) {  // arrival counter per destination: complete when it reaches gridDim.x
        __threadfence_system();
        __syncthreads();
        if (tid < p.n_peers)
            atomicAdd_system(reinterpret_cast<unsigned int *>(p.peers[tid]) + p.flag_off, 1u);
    }
    if (p.trace && tid == 0) {
        long long t1;
        unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        long long *rec = p.trace + (size_t)blockIdx.x * 20;
        for (int i = 0; i < 15; ++i) rec[4 + i] = trace_st[i + 1] - trace_st[i];
        rec[0] = smid;
        rec[1] = trace_t0;
        rec[2] = t1;
        rec[3] = clock64() - trace_c0;
    }
#undef MG_STAMP
}


// =========================================================================================================
// Conv_w_ref (models/Conv_w_ref.py:44-62) in one kernel, same machinery: fp16 hi/lo split operands on
// tcgen05, K-major unswizzled tiles, taps as descriptor shifts, residue tiles for the strided layers.
// =========================================================================================================
//   sig : conv(1->4,k11) conv(4->16,k11) [FFMA]        -> residue-3 tiles -> conv(16->64,k9,s3) [MMA]
//   seq : conv(36->16,k11) on the one-hot = gather-add  -> q1 tile (row = chunk*96 + t)
//         conv(16->32,k11) [MMA, 3 M tiles, tap = shift] -> residue-3 tiles -> conv(32->64,k9,s3) [MMA]
//   merge: cat -> conv(128->64,k5) -> conv(64->64,k5) -> residue-2 tiles -> conv(64->64,k3,s2)
//          -> residue-2 tiles -> conv(64->64,k3,s2) -> fc over the flattened [64][m4] block
// Four chunks per CTA, two CTAs per SM.  Seven MMA layers, 288 MMAs and 59 weight stages (472 KB) per CTA.
namespace cw {
constexpr int KW1 = 11, KW3 = 9, KWM = 5, KW2 = 3;
// Geometry: G chunks per CTA with U tile rows per chunk (G * U = 128 = one M tile) and U1 = 3 U rows per
// chunk in the q1 tile (G * U1 = 384 = three M tiles): chunk_len <= 100 -> G = 4, U = 32; chunk_len <= 200
// (BASELINE config 3) -> G = 2, U = 64.  Every tile has the same size in both cases.
constexpr int Q1_ROWS = 384;                 // 3 M tiles
constexpr int Q1_RP = Q1_ROWS + 16;          // rows stored per K chunk (tap shifts up to 10)
constexpr int Q1_LBO = Q1_RP * 16;           // 6400
constexpr int Q1_HALF = 2 * Q1_LBO;          // 12800: 16 channels, hi (or lo)
constexpr int RP3 = MROWS + 4;               // residue-3 tiles: shifts up to 2
constexpr int LBO3 = RP3 * 16;               // 2112
constexpr int XS_T = 2 * LBO3;               // sig residue tile (16 ch): 4224
constexpr int XQ_T = 4 * LBO3;               // seq residue tile (32 ch): 8448
constexpr int T64 = 8 * LBO_A;               // a 64-channel tile, hi (or lo): 17408
// constants blob (floats)
constexpr int C_WS1 = 0;                     // [j][co]       44 (+4 pad)
constexpr int C_BS1 = 48;                    //                4
constexpr int C_WS2 = 52;                    // [j][ci][co]  704
constexpr int C_BS2 = 756;                   //               16
constexpr int C_BQ1 = 772;                   //               16
constexpr int C_BQ2 = 788;                   //               32
constexpr int C_BS3 = 820;                   //               64
constexpr int C_BQ3 = 884;                   //               64
constexpr int C_BM = 948;                    // m1..m4       256
constexpr int C_SC = 1204;                   // inverse scales: q2, s3, q3, m1, m2, m3, m4 (+1 pad)
constexpr int CONST_FLOATS = 1216;
constexpr int CONST_BYTES = CONST_FLOATS * 4;  // 4864
// weight stages
constexpr int NS_Q2 = (KW1 + 3) / 4;         // 3: four taps ([64 n][16 k] = 2 KB each) per stage
constexpr int NS_S3 = (KW3 + 1) / 2;         // 5: two taps ([128][16] = 4 KB) per stage
constexpr int NS_Q3 = KW3;                   // 9: one tap ([128][32] = 8 KB) per stage
constexpr int NS_M1 = KWM * 4;               // 20: (channel block of 32, tap)
constexpr int NS_M2 = KWM * 2;               // 10: (tap, half of the 64 channels)
constexpr int NS_M3 = KW2 * 2, NS_M4 = KW2 * 2;
constexpr int NS_TOTAL = NS_Q2 + NS_S3 + NS_Q3 + NS_M1 + NS_M2 + NS_M3 + NS_M4;  // 59
// shared memory
constexpr int OFF_CONST = 256;
constexpr int OFF_RING = 5120;
constexpr int OFF_A = OFF_RING + RING * STAGE_BYTES;   // 37888
constexpr int A_BYTES = 77824;
constexpr int SMEM_BYTES = OFF_A + A_BYTES;            // 115712: exactly half an SM
constexpr int A_STG = 0, STG_BYTES = 3072;             // sig 1600, sidx 800, seq <= 256, map <= 384, len 16
constexpr int STG_SIDX = 1600, STG_SEQ = 2400, STG_MAP = 2656, STG_LEN = 3040;
constexpr int A_GS = 3072, GS_CAP = 35200;             // per-base sums of ONE chunk: <= 40 bases x 11 taps x 80 B
constexpr int A_Q1 = 38400;                            // q1 tile hi + lo: 25600
constexpr int A_XS = 0;                                // signal residue tiles (3 x {hi, lo}): 25344, written
                                                       // by sig_conv2 over the dead staging / sums
constexpr int A_S1 = 25600, S1_CAP = 6144;             // sig_conv1 output, inside the dead sums, clear of A_XS
constexpr int A_XQ = A_BYTES - 6 * XQ_T;               // 27136: sequence residue tiles, over the dead s1 / q1
constexpr int MAX_T = 200;
static_assert(STG_LEN + 16 <= STG_BYTES && A_STG + STG_BYTES <= A_GS, "staging does not fit");
static_assert(A_GS + GS_CAP <= A_Q1 && A_Q1 + 2 * Q1_HALF <= A_BYTES, "front-phase tiles overlap");
static_assert(A_XS + 6 * XS_T <= A_S1 && A_S1 + S1_CAP <= A_Q1 && A_XS + 6 * XS_T <= A_XQ, "signal tiles overlap");
static_assert(2 * CAT_HALF <= A_BYTES && 4 * T64 <= A_BYTES, "tiles do not fit");
static_assert(SMEM_BYTES <= 115712, "two CTAs per SM need <= 113 KB each");

struct Bars {
    uint64_t w_full[RING], w_empty[RING], front, done[7];
    uint32_t tmem_base;
};

struct Params {
    const float *sigs;
    const int8_t *seqs;
    const int16_t *maps;
    const int16_t *lens;
    int seq_width, map_width, B, T, kmer_len, num_out, fc_in;
    const float *consts, *gtab;
    int gtab_bytes;
    const uint8_t *wstream;
    const float *fcw, *fcb;
    float *logits;
    float *dbg_cat, *dbg_m4;
    int *flags;
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void load_stage_c(int s, const uint8_t *wstream, uint8_t *ring, Bars *bars) {
    if (s >= NS_TOTAL) return;
    const int slot = s & (RING - 1);
    if (s >= RING) mbar_wait(&bars->w_empty[slot], ((s >> 2) - 1) & 1);
    mbar_expect_tx(&bars->w_full[slot], STAGE_BYTES);
    bulk_g2s(ring + slot * STAGE_BYTES, wstream + (size_t)s * STAGE_BYTES, STAGE_BYTES, &bars->w_full[slot]);
}
// stage `s` (a compile-time number in the unrolled issue loops) has landed; returns its byte offset in the ring
__device__ __forceinline__ uint32_t wait_stage(int s, Bars *bars) {
    const int slot = s & (RING - 1);
    mbar_wait(&bars->w_full[slot], (s >> 2) & 1);
    tc_fence_after();
    return (uint32_t)(slot * STAGE_BYTES);
}

// 64-channel merge-type layer: `taps` taps, tap j reads tile (tile_of[j]) shifted by shift_of[j] rows;
// weights: two stages (K = 32 each) per tap.  A tiles are [8 K chunks][RP rows] hi at `a`, lo at a + T64.
template <int TAPS, int S0>
__device__ __forceinline__ void issue_merge64(const uint32_t (&a_tile)[TAPS], const int (&shift)[TAPS], uint32_t d,
                                              uint8_t *ring, Bars *bars) {
    constexpr uint32_t id_main = idesc_h(128, 128, 0), id_corr = idesc_h(128, 64, 0);
    const uint64_t db = desc_ns(smem_addr(ring), 2048);
#pragma unroll
    for (int j = 0; j < TAPS; ++j) {
        const uint64_t da = desc_ns(a_tile[j] + shift[j] * 16, LBO_A);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int s = S0 + 2 * j + h;
            const uint32_t b0 = wait_stage(s, bars);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const uint64_t a = desc_at(da, (4 * h + 2 * kk) * LBO_A);
                const uint64_t b = desc_at(db, b0 + kk * 2 * 2048);
                if (j == 0 && h == 0 && kk == 0)
                    mma_hc<false>(d, a, b, id_main);
                else
                    mma_hc<true>(d, a, b, id_main);
                mma_hc<true>(d + 64, desc_at(a, T64), b, id_corr);
            }
            umma_commit(&bars->w_empty[s & (RING - 1)]);
        }
    }
}

// accumulator (main + correction columns) -> bias + swish for the 32 channels [32 wh, 32 wh + 32) of row `lane`
__device__ __forceinline__ void drain32(uint32_t tb, int wh, const float *bias, float inv, float (&o)[32]) {
    float v2[32];
    tmem_ld32(tb + 32 * wh, o);
    tmem_ld32(tb + 64 + 32 * wh, v2);
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = swishf_fast(fmaf(o[i] + v2[i], inv, bias[32 * wh + i]));
}

__global__ void __launch_bounds__(THREADS, 2) conv_mega_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(128) uint8_t sm[];
    Bars *bars = reinterpret_cast<Bars *>(sm);
    float *cst = reinterpret_cast<float *>(sm + OFF_CONST);
    uint8_t *ring = sm + OFF_RING;
    uint8_t *ra = sm + OFF_A;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = p.T;
    const int GC = T <= 100 ? 4 : 2;          // chunks per CTA
    const int UR = MROWS / GC;                // tile rows per chunk (32 | 64)
    const int U1 = 3 * UR;                    // q1 tile rows per chunk (96 | 192)
    const int chunk0 = blockIdx.x * GC;
    const int C = min(GC, p.B - chunk0);
    const int S1 = T - (KW1 - 1), S2 = S1 - (KW1 - 1), S3 = (S2 - KW3) / 3 + 1;
    const int M1 = S3 - (KWM - 1), M2 = M1 - (KWM - 1), M3 = (M2 - KW2) / 2 + 1, M4 = (M3 - KW2) / 2 + 1;
    const int seq_width = p.seq_width, map_width = p.map_width, K = p.kmer_len;

    pdl_launch_dependents();
    if (tid == 0) {
        for (int i = 0; i < RING; ++i) {
            mbar_init(&bars->w_full[i], 1);
            mbar_init(&bars->w_empty[i], 1);
        }
        mbar_init(&bars->front, 1);
        for (int i = 0; i < 7; ++i) mbar_init(&bars->done[i], 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_addr(&bars->tmem_base)),
                     "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    if (tid == 0) {  // constants, and the gather table into the (still idle) weight ring
        mbar_expect_tx(&bars->front, CONST_BYTES + p.gtab_bytes);
        bulk_g2s(cst, p.consts, CONST_BYTES, &bars->front);
        bulk_g2s(ring, p.gtab, p.gtab_bytes, &bars->front);
    }
    // ---- stage the compact inputs, move-table expansion ------------------------------------------------------
    float *sig_s = reinterpret_cast<float *>(ra + A_STG);
    int16_t *sidx_s = reinterpret_cast<int16_t *>(ra + A_STG + STG_SIDX);
    int8_t *seq_s = reinterpret_cast<int8_t *>(ra + A_STG + STG_SEQ);
    int16_t *map_s = reinterpret_cast<int16_t *>(ra + A_STG + STG_MAP);
    int *len_s = reinterpret_cast<int *>(ra + A_STG + STG_LEN);
    for (int i = tid; i < C * T; i += THREADS) {
        sig_s[i] = p.sigs[(size_t)chunk0 * T + i];
        sidx_s[i] = -1;
    }
    if (tid < C) {
        int L = p.lens[chunk0 + tid];
        L = max(0, min(L, min(map_width - 1, seq_width - K + 1)));
        len_s[tid] = L;
    }
    __syncthreads();
    for (int i = tid; i < C * seq_width; i += THREADS) {
        const int c = i / seq_width, s = i - c * seq_width;
        seq_s[i] = s < len_s[c] + K - 1 ? p.seqs[(size_t)(chunk0 + c) * seq_width + s] : (int8_t)-1;
    }
    for (int i = tid; i < C * map_width; i += THREADS) {
        const int c = i / map_width, s = i - c * map_width;
        map_s[i] = s <= len_s[c] ? p.maps[(size_t)(chunk0 + c) * map_width + s] : (int16_t)0;
    }
    __syncthreads();
    for (int i = tid; i < C * (map_width - 1); i += THREADS) {
        const int c = i / (map_width - 1), s = i - c * (map_width - 1);
        if (s < len_s[c]) {
            const int st = max((int)map_s[c * map_width + s], 0);
            const int en = min((int)map_s[c * map_width + s + 1], T);
            for (int t = st; t < en; ++t) sidx_s[c * T + t] = (int16_t)s;
        }
    }
    mbar_wait(&bars->front, 0);
    __syncthreads();

    // ---- seq_conv1 (36 -> 16, k11) on the virtual one-hot: two-stage gather-add, one chunk at a time ---------
    {
        const float *gt = reinterpret_cast<const float *>(ring);
        const int zero_off = KW1 * K * 4 * GROW;
        float *gs = reinterpret_cast<float *>(ra + A_GS);
        uint8_t *q_hi = ra + A_Q1, *q_lo = q_hi + Q1_HALF;
        const int LM = map_width - 1;
        for (int c = 0; c < C; ++c) {
            for (int i = tid; i < LM * KW1; i += THREADS) {
                const int sb = i / KW1, j = i - sb * KW1;
                if (sb >= len_s[c]) continue;
                float2 a[8];
#pragma unroll
                for (int o = 0; o < 8; ++o) a[o] = make_float2(0.f, 0.f);
                const int8_t *sp = seq_s + c * seq_width + sb;
                const int joff = j * K * 4 * GROW;
                for (int pp = 0; pp < K; ++pp) {
                    const int base = sp[pp];
                    const int off = (base >= 0 && base <= 3) ? joff + (pp * 4 + base) * GROW : zero_off;
                    const float4 *wv = reinterpret_cast<const float4 *>(gt + off);
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
            // one thread per (output step, 8-channel half): 11 rows of 8 floats, swish, one 16-byte K chunk
            for (int i = tid; i < S1 * 2; i += THREADS) {
                const int t = i >> 1, half = i & 1;
                float acc[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = cst[C_BQ1 + 8 * half + e];
#pragma unroll
                for (int j = 0; j < KW1; ++j) {
                    const int sb = sidx_s[c * T + t + j];
                    if (sb < 0) continue;
                    const float4 *gv = reinterpret_cast<const float4 *>(gs + (size_t)(sb * KW1 + j) * GROW + 8 * half);
                    const float4 v0 = gv[0], v1 = gv[1];
                    acc[0] += v0.x; acc[1] += v0.y; acc[2] += v0.z; acc[3] += v0.w;
                    acc[4] += v1.x; acc[5] += v1.y; acc[6] += v1.z; acc[7] += v1.w;
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = swishf_fast(acc[e]);
                store_chunk8<0>(q_hi, q_lo, half * Q1_LBO + (c * U1 + t) * 16, acc);
            }
            __syncthreads();
        }
    }
    fence_async_smem();
    __syncthreads();  // the gather table is dead: the ring starts streaming weights

    // stage numbers of the unrolled issue loops (compile time: slots, parities, descriptor offsets are immediates)
    constexpr int S_Q2 = 0, S_S3 = NS_Q2, S_Q3 = S_S3 + NS_S3, S_M1 = S_Q3 + NS_Q3, S_M2 = S_M1 + NS_M1,
                  S_M3 = S_M2 + NS_M2, S_M4 = S_M3 + NS_M3;
    int p_next = RING;  // producer (warp 1 lane 0) position
    // ---- seq_conv2 (16 -> 32, k11, stride 1) on the tensor core, under sig_conv1 / sig_conv2 on warps 1..7 ----
    if (warp == 0) {
        if (lane == 0) {
            for (int s = 0; s < RING; ++s) load_stage_c(s, p.wstream, ring, bars);
            constexpr uint32_t id_main = idesc_h(128, 64, 0), id_corr = idesc_h(128, 32, 0);
            const uint64_t dq = desc_ns(smem_addr(ra + A_Q1), Q1_LBO), db = desc_ns(smem_addr(ring), 1024);
#pragma unroll
            for (int st = 0; st < NS_Q2; ++st) {
                const int s = S_Q2 + st;
                const uint32_t b0 = wait_stage(s, bars);
#pragma unroll
                for (int tp = 0; tp < 4; ++tp) {
                    const int j = st * 4 + tp;
                    if (j < KW1) {
#pragma unroll
                        for (int mt = 0; mt < 3; ++mt) {
                            const uint64_t a = desc_at(dq, (mt * 128 + j) * 16);
                            const uint64_t b = desc_at(db, b0 + tp * 2048);
                            if (j == 0)
                                mma_hc<false>(tmem + mt * 64, a, b, id_main);
                            else
                                mma_hc<true>(tmem + mt * 64, a, b, id_main);
                            mma_hc<true>(tmem + mt * 64 + 32, desc_at(a, Q1_HALF), b, id_corr);
                        }
                    }
                }
                umma_commit(&bars->w_empty[s & (RING - 1)]);
                // self-loading while the other warps compute: RING stages ahead would block on this stage's own MMAs
                if (s >= 1) load_stage_c(s + RING - 1, p.wstream, ring, bars);
            }
            umma_commit(&bars->done[0]);
        }
        __syncwarp();
    } else {
        const int wt = tid - 32, NW = THREADS - 32;
        float *s1_s = reinterpret_cast<float *>(ra + A_S1);
        for (int i = wt; i < C * S1; i += NW) {
            const int c = i / S1, t = i - c * S1;
            const float *x = sig_s + c * T + t;
            float4 a = *reinterpret_cast<const float4 *>(cst + C_BS1);
#pragma unroll
            for (int j = 0; j < KW1; ++j) {
                const float xv = x[j];
                const float4 wv = *reinterpret_cast<const float4 *>(cst + C_WS1 + 4 * j);
                a.x = fmaf(wv.x, xv, a.x);
                a.y = fmaf(wv.y, xv, a.y);
                a.z = fmaf(wv.z, xv, a.z);
                a.w = fmaf(wv.w, xv, a.w);
            }
            *reinterpret_cast<float4 *>(s1_s + (size_t)i * 4) =
                make_float4(swishf_fast(a.x), swishf_fast(a.y), swishf_fast(a.z), swishf_fast(a.w));
        }
        nbar_sync(1, NW);
        uint8_t *xs = ra + A_XS;
        for (int i = wt; i < C * S2 * 2; i += NW) {
            const int half = i & 1, ct = i >> 1;
            const int c = ct / S2, t = ct - c * S2;
            float acc[8];
#pragma unroll
            for (int o = 0; o < 8; ++o) acc[o] = cst[C_BS2 + half * 8 + o];
#pragma unroll
            for (int j = 0; j < KW1; ++j) {
                const float4 xv = *reinterpret_cast<const float4 *>(s1_s + (size_t)(c * S1 + t + j) * 4);
                const float xs4[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) {
                    const float4 *wp = reinterpret_cast<const float4 *>(cst + C_WS2 + (j * 4 + ci) * 16 + half * 8);
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
#pragma unroll
            for (int o = 0; o < 8; ++o) acc[o] = swishf_fast(acc[o]);
            const int r = t % 3, u = t / 3;
            uint8_t *t_hi = xs + (2 * r) * XS_T;
            store_chunk8<0>(t_hi, t_hi + XS_T, half * LBO3 + (c * UR + u) * 16, acc);
        }
        fence_async_smem();
    }
    __syncthreads();
    // M1 loaded stages up to NS_Q2 + RING - 2 (it skips the look-ahead of its first stage)
    p_next = NS_Q2 + RING - 1;
    const int q = warp & 3, wh = warp >> 2;
    const int row = q * 32 + lane;            // this thread's row of every M = 128 tile (TMEM lane)
    const int rc = row / UR, rt = row - rc * UR;  // = (chunk, step) of that row
    const bool chunk_ok = rc < C;
    bool overflow = false;

    // ---- epilogue of seq_conv2: 3 M tiles (rows chunk * 96 + t) -> residue-3 tiles of 32 channels -------------
    if (tid == 32) {  // keep the ring full while everybody drains
        for (; p_next < NS_Q2 + RING; ++p_next) load_stage_c(p_next, p.wstream, ring, bars);
    }
    __syncwarp();
    mbar_wait(&bars->done[0], 0);
    tc_fence_after();
    {
        const float inv = cst[C_SC + 0];
        uint8_t *xq = ra + A_XQ;
#pragma unroll 1
        for (int mt = 0; mt < 3; ++mt) {
            const int r384 = mt * 128 + q * 32 + lane;
            const int c = r384 / U1, t = r384 - c * U1;
            float v[16], v2[16];
            const uint32_t tb = tmem + ((uint32_t)(q * 32) << 16) + mt * 64 + 16 * wh;
            tmem_ld16(tb, v);
            tmem_ld16(tb + 32, v2);
            if (c < C && t < S2) {
                const int r = t % 3, u = t / 3;
                uint8_t *t_hi = xq + (2 * r) * XQ_T;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        o[e] = swishf_fast(fmaf(v[8 * j + e] + v2[8 * j + e], inv, cst[C_BQ2 + 16 * wh + 8 * j + e]));
                        if (!(fabsf(o[e]) < 65504.f)) overflow = true;
                    }
                    store_chunk8<0>(t_hi, t_hi + XQ_T, (2 * wh + j) * LBO3 + (c * UR + u) * 16, o);
                }
            }
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- sig_conv3 (16 -> 64) and seq_conv3 (32 -> 64), k9 stride 3, from the residue tiles --------------------
    if (tid == 0) {
        constexpr uint32_t id_main = idesc_h(128, 128, 0), id_corr = idesc_h(128, 64, 0);
        const uint64_t dxs = desc_ns(smem_addr(ra + A_XS), LBO3), dxq = desc_ns(smem_addr(ra + A_XQ), LBO3),
                       db = desc_ns(smem_addr(ring), 2048);
#pragma unroll
        for (int st = 0; st < NS_S3; ++st) {
            const int s = S_S3 + st;
            const uint32_t b0 = wait_stage(s, bars);
#pragma unroll
            for (int tp = 0; tp < 2; ++tp) {
                const int j = st * 2 + tp;
                if (j < KW3) {
                    const uint64_t a = desc_at(dxs, (2 * (j % 3)) * XS_T + (j / 3) * 16);
                    const uint64_t b = desc_at(db, b0 + tp * 4096);
                    if (j == 0)
                        mma_hc<false>(tmem, a, b, id_main);
                    else
                        mma_hc<true>(tmem, a, b, id_main);
                    mma_hc<true>(tmem + 64, desc_at(a, XS_T), b, id_corr);
                }
            }
            umma_commit(&bars->w_empty[s & (RING - 1)]);
        }
#pragma unroll
        for (int j = 0; j < KW3; ++j) {
            const int s = S_Q3 + j;
            const uint32_t b0 = wait_stage(s, bars);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const uint64_t a = desc_at(dxq, (2 * (j % 3)) * XQ_T + (j / 3) * 16 + kk * 2 * LBO3);
                const uint64_t b = desc_at(db, b0 + kk * 2 * 2048);
                if (j == 0 && kk == 0)
                    mma_hc<false>(tmem + 128, a, b, id_main);
                else
                    mma_hc<true>(tmem + 128, a, b, id_main);
                mma_hc<true>(tmem + 192, desc_at(a, XQ_T), b, id_corr);
            }
            umma_commit(&bars->w_empty[s & (RING - 1)]);
        }
        umma_commit(&bars->done[1]);
    } else if (tid == 32) {
        for (; p_next < NS_Q2 + NS_S3 + NS_Q3 + RING; ++p_next) load_stage_c(p_next, p.wstream, ring, bars);
    }
    __syncwarp();
    mbar_wait(&bars->done[1], 0);
    tc_fence_after();
    {   // cat tile: channels 0..63 = signal track, 64..127 = sequence track; this warp: track wh... two passes
        uint8_t *cat_hi = ra, *cat_lo = ra + CAT_HALF;
#pragma unroll 1
        for (int trk = 0; trk < 2; ++trk) {
            float o[32];
            drain32(tmem + ((uint32_t)(q * 32) << 16) + trk * 128, wh, cst + (trk ? C_BQ3 : C_BS3),
                    cst[C_SC + (trk ? 2 : 1)], o);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float o8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    o8[e] = o[8 * j + e];
                    if (chunk_ok && rt < S3 && !(fabsf(o8[e]) < 65504.f)) overflow = true;
                }
                const int kc = trk * 8 + 4 * wh + j;
                store_chunk8<0>(cat_hi, cat_lo, kc * LBO_A + row * 16, o8);
                if (p.dbg_cat && chunk_ok && rt < S3) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) p.dbg_cat[((size_t)(chunk0 + rc) * 128 + kc * 8 + e) * S3 + rt] = o8[e];
                }
            }
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- merge_conv1 (128 -> 64, k5): taps are shifts of the cat tile ----------------------------------------------
    if (tid == 0) {
        constexpr uint32_t id_main = idesc_h(128, 128, 0), id_corr = idesc_h(128, 64, 0);
        const uint64_t dc = desc_ns(smem_addr(ra), LBO_A), db = desc_ns(smem_addr(ring), 2048);
#pragma unroll
        for (int st = 0; st < NS_M1; ++st) {
            const int s = S_M1 + st;
            const int kb = st / KWM, tap = st - kb * KWM;
            const uint32_t b0 = wait_stage(s, bars);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const uint64_t a = desc_at(dc, (kb * 4 + 2 * kk) * LBO_A + tap * 16);
                const uint64_t b = desc_at(db, b0 + kk * 2 * 2048);
                if (st == 0 && kk == 0)
                    mma_hc<false>(tmem, a, b, id_main);
                else
                    mma_hc<true>(tmem, a, b, id_main);
                mma_hc<true>(tmem + 64, desc_at(a, CAT_HALF), b, id_corr);
            }
            umma_commit(&bars->w_empty[s & (RING - 1)]);
        }
        umma_commit(&bars->done[2]);
    } else if (tid == 32) {
        for (; p_next < NS_Q2 + NS_S3 + NS_Q3 + NS_M1 + RING; ++p_next) load_stage_c(p_next, p.wstream, ring, bars);
    }
    __syncwarp();
    mbar_wait(&bars->done[2], 0);
    tc_fence_after();
    {   // -> m1 tile (rows chunk * 32 + t), hi at 0, lo at T64
        float o[32];
        drain32(tmem + ((uint32_t)(q * 32) << 16), wh, cst + C_BM, cst[C_SC + 3], o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float o8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                o8[e] = o[8 * j + e];
                if (chunk_ok && rt < M1 && !(fabsf(o8[e]) < 65504.f)) overflow = true;
            }
            store_chunk8<0>(ra, ra + T64, (4 * wh + j) * LBO_A + row * 16, o8);
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- merge_conv2 (64 -> 64, k5) ---------------------------------------------------------------------------------
    if (tid == 0) {
        const uint32_t a0 = smem_addr(ra);
        const uint32_t a_tile[KWM] = {a0, a0, a0, a0, a0};
        const int shift[KWM] = {0, 1, 2, 3, 4};
        issue_merge64<KWM, S_M2>(a_tile, shift, tmem, ring, bars);
        umma_commit(&bars->done[3]);
    } else if (tid == 32) {
        for (; p_next < NS_TOTAL - NS_M3 - NS_M4 + RING; ++p_next) load_stage_c(p_next, p.wstream, ring, bars);
    }
    __syncwarp();
    mbar_wait(&bars->done[3], 0);
    tc_fence_after();
    {   // -> residue-2 tiles Y_r (row = chunk * 32 + t / 2): tile r at r * 2 * T64 (hi), + T64 (lo)
        float o[32];
        drain32(tmem + ((uint32_t)(q * 32) << 16), wh, cst + C_BM + 64, cst[C_SC + 4], o);
        if (rt < M2) {
            uint8_t *t_hi = ra + (rt & 1) * 2 * T64;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float o8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    o8[e] = o[8 * j + e];
                    if (chunk_ok && !(fabsf(o8[e]) < 65504.f)) overflow = true;
                }
                store_chunk8<0>(t_hi, t_hi + T64, (4 * wh + j) * LBO_A + (rc * UR + (rt >> 1)) * 16, o8);
            }
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- merge_conv3 (64 -> 64, k3, stride 2): tap j = residue j & 1 shifted by j >> 1 -------------------------------
    if (tid == 0) {
        const uint32_t y0 = smem_addr(ra), y1 = y0 + 2 * T64;
        const uint32_t a_tile[KW2] = {y0, y1, y0};
        const int shift[KW2] = {0, 0, 1};
        issue_merge64<KW2, S_M3>(a_tile, shift, tmem, ring, bars);
        umma_commit(&bars->done[4]);
    } else if (tid == 32) {
        for (; p_next < NS_TOTAL - NS_M4 + RING; ++p_next) load_stage_c(p_next, p.wstream, ring, bars);
    }
    __syncwarp();
    mbar_wait(&bars->done[4], 0);
    tc_fence_after();
    {
        float o[32];
        drain32(tmem + ((uint32_t)(q * 32) << 16), wh, cst + C_BM + 128, cst[C_SC + 5], o);
        if (rt < M3) {
            uint8_t *t_hi = ra + (rt & 1) * 2 * T64;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float o8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    o8[e] = o[8 * j + e];
                    if (chunk_ok && !(fabsf(o8[e]) < 65504.f)) overflow = true;
                }
                store_chunk8<0>(t_hi, t_hi + T64, (4 * wh + j) * LBO_A + (rc * UR + (rt >> 1)) * 16, o8);
            }
        }
    }
    if (overflow && p.flags) atomicOr(p.flags, 1);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- merge_conv4 (64 -> 64, k3, stride 2) + fc over the flattened [64][M4] block ---------------------------------
    if (tid == 0) {
        const uint32_t y0 = smem_addr(ra), y1 = y0 + 2 * T64;
        const uint32_t a_tile[KW2] = {y0, y1, y0};
        const int shift[KW2] = {0, 0, 1};
        issue_merge64<KW2, S_M4>(a_tile, shift, tmem, ring, bars);
        umma_commit(&bars->done[5]);
    } else if (tid == 32) {
        for (; p_next < NS_TOTAL; ++p_next) load_stage_c(p_next, p.wstream, ring, bars);
    }
    __syncwarp();
    mbar_wait(&bars->done[5], 0);
    tc_fence_after();
    float *part = reinterpret_cast<float *>(sm + OFF_RING);  // [chunk][num_out <= 8]: the ring is idle now
    if (tid < 4 * 8) part[tid] = 0.f;
    __syncthreads();
    {
        float o[32];
        drain32(tmem + ((uint32_t)(q * 32) << 16), wh, cst + C_BM + 192, cst[C_SC + 6], o);
        if (p.dbg_m4 && chunk_ok && rt < M4) {
#pragma unroll
            for (int e = 0; e < 32; ++e) p.dbg_m4[((size_t)(chunk0 + rc) * SIZE + 32 * wh + e) * M4 + rt] = o[e];
        }
        // warps whose 32 rows hold no valid step of any chunk have nothing to add (uniform per warp)
        if (((q * 32) % UR) < M4) {
            for (int oc = 0; oc < p.num_out; ++oc) {
                float acc = 0.f;
                if (rt < M4) {
                    const float *w = p.fcw + (size_t)oc * p.fc_in + (32 * wh) * M4 + rt;  // flatten index = channel * M4 + t
#pragma unroll
                    for (int e = 0; e < 32; ++e) acc = fmaf(__ldg(w + e * M4), o[e], acc);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
                if (lane == 0) atomicAdd(&part[((q * 32) / UR) * 8 + oc], acc);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
    pdl_wait();
    if (tid < C * p.num_out) {
        const int c = tid / p.num_out, oc = tid - c * p.num_out;
        p.logits[(size_t)(chunk0 + c) * p.num_out + oc] = part[c * 8 + oc] + p.fcb[oc];
    }
}
}  // namespace cw

}  // namespace mega

// =========================================================================================================
// host side
// =========================================================================================================
struct MegaWeights {
    float *dev = nullptr;
    size_t off_consts[2] = {0, 0}, off_gtab = 0, off_stream[2] = {0, 0}, off_whh4 = 0, off_wih2T = 0, off_b2 = 0,
           off_fcw = 0, off_fcb = 0;
    int gtab_bytes = 0, kmer_len = 0, num_out = 0;
    int *flags = nullptr;
};

bool mega_supported(const rb200_model_desc &d) {
    using namespace mega;
    if (d.arch != RB200_ARCH_CONVLSTM_W_REF || d.size != SIZE) return false;
    if (d.n_sig_conv != 3 || d.n_seq_conv != 2 || d.n_merge_conv != 1 || d.n_lstm != 2) return false;
    auto is = [](const rb200_conv_desc &c, int ci, int co, int kw, int st) {
        return c.c_in == ci && c.c_out == co && c.kw == kw && c.stride == st;
    };
    if (d.kmer_len > 16) return false;
    if (KW_SEQ1 * ((d.kmer_len + 3) / 4) * 1024 > TAB_CAP) return false;  // seq_conv1 weight fragments
    return is(d.sig_conv[0], 1, 4, KW_SIG1, 1) && is(d.sig_conv[1], 4, 16, KW_SIG2, 1) &&
           is(d.sig_conv[2], 16, SIZE, KW_SIG3, 3) && is(d.seq_conv[0], 4 * d.kmer_len, 16, KW_SEQ1, 1) &&
           is(d.seq_conv[1], 16, SIZE, KW_SEQ2, 3) && is(d.merge_conv[0], 2 * SIZE, SIZE, KW_MRG, 1);
}

bool mega_shape_ok(const rb200_model *m, int T, int seq_width, int map_width) {
    using namespace mega;
    if (m->mega == nullptr) return false;
    if (T > MAX_T || seq_width > MAX_SEQ_W || map_width > MAX_MAP_W || map_width < 2) return false;
    const int T2 = T - 8, Q1 = T - 4;
    if (T2 < KW_SIG3 || Q1 < KW_SEQ2) return false;
    const int T3 = (T2 - KW_SIG3) / 3 + 1, q2 = (Q1 - KW_SEQ2) / 3 + 1;
    return T3 == q2 && T3 <= MAX_T3 && T3 - (KW_MRG - 1) >= 1;
}

namespace {
uint16_t f2h(float x) {
    __half h = __float2half_rn(x);
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
}
float h2f(uint16_t u) {
    __half h;
    memcpy(&h, &u, 2);
    return __half2float(h);
}
uint16_t f2bf(float x) {
    __nv_bfloat16 h = __float2bfloat16_rn(x);
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
}
// power-of-two scale that brings the largest |w| just under 2^15: the fp16 lo parts stay normal numbers
float pow2_scale(const float *w, size_t n) {
    float mx = 0.f;
    for (size_t i = 0; i < n; ++i) mx = fmaxf(mx, fabsf(w[i]));
    if (!(mx > 0.f) || !std::isfinite(mx)) return 1.f;
    int e = 0;
    frexpf(mx, &e);  // mx = f * 2^e, f in [0.5, 1)
    int s = 15 - e;
    if (s > 24) s = 24;
    if (s < -24) s = -24;
    return ldexpf(1.f, s);
}
}  // namespace

int mega_create(rb200_model *m, const float *blob) {
    using namespace mega;
    const rb200_model_desc &d = m->desc;
    const int K = d.kmer_len;
    MegaWeights *mw = new MegaWeights();
    mw->kmer_len = K;
    mw->num_out = d.num_out;
    std::vector<float> host;
    auto reserve = [&](size_t n_floats) {
        size_t at = host.size();
        host.resize(at + ((n_floats + 31) & ~(size_t)31), 0.f);
        return at;
    };
    const float *w_seq2 = blob + d.seq_conv[1].w_off, *w_sig3 = blob + d.sig_conv[2].w_off,
                *w_mrg = blob + d.merge_conv[0].w_off, *w_ih = blob + d.lstm_w_ih_off[0];
    const float sc[4] = {pow2_scale(w_seq2, (size_t)SIZE * 16 * KW_SEQ2), pow2_scale(w_sig3, (size_t)SIZE * 16 * KW_SIG3),
                         pow2_scale(w_mrg, (size_t)SIZE * 128 * KW_MRG), pow2_scale(w_ih, (size_t)256 * SIZE)};
    // ---- constants (one blob per mode: only the inverse scales differ) ----
    for (int mode = 0; mode < 2; ++mode) {
        mw->off_consts[mode] = reserve(CONST_FLOATS);
        float *f = host.data() + mw->off_consts[mode];
        const float *w = blob + d.sig_conv[0].w_off;  // [4][1][5]
        for (int j = 0; j < KW_SIG1; ++j)
            for (int co = 0; co < 4; ++co) f[C_WSIG1 + j * 4 + co] = w[co * KW_SIG1 + j];
        memcpy(f + C_BSIG1, blob + d.sig_conv[0].b_off, 4 * sizeof(float));
        w = blob + d.sig_conv[1].w_off;  // [16][4][5]
        for (int j = 0; j < KW_SIG2; ++j)
            for (int ci = 0; ci < 4; ++ci)
                for (int co = 0; co < 16; ++co) f[C_WSIG2 + (j * 4 + ci) * 16 + co] = w[(co * 4 + ci) * KW_SIG2 + j];
        memcpy(f + C_BSIG2, blob + d.sig_conv[1].b_off, 16 * sizeof(float));
        memcpy(f + C_BSEQ1, blob + d.seq_conv[0].b_off, 16 * sizeof(float));
        memcpy(f + C_BSIG3, blob + d.sig_conv[2].b_off, SIZE * sizeof(float));
        memcpy(f + C_BSEQ2, blob + d.seq_conv[1].b_off, SIZE * sizeof(float));
        memcpy(f + C_BMRG, blob + d.merge_conv[0].b_off, SIZE * sizeof(float));
        for (int i = 0; i < 4; ++i) f[C_SCALE + i] = mode == 0 ? 1.f / sc[i] : 1.f;
        memcpy(f + C_B1, blob + d.lstm_b_off[0], 256 * sizeof(float));
    }
    // ---- seq_conv1 weights as mma.sync.m16n8k16 B fragments: [tap][k-tile][n-tile][hi, lo][lane][b0, b1] ----
    // lane (g = lane >> 2, t = lane & 3): channel n = 8 * n-tile + g, k-mer position pp = 4 * k-tile + t;
    // b0 = {w[base 0], w[base 1]}, b1 = {w[base 2], w[base 3]} (zero past the k-mer), fp16 of S w and of the rest
    {
        const int KT = (K + 3) / 4;
        const float *w = blob + d.seq_conv[0].w_off;  // [16][4K][5], input row = 4 pp + base
        const float s1 = pow2_scale(w, (size_t)16 * 4 * K * KW_SEQ1);
        mw->gtab_bytes = KW_SEQ1 * KT * 2 * 2 * 32 * 8;
        mw->off_gtab = reserve((size_t)mw->gtab_bytes / 4);
        uint32_t *dst = reinterpret_cast<uint32_t *>(host.data() + mw->off_gtab);
        for (int j = 0; j < KW_SEQ1; ++j)
            for (int kt = 0; kt < KT; ++kt)
                for (int nt = 0; nt < 2; ++nt)
                    for (int part = 0; part < 2; ++part)
                        for (int lane = 0; lane < 32; ++lane) {
                            const int n = 8 * nt + (lane >> 2), pp = 4 * kt + (lane & 3);
                            uint16_t h[4];
                            for (int base = 0; base < 4; ++base) {
                                const float v = pp < K ? w[(n * 4 * K + 4 * pp + base) * KW_SEQ1 + j] * s1 : 0.f;
                                const uint16_t hi = f2h(v);
                                h[base] = part == 0 ? hi : f2h(v - h2f(hi));
                            }
                            uint32_t *o = dst + (((((size_t)j * KT + kt) * 2 + nt) * 2 + part) * 32 + lane) * 2;
                            o[0] = (uint32_t)h[0] | ((uint32_t)h[1] << 16);
                            o[1] = (uint32_t)h[2] | ((uint32_t)h[3] << 16);
                        }
        for (int mode = 0; mode < 2; ++mode) host[mw->off_consts[mode] + C_SCALE + 5] = 1.f / s1;
    }
    // ---- weight streams ----
    for (int mode = 0; mode < 2; ++mode) {
        const int NB = mode == 0 ? 128 : 64, B_LBO = NB * 16, TAPS = STAGE_BYTES / (2 * B_LBO),
                  MRG_KC = STAGE_BYTES / B_LBO;
        const int ns_seq = (KW_SEQ2 + TAPS - 1) / TAPS, ns_sig = (KW_SIG3 + TAPS - 1) / TAPS,
                  ns_mrg = KW_MRG * (16 / MRG_KC), ns_xp = mode == 0 ? 8 : 4;
        const int total = ns_seq + ns_sig + ns_mrg + ns_xp;
        mw->off_stream[mode] = reserve((size_t)total * STAGE_BYTES / 4);
        uint8_t *st0 = reinterpret_cast<uint8_t *>(host.data() + mw->off_stream[mode]);
        // element (row n, k) of a [K chunk][rows][8] tile at `base` with K-chunk distance lbo
        auto put = [&](uint8_t *base, int lbo, int n, int k, uint16_t v) {
            memcpy(base + (size_t)(k >> 3) * lbo + (size_t)n * 16 + (k & 7) * 2, &v, 2);
        };
        auto put_w = [&](uint8_t *base, int lbo, int n, int k, float wv, float scale) {
            if (mode == 0) {
                const uint16_t hi = f2h(wv * scale);
                put(base, lbo, n, k, hi);
                put(base, lbo, n + NB / 2, k, f2h(wv * scale - h2f(hi)));
            } else {
                put(base, lbo, n, k, f2bf(wv));
            }
        };
        int s = 0;
        for (int conv = 0; conv < 2; ++conv) {  // stream order = execution order: seq_conv2, then sig_conv3
            const int KW = conv == 0 ? KW_SEQ2 : KW_SIG3;
            const float *w = conv == 0 ? w_seq2 : w_sig3;  // [64][16][KW]
            const float scale = sc[conv];
            const int ns = conv == 0 ? ns_seq : ns_sig;
            for (int st = 0; st < ns; ++st, ++s)
                for (int tp = 0; tp < TAPS; ++tp) {
                    const int j = st * TAPS + tp;
                    if (j >= KW) continue;
                    uint8_t *base = st0 + (size_t)s * STAGE_BYTES + (size_t)tp * 2 * B_LBO;
                    for (int n = 0; n < SIZE; ++n)
                        for (int k = 0; k < 16; ++k) put_w(base, B_LBO, n, k, w[(n * 16 + k) * KW + j], scale);
                }
        }
        for (int st = 0; st < ns_mrg; ++st, ++s) {  // merge: stage = (channel block kb, tap)
            const int kb = st / KW_MRG, tap = st - kb * KW_MRG;
            uint8_t *base = st0 + (size_t)s * STAGE_BYTES;
            for (int n = 0; n < SIZE; ++n)
                for (int k = 0; k < MRG_KC * 8; ++k)
                    put_w(base, B_LBO, n, k, w_mrg[(n * 128 + kb * MRG_KC * 8 + k) * KW_MRG + tap], sc[2]);
        }
        for (int st = 0; st < ns_xp; ++st, ++s) {  // projection: [K chunk (2)][256 gate rows][8]
            uint8_t *base = st0 + (size_t)s * STAGE_BYTES;
            const int k16 = mode == 0 ? st >> 1 : st;
            for (int n = 0; n < 256; ++n)
                for (int k = 0; k < 16; ++k) {
                    const float wv = w_ih[n * SIZE + k16 * 16 + k];
                    if (mode == 0) {
                        const uint16_t hi = f2h(wv * sc[3]);
                        put(base, 4096, n, k, (st & 1) ? f2h(wv * sc[3] - h2f(hi)) : hi);
                    } else {
                        put(base, 4096, n, k, f2bf(wv));
                    }
                }
        }
    }
    // ---- recurrence / tail weights (same layouts as rb200_fused.cu K3) ----
    // W_hh1 and W_ih2 as the A fragments of mma.sync.m16n8k16 (row-major 16 x 16, fp16): register a of fragment
    // ((m-tile * 4 + k-tile) * 2 + part) of thread (warp w, g = lane >> 2, t = lane & 3) holds
    // W[row][k], W[row][k + 1] with row = gate * 64 + 8 w + g, gate = 2 * m-tile + (a & 1),
    // k = 16 * k-tile + 2 t + 8 * (a >> 1); part 0 = fp16(S w), part 1 = fp16(S w - part 0)
    auto put_gate_frags = [&](size_t off, const float *wsrc, int scale_slot) {
        const float sc = pow2_scale(wsrc, (size_t)256 * SIZE);
        uint32_t *dst = reinterpret_cast<uint32_t *>(host.data() + off);
        for (int tid = 0; tid < 256; ++tid) {
            const int w = tid >> 5, g = (tid & 31) >> 2, t = tid & 3;
            for (int mt = 0; mt < 2; ++mt)
                for (int kt = 0; kt < 4; ++kt)
                    for (int part = 0; part < 2; ++part)
                        for (int a = 0; a < 4; ++a) {
                            const int row = (2 * mt + (a & 1)) * 64 + 8 * w + g;
                            const int k = 16 * kt + 2 * t + 8 * (a >> 1);
                            uint16_t h2[2];
                            for (int e = 0; e < 2; ++e) {
                                const float v = wsrc[row * SIZE + k + e] * sc;
                                const uint16_t hi = f2h(v);
                                h2[e] = part == 0 ? hi : f2h(v - h2f(hi));
                            }
                            const int q = (mt * 4 + kt) * 2 + part;
                            dst[((size_t)q * 256 + tid) * 4 + a] = (uint32_t)h2[0] | ((uint32_t)h2[1] << 16);
                        }
        }
        for (int mode = 0; mode < 2; ++mode) host[mw->off_consts[mode] + C_SCALE + scale_slot] = 1.f / sc;
    };
    mw->off_whh4 = reserve(SIZE * 256);
    put_gate_frags(mw->off_whh4, blob + d.lstm_w_hh_off[0], 4);
    mw->off_wih2T = reserve(SIZE * 256);
    put_gate_frags(mw->off_wih2T, blob + d.lstm_w_ih_off[1], 6);
    mw->off_b2 = reserve(256);
    memcpy(host.data() + mw->off_b2, blob + d.lstm_b_off[1], 256 * sizeof(float));
    mw->off_fcw = reserve((size_t)d.num_out * SIZE);
    memcpy(host.data() + mw->off_fcw, blob + d.fc_w_off, (size_t)d.num_out * SIZE * sizeof(float));
    mw->off_fcb = reserve(d.num_out);
    memcpy(host.data() + mw->off_fcb, blob + d.fc_b_off, d.num_out * sizeof(float));

    cudaError_t e = cudaMalloc(&mw->dev, host.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(mw->dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&mw->flags, 16);
    if (e == cudaSuccess) e = cudaMemset(mw->flags, 0, 16);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(mega_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(mega_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) {
        set_error("single-kernel path set-up failed: %s", cudaGetErrorString(e));
        if (mw->dev) cudaFree(mw->dev);
        if (mw->flags) cudaFree(mw->flags);
        delete mw;
        return RB200_ERR_CUDA;
    }
    m->mega = mw;
    return RB200_OK;
}

void mega_destroy(rb200_model *m) {
    if (!m->mega) return;
    if (m->mega->dev) cudaFree(m->mega->dev);
    if (m->mega->flags) cudaFree(m->mega->flags);
    delete m->mega;
    m->mega = nullptr;
}

struct ConvMegaWeights;
int *conv_mega_flags(rb200_model *m);

int mega_read_flags(rb200_model *m, int *out, bool clear) {
    *out = 0;
    int *flags = m->mega ? m->mega->flags : conv_mega_flags(m);
    if (!flags) return RB200_OK;
    RB200_CUDA_TRY(cudaDeviceSynchronize());
    RB200_CUDA_TRY(cudaMemcpy(out, flags, sizeof(int), cudaMemcpyDeviceToHost));
    if (clear) RB200_CUDA_TRY(cudaMemset(flags, 0, sizeof(int)));
    return RB200_OK;
}

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

int mega_forward_compact(rb200_model *m, Workspace &ws, const float *sigs, const int8_t *seqs, int seq_width,
                         const int16_t *maps, int map_width, const int16_t *lens, int B, int T, float *logits,
                         cudaStream_t stream, int mode, const GatherTarget *gather, const float *enc_dense) {
    using namespace mega;
    const MegaWeights *mw = m->mega;
    if (enc_dense) {  // dense interface: the sizes of the (absent) compact arrays only pass the shape check
        seq_width = mw->kmer_len;
        map_width = 2;
    }
    RB200_REQUIRE(mega_shape_ok(m, T, seq_width, map_width), "chunk shape not supported by the single-kernel path");
    const int T3 = (T - 8 - KW_SIG3) / 3 + 1, TM = T3 - (KW_MRG - 1);
    Params p = {};
    p.sigs = sigs;
    p.seqs = seqs;
    p.maps = maps;
    p.lens = lens;
    p.seq_width = seq_width;
    p.map_width = map_width;
    p.B = B;
    p.T = T;
    p.kmer_len = mw->kmer_len;
    p.num_out = mw->num_out;
    p.consts = mw->dev + mw->off_consts[mode];
    p.gtab = mw->dev + mw->off_gtab;
    p.gtab_bytes = mw->gtab_bytes;
    p.wstream = reinterpret_cast<const uint8_t *>(mw->dev + mw->off_stream[mode]);
    p.whh4 = reinterpret_cast<const float4 *>(mw->dev + mw->off_whh4);
    p.wih2T = mw->dev + mw->off_wih2T;
    p.b2 = mw->dev + mw->off_b2;
    p.fcw = mw->dev + mw->off_fcw;
    p.fcb = mw->dev + mw->off_fcb;
    p.logits = logits;
    p.dbg_cat = p.dbg_m = p.dbg_xp = nullptr;
    p.flags = mw->flags;
    p.q1_in = nullptr;
    size_t q1_bytes = 0;
    if (enc_dense) {
        // K0 (rb200_fused.cu): the honest dense seq_conv1 + BN + swish over the materialised one-hot
        q1_bytes = align256((size_t)B * (T - (KW_SEQ1 - 1)) * 16 * sizeof(float));
        RB200_REQUIRE(!m->keep_debug, "debug tensors are not kept on the dense single-kernel path");
        int rc = ws.ensure(q1_bytes);
        if (rc) return rc;
        float *q1 = reinterpret_cast<float *>(ws.base);
        rc = fused_dense_seq1(m, enc_dense, q1, B, T, stream);
        if (rc) return rc;
        p.q1_in = q1;
    }
    p.stamps = nullptr;
    p.peers = nullptr;
    p.n_peers = 0;
    p.peer_off = 0;
    p.mc_base = nullptr;
    p.flag_off = -1;
    p.ship_cta = 0;
    p.self_rank = -1;
    p.ship_src = nullptr;
    p.ship_off = 0;
    p.ship_n = 0;
    p.ship_flag = -1;
    if (gather != nullptr && gather->deferred) {
        RB200_REQUIRE(gather->n_peers >= 1 && gather->n_peers <= 32 && gather->peers_dev != nullptr &&
                          gather->self_rank >= 0 && gather->self_rank < gather->n_peers,
                      "deferred gather target needs 1..32 peer buffers and this rank's index");
        RB200_REQUIRE(gather->ship_src == nullptr ||
                          (gather->ship_count > 0 && (gather->dst_offset & 3) == 0 &&
                           (reinterpret_cast<uintptr_t>(gather->ship_src) & 15) == 0),
                      "shipped block must be 16-byte aligned in the source and in every buffer");
        p.peers = gather->peers_dev;
        p.n_peers = gather->n_peers;
        p.ship_cta = 1;  // compute CTAs store locally only
        p.self_rank = gather->self_rank;
        p.ship_src = gather->ship_src;
        p.ship_off = gather->dst_offset;
        p.ship_n = (int)gather->ship_count;
        p.ship_flag = gather->flag_offset;
        p.mc_base = gather->multicast_base;
    } else if (gather != nullptr) {
        RB200_REQUIRE(gather->n_peers >= 1 && gather->n_peers <= 32 && gather->peers_dev != nullptr,
                      "gather target needs 1..32 peer buffers");
        p.peers = gather->peers_dev;
        p.n_peers = gather->n_peers;
        p.peer_off = gather->dst_offset;
        p.mc_base = gather->multicast_base;
        p.flag_off = gather->flag_offset;
    }
    // profiling aid: RB200_MEGA_TRACE=<file> records {smid, start, end, cycles} of every CTA of 64 consecutive
    // launches (after 300 earlier ones: steady state) and writes them to <file> as text
    static const char *trace_path = getenv("RB200_MEGA_TRACE");
    static long long *trace_dev = nullptr;
    static int trace_launch = 0;
    constexpr int TRACE_SKIP = 300, TRACE_N = 64, TRACE_GRID = 1024;
    p.trace = nullptr;
    const int grid_ctas = (B + G - 1) / G;
    if (trace_path && grid_ctas <= TRACE_GRID) {
        if (!trace_dev) {
            cudaMalloc(&trace_dev, (size_t)TRACE_N * TRACE_GRID * 20 * sizeof(long long));
            cudaMemset(trace_dev, 0, (size_t)TRACE_N * TRACE_GRID * 20 * sizeof(long long));
        }
        const int k = trace_launch - TRACE_SKIP;
        if (k >= 0 && k < TRACE_N) p.trace = trace_dev + (size_t)k * TRACE_GRID * 20;
        if (k == TRACE_N + 8) {
            cudaStreamSynchronize(stream);
            std::vector<long long> h((size_t)TRACE_N * TRACE_GRID * 20);
            cudaMemcpy(h.data(), trace_dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
            if (FILE *f = fopen(trace_path, "w")) {
                for (int l = 0; l < TRACE_N; ++l)
                    for (int c = 0; c < grid_ctas; ++c) {
                        const long long *r = &h[((size_t)l * TRACE_GRID + c) * 20];
                        fprintf(f, "%d %d", l, c);  // then smid, start ns, end ns, cycles, 14 phase lengths (+1 spare)
                        for (int i = 0; i < 19; ++i) fprintf(f, " %lld", r[i]);
                        fprintf(f, "\n");
                    }
                fclose(f);
            }
        }
        ++trace_launch;
    }
    static const bool want_stamps = getenv("RB200_MEGA_STAMPS") != nullptr;  // profiling aid
    static long long *stamps_dev = nullptr;
    if (want_stamps) {
        if (!stamps_dev) cudaMalloc(&stamps_dev, 32 * sizeof(long long));
        p.stamps = stamps_dev;
    }
    if (m->keep_debug) {
        const size_t n_cat = (size_t)B * 128 * T3, n_m = (size_t)B * SIZE * TM, n_xp = (size_t)B * 256 * TM;
        int rc = ws.ensure(align256(n_cat * 4) + align256(n_m * 4) + align256(n_xp * 4));
        if (rc) return rc;
        p.dbg_cat = reinterpret_cast<float *>(ws.base);
        p.dbg_m = reinterpret_cast<float *>(ws.base + align256(n_cat * 4));
        p.dbg_xp = reinterpret_cast<float *>(ws.base + align256(n_cat * 4) + align256(n_m * 4));
        m->debug.clear();
        m->debug.push_back({"cat", p.dbg_cat, B, 128, T3});
        m->debug.push_back({"merge1", p.dbg_m, B, SIZE, TM});
        m->debug.push_back({"xproj", p.dbg_xp, B, 256, TM});
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((B + G - 1) / G + p.ship_cta);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = m->keep_debug || m->profile ? 0 : 1;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (m->profile) {
        for (int i = 0; i < 4; ++i) RB200_CUDA_TRY(cudaEventCreate(&ev[i]));
        RB200_CUDA_TRY(cudaEventRecord(ev[0], stream));
    }
    if (mode == 0)
        RB200_CUDA_TRY(cudaLaunchKernelEx(&cfg, mega_kernel<0>, p));
    else
        RB200_CUDA_TRY(cudaLaunchKernelEx(&cfg, mega_kernel<1>, p));
    if (m->profile) {
        // the per-kernel profile has three slots (K1, K2, K3 of the three-kernel path): the single kernel
        // reports its whole time in the first one
        for (int i = 1; i < 4; ++i) RB200_CUDA_TRY(cudaEventRecord(ev[i], stream));
        for (int i = 0; i < 4; ++i) m->prof_events.push_back(ev[i]);
    }
    if (want_stamps) {
        long long h[32];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, stamps_dev, sizeof(h), cudaMemcpyDeviceToHost);
        static const char *names[14] = {"prologue", "stage+sidx", "gather", "seq2||sig12", "sig3 mma", "E1 cat",
                                        "merge mma", "E2 m", "xproj mma", "E3 xp", "whh load", "recurrence",
                                        "lstm2", "pdl_wait"};
        fprintf(stderr, "[mega stamps, cycles]");
        for (int i = 0; i < 14; ++i) fprintf(stderr, " %s %lld |", names[i], h[i + 1] - h[i]);
        fprintf(stderr, " total %lld | first weight stage landed after %lld, first three after %lld\n", h[14] - h[0],
                h[15] & 0xFFFFFFFFll, h[15] >> 32);
        fprintf(stderr, "   sig3 stages ready at (cycles after the phase start):");
        for (int i = 0; i < 5; ++i) fprintf(stderr, " %lld", h[16 + i] - h[4]);
        fprintf(stderr, " | all issued %lld | done %lld\n   seq2 stages ready at:", h[24] - h[4], h[5] - h[4]);
        for (int i = 0; i < 7; ++i) fprintf(stderr, " %lld", h[25 + i] - h[3]);
        fprintf(stderr, " | sig_conv1 done %lld, producer done %lld, sig_conv2 done %lld\n", h[21] - h[3], h[22] - h[3],
                h[23] - h[3]);
    }
    m->launches += 1;
    m->last_impl = mode == 0 ? RB200_IMPL_FUSED_MEGA : RB200_IMPL_FUSED_BF16;
    return RB200_OK;
}

// ---- Conv_w_ref single kernel: host side ------------------------------------------------------------------
struct ConvMegaWeights {
    float *dev = nullptr;
    size_t off_consts = 0, off_gtab = 0, off_stream = 0, off_fcw = 0, off_fcb = 0;
    int gtab_bytes = 0, kmer_len = 0, num_out = 0, fc_in = 0;
    int *flags = nullptr;
};

int *conv_mega_flags(rb200_model *m) { return m->conv_mega ? m->conv_mega->flags : nullptr; }

bool conv_mega_supported(const rb200_model_desc &d) {
    using namespace mega;
    if (d.arch != RB200_ARCH_CONV_W_REF || d.size != SIZE || d.num_out > 8) return false;
    if (d.n_sig_conv != 3 || d.n_seq_conv != 3 || d.n_merge_conv != 4 || d.n_lstm != 0) return false;
    auto is = [](const rb200_conv_desc &c, int ci, int co, int kw, int st) {
        return c.c_in == ci && c.c_out == co && c.kw == kw && c.stride == st;
    };
    if ((cw::KW1 * d.kmer_len * 4 + 1) * GROW * 4 > RING * STAGE_BYTES) return false;  // the table borrows the ring
    return is(d.sig_conv[0], 1, 4, cw::KW1, 1) && is(d.sig_conv[1], 4, 16, cw::KW1, 1) &&
           is(d.sig_conv[2], 16, SIZE, cw::KW3, 3) && is(d.seq_conv[0], 4 * d.kmer_len, 16, cw::KW1, 1) &&
           is(d.seq_conv[1], 16, 32, cw::KW1, 1) && is(d.seq_conv[2], 32, SIZE, cw::KW3, 3) &&
           is(d.merge_conv[0], 2 * SIZE, SIZE, cw::KWM, 1) && is(d.merge_conv[1], SIZE, SIZE, cw::KWM, 1) &&
           is(d.merge_conv[2], SIZE, SIZE, cw::KW2, 2) && is(d.merge_conv[3], SIZE, SIZE, cw::KW2, 2);
}

bool conv_mega_shape_ok(const rb200_model *m, int T, int seq_width, int map_width) {
    using namespace mega;
    if (m->conv_mega == nullptr) return false;
    if (T > cw::MAX_T || map_width < 2) return false;
    const int gc = T <= 100 ? 4 : 2, ur = MROWS / gc;
    // staging areas and the per-chunk gather sums
    if (gc * seq_width > 256 || gc * map_width * 2 > 384 || (map_width - 1) * cw::KW1 * GROW * 4 > cw::GS_CAP)
        return false;
    const int S1 = T - (cw::KW1 - 1), S2 = S1 - (cw::KW1 - 1);
    if (S2 < cw::KW3 || gc * S1 * 16 > cw::S1_CAP) return false;
    const int S3 = (S2 - cw::KW3) / 3 + 1, M2 = S3 - 2 * (cw::KWM - 1);
    // tile rows per chunk: the stride-3 residue tiles hold ceil(S2 / 3) rows, the merge taps shift by <= 4
    if ((S2 + 2) / 3 > ur || S3 + (cw::KWM - 1) > ur || S1 > 3 * ur || M2 < cw::KW2) return false;
    const int M3 = (M2 - cw::KW2) / 2 + 1;
    if (M3 < cw::KW2) return false;
    const int M4 = (M3 - cw::KW2) / 2 + 1;
    return m->conv_mega->fc_in == SIZE * M4;  // stock model: chunk_len 100 -> 3 steps -> 192 inputs
}

int conv_mega_create(rb200_model *m, const float *blob) {
    using namespace mega;
    const rb200_model_desc &d = m->desc;
    const int K = d.kmer_len;
    ConvMegaWeights *mw = new ConvMegaWeights();
    mw->kmer_len = K;
    mw->num_out = d.num_out;
    mw->fc_in = d.fc_in;
    std::vector<float> host;
    auto reserve = [&](size_t n_floats) {
        size_t at = host.size();
        host.resize(at + ((n_floats + 31) & ~(size_t)31), 0.f);
        return at;
    };
    const float *w_q2 = blob + d.seq_conv[1].w_off, *w_s3 = blob + d.sig_conv[2].w_off, *w_q3 = blob + d.seq_conv[2].w_off;
    const float *w_m[4] = {blob + d.merge_conv[0].w_off, blob + d.merge_conv[1].w_off, blob + d.merge_conv[2].w_off,
                           blob + d.merge_conv[3].w_off};
    const float sc[7] = {pow2_scale(w_q2, (size_t)32 * 16 * cw::KW1), pow2_scale(w_s3, (size_t)SIZE * 16 * cw::KW3),
                         pow2_scale(w_q3, (size_t)SIZE * 32 * cw::KW3), pow2_scale(w_m[0], (size_t)SIZE * 128 * cw::KWM),
                         pow2_scale(w_m[1], (size_t)SIZE * SIZE * cw::KWM), pow2_scale(w_m[2], (size_t)SIZE * SIZE * cw::KW2),
                         pow2_scale(w_m[3], (size_t)SIZE * SIZE * cw::KW2)};
    mw->off_consts = reserve(cw::CONST_FLOATS);
    {
        float *f = host.data() + mw->off_consts;
        const float *w = blob + d.sig_conv[0].w_off;  // [4][1][11]
        for (int j = 0; j < cw::KW1; ++j)
            for (int co = 0; co < 4; ++co) f[cw::C_WS1 + j * 4 + co] = w[co * cw::KW1 + j];
        memcpy(f + cw::C_BS1, blob + d.sig_conv[0].b_off, 4 * sizeof(float));
        w = blob + d.sig_conv[1].w_off;  // [16][4][11]
        for (int j = 0; j < cw::KW1; ++j)
            for (int ci = 0; ci < 4; ++ci)
                for (int co = 0; co < 16; ++co) f[cw::C_WS2 + (j * 4 + ci) * 16 + co] = w[(co * 4 + ci) * cw::KW1 + j];
        memcpy(f + cw::C_BS2, blob + d.sig_conv[1].b_off, 16 * sizeof(float));
        memcpy(f + cw::C_BQ1, blob + d.seq_conv[0].b_off, 16 * sizeof(float));
        memcpy(f + cw::C_BQ2, blob + d.seq_conv[1].b_off, 32 * sizeof(float));
        memcpy(f + cw::C_BS3, blob + d.sig_conv[2].b_off, SIZE * sizeof(float));
        memcpy(f + cw::C_BQ3, blob + d.seq_conv[2].b_off, SIZE * sizeof(float));
        for (int l = 0; l < 4; ++l) memcpy(f + cw::C_BM + 64 * l, blob + d.merge_conv[l].b_off, SIZE * sizeof(float));
        for (int i = 0; i < 7; ++i) f[cw::C_SC + i] = 1.f / sc[i];
    }
    {
        const int rows = cw::KW1 * K * 4 + 1;
        mw->gtab_bytes = ((rows * GROW * 4) + 15) & ~15;
        mw->off_gtab = reserve((size_t)rows * GROW);
        const float *w = blob + d.seq_conv[0].w_off;  // [16][4K][11]
        for (int j = 0; j < cw::KW1; ++j)
            for (int rw = 0; rw < 4 * K; ++rw)
                for (int co = 0; co < 16; ++co)
                    host[mw->off_gtab + (size_t)(j * 4 * K + rw) * GROW + co] = w[(co * 4 * K + rw) * cw::KW1 + j];
    }
    mw->off_stream = reserve((size_t)cw::NS_TOTAL * STAGE_BYTES / 4);
    {
        uint8_t *st0 = reinterpret_cast<uint8_t *>(host.data() + mw->off_stream);
        // tile [K chunk][n_rows hi + n_rows lo][8 halves]: element (n, k), hi in rows [0, n_out), lo in [n_out, 2 n_out)
        auto put_w = [&](uint8_t *base, int lbo, int n_out, int n, int k, float wv, float scale) {
            const uint16_t hi = f2h(wv * scale);
            const uint16_t lo = f2h(wv * scale - h2f(hi));
            memcpy(base + (size_t)(k >> 3) * lbo + (size_t)n * 16 + (k & 7) * 2, &hi, 2);
            memcpy(base + (size_t)(k >> 3) * lbo + (size_t)(n + n_out) * 16 + (k & 7) * 2, &lo, 2);
        };
        int s = 0;
        for (int st = 0; st < cw::NS_Q2; ++st, ++s)  // seq_conv2 [32][16][11]: four taps per stage
            for (int tp = 0; tp < 4; ++tp) {
                const int j = st * 4 + tp;
                if (j >= cw::KW1) continue;
                uint8_t *base = st0 + (size_t)s * STAGE_BYTES + (size_t)tp * 2048;
                for (int n = 0; n < 32; ++n)
                    for (int k = 0; k < 16; ++k) put_w(base, 1024, 32, n, k, w_q2[(n * 16 + k) * cw::KW1 + j], sc[0]);
            }
        for (int st = 0; st < cw::NS_S3; ++st, ++s)  // sig_conv3 [64][16][9]: two taps per stage
            for (int tp = 0; tp < 2; ++tp) {
                const int j = st * 2 + tp;
                if (j >= cw::KW3) continue;
                uint8_t *base = st0 + (size_t)s * STAGE_BYTES + (size_t)tp * 4096;
                for (int n = 0; n < SIZE; ++n)
                    for (int k = 0; k < 16; ++k) put_w(base, 2048, SIZE, n, k, w_s3[(n * 16 + k) * cw::KW3 + j], sc[1]);
            }
        for (int j = 0; j < cw::KW3; ++j, ++s) {  // seq_conv3 [64][32][9]: one tap per stage
            uint8_t *base = st0 + (size_t)s * STAGE_BYTES;
            for (int n = 0; n < SIZE; ++n)
                for (int k = 0; k < 32; ++k) put_w(base, 2048, SIZE, n, k, w_q3[(n * 32 + k) * cw::KW3 + j], sc[2]);
        }
        for (int st = 0; st < cw::NS_M1; ++st, ++s) {  // merge_conv1 [64][128][5]: (channel block, tap)
            const int kb = st / cw::KWM, tap = st - kb * cw::KWM;
            uint8_t *base = st0 + (size_t)s * STAGE_BYTES;
            for (int n = 0; n < SIZE; ++n)
                for (int k = 0; k < 32; ++k)
                    put_w(base, 2048, SIZE, n, k, w_m[0][(n * 128 + kb * 32 + k) * cw::KWM + tap], sc[3]);
        }
        for (int l = 1; l < 4; ++l) {  // merge_conv2..4 [64][64][kw]: (tap, channel half)
            const int kw = l == 1 ? cw::KWM : cw::KW2;
            for (int j = 0; j < kw; ++j)
                for (int h = 0; h < 2; ++h, ++s) {
                    uint8_t *base = st0 + (size_t)s * STAGE_BYTES;
                    for (int n = 0; n < SIZE; ++n)
                        for (int k = 0; k < 32; ++k)
                            put_w(base, 2048, SIZE, n, k, w_m[l][(n * SIZE + 32 * h + k) * kw + j], sc[3 + l]);
                }
        }
        if (s != cw::NS_TOTAL) {
            set_error("internal: Conv_w_ref weight stream has %d stages, expected %d", s, cw::NS_TOTAL);
            delete mw;
            return RB200_ERR_INVALID;
        }
    }
    mw->off_fcw = reserve((size_t)d.num_out * d.fc_in);
    memcpy(host.data() + mw->off_fcw, blob + d.fc_w_off, (size_t)d.num_out * d.fc_in * sizeof(float));
    mw->off_fcb = reserve(d.num_out);
    memcpy(host.data() + mw->off_fcb, blob + d.fc_b_off, d.num_out * sizeof(float));
    cudaError_t e = cudaMalloc(&mw->dev, host.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(mw->dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&mw->flags, 16);
    if (e == cudaSuccess) e = cudaMemset(mw->flags, 0, 16);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(cw::conv_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cw::SMEM_BYTES);
    if (e != cudaSuccess) {
        set_error("Conv_w_ref single-kernel path set-up failed: %s", cudaGetErrorString(e));
        if (mw->dev) cudaFree(mw->dev);
        if (mw->flags) cudaFree(mw->flags);
        delete mw;
        return RB200_ERR_CUDA;
    }
    m->conv_mega = mw;
    return RB200_OK;
}

void conv_mega_destroy(rb200_model *m) {
    if (!m->conv_mega) return;
    if (m->conv_mega->dev) cudaFree(m->conv_mega->dev);
    if (m->conv_mega->flags) cudaFree(m->conv_mega->flags);
    delete m->conv_mega;
    m->conv_mega = nullptr;
}

int conv_mega_forward_compact(rb200_model *m, Workspace &ws, const float *sigs, const int8_t *seqs, int seq_width,
                              const int16_t *maps, int map_width, const int16_t *lens, int B, int T, float *logits,
                              cudaStream_t stream) {
    using namespace mega;
    const ConvMegaWeights *mw = m->conv_mega;
    RB200_REQUIRE(conv_mega_shape_ok(m, T, seq_width, map_width), "chunk shape not supported by the single-kernel path");
    const int S3 = (T - 2 * (cw::KW1 - 1) - cw::KW3) / 3 + 1;
    const int M3 = (S3 - 2 * (cw::KWM - 1) - cw::KW2) / 2 + 1, M4 = (M3 - cw::KW2) / 2 + 1;
    cw::Params p;
    p.sigs = sigs;
    p.seqs = seqs;
    p.maps = maps;
    p.lens = lens;
    p.seq_width = seq_width;
    p.map_width = map_width;
    p.B = B;
    p.T = T;
    p.kmer_len = mw->kmer_len;
    p.num_out = mw->num_out;
    p.fc_in = mw->fc_in;
    p.consts = mw->dev + mw->off_consts;
    p.gtab = mw->dev + mw->off_gtab;
    p.gtab_bytes = mw->gtab_bytes;
    p.wstream = reinterpret_cast<const uint8_t *>(mw->dev + mw->off_stream);
    p.fcw = mw->dev + mw->off_fcw;
    p.fcb = mw->dev + mw->off_fcb;
    p.logits = logits;
    p.dbg_cat = p.dbg_m4 = nullptr;
    p.flags = mw->flags;
    if (m->keep_debug) {
        const size_t n_cat = (size_t)B * 128 * S3, n_m4 = (size_t)B * SIZE * M4;
        int rc = ws.ensure(align256(n_cat * 4) + align256(n_m4 * 4));
        if (rc) return rc;
        p.dbg_cat = reinterpret_cast<float *>(ws.base);
        p.dbg_m4 = reinterpret_cast<float *>(ws.base + align256(n_cat * 4));
        m->debug.clear();
        m->debug.push_back({"cat", p.dbg_cat, B, 128, S3});
        m->debug.push_back({"merge4", p.dbg_m4, B, SIZE, M4});
    }
    cudaLaunchConfig_t cfg = {};
    const int gc = T <= 100 ? 4 : 2;
    cfg.gridDim = dim3((B + gc - 1) / gc);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = cw::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = m->keep_debug ? 0 : 1;
    RB200_CUDA_TRY(cudaLaunchKernelEx(&cfg, cw::conv_mega_kernel, p));
    m->launches += 1;
    m->last_impl = RB200_IMPL_FUSED_MEGA;
    return RB200_OK;
}

}  // namespace rb200
