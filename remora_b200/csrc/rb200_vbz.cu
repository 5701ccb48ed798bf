// POD5 signal decode on the GPU ("next" row 3, SURVEY.md 8f: "svb16/zig-zag delta decode").
//
// A POD5 signal row ("minknow.vbz") is zstd( svb16( zigzag( delta( int16 samples )))).  zstd is undone
// on the host (pyarrow's codec); this kernel undoes the rest for a batch of rows:
//   svb16   n samples -> ceil(n/8) key bytes (bit i of byte i/8, LSB first: 1 = the value takes two
//           bytes) followed by the little-endian data bytes
//   zigzag  u -> (u >> 1) ^ -(u & 1)
//   delta   running sum modulo 2^16, starting at 0 for every row
// The reference reaches the same code through the pod5 package's C extension (pod5.ReadRecord.signal,
// src/remora/io.py:455); the output is bit-identical to remora_b200.io.decode_vbz and to the samples
// the reference's test files hold.
//
// One CTA per TILE of 8192 samples (grid = tiles x rows), so a batch of rows fills the machine and the
// kernel streams at HBM speed instead of walking each row serially:
//   0  where the tile's data bytes start: one byte per earlier sample of the row + one per earlier
//      two-byte value = popcount of the row's earlier key bytes (re-read by the row's later tiles: 1 KB per
//      tile, L2 hits), CTA reduction - no dependence on other CTAs
//   1  256 key words (32 samples each) -> popcounts -> CTA exclusive scan = data offset of every word
//   2  the tile's data bytes staged into shared memory with aligned 32-bit loads (coalesced)
//   3  thread j decodes the 32 samples of key word j sequentially, two per step: funnel shift to the byte
//      position, one byte permute chosen by the two key bits, zig-zag and running sum on two 16-bit lanes
//      at once (about 10 instructions per sample; the sample's byte position is a running counter)
//   4  CTA scan of the per-thread totals
//   5  the running sum carried in from the row's earlier tiles: every tile publishes the sum of ITS deltas
//      (flag + 16-bit sum in one word) as soon as step 4 is done, and reads its predecessors' words -
//      aggregates, not prefixes, so no tile waits on a chain; predecessors have lower block indices and
//      are resident or finished (dispatch order), so the spin cannot deadlock
//   6  64 contiguous output bytes per thread
// Byte/integer work bound by HBM: about 1.17 B read + 2 B written per sample.
#include "rb200_internal.cuh"

namespace rb200 {

namespace {

constexpr int kThreads = 256;
constexpr int kTile = kThreads * 32;  // samples per tile

__device__ __forceinline__ int cta_exclusive_scan(int v, int *warp_tot, int *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
        const int t = warp_tot[w];
        if (w < warp) base += t;
        tot += t;
    }
    __syncthreads();  // warp_tot reusable
    *total = tot;
    return base + inc - v;
}

__device__ __forceinline__ int cta_sum(int v, int *warp_tot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) warp_tot[warp] = v;
    __syncthreads();
    int tot = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) tot += warp_tot[w];
    __syncthreads();
    return tot;
}

constexpr uint32_t kAggFlag = 0x80000000u;

__global__ void __launch_bounds__(kThreads)
svb16_decode_kernel(const uint8_t *__restrict__ packed, const int64_t *__restrict__ row_off,
                    const int32_t *__restrict__ row_samples, const int64_t *__restrict__ out_off,
                    int16_t *__restrict__ out, int32_t *__restrict__ status, uint32_t *__restrict__ agg,
                    int tiles_per_row) {
    __shared__ int s_warp[kThreads / 32];
    __shared__ __align__(16) uint32_t s_data[(2 * kTile + 64) / 4];

    const int row = blockIdx.y, tile = blockIdx.x;
    const int n = row_samples[row];
    const int t0 = tile * kTile;
    const int j = threadIdx.x;
    const int64_t row_bytes = row_off[row + 1] - row_off[row];
    const int64_t nk = ((int64_t)n + 7) >> 3;
    // the sample count is untrusted input: a row must at least hold its keys and one byte per sample
    const bool sane = n >= 0 && nk + n <= row_bytes && (int64_t)n <= (int64_t)tiles_per_row * kTile;
    if (!sane) {
        if (j == 0 && tile == 0) status[row] = 1;
        if (n <= 0 || t0 >= n) return;
    } else if (t0 >= n) {
        if (tile == 0 && j == 0 && row_bytes != 0) status[row] = 1;  // empty row with bytes
        return;
    }
    uint32_t *my_agg = agg + (size_t)row * tiles_per_row + tile;
    if (!sane) {  // later tiles of the row must not wait for this one
        if (j == 0) atomicExch(my_agg, kAggFlag);
        return;
    }
    const uint8_t *base = packed + row_off[row];
    const uint8_t *data = base + nk;
    int16_t *dst = out + out_off[row];

    // ---- 0 + 1: this thread's key word (32 samples) and its share of the row's EARLIER key bytes, all
    // loads issued before the first use (one memory round trip, not two); the earlier keys are read as
    // aligned 16-byte words with the bytes outside [0, tile * 1024) masked off
    const int64_t kb = (t0 >> 3) + 4 * j;  // first key byte of this thread's word
    uint32_t key = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
        if (kb + b < nk) key |= (uint32_t)base[kb + b] << (8 * b);
    int prior = 0;
    {
        const int n_prior = tile * (kTile / 8);                          // bytes [0, n_prior) of the row
        const int lead = (int)(reinterpret_cast<uintptr_t>(base) & 15);  // bytes before `base` in its 16-byte word
        const uint4 *w16 = reinterpret_cast<const uint4 *>(base - lead);
        const int n16 = (lead + n_prior + 15) >> 4;
        for (int k = j; k < n16; k += kThreads) {
            const uint4 v = __ldg(w16 + k);
            const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int b0 = 16 * k + 4 * e - lead;  // row byte index of this word's first byte
                uint32_t m = 0xFFFFFFFFu;
                if (b0 < 0) m = b0 <= -4 ? 0u : (m << (8 * -b0));
                if (b0 + 4 > n_prior) m = b0 >= n_prior ? 0u : (m & (0xFFFFFFFFu >> (8 * (b0 + 4 - n_prior))));
                prior += __popc(wds[e] & m);
            }
        }
    }
    const int64_t data_pos = (int64_t)t0 + cta_sum(prior, s_warp);
    const int first = t0 + 32 * j;  // first sample of this thread's word
    const int valid = n - first;    // samples of the word inside the row
    if (valid < 32) key = valid > 0 ? (key & ((1u << valid) - 1u)) : 0u;
    const int in_word = valid >= 32 ? 32 : (valid > 0 ? valid : 0);
    int tile_bytes;
    const int pref = cta_exclusive_scan(in_word + __popc(key), s_warp, &tile_bytes);
    if (tile_bytes > row_bytes - nk - data_pos) {  // truncated row: never read past it
        if (j == 0) {
            status[row] = 1;
            atomicExch(my_agg, kAggFlag);
        }
        return;
    }
    // ---- 2: stage the tile's data bytes (aligned words; the buffer is padded by the caller)
    const uint8_t *src = data + data_pos;
    const int mis = (int)(reinterpret_cast<uintptr_t>(src) & 15);
    const uint4 *src_w = reinterpret_cast<const uint4 *>(src - mis);
    const int words = (mis + tile_bytes + 15) >> 4;  // 16-byte words: <= 1025 (the caller pads the buffer)
    uint4 *s_data4 = reinterpret_cast<uint4 *>(s_data);
#pragma unroll 4
    for (int k = j; k < words; k += kThreads) s_data4[k] = __ldg(src_w + k);
    __syncthreads();
    // ---- 3: every thread decodes the 32 samples of its own key word, two at a time: the four bytes at the
    // running byte position are aligned with one funnel shift, ONE byte permute (selector picked by the two
    // key bits, a zero register supplying the absent high bytes) spreads them into two 16-bit lanes, zig-zag
    // and the running sum are done on both lanes at once (d * 0x10001 = {d0, d0 + d1}; VIADD.16x2)
    uint32_t outp[16];
    int total = 0;
    if (in_word > 0) {
        uint32_t pos = (uint32_t)(mis + pref);  // byte offset of this word's first data byte in s_data
        uint32_t kk = key, runb = 0;
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            const uint32_t *wp = s_data + (pos >> 2);
            const uint32_t a = __funnelshift_r(wp[0], wp[1], pos << 3);
            const uint32_t k2 = kk & 3u;
            kk >>= 2;
            // key bits (sample 2p, sample 2p+1) -> output bytes {lo0, hi0, lo1, hi1}; source byte 4 is zero
            // (a two-byte pick from the four 16-bit selectors held in two registers)
            const uint32_t sel = __byte_perm(0x42104140u, 0x32102140u, k2 * 0x22u + 0x10u);
            const uint32_t v = __byte_perm(a, 0u, sel);
            const uint32_t neg = (v & 0x00010001u) * 0xFFFFu;       // 0xFFFF in the lanes whose low bit is set
            const uint32_t d = ((v >> 1) & 0x7FFF7FFFu) ^ neg;      // zig-zag, both lanes
            uint32_t o;
            asm("add.u16x2 %0, %1, %2;" : "=r"(o) : "r"(d * 0x10001u), "r"(runb));
            outp[p] = o;
            runb = __byte_perm(o, 0u, 0x3232);                      // running sum = the high lane, in both lanes
            pos += 2u + __popc(k2);
        }
        total = (int)(runb & 0xFFFFu);  // meaningful for full words (a partial word ends its row)
    } else {
#pragma unroll
        for (int p = 0; p < 16; ++p) outp[p] = 0u;
    }
    // ---- 4: running sum across the threads of the tile
    int tile_sum;
    int before = cta_exclusive_scan(total, s_warp, &tile_sum);
    // ---- 5: publish this tile's aggregate, collect the predecessors' (sums are taken modulo 2^16)
    if (j == 0) {
        __threadfence();
        atomicExch(my_agg, kAggFlag | ((uint32_t)tile_sum & 0xFFFFu));
    }
    int carry = 0;
    for (int k = j; k < tile; k += kThreads) {
        const volatile uint32_t *pa = agg + (size_t)row * tiles_per_row + k;
        uint32_t v;
        while (!((v = *pa) & kAggFlag)) __nanosleep(40);  // the spinning warp leaves the issue slots to the others
        carry += (int)(v & 0xFFFFu);
    }
    if (tile > 0) before += cta_sum(carry, s_warp);
    const uint32_t bb = ((uint32_t)before & 0xFFFFu) * 0x10001u;
#pragma unroll
    for (int p = 0; p < 16; ++p) asm("add.u16x2 %0, %1, %2;" : "=r"(outp[p]) : "r"(outp[p]), "r"(bb));
    // ---- 6: store (64 contiguous bytes per thread)
    const int o0 = t0 + 32 * j;
    if (o0 + 32 <= n && ((reinterpret_cast<uintptr_t>(dst + o0) & 15) == 0)) {
        uint4 *d4 = reinterpret_cast<uint4 *>(dst + o0);
#pragma unroll
        for (int q = 0; q < 4; ++q) d4[q] = make_uint4(outp[4 * q], outp[4 * q + 1], outp[4 * q + 2], outp[4 * q + 3]);
    } else {
#pragma unroll
        for (int c = 0; c < 32; ++c)
            if (o0 + c < n) dst[o0 + c] = (int16_t)(uint16_t)((outp[c >> 1] >> (16 * (c & 1))) & 0xFFFFu);
    }
    // the row's last tile knows where the stream ends: 1 = its length disagrees with the keys
    if (j == 0 && t0 + kTile >= n && data_pos + tile_bytes + nk != row_bytes) status[row] = 1;
}

}  // namespace

size_t svb16_scratch_bytes(int n_rows, int max_row_samples) {
    const size_t tiles = (size_t)((max_row_samples + kTile - 1) / kTile);
    return (size_t)n_rows * (tiles > 0 ? tiles : 1) * sizeof(uint32_t);
}

int launch_svb16_decode(const uint8_t *packed, const int64_t *row_off, const int32_t *row_samples,
                        const int64_t *out_off, int n_rows, int max_row_samples, int16_t *out, int32_t *status,
                        void *scratch, cudaStream_t stream) {
    if (n_rows == 0) return RB200_OK;
    const int tiles = max_row_samples > 0 ? (max_row_samples + kTile - 1) / kTile : 1;
    RB200_REQUIRE(n_rows <= 65535, "at most 65535 rows per call");
    RB200_CUDA_TRY(cudaMemsetAsync(status, 0, (size_t)n_rows * sizeof(int32_t), stream));
    RB200_CUDA_TRY(cudaMemsetAsync(scratch, 0, svb16_scratch_bytes(n_rows, max_row_samples), stream));
    svb16_decode_kernel<<<dim3(tiles, n_rows), kThreads, 0, stream>>>(packed, row_off, row_samples, out_off, out,
                                                                        status, static_cast<uint32_t *>(scratch),
                                                                        tiles);
    RB200_CUDA_TRY(cudaGetLastError());
    return RB200_OK;
}

}  // namespace rb200
