/*
 * remora_b200 — C ABI of the B200 (sm_100a) implementation of Remora's per-chunk
 * modified-base inference hot path.
 *
 * This is the drop-in boundary.  The reference (nanoporetech/remora v3.2.0)
 * has no FFI layer of its own; the native/operator surface this library replaces is
 *
 *   - the Cython op      compute_encoded_kmer_batch          src/remora/encoded_kmers.pyx:13-45
 *   - the model call     network.forward(sigs, seqs)         models/ConvLSTM_w_ref.py:39-58
 *                                                            models/Conv_w_ref.py:44-62
 *     as reached from    RemoraRead.run_model                src/remora/data_chunks.py:516-540
 *                        inference.run_model_batched         src/remora/inference.py:277-316
 *
 * Conventions: every entry point is extern "C", takes plain pointers and sizes (no torch
 * types), returns 0 (RB200_OK) or an error code; rb200_last_error() returns a thread-local
 * message.  Pointers suffixed _dev are device pointers on the handle's device, _host are host
 * pointers.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All
 * launches are asynchronous on `stream` unless stated otherwise.  The caller owns all buffers.
 * A handle may be used from several host threads concurrently (duplex inference does this,
 * src/remora/inference.py:973-982): calls are serialised per handle while they enqueue work.
 *
 * There is NO CPU fallback anywhere behind this interface.
 */
#ifndef REMORA_B200_H
#define REMORA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB200_ABI_VERSION 1

enum rb200_status {
    RB200_OK = 0,
    RB200_ERR_INVALID = 1,     /* bad argument / shape */
    RB200_ERR_CUDA = 2,        /* a CUDA runtime call failed */
    RB200_ERR_UNSUPPORTED = 3, /* valid request this build has no kernel for */
    RB200_ERR_NOMEM = 4
};

enum rb200_arch {
    RB200_ARCH_CONVLSTM_W_REF = 1, /* models/ConvLSTM_w_ref.py */
    RB200_ARCH_CONV_W_REF = 2      /* models/Conv_w_ref.py */
};

enum rb200_impl {
    RB200_IMPL_AUTO = 0,    /* fused kernels when the model/shape qualifies, else tiled layer kernels */
    RB200_IMPL_LAYERS = 1,  /* one plain CUDA kernel per layer, any size / kmer_len / chunk_len */
    RB200_IMPL_FUSED = 2,   /* fused sm_100a kernels (ConvLSTM_w_ref, size 64), fp32 FFMA2; error if n/a */
    RB200_IMPL_FUSED_TC = 3, /* same, merge conv + LSTM input projection on tcgen05 (3xTF32, TMEM) */
    RB200_IMPL_TILED = 4, /* per layer: register-tiled FFMA2 convolutions, gather-form seq_conv1 from the
                             compact arrays; Conv_w_ref and every shape the fused kernels do not take */
    RB200_IMPL_FUSED_MEGA = 5, /* ConvLSTM_w_ref size 64, chunk_len <= 100: ONE kernel per batch, every
                                  GEMM-shaped layer on tcgen05 with fp16 hi/lo split operands (three
                                  products, fp32 TMEM accumulators): fp32 parity (1e-4).  AUTO picks it */
    RB200_IMPL_FUSED_BF16 = 6  /* the same kernel with single-pass bf16 operands (BASELINE configs[1]
                                  "bf16"): fp32 accumulate and gates, its own tolerance; never picked by AUTO */
};

#define RB200_MAX_CONVS 4

/* One Conv1d(+eval BatchNorm folded)+swish layer.  Weights are float32 [c_out][c_in][kw] with
 * the BatchNorm already folded in (W' = W*g/sqrt(var+eps), b' = (b-mean)*g/sqrt(var+eps)+beta;
 * same fold as torch fuse_conv_bn_eval used by the reference at src/remora/model_util.py:216).
 * Offsets count floats from the start of the weight blob. */
typedef struct rb200_conv_desc {
    int32_t c_in, c_out, kw, stride;
    int64_t w_off, b_off;
} rb200_conv_desc;

typedef struct rb200_model_desc {
    int32_t struct_size; /* = sizeof(rb200_model_desc), ABI guard */
    int32_t arch;        /* enum rb200_arch */
    int32_t size;        /* channel width ("size" model parameter) */
    int32_t kmer_len;    /* sequence track has 4*kmer_len rows */
    int32_t num_out;     /* classifier outputs */
    int32_t n_sig_conv, n_seq_conv, n_merge_conv;
    rb200_conv_desc sig_conv[RB200_MAX_CONVS];
    rb200_conv_desc seq_conv[RB200_MAX_CONVS];
    rb200_conv_desc merge_conv[RB200_MAX_CONVS];
    int32_t n_lstm; /* 2 for ConvLSTM_w_ref, 0 for Conv_w_ref */
    /* torch.nn.LSTM layout: w_ih [4H][H], w_hh [4H][H], gate order i,f,g,o;
     * b = bias_ih + bias_hh [4H] (scripts/convert_ts_to_ont_json.py:134-135) */
    int64_t lstm_w_ih_off[2], lstm_w_hh_off[2], lstm_b_off[2];
    int32_t fc_in; /* Linear in_features (size, or size*T_final for Conv_w_ref) */
    int32_t reserved;
    int64_t fc_w_off, fc_b_off; /* [num_out][fc_in], [num_out] */
} rb200_model_desc;

typedef struct rb200_model *rb200_handle;

int rb200_version(void);
const char *rb200_last_error(void);

/* Uploads the (host) weight blob to `device`, builds the kernel-specific weight layouts.
 * Replaces: torch.jit.load + module.to(device) at src/remora/model_util.py:468-481. */
int rb200_create(const rb200_model_desc *desc, const float *weights_host, int64_t n_floats,
                 int device, rb200_handle *out);
int rb200_destroy(rb200_handle h);

/* Test / benchmarking controls. */
int rb200_set_impl(rb200_handle h, int impl); /* enum rb200_impl */
int rb200_last_impl(rb200_handle h);          /* impl used by the last forward on this handle */
uint64_t rb200_launch_count(rb200_handle h);  /* kernels launched through this handle so far */
int rb200_set_debug(rb200_handle h, int keep); /* keep layer outputs of the next LAYERS forward */
/* Copies a kept layer output (float32, [B][C][T] channel-first) to dst_dev.  Names: sig1 sig2
 * sig3 seq1 seq2 seq3 merge1..merge4 lstm1 lstm2.  *n_floats receives the element count. */
int rb200_debug_tensor(rb200_handle h, const char *name, float *dst_dev, int64_t capacity,
                       int64_t *n_floats, int32_t *channels, int32_t *steps, void *stream);

/* Sticky diagnostics of the fp16-split single-kernel path: *flags bit 0 is set once an intermediate
 * activation left the fp16 range (|x| >= 65504; it was saturated, results of that call are not to be
 * trusted - rerun with RB200_IMPL_FUSED_TC, whose operands are fp32-ranged).  Synchronises the device;
 * clear != 0 resets the flags. */
int rb200_get_flags(rb200_handle h, int32_t *flags, int clear);

/* Per-kernel device timing of the fused path (CUDA events recorded on the caller's stream between
 * K1/K2/K3).  rb200_set_profile(h, 1) enables it and zeroes the accumulators;
 * rb200_get_profile synchronises the recorded events and returns the accumulated milliseconds of
 * {K1 front, K2 merge+projection, K3 lstm} and the number of forwards measured. */
int rb200_set_profile(rb200_handle h, int on);
int rb200_get_profile(rb200_handle h, float ms_out[3], int32_t *n_forwards);

/* Dense k-mer one-hot encoding, bit-exact with the reference's Cython op
 *   compute_encoded_kmer_batch(before, after, seqs, seq_mappings, seq_lens)
 *   (src/remora/encoded_kmers.pyx:13-45): out[c, 4*p+base, map[c,s]:map[c,s+1]] = 1.0f for
 * every k-mer offset p < before+after+1 and s < seq_lens[c] with base = seqs[c, s+p] != -1;
 * everything else 0.0f.  out_dev: float32 [n_chunks][4*(before+after+1)][sig_len].
 * sig_len is passed explicitly (the reference reads it from chunk 0, pyx:23); mapping entries are
 * clipped to [0, sig_len].  Never reads seqs past s+p < seq_len+kmer_len-1 nor maps past seq_len. */
int rb200_encode_dense(const int8_t *seqs_dev, int32_t seq_width, const int16_t *maps_dev,
                       int32_t map_width, const int16_t *lens_dev, int32_t n_chunks,
                       int32_t before, int32_t after, int32_t sig_len, float *out_dev,
                       void *stream);

/* model(sigs, enc_kmers): the reference model call with its dense arguments
 * (sigs float32 [B][1][T], enc float32 [B][4*kmer_len][T]) -> logits float32 [B][num_out].
 * Replaces network.forward at models/ConvLSTM_w_ref.py:39-58 / models/Conv_w_ref.py:44-62. */
int rb200_forward_dense(rb200_handle h, const float *sigs_dev, const float *enc_dev, int32_t B,
                        int32_t T, float *logits_dev, void *stream);

/* Fused encode + forward on the reference's compact chunk arrays
 * (CoreRemoraDataset._core_dtypes, src/remora/data_chunks.py:942-948): the one-hot tensor is never
 * materialised.  Precondition (true for every chunk the reference produces): each mapping row
 * is non-decreasing on [0, seq_len] with map[0] == 0 and map[seq_len] == T. */
int rb200_forward_compact(rb200_handle h, const float *sigs_dev, const int8_t *seqs_dev,
                          int32_t seq_width, const int16_t *maps_dev, int32_t map_width,
                          const int16_t *lens_dev, int32_t B, int32_t T, float *logits_dev,
                          void *stream);

/* Fused compute + exchange for the multi-GPU form of the path: rb200_forward_compact whose
 * classifier epilogue stores every chunk's logits straight into the logits buffer of EVERY rank over
 * NVLink / NVSwitch, instead of a separate collective after the kernel (the reference has no multi-device
 * mode; its order contract for gathered batches is src/remora/inference.py:331-367).
 *   peer_bases_dev : device array of n_peers float* (n_peers <= 32): the base of each rank's buffer as mapped
 *                    into THIS process (CUDA IPC / torch symmetric memory `buffer_ptrs_dev`), own rank included
 *   dst_offset     : float offset inside every buffer of this call's [B][num_out] block
 *   multicast_base : NVLS multicast alias of the same buffers (symmetric memory `multicast_ptr`), or NULL;
 *                    when given, ONE multimem.st per value reaches all ranks through the switch
 *   flag_word      : >= 0: index (in 4-byte words from the buffer base) of a uint32 arrival counter that
 *                    every CTA increments on every rank after its stores are visible system-wide; the block
 *                    is complete on a rank when the counter has advanced by ceil(B / 4).  -1: no counter
 *   logits_dev     : optional local copy of the logits (may be NULL)
 * Only the single-kernel path provides it (RB200_ERR_UNSUPPORTED otherwise: use rb200_forward_compact and a
 * collective). */
int rb200_forward_compact_gather(rb200_handle h, const float *sigs_dev, const int8_t *seqs_dev,
                                 int32_t seq_width, const int16_t *maps_dev, int32_t map_width,
                                 const int16_t *lens_dev, int32_t B, int32_t T, float *logits_dev,
                                 void *const *peer_bases_dev, int32_t n_peers, int64_t dst_offset,
                                 void *multicast_base, int64_t flag_word, void *stream);

/* The same exchange one step behind the compute, so that no compute thread block ever waits on a remote
 * store: the kernel's blocks write this call's logits to logits_dev only (this rank's own block of its ring),
 * and ONE extra thread block of the same launch waits for the previous launch on the stream
 * (griddepcontrol.wait) and ships the block an EARLIER call produced - ship_src_dev[0 .. ship_count) - to float
 * offset ship_dst_offset of every other rank's buffer (peer stores, or multimem.st through multicast_base).
 * flag_word >= 0: a uint32 arrival counter bumped by ONE on every rank after the shipped block is visible.
 * B == 0 ships only (the flush after the last step).  The caller keeps track of what is pending
 * (remora_b200.parallel.PeerLogitRing).  ship_src_dev, ship_dst_offset must be 16-byte aligned. */
int rb200_forward_compact_ship(rb200_handle h, const float *sigs_dev, const int8_t *seqs_dev,
                               int32_t seq_width, const int16_t *maps_dev, int32_t map_width,
                               const int16_t *lens_dev, int32_t B, int32_t T, float *logits_dev,
                               void *const *peer_bases_dev, int32_t n_peers, int32_t self_rank,
                               const float *ship_src_dev, int64_t ship_dst_offset, int64_t ship_count,
                               void *multicast_base, int64_t flag_word, void *stream);

/* End-to-end convenience for host callers (the bench's e2e leg and non-torch hosts): copies the
 * compact arrays host->device through pinned staging, runs rb200_forward_compact, copies the
 * logits back and synchronises.  Host buffers may be pageable. */
int rb200_infer_host(rb200_handle h, const float *sigs_host, const int8_t *seqs_host,
                     int32_t seq_width, const int16_t *maps_host, int32_t map_width,
                     const int16_t *lens_host, int32_t B, int32_t T, float *logits_host);

/* Asynchronous form for pipelined callers: every buffer is a PINNED host buffer owned by the caller
 * (cudaHostAlloc / torch pin_memory); the call enqueues H2D copies of the compact arrays, the kernels
 * and the D2H copy of the logits on `stream` and returns immediately.  The caller synchronises the
 * stream (or an event) before reading logits_pinned, and must not touch the input buffers until then.
 * Alternating two streams overlaps step i+1's copies with step i's kernels, which is how the
 * reference's queue-based pipeline (src/remora/inference.py:488-572) keeps its device busy.  Device
 * staging is per (handle, stream). */
int rb200_infer_host_async(rb200_handle h, const float *sigs_pinned, const int8_t *seqs_pinned,
                           int32_t seq_width, const int16_t *maps_pinned, int32_t map_width,
                           const int16_t *lens_pinned, int32_t B, int32_t T, float *logits_pinned,
                           void *stream);

/* On-GPU chunk extraction ("next" row 1): the per-read host loop of
 * RemoraRead.iter_chunks / extract_chunk (src/remora/data_chunks.py:331-466) and the signal
 * normalisation (:191-197) as two kernels.  All pointers are device pointers.
 *   plan: per focus base -> adjusted focus base, chunk centre sample, first overlapping base, number
 *         of overlapping bases (the caller reads min/max of seq_len to size the arrays and to reject
 *         reads with an empty chunk, as the reference's Chunk.check would).
 *   fill: writes signal f32 [n][c0+c1], sequence i8 [n][lmax+kb+ka], mapping i16 [n][lmax+1],
 *         lens i16 [n] - the layout rb200_forward_compact consumes.
 * dacs_dtype: 0 = int16, 1 = float32, 2 = float64 (normalisation follows numpy's precision rules for
 * that dtype so the float32 signal is bit-identical to the reference's). */
int rb200_chunk_plan(const int32_t *seq_to_sig_map_dev, int32_t n_map, int32_t sig_len,
                     const int32_t *focus_bases_dev, int32_t n, int32_t chunk_before, int32_t chunk_after,
                     int32_t base_start_justify, int32_t offset, int32_t *focus_adj_dev,
                     int32_t *focus_sig_dev, int32_t *seq_start_dev, int32_t *seq_len_dev, void *stream);
int rb200_chunk_fill(const void *dacs_dev, int32_t dacs_dtype, int32_t sig_len, double shift, double scale,
                     const int32_t *seq_to_sig_map_dev, int32_t n_map, const int8_t *int_seq_dev,
                     int32_t n_bases, const int32_t *focus_sig_dev, const int32_t *seq_start_dev,
                     const int32_t *seq_len_dev, int32_t n, int32_t chunk_before, int32_t chunk_after,
                     int32_t kmer_before, int32_t kmer_after, int32_t lmax, float *signal_dev,
                     int8_t *sequence_dev, int16_t *mapping_dev, int16_t *lens_dev, void *stream);

/* Post-processing on device ("next" row 2): softmax over num_out, drop class 0,
 * probs float32 [B][num_out-1] (may be NULL) and ML bytes uint8 [B][num_out-1]
 * = min(floor(p*256), 255)  (src/remora/util.py:182-186, 532-535). */
int rb200_softmax_ml(const float *logits_dev, int32_t B, int32_t num_out, float *probs_dev,
                     uint8_t *ml_dev, void *stream);

/* Signal-mapping refinement on the device ("next" row 4): the banded dynamic programme
 * of SigMapRefiner.refine_sig_map / refine_signal_mapping (src/remora/refine_signal_map.py:474-499,
 * 783-840) with the Cython core seq_banded_dp (src/remora/refine_signal_map_core.pyx:403-473:
 * banded_forward_dp + banded_traceback, both "Viterbi" and "dwell_penalty" steps) for a batch of reads.
 * Scores, traceback and the returned path are bit-identical to the reference's.
 *
 * Layout: reads are concatenated.  Read r owns samples [sig_off[r], sig_off[r+1]) of the signal
 * (already trimmed to seq_to_sig_map[0]..seq_to_sig_map[-1]), bases [seq_off[r], seq_off[r+1]) of
 * levels / band_start / band_end (the seq band after adjust_seq_band, signal coordinates relative to
 * the read: start[0] == 0, both strictly increasing, start[b] <= end[b-1], end[last] == read length),
 * traceback words [tb_off[r], tb_off[r+1]) with tb_off[r+1]-tb_off[r] = sum(end-start), and
 * path entries [seq_off[r] + r, seq_off[r+1] + r + 1) (seq_len + 1 values: path[0] = 0, path[b] =
 * first sample of base b, path[seq_len] = read length).  order (may be NULL) = processing order,
 * longest reads first for load balance.
 *
 *   rb200_refine_normalize: signal = (dacs - shift[r]) / scale[r] as float32; dacs_dtype 0 = int16,
 *     1 = float32 (float32 arithmetic), 2 = float64, 3 = float32 samples in float64 arithmetic (numpy
 *     float64 shift/scale), following numpy's promotion so the result matches the reference's bits.
 *   near_cap: capacity (samples) of the per-warp rows kept in shared memory, 64..1024, 0 = chosen from
 *     max_band_width; bases whose band is wider take their rows from wide_scratch_dev.  A smaller
 *     capacity lets more warps (reads) share an SM.  max_band_width: widest band of the batch.
 *   rb200_refine_scratch_bytes: size of wide_scratch_dev (0 when every band fits near_cap).
 *   rb200_refine_dp: algo 0 = "Viterbi", 1 = "dwell_penalty" (penalties on the host, <= 16).
 *     status[r] = 1 when the traceback of read r left its band (undefined behaviour in the reference;
 *     the path of that read is not usable).  queue_dev: one int32 of device scratch. */
int rb200_refine_normalize(const void *dacs_dev, int32_t dacs_dtype, const int64_t *sig_off_dev,
                           const double *shift_dev, const double *scale_dev, int32_t n_reads,
                           int64_t max_len, float *signal_dev, void *stream);
int rb200_refine_scratch_bytes(int32_t near_cap, int32_t max_band_width, int64_t *bytes);
int rb200_refine_dp(const float *signal_dev, const int64_t *sig_off_dev, const float *levels_dev,
                    const int32_t *band_start_dev, const int32_t *band_end_dev, const int64_t *seq_off_dev,
                    const int64_t *tb_off_dev, const int32_t *order_dev, int32_t n_reads,
                    const float *dwell_penalty_host, int32_t n_penalty, int32_t algo, int32_t near_cap,
                    int32_t max_band_width, int32_t *traceback_ws_dev, int32_t *path_dev, float *score_dev,
                    int32_t *status_dev, int32_t *queue_dev, float *wide_scratch_dev, void *stream);

/* POD5 signal decode ("next" row 3): the svb16 + zig-zag + delta layers of the
 * "minknow.vbz" signal rows (the zstd layer is undone on the host) for a batch of rows.  The reference
 * gets these samples from the pod5 package (pod5.ReadRecord.signal, src/remora/io.py:455-462).
 * Row i occupies bytes [row_off[i], row_off[i+1]) of packed_dev: ceil(n/8) key bytes (bit k%8 of byte
 * k/8 set = value k takes two bytes) then the little-endian data bytes; it decodes to row_samples[i]
 * int16 samples written at out_dev + out_off[i] (offsets that are multiples of 8 samples get 16-byte
 * stores).  packed_dev must be readable 16 bytes past row_off[n_rows].  max_row_samples >= every
 * row_samples[i] sizes the launch: one thread block per 8192-sample tile of every row, so rows are decoded
 * in parallel WITHIN a row too; scratch_dev (rb200_svb16_scratch_bytes, 4 B per tile) carries the running
 * sums between the tiles of a row.  status[i] = 1 when the stream length disagrees with the keys or the
 * sample count cannot fit the row's bytes (corrupt row; nothing is read past the row). */
int rb200_svb16_scratch_bytes(int32_t n_rows, int32_t max_row_samples, int64_t *bytes);
int rb200_svb16_decode(const uint8_t *packed_dev, const int64_t *row_off_dev, const int32_t *row_samples_dev,
                       const int64_t *out_off_dev, int32_t n_rows, int32_t max_row_samples, int16_t *out_dev,
                       int32_t *status_dev, void *scratch_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* REMORA_B200_H */
