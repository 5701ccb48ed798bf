// Internal declarations shared by the CUDA translation units of librb200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "remora_b200.h"

namespace rb200 {

void set_error(const char *fmt, ...);

#define RB200_CUDA_TRY(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            rb200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                             __FILE__, __LINE__);                                         \
            return RB200_ERR_CUDA;                                                        \
        }                                                                                 \
    } while (0)

#define RB200_REQUIRE(cond, ...)                                                          \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            rb200::set_error(__VA_ARGS__);                                                \
            return RB200_ERR_INVALID;                                                     \
        }                                                                                 \
    } while (0)

// ---- device math shared by all kernels ------------------------------------------------------
// swish(x) = x * sigmoid(x)   (reference src/remora/activations.py:18)
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float swishf(float x) { return x / (1.0f + expf(-x)); }
// fast variants: ex2.approx based; abs error of sigmoid < 2e-7 (measured in tests)
__device__ __forceinline__ float sigmoidf_fast(float x) {
    return __fdividef(1.0f, 1.0f + __expf(-x));
}
__device__ __forceinline__ float swishf_fast(float x) { return x * sigmoidf_fast(x); }
__device__ __forceinline__ float tanhf_fast(float x) {
    // tanh(x) = 2*sigmoid(2x) - 1
    return 2.0f * sigmoidf_fast(2.0f * x) - 1.0f;
}

// ---- growable device scratch, one per (handle, stream) ---------------------------------------
struct Workspace {
    char *base = nullptr;
    size_t bytes = 0;
    int ensure(size_t need);  // grows (synchronising free + malloc) when too small
    void release();
};

struct DebugTensor {
    std::string name;
    const float *ptr;  // into the workspace of the forward that produced it
    int B, C, T;
};

struct FusedWeights;  // rb200_fused.cu
struct TiledWeights;  // rb200_tiled.cu
struct MegaWeights;   // rb200_mega.cu
struct ConvMegaWeights;  // rb200_mega.cu

}  // namespace rb200

struct rb200_model {
    rb200_model_desc desc;
    int device = 0;
    int sm_count = 148;
    float *blob_dev = nullptr;  // canonical (PyTorch-layout, BN-folded) weights
    int64_t blob_floats = 0;
    int impl = RB200_IMPL_AUTO;
    int last_impl = 0;
    std::atomic<uint64_t> launches{0};
    bool keep_debug = false;
    std::vector<rb200::DebugTensor> debug;
    std::mutex mu;
    std::map<void *, rb200::Workspace> workspaces;  // keyed by stream
    std::map<void *, rb200::Workspace> host_staging;  // device-side input/output staging of rb200_infer_host_async
    std::map<const void *, float *> mapped_out;       // pinned output buffers -> their device address (or null)
    rb200::FusedWeights *fused = nullptr;           // non-null when the fused path applies
    rb200::TiledWeights *tiled = nullptr;           // weight layouts of the register-tiled layer kernels
    rb200::MegaWeights *mega = nullptr;             // single-kernel path (rb200_mega.cu)
    rb200::ConvMegaWeights *conv_mega = nullptr;    // Conv_w_ref single-kernel path (rb200_mega.cu)
    // pinned + device staging for rb200_infer_host
    char *pinned = nullptr;
    size_t pinned_bytes = 0;
    char *staging_dev = nullptr;
    size_t staging_bytes = 0;
    cudaStream_t host_stream = nullptr;
    // tiled layer path: the signal track runs on a side stream next to the sequence track
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int two_tracks = -1;  // -1: decide from RB200_TWO_TRACKS (default on), 0 / 1
    // per-kernel profiling of the fused path
    bool profile = false;
    std::vector<cudaEvent_t> prof_events;  // 4 per profiled forward
    float prof_ms[3] = {0.f, 0.f, 0.f};
    int prof_forwards = 0;
};

namespace rb200 {

// rb200_encode.cu
int launch_encode_dense(const int8_t *seqs, int seq_width, const int16_t *maps, int map_width,
                        const int16_t *lens, int n_chunks, int kmer_len, int T, float *out,
                        int sm_count, cudaStream_t stream, uint64_t *launches);

// rb200_layers.cu : layer-per-kernel path, any architecture / size
size_t layers_workspace_bytes(const rb200_model_desc &d, int B, int T, bool compact);
int layers_forward(rb200_model *m, Workspace &ws, const float *sigs, const float *enc,
                   const int8_t *seqs, int seq_width, const int16_t *maps, int map_width,
                   const int16_t *lens, int B, int T, float *logits, cudaStream_t stream,
                   bool tiled);
int launch_softmax_ml(const float *logits, int B, int num_out, float *probs, uint8_t *ml,
                      cudaStream_t stream);

// rb200_chunks.cu : read -> compact chunk arrays on the device
int launch_chunk_plan(const int32_t *ssm, int n_map, int sig_len, const int32_t *focus, int n, int c0,
                      int c1, int bsj, int offset, int32_t *focus_adj, int32_t *focus_sig,
                      int32_t *seq_start, int32_t *seq_len, cudaStream_t stream);
int launch_chunk_fill(const void *dacs, int dtype, int sig_len, double shift, double scale,
                      const int32_t *ssm, int n_map, const int8_t *int_seq, int n_bases,
                      const int32_t *focus_sig, const int32_t *seq_start, const int32_t *seq_len, int n,
                      int c0, int c1, int kb, int ka, int lmax, float *signal, int8_t *sequence,
                      int16_t *mapping, int16_t *lens, cudaStream_t stream);

// rb200_refine.cu : signal-mapping refinement (banded DP, one warp per read)
int launch_refine_normalise(const void *dacs, int dtype, const int64_t *sig_off, const double *shift,
                            const double *scale, int n_reads, int64_t max_len, float *out,
                            cudaStream_t stream);
size_t refine_wide_scratch_bytes(int sm_count, int near_cap, int max_band_width);
int launch_refine_dp(const float *sig, const int64_t *sig_off, const float *levels, const int32_t *band_st,
                     const int32_t *band_en, const int64_t *seq_off, const int64_t *tb_off,
                     const int32_t *order, int n_reads, const float *pen, int n_pen, int algo, int near_cap,
                     int max_band_width, int32_t *tb, int32_t *path, float *score, int32_t *status,
                     int32_t *counter, float *wide_scratch, int sm_count, cudaStream_t stream);

// rb200_vbz.cu : POD5 signal rows (svb16 + zigzag + delta) -> int16 samples
size_t svb16_scratch_bytes(int n_rows, int max_row_samples);
int launch_svb16_decode(const uint8_t *packed, const int64_t *row_off, const int32_t *row_samples,
                        const int64_t *out_off, int n_rows, int max_row_samples, int16_t *out, int32_t *status,
                        void *scratch, cudaStream_t stream);

// rb200_tiled.cu : register-tiled FFMA2 layer kernels (Conv_w_ref and every non-fused shape);
// each returns RB200_ERR_UNSUPPORTED when the layer / shape has no tiled form
int tiled_create(rb200_model *m, const float *blob_host);
void tiled_destroy(rb200_model *m);
int tiled_conv(rb200_model *m, int track, int layer, const rb200_conv_desc &c, const float *x,
               int64_t x_bstride, int t_in, float *y, int64_t y_bstride, int B, cudaStream_t stream);
bool tiled_gather_ok(const rb200_model *m, int seq_width, int map_width, int T);
int tiled_seq1_gather(rb200_model *m, const int8_t *seqs, int seq_width, const int16_t *maps,
                      int map_width, const int16_t *lens, int B, int T, float *y, int64_t y_bstride,
                      cudaStream_t stream);
int tiled_fc(rb200_model *m, const float *x, int64_t bstride, float *logits, int B,
             cudaStream_t stream);

// rb200_fused.cu : fused sm_100a kernels for ConvLSTM_w_ref size 64
bool fused_supported(const rb200_model_desc &d);
int fused_create(rb200_model *m, const float *blob_host);
void fused_destroy(rb200_model *m);
bool fused_shape_ok(const rb200_model *m, int T, int seq_width, int map_width);
size_t fused_workspace_bytes(const rb200_model *m, int B, int T);
// K0 alone: q1 [B][T-4][16] = swish(seq_conv1(enc)) for the dense interface of the single-kernel path
int fused_dense_seq1(rb200_model *m, const float *enc_dense, float *q1_out, int B, int T, cudaStream_t stream);
int fused_forward_compact(rb200_model *m, Workspace &ws, const float *sigs, const int8_t *seqs,
                          int seq_width, const int16_t *maps, int map_width, const int16_t *lens,
                          int B, int T, float *logits, cudaStream_t stream, bool want_tc,
                          const float *enc_dense = nullptr);

// rb200_mega.cu : ConvLSTM_w_ref size 64 as one kernel per batch (fp16 hi/lo split or bf16 operands)
bool mega_supported(const rb200_model_desc &d);
int mega_create(rb200_model *m, const float *blob_host);
void mega_destroy(rb200_model *m);
bool mega_shape_ok(const rb200_model *m, int T, int seq_width, int map_width);
int mega_read_flags(rb200_model *m, int *out, bool clear);
bool conv_mega_supported(const rb200_model_desc &d);
int conv_mega_create(rb200_model *m, const float *blob_host);
void conv_mega_destroy(rb200_model *m);
bool conv_mega_shape_ok(const rb200_model *m, int T, int seq_width, int map_width);
int conv_mega_forward_compact(rb200_model *m, Workspace &ws, const float *sigs, const int8_t *seqs, int seq_width,
                              const int16_t *maps, int map_width, const int16_t *lens, int B, int T, float *logits,
                              cudaStream_t stream);
struct GatherTarget {            // rb200_forward_compact_gather: where the classifier epilogue stores
    float *const *peers_dev;     // device array of n_peers buffer base pointers (peer mapped)
    int n_peers;
    long long dst_offset;        // float offset of the [B][num_out] block inside every buffer
    float *multicast_base;       // NVLS multicast alias of the buffers, or null
    long long flag_offset;       // uint32 word index of the arrival counter inside every buffer, or -1
    // deferred form (rb200_forward_compact_ship): compute CTAs store locally, one extra CTA ships an EARLIER
    // launch's block [ship_src, ship_src + ship_count) to float offset dst_offset of every other rank's buffer
    bool deferred = false;
    int self_rank = -1;
    const float *ship_src = nullptr;
    long long ship_count = 0;
};
int mega_forward_compact(rb200_model *m, Workspace &ws, const float *sigs, const int8_t *seqs, int seq_width,
                         const int16_t *maps, int map_width, const int16_t *lens, int B, int T, float *logits,
                         cudaStream_t stream, int mode, const GatherTarget *gather = nullptr,
                         const float *enc_dense = nullptr);

}  // namespace rb200
