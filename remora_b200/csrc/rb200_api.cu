// C-ABI entry points (include/remora_b200.h): handle lifetime, argument checks, dispatch between
// the fused sm_100a kernels and the layer-per-kernel path, host-buffer convenience call.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "rb200_internal.cuh"

namespace rb200 {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int Workspace::ensure(size_t need) {
    if (need <= bytes) return RB200_OK;
    if (base) {
        cudaDeviceSynchronize();  // earlier launches may still read the old buffer
        cudaFree(base);
        base = nullptr;
        bytes = 0;
    }
    size_t want = need + need / 4;
    cudaError_t e = cudaMalloc(&base, want);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        base = nullptr;
        return RB200_ERR_NOMEM;
    }
    bytes = want;
    return RB200_OK;
}

void Workspace::release() {
    if (base) cudaFree(base);
    base = nullptr;
    bytes = 0;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
        if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

static bool conv_in_blob(const rb200_conv_desc &c, int64_t n) {
    if (c.c_in <= 0 || c.c_out <= 0 || c.kw <= 0 || c.stride <= 0) return false;
    if (c.w_off < 0 || c.b_off < 0) return false;
    return c.w_off + (int64_t)c.c_in * c.c_out * c.kw <= n && c.b_off + c.c_out <= n;
}

static int check_desc(const rb200_model_desc *d, int64_t n) {
    RB200_REQUIRE(d->struct_size == (int32_t)sizeof(rb200_model_desc),
                  "rb200_model_desc size mismatch (%d vs %zu): header/library ABI skew",
                  d->struct_size, sizeof(rb200_model_desc));
    RB200_REQUIRE(d->arch == RB200_ARCH_CONVLSTM_W_REF || d->arch == RB200_ARCH_CONV_W_REF,
                  "unknown architecture %d", d->arch);
    RB200_REQUIRE(d->size > 0 && d->kmer_len > 0 && d->num_out > 0, "bad model dimensions");
    RB200_REQUIRE(d->n_sig_conv >= 1 && d->n_sig_conv <= RB200_MAX_CONVS && d->n_seq_conv >= 1 &&
                      d->n_seq_conv <= RB200_MAX_CONVS && d->n_merge_conv >= 1 &&
                      d->n_merge_conv <= RB200_MAX_CONVS,
                  "bad conv layer counts");
    for (int i = 0; i < d->n_sig_conv; ++i)
        RB200_REQUIRE(conv_in_blob(d->sig_conv[i], n), "sig_conv%d outside weight blob", i + 1);
    for (int i = 0; i < d->n_seq_conv; ++i)
        RB200_REQUIRE(conv_in_blob(d->seq_conv[i], n), "seq_conv%d outside weight blob", i + 1);
    for (int i = 0; i < d->n_merge_conv; ++i)
        RB200_REQUIRE(conv_in_blob(d->merge_conv[i], n), "merge_conv%d outside weight blob", i + 1);
    RB200_REQUIRE(d->sig_conv[0].c_in == 1, "signal track must have 1 input channel");
    RB200_REQUIRE(d->seq_conv[0].c_in == 4 * d->kmer_len, "seq_conv1 c_in != 4*kmer_len");
    RB200_REQUIRE(d->sig_conv[d->n_sig_conv - 1].c_out == d->size &&
                      d->seq_conv[d->n_seq_conv - 1].c_out == d->size &&
                      d->merge_conv[0].c_in == 2 * d->size,
                  "track widths do not match the merge convolution");
    for (int i = 1; i < d->n_sig_conv; ++i)
        RB200_REQUIRE(d->sig_conv[i].c_in == d->sig_conv[i - 1].c_out, "sig conv chain mismatch");
    for (int i = 1; i < d->n_seq_conv; ++i)
        RB200_REQUIRE(d->seq_conv[i].c_in == d->seq_conv[i - 1].c_out, "seq conv chain mismatch");
    for (int i = 1; i < d->n_merge_conv; ++i)
        RB200_REQUIRE(d->merge_conv[i].c_in == d->merge_conv[i - 1].c_out, "merge chain mismatch");
    if (d->arch == RB200_ARCH_CONVLSTM_W_REF) {
        RB200_REQUIRE(d->n_lstm == 2, "ConvLSTM_w_ref needs 2 LSTM layers");
        RB200_REQUIRE(d->merge_conv[d->n_merge_conv - 1].c_out == d->size && d->fc_in == d->size,
                      "LSTM width mismatch");
        const int64_t H = d->size;
        for (int l = 0; l < 2; ++l)
            RB200_REQUIRE(d->lstm_w_ih_off[l] >= 0 && d->lstm_w_ih_off[l] + 4 * H * H <= n &&
                              d->lstm_w_hh_off[l] >= 0 && d->lstm_w_hh_off[l] + 4 * H * H <= n &&
                              d->lstm_b_off[l] >= 0 && d->lstm_b_off[l] + 4 * H <= n,
                          "lstm%d outside weight blob", l + 1);
    } else {
        RB200_REQUIRE(d->n_lstm == 0, "Conv_w_ref has no LSTM");
    }
    RB200_REQUIRE(d->fc_in > 0 && d->fc_w_off >= 0 &&
                      d->fc_w_off + (int64_t)d->fc_in * d->num_out <= n && d->fc_b_off >= 0 &&
                      d->fc_b_off + d->num_out <= n,
                  "fc outside weight blob");
    return RB200_OK;
}

static int forward_common(rb200_model *m, const float *sigs, const float *enc, const int8_t *seqs,
                          int seq_width, const int16_t *maps, int map_width, const int16_t *lens,
                          int B, int T, float *logits, void *stream_v) {
    RB200_REQUIRE(m != nullptr, "null handle");
    RB200_REQUIRE(B >= 0 && T > 0, "bad batch (%d) / chunk_len (%d)", B, T);
    if (B == 0) return RB200_OK;
    RB200_REQUIRE(sigs && logits, "null buffer");
    const bool compact = enc == nullptr;
    if (compact) {
        RB200_REQUIRE(seqs && maps && lens, "null compact input");
        RB200_REQUIRE(map_width >= 2 && seq_width >= m->desc.kmer_len,
                      "compact arrays too narrow (seq_width %d, map_width %d)", seq_width, map_width);
        RB200_REQUIRE(T < 32768, "chunk_len %d does not fit the int16 mapping", T);
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    DeviceGuard guard(m->device);
    RB200_REQUIRE(guard.ok, "cannot select device %d", m->device);
    std::lock_guard<std::mutex> lock(m->mu);
    Workspace &ws = m->workspaces[stream_v];
    int impl = m->impl;
    const bool fused_ok =
        m->fused != nullptr && fused_shape_ok(m, T, compact ? seq_width : m->desc.kmer_len,
                                              compact ? map_width : 2);
    // AUTO: the single-kernel path when the shape qualifies (compact arrays, chunk_len <= 100), else the
    // three-kernel tensor-core path (falls back to FFMA2 inside when the CTA's rows do not fit two M
    // tiles), else the tiled layer kernels
    const bool conv_mega_ok = compact && conv_mega_shape_ok(m, T, seq_width, map_width);
    // the dense interface reaches the single kernel through the dense seq_conv1 kernel (K0)
    const bool mega_ok = compact ? mega_shape_ok(m, T, seq_width, map_width)
                                 : (m->fused != nullptr && mega_shape_ok(m, T, m->desc.kmer_len, 2));
    if (impl == RB200_IMPL_AUTO)
        impl = (mega_ok || conv_mega_ok) ? RB200_IMPL_FUSED_MEGA : fused_ok ? RB200_IMPL_FUSED_TC : RB200_IMPL_TILED;
    if (impl == RB200_IMPL_FUSED_MEGA && conv_mega_ok)  // Conv_w_ref: its own single kernel
        return conv_mega_forward_compact(m, ws, sigs, seqs, seq_width, maps, map_width, lens, B, T, logits, stream);
    if (impl == RB200_IMPL_FUSED_MEGA || impl == RB200_IMPL_FUSED_BF16) {
        if (!mega_ok) {
            set_error("single-kernel path not available for this model/shape/input form");
            return RB200_ERR_UNSUPPORTED;
        }
        return mega_forward_compact(m, ws, sigs, seqs, seq_width, maps, map_width, lens, B, T, logits, stream,
                                    impl == RB200_IMPL_FUSED_BF16 ? 1 : 0, nullptr, compact ? nullptr : enc);
    }
    if (impl == RB200_IMPL_FUSED || impl == RB200_IMPL_FUSED_TC) {
        if (!fused_ok) {
            set_error("fused kernels not available for this model/shape/input form");
            return RB200_ERR_UNSUPPORTED;
        }
        if (compact)
            return fused_forward_compact(m, ws, sigs, seqs, seq_width, maps, map_width, lens, B, T,
                                         logits, stream, impl == RB200_IMPL_FUSED_TC);
        // dense interface (the reference's model(sigs, enc_kmers)): dense seq_conv1 kernel + the
        // tensor-core fused kernels; shapes they cannot take fall through to the layer kernels
        if (impl == RB200_IMPL_FUSED_TC) {
            int rc = fused_forward_compact(m, ws, sigs, nullptr, 0, nullptr, 0, nullptr, B, T, logits,
                                           stream, true, enc);
            if (rc != RB200_ERR_UNSUPPORTED) return rc;
        }
        if (m->impl != RB200_IMPL_AUTO) {
            set_error("fused kernels not available for the dense interface with this shape");
            return RB200_ERR_UNSUPPORTED;
        }
    }
    // layer by layer: the register-tiled FFMA2 kernels (layers without a tiled form use the plain
    // kernel), or - RB200_IMPL_LAYERS - the plain one-thread-per-output kernels throughout
    const bool tiled = impl != RB200_IMPL_LAYERS;
    m->last_impl = tiled ? RB200_IMPL_TILED : RB200_IMPL_LAYERS;
    return layers_forward(m, ws, sigs, enc, seqs, seq_width, maps, map_width, lens, B, T, logits,
                          stream, tiled);
}

}  // namespace rb200

using namespace rb200;

extern "C" {

int rb200_version(void) { return RB200_ABI_VERSION; }

const char *rb200_last_error(void) { return g_err; }

int rb200_create(const rb200_model_desc *desc, const float *weights_host, int64_t n_floats,
                 int device, rb200_handle *out) {
    RB200_REQUIRE(desc && weights_host && out && n_floats > 0, "null argument");
    int rc = check_desc(desc, n_floats);
    if (rc) return rc;
    int n_dev = 0;
    RB200_CUDA_TRY(cudaGetDeviceCount(&n_dev));
    RB200_REQUIRE(device >= 0 && device < n_dev, "device %d out of range (%d visible)", device,
                  n_dev);
    DeviceGuard guard(device);
    RB200_REQUIRE(guard.ok, "cannot select device %d", device);
    cudaDeviceProp prop;
    RB200_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                  prop.major, prop.minor);
        return RB200_ERR_UNSUPPORTED;
    }
    rb200_model *m = new rb200_model();
    m->desc = *desc;
    m->device = device;
    m->sm_count = prop.multiProcessorCount;
    m->blob_floats = n_floats;
    cudaError_t e = cudaMalloc(&m->blob_dev, n_floats * sizeof(float));
    if (e == cudaSuccess)
        e = cudaMemcpy(m->blob_dev, weights_host, n_floats * sizeof(float),
                       cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error("weight upload failed: %s", cudaGetErrorString(e));
        if (m->blob_dev) cudaFree(m->blob_dev);
        delete m;
        return RB200_ERR_CUDA;
    }
    rc = tiled_create(m, weights_host);
    if (rc == RB200_OK && fused_supported(m->desc)) rc = fused_create(m, weights_host);
    if (rc == RB200_OK && mega_supported(m->desc)) rc = mega_create(m, weights_host);
    if (rc == RB200_OK && conv_mega_supported(m->desc)) rc = conv_mega_create(m, weights_host);
    if (rc) {
        fused_destroy(m);
        mega_destroy(m);
        conv_mega_destroy(m);
        tiled_destroy(m);
        cudaFree(m->blob_dev);
        delete m;
        return rc;
    }
    *out = m;
    return RB200_OK;
}

int rb200_destroy(rb200_handle h) {
    if (!h) return RB200_OK;
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
    fused_destroy(h);
    mega_destroy(h);
    conv_mega_destroy(h);
    tiled_destroy(h);
    for (auto &kv : h->workspaces) kv.second.release();
    for (auto &kv : h->host_staging) kv.second.release();
    if (h->blob_dev) cudaFree(h->blob_dev);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->staging_dev) cudaFree(h->staging_dev);
    if (h->host_stream) cudaStreamDestroy(h->host_stream);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    delete h;
    return RB200_OK;
}

int rb200_set_impl(rb200_handle h, int impl) {
    RB200_REQUIRE(h && impl >= RB200_IMPL_AUTO && impl <= RB200_IMPL_FUSED_BF16, "bad argument");
    if ((impl == RB200_IMPL_FUSED || impl == RB200_IMPL_FUSED_TC) && h->fused == nullptr) {
        set_error("fused kernels not available for this model");
        return RB200_ERR_UNSUPPORTED;
    }
    if ((impl == RB200_IMPL_FUSED_BF16 && h->mega == nullptr) ||
        (impl == RB200_IMPL_FUSED_MEGA && h->mega == nullptr && h->conv_mega == nullptr)) {
        set_error("single-kernel path not available for this model");
        return RB200_ERR_UNSUPPORTED;
    }
    h->impl = impl;
    return RB200_OK;
}

int rb200_last_impl(rb200_handle h) { return h ? h->last_impl : 0; }

int rb200_get_flags(rb200_handle h, int32_t *flags, int clear) {
    RB200_REQUIRE(h && flags, "null argument");
    DeviceGuard guard(h->device);
    std::lock_guard<std::mutex> lock(h->mu);
    int v = 0;
    int rc = mega_read_flags(h, &v, clear != 0);
    *flags = v;
    return rc;
}

uint64_t rb200_launch_count(rb200_handle h) { return h ? h->launches.load() : 0; }

int rb200_set_debug(rb200_handle h, int keep) {
    RB200_REQUIRE(h, "null handle");
    h->keep_debug = keep != 0;
    if (!keep) h->debug.clear();
    return RB200_OK;
}

int rb200_debug_tensor(rb200_handle h, const char *name, float *dst_dev, int64_t capacity,
                       int64_t *n_floats, int32_t *channels, int32_t *steps, void *stream) {
    RB200_REQUIRE(h && name && n_floats, "null argument");
    std::lock_guard<std::mutex> lock(h->mu);
    for (const auto &t : h->debug) {
        if (t.name == name) {
            const int64_t n = (int64_t)t.B * t.C * t.T;
            *n_floats = n;
            if (channels) *channels = t.C;
            if (steps) *steps = t.T;
            if (dst_dev) {
                RB200_REQUIRE(capacity >= n, "debug tensor %s needs %lld floats", name,
                              (long long)n);
                DeviceGuard guard(h->device);
                RB200_CUDA_TRY(cudaMemcpyAsync(dst_dev, t.ptr, n * sizeof(float),
                                               cudaMemcpyDeviceToDevice,
                                               static_cast<cudaStream_t>(stream)));
            }
            return RB200_OK;
        }
    }
    set_error("no kept tensor named %s (enable rb200_set_debug and run a LAYERS forward)", name);
    return RB200_ERR_INVALID;
}

static int drain_profile(rb200_model *h) {
    for (size_t i = 0; i + 3 < h->prof_events.size(); i += 4) {
        RB200_CUDA_TRY(cudaEventSynchronize(h->prof_events[i + 3]));
        for (int k = 0; k < 3; ++k) {
            float ms = 0.f;
            RB200_CUDA_TRY(cudaEventElapsedTime(&ms, h->prof_events[i + k], h->prof_events[i + k + 1]));
            h->prof_ms[k] += ms;
        }
        h->prof_forwards++;
    }
    for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
    h->prof_events.clear();
    return RB200_OK;
}

int rb200_set_profile(rb200_handle h, int on) {
    RB200_REQUIRE(h, "null handle");
    DeviceGuard guard(h->device);
    std::lock_guard<std::mutex> lock(h->mu);
    int rc = drain_profile(h);
    if (rc) return rc;
    h->profile = on != 0;
    h->prof_ms[0] = h->prof_ms[1] = h->prof_ms[2] = 0.f;
    h->prof_forwards = 0;
    return RB200_OK;
}

int rb200_get_profile(rb200_handle h, float ms_out[3], int32_t *n_forwards) {
    RB200_REQUIRE(h && ms_out, "null argument");
    DeviceGuard guard(h->device);
    std::lock_guard<std::mutex> lock(h->mu);
    int rc = drain_profile(h);
    if (rc) return rc;
    for (int k = 0; k < 3; ++k) ms_out[k] = h->prof_ms[k];
    if (n_forwards) *n_forwards = h->prof_forwards;
    return RB200_OK;
}

int rb200_encode_dense(const int8_t *seqs_dev, int32_t seq_width, const int16_t *maps_dev,
                       int32_t map_width, const int16_t *lens_dev, int32_t n_chunks,
                       int32_t before, int32_t after, int32_t sig_len, float *out_dev,
                       void *stream) {
    RB200_REQUIRE(n_chunks >= 0 && before >= 0 && after >= 0 && sig_len > 0, "bad argument");
    if (n_chunks == 0) return RB200_OK;
    RB200_REQUIRE(seqs_dev && maps_dev && lens_dev && out_dev, "null buffer");
    const int kmer_len = before + after + 1;
    RB200_REQUIRE(map_width >= 2 && seq_width >= kmer_len, "compact arrays too narrow");
    int dev = 0, sms = 148;
    RB200_CUDA_TRY(cudaGetDevice(&dev));
    RB200_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    return launch_encode_dense(seqs_dev, seq_width, maps_dev, map_width, lens_dev, n_chunks,
                               kmer_len, sig_len, out_dev, sms, static_cast<cudaStream_t>(stream),
                               nullptr);
}

int rb200_forward_dense(rb200_handle h, const float *sigs_dev, const float *enc_dev, int32_t B,
                        int32_t T, float *logits_dev, void *stream) {
    RB200_REQUIRE(enc_dev || B == 0, "null enc_kmers");
    return forward_common(h, sigs_dev, enc_dev, nullptr, 0, nullptr, 0, nullptr, B, T, logits_dev,
                          stream);
}

int rb200_forward_compact(rb200_handle h, const float *sigs_dev, const int8_t *seqs_dev,
                          int32_t seq_width, const int16_t *maps_dev, int32_t map_width,
                          const int16_t *lens_dev, int32_t B, int32_t T, float *logits_dev,
                          void *stream) {
    return forward_common(h, sigs_dev, nullptr, seqs_dev, seq_width, maps_dev, map_width, lens_dev,
                          B, T, logits_dev, stream);
}

int rb200_forward_compact_gather(rb200_handle h, const float *sigs_dev, const int8_t *seqs_dev,
                                 int32_t seq_width, const int16_t *maps_dev, int32_t map_width,
                                 const int16_t *lens_dev, int32_t B, int32_t T, float *logits_dev,
                                 void *const *peer_bases_dev, int32_t n_peers, int64_t dst_offset,
                                 void *multicast_base, int64_t flag_word, void *stream_v) {
    RB200_REQUIRE(h != nullptr, "null handle");
    RB200_REQUIRE(B >= 0 && T > 0, "bad batch (%d) / chunk_len (%d)", B, T);
    if (B == 0) return RB200_OK;
    RB200_REQUIRE(sigs_dev && seqs_dev && maps_dev && lens_dev && peer_bases_dev, "null buffer");
    RB200_REQUIRE(n_peers >= 1 && n_peers <= 32 && dst_offset >= 0, "bad gather target");
    RB200_REQUIRE(map_width >= 2 && seq_width >= h->desc.kmer_len, "compact arrays too narrow");
    DeviceGuard guard(h->device);
    RB200_REQUIRE(guard.ok, "cannot select device %d", h->device);
    std::lock_guard<std::mutex> lock(h->mu);
    const int impl = h->impl;
    if (!(impl == RB200_IMPL_AUTO || impl == RB200_IMPL_FUSED_MEGA || impl == RB200_IMPL_FUSED_BF16) ||
        !mega_shape_ok(h, T, seq_width, map_width)) {
        set_error("the fused exchange needs the single-kernel path (model / shape / selected implementation)");
        return RB200_ERR_UNSUPPORTED;
    }
    GatherTarget g;
    g.peers_dev = reinterpret_cast<float *const *>(peer_bases_dev);
    g.n_peers = n_peers;
    g.dst_offset = dst_offset;
    g.multicast_base = static_cast<float *>(multicast_base);
    g.flag_offset = flag_word;
    return mega_forward_compact(h, h->workspaces[stream_v], sigs_dev, seqs_dev, seq_width, maps_dev, map_width,
                                lens_dev, B, T, logits_dev, static_cast<cudaStream_t>(stream_v),
                                impl == RB200_IMPL_FUSED_BF16 ? 1 : 0, &g);
}

int rb200_forward_compact_ship(rb200_handle h, const float *sigs_dev, const int8_t *seqs_dev,
                               int32_t seq_width, const int16_t *maps_dev, int32_t map_width,
                               const int16_t *lens_dev, int32_t B, int32_t T, float *logits_dev,
                               void *const *peer_bases_dev, int32_t n_peers, int32_t self_rank,
                               const float *ship_src_dev, int64_t ship_dst_offset, int64_t ship_count,
                               void *multicast_base, int64_t flag_word, void *stream_v) {
    RB200_REQUIRE(h != nullptr, "null handle");
    RB200_REQUIRE(B >= 0 && T > 0, "bad batch (%d) / chunk_len (%d)", B, T);
    RB200_REQUIRE(B == 0 || (sigs_dev && seqs_dev && maps_dev && lens_dev && logits_dev), "null buffer");
    RB200_REQUIRE(peer_bases_dev != nullptr && n_peers >= 1 && n_peers <= 32 && self_rank >= 0 &&
                      self_rank < n_peers && ship_dst_offset >= 0 && ship_count >= 0 &&
                      ship_count <= (int64_t)1 << 30,
                  "bad gather target");
    if (B == 0 && (ship_src_dev == nullptr || ship_count == 0)) return RB200_OK;
    RB200_REQUIRE(map_width >= 2 && seq_width >= h->desc.kmer_len, "compact arrays too narrow");
    DeviceGuard guard(h->device);
    RB200_REQUIRE(guard.ok, "cannot select device %d", h->device);
    std::lock_guard<std::mutex> lock(h->mu);
    const int impl = h->impl;
    if (!(impl == RB200_IMPL_AUTO || impl == RB200_IMPL_FUSED_MEGA || impl == RB200_IMPL_FUSED_BF16) ||
        !mega_shape_ok(h, T, seq_width, map_width)) {
        set_error("the fused exchange needs the single-kernel path (model / shape / selected implementation)");
        return RB200_ERR_UNSUPPORTED;
    }
    GatherTarget g;
    g.peers_dev = reinterpret_cast<float *const *>(peer_bases_dev);
    g.n_peers = n_peers;
    g.dst_offset = ship_dst_offset;
    g.multicast_base = static_cast<float *>(multicast_base);
    g.flag_offset = flag_word;
    g.deferred = true;
    g.self_rank = self_rank;
    g.ship_src = ship_count > 0 ? ship_src_dev : nullptr;
    g.ship_count = ship_count;
    return mega_forward_compact(h, h->workspaces[stream_v], sigs_dev, seqs_dev, seq_width, maps_dev, map_width,
                                lens_dev, B, T, logits_dev, static_cast<cudaStream_t>(stream_v),
                                impl == RB200_IMPL_FUSED_BF16 ? 1 : 0, &g);
}

int rb200_infer_host(rb200_handle h, const float *sigs_host, const int8_t *seqs_host,
                     int32_t seq_width, const int16_t *maps_host, int32_t map_width,
                     const int16_t *lens_host, int32_t B, int32_t T, float *logits_host) {
    RB200_REQUIRE(h, "null handle");
    RB200_REQUIRE(B >= 0 && T > 0, "bad batch / chunk_len");
    if (B == 0) return RB200_OK;
    RB200_REQUIRE(sigs_host && seqs_host && maps_host && lens_host && logits_host, "null buffer");
    DeviceGuard guard(h->device);
    RB200_REQUIRE(guard.ok, "cannot select device %d", h->device);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t b_sig = up((size_t)B * T * 4), b_seq = up((size_t)B * seq_width),
                 b_map = up((size_t)B * map_width * 2), b_len = up((size_t)B * 2),
                 b_out = up((size_t)B * h->desc.num_out * 4);
    const size_t total = b_sig + b_seq + b_map + b_len + b_out;
    {
        std::lock_guard<std::mutex> lock(h->mu);
        if (!h->host_stream)
            RB200_CUDA_TRY(cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking));
        if (total > h->pinned_bytes) {
            RB200_CUDA_TRY(cudaStreamSynchronize(h->host_stream));
            if (h->pinned) cudaFreeHost(h->pinned);
            if (h->staging_dev) cudaFree(h->staging_dev);
            h->pinned = nullptr;
            h->staging_dev = nullptr;
            h->pinned_bytes = 0;
            RB200_CUDA_TRY(cudaMallocHost(&h->pinned, total));
            RB200_CUDA_TRY(cudaMalloc(&h->staging_dev, total));
            h->pinned_bytes = total;
            h->staging_bytes = total;
        }
    }
    // NB: the staging buffers make this call non-reentrant per handle; callers that need
    // concurrency use rb200_forward_compact with their own device buffers.
    char *p = h->pinned;
    char *d = h->staging_dev;
    cudaStream_t s = h->host_stream;
    memcpy(p, sigs_host, (size_t)B * T * 4);
    memcpy(p + b_sig, seqs_host, (size_t)B * seq_width);
    memcpy(p + b_sig + b_seq, maps_host, (size_t)B * map_width * 2);
    memcpy(p + b_sig + b_seq + b_map, lens_host, (size_t)B * 2);
    const size_t in_bytes = b_sig + b_seq + b_map + b_len;
    RB200_CUDA_TRY(cudaMemcpyAsync(d, p, in_bytes, cudaMemcpyHostToDevice, s));
    int rc = rb200_forward_compact(
        h, reinterpret_cast<const float *>(d), reinterpret_cast<const int8_t *>(d + b_sig),
        seq_width, reinterpret_cast<const int16_t *>(d + b_sig + b_seq), map_width,
        reinterpret_cast<const int16_t *>(d + b_sig + b_seq + b_map), B, T,
        reinterpret_cast<float *>(d + in_bytes), s);
    if (rc) return rc;
    RB200_CUDA_TRY(cudaMemcpyAsync(p + in_bytes, d + in_bytes, (size_t)B * h->desc.num_out * 4,
                                   cudaMemcpyDeviceToHost, s));
    RB200_CUDA_TRY(cudaStreamSynchronize(s));
    memcpy(logits_host, p + in_bytes, (size_t)B * h->desc.num_out * 4);
    return RB200_OK;
}

int rb200_infer_host_async(rb200_handle h, const float *sigs_pinned, const int8_t *seqs_pinned,
                           int32_t seq_width, const int16_t *maps_pinned, int32_t map_width,
                           const int16_t *lens_pinned, int32_t B, int32_t T, float *logits_pinned,
                           void *stream_v) {
    RB200_REQUIRE(h, "null handle");
    RB200_REQUIRE(B >= 0 && T > 0, "bad batch / chunk_len");
    if (B == 0) return RB200_OK;
    RB200_REQUIRE(sigs_pinned && seqs_pinned && maps_pinned && lens_pinned && logits_pinned,
                  "null buffer");
    DeviceGuard guard(h->device);
    RB200_REQUIRE(guard.ok, "cannot select device %d", h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream_v);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t n_sig = (size_t)B * T * 4, n_seq = (size_t)B * seq_width,
                 n_map = (size_t)B * map_width * 2, n_len = (size_t)B * 2,
                 n_out = (size_t)B * h->desc.num_out * 4;
    const size_t o_seq = up(n_sig), o_map = o_seq + up(n_seq), o_len = o_map + up(n_map),
                 o_out = o_len + up(n_len), total = o_out + up(n_out);
    char *d = nullptr;
    {
        std::lock_guard<std::mutex> lock(h->mu);
        Workspace &st = h->host_staging[stream_v];
        int rc = st.ensure(total);
        if (rc) return rc;
        d = st.base;
    }
    // a caller that keeps the four arrays in ONE pinned block, each starting on the next 256-byte boundary after the
    // previous one (the layout of the device staging), gets one copy instead of four: the host side of a pipelined
    // step is a handful of runtime calls, and a short run is bound by them (B200Model.pinned_batch builds such blocks)
    const char *h0 = reinterpret_cast<const char *>(sigs_pinned);
    if (reinterpret_cast<const char *>(seqs_pinned) == h0 + o_seq &&
        reinterpret_cast<const char *>(maps_pinned) == h0 + o_map &&
        reinterpret_cast<const char *>(lens_pinned) == h0 + o_len) {
        RB200_CUDA_TRY(cudaMemcpyAsync(d, h0, o_len + n_len, cudaMemcpyHostToDevice, s));
    } else {
        RB200_CUDA_TRY(cudaMemcpyAsync(d, sigs_pinned, n_sig, cudaMemcpyHostToDevice, s));
        RB200_CUDA_TRY(cudaMemcpyAsync(d + o_seq, seqs_pinned, n_seq, cudaMemcpyHostToDevice, s));
        RB200_CUDA_TRY(cudaMemcpyAsync(d + o_map, maps_pinned, n_map, cudaMemcpyHostToDevice, s));
        RB200_CUDA_TRY(cudaMemcpyAsync(d + o_len, lens_pinned, n_len, cudaMemcpyHostToDevice, s));
    }
    // the 8 B/chunk of logits go straight into the caller's pinned buffer when the device can address it (pinned
    // memory is mapped under unified addressing): the kernel's own stores are the device-to-host transfer, one
    // runtime call less per step.  The mapping of a buffer is looked up once and remembered.
    float *out_dev = nullptr;
    {
        std::lock_guard<std::mutex> lock(h->mu);
        auto it = h->mapped_out.find(logits_pinned);
        if (it == h->mapped_out.end()) {
            cudaPointerAttributes pa = {};
            float *mapped = nullptr;
            if (cudaPointerGetAttributes(&pa, logits_pinned) == cudaSuccess && pa.type == cudaMemoryTypeHost &&
                pa.devicePointer != nullptr)
                mapped = static_cast<float *>(pa.devicePointer);
            else
                (void)cudaGetLastError();
            if (h->mapped_out.size() > 1024) h->mapped_out.clear();
            it = h->mapped_out.emplace(logits_pinned, mapped).first;
        }
        out_dev = it->second;
    }
    int rc = rb200_forward_compact(h, reinterpret_cast<const float *>(d),
                                   reinterpret_cast<const int8_t *>(d + o_seq), seq_width,
                                   reinterpret_cast<const int16_t *>(d + o_map), map_width,
                                   reinterpret_cast<const int16_t *>(d + o_len), B, T,
                                   out_dev != nullptr ? out_dev : reinterpret_cast<float *>(d + o_out), s);
    if (rc) return rc;
    if (out_dev == nullptr)
        RB200_CUDA_TRY(cudaMemcpyAsync(logits_pinned, d + o_out, n_out, cudaMemcpyDeviceToHost, s));
    return RB200_OK;
}

int rb200_chunk_plan(const int32_t *seq_to_sig_map_dev, int32_t n_map, int32_t sig_len,
                     const int32_t *focus_bases_dev, int32_t n, int32_t chunk_before, int32_t chunk_after,
                     int32_t base_start_justify, int32_t offset, int32_t *focus_adj_dev,
                     int32_t *focus_sig_dev, int32_t *seq_start_dev, int32_t *seq_len_dev, void *stream) {
    RB200_REQUIRE(n >= 0 && n_map >= 2 && sig_len >= 0 && chunk_before >= 0 && chunk_after >= 0 &&
                      chunk_before + chunk_after > 0,
                  "bad argument");
    if (n == 0) return RB200_OK;
    RB200_REQUIRE(seq_to_sig_map_dev && focus_bases_dev && focus_adj_dev && focus_sig_dev &&
                      seq_start_dev && seq_len_dev,
                  "null buffer");
    return launch_chunk_plan(seq_to_sig_map_dev, n_map, sig_len, focus_bases_dev, n, chunk_before,
                             chunk_after, base_start_justify, offset, focus_adj_dev, focus_sig_dev,
                             seq_start_dev, seq_len_dev, static_cast<cudaStream_t>(stream));
}

int rb200_chunk_fill(const void *dacs_dev, int32_t dacs_dtype, int32_t sig_len, double shift, double scale,
                     const int32_t *seq_to_sig_map_dev, int32_t n_map, const int8_t *int_seq_dev,
                     int32_t n_bases, const int32_t *focus_sig_dev, const int32_t *seq_start_dev,
                     const int32_t *seq_len_dev, int32_t n, int32_t chunk_before, int32_t chunk_after,
                     int32_t kmer_before, int32_t kmer_after, int32_t lmax, float *signal_dev,
                     int8_t *sequence_dev, int16_t *mapping_dev, int16_t *lens_dev, void *stream) {
    RB200_REQUIRE(n >= 0 && dacs_dtype >= 0 && dacs_dtype <= 2 && lmax >= 1 && kmer_before >= 0 &&
                      kmer_after >= 0 && chunk_before + chunk_after > 0 &&
                      chunk_before + chunk_after < 32768,
                  "bad argument");
    if (n == 0) return RB200_OK;
    RB200_REQUIRE(dacs_dev && seq_to_sig_map_dev && int_seq_dev && focus_sig_dev && seq_start_dev &&
                      seq_len_dev && signal_dev && sequence_dev && mapping_dev && lens_dev,
                  "null buffer");
    return launch_chunk_fill(dacs_dev, dacs_dtype, sig_len, shift, scale, seq_to_sig_map_dev, n_map,
                             int_seq_dev, n_bases, focus_sig_dev, seq_start_dev, seq_len_dev, n,
                             chunk_before, chunk_after, kmer_before, kmer_after, lmax, signal_dev,
                             sequence_dev, mapping_dev, lens_dev, static_cast<cudaStream_t>(stream));
}

int rb200_softmax_ml(const float *logits_dev, int32_t B, int32_t num_out, float *probs_dev,
                     uint8_t *ml_dev, void *stream) {
    RB200_REQUIRE(B >= 0 && num_out >= 2, "bad argument");
    if (B == 0) return RB200_OK;
    RB200_REQUIRE(logits_dev && (probs_dev || ml_dev), "null buffer");
    return launch_softmax_ml(logits_dev, B, num_out, probs_dev, ml_dev,
                             static_cast<cudaStream_t>(stream));
}

static int current_sm_count(int *out) {
    int dev = 0, sms = 0;
    RB200_CUDA_TRY(cudaGetDevice(&dev));
    RB200_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    *out = sms;
    return RB200_OK;
}

int rb200_refine_normalize(const void *dacs_dev, int32_t dacs_dtype, const int64_t *sig_off_dev,
                           const double *shift_dev, const double *scale_dev, int32_t n_reads,
                           int64_t max_len, float *signal_dev, void *stream) {
    RB200_REQUIRE(n_reads >= 0 && max_len >= 0 && dacs_dtype >= 0 && dacs_dtype <= 3, "bad argument");
    if (n_reads == 0 || max_len == 0) return RB200_OK;
    RB200_REQUIRE(dacs_dev && sig_off_dev && shift_dev && scale_dev && signal_dev, "null buffer");
    RB200_REQUIRE(n_reads <= 65535, "at most 65535 reads per call");
    return launch_refine_normalise(dacs_dev, dacs_dtype, sig_off_dev, shift_dev, scale_dev, n_reads,
                                   max_len, signal_dev, static_cast<cudaStream_t>(stream));
}

int rb200_refine_scratch_bytes(int32_t near_cap, int32_t max_band_width, int64_t *bytes) {
    RB200_REQUIRE(bytes && max_band_width >= 0 && near_cap >= 0, "bad argument");
    int sms = 0;
    int rc = current_sm_count(&sms);
    if (rc != RB200_OK) return rc;
    *bytes = (int64_t)refine_wide_scratch_bytes(sms, near_cap, max_band_width);
    return RB200_OK;
}

int rb200_refine_dp(const float *signal_dev, const int64_t *sig_off_dev, const float *levels_dev,
                    const int32_t *band_start_dev, const int32_t *band_end_dev, const int64_t *seq_off_dev,
                    const int64_t *tb_off_dev, const int32_t *order_dev, int32_t n_reads,
                    const float *dwell_penalty_host, int32_t n_penalty, int32_t algo, int32_t near_cap,
                    int32_t max_band_width, int32_t *traceback_ws_dev, int32_t *path_dev, float *score_dev,
                    int32_t *status_dev, int32_t *queue_dev, float *wide_scratch_dev, void *stream) {
    RB200_REQUIRE(n_reads >= 0 && (algo == 0 || algo == 1) && max_band_width >= 1 && near_cap >= 0,
                  "bad argument");
    RB200_REQUIRE(algo == 0 || (n_penalty >= 1 && n_penalty <= 16 && dwell_penalty_host),
                  "the dwell_penalty algorithm needs 1..16 penalties");
    if (n_reads == 0) return RB200_OK;
    RB200_REQUIRE(signal_dev && sig_off_dev && levels_dev && band_start_dev && band_end_dev &&
                      seq_off_dev && tb_off_dev && traceback_ws_dev && path_dev && score_dev &&
                      status_dev && queue_dev,
                  "null buffer");
    int sms = 0;
    int rc = current_sm_count(&sms);
    if (rc != RB200_OK) return rc;
    RB200_REQUIRE(refine_wide_scratch_bytes(sms, near_cap, max_band_width) == 0 || wide_scratch_dev,
                  "bands wider than the shared-memory rows need wide_scratch_dev "
                  "(rb200_refine_scratch_bytes)");
    return launch_refine_dp(signal_dev, sig_off_dev, levels_dev, band_start_dev, band_end_dev, seq_off_dev,
                            tb_off_dev, order_dev, n_reads, dwell_penalty_host, algo == 1 ? n_penalty : 1,
                            algo, near_cap, max_band_width, traceback_ws_dev, path_dev, score_dev, status_dev,
                            queue_dev, wide_scratch_dev, sms, static_cast<cudaStream_t>(stream));
}

int rb200_svb16_scratch_bytes(int32_t n_rows, int32_t max_row_samples, int64_t *bytes) {
    RB200_REQUIRE(bytes && n_rows >= 0 && max_row_samples >= 0, "bad argument");
    *bytes = (int64_t)svb16_scratch_bytes(n_rows, max_row_samples);
    return RB200_OK;
}

int rb200_svb16_decode(const uint8_t *packed_dev, const int64_t *row_off_dev, const int32_t *row_samples_dev,
                       const int64_t *out_off_dev, int32_t n_rows, int32_t max_row_samples, int16_t *out_dev,
                       int32_t *status_dev, void *scratch_dev, void *stream) {
    RB200_REQUIRE(n_rows >= 0 && max_row_samples >= 0, "bad argument");
    if (n_rows == 0) return RB200_OK;
    RB200_REQUIRE(packed_dev && row_off_dev && row_samples_dev && out_off_dev && out_dev && status_dev &&
                      scratch_dev,
                  "null buffer");
    return launch_svb16_decode(packed_dev, row_off_dev, row_samples_dev, out_off_dev, n_rows, max_row_samples,
                               out_dev, status_dev, scratch_dev, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
