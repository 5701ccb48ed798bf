// Layer-per-kernel CUDA path: one kernel per reference layer, any size / kmer_len / chunk_len,
// both architectures.  It is the always-available path (Conv_w_ref, non-64 widths) and the
// independent cross-check of the fused kernels: it evaluates the network exactly as the reference
// writes it (models/ConvLSTM_w_ref.py:39-58, models/Conv_w_ref.py:44-62), including the FULL
// second LSTM over the flipped sequence, with BatchNorm folded into the convolutions.
// Activations are channel-first float32 [B][C][T] like the reference tensors.
#include "rb200_internal.cuh"

#include <cstdlib>

#include <algorithm>

namespace rb200 {

// y[b][co][t] = swish(bias[co] + sum_{ci,j} w[co][ci][j] * x[b][ci][t*stride + j])
// One thread per output element; a warp covers consecutive t of one (b, co) so x loads coalesce
// and w loads are warp-uniform.
__global__ void __launch_bounds__(256)
conv1d_swish_kernel(const float *__restrict__ x, int64_t x_bstride, const float *__restrict__ w,
                    const float *__restrict__ bias, float *__restrict__ y, int64_t y_bstride,
                    int B, int c_in, int t_in, int c_out, int t_out, int kw, int stride) {
    const int64_t total = (int64_t)B * c_out * t_out;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(idx % t_out);
        const int co = (int)((idx / t_out) % c_out);
        const int b = (int)(idx / ((int64_t)t_out * c_out));
        const float *xb = x + b * x_bstride + t * stride;
        const float *wc = w + (size_t)co * c_in * kw;
        float acc = bias[co];
        for (int ci = 0; ci < c_in; ++ci) {
            const float *xr = xb + (size_t)ci * t_in;
            const float *wr = wc + ci * kw;
            for (int j = 0; j < kw; ++j) acc = fmaf(__ldg(wr + j), __ldg(xr + j), acc);
        }
        y[b * y_bstride + (size_t)co * t_out + t] = swishf(acc);
    }
}

// Single-layer LSTM over one chunk per CTA (torch.nn.LSTM semantics, gates i,f,g,o, h0=c0=0).
// x: [B][H][steps] channel-first.  out: [B][H][steps], out[., ., t] = swish(h_t) as the reference
// applies swish to the LSTM output (ConvLSTM_w_ref.py:52-53).  reverse=1 walks t = steps-1 .. 0,
// which equals flip -> lstm -> flip (ConvLSTM_w_ref.py:53).  blockDim.x = min(4H, 1024).
__global__ void lstm_layer_kernel(const float *__restrict__ x, const float *__restrict__ w_ih,
                                  const float *__restrict__ w_hh, const float *__restrict__ bias,
                                  float *__restrict__ out, int H, int steps, int reverse) {
    extern __shared__ float lsm[];
    float *xs = lsm;            // [H]
    float *hs = xs + H;         // [H]
    float *gs = hs + H;         // [4H]
    float *cs = gs + 4 * H;     // [H]
    const int b = blockIdx.x;
    const float *xb = x + (size_t)b * H * steps;
    float *ob = out + (size_t)b * H * steps;
    for (int k = threadIdx.x; k < H; k += blockDim.x) {
        hs[k] = 0.f;
        cs[k] = 0.f;
    }
    for (int it = 0; it < steps; ++it) {
        const int t = reverse ? steps - 1 - it : it;
        __syncthreads();
        for (int k = threadIdx.x; k < H; k += blockDim.x) xs[k] = xb[(size_t)k * steps + t];
        __syncthreads();
        for (int r = threadIdx.x; r < 4 * H; r += blockDim.x) {
            const float *wi = w_ih + (size_t)r * H;
            const float *wh = w_hh + (size_t)r * H;
            float acc = bias[r];
            for (int k = 0; k < H; ++k) acc = fmaf(__ldg(wi + k), xs[k], acc);
            for (int k = 0; k < H; ++k) acc = fmaf(__ldg(wh + k), hs[k], acc);
            gs[r] = acc;
        }
        __syncthreads();
        for (int u = threadIdx.x; u < H; u += blockDim.x) {
            const float ig = sigmoidf_acc(gs[u]);
            const float fg = sigmoidf_acc(gs[H + u]);
            const float gg = tanhf(gs[2 * H + u]);
            const float og = sigmoidf_acc(gs[3 * H + u]);
            const float c = fg * cs[u] + ig * gg;
            const float h = og * tanhf(c);
            cs[u] = c;
            hs[u] = h;
            ob[(size_t)u * steps + t] = swishf(h);
        }
    }
}

// logits[b][o] = fc_b[o] + sum_k fc_w[o][k] * feat(b, k); feat is x[b*bstride + k*kstride]
__global__ void fc_kernel(const float *__restrict__ x, int64_t bstride, int64_t kstride,
                          const float *__restrict__ w, const float *__restrict__ bias,
                          float *__restrict__ logits, int B, int fc_in, int num_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * num_out) return;
    const int b = idx / num_out, o = idx - b * num_out;
    const float *xb = x + (size_t)b * bstride;
    const float *wo = w + (size_t)o * fc_in;
    float acc = bias[o];
    for (int k = 0; k < fc_in; ++k) acc = fmaf(__ldg(wo + k), xb[(size_t)k * kstride], acc);
    logits[idx] = acc;
}

__global__ void softmax_ml_kernel(const float *__restrict__ logits, int B, int num_out,
                                  float *__restrict__ probs, uint8_t *__restrict__ ml) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float *l = logits + (size_t)b * num_out;
    float mx = l[0];
    for (int o = 1; o < num_out; ++o) mx = fmaxf(mx, l[o]);
    float den = 0.f;
    for (int o = 0; o < num_out; ++o) den += expf(l[o] - mx);
    for (int o = 1; o < num_out; ++o) {
        const float p = expf(l[o] - mx) / den;
        if (probs) probs[(size_t)b * (num_out - 1) + o - 1] = p;
        if (ml) {
            // util.py:532-535 computes floor(p*256) in float64 from the float32 softmax
            const double sc = floor((double)p * 256.0);
            ml[(size_t)b * (num_out - 1) + o - 1] = (uint8_t)(sc >= 256.0 ? 255.0 : sc);
        }
    }
}

int launch_softmax_ml(const float *logits, int B, int num_out, float *probs, uint8_t *ml,
                      cudaStream_t stream) {
    if (B == 0) return RB200_OK;
    softmax_ml_kernel<<<(B + 127) / 128, 128, 0, stream>>>(logits, B, num_out, probs, ml);
    RB200_CUDA_TRY(cudaGetLastError());
    return RB200_OK;
}

static inline int conv_out_len(int t_in, int kw, int stride) {
    return t_in < kw ? 0 : (t_in - kw) / stride + 1;
}

struct Plan {
    // time lengths after each conv of each track
    int sig_t[RB200_MAX_CONVS + 1], seq_t[RB200_MAX_CONVS + 1], mrg_t[RB200_MAX_CONVS + 1];
    bool ok;
};

static Plan make_plan(const rb200_model_desc &d, int T) {
    Plan p;
    p.ok = true;
    p.sig_t[0] = T;
    for (int i = 0; i < d.n_sig_conv; ++i)
        p.sig_t[i + 1] = conv_out_len(p.sig_t[i], d.sig_conv[i].kw, d.sig_conv[i].stride);
    p.seq_t[0] = T;
    for (int i = 0; i < d.n_seq_conv; ++i)
        p.seq_t[i + 1] = conv_out_len(p.seq_t[i], d.seq_conv[i].kw, d.seq_conv[i].stride);
    if (p.sig_t[d.n_sig_conv] != p.seq_t[d.n_seq_conv] || p.sig_t[d.n_sig_conv] <= 0) p.ok = false;
    p.mrg_t[0] = p.sig_t[d.n_sig_conv];
    for (int i = 0; i < d.n_merge_conv; ++i)
        p.mrg_t[i + 1] = conv_out_len(p.mrg_t[i], d.merge_conv[i].kw, d.merge_conv[i].stride);
    if (p.mrg_t[d.n_merge_conv] <= 0) p.ok = false;
    return p;
}

static inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

size_t layers_workspace_bytes(const rb200_model_desc &d, int B, int T, bool compact) {
    Plan p = make_plan(d, T);
    if (!p.ok) return 0;
    size_t total = 0;
    if (compact) total += align_up((size_t)B * 4 * d.kmer_len * T * 4);
    for (int i = 0; i < d.n_sig_conv - 1; ++i)
        total += align_up((size_t)B * d.sig_conv[i].c_out * p.sig_t[i + 1] * 4);
    for (int i = 0; i < d.n_seq_conv - 1; ++i)
        total += align_up((size_t)B * d.seq_conv[i].c_out * p.seq_t[i + 1] * 4);
    total += align_up((size_t)B * 2 * d.size * p.mrg_t[0] * 4);  // cat
    for (int i = 0; i < d.n_merge_conv; ++i)
        total += align_up((size_t)B * d.merge_conv[i].c_out * p.mrg_t[i + 1] * 4);
    if (d.n_lstm) total += 2 * align_up((size_t)B * d.size * p.mrg_t[d.n_merge_conv] * 4);
    return total + 1024;
}

static int run_conv(rb200_model *m, int track, int layer, bool tiled, const rb200_conv_desc &c,
                    const float *x, int64_t x_bstride, int t_in, float *y, int64_t y_bstride, int B,
                    cudaStream_t stream) {
    if (tiled) {
        const int rc = tiled_conv(m, track, layer, c, x, x_bstride, t_in, y, y_bstride, B, stream);
        if (rc != RB200_ERR_UNSUPPORTED) return rc;
    }
    const int t_out = conv_out_len(t_in, c.kw, c.stride);
    const int64_t total = (int64_t)B * c.c_out * t_out;
    const int threads = 256;
    int64_t blocks = (total + threads - 1) / threads;
    if (blocks > (int64_t)m->sm_count * 32) blocks = (int64_t)m->sm_count * 32;
    conv1d_swish_kernel<<<(int)blocks, threads, 0, stream>>>(
        x, x_bstride, m->blob_dev + c.w_off, m->blob_dev + c.b_off, y, y_bstride, B, c.c_in, t_in,
        c.c_out, t_out, c.kw, c.stride);
    m->launches++;
    RB200_CUDA_TRY(cudaGetLastError());
    return RB200_OK;
}

int layers_forward(rb200_model *m, Workspace &ws, const float *sigs, const float *enc,
                   const int8_t *seqs, int seq_width, const int16_t *maps, int map_width,
                   const int16_t *lens, int B, int T, float *logits, cudaStream_t stream,
                   bool tiled) {
    const rb200_model_desc &d = m->desc;
    Plan p = make_plan(d, T);
    RB200_REQUIRE(p.ok, "chunk_len %d too short / inconsistent for this architecture", T);
    const int t_final = p.mrg_t[d.n_merge_conv];
    if (d.arch == RB200_ARCH_CONV_W_REF)
        RB200_REQUIRE(d.fc_in == d.merge_conv[d.n_merge_conv - 1].c_out * t_final,
                      "Conv_w_ref fc expects %d features but chunk_len %d gives %d "
                      "(reference models/Conv_w_ref.py:42 only supports one chunk_len)",
                      d.fc_in, T, d.merge_conv[d.n_merge_conv - 1].c_out * t_final);
    const bool compact = enc == nullptr;
    const size_t need = layers_workspace_bytes(d, B, T, compact);
    int rc = ws.ensure(need);
    if (rc) return rc;
    char *cur = ws.base;
    auto take = [&](size_t bytes) {
        float *ptr = reinterpret_cast<float *>(cur);
        cur += align_up(bytes);
        return ptr;
    };
    if (m->keep_debug) m->debug.clear();
    auto keep = [&](const char *name, const float *ptr, int C, int Tn) {
        if (m->keep_debug) m->debug.push_back({name, ptr, B, C, Tn});
    };
    static const char *sig_names[] = {"sig1", "sig2", "sig3", "sig4"};
    static const char *seq_names[] = {"seq1", "seq2", "seq3", "seq4"};
    static const char *mrg_names[] = {"merge1", "merge2", "merge3", "merge4"};

    // tiled mode: the first sequence convolution reads the compact arrays itself (gather form)
    bool gather_seq1 = false;
    if (compact && tiled) gather_seq1 = tiled_gather_ok(m, seq_width, map_width, T);
    if (compact && !gather_seq1) {
        float *enc_buf = take((size_t)B * 4 * d.kmer_len * T * 4);
        uint64_t l = 0;
        rc = launch_encode_dense(seqs, seq_width, maps, map_width, lens, B, d.kmer_len, T, enc_buf,
                                 m->sm_count, stream, &l);
        if (rc) return rc;
        m->launches += l;
        enc = enc_buf;
    }
    // cat buffer [B][2*size][T3]; the last conv of each track writes straight into its half
    const int t_cat = p.mrg_t[0];
    const int64_t cat_bstride = (int64_t)2 * d.size * t_cat;
    // signal track
    const float *x = sigs;
    int64_t xb = (int64_t)d.sig_conv[0].c_in * T;
    std::vector<float *> sig_bufs, seq_bufs;
    for (int i = 0; i < d.n_sig_conv - 1; ++i)
        sig_bufs.push_back(take((size_t)B * d.sig_conv[i].c_out * p.sig_t[i + 1] * 4));
    for (int i = 0; i < d.n_seq_conv - 1; ++i)
        seq_bufs.push_back(take((size_t)B * d.seq_conv[i].c_out * p.seq_t[i + 1] * 4));
    float *cat = take((size_t)B * cat_bstride * 4);
    // The two tracks are independent until the concatenation: in tiled mode the signal track goes to a
    // side stream (fork / join with events) so that its kernels share the SMs with the sequence track's
    // (each tiled convolution alone keeps the FMA pipe 54-63 % busy with one 8-warp CTA per SM).
    if (m->two_tracks < 0) {
        const char *e = getenv("RB200_TWO_TRACKS");
        m->two_tracks = (e && e[0] == '0') ? 0 : 1;
    }
    const bool fork = tiled && m->two_tracks == 1 && !m->keep_debug;
    cudaStream_t sig_stream = stream;
    if (fork) {
        if (!m->side_stream) {
            RB200_CUDA_TRY(cudaStreamCreateWithFlags(&m->side_stream, cudaStreamNonBlocking));
            RB200_CUDA_TRY(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
            RB200_CUDA_TRY(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
        }
        RB200_CUDA_TRY(cudaEventRecord(m->ev_fork, stream));
        RB200_CUDA_TRY(cudaStreamWaitEvent(m->side_stream, m->ev_fork, 0));
        sig_stream = m->side_stream;
    }
    for (int i = 0; i < d.n_sig_conv; ++i) {
        const bool last = i == d.n_sig_conv - 1;
        float *y = last ? cat : sig_bufs[i];
        const int64_t yb = last ? cat_bstride : (int64_t)d.sig_conv[i].c_out * p.sig_t[i + 1];
        rc = run_conv(m, 0, i, tiled, d.sig_conv[i], x, xb, p.sig_t[i], y, yb, B, sig_stream);
        if (rc) return rc;
        if (!last) keep(sig_names[i], y, d.sig_conv[i].c_out, p.sig_t[i + 1]);
        x = y;
        xb = yb;
    }
    x = enc;
    xb = (int64_t)d.seq_conv[0].c_in * T;
    for (int i = 0; i < d.n_seq_conv; ++i) {
        const bool last = i == d.n_seq_conv - 1;
        float *y = last ? cat + (size_t)d.size * t_cat : seq_bufs[i];
        const int64_t yb = last ? cat_bstride : (int64_t)d.seq_conv[i].c_out * p.seq_t[i + 1];
        if (i == 0 && gather_seq1)
            rc = tiled_seq1_gather(m, seqs, seq_width, maps, map_width, lens, B, T, y, yb, stream);
        else
            rc = run_conv(m, 1, i, tiled, d.seq_conv[i], x, xb, p.seq_t[i], y, yb, B, stream);
        if (rc) {
            if (rc == RB200_ERR_UNSUPPORTED) set_error("sequence convolution %d has no kernel", i);
            return rc;
        }
        if (!last) keep(seq_names[i], y, d.seq_conv[i].c_out, p.seq_t[i + 1]);
        x = y;
        xb = yb;
    }
    if (fork) {  // join: the merge convolutions read both halves of cat
        RB200_CUDA_TRY(cudaEventRecord(m->ev_join, m->side_stream));
        RB200_CUDA_TRY(cudaStreamWaitEvent(stream, m->ev_join, 0));
    }
    keep("cat", cat, 2 * d.size, t_cat);
    x = cat;
    xb = cat_bstride;
    int t_cur = t_cat;
    for (int i = 0; i < d.n_merge_conv; ++i) {
        float *y = take((size_t)B * d.merge_conv[i].c_out * p.mrg_t[i + 1] * 4);
        const int64_t yb = (int64_t)d.merge_conv[i].c_out * p.mrg_t[i + 1];
        rc = run_conv(m, 2, i, tiled, d.merge_conv[i], x, xb, t_cur, y, yb, B, stream);
        if (rc) return rc;
        keep(mrg_names[i], y, d.merge_conv[i].c_out, p.mrg_t[i + 1]);
        x = y;
        xb = yb;
        t_cur = p.mrg_t[i + 1];
    }
    if (d.n_lstm == 2) {
        const int H = d.size;
        RB200_REQUIRE(H <= 4096, "LSTM width %d not supported", H);
        float *l1 = take((size_t)B * H * t_cur * 4);
        float *l2 = take((size_t)B * H * t_cur * 4);
        const int threads = 4 * H < 1024 ? (4 * H < 32 ? 32 : 4 * H) : 1024;
        const size_t smem = (size_t)7 * H * sizeof(float);
        RB200_REQUIRE(smem <= 48 * 1024, "LSTM width %d not supported", H);
        lstm_layer_kernel<<<B, threads, smem, stream>>>(x, m->blob_dev + d.lstm_w_ih_off[0],
                                                        m->blob_dev + d.lstm_w_hh_off[0],
                                                        m->blob_dev + d.lstm_b_off[0], l1, H,
                                                        t_cur, 0);
        m->launches++;
        RB200_CUDA_TRY(cudaGetLastError());
        keep("lstm1", l1, H, t_cur);
        lstm_layer_kernel<<<B, threads, smem, stream>>>(l1, m->blob_dev + d.lstm_w_ih_off[1],
                                                        m->blob_dev + d.lstm_w_hh_off[1],
                                                        m->blob_dev + d.lstm_b_off[1], l2, H,
                                                        t_cur, 1);
        m->launches++;
        RB200_CUDA_TRY(cudaGetLastError());
        keep("lstm2", l2, H, t_cur);
        // z[-1] of the re-flipped sequence = time index t_cur-1 (ConvLSTM_w_ref.py:53-54)
        const int n = B * d.num_out;
        fc_kernel<<<(n + 127) / 128, 128, 0, stream>>>(l2 + (t_cur - 1), (int64_t)H * t_cur, t_cur,
                                                       m->blob_dev + d.fc_w_off,
                                                       m->blob_dev + d.fc_b_off, logits, B, H,
                                                       d.num_out);
        m->launches++;
    } else if (tiled) {
        rc = tiled_fc(m, x, xb, logits, B, stream);
        if (rc) return rc;
    } else {
        // torch.flatten([B][C][T]) -> feature index c*T + t = channel-first buffer as is
        const int n = B * d.num_out;
        fc_kernel<<<(n + 127) / 128, 128, 0, stream>>>(x, xb, 1, m->blob_dev + d.fc_w_off,
                                                       m->blob_dev + d.fc_b_off, logits, B,
                                                       d.fc_in, d.num_out);
        m->launches++;
    }
    RB200_CUDA_TRY(cudaGetLastError());
    return RB200_OK;
}

}  // namespace rb200
