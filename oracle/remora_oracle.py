"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's per-chunk inference hot path, used only as the
checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
remora_b200/ never imports this module.

Parity status: PINNED.  The reference holds no golden vectors for this path, so the
oracle is pinned against outputs of the reference itself run in the build container:
  * encoder: bit-for-bit against the reference's compiled Cython encoder (oracle/_ref) and the
    committed vectors tests/golden/encode_cases.npz (incl. the SURVEY §8c known-answer case);
  * forward: against logits produced by the reference's own TorchScript modules
    (tests/golden/*.pt exported with reference model_util.export_model_torchscript, outputs in
    tests/golden/forward_cases.npz), tolerance 2e-6 (fp32 re-association only).
Generating script: tests/golden/make_golden.py.

Each function cites the reference file:line it restates.
"""
import ctypes
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------------------
# k-mer one-hot encoder   (reference src/remora/encoded_kmers.pyx:13-45)
# --------------------------------------------------------------------------------------
def encode_kmers_numpy(before, after, seqs, seq_mappings, seq_lens):
    """Pure-numpy/Python-loop restatement; small cases only.

    out[c, 4*p + base, map[c,s]:map[c,s+1]] = 1 for every k-mer offset p and base index s with
    base = seqs[c, s+p] != -1 (pyx:33-44); sig_len from chunk 0 (pyx:23); zero init (pyx:26)."""
    seqs = np.asarray(seqs, dtype=np.int8)
    seq_mappings = np.asarray(seq_mappings, dtype=np.int16)
    seq_lens = np.asarray(seq_lens, dtype=np.int16)
    n = seq_lens.size
    kmer_len = before + after + 1
    sig_len = int(seq_mappings[0, seq_lens[0]])
    out = np.zeros((n, 4 * kmer_len, sig_len), dtype=np.float32)
    for c in range(n):
        for p in range(kmer_len):
            for s in range(int(seq_lens[c])):
                b = int(seqs[c, s + p])
                if b == -1:
                    continue
                out[c, 4 * p + b, int(seq_mappings[c, s]):int(seq_mappings[c, s + 1])] = 1.0
    return out


_LIB = None


def _liboracle():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.isfile(path):
            from build_ref import build_liboracle  # noqa  (oracle/ is on sys.path for callers)
            build_liboracle()
        lib = ctypes.CDLL(path)
        lib.oracle_compute_encoded_kmer_batch.argtypes = [
            ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
            ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        lib.oracle_compute_encoded_kmer_batch.restype = None
        _LIB = lib
    return _LIB


def encode_kmers_c(before, after, seqs, seq_mappings, seq_lens):
    """C restatement (oracle/oracle_encode.c); same arguments/result as the reference function."""
    seqs = np.ascontiguousarray(seqs, dtype=np.int8)
    seq_mappings = np.ascontiguousarray(seq_mappings, dtype=np.int16)
    seq_lens = np.ascontiguousarray(seq_lens, dtype=np.int16)
    n = seq_lens.size
    kmer_len = before + after + 1
    sig_len = int(seq_mappings[0, seq_lens[0]])
    out = np.empty((n, 4 * kmer_len, sig_len), dtype=np.float32)
    _liboracle().oracle_compute_encoded_kmer_batch(
        before, after, seqs.ctypes.data, seqs.shape[1], seq_mappings.ctypes.data,
        seq_mappings.shape[1], seq_lens.ctypes.data, n, sig_len, out.ctypes.data)
    return out


def load_ref_encoder():
    """The reference's own compiled Cython encoder from oracle/_ref (None when not built)."""
    import sysconfig
    path = os.path.join(HERE, "_ref", "encoded_kmers" + sysconfig.get_config_var("EXT_SUFFIX"))
    if not os.path.isfile(path):
        return None
    spec = importlib.util.spec_from_file_location("encoded_kmers", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def encode_kmers(before, after, seqs, seq_mappings, seq_lens):
    """Best available CPU encoder: the compiled reference if present, else the C restatement."""
    ref = load_ref_encoder()
    if ref is not None:
        return ref.compute_encoded_kmer_batch(before, after, seqs, seq_mappings, seq_lens)
    return encode_kmers_c(before, after, seqs, seq_mappings, seq_lens)


# --------------------------------------------------------------------------------------
# network forward   (reference models/ConvLSTM_w_ref.py:39-58, models/Conv_w_ref.py:44-62)
# --------------------------------------------------------------------------------------
def _swish(x):
    """reference src/remora/activations.py:18"""
    import torch
    return x * torch.sigmoid(x)


def _conv_bn_swish(x, sd, conv, bn, stride=1):
    """conv1d (valid) -> eval-mode BatchNorm1d (running stats, eps 1e-5) -> swish."""
    import torch.nn.functional as F
    y = F.conv1d(x, sd[f"{conv}.weight"], sd[f"{conv}.bias"], stride=stride)
    y = F.batch_norm(y, sd[f"{bn}.running_mean"], sd[f"{bn}.running_var"], sd[f"{bn}.weight"],
                     sd[f"{bn}.bias"], training=False, eps=1e-5)
    return _swish(y)


def _lstm_forward(x, sd, name):
    """Single-layer uni-directional LSTM, zero initial state, gate order i,f,g,o
    (torch.nn.LSTM semantics used at ConvLSTM_w_ref.py:32-33,51-53).  x: [steps, B, H]."""
    import torch
    w_ih, w_hh = sd[f"{name}.weight_ih_l0"], sd[f"{name}.weight_hh_l0"]
    b = sd[f"{name}.bias_ih_l0"] + sd[f"{name}.bias_hh_l0"]
    H = w_hh.shape[1]
    h = x.new_zeros(x.shape[1], H)
    c = x.new_zeros(x.shape[1], H)
    outs = []
    for t in range(x.shape[0]):
        g = x[t] @ w_ih.T + h @ w_hh.T + b
        i, f, gg, o = g[:, :H], g[:, H:2 * H], g[:, 2 * H:3 * H], g[:, 3 * H:]
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs)


def forward_convlstm(sd, sigs, seqs):
    """ConvLSTM_w_ref.network.forward (models/ConvLSTM_w_ref.py:39-58), full (un-shortcut) form:
    both LSTMs run over every step, flips included."""
    import torch
    with torch.no_grad():
        s = _conv_bn_swish(sigs, sd, "sig_conv1", "sig_bn1")
        s = _conv_bn_swish(s, sd, "sig_conv2", "sig_bn2")
        s = _conv_bn_swish(s, sd, "sig_conv3", "sig_bn3", stride=3)
        q = _conv_bn_swish(seqs, sd, "seq_conv1", "seq_bn1")
        q = _conv_bn_swish(q, sd, "seq_conv2", "seq_bn2", stride=3)
        z = torch.cat((s, q), 1)
        z = _conv_bn_swish(z, sd, "merge_conv1", "merge_bn")
        z = z.permute(2, 0, 1)
        z = _swish(_lstm_forward(z, sd, "lstm1"))
        z = torch.flip(_swish(_lstm_forward(torch.flip(z, (0,)), sd, "lstm2")), (0,))
        z = z[-1]
        return z @ sd["fc.weight"].T + sd["fc.bias"]


def forward_conv(sd, sigs, seqs):
    """Conv_w_ref.network.forward (models/Conv_w_ref.py:44-62)."""
    import torch
    with torch.no_grad():
        s = _conv_bn_swish(sigs, sd, "sig_conv1", "sig_bn1")
        s = _conv_bn_swish(s, sd, "sig_conv2", "sig_bn2")
        s = _conv_bn_swish(s, sd, "sig_conv3", "sig_bn3", stride=3)
        q = _conv_bn_swish(seqs, sd, "seq_conv1", "seq_bn1")
        q = _conv_bn_swish(q, sd, "seq_conv2", "seq_bn2")
        q = _conv_bn_swish(q, sd, "seq_conv3", "seq_bn3", stride=3)
        z = torch.cat((s, q), 1)
        z = _conv_bn_swish(z, sd, "merge_conv1", "merge_bn1")
        z = _conv_bn_swish(z, sd, "merge_conv2", "merge_bn2")
        z = _conv_bn_swish(z, sd, "merge_conv3", "merge_bn3", stride=2)
        z = _conv_bn_swish(z, sd, "merge_conv4", "merge_bn4", stride=2)
        z = torch.flatten(z, start_dim=1)
        return z @ sd["fc.weight"].T + sd["fc.bias"]


def forward_from_state_dict(sd, sigs, seqs):
    """Dispatch on the module names exactly like the reference's exporter does
    (model_util.py:231-263: lstm vs conv layer sets)."""
    import torch
    sd = {k: v.detach().to("cpu", torch.float32) for k, v in sd.items() if v.dtype.is_floating_point}
    sigs = torch.as_tensor(sigs, dtype=torch.float32)
    seqs = torch.as_tensor(seqs, dtype=torch.float32)
    if "lstm1.weight_ih_l0" in sd:
        return forward_convlstm(sd, sigs, seqs)
    return forward_conv(sd, sigs, seqs)


def oracle_infer_compact(sd, kmer_context_bases, signal, sequence, mapping, seq_lens):
    """encode (C restatement / compiled reference) + forward; returns float32 logits [N, num_out]."""
    enc = encode_kmers(kmer_context_bases[0], kmer_context_bases[1], sequence, mapping, seq_lens)
    return forward_from_state_dict(sd, signal, enc).numpy()


# --------------------------------------------------------------------------------------
# post-processing restatements (reference src/remora/util.py:182-186, 532-535)
# --------------------------------------------------------------------------------------
def softmax_axis1(x):
    e = np.exp((x.T - np.max(x, axis=1)).T)
    return (e.T / e.sum(axis=1)).T


def ml_bytes(probs):
    scaled = np.floor(np.asarray(probs, dtype=np.float64) * 256)
    scaled[scaled == 256] = 255
    return scaled.astype(np.uint8)
