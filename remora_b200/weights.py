"""TorchScript ``state_dict`` -> canonical BN-folded weight blob + ``rb200_model_desc``.

The reference ships models as TorchScript modules (src/remora/model_util.py:468-481); their
``state_dict()`` layout is listed in SURVEY.md Appendix C.  Architecture detection follows the
reference's own exporter (model_util.py:231-263): ``lstm1`` present -> ConvLSTM_w_ref, else Conv_w_ref.
BatchNorm (eval, eps 1e-5) is folded into the preceding convolution exactly like
``torch.nn.utils.fusion.fuse_conv_bn_eval`` (used by the reference at model_util.py:216); LSTM biases
are summed (scripts/convert_ts_to_ont_json.py:134-135).  Folding is done in float64 and rounded
once to float32.
"""
import ctypes

import numpy as np

from . import RemoraError, _native

BN_EPS = 1e-5

# (conv name, bn name, stride) per track, fixed by the reference model files
_CONVLSTM_LAYERS = {
    "sig": [("sig_conv1", "sig_bn1", 1), ("sig_conv2", "sig_bn2", 1), ("sig_conv3", "sig_bn3", 3)],
    "seq": [("seq_conv1", "seq_bn1", 1), ("seq_conv2", "seq_bn2", 3)],
    "merge": [("merge_conv1", "merge_bn", 1)],
}  # models/ConvLSTM_w_ref.py:18-31
_CONV_LAYERS = {
    "sig": [("sig_conv1", "sig_bn1", 1), ("sig_conv2", "sig_bn2", 1), ("sig_conv3", "sig_bn3", 3)],
    "seq": [("seq_conv1", "seq_bn1", 1), ("seq_conv2", "seq_bn2", 1), ("seq_conv3", "seq_bn3", 3)],
    "merge": [("merge_conv1", "merge_bn1", 1), ("merge_conv2", "merge_bn2", 1),
              ("merge_conv3", "merge_bn3", 2), ("merge_conv4", "merge_bn4", 2)],
}  # models/Conv_w_ref.py:18-40


def _np(sd, key):
    try:
        return sd[key].detach().cpu().numpy().astype(np.float64)
    except KeyError:
        raise RemoraError(f"model state_dict lacks {key}; unsupported architecture")


def fold_conv_bn(sd, conv, bn):
    w, b = _np(sd, f"{conv}.weight"), _np(sd, f"{conv}.bias")
    g, beta = _np(sd, f"{bn}.weight"), _np(sd, f"{bn}.bias")
    mu, var = _np(sd, f"{bn}.running_mean"), _np(sd, f"{bn}.running_var")
    scale = g / np.sqrt(var + BN_EPS)
    return ((w * scale[:, None, None]).astype(np.float32),
            ((b - mu) * scale + beta).astype(np.float32))


def pack_state_dict(sd):
    """Returns (ModelDesc, blob float32 ndarray, info dict)."""
    is_lstm = "lstm1.weight_ih_l0" in sd
    layers = _CONVLSTM_LAYERS if is_lstm else _CONV_LAYERS
    desc = _native.ModelDesc()
    desc.struct_size = ctypes.sizeof(_native.ModelDesc)
    desc.arch = _native.ARCH_CONVLSTM_W_REF if is_lstm else _native.ARCH_CONV_W_REF
    parts, cursor = [], 0

    def push(arr):
        nonlocal cursor
        arr = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1)
        off = cursor
        parts.append(arr)
        pad = (-arr.size) % 4  # keep every tensor 16-byte aligned inside the blob
        if pad:
            parts.append(np.zeros(pad, dtype=np.float32))
        cursor += arr.size + pad
        return off

    for track, field, count in (("sig", desc.sig_conv, "n_sig_conv"),
                                ("seq", desc.seq_conv, "n_seq_conv"),
                                ("merge", desc.merge_conv, "n_merge_conv")):
        setattr(desc, count, len(layers[track]))
        for i, (conv, bn, stride) in enumerate(layers[track]):
            w, b = fold_conv_bn(sd, conv, bn)
            field[i].c_out, field[i].c_in, field[i].kw = w.shape
            field[i].stride = stride
            field[i].w_off = push(w)
            field[i].b_off = push(b)
    size = desc.merge_conv[0].c_in // 2
    desc.size = size
    if desc.seq_conv[0].c_in % 4:
        raise RemoraError("seq_conv1 input channels not a multiple of 4")
    desc.kmer_len = desc.seq_conv[0].c_in // 4
    if is_lstm:
        desc.n_lstm = 2
        for l, name in enumerate(("lstm1", "lstm2")):
            desc.lstm_w_ih_off[l] = push(_np(sd, f"{name}.weight_ih_l0"))
            desc.lstm_w_hh_off[l] = push(_np(sd, f"{name}.weight_hh_l0"))
            desc.lstm_b_off[l] = push(_np(sd, f"{name}.bias_ih_l0") + _np(sd, f"{name}.bias_hh_l0"))
    else:
        desc.n_lstm = 0
    fc_w, fc_b = _np(sd, "fc.weight"), _np(sd, "fc.bias")
    desc.num_out, desc.fc_in = fc_w.shape
    desc.fc_w_off = push(fc_w)
    desc.fc_b_off = push(fc_b)
    blob = np.concatenate(parts)
    info = {"arch": "ConvLSTM_w_ref" if is_lstm else "Conv_w_ref", "size": int(size),
            "kmer_len": int(desc.kmer_len), "num_out": int(desc.num_out)}
    return desc, blob, info
