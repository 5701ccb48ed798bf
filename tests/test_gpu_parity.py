"""Parity tests proper (run on the B200 box): every call goes through the C-ABI library.

Tolerances: k-mer one-hot encode bit-exact; float32 logits <= 1e-4 max-abs against the reference's
own CPU logits (BASELINE.json north_star), in practice ~1e-6; per-layer activations <= 2e-5."""
import os
import threading

import numpy as np
import pytest
import torch

import remora_oracle as ro
from conftest import GOLDEN, load_golden_model, unpack_encode_case
from remora_b200 import RemoraError, data_chunks, encoded_kmers, inference, model_util
from remora_b200.synth import synth_chunks, synth_read

pytestmark = pytest.mark.gpu
LOGIT_TOL = 1e-4  # north_star tolerance

_MODELS = {}


def gpu_model(name):
    if name not in _MODELS:
        _MODELS[name] = model_util.load_model(os.path.join(GOLDEN, name + ".pt"),
                                              device=torch.device("cuda:0"), eval_only=True)
    return _MODELS[name]


def impls_for(model, T=100):
    """CUDA paths to check for this model: the plain and the register-tiled layer kernels always; for
    ConvLSTM_w_ref/64 also the fused fp32 FFMA2 kernels, the tcgen05 (3xTF32) variant and - for
    chunk_len <= 100 - the single-kernel path (fp16 hi/lo split operands on tcgen05, fp32 parity)."""
    if model.info["arch"] == "Conv_w_ref" and model.info["size"] == 64 and T in (100, 200):
        return ["layers", "tiled", "fused_mega"]  # Conv_w_ref has its own single kernel
    if not (model.info["arch"] == "ConvLSTM_w_ref" and model.info["size"] == 64 and fused_available(model)):
        return ["layers", "tiled"]
    return ["layers", "tiled", "fused", "fused_tc"] + (["fused_mega"] if T <= 100 else [])


def fused_available(model):
    try:
        model.set_impl("fused")
        model.set_impl("auto")
        return True
    except RemoraError:
        return False


# ---------------------------------------------------------------------------------------------
# encoder: bit-exact
# ---------------------------------------------------------------------------------------------
def test_encode_known_answer(encode_cases):
    got = encoded_kmers.compute_encoded_kmer_batch(4, 4, encode_cases["kat_seqs"],
                                                   encode_cases["kat_maps"],
                                                   encode_cases["kat_lens"])
    assert got.dtype == np.float32 and np.array_equal(got, encode_cases["kat_out"])


def test_encode_bit_exact_vs_golden(encode_cases):
    for cid, kb, ka, T in encode_cases["cases"]:
        seqs, maps, lens, want = unpack_encode_case(encode_cases, cid)
        got = encoded_kmers.compute_encoded_kmer_batch(int(kb), int(ka), seqs, maps, lens)
        assert got.shape == want.shape and np.array_equal(got, want), f"case {cid} k=({kb},{ka}) T={T}"


@pytest.mark.parametrize("n,T,ctx", [(1, 100, (4, 4)), (1000, 100, (4, 4)), (333, 200, (4, 4)),
                                     (77, 400, (2, 3)), (50, 1000, (4, 4)), (40, 3000, (1, 1))])
def test_encode_bit_exact_vs_oracle(n, T, ctx):
    d = synth_chunks(n, T, ctx, seed=n + T, frac_n=0.03, frac_edge=0.2)
    want = ro.encode_kmers_c(ctx[0], ctx[1], d["sequence"], d["sequence_to_signal_mapping"],
                             d["sequence_lengths"])
    got = encoded_kmers.compute_encoded_kmer_batch(ctx[0], ctx[1], d["sequence"],
                                                   d["sequence_to_signal_mapping"],
                                                   d["sequence_lengths"])
    assert np.array_equal(got, want)


def test_encode_full_size_properties():
    """BASELINE-size batch (8192 x 36 x 100): size-independent properties of the one-hot tensor."""
    n, T = 8192, 100
    d = synth_chunks(n, T, (4, 4), seed=4242)
    out = encoded_kmers.compute_encoded_kmer_batch_torch(
        4, 4, torch.from_numpy(d["sequence"]), torch.from_numpy(d["sequence_to_signal_mapping"]),
        torch.from_numpy(d["sequence_lengths"]), sig_len=T, device=torch.device("cuda:0"))
    assert out.shape == (n, 36, T)
    assert bool(((out == 0) | (out == 1)).all())
    per_group = out.view(n, 9, 4, T).sum(dim=2)  # at most one base per (k-mer offset, time)
    assert float(per_group.max()) == 1.0
    # column sums: number of non-N bases in the k-mer covering each sample = checksum of oracle
    want = ro.encode_kmers_c(4, 4, d["sequence"][:64], d["sequence_to_signal_mapping"][:64],
                             d["sequence_lengths"][:64])
    assert np.array_equal(out[:64].cpu().numpy(), want)
    total = int(out.sum().item())
    # checksum of checksums: total ones = sum over bases of dwell * valid k-mer entries
    exp = 0
    for c in range(n):
        L = int(d["sequence_lengths"][c])
        dw = np.diff(d["sequence_to_signal_mapping"][c, :L + 1].astype(np.int64))
        sq = d["sequence"][c]
        for p in range(9):
            exp += int(dw[sq[p:p + L] != -1].sum())
    assert total == exp


# ---------------------------------------------------------------------------------------------
# forward: logits within 1e-4 of the reference's CPU logits
# ---------------------------------------------------------------------------------------------
def _case_inputs(forward_cases, key):
    return (forward_cases[key + "__signal"], forward_cases[key + "__seqs"],
            forward_cases[key + "__maps"], forward_cases[key + "__lens"],
            forward_cases[key + "__logits"])


def test_forward_compact_vs_reference_logits(forward_cases):
    for key in forward_cases["index"]:
        key = str(key)
        model, md = gpu_model(key.split("__")[0])
        sig, seqs, maps, lens, want = _case_inputs(forward_cases, key)
        for impl in impls_for(model, sig.shape[-1]):
            model.set_impl(impl)
            got = model.forward_compact(torch.from_numpy(sig), torch.from_numpy(seqs),
                                        torch.from_numpy(maps), torch.from_numpy(lens))
            # the implementation that RAN is the one requested: every golden shape (T = 100 / 200 / 400)
            # fits the tensor-core kernels, so a silent FFMA2 fallback would fail here
            assert model.last_impl == impl, (key, impl, model.last_impl)
            err = np.abs(got.cpu().numpy() - want).max()
            assert err < LOGIT_TOL, f"{key} [{impl}] max-abs err {err}"
        model.set_impl("auto")
        if "fused_mega" in impls_for(model, sig.shape[-1]):
            assert model.get_flags(clear=True) == 0, "an activation left the fp16 range"


def test_forward_dense_vs_reference_logits(forward_cases):
    """model(sigs, enc_kmers): the reference's own call form, dense one-hot input."""
    for key in forward_cases["index"]:
        key = str(key)
        model, md = gpu_model(key.split("__")[0])
        sig, seqs, maps, lens, want = _case_inputs(forward_cases, key)
        ctx = md["kmer_context_bases"]
        enc = ro.encode_kmers_c(ctx[0], ctx[1], seqs, maps, lens)
        got = model(torch.from_numpy(sig).cuda(), torch.from_numpy(enc).cuda())
        assert got.shape == want.shape and got.dtype == torch.float32 and got.is_cuda
        assert np.abs(got.detach().cpu().numpy() - want).max() < LOGIT_TOL, key


def test_per_layer_activations_vs_oracle(forward_cases):
    """Layer kernels against the oracle's torch restatement, layer by layer."""
    key = "convlstm_s64_k9_hot__n64_T100"
    model, md = gpu_model("convlstm_s64_k9_hot")
    sd, _ = load_golden_model("convlstm_s64_k9_hot")
    sd = {k: v.float() for k, v in sd.items() if v.dtype.is_floating_point}
    sig, seqs, maps, lens, _ = _case_inputs(forward_cases, key)
    enc = torch.from_numpy(ro.encode_kmers_c(4, 4, seqs, maps, lens))
    with torch.no_grad():
        s1 = ro._conv_bn_swish(torch.from_numpy(sig), sd, "sig_conv1", "sig_bn1")
        s2 = ro._conv_bn_swish(s1, sd, "sig_conv2", "sig_bn2")
        s3 = ro._conv_bn_swish(s2, sd, "sig_conv3", "sig_bn3", stride=3)
        q1 = ro._conv_bn_swish(enc, sd, "seq_conv1", "seq_bn1")
        q2 = ro._conv_bn_swish(q1, sd, "seq_conv2", "seq_bn2", stride=3)
        cat = torch.cat((s3, q2), 1)
        m1 = ro._conv_bn_swish(cat, sd, "merge_conv1", "merge_bn")
        l1 = ro._swish(ro._lstm_forward(m1.permute(2, 0, 1), sd, "lstm1")).permute(1, 2, 0)
    for impl in ("tiled", "layers"):
        model.set_impl(impl)
        model.set_debug(True)
        model.forward_compact(torch.from_numpy(sig), torch.from_numpy(seqs), torch.from_numpy(maps),
                              torch.from_numpy(lens))
        for name, want in (("sig1", s1), ("sig2", s2), ("seq1", q1), ("cat", cat), ("merge1", m1),
                           ("lstm1", l1)):
            got = model.debug_tensor(name).cpu()
            assert got.shape == want.shape, (impl, name)
            err, scale = float((got - want).abs().max()), max(1.0, float(want.abs().max()))
            assert err < 2e-5 * (scale if impl == "tiled" else 1.0), (impl, name, err, scale)
    cat_layers = model.debug_tensor("cat").cpu()
    if "fused" in impls_for(model):
        # intermediates of the fused kernels: cat (K1 output) and the LSTM1 input projection (K2)
        model.set_impl("fused")
        model.set_debug(True)
        model.forward_compact(torch.from_numpy(sig), torch.from_numpy(seqs),
                              torch.from_numpy(maps), torch.from_numpy(lens))
        got_cat = model.debug_tensor("cat").cpu()
        assert (got_cat[:, :64] - cat[:, :64]).abs().max() < 2e-5, "fused sig track"
        assert (got_cat[:, 64:] - cat[:, 64:]).abs().max() < 2e-5, "fused seq track"
        assert (got_cat - cat_layers).abs().max() < 2e-5
        with torch.no_grad():
            want_xp = torch.einsum("rk,bkt->brt", sd["lstm1.weight_ih_l0"], m1) + \
                (sd["lstm1.bias_ih_l0"] + sd["lstm1.bias_hh_l0"])[None, :, None]
        got_xp = model.debug_tensor("xproj").cpu()
        assert got_xp.shape == want_xp.shape
        scale = float(want_xp.abs().max())  # the hot fixture's projection spans +-40
        assert float((got_xp - want_xp).abs().max()) < 1e-5 * scale + 1e-5, "fused merge conv + projection"
        model.set_impl("fused_tc")
        model.forward_compact(torch.from_numpy(sig), torch.from_numpy(seqs),
                              torch.from_numpy(maps), torch.from_numpy(lens))
        got_xp = model.debug_tensor("xproj").cpu()
        err = float((got_xp - want_xp).abs().max())
        assert err < 1e-5 * scale + 1e-5, f"tcgen05 convs + projection: err {err:.3e} at scale {scale:.1f}"
    if "fused_mega" in impls_for(model):
        # intermediates of the single-kernel path: cat (both stride-3 convs on tcgen05), merge conv
        # output, LSTM1 input projection - fp16 hi/lo split operands, fp32 accumulate
        model.set_impl("fused_mega")
        model.set_debug(True)
        model.forward_compact(torch.from_numpy(sig), torch.from_numpy(seqs),
                              torch.from_numpy(maps), torch.from_numpy(lens))
        assert model.last_impl == "fused_mega"
        got_cat = model.debug_tensor("cat").cpu()
        assert got_cat.shape == cat.shape
        assert (got_cat[:, :64] - cat[:, :64]).abs().max() < 2e-5, "single kernel: sig track"
        assert (got_cat[:, 64:] - cat[:, 64:]).abs().max() < 2e-5, "single kernel: seq track"
        got_m = model.debug_tensor("merge1").cpu()
        assert got_m.shape == m1.shape and (got_m - m1).abs().max() < 2e-5 * max(1.0, float(m1.abs().max()))
        got_xp = model.debug_tensor("xproj").cpu()
        err = float((got_xp - want_xp).abs().max())
        assert err < 1e-5 * scale + 1e-5, f"single kernel projection: err {err:.3e} at scale {scale:.1f}"
    model.set_debug(False)
    model.set_impl("auto")


def test_bf16_variant_tolerance(forward_cases):
    """BASELINE configs[1] names a bf16 ConvLSTM_w_ref: the single kernel with one-pass bf16 tensor-core
    operands (activations and weights of the four GEMM-shaped layers rounded to bf16, fp32 accumulate,
    fp32 signal convs / gates / classifier).  Its own stated tolerance against the reference's fp32 CPU
    logits on the hot fixture (logits span +-4): max-abs logit error < 0.15, max-abs probability error
    < 0.03, ML-byte mismatch (|delta| > 1) rate < 15 %, never selected by AUTO."""
    model, md = gpu_model("convlstm_s64_k9_hot")
    worst_logit = worst_prob = 0.0
    n = n_ml_bad = 0
    for key in forward_cases["index"]:
        key = str(key)
        if not key.startswith("convlstm_s64_k9_hot__") or not key.endswith("T100"):
            continue
        sig, seqs, maps, lens, want = _case_inputs(forward_cases, key)
        model.set_impl("fused_bf16")
        got = model.forward_compact(torch.from_numpy(sig), torch.from_numpy(seqs), torch.from_numpy(maps),
                                    torch.from_numpy(lens)).cpu().numpy()
        assert model.last_impl == "fused_bf16" and np.isfinite(got).all()
        from remora_b200 import util
        pg, pw = util.softmax_axis1(got)[:, 1:], util.softmax_axis1(want)[:, 1:]
        worst_logit = max(worst_logit, float(np.abs(got - want).max()))
        worst_prob = max(worst_prob, float(np.abs(pg - pw).max()))
        n_ml_bad += int((np.abs(ro.ml_bytes(pg).astype(int) - ro.ml_bytes(pw).astype(int)) > 1).sum())
        n += pg.size
    model.set_impl("auto")
    print(f"bf16 variant: max|dlogit| {worst_logit:.4f} max|dprob| {worst_prob:.4f} "
          f"ML bytes off by >1: {n_ml_bad}/{n}")
    assert n > 0 and worst_logit < 0.15 and worst_prob < 0.03 and n_ml_bad / n < 0.15


@pytest.mark.parametrize("B", [1, 2, 63, 64, 65, 1024, 4096])
def test_batch_sizes_and_ragged_batches(B):
    """Any B in [1, batch_size] (ragged last batch, inference.py:308-310); per-chunk independence:
    the logits of chunk i do not depend on the batch it is in."""
    model, md = gpu_model("convlstm_s64_k9_hot")
    sd, _ = load_golden_model("convlstm_s64_k9_hot")
    d = synth_chunks(B, 100, (4, 4), seed=B)
    args = [torch.from_numpy(d[k]) for k in ("signal", "sequence", "sequence_to_signal_mapping",
                                             "sequence_lengths")]
    outs = {}
    for impl in impls_for(model):
        model.set_impl(impl)
        outs[impl] = model.forward_compact(*args).cpu().numpy()
    model.set_impl("auto")
    n_chk = min(B, 48)
    want = ro.oracle_infer_compact(sd, (4, 4), d["signal"][:n_chk], d["sequence"][:n_chk],
                                   d["sequence_to_signal_mapping"][:n_chk],
                                   d["sequence_lengths"][:n_chk])
    for impl, got in outs.items():
        assert got.shape == (B, 2) and np.isfinite(got).all()
        assert np.abs(got[:n_chk] - want).max() < LOGIT_TOL, impl
    # independence / determinism: tail chunks re-run as their own small batch
    k = min(B, 5)
    tail = [a[B - k:] for a in args]
    for impl in impls_for(model):
        model.set_impl(impl)
        again = model.forward_compact(*tail).cpu().numpy()
        assert np.abs(again - outs[impl][B - k:]).max() < 2e-6, impl
    model.set_impl("auto")


@pytest.mark.parametrize("T,max_seq_len,lens,stride", [
    (100, 20, None, 5), (100, 20, (15, 20), 3), (100, 21, None, 5), (100, 40, (25, 40), 2), (100, 100, (60, 100), 1),
    (64, 30, (10, 30), 2), (37, 12, None, 5)])
def test_single_kernel_mapping_widths(T, max_seq_len, lens, stride):
    """The single kernel has two forms of seq_conv1: per-base sums (up to 20 bases per chunk fit their buffer)
    and, for wider mapping arrays, the layer as one implicit GEMM; both, at several chunk lengths, against
    the oracle, and the AUTO choice must stay on the single kernel."""
    model, md = gpu_model("convlstm_s64_k9_hot")
    sd, _ = load_golden_model("convlstm_s64_k9_hot")
    B = 37
    d = synth_chunks(B, T, (4, 4), seed=1000 * T + max_seq_len, max_seq_len=max_seq_len, seq_len_range=lens,
                     stride=stride)
    args = [torch.from_numpy(d[k]) for k in ("signal", "sequence", "sequence_to_signal_mapping",
                                             "sequence_lengths")]
    want = ro.oracle_infer_compact(sd, (4, 4), d["signal"], d["sequence"],
                                   d["sequence_to_signal_mapping"], d["sequence_lengths"])
    model.set_impl("auto")
    got = model.forward_compact(*args).cpu().numpy()
    assert model.last_impl == "fused_mega"
    err = float(np.abs(got - want).max())
    assert err < LOGIT_TOL, f"T={T} max_seq_len={max_seq_len}: max-abs err {err:.3e}"


def test_full_baseline_batch_against_oracle():
    """The whole BASELINE batch (1024 chunks, T=100) against the oracle (the reference's arithmetic on
    CPU), every chunk, for every CUDA implementation - not a prefix and not another CUDA path."""
    model, md = gpu_model("convlstm_s64_k9_hot")
    sd, _ = load_golden_model("convlstm_s64_k9_hot")
    d = synth_chunks(1024, 100, (4, 4), seed=20261017)
    args = [torch.from_numpy(d[k]) for k in ("signal", "sequence", "sequence_to_signal_mapping",
                                             "sequence_lengths")]
    want = ro.oracle_infer_compact(sd, (4, 4), d["signal"], d["sequence"],
                                   d["sequence_to_signal_mapping"], d["sequence_lengths"])
    assert want.shape == (1024, 2) and float(np.abs(want).max()) > 1.0  # logits span several units
    for impl in impls_for(model):
        model.set_impl(impl)
        got = model.forward_compact(*args).cpu().numpy()
        assert model.last_impl == impl
        err = float(np.abs(got - want).max())
        assert err < LOGIT_TOL, f"[{impl}] max-abs err {err:.3e} over 1024 chunks"
    model.set_impl("auto")


@pytest.mark.parametrize("name,T,B", [("conv_s64_k9", 100, 1), ("conv_s64_k9", 100, 7),
                                      ("conv_s64_k9", 100, 1024), ("conv_s64_k9", 100, 4099),
                                      ("convlstm_s16_k6_o3", 60, 33), ("convlstm_s16_k6_o3", 400, 50),
                                      ("convlstm_s64_k9_hot", 200, 130),
                                      ("convlstm_s64_k9_hot", 400, 70)])
def test_tiled_layer_kernels_match_plain_layer_kernels(name, T, B):
    """Register-tiled FFMA2 convolutions + gather-form seq_conv1 against the one-thread-per-output
    kernels, layer by layer (both compact and dense input), over shapes that exercise every tile plan."""
    model, md = gpu_model(name)
    ctx = tuple(md["kmer_context_bases"])
    d = synth_chunks(B, T, ctx, seed=T + B)
    args = [torch.from_numpy(d[k]) for k in ("signal", "sequence", "sequence_to_signal_mapping",
                                             "sequence_lengths")]
    layer_names = ["sig1", "sig2", "sig3", "seq1", "seq2", "seq3", "cat", "merge1", "merge2", "merge3",
                   "merge4"]
    kept, logits = {}, {}
    for impl in ("layers", "tiled"):
        model.set_impl(impl)
        model.set_debug(True)
        logits[impl] = model.forward_compact(*args).cpu()
        assert model.last_impl == impl
        kept[impl] = {}
        for ln in layer_names:
            try:
                kept[impl][ln] = model.debug_tensor(ln).cpu()
            except RemoraError:
                pass
    model.set_debug(False)
    assert set(kept["layers"]) == set(kept["tiled"]) and "cat" in kept["tiled"]
    for ln, want in kept["layers"].items():
        scale = max(1.0, float(want.abs().max()))
        assert float((kept["tiled"][ln] - want).abs().max()) < 2e-5 * scale, ln
    assert float((logits["tiled"] - logits["layers"]).abs().max()) < 5e-5
    # dense input through the tiled kernels (no gather form for seq_conv1)
    enc = torch.from_numpy(ro.encode_kmers_c(ctx[0], ctx[1], d["sequence"],
                                             d["sequence_to_signal_mapping"],
                                             d["sequence_lengths"])).cuda()
    model.set_impl("tiled")
    dense = model(args[0].cuda(), enc).cpu()
    model.set_impl("auto")
    assert float((dense - logits["layers"]).abs().max()) < 5e-5


@pytest.mark.parametrize("B", [1, 3, 300, 4099])
def test_conv_w_ref_single_kernel_matches_oracle(B):
    """Conv_w_ref (stock chunk_len 100): AUTO = the single kernel (seven GEMM-shaped layers on tcgen05 with
    fp16 hi/lo split operands); every CUDA path against the oracle on every chunk, intermediates of the
    single kernel (cat, last merge conv) against the plain layer kernels."""
    model, md = gpu_model("conv_s64_k9")
    sd, _ = load_golden_model("conv_s64_k9")
    d = synth_chunks(B, 100, (4, 4), seed=5 + B)
    args = [torch.from_numpy(d[k]) for k in ("signal", "sequence", "sequence_to_signal_mapping",
                                             "sequence_lengths")]
    want = ro.oracle_infer_compact(sd, (4, 4), d["signal"], d["sequence"],
                                   d["sequence_to_signal_mapping"], d["sequence_lengths"])
    model.set_impl("auto")
    got = model.forward_compact(*args).cpu().numpy()
    assert model.last_impl == "fused_mega"
    assert np.abs(got - want).max() < LOGIT_TOL
    kept = {}
    for impl in ("layers", "fused_mega", "tiled"):
        model.set_impl(impl)
        model.set_debug(True)
        out = model.forward_compact(*args).cpu().numpy()
        assert model.last_impl == impl and np.abs(out - want).max() < LOGIT_TOL, impl
        if impl != "tiled":
            kept[impl] = {n: model.debug_tensor(n).cpu() for n in ("cat", "merge4")}
    model.set_debug(False)
    model.set_impl("auto")
    for n in ("cat", "merge4"):
        a, b = kept["fused_mega"][n], kept["layers"][n]
        assert a.shape == b.shape and float((a - b).abs().max()) < 2e-5 * max(1.0, float(b.abs().max())), n
    assert model.get_flags(clear=True) == 0


def test_conv_w_ref_chunk_len_200_matches_reference():
    """BASELINE config 3 as literally stated: Conv_w_ref at chunk_len 200 (classifier of 64*11 inputs),
    logits of the reference's TorchScript module on CPU; compact and dense interfaces, plain and tiled
    layer kernels, batch 33 / 1 / 4096 (the config's batch size: finite and batch-invariant)."""
    model, md = gpu_model("conv_s64_k9_T200")
    assert md["chunk_len"] == 200
    g = np.load(os.path.join(GOLDEN, "conv_T200_cases.npz"))
    for key in ("n33", "n1"):
        args = [torch.from_numpy(g[key + k]) for k in ("_signal", "_seqs", "_maps", "_lens")]
        for impl in ("layers", "tiled", "fused_mega", "auto"):
            model.set_impl(impl)
            got = model.forward_compact(*args).cpu().numpy()
            assert model.last_impl == ("fused_mega" if impl == "auto" else impl)
            assert np.abs(got - g[key + "_logits"]).max() < LOGIT_TOL, (key, impl)
        enc = encoded_kmers.compute_encoded_kmer_batch(4, 4, g[key + "_seqs"], g[key + "_maps"], g[key + "_lens"])
        dense = model(torch.from_numpy(g[key + "_signal"]).cuda(), torch.from_numpy(enc).cuda()).cpu().numpy()
        assert np.abs(dense - g[key + "_logits"]).max() < LOGIT_TOL
    d = synth_chunks(4096, 200, (4, 4), seed=77)
    args = [torch.from_numpy(d[k]) for k in ("signal", "sequence", "sequence_to_signal_mapping",
                                             "sequence_lengths")]
    model.set_impl("auto")
    big = model.forward_compact(*args).cpu().numpy()
    assert np.isfinite(big).all()
    part = model.forward_compact(*[a[1000:1033] for a in args]).cpu().numpy()
    assert np.array_equal(big[1000:1033], part)  # batch-invariant kernels


def test_empty_batch():
    model, _ = gpu_model("convlstm_s64_k9")
    out = model.forward_compact(torch.zeros((0, 1, 100)), torch.zeros((0, 28), dtype=torch.int8),
                                torch.zeros((0, 21), dtype=torch.int16),
                                torch.zeros((0,), dtype=torch.int16))
    assert out.shape == (0, 2)


def test_conv_w_ref_rejects_other_chunk_len():
    """Stock Conv_w_ref only accepts chunk_len 100 (models/Conv_w_ref.py:42): loud error."""
    model, _ = gpu_model("conv_s64_k9")
    d = synth_chunks(4, 200, (4, 4), seed=1)
    with pytest.raises(RemoraError):
        model.forward_compact(torch.from_numpy(d["signal"]), torch.from_numpy(d["sequence"]),
                              torch.from_numpy(d["sequence_to_signal_mapping"]),
                              torch.from_numpy(d["sequence_lengths"]))


def test_bad_arguments_raise():
    model, _ = gpu_model("convlstm_s64_k9")
    with pytest.raises(RemoraError):
        model(torch.zeros((2, 1, 100)).cuda(), torch.zeros((2, 35, 100)).cuda())
    with pytest.raises(RemoraError):
        model(torch.zeros((2, 1, 100), dtype=torch.float64).cuda(),
              torch.zeros((2, 36, 100)).cuda())
    with pytest.raises(RemoraError):  # shorter than the receptive field
        model(torch.zeros((2, 1, 12)).cuda(), torch.zeros((2, 36, 12)).cuda())


def test_model_quacks_like_the_reference_module():
    model, md = gpu_model("convlstm_s64_k9")
    assert next(model.parameters()).device == torch.device("cuda:0")
    assert all(not p.requires_grad for p in model.parameters())
    assert model.eval() is model
    assert md["chunk_len"] == 100 and md["kmer_len"] == 9


# ---------------------------------------------------------------------------------------------
# boundary: call_read_mods / run_model_batched / infer_host / softmax+ML
# ---------------------------------------------------------------------------------------------
def test_call_read_mods_matches_reference(read_cases):
    meta, g = read_cases
    for i, m in enumerate(meta):
        model, md = gpu_model(m["model"])
        def fresh():
            return data_chunks.RemoraRead(dacs=g[f"r{i}_dacs"], shift=m["shift"],
                                          scale=m["scale"], seq_to_sig_map=g[f"r{i}_ssm"],
                                          int_seq=g[f"r{i}_int_seq"])
        order = np.argsort(g[f"r{i}_pos"])
        nn_out, labels, pos = inference.call_read_mods(fresh(), model, md)
        assert np.array_equal(pos, g[f"r{i}_pos"][order])
        assert np.array_equal(labels, g[f"r{i}_labels"][order])
        assert np.abs(nn_out - g[f"r{i}_nn_out"][order]).max() < LOGIT_TOL
        probs, _, _ = inference.call_read_mods(fresh(), model, md, return_mod_probs=True)
        assert probs.dtype == np.float64
        assert np.abs(probs - g[f"r{i}_probs"][order]).max() < LOGIT_TOL
        mm, ml = inference.call_read_mods(fresh(), model, md, return_mm_ml_tags=True)
        assert mm == m["mm"]
        ml = np.frombuffer(ml, dtype=np.uint8).astype(int)
        # a 1e-6 logit difference can move a probability across a 1/256 bin edge
        diff = np.abs(ml - g[f"r{i}_ml"].astype(int))
        assert diff.max() <= 1 and (diff != 0).mean() <= 0.02


def test_gpu_chunk_extraction_is_bit_identical(read_cases):
    """rb200_chunk_plan / rb200_chunk_fill against the host chunk extraction (itself pinned to the
    reference's call_read_mods by tests/test_host.py): all four compact arrays bit for bit, for int16,
    float32 and float64 DAC arrays, short reads with padding on both sides, N bases."""
    from remora_b200 import util
    _, md = load_golden_model("convlstm_s64_k9")
    cases = []
    for n_bases, seed, dt in ((400, 1, np.int16), (37, 2, np.int16), (8, 3, np.int16),
                              (250, 4, np.float32), (120, 5, np.float64), (300, 6, np.int32)):
        dacs, shift, scale, ssm, int_seq = synth_read(n_bases, seed=seed, frac_n=0.02)
        cases.append((dacs.astype(dt), shift, scale, ssm, int_seq))
    for dacs, shift, scale, ssm, int_seq in cases:
        for focus_kind in ("motif", "all"):
            def fresh():
                r = data_chunks.RemoraRead(dacs=dacs, shift=shift, scale=scale, seq_to_sig_map=ssm,
                                           int_seq=int_seq)
                if focus_kind == "motif":
                    r.set_motif_focus_bases([util.Motif("C", 0)])
                else:
                    r.focus_bases = np.arange(int_seq.size)
                return r
            host, dev = fresh(), fresh()
            host.prepare_batches(md, 64)
            dev.prepare_batches_gpu(md, 64, device=torch.device("cuda:0"))
            assert len(host.batches) == len(dev.batches)
            for hb, db in zip(host.batches, dev.batches):
                assert np.array_equal(hb.signal, db.signal.cpu().numpy())
                assert np.array_equal(hb.seq_lens, db.seq_lens.cpu().numpy())
                assert np.array_equal(hb.seq_to_sig_map, db.seq_to_sig_map.cpu().numpy())
                assert np.array_equal(hb.sequence, db.sequence.cpu().numpy())
                assert np.array_equal(hb.read_focus_bases, db.read_focus_bases)


def test_call_read_mods_with_device_extraction(read_cases):
    meta, g = read_cases
    for i, m in enumerate(meta):
        model, md = gpu_model(m["model"])
        read = data_chunks.RemoraRead(dacs=g[f"r{i}_dacs"], shift=m["shift"], scale=m["scale"],
                                      seq_to_sig_map=g[f"r{i}_ssm"], int_seq=g[f"r{i}_int_seq"])
        nn_out, labels, pos = inference.call_read_mods(read, model, md, extract_on_device=True)
        order = np.argsort(g[f"r{i}_pos"])
        assert np.array_equal(pos, g[f"r{i}_pos"][order])
        assert np.abs(nn_out - g[f"r{i}_nn_out"][order]).max() < LOGIT_TOL


def test_reference_test_read(read_cases):
    """scripts/api_example.py flow: load_model -> RemoraRead.test_read() -> call_read_mods."""
    _, g = read_cases
    model, md = gpu_model("convlstm_s64_k9")
    nn_out, labels, pos = inference.call_read_mods(data_chunks.RemoraRead.test_read(), model, md)
    order = np.argsort(g["test_read_pos"])
    assert np.array_equal(pos, g["test_read_pos"][order])
    assert np.abs(nn_out - g["test_read_nn_out"][order]).max() < LOGIT_TOL


def test_run_model_batched_generator(forward_cases):
    key = "convlstm_s64_k9_hot__n64_T100"
    model, md = gpu_model("convlstm_s64_k9_hot")
    sig, seqs, maps, lens, want = _case_inputs(forward_cases, key)
    enc = ro.encode_kmers_c(4, 4, seqs, maps, lens)
    items = [("C", sig[:32], enc[:32], np.arange(32), ["r"] * 32),   # full batch -> pinned path
             ("C", sig[32:50], enc[32:50], np.arange(18), ["r"] * 18)]  # ragged
    outs = list(inference.run_model_batched(items, {"C": model}, [md], batch_size=32))
    got = np.concatenate([o[1].cpu().numpy() for o in outs])
    assert np.abs(got - want[:50]).max() < LOGIT_TOL


def test_infer_host_roundtrip(forward_cases):
    key = "convlstm_s64_k9_hot__n64_T100"
    model, _ = gpu_model("convlstm_s64_k9_hot")
    sig, seqs, maps, lens, want = _case_inputs(forward_cases, key)
    got = model.infer_host(sig, seqs, maps, lens)
    assert np.abs(got - want).max() < LOGIT_TOL
    got2 = model.infer_host(sig[:3], seqs[:3], maps[:3], lens[:3])  # smaller batch reuses staging
    assert np.abs(got2 - want[:3]).max() < LOGIT_TOL


def test_infer_host_async_pipelined(forward_cases):
    """rb200_infer_host_async: pinned buffers, two alternating streams, results per slot."""
    key = "convlstm_s64_k9_hot__n64_T100"
    model, _ = gpu_model("convlstm_s64_k9_hot")
    sig, seqs, maps, lens, want = _case_inputs(forward_cases, key)
    pins = [torch.from_numpy(a).pin_memory() for a in (sig, seqs, maps, lens)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [torch.zeros((64, 2)).pin_memory() for _ in range(2)]
    for i in range(6):
        slot = i & 1
        streams[slot].synchronize()
        if i >= 2:
            assert np.abs(outs[slot].numpy() - want).max() < LOGIT_TOL
        outs[slot].zero_()
        model.infer_host_async(*pins, outs[slot], stream=streams[slot])
    for slot in range(2):
        streams[slot].synchronize()
        assert np.abs(outs[slot].numpy() - want).max() < LOGIT_TOL
    with pytest.raises(RemoraError):  # pageable buffers are refused
        model.infer_host_async(torch.from_numpy(sig), *pins[1:], outs[0])
    # the four arrays in one pinned block (one host-to-device copy instead of four): same logits, and the views
    # really are laid out the way rb200_infer_host_async recognises
    packed = model.pinned_batch(sig, seqs, maps, lens)
    up = lambda n: (n + 255) // 256 * 256  # noqa: E731
    sizes = [t.numel() * t.element_size() for t in packed]
    assert packed[1].data_ptr() == packed[0].data_ptr() + up(sizes[0])
    assert packed[2].data_ptr() == packed[1].data_ptr() + up(sizes[1])
    assert packed[3].data_ptr() == packed[2].data_ptr() + up(sizes[2])
    assert all(t.is_pinned() for t in packed)
    outs[0].zero_()
    model.infer_host_async(*packed, outs[0], stream=streams[0])
    streams[0].synchronize()
    assert np.abs(outs[0].numpy() - want).max() < LOGIT_TOL


def test_softmax_ml_kernel():
    import ctypes
    from remora_b200 import _native, util
    lib = _native.load_library()
    rng = np.random.default_rng(0)
    logits = (rng.standard_normal((1000, 3)) * 4).astype(np.float32)
    logits[0] = [0, 50, -50]  # p -> 1.0 must clip to 255
    lg = torch.from_numpy(logits).cuda()
    probs = torch.empty((1000, 2), dtype=torch.float32, device="cuda")
    ml = torch.empty((1000, 2), dtype=torch.uint8, device="cuda")
    rc = lib.rb200_softmax_ml(ctypes.c_void_p(lg.data_ptr()), 1000, 3,
                              ctypes.c_void_p(probs.data_ptr()), ctypes.c_void_p(ml.data_ptr()),
                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _native.check(rc, "rb200_softmax_ml")
    want_p = util.softmax_axis1(logits)[:, 1:]
    assert np.abs(probs.cpu().numpy() - want_p).max() < 1e-6
    want_ml = ro.ml_bytes(want_p)
    diff = np.abs(ml.cpu().numpy().astype(int) - want_ml.astype(int))
    assert diff.max() <= 1 and (diff != 0).mean() < 0.01 and ml[0, 0] == 255


def test_concurrent_threads_share_one_model(forward_cases):
    """Duplex inference calls one model from several threads (inference.py:973-982)."""
    key = "convlstm_s64_k9_hot__n64_T100"
    model, _ = gpu_model("convlstm_s64_k9_hot")
    sig, seqs, maps, lens, want = _case_inputs(forward_cases, key)
    errs = []

    def worker(seed):
        try:
            rng = np.random.default_rng(seed)
            for _ in range(20):
                idx = rng.permutation(64)[: rng.integers(1, 64)]
                got = model.forward_compact(torch.from_numpy(sig[idx]), torch.from_numpy(seqs[idx]),
                                            torch.from_numpy(maps[idx]),
                                            torch.from_numpy(lens[idx])).cpu().numpy()
                if np.abs(got - want[idx]).max() >= LOGIT_TOL:
                    errs.append("mismatch")
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(4)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errs, errs


def test_full_size_batch_properties():
    """BASELINE-size step (1024 chunks, T=100) and a large one (8192): permutation equivariance
    (chunks are independent) and agreement between the two CUDA paths."""
    model, _ = gpu_model("convlstm_s64_k9_hot")
    for B in (1024, 8192):
        d = synth_chunks(B, 100, (4, 4), seed=B + 1)
        args = [torch.from_numpy(d[k]).cuda() for k in
                ("signal", "sequence", "sequence_to_signal_mapping", "sequence_lengths")]
        out = model.forward_compact(*args)
        perm = torch.randperm(B, device="cuda")
        out_p = model.forward_compact(*[a[perm] for a in args])
        assert (out[perm] - out_p).abs().max() < 2e-6
        if "fused" in impls_for(model):
            model.set_impl("layers")
            ref = model.forward_compact(*args)
            assert (out - ref).abs().max() < LOGIT_TOL  # two fp32 paths, both within tolerance of the oracle
            model.set_impl("fused_tc")
            out_tc = model.forward_compact(*args)
            assert (out_tc - ref).abs().max() < LOGIT_TOL  # tensor-core (3xTF32) path
            model.set_impl("fused_mega")
            out_mega = model.forward_compact(*args)
            model.set_impl("auto")
            assert model.last_impl == "fused_mega"
            assert (out_mega - ref).abs().max() < LOGIT_TOL  # single kernel, fp16 hi/lo split operands
