"""The oracle (oracle/) against the committed golden vectors produced by the reference, and -
when /root/reference is present - against the reference itself.  CPU only."""
import os

import numpy as np
import pytest

import remora_oracle as ro
from conftest import load_golden_model, unpack_encode_case


def test_encoder_known_answer(encode_cases):
    # SURVEY.md §8c KAT, frozen from the reference's Cython encoder
    seqs, maps, lens = encode_cases["kat_seqs"], encode_cases["kat_maps"], encode_cases["kat_lens"]
    want = encode_cases["kat_out"]
    assert want.shape == (1, 36, 5) and want.sum() == 40
    for t, rows in ((0, [0, 5, 10, 15, 20, 25, 30, 35]), (4, [1, 6, 11, 16, 21, 26, 31, 32])):
        assert np.flatnonzero(want[0, :, t]).tolist() == rows
    for fn in (ro.encode_kmers_numpy, ro.encode_kmers_c):
        assert np.array_equal(fn(4, 4, seqs, maps, lens), want)


def test_encoder_restatements_match_golden(encode_cases):
    for cid, kb, ka, T in encode_cases["cases"]:
        seqs, maps, lens, want = unpack_encode_case(encode_cases, cid)
        assert want.shape == (len(lens), 4 * (kb + ka + 1), T)
        got_c = ro.encode_kmers_c(int(kb), int(ka), seqs, maps, lens)
        assert got_c.dtype == np.float32 and np.array_equal(got_c, want), f"C case {cid}"
        if cid % 4 == 0:  # python loops are slow; a quarter of the cases is enough
            assert np.array_equal(ro.encode_kmers_numpy(int(kb), int(ka), seqs, maps, lens), want)


def test_compiled_reference_encoder_matches_golden(encode_cases):
    ref = ro.load_ref_encoder()
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    for cid, kb, ka, T in encode_cases["cases"]:
        seqs, maps, lens, want = unpack_encode_case(encode_cases, cid)
        got = ref.compute_encoded_kmer_batch(int(kb), int(ka), seqs, maps, lens)
        assert np.array_equal(np.asarray(got), want)


def test_forward_restatement_matches_reference_logits(forward_cases):
    # fp32 re-association only: the restatement un-fuses nothing, tolerance 5e-6
    for key in forward_cases["index"]:
        key = str(key)
        sd, md = load_golden_model(key.split("__")[0])
        got = ro.oracle_infer_compact(sd, md["kmer_context_bases"], forward_cases[key + "__signal"],
                                      forward_cases[key + "__seqs"], forward_cases[key + "__maps"],
                                      forward_cases[key + "__lens"])
        want = forward_cases[key + "__logits"]
        assert got.shape == want.shape
        assert np.abs(got - want).max() < 5e-6, key


def test_golden_logits_depend_on_the_input(forward_cases):
    # guards the fixtures themselves: the "hot" models must spread logits far beyond the 1e-4
    # tolerance, otherwise a parity test could pass on a kernel that ignores its input
    for key in forward_cases["index"]:
        key = str(key)
        if "_hot" in key and "n64" in key:
            assert forward_cases[key + "__logits"].std(axis=0).min() > 0.1


def test_lstm2_single_step_shortcut_is_exact(forward_cases):
    """only the first step of the reversed lstm2 pass is consumed."""
    import torch
    key = "convlstm_s64_k9_hot__n64_T100"
    sd, md = load_golden_model("convlstm_s64_k9_hot")
    sd = {k: v.float() for k, v in sd.items() if v.dtype.is_floating_point}
    enc = ro.encode_kmers_c(4, 4, forward_cases[key + "__seqs"], forward_cases[key + "__maps"],
                            forward_cases[key + "__lens"])
    sig = torch.from_numpy(forward_cases[key + "__signal"])
    with torch.no_grad():
        s = ro._conv_bn_swish(sig, sd, "sig_conv1", "sig_bn1")
        s = ro._conv_bn_swish(s, sd, "sig_conv2", "sig_bn2")
        s = ro._conv_bn_swish(s, sd, "sig_conv3", "sig_bn3", stride=3)
        q = ro._conv_bn_swish(torch.from_numpy(enc), sd, "seq_conv1", "seq_bn1")
        q = ro._conv_bn_swish(q, sd, "seq_conv2", "seq_bn2", stride=3)
        z = ro._conv_bn_swish(torch.cat((s, q), 1), sd, "merge_conv1", "merge_bn").permute(2, 0, 1)
        h1 = ro._lstm_forward(z, sd, "lstm1")
        x = ro._swish(h1[-1])
        g = x @ sd["lstm2.weight_ih_l0"].T + sd["lstm2.bias_ih_l0"] + sd["lstm2.bias_hh_l0"]
        H = 64
        c = torch.sigmoid(g[:, :H]) * torch.tanh(g[:, 2 * H:3 * H])
        h2 = torch.sigmoid(g[:, 3 * H:]) * torch.tanh(c)
        out = ro._swish(h2) @ sd["fc.weight"].T + sd["fc.bias"]
    assert np.abs(out.numpy() - forward_cases[key + "__logits"]).max() < 5e-6


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/remora"), reason="no reference tree")
def test_oracle_against_live_reference():
    """Fresh seeded batch through the reference itself (Cython encoder + TorchScript module)."""
    import torch
    import ref_harness
    ref_harness.import_reference()
    from remora import encoded_kmers, model_util as ref_model_util
    from remora_b200.synth import synth_chunks
    here = os.path.join(os.path.dirname(__file__), "golden")
    model, md = ref_model_util.load_model(os.path.join(here, "convlstm_s64_k9_hot.pt"),
                                          eval_only=True)
    d = synth_chunks(50, 100, (4, 4), seed=777)
    enc = encoded_kmers.compute_encoded_kmer_batch(4, 4, d["sequence"],
                                                   d["sequence_to_signal_mapping"],
                                                   d["sequence_lengths"])
    assert np.array_equal(enc, ro.encode_kmers_c(4, 4, d["sequence"],
                                                 d["sequence_to_signal_mapping"],
                                                 d["sequence_lengths"]))
    with torch.no_grad():
        want = model(torch.from_numpy(d["signal"]), torch.from_numpy(enc)).numpy()
    got = ro.forward_from_state_dict(model.state_dict(), d["signal"], enc).numpy()
    assert np.abs(got - want).max() < 5e-6


def test_oracle_conv_w_ref_chunk_len_200():
    """BASELINE config 3 as stated (Conv_w_ref, chunk_len 200): classifier widened to 64*11 inputs in the
    reference module before its own exporter scripted it (tests/golden/make_golden.py)."""
    import remora_oracle as ro
    from conftest import GOLDEN, load_golden_model
    sd, md = load_golden_model("conv_s64_k9_T200")
    assert md["chunk_len"] == 200 and sd["fc.weight"].shape == (2, 704)
    g = np.load(os.path.join(GOLDEN, "conv_T200_cases.npz"))
    for key in ("n33", "n1"):
        enc = ro.encode_kmers_c(4, 4, g[key + "_seqs"], g[key + "_maps"], g[key + "_lens"])
        out = ro.forward_from_state_dict(sd, g[key + "_signal"], enc).numpy()
        assert np.abs(out - g[key + "_logits"]).max() < 2e-6
