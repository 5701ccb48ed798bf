"""Host-side logic (no GPU): sequence coding, motifs, tags, metadata derivation, weight packing,
vectorised chunk extraction against the reference's golden call_read_mods outputs."""
import ctypes
import os

import numpy as np
import pytest
import torch

import remora_oracle as ro
from conftest import load_golden_model
from remora_b200 import RemoraError, _native, data_chunks, model_util, util, weights


def test_seq_coding_roundtrip():
    s = "ACGTNACGTTGCA"
    ints = util.seq_to_int(s)
    assert ints.tolist() == [0, 1, 2, 3, -1, 0, 1, 2, 3, 3, 2, 1, 0]
    assert util.int_to_seq(ints) == s
    assert util.int_to_seq(np.array([], dtype=int)) == ""
    with pytest.raises(RemoraError):
        util.int_to_seq(np.array([5]))


def test_motif_findall_and_focus_bases():
    seq = util.seq_to_int("ACGCGTTCGNCGA")
    cg = util.Motif("CG", 0)
    assert cg.findall(seq).tolist() == [1, 3, 7, 10]
    assert util.Motif("NCG", 1).to_tuple() == ("CG", 0)  # leading N clipped
    assert util.Motif("CGN", 0).to_tuple() == ("CG", 0)
    drach = util.Motif("DRACH", 2)
    hits = drach.findall(util.seq_to_int("GGACTAGACATT"))
    assert hits.tolist() == [0, 5]
    fb = util.find_focus_bases_in_int_sequence(seq, [cg, util.Motif("C", 0)])
    assert fb.tolist() == [1, 3, 7, 10]
    assert cg.match(seq, 1) and not cg.match(seq, 2)
    with pytest.raises(RemoraError):
        util.Motif("CZ", 0)
    with pytest.raises(RemoraError):
        util.Motif("CG", 2)


def test_softmax_and_ml_rounding():
    x = np.array([[0.0, 0.0], [10.0, -10.0], [-3.0, 3.0]], dtype=np.float32)
    p = util.softmax_axis1(x)
    assert np.allclose(p.sum(axis=1), 1.0) and p.dtype == np.float32
    mm, ml = util.format_mm_ml_tags("ACGCG", [1, 3], np.array([[0.5], [1.0]]), "m", "C")
    assert mm == "C+m?,0,0;" and list(ml) == [128, 255]  # floor(p*256), 256 -> 255


def test_mm_ml_tags_match_reference(read_cases):
    meta, g = read_cases
    for i, m in enumerate(meta):
        _, md = load_golden_model(m["model"])
        probs = util.softmax_axis1(g[f"r{i}_nn_out"])[:, 1:].astype(np.float64)
        assert np.array_equal(probs, g[f"r{i}_probs"])
        mm, ml = util.format_mm_ml_tags(m["str_seq"], g[f"r{i}_pos"], probs, md["mod_bases"],
                                        md["can_base"])
        assert mm == m["mm"]
        assert np.array_equal(np.frombuffer(ml, dtype=np.uint8), g[f"r{i}_ml"])


def test_derived_metadata():
    _, md = load_golden_model("convlstm_s64_k9")
    assert md["kmer_len"] == 9 and md["chunk_len"] == 100 and md["can_base"] == "C"
    assert md["motifs"] == [("CG", 0)] and md["motif"] == ("CG", 0)
    assert md["mod_long_names"] == ["5mC"] and md["base_start_justify"] is False
    assert md["offset"] == 0 and not md["sig_map_refiner"].is_loaded
    assert not any(k.startswith("refine_") for k in md)
    # older Dorado-style key layout (reference model_util.py:362-379)
    old = {"mod_bases": "m", "mod_long_names_0": "5mC", "kmer_context_bases_0": "4",
           "kmer_context_bases_1": "4", "chunk_context_0": "50", "chunk_context_1": "50",
           "motif": "CG", "motif_offset": "0"}
    model_util.add_derived_metadata(old)
    assert old["kmer_context_bases"] == (4, 4) and old["chunk_len"] == 100
    assert old["motifs"] == [("CG", 0)] and old["reverse_signal"] is False


def test_load_model_errors():
    with pytest.raises(RemoraError):
        model_util.load_model("/nonexistent/model.pt")
    with pytest.raises(RemoraError):
        model_util.load_model()
    if not torch.cuda.is_available():
        with pytest.raises(RemoraError):  # no silent CPU fallback
            model_util.load_model(os.path.join(os.path.dirname(__file__), "golden",
                                               "convlstm_s64_k9.pt"))


@pytest.mark.parametrize("name", ["convlstm_s64_k9_hot", "convlstm_s16_k6_o3", "conv_s64_k9"])
def test_weight_packing_folds_batchnorm(name, forward_cases):
    """Folded convs from the blob reproduce conv+BN of the oracle (<=2e-6)."""
    import torch.nn.functional as F
    sd, md = load_golden_model(name)
    desc, blob, info = weights.pack_state_dict(sd)
    assert desc.struct_size == ctypes.sizeof(_native.ModelDesc)
    assert info["kmer_len"] == md["kmer_len"]
    assert desc.seq_conv[0].c_in == 4 * md["kmer_len"]
    sdf = {k: v.float() for k, v in sd.items() if v.dtype.is_floating_point}
    c = desc.sig_conv[0]
    w = torch.from_numpy(blob[c.w_off:c.w_off + c.c_out * c.c_in * c.kw]).view(c.c_out, c.c_in, c.kw)
    b = torch.from_numpy(blob[c.b_off:c.b_off + c.c_out])
    x = torch.randn(3, 1, 60)
    want = F.batch_norm(F.conv1d(x, sdf["sig_conv1.weight"], sdf["sig_conv1.bias"]),
                        sdf["sig_bn1.running_mean"], sdf["sig_bn1.running_var"],
                        sdf["sig_bn1.weight"], sdf["sig_bn1.bias"], training=False, eps=1e-5)
    assert (F.conv1d(x, w, b) - want).abs().max() < 2e-6
    if desc.n_lstm:
        H = desc.size
        off = desc.lstm_b_off[1]
        want_b = (sdf["lstm2.bias_ih_l0"] + sdf["lstm2.bias_hh_l0"]).numpy()
        assert np.allclose(blob[off:off + 4 * H], want_b, atol=1e-7)
    for off in (desc.fc_w_off, desc.fc_b_off):
        assert off % 4 == 0  # 16-byte aligned tensors


def test_chunk_extraction_matches_reference(read_cases):
    """RemoraRead.prepare_batches (vectorised) + oracle forward == reference call_read_mods."""
    meta, g = read_cases
    for i, m in enumerate(meta):
        sd, md = load_golden_model(m["model"])
        read = data_chunks.RemoraRead(dacs=g[f"r{i}_dacs"], shift=m["shift"], scale=m["scale"],
                                      seq_to_sig_map=g[f"r{i}_ssm"], int_seq=g[f"r{i}_int_seq"])
        assert read.str_seq == m["str_seq"]
        read.set_motif_focus_bases([util.Motif(*mot) for mot in md["motifs"]])
        read.prepare_batches(md, 16)
        assert all(len(b) <= 16 for b in read.batches)
        out = np.concatenate([ro.oracle_infer_compact(sd, md["kmer_context_bases"], b.signal,
                                                      b.sequence, b.seq_to_sig_map, b.seq_lens)
                              for b in read.batches])
        pos = np.concatenate([b.read_focus_bases for b in read.batches])
        order = np.argsort(g[f"r{i}_pos"])
        assert np.array_equal(pos, g[f"r{i}_pos"][order])
        assert np.abs(out - g[f"r{i}_nn_out"][order]).max() < 5e-6
        assert read.batches[0].sequence.dtype == np.int8
        assert read.batches[0].seq_to_sig_map.dtype == np.int16


def test_single_chunk_api_agrees_with_vectorised_path():
    from remora_b200.synth import synth_read
    dacs, shift, scale, ssm, int_seq = synth_read(60, seed=9)
    read = data_chunks.RemoraRead(dacs, shift, scale, ssm, int_seq=int_seq)
    read.focus_bases = np.array([0, 5, 30, 59])
    chunks = list(read.iter_chunks((50, 50), (4, 4)))
    assert len(chunks) == 4
    for ch in chunks:
        ch.check()
        assert ch.signal.shape == (100,) and ch.seq_to_sig_map[0] == 0
        assert ch.seq_to_sig_map[-1] == 100 and ch.seq_w_context.size == ch.seq_len + 8
    # first chunk sticks out on the left: zero padded signal, -1 padded sequence
    assert np.all(chunks[0].signal[: 50 - int(ssm[1]) // 2 - 1] == 0) or chunks[0].signal[0] == 0
    assert np.all(chunks[0].seq_w_context[:4] == -1)
    assert np.all(chunks[-1].seq_w_context[-4:] == -1)


def test_empty_read_returns_three_empty_arrays():
    from remora_b200 import inference
    _, md = load_golden_model("convlstm_s64_k9")
    read = data_chunks.RemoraRead(np.zeros(50), 0.0, 1.0, np.arange(0, 51, 10),
                                  int_seq=np.array([0, 0, 3, 3, 0]))  # no CG
    out = inference.call_read_mods(read, model=None, model_metadata=md)
    assert all(isinstance(a, np.ndarray) and a.size == 0 for a in out)
