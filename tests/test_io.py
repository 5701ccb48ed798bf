"""POD5 / BAM input path.

CPU: format round trips through our own writers (VBZ signal codec incl. extreme values, POD5 container
+ FlatBuffers footer, BGZF/BAM records and tags), reference-sequence reconstruction from MD, the
move-table / CIGAR coordinate maps; and - when the reference's test files are present (build container)
- the readers on the real files, checked against tests/golden/io_cases.npz (arrays produced by the
reference's own io.Read from the same records) and against the live reference classes.

GPU: the golden real reads through load_model -> call_read_mods (with and without signal-mapping
refinement) against the reference's CPU calls, and infer_from_pod5_and_bam end to end on a synthetic
POD5 + BAM pair written here.
"""
import os
import uuid
import zlib

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from remora_b200 import RemoraError, data_chunks, inference, io, model_util
from remora_b200.synth import synth_pod5_bam_run

REF_DATA = "/root/reference/tests/data"
have_ref_data = os.path.isfile(os.path.join(REF_DATA, "can_reads.pod5"))


@pytest.fixture(scope="module")
def io_cases():
    return np.load(os.path.join(GOLDEN, "io_cases.npz"))


# ---------------------------------------------------------------------------------------------
# codecs and containers: round trips
# ---------------------------------------------------------------------------------------------
def test_vbz_round_trip():
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 8, 9, 1000, 102400):
        sig = rng.integers(-32768, 32767, size=n).astype(np.int16)
        if n > 5:
            sig[:5] = [-32768, 32767, 0, -1, 1]  # deltas that wrap around int16
        blob = io.encode_vbz(sig)
        assert np.array_equal(io.decode_vbz(blob, n), sig), n
    smooth = (np.cumsum(rng.integers(-20, 20, size=50000)) + 900).astype(np.int16)
    assert len(io.encode_vbz(smooth)) < smooth.nbytes * 0.6  # one byte per sample + zstd
    with pytest.raises(RemoraError):
        io.decode_vbz(b"not a zstd frame", 4)


def make_ids(n, seed=0):
    rng = np.random.default_rng(seed)
    return [str(uuid.UUID(int=int(rng.integers(1, 2 ** 62)))) for _ in range(n)]


# Known-answer vector built BY HAND from the published "minknow.vbz" / svb16 format description (POD5
# format spec: delta -> zig-zag -> StreamVByte-16 with one key bit per value, keys first, little-endian
# data, then a zstd frame), independent of this package's encoder:
#   samples  10, 12, 9, 300, -200, -200, 32767, -32768, 0
#   deltas   10,  2, -3, 291, -500,   0, 32967->(wraps mod 2^16: -32569), -65535->(+1), 32768->(-32768)
#   zig-zag  20,  4,  5, 582,  999,   0, 65137, 2, 65535
#   key bits (1 = two bytes): values 3, 4, 6, 8 -> byte0 = 0b01011000 = 0x58, byte1 = 0b00000001
SVB16_KAT_SAMPLES = np.array([10, 12, 9, 300, -200, -200, 32767, -32768, 0], dtype=np.int16)
SVB16_KAT_STREAM = bytes([0x58, 0x01,
                          0x14, 0x04, 0x05, 0x46, 0x02, 0xE7, 0x03, 0x00, 0x71, 0xFE, 0x02, 0xFF, 0xFF])


def _kat_blob():
    import pyarrow as pa
    return pa.Codec("zstd").compress(SVB16_KAT_STREAM, asbytes=True)


def test_svb16_known_answer_numpy():
    assert np.array_equal(io.decode_vbz(_kat_blob(), SVB16_KAT_SAMPLES.size), SVB16_KAT_SAMPLES)
    # and the package's own encoder produces exactly this stream
    assert bytes(io._zstd_frame_content(io.encode_vbz(SVB16_KAT_SAMPLES))) == SVB16_KAT_STREAM


@pytest.mark.gpu
def test_svb16_known_answer_gpu():
    dev = torch.device("cuda:0")
    d_out, spans = io.decode_vbz_rows_gpu([_kat_blob()], [SVB16_KAT_SAMPLES.size], dev)
    st, ln = spans[0]
    assert ln == SVB16_KAT_SAMPLES.size
    assert np.array_equal(d_out[st:st + ln].cpu().numpy(), SVB16_KAT_SAMPLES)


def test_pod5_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    sig = (np.cumsum(rng.integers(-30, 30, size=250000)) + 900).astype(np.int16)
    ids = make_ids(3)
    reads = [(ids[0], sig, -240.0, 0.1755), (ids[1], sig[:1234], 3.0, 0.25), (ids[2], sig[:0], 0.0, 1.0)]
    path = str(tmp_path / "t.pod5")
    io.write_pod5(path, reads)
    with io.Pod5Reader(path) as reader:
        assert reader.read_ids == ids and reader.num_reads == 3
        for rid, s, off, sc in reads:
            got = reader.get_read(rid)
            assert str(got.read_id) == rid and np.array_equal(got.signal, s) and got.num_samples == s.size
            assert got.calibration.offset == np.float32(off) and got.calibration.scale == np.float32(sc)
        assert [str(r.read_id) for r in reader.reads(selection=[ids[1], "missing"])] == [ids[1]]
        with pytest.raises(RemoraError):
            reader.get_read("missing")
    assert [str(r.read_id) for r in io.iter_pod5_reads(path, num_reads=2)] == ids[:2]
    # several Arrow record batches and small signal chunks (what files written by MinKNOW look like)
    many = str(tmp_path / "batches.pod5")
    io.write_pod5(many, reads, chunk=4096, rows_per_batch=3)
    with io.Pod5Reader(many) as reader:
        assert reader._tables["signal"].num_record_batches > 5 and reader.read_ids == ids
        for rid, s, off, sc in reads:
            assert np.array_equal(reader.get_read(rid).signal, s)
    bad = tmp_path / "bad.pod5"
    bad.write_bytes(b"x" * 100)
    with pytest.raises(RemoraError):
        io.Pod5Reader(str(bad))


def test_bam_round_trip(tmp_path):
    rng = np.random.default_rng(2)
    ids = make_ids(3, seed=5)
    mv = np.r_[5, rng.integers(0, 2, size=200)].astype(np.int8)
    recs = [
        dict(query_name=ids[0], flag=0, reference_id=0, reference_start=99, mapping_quality=60,
             cigartuples=[(4, 3), (0, 10), (1, 2), (2, 1), (0, 5)], query_sequence="ACGTNACGTAACCGGTTACG",
             tags=[("mv", "Bc", mv), ("ts", "i", 10), ("ns", "I", 1010), ("sm", "f", 87.5), ("sd", "f", 26.25),
                   ("MD", "Z", "10^A5"), ("tp", "A", "P"), ("pi", "Z", ids[1]), ("xs", "s", -7), ("xc", "C", 200),
                   ("xf", "Bf", np.float32([1.5, -2.25]))]),
        dict(query_name=ids[1], query_sequence="ACG"),
        dict(query_name=ids[2], flag=0x900, reference_id=0, reference_start=5, cigartuples=[(0, 4)],
             query_sequence="ACGT"),
    ]
    recs += [dict(query_name=f"bulk{i}", query_sequence="ACGT" * 400) for i in range(120)]  # > one BGZF block
    path = str(tmp_path / "t.bam")
    io.write_bam(path, "@HD\tVN:1.6\n@SQ\tSN:chr1\tLN:1000\n", [("chr1", 1000)], recs)
    with io.BamReader(path) as bam:
        assert bam.references == ["chr1"] and bam.lengths == [1000] and bam.header_text.startswith("@HD")
        out = list(bam)
    assert len(out) == len(recs)
    a = out[0]
    assert a.query_name == ids[0] and a.query_sequence == recs[0]["query_sequence"]
    assert a.cigartuples == recs[0]["cigartuples"] and a.reference_name == "chr1" and a.reference_start == 99
    assert a.mapping_quality == 60 and not a.is_reverse and not a.is_unmapped
    assert a.get_tag("ts") == 10 and a.get_tag("ns") == 1010 and a.get_tag("xs") == -7 and a.get_tag("xc") == 200
    assert list(a.get_tag("mv")) == list(mv) and isinstance(a.get_tag("mv")[0], int)
    assert a.get_tag("sd") == 26.25 and a.get_tag("tp") == "P" and a.get_tag("pi") == ids[1]
    assert list(a.get_tag("xf")) == [1.5, -2.25] and a.has_tag("MD") and not a.has_tag("zz")
    with pytest.raises(KeyError):
        a.get_tag("zz")
    # MD + CIGAR -> reference bases: 10 matches, deleted A, 5 matches (the 2-base insertion drops out)
    assert a.get_reference_sequence() == "TNACGTAACC" + "A" + "TTACG"
    assert out[1].is_unmapped and out[1].reference_name is None and out[1].query_sequence == "ACG"
    assert out[2].is_secondary and out[2].is_supplementary and not io.read_is_primary(out[2])
    sam = a.to_sam(drop_tags=("mv",), extra_tags=["MM:Z:C+m?,1;"]).split("\t")
    assert sam[:6] == [ids[0], "0", "chr1", "100", "60", "3S10M2I1D5M"] and sam[-1] == "MM:Z:C+m?,1;"
    assert "ts:i:10" in sam and "sd:f:26.25" in sam and not any(f.startswith("mv:") for f in sam)
    # the index keys split reads by their parent id and skips non-primary records
    idx = io.ReadIndexedBam(path)
    assert ids[1] in idx and idx.get_first_alignment(ids[1]).query_name == ids[0]  # pi tag -> parent id
    assert ids[2] not in idx and idx.skip_reasons["Non-primary alignment"] == 1
    assert io.ReadIndexedBam(path, req_tags={"mv"}).num_reads == 1
    with pytest.raises(RemoraError):
        next(idx.get_alignments("missing"))
    # pointer index (virtual offsets, records re-read on demand) == in-memory index, also across blocks
    lazy = io.ReadIndexedBam(path, in_memory=False)
    assert lazy.read_ids == idx.read_ids and lazy.num_records == idx.num_records
    for rid in (ids[1], "bulk0", "bulk77", "bulk119"):
        a_mem, a_lazy = idx.get_first_alignment(rid), lazy.get_first_alignment(rid)
        assert a_mem.query_name == a_lazy.query_name and a_mem.query_sequence == a_lazy.query_sequence
        assert a_mem.tags == a_lazy.tags or [t for t, _ in a_mem.tags] == [t for t, _ in a_lazy.tags]
    assert sum(1 for _ in lazy) == idx.num_records
    lazy.close()
    not_bam = tmp_path / "plain.bam"
    not_bam.write_bytes(b"plain text, not BGZF" * 4)
    with pytest.raises(RemoraError):
        io.BamReader(str(not_bam))


def test_md_mismatches_and_errors():
    rec = io.AlignedSegment("r", 0, 0, "c", 0, 0, [(0, 6)], "ACGTAC", np.zeros(6, np.uint8), [("MD", "2A3")])
    assert rec.get_reference_sequence() == "ACaTAC"
    rec.tags = [("MD", "2A9")]
    with pytest.raises(ValueError):
        rec.get_reference_sequence()
    rec.tags = []
    with pytest.raises(ValueError):
        rec.get_reference_sequence()


def test_move_table_and_cigar_maps():
    mv = [5, 1, 0, 1, 1, 0, 0, 1]
    q2s, table, stride = io.parse_move_tag(mv, sig_len=35, seq_len=4)
    assert stride == 5 and np.array_equal(q2s, [0, 10, 15, 30, 35]) and table.size == 7
    rq2s = io.parse_move_tag(mv, sig_len=35, seq_len=4, reverse_signal=True)[0]
    assert np.array_equal(rq2s, [0, 5, 20, 25, 35])
    with pytest.raises(RemoraError):
        io.parse_move_tag(mv, sig_len=35, seq_len=5)
    with pytest.raises(RemoraError):
        io.parse_move_tag(mv, sig_len=80, seq_len=4)
    # 2 matches, 1 inserted query base, 1 match, 1 deleted reference base, 1 match
    knots = io.make_sequence_coordinate_mapping([(0, 2), (1, 1), (0, 1), (2, 1), (0, 1)])
    assert np.allclose(knots, [0, 1, 3, 3.5, 4, 5])
    r2s = io.compute_ref_to_signal(np.array([0, 10, 20, 30, 40, 50]), [(0, 2), (1, 1), (0, 1), (2, 1), (0, 1)])
    assert np.array_equal(r2s, [0, 10, 30, 35, 40, 50])
    with pytest.raises(RemoraError):
        io.make_sequence_coordinate_mapping([(4, 5)])


# ---------------------------------------------------------------------------------------------
# the reference's real test files (build container only)
# ---------------------------------------------------------------------------------------------
@pytest.mark.skipif(not have_ref_data, reason="reference test data not present")
def test_real_files_match_golden(io_cases):
    for stem, rid, n_seq, n_sig, strand in io_cases["index"]:
        bam_idx = io.ReadIndexedBam(os.path.join(REF_DATA, f"{stem}_mappings.bam"))
        assert bam_idx.num_reads == 14 and rid in bam_idx
        with io.Pod5Reader(os.path.join(REF_DATA, f"{stem}_reads.pod5")) as pod5:
            read = io.Read.from_pod5_and_alignment(pod5.get_read(rid), bam_idx.get_first_alignment(rid))
        assert len(read.seq) == int(n_seq) and read.dacs.size == int(n_sig) and read.ref_reg.strand == strand
        k = next(key[: -len("bc_dacs")] for key in io_cases.files if key.startswith(stem) and key.endswith("bc_dacs")
                 and io_cases[key].size == read.query_to_signal[-1] - read.query_to_signal[0]
                 and np.array_equal(io_cases[key][:50], read.dacs[read.query_to_signal[0]:][:50]))
        for anchor, tag in ((False, "bc"), (True, "ref")):
            rr = read.into_remora_read(anchor)
            if tag == "bc":
                assert np.array_equal(rr.dacs, io_cases[k + "bc_dacs"])
            else:
                assert [rr.dacs.size, zlib.crc32(rr.dacs.astype(np.int16).tobytes())] == list(io_cases[k + "ref_dacs"])
            assert np.array_equal(rr.seq_to_sig_map, io_cases[k + tag + "_ssm"])
            assert np.array_equal(rr.int_seq, io_cases[k + tag + "_int_seq"])
            assert [rr.shift, rr.scale] == list(io_cases[k + tag + "_shift_scale"])


@pytest.mark.skipif(not have_ref_data, reason="reference test data not present")
def test_read_join_equals_live_reference():
    import ref_harness
    ref_harness.import_reference()
    from remora import io as ref_io
    for stem in ("can", "mod"):
        bam_idx = io.ReadIndexedBam(os.path.join(REF_DATA, f"{stem}_mappings.bam"))
        with io.Pod5Reader(os.path.join(REF_DATA, f"{stem}_reads.pod5")) as pod5:
            for rid in pod5.read_ids[:6]:
                rec = bam_idx.get_first_alignment(rid)
                mine = io.Read.from_pod5_and_alignment(pod5.get_read(rid), rec)
                theirs = ref_io.Read.from_pod5_and_alignment(pod5.get_read(rid), rec)
                assert mine.ref_seq == theirs.ref_seq and mine.ref_reg.end == theirs.ref_reg.end
                for anchor in (False, True):
                    a, b = mine.into_remora_read(anchor), theirs.into_remora_read(anchor)
                    assert np.array_equal(a.dacs, b.dacs) and (a.shift, a.scale) == (b.shift, b.scale)
                    assert np.array_equal(a.seq_to_sig_map, b.seq_to_sig_map) and a.str_seq == b.str_seq


# ---------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------
def golden_reads(io_cases):
    for key in sorted(k for k in io_cases.files if k.endswith("bc_dacs")):
        k = key[: -len("bc_dacs")]
        shift, scale = io_cases[k + "bc_shift_scale"]
        yield k, dict(dacs=io_cases[k + "bc_dacs"], shift=float(shift), scale=float(scale),
                      seq_to_sig_map=io_cases[k + "bc_ssm"], int_seq=io_cases[k + "bc_int_seq"].astype(np.int64))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["convlstm_s64_k9_hot", "convlstm_s64_k9_refine"])
def test_gpu_real_reads_match_reference_cpu_calls(io_cases, name):
    """BASELINE config 5 in miniature: real reads (reference tests/data), mod-call parity against the
    reference's CPU path, without and with signal-mapping refinement."""
    dev = torch.device("cuda:0")
    model, md = model_util.load_model(os.path.join(GOLDEN, name + ".pt"), device=dev, eval_only=True)
    n = 0
    for k, kw in golden_reads(io_cases):
        for on_device in (False, True):
            read = data_chunks.RemoraRead(**{a: (v.copy() if hasattr(v, "copy") else v) for a, v in kw.items()})
            nn_out, _, pos = inference.call_read_mods(read, model, md, extract_on_device=on_device)
            assert np.array_equal(read.seq_to_sig_map, io_cases[k + name + "_ssm"])
            assert [read.shift, read.scale] == list(io_cases[k + name + "_shift_scale"])
            order, want = np.argsort(pos), np.argsort(io_cases[k + name + "_pos"])
            assert np.array_equal(pos[order], io_cases[k + name + "_pos"][want])
            assert np.abs(nn_out[order] - io_cases[k + name + "_nn_out"][want]).max() < 1e-4
        read = data_chunks.RemoraRead(**kw)
        mm, ml = inference.call_read_mods(read, model, md, return_mm_ml_tags=True)
        assert mm == str(io_cases[k + name + "_mm"])
        got_ml, want_ml = np.frombuffer(ml, dtype=np.uint8).astype(int), io_cases[k + name + "_ml"].astype(int)
        assert got_ml.size == want_ml.size and np.abs(got_ml - want_ml).max() <= 1  # bin edges, SURVEY App. D
        n += 1
    assert n == 4


@pytest.mark.gpu
def test_gpu_reference_anchored_calls_match_reference(io_cases):
    """`remora infer --reference-anchored`: reads rebuilt from the golden arrays of the reference's
    reference-anchored RemoraRead (dacs are a slice of the basecall-anchored samples)."""
    dev = torch.device("cuda:0")
    model, md = model_util.load_model(os.path.join(GOLDEN, "convlstm_s64_k9_hot.pt"), device=dev, eval_only=True)
    n = 0
    for key in sorted(k for k in io_cases.files if k.endswith("bc_dacs")):
        k = key[: -len("bc_dacs")]
        full = io_cases[key]
        size, crc = (int(v) for v in io_cases[k + "ref_dacs"])
        ssm = io_cases[k + "ref_ssm"]
        # locate the slice: the reference-anchored samples start where the first aligned base starts
        starts = [st for st in range(0, full.size - size + 1)
                  if zlib.crc32(full[st:st + size].tobytes()) == crc] if size <= full.size else []
        if not starts:
            continue
        shift, scale = io_cases[k + "ref_shift_scale"]
        read = data_chunks.RemoraRead(dacs=full[starts[0]:starts[0] + size].copy(), shift=float(shift),
                                      scale=float(scale), seq_to_sig_map=ssm.copy(),
                                      int_seq=io_cases[k + "ref_int_seq"].astype(np.int64))
        nn_out, _, pos = inference.call_read_mods(read, model, md)
        order, want = np.argsort(pos), np.argsort(io_cases[k + "ref_pos"])
        assert np.array_equal(pos[order], io_cases[k + "ref_pos"][want])
        assert np.abs(nn_out[order] - io_cases[k + "ref_nn_out"][want]).max() < 1e-4
        n += 1
    assert n >= 1


def write_synthetic_run(tmp_path, n_reads=12, seed=0):
    return synth_pod5_bam_run(str(tmp_path / "run.pod5"), str(tmp_path / "run.bam"), n_reads=n_reads, seed=seed)


def test_synthetic_run_reads_back(tmp_path):
    pod5, bam, truth = write_synthetic_run(tmp_path, n_reads=4)
    idx = io.ReadIndexedBam(bam, req_tags={"mv"})
    got = list(io.iter_io_reads(pod5, idx))
    assert len(got) == 4 and all(err is None for _, err in got)
    for read, _ in got:
        t = truth[read.read_id]
        rr = read.into_remora_read(False)
        assert np.array_equal(rr.dacs, t["dacs"]) and rr.str_seq == t["seq"]
        assert np.array_equal(rr.seq_to_sig_map, t["ssm"])
        assert np.isclose(rr.shift, t["shift"]) and np.isclose(rr.scale, t["scale"])
    with pytest.raises(RemoraError):  # unmapped reads cannot be reference anchored
        got[0][0].into_remora_read(True)


class _OracleModel(torch.nn.Module):
    """CPU stand-in with the B200Model call surface (tests only): logits from the oracle forward."""

    def __init__(self, name):
        super().__init__()
        import remora_oracle as ro
        from conftest import load_golden_model
        self.ro = ro
        self.sd, self.md = load_golden_model(name)
        self.anchor = torch.nn.Parameter(torch.zeros(1), requires_grad=False)

    def forward_compact(self, sig, seq, mp, ln, out=None):
        res = self.ro.oracle_infer_compact(self.sd, self.md["kmer_context_bases"], sig.numpy(), seq.numpy(),
                                           mp.numpy(), ln.numpy())
        return torch.from_numpy(np.ascontiguousarray(res))

    def softmax_ml(self, logits, want_probs=True):
        from remora_b200 import util
        probs = util.softmax_axis1(logits.numpy())[:, 1:]
        ml = torch.from_numpy(self.ro.ml_bytes(probs))
        return (torch.from_numpy(probs.astype(np.float32)) if want_probs else None), ml


def test_pipeline_host_logic_on_cpu(tmp_path):
    """infer_from_pod5_and_bam with every GPU stage switched off or stubbed (oracle forward, numpy signal
    decode, host chunk extraction, a model without a refiner): the host plumbing - grouping reads, merging
    chunk batches of different widths, tags, SAM / BAM output, rank sharding - against per-read calls."""
    pod5, bam, truth = write_synthetic_run(tmp_path, n_reads=7)
    model = _OracleModel("convlstm_s16_k6_o3")
    md = dict(model.md)
    assert not md["sig_map_refiner"].is_loaded
    kw = dict(decode_on_device=False, extract_on_device=False)
    out_bam = str(tmp_path / "cpu.bam")
    res = inference.infer_from_pod5_and_bam(pod5, bam, (model, md), out_path=out_bam, reads_per_batch=3,
                                            batch_size=50, return_probs=True, **kw)
    assert [r["read_id"] for r in res] == list(truth) and all(r["error"] is None for r in res)
    with io.BamReader(out_bam) as reader:
        recs = {r.query_name: r for r in reader}
    n_called = 0
    for r in res:
        t = truth[r["read_id"]]
        read = data_chunks.RemoraRead(dacs=t["dacs"].copy(), shift=float(t["shift"]), scale=float(t["scale"]),
                                      seq_to_sig_map=t["ssm"].copy(), str_seq=t["seq"])
        probs, _, pos = inference.call_read_mods(read, model, md, return_mod_probs=True)
        if len(pos) == 0:
            assert r["mm"] == "" and not r["calls"]
            continue
        n_called += 1
        got_pos, got_probs = r["calls"][md["can_base"]]
        assert np.array_equal(got_pos, pos) and np.allclose(got_probs, probs, atol=1e-5)
        assert recs[r["read_id"]].get_tag("MM") == r["mm"]
        assert list(recs[r["read_id"]].get_tag("ML")) == list(r["ml"])
        assert len(r["ml"]) == len(pos) * len(md["mod_bases"])  # two modified bases -> two MM sections
        assert r["mm"].count(";") == len(md["mod_bases"])
    assert n_called >= 5
    # two ranks: disjoint shards whose union is the single-process result
    shards = [inference.infer_from_pod5_and_bam(pod5, bam, (model, md), rank=k, world_size=2, **kw)
              for k in range(2)]
    ids = [r["read_id"] for sh in shards for r in sh]
    assert sorted(ids) == sorted(truth) and len(set(ids)) == len(ids) and min(len(sh) for sh in shards) >= 2
    by_id = {r["read_id"]: r for sh in shards for r in sh}
    assert all(by_id[r["read_id"]]["mm"] == r["mm"] and by_id[r["read_id"]]["ml"] == r["ml"] for r in res)
    # SAM output, num_reads
    out_sam = str(tmp_path / "cpu.sam")
    few = inference.infer_from_pod5_and_bam(pod5, bam, (model, md), out_path=out_sam, num_reads=2, **kw)
    assert len(few) == 2
    lines = [ln for ln in open(out_sam).read().splitlines() if not ln.startswith("@")]
    assert len(lines) == 2 and all(ln.split("\t")[0] == r["read_id"] for ln, r in zip(lines, few))


@pytest.mark.gpu
def test_gpu_svb16_decode_bit_exact(io_cases, tmp_path):
    """rb200_svb16_decode against the numpy decoder: ragged row lengths around the 32-sample word and
    8192-sample tile boundaries, int16 extremes (deltas that wrap), empty rows, multi-row reads, real
    signal statistics (the reference's test reads), and a corrupted row."""
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    sizes = [0, 1, 5, 31, 32, 33, 255, 256, 257, 8191, 8192, 8193, 16384, 20001, 102400, 70000]
    rows, owner = [], []
    for k, n in enumerate(sizes):
        sig = (np.cumsum(rng.integers(-400, 400, size=n)) + 700).astype(np.int16)
        if n > 4:
            sig[1:4] = [-32768, 32767, -1]
        rows.append(sig)
        owner.append(k // 2)  # two rows per "read"
    real = io_cases[sorted(k for k in io_cases.files if k.endswith("bc_dacs"))[0]]
    rows.append(real)
    owner.append(99)
    blobs = [io.encode_vbz(r) for r in rows]
    d_out, spans = io.decode_vbz_rows_gpu(blobs, [r.size for r in rows], dev, owner)
    host = d_out.cpu().numpy()
    assert len(spans) == len(set(owner)) and all(st % 8 == 0 for st, _ in spans)
    for read_idx, (st, ln) in zip(sorted(set(owner)), spans):
        want = np.concatenate([r for r, o in zip(rows, owner) if o == read_idx])
        assert ln == want.size and np.array_equal(host[st:st + ln], want), read_idx
        for r, blob in zip(rows, blobs):
            assert np.array_equal(io.decode_vbz(blob, r.size), r)
    # through the POD5 reader: one launch for several reads
    ids = make_ids(3, seed=9)
    path = str(tmp_path / "g.pod5")
    io.write_pod5(path, [(ids[0], rows[14], -240.0, 0.18), (ids[1], rows[9], 1.0, 0.2), (ids[2], rows[0], 0.0, 1.0)])
    with io.Pod5Reader(path) as reader:
        got = reader.get_reads(ids, device=dev)
        assert [str(g.read_id) for g in got] == ids
        assert np.array_equal(got[0].signal, rows[14]) and np.array_equal(got[1].signal, rows[9])
        assert got[2].signal.size == 0
    # a truncated row is reported, not read past
    raw = io._zstd_frame_content(blobs[10])
    import pyarrow as pa
    bad = pa.Codec("zstd").compress(bytes(raw)[:-5], asbytes=True)
    with pytest.raises(RemoraError):
        io.decode_vbz_rows_gpu([bad], [rows[10].size], dev)


@pytest.mark.gpu
def test_gpu_infer_from_pod5_and_bam(tmp_path):
    dev = torch.device("cuda:0")
    pod5, bam, truth = write_synthetic_run(tmp_path, n_reads=12)
    model, md = model_util.load_model(os.path.join(GOLDEN, "convlstm_s64_k9_refine.pt"), device=dev,
                                      eval_only=True)
    out_sam = str(tmp_path / "calls.sam")
    res = inference.infer_from_pod5_and_bam(pod5, bam, (model, md), out_path=out_sam, reads_per_batch=5,
                                            return_probs=True, extract_on_device=True)
    assert len(res) == 12 and all(r["error"] is None for r in res)
    lines = [ln for ln in open(out_sam).read().splitlines() if not ln.startswith("@")]
    assert len(lines) == 12
    for r, line in zip(res, lines):
        t = truth[r["read_id"]]
        # the same read through the single-read API (refinement batch of one) gives the same calls
        read = data_chunks.RemoraRead(dacs=t["dacs"].copy(), shift=float(t["shift"]), scale=float(t["scale"]),
                                      seq_to_sig_map=t["ssm"].copy(), str_seq=t["seq"])
        probs, _, pos = inference.call_read_mods(read, model, md, return_mod_probs=True)
        if len(pos) == 0:
            assert r["mm"] == ""
            continue
        got_pos, got_probs = r["calls"]["C"]
        assert np.array_equal(got_pos, pos) and np.allclose(got_probs, probs, atol=1e-6)
        fields = line.split("\t")
        assert fields[0] == r["read_id"] and f"MM:Z:{r['mm']}" in fields
        assert any(f.startswith("ML:B:C,") for f in fields)
        assert any(f.startswith("mv:B:c,") for f in fields)  # the move table is kept, as the reference keeps it
        assert r["mm"].startswith("C+m?,") and len(r["ml"]) == len(pos)
    # BAM output carries the same tags; two models (one per canonical base) concatenate their sections
    out_bam = str(tmp_path / "calls.bam")
    model2, md2 = model_util.load_model(os.path.join(GOLDEN, "convlstm_s64_k9_hot.pt"), device=dev, eval_only=True)
    md2 = dict(md2, can_base="G", motifs=[("GC", 0)], motif=("GC", 0), mod_bases="x", mod_long_names=["test"])
    res2 = inference.infer_from_pod5_and_bam(pod5, bam, {"C": (model, md), "G": (model2, md2)}, out_path=out_bam,
                                             decode_on_device=False, extract_on_device=False, drop_move_tag=True)
    with io.BamReader(out_bam) as reader:
        recs = {r.query_name: r for r in reader}
    assert len(recs) == 12
    for r1, r2 in zip(res, res2):
        assert r2["read_id"] == r1["read_id"] and r2["mm"].startswith(r1["mm"])  # host paths: same C calls
        rec = recs[r2["read_id"]]
        if r2["mm"]:
            assert rec.get_tag("MM") == r2["mm"] and list(rec.get_tag("ML")) == list(r2["ml"])
            assert "G+x?," in r2["mm"] or "GC" not in truth[r2["read_id"]]["seq"]
        assert not rec.has_tag("mv") and rec.query_sequence == truth[r2["read_id"]]["seq"]
