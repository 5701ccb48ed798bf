"""Generate the committed golden vectors of the signal-mapping refinement by
RUNNING THE REFERENCE (nanoporetech/remora at /root/reference) in the build container.

    python tests/golden/make_golden_refine.py

Outputs (tests/golden/):
  convlstm_s64_k9_refine.pt   the "hot" ConvLSTM_w_ref fixture exported by the reference with a loaded
                              SigMapRefiner in its metadata (seeded 6-mer level table, central position 2,
                              rough re-scaling, one round of dwell_penalty refinement, half bandwidth 5)
  refine_cases.npz            synthetic reads (remora_b200.synth.synth_refine_read) and, from the
                              reference: levels, rough re-scaled shift/scale, seq band, path, final score,
                              checksums of the full score / traceback arrays (refine_signal_mapping), and
                              the result of RemoraRead.refine_signal_mapping
  refine_read_cases.npz       reads through reference inference.call_read_mods with the model above
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)

import make_golden  # noqa: E402  (imports the reference through oracle/ref_harness.py)
from remora import inference, model_util  # noqa: E402
from remora import refine_signal_map as ref_rsm  # noqa: E402
from remora.data_chunks import RemoraRead  # noqa: E402
from remora_b200.synth import synth_levels_table, synth_refine_read  # noqa: E402

KMER_LEN, CENTER = 6, 2
TABLE = synth_levels_table(KMER_LEN, seed=0)
SD_ARR = ref_rsm.DEFAULT_REFINE_SHORT_DWELL_PEN

# (n_bases, seed, algo, extra synth arguments, leading/trailing unmapped samples)
CASES = [
    (12, 1, "dwell_penalty", {}, (0, 0)),
    (40, 2, "dwell_penalty", {}, (0, 0)),
    (300, 3, "dwell_penalty", {}, (37, 11)),
    (1200, 4, "dwell_penalty", {}, (0, 0)),
    (900, 5, "dwell_penalty", dict(frac_zero_dwell=0.15, jitter=12), (0, 0)),
    (500, 6, "dwell_penalty", dict(frac_stall=0.02, stall_range=(500, 1500)), (5, 0)),  # bands > 1024 samples
    (700, 7, "dwell_penalty", dict(mean_dwell=4, frac_stall=0.0), (0, 0)),
    (40, 8, "Viterbi", {}, (0, 0)),
    (800, 9, "Viterbi", {}, (3, 9)),
    (600, 10, "Viterbi", dict(frac_stall=0.03), (0, 0)),
]


def checksum_i(a):
    a = a.astype(np.int64)
    return np.array([a.sum(), (a * (np.arange(a.size) % 1009)).sum()], dtype=np.int64)


def refine_cases():
    out = {}
    index = []
    for cid, (n, seed, algo, kw, (lead, trail)) in enumerate(CASES):
        dacs, shift, scale, ssm, int_seq = synth_refine_read(n, TABLE, KMER_LEN, CENTER, seed=seed, **kw)
        if lead or trail:  # signal outside the mapped range (a trimmed / clipped read)
            rng = np.random.default_rng(1000 + seed)
            dacs = np.concatenate([rng.integers(300, 600, lead).astype(np.int16), dacs,
                                   rng.integers(300, 600, trail).astype(np.int16)])
            ssm = ssm + lead
        refiner = ref_rsm.SigMapRefiner(_levels_array=TABLE, center_idx=CENTER, do_rough_rescale=True,
                                        scale_iters=0, algo=algo, half_bandwidth=5, sd_arr=SD_ARR)
        levels = refiner.extract_levels(int_seq)
        sh, sc = refiner.rough_rescale(shift, scale, ssm, int_seq, dacs)
        path, all_scores, traceback, seq_band, base_offsets = ref_rsm.refine_signal_mapping(
            (dacs - sh) / sc, ssm, levels, 5, algo, SD_ARR)
        read = RemoraRead(dacs.copy(), shift, scale, ssm.copy(), int_seq.copy())
        read.refine_signal_mapping(refiner)
        assert np.array_equal(read.seq_to_sig_map, path) and read.shift == sh and read.scale == sc
        k = f"c{cid}_"
        out.update({k + "dacs": dacs, k + "ssm": ssm, k + "int_seq": int_seq.astype(np.int8),
                    k + "shift_scale": np.array([shift, scale, sh, sc], dtype=np.float64),
                    k + "levels": levels, k + "seq_band": seq_band, k + "path": path.astype(np.int32),
                    k + "final_score": np.float32(all_scores[-1]),
                    k + "scores_sum": np.float64(all_scores.astype(np.float64).sum()),
                    k + "tb_check": checksum_i(traceback)})
        index.append([cid, n, {"Viterbi": 0, "dwell_penalty": 1}[algo],
                      int((seq_band[1] - seq_band[0]).max())])
        print(cid, n, algo, "changed", float((path != ssm).mean()), "max band", index[-1][3])
    # NaN levels (bases without a level keep their samples, refine_signal_map.py:673-680, 829-830)
    dacs, shift, scale, ssm, int_seq = synth_refine_read(400, TABLE, KMER_LEN, CENTER, seed=20)
    refiner = ref_rsm.SigMapRefiner(_levels_array=TABLE, center_idx=CENTER, do_rough_rescale=False,
                                    scale_iters=0)
    levels = refiner.extract_levels(int_seq)
    levels[[50, 51, 52, 200]] = np.nan
    path, all_scores, traceback, seq_band, _ = ref_rsm.refine_signal_mapping(
        (dacs - shift) / scale, ssm, levels, 5, "dwell_penalty", SD_ARR)
    out.update({"nan_dacs": dacs, "nan_ssm": ssm, "nan_levels": levels,
                "nan_shift_scale": np.array([shift, scale]), "nan_seq_band": seq_band,
                "nan_path": path.astype(np.int32), "nan_final_score": np.float32(all_scores[-1]),
                "nan_tb_check": checksum_i(traceback)})
    out["index"] = np.array(index)
    out["table"] = TABLE
    np.savez_compressed(os.path.join(HERE, "refine_cases.npz"), **out)


def refine_model_and_reads():
    path = os.path.join(HERE, "convlstm_s64_k9_refine.pt")
    make_golden.make_model(
        arch="ConvLSTM_w_ref", size=64, kmer_context=(4, 4), num_out=2, chunk_context=(50, 50),
        motifs=[("CG", 0)], mod_bases="m", mod_long_names=["5mC"], seed=3, hot=(3.0, 8.0, 16.0), path=path,
        refine=dict(refine_kmer_levels=TABLE, refine_kmer_center_idx=CENTER, refine_do_rough_rescale=True,
                    refine_scale_iters=0, refine_algo="dwell_penalty", refine_half_bandwidth=5,
                    refine_sd_arr=SD_ARR))
    model, md = model_util.load_model(path, eval_only=True)
    assert md["sig_map_refiner"].is_loaded
    arrays = {}
    for idx, (n, seed) in enumerate([(500, 60), (90, 61), (1500, 62)]):
        dacs, shift, scale, ssm, int_seq = synth_refine_read(n, TABLE, KMER_LEN, CENTER, seed=seed)
        read = RemoraRead(dacs=dacs.copy(), shift=shift, scale=scale, seq_to_sig_map=ssm.copy(),
                          int_seq=int_seq.copy(), read_id=f"refine{idx}")
        nn_out, labels, pos = inference.call_read_mods(read, model, md)
        k = f"r{idx}_"
        arrays.update({k + "dacs": dacs, k + "ssm": ssm, k + "int_seq": int_seq.astype(np.int8),
                       k + "shift_scale": np.array([shift, scale, read.shift, read.scale]),
                       k + "refined_ssm": np.asarray(read.seq_to_sig_map),
                       k + "nn_out": nn_out.astype(np.float32), k + "pos": pos})
        print("read", idx, n, "calls", pos.size)
    np.savez_compressed(os.path.join(HERE, "refine_read_cases.npz"), **arrays)


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    refine_cases()
    refine_model_and_reads()
    real_read_cases()  # needs tests/golden/io_cases.npz (make_golden_io.py)


def real_read_cases():
    """The reference's own 9-mer level table (tests/data/levels.txt) and two of its real test reads
    (tests/golden/io_cases.npz: samples, move-table mapping and basecalls as the reference's io.Read
    gives them): k-mer table statistics, per-base levels, rough re-scaling and the refined mapping from
    the reference's SigMapRefiner.  The 1 MB table itself is not committed; the per-base levels are."""
    levels_path = os.path.join(make_golden.ref_harness.REFERENCE_ROOT, "tests", "data", "levels.txt")
    refiner = ref_rsm.SigMapRefiner(kmer_model_filename=levels_path, do_rough_rescale=True, scale_iters=0,
                                    algo="dwell_penalty", half_bandwidth=5, sd_arr=SD_ARR)
    io_cases = np.load(os.path.join(HERE, "io_cases.npz"))
    out = {"kmer_len": np.array(refiner.kmer_len), "center_idx": np.array(refiner.center_idx),
           "kmer_idx_stats": np.array(refiner.kmer_idx_stats, dtype=np.float64),
           "table_head": refiner.levels_array[:4096].copy(),
           "table_crc": np.array(__import__("zlib").crc32(refiner.levels_array.tobytes()))}
    keys = sorted(k for k in io_cases.files if k.endswith("bc_dacs"))
    for idx, key in enumerate((keys[0], keys[2])):
        k = key[: -len("bc_dacs")]
        dacs, ssm = io_cases[key], io_cases[k + "bc_ssm"]
        int_seq = io_cases[k + "bc_int_seq"].astype(np.int64)
        shift, scale = io_cases[k + "bc_shift_scale"]
        for algo in ("dwell_penalty", "Viterbi"):
            refiner.algo = algo
            read = RemoraRead(dacs.copy(), float(shift), float(scale), ssm.copy(), int_seq.copy())
            read.refine_signal_mapping(refiner)
            out[f"real{idx}_{algo}_ssm"] = np.asarray(read.seq_to_sig_map).astype(np.int32)
            out[f"real{idx}_{algo}_shift_scale"] = np.array([read.shift, read.scale])
        out[f"real{idx}_key"] = np.array(k)
        out[f"real{idx}_levels"] = refiner.extract_levels(int_seq)
        print("real read", k, "bases", int_seq.size, "changed",
              float((out[f"real{idx}_dwell_penalty_ssm"] != ssm).mean()))
    np.savez_compressed(os.path.join(HERE, "refine_real_cases.npz"), **out)
