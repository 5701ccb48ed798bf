"""Small, ragged invocations of every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_step.py
Sizes are tiny on purpose (the tools slow kernels down 10-100x); correctness of the outputs is checked
by the pytest suite, this script only has to touch every kernel with awkward shapes."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from remora_b200 import data_chunks, encoded_kmers, inference, model_util  # noqa: E402
from remora_b200 import refine_signal_map as rsm  # noqa: E402
from remora_b200.synth import synth_chunks, synth_levels_table, synth_read, synth_refine_read  # noqa: E402

dev = torch.device("cuda:0")
G = os.path.join(ROOT, "tests", "golden")
only = set(sys.argv[1:])


def want(name):
    return not only or name in only


if want("forward"):
    for name, T in (("convlstm_s64_k9_hot", 100), ("convlstm_s64_k9_hot", 64), ("convlstm_s16_k6_o3", 55),
                    ("conv_s64_k9", 100)):
        model, md = model_util.load_model(os.path.join(G, name + ".pt"), device=dev, eval_only=True)
        impls = ["layers", "tiled"] + (["fused", "fused_tc", "fused_mega", "fused_bf16"]
                                       if name.startswith("convlstm_s64") else []) + \
            (["fused_mega"] if name == "conv_s64_k9" else [])
        for B in (1, 7, 13):
            d = synth_chunks(B, T, tuple(md["kmer_context_bases"]), seed=B)
            args = [torch.from_numpy(d[k]).to(dev) for k in ("signal", "sequence", "sequence_to_signal_mapping",
                                                             "sequence_lengths")]
            enc = encoded_kmers.compute_encoded_kmer_batch_torch(
                *md["kmer_context_bases"], *args[1:], sig_len=T, device=dev)
            for impl in impls + ["auto"]:
                model.set_impl(impl)
                out = model.forward_compact(*args)
                torch.cuda.synchronize()
                assert torch.isfinite(out).all()
                if impl in ("layers", "auto"):  # the dense interface: reference call form
                    out2 = model(args[0], enc)
                    torch.cuda.synchronize()
                    assert torch.isfinite(out2).all()
        if name == "convlstm_s64_k9_hot" and T == 100:
            # mapping arrays wider than the per-base sums' buffer: seq_conv1 as one implicit GEMM
            d = synth_chunks(9, T, tuple(md["kmer_context_bases"]), seed=77, max_seq_len=40, seq_len_range=(20, 40),
                             stride=2)
            args = [torch.from_numpy(d[k]).to(dev) for k in ("signal", "sequence", "sequence_to_signal_mapping",
                                                             "sequence_lengths")]
            model.set_impl("auto")
            out = model.forward_compact(*args)
            torch.cuda.synchronize()
            assert torch.isfinite(out).all() and model.last_impl == "fused_mega"
        print("forward ok", name, T, flush=True)

if want("read"):
    model, md = model_util.load_model(os.path.join(G, "convlstm_s64_k9_refine.pt"), device=dev, eval_only=True)
    table = synth_levels_table(6, 0)
    for seed, n in ((1, 60), (2, 333)):
        dacs, shift, scale, ssm, int_seq = synth_refine_read(n, table, 6, 2, seed=seed)
        for on_dev in (False, True):
            read = data_chunks.RemoraRead(dacs.copy(), shift, scale, ssm.copy(), int_seq.copy())
            inference.call_read_mods(read, model, md, return_mm_ml_tags=True, extract_on_device=on_dev)
    print("read ok", flush=True)

if want("refine"):
    table = synth_levels_table(6, 0)
    for algo in ("dwell_penalty", "Viterbi"):
        refiner = rsm.SigMapRefiner(_levels_array=table, center_idx=2, do_rough_rescale=True, scale_iters=0,
                                    algo=algo, device=dev)
        reads = []
        for i, n in enumerate((12, 40, 130, 77, 260, 31, 19, 90, 55)):
            kw = dict(frac_stall=0.05, stall_range=(300, 1300)) if i % 3 == 0 else {}
            dacs, shift, scale, ssm, int_seq = synth_refine_read(n, table, 6, 2, seed=40 + i, **kw)
            reads.append(data_chunks.RemoraRead(dacs, shift, scale, ssm, int_seq))
        refiner.refine_reads(reads)
        # small shared rows: most bases take the global-scratch rows
        levels = [refiner.extract_levels(r.int_seq) for r in reads]
        bands = [rsm.compute_seq_band(r.seq_to_sig_map - r.seq_to_sig_map[0], lv, 5) for r, lv in zip(reads, levels)]
        batch = rsm.DeviceRefineBatch([r.dacs for r in reads], [r.shift for r in reads], [r.scale for r in reads],
                                      levels, bands, algo, refiner.sd_arr, dev, near_cap=64)
        batch.run()
        batch.paths()
    print("refine ok", flush=True)
