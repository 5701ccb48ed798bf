"""Golden vectors for the POD5/BAM input path, generated in the build container.

    python tests/golden/make_golden_io.py

The reference reads these formats through pysam / pod5, which are not installed here, so the records
come from remora_b200.io's own parsers and are then handed to the REFERENCE's code:
reference io.Read.from_pod5_and_alignment -> into_remora_read -> inference.call_read_mods (CPU) with the
committed fixture models.  Output tests/golden/io_cases.npz holds, for two reads of each of the
reference's test files (tests/data/{can,mod}_reads.pod5 + {can,mod}_mappings.bam): the arrays of the
reference's RemoraRead (what the readers + join must reproduce) and the reference's calls with and
without signal-mapping refinement.
"""
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_harness  # noqa: E402

ref_harness.import_reference()
from remora import inference, io as ref_io, model_util  # noqa: E402
from remora_b200 import io as rio  # noqa: E402

DATA = os.path.join(ref_harness.REFERENCE_ROOT, "tests", "data")
torch.set_grad_enabled(False)
torch.set_num_threads(4)


def main():
    out = {}
    index = []
    models = {name: model_util.load_model(os.path.join(HERE, name + ".pt"), eval_only=True)
              for name in ("convlstm_s64_k9_hot", "convlstm_s64_k9_refine")}
    for stem in ("can", "mod"):
        bam_idx = rio.ReadIndexedBam(os.path.join(DATA, f"{stem}_mappings.bam"))
        with rio.Pod5Reader(os.path.join(DATA, f"{stem}_reads.pod5")) as pod5:
            picked = sorted(rid for rid in pod5.read_ids if rid in bam_idx)[:2]
            for rid in picked:
                rec = bam_idx.get_first_alignment(rid)
                io_read = ref_io.Read.from_pod5_and_alignment(pod5.get_read(rid), rec)
                k = f"{stem}_{len(index)}_"
                for anchor, tag in ((False, "bc"), (True, "ref")):
                    rr = io_read.into_remora_read(anchor)
                    dacs = rr.dacs.astype(np.int16)
                    # the reference-anchored samples are a slice of the same signal: keep a checksum only
                    out.update({k + tag + "_dacs": dacs if tag == "bc" else
                                np.array([dacs.size, zlib.crc32(dacs.tobytes())], dtype=np.int64),
                                k + tag + "_ssm": rr.seq_to_sig_map,
                                k + tag + "_int_seq": rr.int_seq.astype(np.int8),
                                k + tag + "_shift_scale": np.array([rr.shift, rr.scale], dtype=np.float64)})
                for name, (model, md) in models.items():
                    rr = io_read.into_remora_read(False)
                    nn_out, _, pos = inference.call_read_mods(rr, model, md)
                    mm, ml = inference.call_read_mods(io_read.into_remora_read(False), model, md,
                                                      return_mm_ml_tags=True)
                    out.update({k + name + "_nn_out": nn_out.astype(np.float32), k + name + "_pos": pos,
                                k + name + "_ssm": np.asarray(rr.seq_to_sig_map),
                                k + name + "_shift_scale": np.array([rr.shift, rr.scale]),
                                k + name + "_mm": np.array(mm), k + name + "_ml": np.frombuffer(ml, dtype=np.uint8)})
                    print(stem, rid, name, "calls", pos.size)
                # reference-anchored calls (remora infer --reference-anchored): sequence and mapping from
                # the alignment instead of the basecalls
                model, md = models["convlstm_s64_k9_hot"]
                nn_out, _, pos = inference.call_read_mods(io_read.into_remora_read(True), model, md)
                out.update({k + "ref_nn_out": nn_out.astype(np.float32), k + "ref_pos": pos})
                index.append([stem, rid, len(io_read.seq), io_read.dacs.size, io_read.ref_reg.strand])
    out["index"] = np.array(index)
    np.savez_compressed(os.path.join(HERE, "io_cases.npz"), **out)


if __name__ == "__main__":
    main()
