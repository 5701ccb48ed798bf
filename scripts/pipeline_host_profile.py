"""Host side of the file pipeline without a GPU: infer_from_pod5_and_bam on a synthetic run with the three GPU
stages replaced by instant stand-ins (network = zeros, banded DP = the mapping it was given, signal decode =
numpy), so that cProfile shows what the HOST costs per read - the part that bounds the pipeline.

    python scripts/pipeline_host_profile.py [--reads 256] [--bases 2000] [--top 30]
"""
import argparse
import cProfile
import os
import pstats
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from remora_b200 import inference, model_util, util  # noqa: E402
from remora_b200 import refine_signal_map as rsm  # noqa: E402
from remora_b200.synth import synth_pod5_bam_run  # noqa: E402


class InstantModel(torch.nn.Module):
    def __init__(self, num_out):
        super().__init__()
        self.num_out = num_out
        self.anchor = torch.nn.Parameter(torch.zeros(1), requires_grad=False)

    def forward_compact(self, sig, seq, mp, ln, out=None):
        return torch.zeros((sig.shape[0], self.num_out), dtype=torch.float32)

    def softmax_ml(self, logits, want_probs=True):
        n = logits.shape[0]
        ml = torch.full((n, self.num_out - 1), 128, dtype=torch.uint8)
        return (torch.full((n, self.num_out - 1), 0.5) if want_probs else None), ml


def instant_dp(dacs_list, shifts, scales, levels_list, seq_bands, *a, **k):
    # a valid path inside every band: the band's lower edge made strictly usable is not needed for timing -
    # return evenly spread boundaries
    out = []
    for d, lv in zip(dacs_list, levels_list):
        out.append(np.linspace(0, len(d), len(lv) + 1).astype(np.int64))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=256)
    ap.add_argument("--bases", type=int, default=2000)
    ap.add_argument("--top", type=int, default=30)
    args = ap.parse_args()
    rsm.banded_dp_batch = instant_dp
    sd, md = model_util._raw_load_torchscript(os.path.join(ROOT, "tests", "golden", "convlstm_s64_k9_refine.pt"))
    model_util.add_derived_metadata(md)
    model = InstantModel(len(md["mod_bases"]) + 1)
    with tempfile.TemporaryDirectory() as tmp:
        pod5, bam, truth = synth_pod5_bam_run(os.path.join(tmp, "r.pod5"), os.path.join(tmp, "r.bam"),
                                              n_reads=args.reads, bases=(args.bases // 2, args.bases * 3 // 2))
        kw = dict(decode_on_device=False, extract_on_device=False, reads_per_batch=256,
                  out_path=os.path.join(tmp, "o.bam"))
        inference.infer_from_pod5_and_bam(pod5, bam, (model, md), num_reads=8, **kw)
        dts = []
        for _ in range(5):
            t0 = time.perf_counter()
            res = inference.infer_from_pod5_and_bam(pod5, bam, (model, md), **kw)
            dts.append(time.perf_counter() - t0)
        dt = min(dts)
        ok = sum(r["error"] is None for r in res)
        print(f"{len(res)} reads ({ok} called) in {dt:.3f} s -> {len(res) / dt:.0f} reads/s host side "
              f"({dt / len(res) * 1e6:.0f} us per read)", flush=True)
        pr = cProfile.Profile()
        pr.enable()
        inference.infer_from_pod5_and_bam(pod5, bam, (model, md), **kw)
        pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(args.top)


if __name__ == "__main__":
    main()
