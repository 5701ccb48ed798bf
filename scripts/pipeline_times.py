"""End-to-end file pipeline: POD5 + BAM -> signal-mapping refinement -> chunks -> ConvLSTM_w_ref -> MM/ML
tags (`inference.infer_from_pod5_and_bam`) on a synthetic run written by remora_b200.io's own writers.
Wall clock, host stages included (this is the row where the host, not the GPU, is the limit).

    python scripts/pipeline_times.py [--reads 256] [--bases 2000]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from remora_b200 import inference, model_util  # noqa: E402
from remora_b200.synth import synth_pod5_bam_run  # noqa: E402


def pipeline_bench(n_reads=128, n_bases=2000):
    """Compact form for bench.py's `next_rows`: wall clock of infer_from_pod5_and_bam (refiner-carrying
    fixture model, BAM output) on a synthetic run written by our own POD5/BAM writers."""
    dev = torch.device("cuda:0")
    with tempfile.TemporaryDirectory() as tmp:
        pod5, bam, truth = synth_pod5_bam_run(os.path.join(tmp, "run.pod5"), os.path.join(tmp, "run.bam"),
                                              n_reads=n_reads, bases=(n_bases // 2, n_bases * 3 // 2))
        model, md = model_util.load_model(os.path.join(ROOT, "tests", "golden", "convlstm_s64_k9_refine.pt"),
                                          device=dev, eval_only=True)
        inference.infer_from_pod5_and_bam(pod5, bam, (model, md), num_reads=8)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = inference.infer_from_pod5_and_bam(pod5, bam, (model, md), out_path=os.path.join(tmp, "o.bam"),
                                                reads_per_batch=256)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    calls = sum(len(r["ml"]) for r in res)
    bases = sum(len(t["seq"]) for t in truth.values())
    return {"unit": "reads/s", "value": len(res) / dt, "bases_per_s": bases / dt, "calls_per_s": calls / dt,
            "workload": f"{len(res)} synthetic reads, {bases} bases, {calls} CG calls: POD5 decode (GPU) + BAM join "
                        "+ rough re-scaling (host) + banded-DP refinement (GPU) + chunk extraction (host) + "
                        "ConvLSTM_w_ref (GPU) + MM/ML tags + BAM output, one process, wall clock",
            "bound": "host (numpy re-scaling, record parsing, tag formatting); the GPU stages are a few percent"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=256)
    ap.add_argument("--bases", type=int, default=2000)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        t0 = time.perf_counter()
        pod5, bam, truth = synth_pod5_bam_run(os.path.join(tmp, "run.pod5"), os.path.join(tmp, "run.bam"),
                                              n_reads=args.reads, bases=(args.bases // 2, args.bases * 3 // 2))
        print(f"wrote {args.reads} reads ({os.path.getsize(pod5) / 1e6:.1f} MB pod5, "
              f"{os.path.getsize(bam) / 1e6:.1f} MB bam) in {time.perf_counter() - t0:.1f} s", flush=True)
        for name in ("convlstm_s64_k9_hot", "convlstm_s64_k9_refine"):
            model, md = model_util.load_model(os.path.join(ROOT, "tests", "golden", name + ".pt"), device=dev,
                                              eval_only=True)
            for on_dev in (True, False):
                inference.infer_from_pod5_and_bam(pod5, bam, (model, md), num_reads=8, extract_on_device=on_dev)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                res = inference.infer_from_pod5_and_bam(pod5, bam, (model, md), out_path=os.path.join(tmp, "o.sam"),
                                                        extract_on_device=on_dev, reads_per_batch=256)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                calls = sum(len(r["ml"]) for r in res)
                bases = sum(len(t["seq"]) for t in truth.values())
                key = f"{name}{'' if on_dev else '_host_chunks'}"
                out[key] = dict(reads=len(res), bases=bases, calls=calls, seconds=dt, reads_per_s=len(res) / dt,
                                bases_per_s=bases / dt, calls_per_s=calls / dt)
                print(f"[pipeline] {key}: {len(res)} reads, {bases} bases, {calls} calls in {dt:.2f} s -> "
                      f"{len(res) / dt:.0f} reads/s, {bases / dt / 1e3:.0f} k bases/s, {calls / dt / 1e3:.1f} k calls/s",
                      flush=True)
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
