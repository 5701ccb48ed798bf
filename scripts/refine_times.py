"""Signal-mapping refinement throughput: rb200_refine_normalize + rb200_refine_dp (CUDA events, inputs
resident) against the reference's own banded DP (oracle/_ref Cython core when it was built, else the
oracle C restatement) on the box's host cores.

    python scripts/refine_times.py [--reads 4096] [--bases 1000] [--json out.json]

Units: reads/s, bases/s and DP cells/s (one cell = one (base, sample) pair inside the band; the
algorithmic work unit of refine_signal_map_core.pyx).  End to end (host numpy set-up + uploads + kernels
+ read-back) is reported separately.
"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from remora_b200 import refine_signal_map as rsm  # noqa: E402
from remora_b200.synth import synth_levels_table, synth_refine_read  # noqa: E402

K, C = 6, 2


def make_reads(n_reads, n_bases, seed=0, distinct=64):
    """`distinct` different synthetic reads, repeated to n_reads (generation is the slow part)."""
    table = synth_levels_table(K, 0)
    rng = np.random.default_rng(seed)
    base = []
    for i in range(min(distinct, n_reads)):
        n = int(rng.integers(max(12, n_bases // 2), n_bases * 3 // 2 + 1))
        base.append(synth_refine_read(n, table, K, C, seed=seed * 1000 + i, frac_stall=0.002))
    return table, [base[i % len(base)] for i in range(n_reads)]


def gpu_leg(table, reads, algo, iters=5, near_cap=None):
    refiner = rsm.SigMapRefiner(_levels_array=table, center_idx=C, do_rough_rescale=True, scale_iters=0,
                                algo=algo, device=torch.device("cuda:0"))
    t0 = time.perf_counter()
    levels = [refiner.extract_levels(r[4]) for r in reads]
    resc = [refiner.rough_rescale(r[1], r[2], r[3], r[4], r[0]) for r in reads]
    bands = [rsm.compute_seq_band(r[3] - r[3][0], lv, 5) for r, lv in zip(reads, levels)]
    t_host = time.perf_counter() - t0
    t0 = time.perf_counter()
    batch = rsm.DeviceRefineBatch([r[0][r[3][0]:r[3][-1]] for r in reads], [x[0] for x in resc],
                                  [x[1] for x in resc], levels, bands, algo, refiner.sd_arr,
                                  torch.device("cuda:0"), near_cap=near_cap)
    torch.cuda.synchronize()
    t_up = time.perf_counter() - t0
    for _ in range(2):
        batch.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        batch.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    t0 = time.perf_counter()
    paths = batch.paths()
    t_down = time.perf_counter() - t0
    n_bases = int(batch.seq_off[-1])
    return dict(algo=algo, reads=len(reads), bases=n_bases, cells=batch.cells, widest_band=batch.widest,
                near_cap=batch.near_cap,
                kernel_ms=ms, reads_per_s=len(reads) / ms * 1e3, bases_per_s=n_bases / ms * 1e3,
                cells_per_s=batch.cells / ms * 1e3, host_setup_s=t_host, upload_s=t_up, readback_s=t_down,
                e2e_reads_per_s=len(reads) / (t_host + t_up + ms / 1e3 + t_down)), paths, levels, resc, bands


def cpu_leg(reads, levels, resc, bands, algo, budget_s=10.0, threads=None):
    """The reference's banded DP on host threads (the Cython core releases nothing, so processes would
    be needed for the reference; the C restatement is called through ctypes, which drops the GIL)."""
    import build_ref
    import refine_oracle as ro
    core = build_ref.load_ref_refine_core()
    sigs = [((r[0][r[3][0]:r[3][-1]] - x[0]) / x[1]).astype(np.float32) for r, x in zip(reads, resc)]
    lv0 = [np.where(np.isnan(lv), 0, lv).astype(np.float32) for lv in levels]
    out = {}
    # single thread, reference core
    if core is not None:
        t0 = time.perf_counter()
        done = cells = 0
        for sig, lv, band in zip(sigs, lv0, bands):
            core.seq_banded_dp(sig, lv, band, ro.DEFAULT_SD_ARR, algo)
            done += 1
            cells += int((band[1] - band[0]).sum())
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
        out["reference_1_thread"] = dict(reads=done, seconds=dt, reads_per_s=done / dt, cells_per_s=cells / dt)
    # all host threads, oracle C restatement (same algorithm, GIL-free)
    threads = threads or min(os.cpu_count() or 1, 64)
    ro.seq_banded_dp(sigs[0], lv0[0], bands[0], ro.DEFAULT_SD_ARR, algo)
    sample = list(zip(sigs, lv0, bands))
    t_one = time.perf_counter()
    ro.seq_banded_dp(*sample[0], ro.DEFAULT_SD_ARR, algo)
    t_one = time.perf_counter() - t_one
    n = int(min(len(sample), max(threads, budget_s / max(t_one, 1e-6) * threads / 2)))
    sample = sample[:n]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda a: ro.seq_banded_dp(a[0], a[1], a[2], ro.DEFAULT_SD_ARR, algo), sample))
    dt = time.perf_counter() - t0
    cells = sum(int((b[1] - b[0]).sum()) for _, _, b in sample)
    out["port_all_threads"] = dict(reads=n, threads=threads, seconds=dt, reads_per_s=n / dt,
                                   cells_per_s=cells / dt)
    return out


def refine_bench(n_reads=2048, n_bases=600, cpu_seconds=2.0):
    """Compact form for bench.py's `next_rows` object: the default (dwell_penalty) refinement of a
    synthetic batch on the GPU and a bounded sample of the same reads through the reference's DP."""
    table, reads = make_reads(n_reads, n_bases)
    g, paths, levels, resc, bands = gpu_leg(table, reads, "dwell_penalty", iters=3)
    cpu = cpu_leg(reads, levels, resc, bands, "dwell_penalty", budget_s=cpu_seconds)
    ref = cpu.get("reference_1_thread") or {}
    port = cpu.get("port_all_threads") or {}
    return {
        "kernel": "refine_dp_kernel (+ refine_normalise_kernel)", "unit": "DP cells/s",
        "workload": f"{g['reads']} synthetic reads, {g['bases']} bases, {g['cells']} band cells, "
                    f"half bandwidth 5, dwell_penalty, widest band {g['widest_band']} samples",
        "value": g["cells_per_s"], "ms": g["kernel_ms"], "reads_per_s": g["reads_per_s"],
        "bases_per_s": g["bases_per_s"], "near_cap": g["near_cap"],
        "bound": "instruction issue / dependent FADD+FMNMX chain (exact fp32 rounding order forbids a scan)",
        "cpu_reference_1_thread_cells_per_s": ref.get("cells_per_s"),
        "cpu_port_threads": port.get("threads"), "cpu_port_cells_per_s": port.get("cells_per_s"),
        "parity": "paths bit-identical to the reference (tests/test_refine.py)",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=4096)
    ap.add_argument("--bases", type=int, default=1000)
    ap.add_argument("--json", default=None)
    ap.add_argument("--cpu-seconds", type=float, default=8.0)
    ap.add_argument("--near-caps", default="", help="comma list: also time these shared-memory row capacities")
    args = ap.parse_args()
    table, reads = make_reads(args.reads, args.bases)
    result = {}
    for algo in ("dwell_penalty", "Viterbi"):
        g, paths, levels, resc, bands = gpu_leg(table, reads, algo)
        print(f"[gpu] {algo}: {g['reads']} reads, {g['bases']} bases, {g['cells'] / 1e6:.1f} M cells, "
              f"widest band {g['widest_band']}, near_cap {g['near_cap']}: {g['kernel_ms']:.2f} ms -> {g['reads_per_s'] / 1e3:.1f} k reads/s, "
              f"{g['bases_per_s'] / 1e6:.1f} M bases/s, {g['cells_per_s'] / 1e9:.2f} G cells/s; "
              f"host set-up {g['host_setup_s']:.2f} s, upload {g['upload_s']:.2f} s, read-back "
              f"{g['readback_s']:.2f} s -> e2e {g['e2e_reads_per_s']:.0f} reads/s", flush=True)
        result[algo] = g
        if algo == "dwell_penalty":
            import refine_oracle as ro
            for i in (0, 1, len(reads) - 1):  # spot check against the oracle
                want = ro.seq_banded_dp(
                    ((reads[i][0][reads[i][3][0]:reads[i][3][-1]] - resc[i][0]) / resc[i][1]).astype(np.float32),
                    np.where(np.isnan(levels[i]), 0, levels[i]), bands[i], ro.DEFAULT_SD_ARR, algo)[1]
                assert np.array_equal(paths[i], want), i
            c = cpu_leg(reads, levels, resc, bands, algo, args.cpu_seconds)
            for k, v in c.items():
                print(f"[cpu] {k}: {v}", flush=True)
            result["cpu"] = c
    for cap in [int(x) for x in args.near_caps.split(",") if x]:
        g = gpu_leg(table, reads, "dwell_penalty", near_cap=cap)[0]
        print(f"[gpu] dwell_penalty near_cap {cap}: {g['kernel_ms']:.2f} ms -> {g['cells_per_s'] / 1e9:.2f} G cells/s",
              flush=True)
        result[f"near_cap_{cap}"] = g
    if args.json:
        with open(args.json, "w") as fh:
            json.dump(result, fh, indent=1)


if __name__ == "__main__":
    main()
