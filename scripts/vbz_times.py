"""POD5 signal decode: rb200_svb16_decode (CUDA events, packed rows resident in HBM) against the numpy
decoder on one host core.  Bytes counted for the roofline: the packed svb16 stream read + 2 B per sample
written (the zstd layer is host work in both arms and is timed separately).

    python scripts/vbz_times.py [--reads 512] [--samples 100000]
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from remora_b200 import _native, io  # noqa: E402


def vbz_bench(n_reads=512, n_samples=100000, cpu_seconds=3.0, verbose=True):
    class args:  # noqa: N801
        reads, samples = n_reads, n_samples
    rng = np.random.default_rng(0)
    dev = torch.device("cuda:0")
    distinct = []
    for i in range(16):  # nanopore-like: level steps + noise, ~1.1 bytes per sample after svb16
        n = int(rng.integers(args.samples // 2, args.samples * 3 // 2))
        lv = np.repeat(rng.normal(900, 120, size=n // 10 + 1), 10)[:n]
        distinct.append(np.clip(lv + rng.normal(0, 12, size=n), 0, 2047).astype(np.int16))
    sigs = [distinct[i % 16] for i in range(args.reads)]
    rows, counts, owner, blobs = [], [], [], []
    cache = {}
    for k, s in enumerate(sigs):
        for st in range(0, s.size, 102400):
            rows.append(s[st:st + 102400])
            counts.append(rows[-1].size)
            owner.append(k)
            key = (k % 16, st)
            if key not in cache:
                cache[key] = io.encode_vbz(rows[-1])
            blobs.append(cache[key])
    t0 = time.perf_counter()
    raws = [io._zstd_frame_content(b) for b in blobs]
    t_zstd = time.perf_counter() - t0
    n_total = int(sum(counts))
    packed_bytes = int(sum(r.size for r in raws))
    # numpy decoder, one core, bounded sample
    t0 = time.perf_counter()
    done = 0
    for b, c in zip(blobs, counts):
        io.decode_vbz(b, c)
        done += c
        if time.perf_counter() - t0 > cpu_seconds:
            break
    cpu_rate = done / (time.perf_counter() - t0)
    # GPU: one full call (host zstd + upload + kernel), then the kernel alone on resident buffers
    t0 = time.perf_counter()
    d_out, spans = io.decode_vbz_rows_gpu(blobs, counts, dev, owner)
    torch.cuda.synchronize()
    t_call = time.perf_counter() - t0
    for k in (0, 1, len(sigs) - 1):
        st, ln = spans[k]
        assert np.array_equal(d_out[st:st + ln].cpu().numpy(), sigs[k])
    lib = _native.load_library()
    row_off = np.zeros(len(raws) + 1, dtype=np.int64)
    row_off[1:] = np.cumsum([r.size for r in raws])
    packed = np.zeros(int(row_off[-1]) + 16, dtype=np.uint8)
    for r, raw in enumerate(raws):
        packed[row_off[r]:row_off[r + 1]] = np.frombuffer(raw, dtype=np.uint8)
    out_off = np.concatenate([[0], np.cumsum((np.asarray(counts) + 7) & ~7)[:-1]]).astype(np.int64)
    up = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    d_packed, d_row_off, d_n, d_out_off = up(packed), up(row_off), up(np.asarray(counts, np.int32)), up(out_off)
    d_o = torch.empty(int(out_off[-1]) + counts[-1] + 8, dtype=torch.int16, device=dev)
    d_status = torch.zeros(len(raws), dtype=torch.int32, device=dev)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    max_n = int(max(counts))
    need = ctypes.c_int64()
    _native.check(lib.rb200_svb16_scratch_bytes(len(raws), max_n, ctypes.byref(need)), "svb16 scratch")
    d_scratch = torch.empty(max(int(need.value), 4), dtype=torch.uint8, device=dev)
    run = lambda: _native.check(lib.rb200_svb16_decode(ptr(d_packed), ptr(d_row_off), ptr(d_n), ptr(d_out_off),  # noqa: E731
                                                       len(raws), max_n, ptr(d_o), ptr(d_status), ptr(d_scratch),
                                                       stream), "svb16")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    times = []
    for _ in range(7):
        flush.fill_(1.0)  # write more than L2 between timed launches: inputs and outputs come from / go to HBM
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    del flush
    ms = float(np.median(times))
    assert int(d_status.max()) == 0
    gbs = (packed_bytes + 2 * n_total) / (ms * 1e-3) / 1e9
    res = dict(rows=len(raws), samples=n_total, packed_bytes=packed_bytes, bytes_per_sample=packed_bytes / n_total,
               kernel_ms=ms, samples_per_s=n_total / ms * 1e3, algorithmic_gb_per_s=gbs, l2="flushed between launches",
               numpy_samples_per_s_1_core=cpu_rate, host_zstd_samples_per_s=n_total / t_zstd,
               full_call_samples_per_s=n_total / t_call)
    if verbose:
        print(f"[vbz] {len(raws)} rows, {n_total / 1e6:.1f} M samples, {packed_bytes / n_total:.2f} B/sample packed: kernel "
          f"{ms:.3f} ms -> {n_total / ms / 1e6:.1f} G samples/s, {gbs:.0f} GB/s algorithmic; numpy 1 core "
          f"{cpu_rate / 1e6:.1f} M samples/s; host zstd {n_total / t_zstd / 1e6:.0f} M samples/s; full call (zstd + "
          f"upload + kernel) {n_total / t_call / 1e6:.0f} M samples/s", flush=True)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=512)
    ap.add_argument("--samples", type=int, default=100000)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    res = vbz_bench(args.reads, args.samples)
    if args.json:
        json.dump(res, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
