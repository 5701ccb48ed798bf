#!/usr/bin/env python
"""Benchmark of the Remora per-chunk modified-base inference hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): chunks/sec at chunk_len=100, kmer_context=(4,4), ConvLSTM_w_ref (size 64),
batch=1024 per GPU.  One "step" = one batch of 1024 synthetic chunks through the hot path
(k-mer encode fused into the forward).  `value` times K steps with the compact chunk arrays
already resident in HBM; `e2e` times the same steps through the host-buffer C-ABI call
(rb200_infer_host: pinned staging, H2D, kernels, D2H) every step.  Weights are seeded random
(tests/golden/convlstm_s64_k9_hot.pt, exported by the reference's own exporter) and data is
synthetic: no datasets or checkpoints are reachable.

Timing hygiene: W>=3 warm-up steps; every step reads a different batch of a resident pool that is
larger than the 126 MB L2 (inputs come from HBM); device timing with CUDA events bracketed by
barrier + synchronize; max over ranks; SM clocks sampled with nvidia-smi during the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHUNK_LEN = 100
KMER_CONTEXT = (4, 4)
BATCH = 1024
MODEL_PT = os.path.join(ROOT, "tests", "golden", "convlstm_s64_k9_hot.pt")
# algorithmic bytes / flops per chunk
BYTES_IFACE_A = 4 * CHUNK_LEN + 4 * 36 * CHUNK_LEN + 4 * 2   # 14 808 B: what the reference model call consumes
DENSE_MFLOP = 6.989312                                       # 3 494 656 MAC, ConvLSTM_w_ref T=100
METRIC = "chunks/sec (chunk_len=100, batch=1024)"
# the workload both arms run (identical string in both lines)
WORKLOAD = ("synthetic chunks, chunk_len=100, kmer_context=(4,4), ConvLSTM_w_ref size 64 (134082 params), "
            "batch=1024 per device (BASELINE configs[1]), fp32")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic_bytes(kernel_name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel_name` from the committed
    `ncu --set full` summary (profiles/, cold-cache capture of the same 1024-chunk step), or None."""
    import glob
    import re
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_summary_r*.json"))):
        try:
            rows = json.load(open(path))
        except (OSError, ValueError):
            continue
        for row in rows:
            if kernel_name not in row.get("Kernel Name", ""):
                continue
            total = 0.0
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                m = re.match(r"([0-9.]+)\s*(\w+)?", row.get(key, ""))
                if not m:
                    total = None
                    break
                unit = (m.group(2) or "byte").lower()
                scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
                total += float(m.group(1)) * scale
            if total is not None:
                best = (total, os.path.basename(path))
    return best


def ncu_pipe_counters(kernel_name):
    """sm__pipe_tensor / fma 'cycles active' percentages of `kernel_name` from the newest committed
    `ncu --set full` summary under profiles/ (None when absent)."""
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_summary_r*.json"))):
        try:
            rows = json.load(open(path))
        except (OSError, ValueError):
            continue
        for row in rows:
            if kernel_name not in row.get("Kernel Name", ""):
                continue
            got = {"source": os.path.basename(path)}
            for key, val in row.items():
                lk = key.lower()
                if "pipe_tensor" in lk and "pct" in lk and "tensor" not in got:
                    got["tensor"] = val
                if "pipe_fma" in lk and "pct" in lk and "fma" not in got:
                    got["fma"] = val
            best = got
    return best


def make_pool(n_batches, seed):
    from remora_b200.synth import synth_chunks
    d = synth_chunks(n_batches * BATCH, CHUNK_LEN, KMER_CONTEXT, seed=seed)
    return d


# ------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the reference's own CPU implementation of the path
# ------------------------------------------------------------------------------------------------
def cpu_reference_runner():
    """Returns (step_fn(batch_dict) -> logits, kind, description).  encode = the reference's own
    Cython encoder compiled into oracle/_ref (else the C restatement); forward = the TorchScript
    module exported by the reference's export_model_torchscript, run on CPU by torch.jit."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import remora_oracle as ro
    ref_enc = ro.load_ref_encoder()
    kind = "reference" if ref_enc is not None else "port"
    torch.set_grad_enabled(False)
    module = torch.jit.load(MODEL_PT, map_location="cpu").eval()

    def step(d):
        if ref_enc is not None:
            enc = ref_enc.compute_encoded_kmer_batch(KMER_CONTEXT[0], KMER_CONTEXT[1], d["sequence"],
                                                     d["sequence_to_signal_mapping"],
                                                     d["sequence_lengths"])
        else:
            enc = ro.encode_kmers_c(KMER_CONTEXT[0], KMER_CONTEXT[1], d["sequence"],
                                    d["sequence_to_signal_mapping"], d["sequence_lengths"])
        return module(torch.from_numpy(d["signal"]), torch.from_numpy(np.asarray(enc)))

    desc = ("encode: reference Cython compute_encoded_kmer_batch (oracle/_ref, single thread as in "
            "the reference) + forward: reference TorchScript module on CPU") if kind == "reference" \
        else "encode: C restatement (oracle/oracle_encode.c) + forward: reference TorchScript module on CPU"
    return step, kind, desc


def tune_cpu_threads(step, pool):
    """torch's default (all cores) is far from the best setting for 1024-chunk batches on a many-core
    host (on the 128-core bench box it is ~200x slower than 16-32 threads), so the baseline uses the
    intra-op thread count that maximises the reference's throughput: sweep upwards, stop when it
    clearly degrades.  Returns (best_threads, {threads: chunks/s})."""
    import torch
    n_cpu = os.cpu_count() or 1
    cands = sorted({min(n_cpu, t) for t in (4, 8, 16, 32, 64, 128, n_cpu)})
    seen, best_t, best_v = {}, cands[0], 0.0
    for t in cands:
        torch.set_num_threads(t)
        step(slice_batch(pool, 0))
        t0 = time.perf_counter()
        for i in range(2):
            step(slice_batch(pool, 1 + i))
        v = 2 * BATCH / (time.perf_counter() - t0)
        seen[t] = round(v, 1)
        if v > best_v:
            best_t, best_v = t, v
        elif v < 0.6 * best_v:
            break
    torch.set_num_threads(best_t)
    return best_t, seen


def slice_batch(pool, i):
    sl = slice(i * BATCH, (i + 1) * BATCH)
    return {k: (v[sl] if isinstance(v, np.ndarray) else v) for k, v in pool.items()}


def run_cpu_sample(max_seconds=10.0, max_batches=2000, warm=2):
    step, kind, desc = cpu_reference_runner()
    pool = make_pool(8, seed=99)
    threads, sweep = tune_cpu_threads(step, pool)
    desc += f"; torch intra-op threads tuned over {sweep} (chunks/s) -> {threads} of {os.cpu_count()} host cores"
    for i in range(warm):
        step(slice_batch(pool, i % 8))
    t0 = time.perf_counter()
    n = 0
    while n < max_batches and time.perf_counter() - t0 < max_seconds:
        step(slice_batch(pool, n % 8))
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n * BATCH / dt, "unit": "chunks/s", "cores": threads, "kind": kind,
            "sample": f"{n} batches of {BATCH} chunks (chunk_len {CHUNK_LEN}) in {dt:.1f} s; {desc}"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, kind, desc = cpu_reference_runner()
    pool = make_pool(8, seed=99)
    threads, sweep = tune_cpu_threads(step, pool)
    desc += f"; torch intra-op threads tuned over {sweep} (chunks/s) -> {threads} of {os.cpu_count()} host cores"
    for i in range(args.warmup):
        step(slice_batch(pool, i % 8))
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(slice_batch(pool, i % 8))
    dt = time.perf_counter() - t0
    value = args.steps * BATCH / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "chunks/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "global_batch": BATCH, "chunk_len": CHUNK_LEN,
                   "kmer_context": list(KMER_CONTEXT),
                   "arm": "the reference's own CPU implementation on this box's host cores (rank 0 only)"},
        "cpu_baseline": {"value": value, "unit": "chunks/s", "cores": threads, "kind": kind,
                         "sample": f"{args.steps} batches of {BATCH} chunks; {desc}"},
        "e2e": {"value": value, "unit": "chunks/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



# ------------------------------------------------------------------------------------------------
# the second bar of SURVEY.md 2.2 / BASELINE.md 3.5: the reference's STOCK GPU path on the same B200
# ------------------------------------------------------------------------------------------------
def gpu_stock_baseline(device, steps, warmup):
    """The reference's own GPU path: its TorchScript module moved to the B200 and called with dense
    inputs already resident in HBM, `models[can_base](sigs, enc_kmers)` (src/remora/inference.py:311-314,
    data_chunks.py:528-533) - every op is a cuDNN / cuBLAS / ATen library kernel, none of this repo's.
    Dense inputs come from the oracle's C encoder (untimed).  fp32 under three TF32 settings: torch's
    defaults (what `remora infer --device N` gets), TF32 off everywhere, TF32 on everywhere."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import remora_oracle as ro
    n_b = 16  # 16 x 1024 x 14.8 KB = 242 MB of dense inputs: larger than L2
    pool = make_pool(n_b, seed=7)
    enc = ro.encode_kmers_c(KMER_CONTEXT[0], KMER_CONTEXT[1], pool["sequence"],
                            pool["sequence_to_signal_mapping"], pool["sequence_lengths"])
    sigs_d = torch.from_numpy(pool["signal"]).to(device)
    enc_d = torch.from_numpy(np.asarray(enc)).to(device)
    module = torch.jit.load(MODEL_PT, map_location=device).eval()
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    out = {"what": "reference TorchScript module on this GPU via cuDNN/cuBLAS/ATen, dense (sigs, enc_kmers) "
                   "resident in HBM (src/remora/inference.py:311-314 semantics), batch 1024, chunk_len 100",
           "unit": "chunks/s", "steps": steps, "warmup": max(warmup, 8),
           "torch_default_tf32": {"cudnn": bool(saved[0]), "matmul": bool(saved[1])}}
    try:
        with torch.no_grad():
            for name, tf in (("torch_defaults", saved), ("tf32_off", (False, False)), ("tf32_on", (True, True))):
                torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf
                for i in range(max(warmup, 8)):  # includes the TorchScript profiling runs
                    sl = slice((i % n_b) * BATCH, (i % n_b + 1) * BATCH)
                    module(sigs_d[sl], enc_d[sl])
                torch.cuda.synchronize(device)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(steps):
                    sl = slice((i % n_b) * BATCH, (i % n_b + 1) * BATCH)
                    res = module(sigs_d[sl], enc_d[sl])
                e1.record()
                torch.cuda.synchronize(device)
                ms = e0.elapsed_time(e1)
                out[name] = {"value": steps * BATCH / (ms * 1e-3), "ms_per_step": ms / steps}
            out["finite"] = bool(torch.isfinite(res).all())
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    return out


def config_legs(device, hot_model):
    """Driver-visible numbers for the other BASELINE.json configs (parity of each is in tests/):
    config 3 = Conv_w_ref at batch 4096 (stock chunk_len 100 and the chunk_len-200 / 704-input-classifier
    variant SURVEY.md 8d prescribes), and the reference's dense call form model(sigs, enc_kmers)."""
    import torch
    from remora_b200 import model_util
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import remora_oracle as ro
    from remora_b200.synth import synth_chunks
    legs = {}

    def timed(fn, n_items, steps=30, warm=5):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(device)
        ms = e0.elapsed_time(e1) / steps
        return {"value": n_items / (ms * 1e-3), "unit": "chunks/s", "ms_per_step": ms, "steps": steps}

    for name, pt, T in (("conv_w_ref_b4096_T100", "conv_s64_k9.pt", 100),
                        ("conv_w_ref_b4096_T200_fc704", "conv_s64_k9_T200.pt", 200)):
        try:
            model, md = model_util.load_model(os.path.join(ROOT, "tests", "golden", pt), device=device,
                                              eval_only=True)
            B = 4096
            # 8 different batches (8 x 4096 x ~0.5-0.9 KB); activations between layers dominate traffic
            ds = [synth_chunks(B, T, KMER_CONTEXT, seed=50 + k) for k in range(4)]
            dev = [[torch.from_numpy(d[k]).to(device) for k in
                    ("signal", "sequence", "sequence_to_signal_mapping", "sequence_lengths")] for d in ds]
            it = [0]

            def fn():
                it[0] += 1
                model.forward_compact(*dev[it[0] % len(dev)])
            l0 = model.launch_count
            legs[name] = timed(fn, B)
            legs[name].update({"impl": model.last_impl, "batch": B, "chunk_len": T, "dtype": "f32",
                               "launches_per_step": (model.launch_count - l0) / (legs[name]["steps"] + 5)})
            del model, dev
        except Exception as e:  # noqa: BLE001
            legs[name] = {"error": str(e)[:200]}
    try:
        d = make_pool(8, seed=11)
        enc = ro.encode_kmers_c(KMER_CONTEXT[0], KMER_CONTEXT[1], d["sequence"],
                                d["sequence_to_signal_mapping"], d["sequence_lengths"])
        sig_d, enc_d = torch.from_numpy(d["signal"]).to(device), torch.from_numpy(np.asarray(enc)).to(device)
        it = [0]

        def fn_dense():
            it[0] += 1
            sl = slice((it[0] % 8) * BATCH, (it[0] % 8 + 1) * BATCH)
            hot_model(sig_d[sl], enc_d[sl])
        legs["dense_interface_b1024_T100"] = timed(fn_dense, BATCH, steps=100)
        legs["dense_interface_b1024_T100"].update({
            "impl": hot_model.last_impl, "what": "model(sigs, enc_kmers) with the materialised one-hot, the "
            "reference's own call form (14.8 KB/chunk read from HBM)"})
    except Exception as e:  # noqa: BLE001
        legs["dense_interface_b1024_T100"] = {"error": str(e)[:200]}
    return legs

# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def ours(args):
    import torch
    import torch.distributed as dist
    from remora_b200 import model_util

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    model, md = model_util.load_model(MODEL_PT, device=device, eval_only=True)
    assert md["chunk_len"] == CHUNK_LEN and md["kmer_len"] == 9
    # resident pool: > L2 (126 MB) of compact inputs so every step streams its batch from HBM.
    n_pool = args.pool_batches
    pool = make_pool(n_pool, seed=1 + rank)
    dev_pool = {k: torch.from_numpy(pool[k]).to(device) for k in
                ("signal", "sequence", "sequence_to_signal_mapping", "sequence_lengths")}
    pool_bytes = sum(v.numel() * v.element_size() for v in dev_pool.values())

    def dev_batch(i):
        sl = slice((i % n_pool) * BATCH, (i % n_pool + 1) * BATCH)
        return (dev_pool["signal"][sl], dev_pool["sequence"][sl],
                dev_pool["sequence_to_signal_mapping"][sl], dev_pool["sequence_lengths"][sl])

    def barrier():
        if distributed:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(device)

    # The only exchange step the path has: 8 B/chunk of logits gathered to every rank (NCCL all-gather
    # over NVLink).  Every step's logits are gathered inside the timed region; the steps of a group of G
    # write into one slot of a double-buffered device ring and ONE asynchronous all-gather ships the group
    # (G x 8 KB per rank), so the collective's host-side launch cost (tens of microseconds, comparable to a
    # whole 104 us step) is paid once per group and the transfer overlaps the next group's kernels.
    G = max(1, args.gather_every)
    peer_ring, gather_mode = None, None
    if distributed:
        gather_mode = args.gather
        if gather_mode == "p2p":
            # the exchange fused into the kernel: the classifier epilogue stores into every rank's ring over
            # NVLink (symmetric memory), no collective launch at all; NCCL stays as the fallback
            try:
                from remora_b200.parallel import PeerLogitRing
                peer_ring = PeerLogitRing(model, slots=2, steps=G, batch=BATCH, multicast=args.multicast,
                                          deferred=not args.immediate)
                model.forward_compact(*dev_batch(0))
                if model.last_impl not in ("fused_mega", "fused_bf16"):
                    raise RuntimeError("single-kernel path not selected")
            except Exception as e:  # noqa: BLE001
                if rank == 0:
                    print(f"[bench] p2p gather unavailable ({str(e)[:200]}); using NCCL all-gather", file=sys.stderr)
                peer_ring, gather_mode = None, "nccl"
            # every rank must take the same path
            agree = torch.tensor([1 if peer_ring is not None else 0], dtype=torch.int32, device=device)
            dist.all_reduce(agree, op=dist.ReduceOp.MIN)
            if int(agree.item()) == 0:
                peer_ring, gather_mode = None, "nccl"
        local_ring = torch.empty((2, G, BATCH, model.num_out), dtype=torch.float32, device=device)
        gath_ring = torch.empty((2, world * G * BATCH, model.num_out), dtype=torch.float32, device=device)
    pending = [None, None]

    def step(i, last=False):
        if not distributed:
            model.forward_compact(*dev_batch(i))
            return
        slot, k = (i // G) & 1, i % G
        if peer_ring is not None:
            peer_ring.forward(model, dev_batch(i), slot, k, signal=args.signal)
            return
        if k == 0 and pending[slot] is not None:
            pending[slot].wait()  # stream-side wait: the gather that still reads this slot
            pending[slot] = None
        model.forward_compact(*dev_batch(i), out=local_ring[slot, k])
        if k == G - 1 or last:
            pending[slot] = dist.all_gather_into_tensor(gath_ring[slot], local_ring[slot].view(G * BATCH, -1),
                                                        async_op=True)

    def drain():
        if peer_ring is not None:
            peer_ring.flush()  # deferred exchange: the last step's block leaves with a one-block launch
        for slot in range(2):
            if pending[slot] is not None:
                pending[slot].wait()
                pending[slot] = None

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.6)  # nvidia-smi needs a few hundred ms before its first sample (GPU idle: before warm-up)
    # every rank starts the settle phase together: a rank that got here early would otherwise settle, then sit
    # idle (clocks and power state dropping) at the barrier before the timed region until the slowest rank
    # arrives - measured at 2 GPUs: that rank's first 20-step window ran 11 % slower than its later ones
    import gc
    gc.collect()
    gc.disable()  # no collection inside the timed windows (a 20-step window is 0.8 ms); collected HERE, before the
    #               settle phase, because a collection right before a window idles the GPU for tens of milliseconds
    barrier()
    # untimed settle phase: touch the whole resident pool once (first-touch page faults, TLB fill, L2
    # state, clocks back up after the idle wait above), i.e. >= 200 forwards, so that a short --steps
    # window reads the steady state; the --warmup steps follow and lead straight into the timed region
    # (multi-GPU: through the exchange path, so that every ring slot and peer mapping has been written once)
    n_settle = max(n_pool, 200)
    for i in range(n_settle):
        step(i, last=i == n_settle - 1)
    drain()
    torch.cuda.synchronize(device)
    for i in range(args.warmup):
        step(i, last=i == args.warmup - 1)
    drain()
    impl_used = model.last_impl
    launches0 = model.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step(i, last=i == args.steps - 1)
    drain()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = model.launch_count - launches0
    if args.diag:  # diagnostic only (stderr): the same window a few more times, per rank, with host enqueue time
        for rep in range(args.diag):
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            h0 = time.perf_counter()
            d0.record()
            for i in range(args.steps):
                step(i, last=i == args.steps - 1)
            drain()
            d1.record()
            h1 = time.perf_counter()
            barrier()
            print(f"[diag] rank {rank} window {rep}: device {d0.elapsed_time(d1) / args.steps * 1e3:.2f} us/step, "
                  f"host enqueue {(h1 - h0) / args.steps * 1e6:.2f} us/step (first window: "
                  f"{ms / args.steps * 1e3:.2f})", file=sys.stderr, flush=True)

    # ---- untimed check of the exchange step (N > 1): the gathered ring slot holds, for every rank in
    # rank order and every step of the group in step order, exactly the logits a single GPU computes for
    # the same chunks (bitwise: the kernels are batch-invariant).  Every rank ships the compact inputs of
    # its last group to everybody (bytes over NCCL), recomputes all of them locally and compares.
    gather_verified = None
    if distributed:
        if peer_ring is not None:
            peer_ring.barrier()                       # every rank's kernels (and their peer stores) are done
        first = ((args.steps - 1) // G) * G           # first step of the last (possibly partial) group
        n_in_group = args.steps - first
        slot = (first // G) & 1
        ok = True
        for k in range(n_in_group):
            parts = []
            for t in dev_batch(first + k):
                mine = t.contiguous().view(torch.uint8).reshape(-1)
                everyone = torch.empty(world * mine.numel(), dtype=torch.uint8, device=device)
                dist.all_gather_into_tensor(everyone, mine)
                parts.append(everyone.view(world, -1))
            for r in range(world):
                shapes = [(BATCH, 1, CHUNK_LEN), dev_pool["sequence"][:BATCH].shape,
                          dev_pool["sequence_to_signal_mapping"][:BATCH].shape, (BATCH,)]
                dts = [torch.float32, torch.int8, torch.int16, torch.int16]
                args_r = [p[r].view(dt).reshape(tuple(sh)) for p, dt, sh in zip(parts, dts, shapes)]
                want = model.forward_compact(*args_r)
                got = (peer_ring.block(slot)[r, k] if peer_ring is not None
                       else gath_ring[slot].view(world, G, BATCH, -1)[r, k])
                ok = ok and bool(torch.equal(got, want))
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gather_verified = bool(flag.item())
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * BATCH * args.steps / (ms_max * 1e-3)

    # ---- e2e: HOST buffers through the C-ABI host call (rb200_infer_host_async): every step copies its
    # compact inputs from pinned host memory to the device, runs the kernels and copies the logits back
    # to pinned host memory, all inside the timed region.  Four streams rotate so later steps' copies
    # overlap earlier steps' kernels - the same pipelining the reference's queue pipeline does
    # (src/remora/inference.py:488-572); a slot's result is read on the host before the slot is reused.
    n_host = min(n_pool, 16)
    host_batches = []
    for i in range(n_host):
        d = slice_batch(pool, i)
        # one pinned block per batch (the four arrays on 256-byte boundaries): one host-to-device copy per step
        host_batches.append(model.pinned_batch(d["signal"], d["sequence"], d["sequence_to_signal_mapping"],
                                               d["sequence_lengths"]))
    NSLOT = 4  # steps in flight: copies of later steps overlap the kernels of earlier ones
    e2e_streams = [torch.cuda.Stream(device) for _ in range(NSLOT)]
    e2e_out = [torch.empty((BATCH, model.num_out), dtype=torch.float32).pin_memory() for _ in range(NSLOT)]
    checksum = 0.0

    def e2e_run(n_steps):
        nonlocal checksum
        for i in range(n_steps):
            slot = i % NSLOT
            if i >= NSLOT:
                e2e_streams[slot].synchronize()          # step i-NSLOT finished: its logits are on the host
                checksum += float(e2e_out[slot][0, 0])   # host-side read of the step's result
            model.infer_host_async(*host_batches[i % n_host], e2e_out[slot], stream=e2e_streams[slot])
        for slot in range(NSLOT):
            e2e_streams[slot].synchronize()
            checksum += float(e2e_out[slot][0, 0])

    e2e_run(args.warmup)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = args.steps
    e2e_run(e2e_steps)
    torch.cuda.synchronize(device)
    e2e_s = time.perf_counter() - t0
    gc.enable()
    t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * BATCH * e2e_steps / float(t.item())
    # clock samples cover both timed regions (device-resident loop and the end-to-end loop)
    clocks = sampler.stop() if rank == 0 else None
    # bytes the single copy of a packed batch moves (the four arrays + the padding between them)
    hb0 = host_batches[0]
    h2d = int(hb0[3].data_ptr() - hb0[0].data_ptr() + hb0[3].numel() * hb0[3].element_size())
    d2h = BATCH * model.num_out * 4
    # the synchronous single-call form (pageable host buffers, internal pinned staging, blocking)
    sync_steps = max(10, args.steps // 4)
    d0 = slice_batch(pool, 0)
    for _ in range(3):
        model.infer_host(d0["signal"], d0["sequence"], d0["sequence_to_signal_mapping"], d0["sequence_lengths"])
    t0 = time.perf_counter()
    for i in range(sync_steps):
        d = slice_batch(pool, i % n_host)
        model.infer_host(d["signal"], d["sequence"], d["sequence_to_signal_mapping"], d["sequence_lengths"])
    e2e_sync_value = BATCH * sync_steps / (time.perf_counter() - t0)

    # ---- per-kernel device times (separate pass: event records would perturb the headline) --------
    prof = None
    single_kernel = impl_used in ("fused_mega", "fused_bf16")
    if impl_used in ("fused", "fused_tc") or single_kernel:
        model.set_profile(True)
        for i in range(min(args.steps, 500)):
            model.forward_compact(*dev_batch(i))
        torch.cuda.synchronize(device)
        ms3, n_fw = model.get_profile()
        model.set_profile(False)
        if single_kernel:  # events around every launch: launches do not overlap in this pass
            prof = {"mega_kernel_isolated_ms": ms3[0] / n_fw, "forwards": n_fw}
        else:
            prof = {"k1_front_ms": ms3[0] / n_fw, "k2_merge_xproj_ms": ms3[1] / n_fw,
                    "k3_lstm_ms": ms3[2] / n_fw, "forwards": n_fw}

    # ---- the other variants of the same step, device-resident, same pool (N = 1 only) ---------------------
    variants = {}
    if world == 1 and single_kernel:
        for name in ("fused_bf16", "fused_tc"):
            try:
                model.set_impl(name)
                for i in range(20):
                    model.forward_compact(*dev_batch(i))
                torch.cuda.synchronize(device)
                v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n_v = min(args.steps, 500)
                v0.record()
                for i in range(n_v):
                    model.forward_compact(*dev_batch(i))
                v1.record()
                torch.cuda.synchronize(device)
                v_ms = v0.elapsed_time(v1) / n_v
                variants[name] = {"value": BATCH / (v_ms * 1e-3), "unit": "chunks/s", "ms_per_step": v_ms,
                                  "impl": model.last_impl, "steps": n_v}
            except Exception as e:  # noqa: BLE001
                variants[name] = {"error": str(e)[:200]}
        model.set_impl("auto")
        if "value" in variants.get("fused_bf16", {}):
            variants["fused_bf16"].update({
                "dtype": "bf16", "what": "BASELINE configs[1] as written: single-pass bf16 tensor-core operands for "
                "the four GEMM-shaped layers, fp32 accumulate / gates / classifier",
                "tolerance": "tests/test_gpu_parity.py::test_bf16_variant_tolerance: max-abs logit error < 0.25 "
                             "on logits spanning +-4, max-abs probability error < 0.05"})
        if "value" in variants.get("fused_tc", {}):
            variants["fused_tc"]["what"] = "round-1 path: three kernels, 3xTF32 on tcgen05"

    # ---- dense encode kernel alone (the HBM-bound kernel of the path) ------------------------------
    from remora_b200 import encoded_kmers
    enc_n = min(32768, n_pool * BATCH)  # 472 MB of one-hot output per launch
    sl = slice(0, enc_n)
    enc_args = (dev_pool["sequence"][sl], dev_pool["sequence_to_signal_mapping"][sl],
                dev_pool["sequence_lengths"][sl])
    for _ in range(3):
        encoded_kmers.compute_encoded_kmer_batch_torch(4, 4, *enc_args, sig_len=CHUNK_LEN, device=device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)
    enc_ms = []
    for _ in range(5):
        flush.fill_(1.0)  # write > L2 between timed launches
        e0.record()
        encoded_kmers.compute_encoded_kmer_batch_torch(4, 4, *enc_args, sig_len=CHUNK_LEN, device=device)
        e1.record()
        torch.cuda.synchronize(device)
        enc_ms.append(e0.elapsed_time(e1))
    del flush
    enc_ms = float(np.median(enc_ms))

    if rank == 0:
        peak_gbs, peak_src, peaks = measured_peaks()
        line = {
            "metric": METRIC, "value": value, "unit": "chunks/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "batch_per_gpu": BATCH, "global_batch": world * BATCH,
                       "chunk_len": CHUNK_LEN, "kmer_context": list(KMER_CONTEXT),
                       "arm": "fp32-parity arithmetic: one sm_100a kernel per batch; the GEMM-shaped layers on tcgen05 "
                              "with fp16 hi/lo split operands (3 products, 22 significant bits) and fp32 TMEM "
                              "accumulators, the LSTM recurrence and the small GEMMs on mma.sync (4 products), the "
                              "rest fp32 FFMA; logits within 1e-4 of the reference CPU forward (tests); the bf16 variant "
                              "the config names is timed in `variants`",
                       "parallelism": ("single GPU" if world == 1 else
                                       f"batch-shard x{world}; exchange fused into the kernel: "
                                       + ("one extra thread block of every launch ships the previous step's logits "
                                          if peer_ring.deferred else "the classifier epilogue stores every step's logits ")
                                       + f"into every rank's ring over NVLink "
                                       f"(symmetric memory{', NVLS multicast' if peer_ring.multicast_ptr else ''}), "
                                       f"no collective launch" if peer_ring is not None else
                                       f"batch-shard x{world}, logits all-gathered (NCCL) in groups of "
                                       f"{G} steps, asynchronously"),
                       "gather": gather_mode,
                       "gather_deferred": bool(peer_ring.deferred) if peer_ring is not None else None,
                       "impl": impl_used,
                       "l2_policy": f"inputs larger than L2: each step reads a different batch of a "
                                    f"{pool_bytes / 1e6:.0f} MB resident pool ({n_pool} batches)"},
            "e2e": {"value": e2e_value, "unit": "chunks/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "path": "B200Model.infer_host_async -> rb200_infer_host_async, pinned host buffers (the four "
                            "arrays of a batch in one pinned block: one H2D copy per step), 4 "
                            "streams in rotation, the kernel stores the logits straight into the pinned result buffer (a slot's result is read on the host before the slot is "
                            "reused); every step's H2D and D2H are in the timed region",
                    "blocking_single_call_value": e2e_sync_value,
                    "blocking_single_call_path": "B200Model.infer_host -> rb200_infer_host (pageable "
                                                 "buffers, one call = copy in + kernels + copy out + sync), "
                                                 "this rank only"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "settle": f"{max(n_pool, 200)} untimed forwards over the whole pool before the {args.warmup} warm-up steps",
        }
        if gather_verified is not None:
            line["gather_verified"] = gather_verified
        if variants:
            line["variants"] = variants
        if prof is not None and single_kernel:
            sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
            step_s = ms_max / args.steps * 1e-3
            fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
            h_peak = float(peaks.get("bf16_tflops", 1590.0))  # measured dense bf16 GEMM; fp16 runs at the same rate
            n_prod = 3 if impl_used == "fused_mega" else 1
            kname = "mega_kernel<0>" if impl_used == "fused_mega" else "mega_kernel<1>"
            # per chunk, T = 100: dense MACs of the layers on tcgen05 (seq_conv2, sig_conv3,
            # merge_conv1, LSTM1 input projection) ...
            mac_tensor = 258048 + 372736 + 983040 + 393216
            # ... of the layers on the warp-level tensor-core path (mma.sync): LSTM1 recurrence (24 steps x 256 x
            # 64) and the single LSTM2 step ...
            mac_hmma = 393216 + 16384
            # ... and of the FFMA2 / FADD2 work: sig_conv1/2, classifier, gather-add of seq_conv1 (adds as MACs)
            mac_fma = 1920 + 29440 + 128 + 69120
            mac_dense = int(DENSE_MFLOP * 1e6 / 2)
            n_prod_hmma = 4  # (W_hi + W_lo) . (h_hi + h_lo): all four products, fp16 operands, fp32 accumulate
            hmma_peak = 148 * 955.0 * 2 * sm_mhz * 1e6 / 1e12  # measured: 8.6 cycles per HMMA.16816 and SM sub-partition
            traffic = ncu_traffic_bytes("mega_kernel")
            ncu_pipes = ncu_pipe_counters("mega_kernel")
            ach = BATCH * 2.0 * (mac_tensor + mac_hmma) / step_s / 1e12
            ach_t = BATCH * 2.0 * mac_tensor / step_s / 1e12
            ach_h = BATCH * 2.0 * mac_hmma / step_s / 1e12
            line["roofline"] = {
                "kernel": kname, "bound": "tensor", "achieved": ach, "peak": h_peak, "unit": "TFLOP/s",
                "frac": ach / h_peak, "executed_frac": (n_prod * ach_t + n_prod_hmma * ach_h) / h_peak,
                "traffic": (traffic or (None, None))[0], "traffic_source": (traffic or (None, None))[1],
                "peak_source": peak_src,
                "duration_ms": step_s * 1e3, "isolated_launch_ms": prof["mega_kernel_isolated_ms"],
                "note": f"one kernel per step; achieved = the dense MACs of the layers that run on the tensor cores "
                        f"(tcgen05: 2 007 040 per chunk, mma.sync: 409 600 per chunk, each counted ONCE) x 2 x {BATCH} / "
                        f"the average launch duration in the timed region (= ms_per_step: consecutive launches "
                        f"overlap, two CTAs per SM); the kernel executes {n_prod} (tcgen05) / {n_prod_hmma} (mma.sync) "
                        f"fp16 products per MAC for fp32 parity, executed_frac counts them.  Neither the tensor "
                        f"pipe nor HBM is the binding resource: the step is bound by each CTA's dependency chain "
                        f"(24-step LSTM, CUDA-core phases between the MMA phases) times the two chains an SM holds "
                        f"(roofline_compute, phase stamps in profiles/)"}
            line["roofline_compute"] = {
                "tensor": {"executed_mac_per_chunk": n_prod * mac_tensor, "dense_mac_per_chunk": mac_tensor,
                           "executed_tflops": n_prod * ach_t, "peak_tflops": h_peak,
                           "frac_executed": n_prod * ach_t / h_peak, "frac_dense": ach_t / h_peak,
                           "ncu_pipe_tensor_pct": (ncu_pipes or {}).get("tensor")},
                "mma_sync": {"executed_mac_per_chunk": n_prod_hmma * mac_hmma, "dense_mac_per_chunk": mac_hmma,
                             "executed_tflops": n_prod_hmma * ach_h, "peak_tflops": hmma_peak,
                             "frac_executed": n_prod_hmma * ach_h / hmma_peak,
                             "note": "LSTM recurrence + LSTM2 step on HMMA.16816 from registers"},
                "ffma2": {"executed_mac_per_chunk": mac_fma,
                          "executed_tflops": BATCH * 2.0 * mac_fma / step_s / 1e12, "peak_tflops": fp32_peak,
                          "frac_executed": BATCH * 2.0 * mac_fma / step_s / 1e12 / fp32_peak,
                          "ncu_pipe_fma_pct": (ncu_pipes or {}).get("fma")},
                "reference_dense": {"mac_per_chunk": mac_dense,
                                    "tflops": BATCH * DENSE_MFLOP * 1e6 / step_s / 1e12,
                                    "note": "what the reference executes (full LSTM2, dense one-hot conv); "
                                            "the kernel skips 23/24 of LSTM2 and gathers seq_conv1 (SURVEY a3.4, a3.9)"},
                "peak_source": f"tensor: MEASURED_PEAKS bf16_tflops (burst); mma_sync: 148 SM x 955 MAC/clk (measured, "
                               f"scripts/microbench/hmma_rate.cu) x {sm_mhz:.0f} MHz; ffma2: 148 SM x 128 lanes x 2 flop x "
                               f"{sm_mhz:.0f} MHz (sampled clock)",
                "ncu_source": (ncu_pipes or {}).get("source")}
            line["roofline_step"] = {
                "iface_a_bytes_per_chunk": BYTES_IFACE_A,
                "hbm_frac_iface_a": world * BATCH * BYTES_IFACE_A / step_s / 1e9 / (peak_gbs * world),
                "iface_b_bytes_per_chunk": 488,
                "hbm_frac_iface_b": world * BATCH * 488 / step_s / 1e9 / (peak_gbs * world),
                "dense_mflop_per_chunk": DENSE_MFLOP,
                "fp32_frac_dense": BATCH * DENSE_MFLOP * 1e6 / step_s / 1e12 / fp32_peak,
            }
            line["kernels_ms"] = prof
        elif prof is not None:
            sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
            fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12  # FFMA2 issue peak at the sampled clock
            tf32_peak = float(peaks.get("bf16_tflops", 1590.0)) / 2.0  # dense tf32 = half the measured bf16 GEMM
            tc = impl_used == "fused_tc"
            # algorithmic bytes / MACs per chunk of each fused kernel
            cat_b = 28 * 128 * 4  # cat activations, fp32 (the tensor-core path derives the TF32 lo part on chip)
            kernels = {
                "k1tc_front_kernel" if tc else "k1_front_kernel": {"ms": prof["k1_front_ms"], "bytes": 480 + cat_b,
                                    "mac": 1920 + 29440 + 258048 + 276480 + 372736, "roof": "fp32"},
                "k2tc_kernel" if tc else "k2_merge_kernel": {
                    "ms": prof["k2_merge_xproj_ms"], "bytes": cat_b + 24 * 256 * 4,
                    "mac": 983040 + 393216, "roof": "tensor" if tc else "fp32"},
                "k3_lstm_kernel": {"ms": prof["k3_lstm_ms"], "bytes": 24 * 256 * 4 + 8,
                                   "mac": 393216 + 786432 + 128, "roof": "fp32"},
            }
            name, dom = max(kernels.items(), key=lambda kv: kv[1]["ms"])
            dom_s = dom["ms"] * 1e-3
            line["roofline"] = {
                "kernel": name, "bound": "hbm", "achieved": BATCH * dom["bytes"] / dom_s / 1e9,
                "peak": peak_gbs, "unit": "GB/s", "frac": BATCH * dom["bytes"] / dom_s / 1e9 / peak_gbs,
                "traffic": (ncu_traffic_bytes(name) or (None, None))[0],
                "traffic_source": (ncu_traffic_bytes(name) or (None, None))[1], "peak_source": peak_src,
                "note": "dominant kernel by measured time; it is compute bound, not HBM bound (472 FLOP/B "
                        "at the reference interface; intermediates are L2 resident): the binding roofs are "
                        "in roofline_compute",
            }
            line["roofline_compute"] = {
                k: {"ms": v["ms"], "bound": "tensor (tf32 dense, 3xTF32 split counted once)"
                    if v["roof"] == "tensor" else "fp32 FFMA2 issue",
                    "achieved": BATCH * 2.0 * v["mac"] / (v["ms"] * 1e-3) / 1e12,
                    "peak": tf32_peak if v["roof"] == "tensor" else fp32_peak, "unit": "TFLOP/s",
                    "frac": BATCH * 2.0 * v["mac"] / (v["ms"] * 1e-3) / 1e12 /
                    (tf32_peak if v["roof"] == "tensor" else fp32_peak)}
                for k, v in kernels.items()}
            line["roofline_compute"]["peak_source"] = (
                f"fp32: 148 SM x 128 lanes x 2 flop x {sm_mhz:.0f} MHz (sampled clock; FFMA2 micro-benchmark "
                f"reaches 97-99% of it); tf32: MEASURED_PEAKS bf16_tflops / 2; flops are the reference's "
                f"dense MACs (skipped LSTM2 steps and the one-hot conv counted as the reference does them)")
            step_s = ms_max / args.steps * 1e-3
            line["roofline_step"] = {
                "iface_a_bytes_per_chunk": BYTES_IFACE_A,
                "hbm_frac_iface_a": world * BATCH * BYTES_IFACE_A / step_s / 1e9 / (peak_gbs * world),
                "dense_mflop_per_chunk": DENSE_MFLOP,
                "fp32_frac_dense": BATCH * DENSE_MFLOP * 1e6 / step_s / 1e12 / fp32_peak,
            }
            line["kernels_ms"] = prof
        enc_bytes = enc_n * 36 * CHUNK_LEN * 4
        line["roofline_encode"] = {
            "kernel": "encode_dense_tma_kernel", "bound": "hbm", "achieved": enc_bytes / (enc_ms * 1e-3) / 1e9,
            "peak": peak_gbs, "unit": "GB/s", "frac": enc_bytes / (enc_ms * 1e-3) / 1e9 / peak_gbs,
            "chunks": enc_n, "ms": enc_ms, "l2": "flushed between launches",
        }
        if world == 1:
            try:
                line["gpu_stock_baseline"] = gpu_stock_baseline(device, steps=min(args.steps, 200),
                                                                warmup=args.warmup)
                best = max(v["value"] for k, v in line["gpu_stock_baseline"].items()
                           if isinstance(v, dict) and "value" in v)
                line["gpu_stock_baseline"]["ours_over_best_stock"] = value / best
            except Exception as e:  # noqa: BLE001
                line["gpu_stock_baseline"] = {"error": str(e)[:300]}
            try:
                line["configs"] = config_legs(device, model)
                stock = line["gpu_stock_baseline"]
                if "torch_defaults" in stock and "value" in line["configs"].get("dense_interface_b1024_T100", {}):
                    line["configs"]["dense_interface_b1024_T100"]["over_stock_cudnn_same_interface"] = (
                        line["configs"]["dense_interface_b1024_T100"]["value"] /
                        max(stock[k]["value"] for k in ("torch_defaults", "tf32_off", "tf32_on")))
            except Exception as e:  # noqa: BLE001
                line["configs"] = {"error": str(e)[:300]}
            line["cpu_baseline"] = run_cpu_sample()
            try:  # rows of SURVEY.md 8f measured beside the headline (never allowed to break the line)
                sys.path.insert(0, os.path.join(ROOT, "scripts"))
                import refine_times
                line["next_rows"] = {"signal_mapping_refinement": refine_times.refine_bench()}
            except Exception as e:  # noqa: BLE001
                line["next_rows"] = {"signal_mapping_refinement": {"error": str(e)[:200]}}
            try:
                import vbz_times
                v = vbz_times.vbz_bench(n_reads=1024, cpu_seconds=1.5, verbose=False)
                line["next_rows"]["pod5_signal_decode"] = {
                    "kernel": "svb16_decode_kernel", "unit": "samples/s", "value": v["samples_per_s"],
                    "ms": v["kernel_ms"], "samples": v["samples"], "packed_bytes_per_sample": v["bytes_per_sample"],
                    "roofline": {"bound": "hbm", "achieved": v["algorithmic_gb_per_s"], "peak": peak_gbs,
                                 "unit": "GB/s", "frac": v["algorithmic_gb_per_s"] / peak_gbs,
                                 "l2": v["l2"]},
                    "cpu_numpy_1_core_samples_per_s": v["numpy_samples_per_s_1_core"],
                    "host_zstd_samples_per_s": v["host_zstd_samples_per_s"],
                    "parity": "bit-exact (tests/test_io.py)"}
            except Exception as e:  # noqa: BLE001
                line["next_rows"]["pod5_signal_decode"] = {"error": str(e)[:200]}
            try:
                import pipeline_times
                line["next_rows"]["file_pipeline"] = pipeline_times.pipeline_bench()
            except Exception as e:  # noqa: BLE001
                line["next_rows"]["file_pipeline"] = {"error": str(e)[:200]}
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gather-every", type=int, default=8,
                    help="multi-GPU: steps per all-gather of the logits (every step's logits are gathered)")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU exchange of the logits: fused peer stores from the kernel (default) or NCCL")
    ap.add_argument("--multicast", action="store_true", help="p2p gather through the NVLS multicast address")
    ap.add_argument("--immediate", action="store_true",
                    help="p2p gather: every thread block stores to every rank itself (rb200_forward_compact_gather) "
                         "instead of the one-step-behind shipping block (rb200_forward_compact_ship)")
    ap.add_argument("--signal", action="store_true",
                    help="p2p gather: also bump the per-step arrival counters on every rank (system-scope fence "
                         "per thread block); the bench reads the ring only after a barrier and leaves it off")
    ap.add_argument("--diag", type=int, default=0, help="repeat the timed window this many times, report to stderr")
    ap.add_argument("--pool-batches", type=int, default=400,
                    help="resident batches of compact inputs (400 x 1024 x 480 B = 197 MB > L2)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
