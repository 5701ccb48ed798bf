"""Per-CTA trace of the single-kernel path in steady state (profiling aid, RB200_MEGA_TRACE): how long a
CTA lives, how many SM slots are busy, how far apart in phase the two CTAs of an SM are, and the host's
launch rate (tiny batches: the GPU is never the limit).

    RB200_MEGA_TRACE=gpurun_out/mega_trace.txt python scripts/mega_trace.py
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from remora_b200 import model_util  # noqa: E402
from remora_b200.synth import synth_chunks  # noqa: E402

pt = os.path.join(ROOT, "tests", "golden", "convlstm_s64_k9_hot.pt")
model, md = model_util.load_model(pt, device=torch.device("cuda:0"), eval_only=True)
B = int(os.environ.get("TRACE_BATCH", "1024"))
pool = [[torch.from_numpy(d[k]).cuda() for k in ("signal", "sequence", "sequence_to_signal_mapping",
                                                 "sequence_lengths")]
        for d in (synth_chunks(B, 100, (4, 4), seed=s) for s in range(8))]
out = torch.empty((B, 2), dtype=torch.float32, device="cuda")
n = 600
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n):
    model.forward_compact(*pool[i % 8], out=out)
e1.record()
torch.cuda.synchronize()
print(f"B={B}: {e0.elapsed_time(e1) / n * 1e3:.1f} us/step back to back [{model.last_impl}]", flush=True)

# host launch rate: batches of 4 chunks (one CTA)
tiny = [[t[:4] for t in a] for a in pool]
out4 = torch.empty((4, 2), dtype=torch.float32, device="cuda")
for i in range(50):
    model.forward_compact(*tiny[i % 8], out=out4)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(2000):
    model.forward_compact(*tiny[i % 8], out=out4)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host: {(t1 - t0) / 2000 * 1e6:.1f} us per forward_compact call (enqueue only), "
      f"{(t2 - t0) / 2000 * 1e6:.1f} us incl. drain", flush=True)

path = os.environ.get("RB200_MEGA_TRACE")
if path and os.path.isfile(path):
    a = np.loadtxt(path, dtype=np.int64)
    a = a[a[:, 4] > 0]
    launch, cta, smid, t_start, t_end, cyc = a.T[:6]
    if a.shape[1] >= 20:
        names = ["prologue", "stage", "gather", "seq2||sig12", "sig3", "E1", "merge", "E2", "xproj", "E3", "whh",
                 "recurrence", "lstm2", "pdl_wait"]
        ph = a[:, 6:20]
        print("steady-state phase lengths (cycles, mean over all CTAs): " +
              ", ".join(f"{n} {ph[:, i].mean():.0f}" for i, n in enumerate(names)) + f"; sum {ph.sum(axis=1).mean():.0f}", flush=True)
    dur = (t_end - t_start) / 1e3
    print(f"trace: {len(a)} CTAs of {len(np.unique(launch))} launches; CTA life us: mean {dur.mean():.1f} "
          f"p10 {np.percentile(dur, 10):.1f} p50 {np.percentile(dur, 50):.1f} p90 {np.percentile(dur, 90):.1f} "
          f"max {dur.max():.1f}; cycles mean {cyc.mean():.0f}", flush=True)
    span = (t_end.max() - t_start.min()) / 1e3
    busy = dur.sum()
    n_sm = len(np.unique(smid))
    print(f"span {span:.1f} us over {n_sm} SMs: {span / len(np.unique(launch)):.2f} us per launch; slot occupancy "
          f"{busy / (span * n_sm * 2) * 100:.1f} % of {n_sm * 2} slots", flush=True)
    # per launch: first start, last start, last end
    for l in np.unique(launch)[:6]:
        m = launch == l
        print(f"  launch {l}: starts {t_start[m].min() - t_start.min():>8} .. {t_start[m].max() - t_start.min():>8} ns, "
              f"last end {t_end[m].max() - t_start.min():>8} ns, SMs {len(np.unique(smid[m]))}")
    # phase offset between the two CTAs that share an SM: for each CTA start, the time since the start of the
    # CTA that is still resident on the same SM
    offs = []
    for s in np.unique(smid):
        m = smid == s
        st, en = t_start[m], t_end[m]
        order = np.argsort(st)
        st, en = st[order], en[order]
        for i in range(1, len(st)):
            live = (st[:i] <= st[i]) & (en[:i] > st[i])
            if live.any():
                offs.append((st[i] - st[:i][live].max()) / 1e3)
    offs = np.array(offs)
    if len(offs):
        hist, edges = np.histogram(offs, bins=[0, 2, 5, 10, 15, 20, 25, 30, 40, 60, 1e9])
        print("phase offset to the resident partner (us): " +
              ", ".join(f"{edges[i]:.0f}-{edges[i + 1]:.0f}: {hist[i]}" for i in range(len(hist))), flush=True)
        # CTA life against its offset
        print("mean CTA life by partner offset:")
        k = 0
        lifes = []
        for s in np.unique(smid):
            m = smid == s
            st, en = t_start[m], t_end[m]
            order = np.argsort(st)
            st, en = st[order], en[order]
            for i in range(1, len(st)):
                live = (st[:i] <= st[i]) & (en[:i] > st[i])
                if live.any():
                    lifes.append(((st[i] - st[:i][live].max()) / 1e3, (en[i] - st[i]) / 1e3))
        lifes = np.array(lifes)
        for lo, hi in ((0, 5), (5, 15), (15, 25), (25, 40), (40, 1e9)):
            m = (lifes[:, 0] >= lo) & (lifes[:, 0] < hi)
            if m.any():
                print(f"  offset {lo}-{hi} us: n={m.sum()} life {lifes[m, 1].mean():.1f} us")
