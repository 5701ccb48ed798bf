"""Bring-up check of the Conv_w_ref single kernel: intermediates against the plain layer kernels, logits
against the oracle, CUDA-event timings next to the tiled layer kernels."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import remora_oracle as ro  # noqa: E402
from remora_b200 import model_util  # noqa: E402
from remora_b200.synth import synth_chunks  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "conv_s64_k9"
pt = os.path.join(ROOT, "tests", "golden", name + ".pt")
model, md = model_util.load_model(pt, device=torch.device("cuda:0"), eval_only=True)
sd, _ = model_util._raw_load_torchscript(pt)
TT = md["chunk_len"]


def args_of(d):
    return [torch.from_numpy(d[k]).cuda() for k in ("signal", "sequence", "sequence_to_signal_mapping",
                                                    "sequence_lengths")]


for B in (3, 64, 1024):
    d = synth_chunks(B, TT, (4, 4), seed=200 + B)
    a = args_of(d)
    want = ro.oracle_infer_compact(sd, (4, 4), d["signal"], d["sequence"], d["sequence_to_signal_mapping"],
                                   d["sequence_lengths"])
    model.set_impl("layers")
    model.set_debug(True)
    ref = model.forward_compact(*a).cpu().numpy()
    kept = {n: model.debug_tensor(n).cpu() for n in ("cat", "merge4")}
    model.set_impl("fused_mega")
    model.set_debug(True)
    got = model.forward_compact(*a).cpu().numpy()
    torch.cuda.synchronize()
    line = f"B={B} [{model.last_impl}] logits vs oracle {np.abs(got - want).max():.3e} vs layers {np.abs(got - ref).max():.3e} (|logit| max {np.abs(want).max():.2f})"
    for n in ("cat", "merge4"):
        t = model.debug_tensor(n).cpu()
        line += f" | {n} {float((t - kept[n]).abs().max()):.3e} (scale {float(kept[n].abs().max()):.1f})"
    print(line, "flags", model.get_flags(clear=True), flush=True)
    model.set_debug(False)
    got2 = model.forward_compact(*a).cpu().numpy()
    print(f"     no-debug launch equals debug launch: {np.array_equal(got, got2)}", flush=True)

for B in (1024, 4096):
    pool = [args_of(synth_chunks(B, TT, (4, 4), seed=s)) for s in range(4)]
    for impl in ("tiled", "fused_mega"):
        model.set_impl(impl)
        for i in range(10):
            model.forward_compact(*pool[i % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 100
        e0.record()
        for i in range(n):
            model.forward_compact(*pool[i % 4])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"[{model.last_impl}] B={B}: {ms * 1e3:.1f} us/step = {B / ms / 1e3:.2f} M chunks/s", flush=True)
