"""Conv_w_ref (BASELINE config 3) and non-fused shapes: plain vs register-tiled layer kernels."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from remora_b200 import model_util
from remora_b200.synth import synth_chunks


import time

ENQ = [0.0]


def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    ENQ[0] = (time.perf_counter() - t0) / n * 1e6
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, T, batches in (("conv_s64_k9", 100, (1024, 4096)), ("conv_s64_k9_T200", 200, (4096,)),
                         ("convlstm_s16_k6_o3", 100, (1024,))):
    model, md = model_util.load_model(os.path.join(ROOT, f"tests/golden/{name}.pt"),
                                      device=torch.device("cuda:0"), eval_only=True)
    for B in batches:
        d = synth_chunks(B, T, tuple(md["kmer_context_bases"]), seed=3)
        args = [torch.from_numpy(d[k]).cuda() for k in ("signal", "sequence", "sequence_to_signal_mapping",
                                                        "sequence_lengths")]
        for impl in ("layers", "tiled"):
            model.set_impl(impl)
            ms = timed(lambda: model.forward_compact(*args))
            print(f"{name} B={B} T={T} [{impl}]: {ms*1e3:.0f} us -> {B/ms*1e3/1e6:.3f} M chunks/s (host enqueue {ENQ[0]:.0f} us)", flush=True)
