"""Per-kernel CUDA-event times of the fused path for a few (B, T) shapes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from remora_b200 import model_util  # noqa: E402
from remora_b200.synth import synth_chunks  # noqa: E402

model, md = model_util.load_model(os.path.join(ROOT, "tests/golden/convlstm_s64_k9_hot.pt"),
                                  device=torch.device("cuda:0"), eval_only=True)
shapes = [(1024, 100), (1024, 196), (1024, 52), (512, 100), (2048, 100), (4096, 100), (8192, 100)]
impl = os.environ.get("RB200_IMPL", "auto")
model.set_impl(impl)
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
for B, T in shapes:
    d = synth_chunks(B, T, (4, 4), seed=3)
    args = [torch.from_numpy(d[k]).cuda() for k in
            ("signal", "sequence", "sequence_to_signal_mapping", "sequence_lengths")]
    for _ in range(5):
        model.forward_compact(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        model.forward_compact(*args)
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / n
    model.set_profile(True)
    for _ in range(n):
        model.forward_compact(*args)
    torch.cuda.synchronize()
    ms, nf = model.get_profile()
    model.set_profile(False)
    print(f"[{model.last_impl}] B={B} T={T} step={total*1e3:.1f}us K1={ms[0]/nf*1e3:.1f} K2={ms[1]/nf*1e3:.1f} "
          f"K3={ms[2]/nf*1e3:.1f}us  -> {B/total*1e3/1e6:.2f} M chunks/s", flush=True)
