"""Steady-state step time of the single-kernel path for the library named by RB200_LIB (A/B runs of kernel
variants built with scripts/build_variant.sh): 3 x 1500 back-to-back launches of the BASELINE batch over a
pool larger than L2, CUDA events; also checks the logits against the oracle once."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from remora_b200 import model_util  # noqa: E402
from remora_b200.synth import synth_chunks  # noqa: E402

pt = os.path.join(ROOT, "tests", "golden", "convlstm_s64_k9_hot.pt")
model, md = model_util.load_model(pt, device=torch.device("cuda:0"), eval_only=True)
B = 1024
NP = 64
d = synth_chunks(NP * B, 100, (4, 4), seed=5)
keys = ("signal", "sequence", "sequence_to_signal_mapping", "sequence_lengths")
dev = {k: torch.from_numpy(d[k]).cuda() for k in keys}
out = torch.empty((B, 2), dtype=torch.float32, device="cuda")


def batch(i):
    sl = slice((i % NP) * B, (i % NP + 1) * B)
    return [dev[k][sl] for k in keys]


if os.environ.get("AB_CHECK", "1") == "1":
    import remora_oracle as ro
    sd, _ = model_util._raw_load_torchscript(pt)
    want = ro.oracle_infer_compact(sd, (4, 4), *[d[k][:256] for k in keys])
    got = model.forward_compact(*[dev[k][:256] for k in keys]).cpu().numpy()
    print(f"max |logit - oracle| over 256 chunks: {np.abs(got - want).max():.3e}", flush=True)
for i in range(300):
    model.forward_compact(*batch(i), out=out)
torch.cuda.synchronize()
res = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(1500):
        model.forward_compact(*batch(i), out=out)
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 1500 * 1e3)
print(f"{os.environ.get('RB200_LIB', 'default lib')}: [{model.last_impl}] us/step {', '.join(f'{r:.2f}' for r in res)} "
      f"-> {B / min(res):.2f} M chunks/s", flush=True)
