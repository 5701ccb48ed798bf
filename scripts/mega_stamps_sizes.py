"""Phase stamps of CTA 0 (RB200_MEGA_STAMPS=1) for single launches of different batch sizes: one CTA on the
whole GPU, one CTA per SM, two CTAs per SM - separates per-SM sharing from chip-wide (L2) contention."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from remora_b200 import model_util  # noqa: E402
from remora_b200.synth import synth_chunks  # noqa: E402

model, md = model_util.load_model(os.path.join(ROOT, "tests/golden/convlstm_s64_k9_hot.pt"),
                                  device=torch.device("cuda:0"), eval_only=True)
for B in [int(x) for x in sys.argv[1:]] or [4, 592, 1184]:
    d = synth_chunks(B, 100, (4, 4), seed=3)
    args = [torch.from_numpy(d[k]).cuda() for k in ("signal", "sequence", "sequence_to_signal_mapping",
                                                    "sequence_lengths")]
    for _ in range(3):
        print(f"B={B}:", file=sys.stderr, end=" ", flush=True)
        model.forward_compact(*args)
        torch.cuda.synchronize()
