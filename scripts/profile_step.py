"""Small driver for ncu captures: a few forwards of the BASELINE step (1024 chunks, T=100)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from remora_b200 import encoded_kmers, model_util  # noqa: E402
from remora_b200.synth import synth_chunks  # noqa: E402

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 5
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
model, md = model_util.load_model(os.path.join(ROOT, "tests/golden/convlstm_s64_k9_hot.pt"),
                                  device=torch.device("cuda:0"), eval_only=True)
d = synth_chunks(B, 100, (4, 4), seed=3)
args = [torch.from_numpy(d[k]).cuda() for k in
        ("signal", "sequence", "sequence_to_signal_mapping", "sequence_lengths")]
for _ in range(n_iter):
    out = model.forward_compact(*args)
enc = encoded_kmers.compute_encoded_kmer_batch_torch(4, 4, *args[1:], sig_len=100,
                                                     device=torch.device("cuda:0"))
torch.cuda.synchronize()
print(model.last_impl, float(out.sum()), float(enc.sum()))
