"""Timing of the reference's own call form model(sigs, enc_kmers) (dense one-hot input)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from remora_b200 import encoded_kmers, model_util
from remora_b200.synth import synth_chunks
for name in ("convlstm_s64_k9_hot", "conv_s64_k9"):
    model, md = model_util.load_model(os.path.join(ROOT, f"tests/golden/{name}.pt"), device=torch.device("cuda:0"), eval_only=True)
    d = synth_chunks(1024, 100, (4, 4), seed=3)
    args = [torch.from_numpy(d[k]).cuda() for k in ("signal", "sequence", "sequence_to_signal_mapping", "sequence_lengths")]
    enc = encoded_kmers.compute_encoded_kmer_batch_torch(4, 4, *args[1:], sig_len=100, device=torch.device("cuda:0"))
    for _ in range(3):
        model(args[0], enc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = model(args[0], enc)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: dense model(sigs, enc) {ms*1e3:.0f} us per 1024 chunks -> {1024/ms*1e3/1e6:.3f} M chunks/s [{model.last_impl}]")
