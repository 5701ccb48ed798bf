import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from remora_b200 import model_util
from remora_b200.synth import synth_chunks
import remora_oracle as ro
pt = os.path.dirname(os.path.dirname(os.path.abspath(__file__))) + "/tests/golden/convlstm_s64_k9_hot.pt"
model, md = model_util.load_model(pt, device=torch.device("cuda:0"), eval_only=True)
sd, _ = model_util._raw_load_torchscript(pt)
for B, T in [(64, 100), (7, 100), (1024, 100), (33, 200)]:
    d = synth_chunks(B, T, (4, 4), seed=B)
    args = [torch.from_numpy(d[k]) for k in ("signal", "sequence", "sequence_to_signal_mapping", "sequence_lengths")]
    model.set_impl("fused"); a = model.forward_compact(*args).cpu().numpy()
    model.set_impl("fused_tc"); b = model.forward_compact(*args).cpu().numpy(); torch.cuda.synchronize()
    n = min(B, 64)
    want = ro.oracle_infer_compact(sd, (4, 4), d["signal"][:n], d["sequence"][:n], d["sequence_to_signal_mapping"][:n], d["sequence_lengths"][:n])
    print(B, T, model.last_impl, "tc-vs-fused", np.abs(a - b).max(), "tc-vs-oracle", np.abs(b[:n] - want).max(), "fused-vs-oracle", np.abs(a[:n] - want).max(), flush=True)
