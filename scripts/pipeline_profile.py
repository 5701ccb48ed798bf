"""cProfile of inference.infer_from_pod5_and_bam on a synthetic run (where does the host time go)."""
import cProfile
import os
import pstats
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from remora_b200 import inference, model_util  # noqa: E402
from remora_b200.synth import synth_pod5_bam_run  # noqa: E402

dev = torch.device("cuda:0")
with tempfile.TemporaryDirectory() as tmp:
    pod5, bam, _ = synth_pod5_bam_run(os.path.join(tmp, "r.pod5"), os.path.join(tmp, "r.bam"), n_reads=512,
                                      bases=(1000, 3000))
    model, md = model_util.load_model(os.path.join(ROOT, "tests/golden/convlstm_s64_k9_refine.pt"), device=dev,
                                      eval_only=True)
    for kw in (dict(extract_on_device=True), dict(extract_on_device=False)):
        inference.infer_from_pod5_and_bam(pod5, bam, (model, md), num_reads=16, **kw)
        pr = cProfile.Profile()
        pr.enable()
        inference.infer_from_pod5_and_bam(pod5, bam, (model, md), reads_per_batch=256, **kw)
        torch.cuda.synchronize()
        pr.disable()
        print("=====", kw)
        pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
