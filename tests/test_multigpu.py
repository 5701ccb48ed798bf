"""N > 1 on real hardware (needs >= 2 visible GPUs; `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`):
one process per GPU over NCCL with the CUDA model as compute.

* batch sharding: every rank ends with the gathered [B, num_out] logits in chunk order, BITWISE equal
  to what one GPU computes for the same chunks (the kernels are batch-invariant) - the order contract
  the reference's `unbatch` stage relies on (src/remora/inference.py:331-367);
* file pipeline (BASELINE config 5 shape): `infer_from_pod5_and_bam(rank=, world_size=)` on two GPUs -
  the union of the ranks' calls equals the single-GPU calls, read by read (MM string and ML bytes).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shard_worker(rank, world, port, n_chunks, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from remora_b200 import model_util, parallel
    from remora_b200.synth import synth_chunks
    model, md = model_util.load_model(os.path.join(GOLDEN, "convlstm_s64_k9_hot.pt"), device=dev, eval_only=True)
    d = synth_chunks(n_chunks, md["chunk_len"], md["kmer_context_bases"], seed=11)
    arrays = [torch.from_numpy(d[k]).to(dev) for k in ("signal", "sequence", "sequence_to_signal_mapping",
                                                       "sequence_lengths")]
    caller = parallel.ShardedCaller(model.forward_compact, num_out=model.num_out)
    out = caller.call(*arrays)
    assert out.shape == (n_chunks, model.num_out) and out.is_cuda
    single = model.forward_compact(*arrays)  # the whole batch on this GPU alone
    np.save(os.path.join(out_dir, f"gathered{rank}.npy"), out.cpu().numpy())
    np.save(os.path.join(out_dir, f"single{rank}.npy"), single.cpu().numpy())
    with open(os.path.join(out_dir, f"impl{rank}.txt"), "w") as fh:
        fh.write(model.last_impl)
    dist.barrier(device_ids=[rank])
    dist.destroy_process_group()


@pytest.mark.parametrize("n_chunks", [2048, 1001])  # even split and ragged split
def test_nccl_shard_gather_is_bitwise_single_gpu(tmp_path, n_chunks):
    import torch.multiprocessing as mp
    mp.spawn(_shard_worker, args=(2, _free_port(), n_chunks, str(tmp_path)), nprocs=2, join=True)
    g = [np.load(tmp_path / f"gathered{r}.npy") for r in range(2)]
    s = [np.load(tmp_path / f"single{r}.npy") for r in range(2)]
    assert open(tmp_path / "impl0.txt").read() == "fused_mega"
    assert np.array_equal(g[0], g[1])      # every rank holds the same gathered tensor
    assert np.array_equal(s[0], s[1])      # both GPUs compute the same bits
    assert np.array_equal(g[0], s[0])      # shard + gather == one GPU, bit for bit, in chunk order


def _p2p_worker(rank, world, port, out_dir, deferred):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from remora_b200 import model_util, parallel
    from remora_b200.synth import synth_chunks
    model, md = model_util.load_model(os.path.join(GOLDEN, "convlstm_s64_k9_hot.pt"), device=dev, eval_only=True)
    B, steps = 1000 if not deferred else 1001, 3
    ring = parallel.PeerLogitRing(model, slots=2, steps=steps, batch=B, deferred=deferred)
    mine, wants = [], []
    for k in range(steps):
        d = synth_chunks(B, 100, (4, 4), seed=100 * rank + k)
        arrays = [torch.from_numpy(d[x]).to(dev) for x in ("signal", "sequence", "sequence_to_signal_mapping",
                                                           "sequence_lengths")]
        ring.forward(model, arrays, slot=1, step=k)
        mine.append(arrays)
    ring.flush()
    torch.cuda.synchronize(dev)
    ring.barrier()
    # what a single GPU computes for EVERY rank's batches (inputs regenerated from the seeds)
    ok = True
    for r in range(world):
        for k in range(steps):
            d = synth_chunks(B, 100, (4, 4), seed=100 * r + k)
            arrays = [torch.from_numpy(d[x]).to(dev) for x in ("signal", "sequence", "sequence_to_signal_mapping",
                                                               "sequence_lengths")]
            want = model.forward_compact(*arrays)
            ok = ok and bool(torch.equal(ring.block(1)[r, k], want))
    arrivals = ring.arrivals(1).cpu().numpy()
    per_step = 1 if deferred else (B + 3) // 4   # one increment per shipped block | per thread block
    np.save(os.path.join(out_dir, f"p2p{rank}.npy"), np.array([int(ok), int((arrivals == per_step).all()),
                                                              int(ring.block(0).abs().sum().item() == 0)]))
    dist.barrier(device_ids=[rank])
    dist.destroy_process_group()


@pytest.mark.parametrize("deferred", [False, True])
def test_kernel_fused_peer_gather_is_bitwise_single_gpu(tmp_path, deferred):
    """rb200_forward_compact_gather / rb200_forward_compact_ship: the kernel stores its logits into every
    rank's symmetric-memory ring over NVLink - every thread block itself, or (deferred) one extra thread
    block of the next launch.  Every rank must hold, in rank and step order, exactly the bits a single GPU
    computes; the arrival counters must read one increment per CTA (per shipped block when deferred); the
    untouched slot stays zero."""
    import torch.multiprocessing as mp
    mp.spawn(_p2p_worker, args=(2, _free_port(), str(tmp_path), deferred), nprocs=2, join=True)
    for r in range(2):
        ok, counted, clean = np.load(tmp_path / f"p2p{r}.npy")
        assert ok == 1 and counted == 1 and clean == 1, (r, ok, counted, clean)


def _pipeline_worker(rank, world, port, pod5, bam, out_dir):
    sys.path.insert(0, ROOT)
    import pickle
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from remora_b200 import inference, model_util
    model, md = model_util.load_model(os.path.join(GOLDEN, "convlstm_s64_k9_refine.pt"), device=dev, eval_only=True)
    res = inference.infer_from_pod5_and_bam(pod5, bam, (model, md), reads_per_batch=4, rank=rank, world_size=world,
                                            out_path=os.path.join(out_dir, f"calls.rank{rank}.bam"))
    with open(os.path.join(out_dir, f"res{rank}.pkl"), "wb") as fh:
        pickle.dump([(r["read_id"], r["mm"], bytes(r["ml"]), r["error"]) for r in res], fh)
    if rank == 0:
        whole = inference.infer_from_pod5_and_bam(pod5, bam, (model, md), reads_per_batch=4)
        with open(os.path.join(out_dir, "single.pkl"), "wb") as fh:
            pickle.dump([(r["read_id"], r["mm"], bytes(r["ml"]), r["error"]) for r in whole], fh)
    dist.barrier(device_ids=[rank])
    dist.destroy_process_group()


def test_file_pipeline_on_two_gpus_equals_one_gpu(tmp_path):
    import pickle
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from remora_b200 import io
    from remora_b200.synth import synth_pod5_bam_run
    pod5, bam, truth = synth_pod5_bam_run(str(tmp_path / "run.pod5"), str(tmp_path / "run.bam"), n_reads=14)
    mp.spawn(_pipeline_worker, args=(2, _free_port(), pod5, bam, str(tmp_path)), nprocs=2, join=True)
    parts = [pickle.load(open(tmp_path / f"res{r}.pkl", "rb")) for r in range(2)]
    single = {r[0]: r for r in pickle.load(open(tmp_path / "single.pkl", "rb"))}
    ids = [r[0] for p in parts for r in p]
    assert sorted(ids) == sorted(truth) and len(set(ids)) == len(ids)   # disjoint shards, union = the run
    assert min(len(p) for p in parts) >= 3
    for p in parts:
        for rid, mm, ml, err in p:
            assert err is None and (rid, mm, ml, err) == single[rid]     # identical calls, read by read
    # every rank wrote a real BAM holding exactly its reads
    for r, p in enumerate(parts):
        with io.BamReader(str(tmp_path / f"calls.rank{r}.bam")) as reader:
            names = [rec.query_name for rec in reader]
        assert sorted(names) == sorted(x[0] for x in p)
