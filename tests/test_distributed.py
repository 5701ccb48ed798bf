"""N>1 path on CPU: world_size-2 gloo processes shard a batch, run the (oracle) forward on their
shard and all-gather the logits; the result must equal the single-process result bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from remora_b200 import parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 1024, 1025, 8191):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = parallel.shard_sizes(n, w)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_chunks, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import remora_oracle as ro
    from remora_b200 import model_util
    from remora_b200.synth import synth_chunks
    sd, md = model_util._raw_load_torchscript(os.path.join(ROOT, "tests/golden/convlstm_s16_k6_o3.pt"))
    model_util.add_derived_metadata(md)
    ctx = md["kmer_context_bases"]
    d = synth_chunks(n_chunks, md["chunk_len"], ctx, seed=5)

    def compute(sig, seq, mp_, ln):
        return torch.from_numpy(ro.oracle_infer_compact(sd, ctx, sig.numpy(), seq.numpy(),
                                                        mp_.numpy(), ln.numpy()))

    caller = parallel.ShardedCaller(compute, num_out=3)
    arrays = [torch.from_numpy(d[k]) for k in ("signal", "sequence", "sequence_to_signal_mapping",
                                               "sequence_lengths")]
    out = caller.call(*arrays)
    assert out.shape == (n_chunks, 3)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), out.numpy())
    if rank == 0:
        np.save(os.path.join(out_dir, "single.npy"), compute(*arrays).numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_chunks", [24, 25])  # even split and ragged split
def test_two_rank_gloo_shard_gather(tmp_path, n_chunks):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_chunks, str(tmp_path)), nprocs=2, join=True)
    single = np.load(tmp_path / "single.npy")
    got = [np.load(tmp_path / f"rank{r}.npy") for r in range(2)]
    assert np.array_equal(got[0], got[1])  # every rank ends with the same gathered tensor
    # the CPU oracle picks batch-size dependent oneDNN kernels, so shard-vs-whole is equal only to
    # fp32 rounding; the CUDA kernels are batch-invariant (tests/test_gpu_parity.py checks that)
    assert np.abs(got[0] - single).max() < 2e-6


def test_shard_by_work_is_a_balanced_partition():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        work = rng.integers(100, 50000, size=101).tolist()
        parts = [parallel.shard_by_work(work, world, r) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(101))  # every read exactly once
        loads = [sum(work[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(work)  # greedy longest-first bound
    assert parallel.shard_by_work([], 4, 2) == []


def _read_shard_worker(rank, world, port, bam, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from remora_b200 import io
    idx = io.ReadIndexedBam(bam, req_tags={"mv"})
    ids = idx.read_ids
    mine = [ids[i] for i in parallel.shard_by_work([len(idx.get_first_alignment(i).query_sequence) for i in ids],
                                                   world, rank)]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        np.save(os.path.join(out_dir, "ids.npy"), np.array([len(g) for g in gathered]))
        flat = sorted(i for g in gathered for i in g)
        assert flat == sorted(ids) and len(set(flat)) == len(flat)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_read_sharding_of_a_bam(tmp_path):
    """The file pipeline's multi-GPU form: both ranks index the same BAM and take disjoint read sets
    whose union is the file (no data-path collective; the gather here is only the test's check)."""
    from remora_b200.synth import synth_pod5_bam_run
    _, bam, truth = synth_pod5_bam_run(str(tmp_path / "r.pod5"), str(tmp_path / "r.bam"), n_reads=9,
                                       bases=(50, 200))
    mp.spawn(_read_shard_worker, args=(2, _free_port(), bam, str(tmp_path)), nprocs=2, join=True)
    counts = np.load(tmp_path / "ids.npy")
    assert counts.sum() == 9 and counts.min() >= 3


def test_logit_ring_layout_blocks_are_disjoint_and_aligned():
    """Layout of the peer-store exchange ring (remora_b200.parallel.LogitRingLayout): every (slot, rank, step)
    block is 16-byte aligned (the shipping thread block moves float4s), blocks and arrival counters never
    overlap, and the rank-major order inside a slot is the order an all-gather would produce."""
    from remora_b200.parallel import LogitRingLayout
    for world, slots, steps, batch, num_out in ((2, 2, 8, 1024, 2), (8, 2, 3, 1001, 3), (3, 1, 1, 1, 2)):
        lay = LogitRingLayout(world, slots, steps, batch, num_out)
        assert lay.block_floats % 4 == 0 and lay.block_floats >= batch * num_out
        seen = set()
        last = -1
        for slot in range(slots):
            for rank in range(world):
                for step in range(steps):
                    off = lay.offset(slot, step, rank)
                    assert off % 4 == 0 and off > last        # aligned, strictly increasing in (slot, rank, step)
                    last = off
                    span = range(off, off + batch * num_out)
                    assert span[-1] < lay.n_data and not (seen & {span[0], span[-1]})
                    seen.update((span[0], span[-1]))
        words = {lay.flag_word(s_, k, r) for s_ in range(slots) for k in range(steps) for r in range(world)}
        assert len(words) == lay.n_flags and min(words) == lay.n_data and max(words) == lay.n_data + lay.n_flags - 1
        with pytest.raises(IndexError):
            lay.offset(slots, 0, 0)
        with pytest.raises(IndexError):
            lay.flag_word(0, steps, 0)


def test_peer_ring_ships_one_step_behind_and_flushes():
    """Bookkeeping of PeerLogitRing's deferred form without GPUs: every forward hands the kernel the block of the
    PREVIOUS forward to ship (source view, destination offset, arrival counter), the first one ships nothing,
    flush() ships the last one exactly once - emulated with a stand-in model that performs the 'peer stores'
    into plain tensors."""
    from remora_b200.parallel import LogitRingLayout, PeerLogitRing
    world, rank, slots, steps, batch, num_out = 3, 1, 2, 2, 5, 2
    lay = LogitRingLayout(world, slots, steps, batch, num_out)
    peers = [torch.zeros(lay.n_data + lay.n_flags) for _ in range(world)]   # every rank's ring

    class FakeModel:
        num_out = 2
        calls = []

        def forward_compact_ship(self, arrays, out, peer_bases, n_peers, self_rank, ship_src=None, ship_dst_offset=0,
                                 multicast_ptr=0, flag_word=-1, shape_hint=None):
            self.calls.append((arrays is None, None if ship_src is None else int(ship_dst_offset), int(flag_word)))
            if ship_src is not None:                     # the extra thread block: previous block -> every other rank
                for r in range(n_peers):
                    if r != self_rank:
                        peers[r][ship_dst_offset:ship_dst_offset + ship_src.numel()] = ship_src
                        if flag_word >= 0:
                            peers[r][flag_word] += 1
            if arrays is not None:                       # the compute blocks: this rank's own block only
                out.copy_(arrays[0].reshape(-1)[:out.numel()])
            else:
                assert shape_hint == (7, 11, 4)

    ring = object.__new__(PeerLogitRing)                 # the constructor needs symmetric memory: fill in by hand
    model = FakeModel()
    ring.world, ring.rank, ring.slots, ring.steps, ring.batch, ring.num_out = world, rank, slots, steps, batch, num_out
    ring.layout, ring.block_floats, ring.n_data, ring.n_flags = lay, lay.block_floats, lay.n_data, lay.n_flags
    ring.buf, ring.peers_dev, ring.multicast_ptr, ring.deferred = peers[rank], 0, 0, True
    ring._pending, ring._shape_hint, ring._model = None, None, model
    sent = []
    for i in range(slots * steps):
        slot, step = i // steps, i % steps
        sig = torch.full((batch, 1, 7), float(i + 1))
        arrays = (sig, torch.zeros((batch, 11), dtype=torch.int8), torch.zeros((batch, 4), dtype=torch.int16),
                  torch.zeros(batch, dtype=torch.int16))
        ring.forward(model, arrays, slot, step, signal=True)
        sent.append((slot, step, float(i + 1)))
    assert model.calls[0][1] is None                                    # nothing to ship with the first forward
    assert [c[1] for c in model.calls[1:]] == [lay.offset(s_, k, rank) for s_, k, _ in sent[:-1]]   # one step behind
    last = lay.offset(*sent[-1][:2], rank)
    assert not torch.equal(peers[0][last:last + batch * num_out], peers[rank][last:last + batch * num_out])
    ring.flush()
    ring.flush()                                                        # idempotent: nothing pending any more
    assert model.calls[-1][0] is True and sum(c[0] for c in model.calls) == 1
    for r in range(world):                                              # every rank holds every block of this rank
        for s_, k, val in sent:
            off = lay.offset(s_, k, rank)
            assert torch.all(peers[r][off:off + batch * num_out] == val)
            if r != rank:
                assert peers[r][lay.flag_word(s_, k, rank)] == 1       # one arrival per shipped block
