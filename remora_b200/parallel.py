"""Multi-GPU batch sharding: one process per GPU, chunks are independent, so a batch
is split into contiguous shards, every rank runs the whole network on its shard with replicated
weights, and the only exchange the path has is the gather of float32 logits (8 B/chunk for two
classes) back to the consumer rank(s).  ``torch.distributed`` over NCCL (NVLink/NVSwitch) on the
GPU box; the same code runs over gloo on CPU for the host-logic tests.

The reference has no multi-device mode at all (one ``--device`` int, src/remora/parsers.py:1373-1377);
``ShardedCaller.call`` produces, on every rank, exactly the [B, num_out] tensor a single-device
``model(sigs, enc_kmers)`` would have produced, in the original chunk order, so that the reference's
order-sensitive ``unbatch`` stage (src/remora/inference.py:331-367) can consume it unchanged.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, world_size, rank):
    """Contiguous, balanced split: the first ``n_items % world_size`` ranks get one extra item."""
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_items, world_size):
    return [shard_bounds(n_items, world_size, r)[1] - shard_bounds(n_items, world_size, r)[0]
            for r in range(world_size)]


class ShardedCaller:
    """Runs ``compute(*shard_arrays) -> [n_shard, num_out]`` on this rank's shard of a global batch
    and returns the all-gathered [B, num_out] logits in chunk order on every rank."""

    def __init__(self, compute, num_out, group=None):
        self.compute = compute
        self.num_out = num_out
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def my_slice(self, n_items):
        return slice(*shard_bounds(n_items, self.world, self.rank))

    def call(self, *global_arrays, async_op=False):
        """``global_arrays``: per-chunk tensors with the full batch on dim 0 (each rank holds, or can
        slice, its own part; nothing but its shard is touched)."""
        n = global_arrays[0].shape[0]
        sl = self.my_slice(n)
        local = self.compute(*[a[sl] for a in global_arrays])
        return self.gather(local, n, async_op=async_op)

    def gather(self, local_logits, n_items, async_op=False):
        if self.world == 1:
            return local_logits
        sizes = shard_sizes(n_items, self.world)
        if len(set(sizes)) == 1:  # even split: one all_gather into a contiguous tensor
            out = torch.empty((n_items, self.num_out), dtype=local_logits.dtype,
                              device=local_logits.device)
            work = dist.all_gather_into_tensor(out, local_logits.contiguous(), group=self.group,
                                               async_op=async_op)
            return (out, work) if async_op else out
        # ragged last shards: pad to the largest shard, gather, trim
        pad = max(sizes)
        buf = torch.zeros((pad, self.num_out), dtype=local_logits.dtype, device=local_logits.device)
        buf[: local_logits.shape[0]] = local_logits
        parts = [torch.empty_like(buf) for _ in range(self.world)]
        dist.all_gather(parts, buf, group=self.group)
        out = torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
        return (out, None) if async_op else out


def shard_by_work(work, world_size, rank):
    """Indices of the items this rank handles when items carry unequal work (reads of different lengths:
    the refinement DP and the chunk count both grow with the number of bases).  Longest-first greedy
    assignment to the least loaded rank; deterministic, every rank computes the same partition, no
    communication.  Reads are independent, so the file pipeline needs no collective at
    all: each rank writes the calls of its own reads."""
    work = [int(w) for w in work]
    order = sorted(range(len(work)), key=lambda i: (-work[i], i))
    load = [0] * world_size
    mine = []
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        load[r] += work[i]
        if r == rank:
            mine.append(i)
    return sorted(mine)


class LogitRingLayout:
    """Where every block of the exchange ring lives (pure arithmetic, shared by ``PeerLogitRing`` and its CPU
    tests): float32 ``[slots][world][steps][block_floats]`` then one uint32 arrival counter per
    (slot, step, source rank).  ``block_floats`` = batch x num_out rounded up to a multiple of 4, so that every
    block starts on a 16-byte boundary (the shipping thread block moves float4s)."""

    def __init__(self, world, slots, steps, batch, num_out):
        self.world, self.slots, self.steps, self.batch, self.num_out = world, slots, steps, batch, num_out
        self.block_floats = (batch * num_out + 3) // 4 * 4
        self.n_data = slots * world * steps * self.block_floats
        self.n_flags = slots * steps * world

    def offset(self, slot, step, rank):
        if not (0 <= slot < self.slots and 0 <= step < self.steps and 0 <= rank < self.world):
            raise IndexError(f"block (slot {slot}, step {step}, rank {rank}) outside the ring")
        return ((slot * self.world + rank) * self.steps + step) * self.block_floats

    def flag_word(self, slot, step, rank):
        if not (0 <= slot < self.slots and 0 <= step < self.steps and 0 <= rank < self.world):
            raise IndexError(f"counter (slot {slot}, step {step}, rank {rank}) outside the ring")
        return self.n_data + (slot * self.steps + step) * self.world + rank


class PeerLogitRing:
    """The exchange step without a collective call: a ring of logits blocks in symmetric memory
    (``torch.distributed._symmetric_memory``: every rank's buffer is mapped into every process over
    NVLink / NVSwitch) that the forward kernel itself writes into on ALL ranks
    (``B200Model.forward_compact_gather`` -> ``rb200_forward_compact_gather``).

    Layout of every rank's buffer: float32 [slots][world][steps][batch][num_out], then one uint32 arrival
    counter per (slot, step, source rank).  ``forward(model, arrays, slot, step)`` runs this rank's
    batch and lands its logits in block [slot][rank][step] everywhere; ``block(slot)`` is the local view
    [world, steps, batch, num_out] in rank order - the same tensor an all-gather would have produced.
    The NCCL path (``ShardedCaller.gather``) stays as the checked fallback.

    ``deferred=True`` (default): the kernel's compute blocks store into this rank's own block only and one
    extra thread block of the NEXT launch ships the finished block to the peers
    (``rb200_forward_compact_ship``), so a remote store never sits between a compute block and the SM slot
    it frees; ``flush()`` ships the last block.  ``deferred=False``: every compute block stores to every rank
    itself (``rb200_forward_compact_gather``)."""

    def __init__(self, model, slots, steps, batch, group=None, multicast=False, deferred=True):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.slots, self.steps, self.batch, self.num_out = slots, steps, batch, model.num_out
        self.layout = LogitRingLayout(self.world, slots, steps, batch, model.num_out)
        self.block_floats, self.n_data, self.n_flags = (self.layout.block_floats, self.layout.n_data,
                                                        self.layout.n_flags)
        self.buf = symm_mem.empty(self.n_data + self.n_flags, dtype=torch.float32, device=model.device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        self.peers_dev = int(self.handle.buffer_ptrs_dev)
        self.multicast_ptr = 0
        if multicast and getattr(self.handle, "has_multicast_support", False):
            self.multicast_ptr = int(self.handle.multicast_ptr or 0)
        self.deferred = deferred
        self._pending = None     # (local block tensor, float offset, flag word) still to be shipped
        self._shape_hint = None
        self._model = model
        self.handle.barrier()

    def offset(self, slot, step, rank=None):
        return self.layout.offset(slot, step, self.rank if rank is None else rank)

    def flag_word(self, slot, step, rank=None):
        return self.layout.flag_word(slot, step, self.rank if rank is None else rank)

    def forward(self, model, arrays, slot, step, signal=True):
        """``signal``: also bump the per-(slot, step, source) arrival counters on every rank (one
        system-scope fence + atomic per thread block); consumers that only read after a stream / group
        synchronisation can switch it off."""
        if not self.deferred:
            model.forward_compact_gather(*arrays, self.peers_dev, self.world, self.offset(slot, step),
                                         multicast_ptr=self.multicast_ptr,
                                         flag_word=self.flag_word(slot, step) if signal else -1)
            return
        off = self.offset(slot, step)
        mine = self.buf[off:off + arrays[0].shape[0] * self.num_out]
        pend = self._pending or (None, 0, -1)
        model.forward_compact_ship(arrays, mine, self.peers_dev, self.world, self.rank, ship_src=pend[0],
                                   ship_dst_offset=pend[1], multicast_ptr=self.multicast_ptr, flag_word=pend[2])
        self._pending = (mine, off, self.flag_word(slot, step) if signal else -1)
        self._shape_hint = (arrays[0].shape[-1], arrays[1].shape[1], arrays[2].shape[1])

    def flush(self):
        """Ship the block of the last ``forward`` (deferred mode; a one-block launch on the current stream)."""
        if self._pending is None:
            return
        pend, self._pending = self._pending, None
        self._model.forward_compact_ship(None, None, self.peers_dev, self.world, self.rank, ship_src=pend[0],
                                         ship_dst_offset=pend[1], multicast_ptr=self.multicast_ptr,
                                         flag_word=pend[2], shape_hint=self._shape_hint)

    def block(self, slot):
        n = self.world * self.steps * self.block_floats
        return self.buf[slot * n:(slot + 1) * n].view(self.world, self.steps, self.block_floats)[
            :, :, :self.batch * self.num_out].view(self.world, self.steps, self.batch, self.num_out)

    def arrivals(self, slot):
        """uint32 arrival counters [steps, world] of a slot (CTAs that have delivered, cumulative)."""
        st = self.n_data + slot * self.steps * self.world
        return self.buf[st:st + self.steps * self.world].view(torch.int32).view(self.steps, self.world)

    def barrier(self):
        self.handle.barrier()
