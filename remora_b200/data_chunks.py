"""Read -> chunk batches, behind the reference's ``remora.data_chunks.RemoraRead`` surface
(src/remora/data_chunks.py:125-540).

Differences from the reference that matter for speed, none for results:
  * chunk extraction is vectorised over all focus bases of a read (numpy index arithmetic) instead
    of one Python ``extract_chunk`` call + in-memory dataset write per chunk;
  * batches carry the reference's COMPACT chunk arrays (``CoreRemoraDataset._core_dtypes``,
    data_chunks.py:942-948) and the one-hot k-mer tensor is never built on the host: ``run_model``
    hands the compact arrays to ``model.forward_compact`` (fused encode+forward CUDA kernels), or,
    for a model without that method, encodes on the GPU with ``rb200_encode_dense`` first.
"""
import dataclasses

import numpy as np
import torch

from . import RemoraError, constants, util


@dataclasses.dataclass
class ChunkBatch:
    """One batch of chunks in the reference's compact array format."""

    signal: np.ndarray            # float32 [B, 1, T]
    sequence: np.ndarray          # int8    [B, Lmax + kmer_len - 1]
    seq_to_sig_map: np.ndarray    # int16   [B, Lmax + 1]
    seq_lens: np.ndarray          # int16   [B]
    labels: np.ndarray            # int64   [B]
    read_focus_bases: np.ndarray  # int64   [B]
    kmer_context_bases: tuple

    def __len__(self):
        return self.seq_lens.shape[0]

    def enc_kmers(self, device):
        """Dense one-hot tensor on ``device`` (GPU encode kernel); only needed for models that
        lack ``forward_compact``."""
        from .encoded_kmers import compute_encoded_kmer_batch_torch
        return compute_encoded_kmer_batch_torch(
            self.kmer_context_bases[0], self.kmer_context_bases[1], torch.from_numpy(self.sequence),
            torch.from_numpy(self.seq_to_sig_map), torch.from_numpy(self.seq_lens),
            sig_len=self.signal.shape[-1], device=device)


@dataclasses.dataclass
class DeviceChunkBatch:
    """Same compact chunk arrays as :class:`ChunkBatch`, already resident on the GPU (produced by
    ``RemoraRead.prepare_batches_gpu``); ``labels`` / ``read_focus_bases`` stay on the host."""

    signal: torch.Tensor          # float32 [B, 1, T]   (device)
    sequence: torch.Tensor        # int8    [B, Lmax + kmer_len - 1]
    seq_to_sig_map: torch.Tensor  # int16   [B, Lmax + 1]
    seq_lens: torch.Tensor        # int16   [B]
    labels: np.ndarray
    read_focus_bases: np.ndarray
    kmer_context_bases: tuple

    def __len__(self):
        return self.seq_lens.shape[0]

    def enc_kmers(self, device):
        from .encoded_kmers import compute_encoded_kmer_batch_torch
        return compute_encoded_kmer_batch_torch(
            self.kmer_context_bases[0], self.kmer_context_bases[1], self.sequence, self.seq_to_sig_map,
            self.seq_lens, sig_len=self.signal.shape[-1], device=device)


@dataclasses.dataclass
class Chunk:
    """Single chunk, same fields as the reference's ``Chunk`` (data_chunks.py:543-641)."""

    signal: np.ndarray
    seq_w_context: np.ndarray
    seq_to_sig_map: np.ndarray
    kmer_context_bases: tuple
    chunk_sig_focus_idx: int
    chunk_focus_base: int
    read_focus_base: int
    read_id: str = None
    label: int = None

    @property
    def kmer_len(self):
        return sum(self.kmer_context_bases) + 1

    @property
    def seq_len(self):
        return self.seq_w_context.size - sum(self.kmer_context_bases)

    def check(self):
        if self.signal.size <= 0:
            raise RemoraError("No signal for chunk")
        if np.any(np.isnan(self.signal)):
            raise RemoraError("Signal contains NaN")
        if self.seq_w_context.size - sum(self.kmer_context_bases) != self.seq_to_sig_map.size - 1:
            raise RemoraError("Invalid sig to seq map length")
        if self.seq_to_sig_map[0] < 0:
            raise RemoraError("Seq to sig map starts before 0")
        if self.seq_to_sig_map[-1] > self.signal.size:
            raise RemoraError("Seq to sig map ends after signal")


@dataclasses.dataclass
class RemoraRead:
    """Same dataclass fields, in the same order, as the reference (data_chunks.py:151-160).

    ``dacs`` un-normalised signal (already reversed for reverse_signal models); normalised signal is
    ``(dacs - shift) / scale`` as float32 (data_chunks.py:191-197); ``seq_to_sig_map`` has one more
    entry than the sequence; ``int_seq`` codes A0 C1 G2 T3 N-1."""

    dacs: np.ndarray
    shift: float
    scale: float
    seq_to_sig_map: np.ndarray
    int_seq: np.ndarray = None
    str_seq: str = None
    read_id: str = None
    labels: np.ndarray = None
    focus_bases: np.ndarray = None
    batches: list = None

    def __post_init__(self):
        if self.int_seq is None:
            if self.str_seq is None:
                raise RemoraError("Must provide sequence to initialize RemoraRead")
            self.int_seq = util.seq_to_int(self.str_seq)
        else:
            self.str_seq = util.int_to_seq(self.int_seq)
        self._sig = None

    @classmethod
    def test_read(cls, nbases=20, signal_per_base=10):
        """Spoofed read: zero signal, ``nbases`` x ``signal_per_base`` samples, ACGT repeats
        (reference data_chunks.py:178-189; the reference passes its positional arguments so that
        the label zeros land in ``read_id`` and the id string in ``str_seq`` - reproduced)."""
        return cls(np.zeros(nbases * signal_per_base), 0.0, 1.0,
                   np.arange(nbases * signal_per_base + 1, step=signal_per_base),
                   np.arange(nbases) % 4, "test_read", np.zeros(nbases, dtype=np.int64))

    @property
    def sig(self):
        if self._sig is None:
            self._sig = ((self.dacs - self.shift) / self.scale).astype(np.float32)
        return self._sig

    def check(self):
        """Same validity rules as the reference (data_chunks.py:222-251)."""
        if self.seq_to_sig_map.size != self.int_seq.size + 1:
            raise RemoraError(f"Invalid read: seq ({self.int_seq.size}) and mapping "
                              f"({self.seq_to_sig_map.size}) sizes incompatible")
        if self.seq_to_sig_map[0] != 0:
            raise RemoraError("Invalid read: mapping start")
        if self.seq_to_sig_map[-1] != np.asarray(self.dacs).size:  # (not self.sig.size: that would normalise the signal)
            raise RemoraError("Invalid read: mapping end")
        if self.int_seq.max() > 3 or self.int_seq.min() < -1:
            raise RemoraError("Invalid read: Invalid base")

    def copy(self):
        return RemoraRead(
            dacs=self.dacs.copy(), shift=self.shift, scale=self.scale,
            seq_to_sig_map=self.seq_to_sig_map,
            int_seq=None if self.int_seq is None else self.int_seq.copy(), str_seq=self.str_seq,
            read_id=self.read_id, labels=None if self.labels is None else self.labels.copy(),
            focus_bases=None if self.focus_bases is None else self.focus_bases.copy())

    def refine_signal_mapping(self, sig_map_refiner, check_read=False):
        """No-op for an unloaded refiner, like the reference (data_chunks.py:267-269); a loaded
        refiner (``remora_b200.refine_signal_map.SigMapRefiner``: re-scaling on the host, banded DP on
        the GPU) is applied the same way the reference applies it (data_chunks.py:270-308).  Many
        reads at once: ``SigMapRefiner.refine_reads`` (one launch for the whole batch)."""
        if not sig_map_refiner.is_loaded:
            return
        if sig_map_refiner.do_rough_rescale:
            self.shift, self.scale = sig_map_refiner.rough_rescale(
                self.shift, self.scale, self.seq_to_sig_map, self.int_seq, self.dacs)
            self._sig = None
        if sig_map_refiner.scale_iters >= 0:
            try:
                self.seq_to_sig_map, self.shift, self.scale = sig_map_refiner.refine_sig_map(
                    self.shift, self.scale, self.seq_to_sig_map, self.int_seq, self.dacs)
            except IndexError:
                pass
            self._sig = None
        if check_read:
            self.check()

    def set_motif_focus_bases(self, motifs):
        """focus_bases <- all hits of any motif in int_seq (data_chunks.py:310-317)."""
        self.focus_bases = util.find_focus_bases_in_int_sequence(self.int_seq, motifs)

    # ------------------------------------------------------------------------------------------
    # chunk extraction
    # ------------------------------------------------------------------------------------------
    def _focus_signal_positions(self, base_start_justify, offset):
        """Signal index each focus base's chunk is centred on (data_chunks.py:443-453)."""
        ssm = np.asarray(self.seq_to_sig_map)
        fb = np.clip(np.asarray(self.focus_bases, dtype=np.int64) + offset, 0, ssm.size - 2)
        if base_start_justify:
            return fb, ssm[fb].astype(np.int64)
        return fb, ((ssm[fb] + ssm[fb + 1]) // 2).astype(np.int64)

    def extract_chunk(self, focus_sig_idx, chunk_context, kmer_context_bases, label=-1,
                      read_focus_base=-1, check_chunk=False, signal_padding=False):
        """Single-chunk form with the reference's signature (data_chunks.py:331-423)."""
        if signal_padding:
            raise RemoraError("signal_padding is a training-time option not used by inference")
        arrays = self._extract_arrays(np.array([focus_sig_idx], dtype=np.int64), chunk_context,
                                      kmer_context_bases)
        sig, seq, ssm, lens, seq_start, sig_start = arrays
        L = int(lens[0])
        chunk = Chunk(signal=sig[0], seq_w_context=seq[0, :L + sum(kmer_context_bases)],
                      seq_to_sig_map=ssm[0, :L + 1].astype(np.int32),
                      kmer_context_bases=tuple(kmer_context_bases),
                      chunk_sig_focus_idx=int(focus_sig_idx - sig_start[0]),
                      chunk_focus_base=int(read_focus_base - seq_start[0]),
                      read_focus_base=read_focus_base, read_id=self.read_id, label=label)
        if check_chunk:
            chunk.check()
        return chunk

    def _extract_arrays(self, focus_sig_idx, chunk_context, kmer_context_bases):
        """Vectorised restatement of ``extract_chunk`` (data_chunks.py:331-423) for many focus
        positions at once.  Returns (signal f32 [N,T], sequence i8 [N,Lmax+k-1] (-1 padded),
        mapping int64 [N,Lmax+1], seq_lens int64 [N], seq_start, clipped sig_start)."""
        sig = self.sig
        ssm = np.asarray(self.seq_to_sig_map, dtype=np.int64)
        int_seq = np.asarray(self.int_seq)
        n = focus_sig_idx.size
        T = int(sum(chunk_context))
        kb, ka = int(kmer_context_bases[0]), int(kmer_context_bases[1])
        raw_start = focus_sig_idx - int(chunk_context[0])
        raw_end = focus_sig_idx + int(chunk_context[1])
        # zero padding where the chunk sticks out of the read signal (data_chunks.py:346-361)
        pad_left = np.maximum(-raw_start, 0)
        sig_start = np.maximum(raw_start, 0)
        sig_end = np.minimum(raw_end, sig.size)
        cols = np.arange(T, dtype=np.int64)[None, :]
        src = raw_start[:, None] + cols
        valid = (src >= 0) & (src < sig.size)
        chunk_sig = np.where(valid, sig[np.clip(src, 0, max(sig.size - 1, 0))],
                             np.float32(0)).astype(np.float32)
        # bases overlapping [sig_start, sig_end) (data_chunks.py:370-373)
        seq_start = np.searchsorted(ssm, sig_start, side="right") - 1
        seq_end = np.searchsorted(ssm, sig_end, side="left")
        seq_lens = seq_end - seq_start
        if n and seq_lens.min() < 1:
            raise RemoraError("chunk without sequence")
        lmax = int(seq_lens.max()) if n else 0
        # mapping relative to the chunk, ends pinned to the chunk boundaries (:376-382)
        mcols = np.arange(lmax + 1, dtype=np.int64)[None, :]
        midx = np.minimum(seq_start[:, None] + mcols, ssm.size - 1)
        mapping = ssm[midx] - (sig_start - pad_left)[:, None]
        mapping[:, 0] = 0
        mapping[np.arange(n), seq_lens] = T
        mapping = np.where(mcols <= seq_lens[:, None], mapping, 0)
        # sequence with k-mer context, -1 beyond the read ends (:385-409)
        width = lmax + kb + ka
        scols = np.arange(width, dtype=np.int64)[None, :]
        sidx = seq_start[:, None] - kb + scols
        in_read = (sidx >= 0) & (sidx < int_seq.size) & (scols < (seq_lens + kb + ka)[:, None])
        sequence = np.where(in_read, int_seq[np.clip(sidx, 0, max(int_seq.size - 1, 0))], -1)
        return chunk_sig, sequence.astype(np.int8), mapping, seq_lens, seq_start, sig_start

    def iter_chunks(self, chunk_context, kmer_context_bases, base_start_justify=False, offset=0,
                    check_chunks=False, motifs=None):
        """Generator of ``Chunk`` objects, reference signature (data_chunks.py:425-466)."""
        for focus_base in self.focus_bases:
            if motifs is not None and not any(m.match(self.int_seq, focus_base) for m in motifs):
                continue
            label = -1 if self.labels is None else self.labels[focus_base]
            fb = max(min(focus_base + offset, self.seq_to_sig_map.size - 2), 0)
            if base_start_justify:
                idx = self.seq_to_sig_map[fb]
            else:
                idx = (self.seq_to_sig_map[fb] + self.seq_to_sig_map[fb + 1]) // 2
            try:
                yield self.extract_chunk(idx, chunk_context, kmer_context_bases, label=label,
                                         read_focus_base=fb, check_chunk=check_chunks)
            except RemoraError:
                continue

    def prepare_batches(self, model_metadata, batch_size=constants.DEFAULT_BATCH_SIZE):
        """Build ``self.batches`` (list of :class:`ChunkBatch`) for every focus base
        (reference data_chunks.py:468-514).  The reference ignores ``batch_size`` and always
        uses 2048 (data_chunks.py:489-503); batch boundaries do not affect any result, so the
        argument is honoured here."""
        self.batches = []
        self.refine_signal_mapping(model_metadata["sig_map_refiner"])
        if self.focus_bases is None or len(self.focus_bases) == 0:
            return
        chunk_context = tuple(model_metadata["chunk_context"])
        kmer_context = tuple(model_metadata["kmer_context_bases"])
        fb, focus_sig = self._focus_signal_positions(model_metadata["base_start_justify"],
                                                     model_metadata["offset"])
        labels = (np.full(fb.size, -1, dtype=np.int64) if self.labels is None
                  else np.asarray(self.labels)[np.asarray(self.focus_bases)].astype(np.int64))
        try:
            sig, seq, mapping, lens, _, _ = self._extract_arrays(focus_sig, chunk_context,
                                                                 kmer_context)
        except RemoraError:
            return
        T = sum(chunk_context)
        if T >= 32768:
            raise RemoraError("chunk_len does not fit the int16 mapping")
        batch_size = max(1, int(batch_size or constants.DEFAULT_BATCH_SIZE))
        for st in range(0, fb.size, batch_size):
            en = min(st + batch_size, fb.size)
            self.batches.append(ChunkBatch(
                signal=np.ascontiguousarray(sig[st:en, None, :]),
                sequence=np.ascontiguousarray(seq[st:en]),
                seq_to_sig_map=np.ascontiguousarray(mapping[st:en].astype(np.int16)),
                seq_lens=np.ascontiguousarray(lens[st:en].astype(np.int16)),
                labels=labels[st:en], read_focus_bases=fb[st:en].astype(np.int64),
                kmer_context_bases=kmer_context))

    def prepare_batches_gpu(self, model_metadata, batch_size=constants.DEFAULT_BATCH_SIZE, device=None):
        """``prepare_batches`` with the chunk extraction itself on the GPU ("next" row 1):
        the read's raw arrays go to the device once (DAC samples, mapping, sequence: a few bytes per
        sample instead of ~480 B per chunk) and two kernels (``rb200_chunk_plan`` / ``rb200_chunk_fill``)
        build the compact chunk arrays there, bit-identical to the host path above.  Batches are
        :class:`DeviceChunkBatch` objects; ``run_model`` consumes them without further copies."""
        import ctypes
        from . import _native
        lib = _native.load_library()
        self.batches = []
        self.refine_signal_mapping(model_metadata["sig_map_refiner"])
        if self.focus_bases is None or len(self.focus_bases) == 0:
            return
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        device = torch.device(device)
        c0, c1 = (int(x) for x in model_metadata["chunk_context"])
        kb, ka = (int(x) for x in model_metadata["kmer_context_bases"])
        T = c0 + c1
        if T >= 32768:
            raise RemoraError("chunk_len does not fit the int16 mapping")
        dacs = np.ascontiguousarray(self.dacs)
        if dacs.dtype == np.int16:
            code = 0
        elif dacs.dtype == np.float32 and ((dacs[:1] - self.shift) / self.scale).dtype == np.float32:
            code = 1  # python-float scalars: numpy stays in float32 (two float32 roundings)
        else:  # everything else - incl. float32 samples with np.float64 shift / scale, as rough re-scaling leaves
            # them - is computed in float64 by numpy in (dacs - shift) / scale; float32 -> float64 is exact
            dacs, code = dacs.astype(np.float64), 2
        ssm = np.ascontiguousarray(self.seq_to_sig_map, dtype=np.int32)
        focus = np.ascontiguousarray(self.focus_bases, dtype=np.int32)
        n = focus.size
        with torch.cuda.device(device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
            d_dacs = torch.from_numpy(dacs).to(device)
            d_ssm = torch.from_numpy(ssm).to(device)
            d_seq = torch.from_numpy(np.ascontiguousarray(self.int_seq, dtype=np.int8)).to(device)
            d_focus = torch.from_numpy(focus).to(device)
            plan = torch.empty((4, n), dtype=torch.int32, device=device)  # focus_adj, focus_sig, start, len
            ptr = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
            _native.check(lib.rb200_chunk_plan(ptr(d_ssm), ssm.size, dacs.size, ptr(d_focus), n, c0, c1,
                                               int(bool(model_metadata["base_start_justify"])),
                                               int(model_metadata["offset"]), ptr(plan[0]), ptr(plan[1]),
                                               ptr(plan[2]), ptr(plan[3]), stream), "rb200_chunk_plan")
            lo, hi = int(plan[3].min()), int(plan[3].max())
            if lo < 1:  # a chunk without sequence: the host path drops the read as well
                return
            signal = torch.empty((n, 1, T), dtype=torch.float32, device=device)
            sequence = torch.empty((n, hi + kb + ka), dtype=torch.int8, device=device)
            mapping = torch.empty((n, hi + 1), dtype=torch.int16, device=device)
            lens = torch.empty((n,), dtype=torch.int16, device=device)
            _native.check(lib.rb200_chunk_fill(ptr(d_dacs), code, dacs.size, float(self.shift),
                                               float(self.scale), ptr(d_ssm), ssm.size, ptr(d_seq),
                                               d_seq.numel(), ptr(plan[1]), ptr(plan[2]), ptr(plan[3]), n,
                                               c0, c1, kb, ka, hi, ptr(signal), ptr(sequence), ptr(mapping),
                                               ptr(lens), stream), "rb200_chunk_fill")
            fb = plan[0].cpu().numpy().astype(np.int64)
        labels = (np.full(n, -1, dtype=np.int64) if self.labels is None
                  else np.asarray(self.labels)[np.asarray(self.focus_bases)].astype(np.int64))
        batch_size = max(1, int(batch_size or constants.DEFAULT_BATCH_SIZE))
        for st in range(0, n, batch_size):
            en = min(st + batch_size, n)
            self.batches.append(DeviceChunkBatch(
                signal=signal[st:en], sequence=sequence[st:en], seq_to_sig_map=mapping[st:en],
                seq_lens=lens[st:en], labels=labels[st:en], read_focus_bases=fb[st:en],
                kmer_context_bases=(kb, ka)))

    def run_model(self, model, keep_on_device=False):
        """Call modified bases on this read's prepared batches (reference data_chunks.py:516-540).
        Returns (nn_out float32 [N,num_out], labels int64 [N], read positions int64 [N]);
        ``keep_on_device`` (extension) leaves nn_out as one device tensor for device post-processing."""
        device = next(model.parameters()).device
        outputs, labels, poss = [], [], []
        compact = hasattr(model, "forward_compact")
        for batch in self.batches:
            on_device = isinstance(batch, DeviceChunkBatch)
            sigs = batch.signal.to(device) if on_device else torch.from_numpy(batch.signal).to(device)
            if compact and on_device:
                out = model.forward_compact(sigs, batch.sequence, batch.seq_to_sig_map, batch.seq_lens)
            elif compact:
                out = model.forward_compact(sigs, torch.from_numpy(batch.sequence),
                                            torch.from_numpy(batch.seq_to_sig_map),
                                            torch.from_numpy(batch.seq_lens))
            else:
                out = model(sigs, batch.enc_kmers(device))
            outputs.append(out.detach() if keep_on_device else out.detach().cpu().numpy())
            labels.append(batch.labels)
            poss.append(batch.read_focus_bases)
        nn_out = torch.cat(outputs, dim=0) if keep_on_device else np.concatenate(outputs, axis=0)
        return nn_out, np.concatenate(labels), np.concatenate(poss)
