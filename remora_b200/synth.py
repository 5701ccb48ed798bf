"""Seeded synthetic chunk batches and reads (SURVEY.md §8d "Synthetic inputs").

Used by tests/, tests/golden/make_golden.py and bench.py.  Everything is generated with
``numpy.random.default_rng(seed)`` so the same seed gives the same arrays here and on the GPU box.

Array layout follows the reference's compact chunk format
(``CoreRemoraDataset._core_dtypes``, reference src/remora/data_chunks.py:942-948):
  signal                      float32 [N, 1, T]
  sequence                    int8    [N, Lmax + kmer_len - 1]   values -1..3 (-1 = N / beyond read end)
  sequence_to_signal_mapping  int16   [N, Lmax + 1]              monotone, [0] == 0, [seq_len] == T
  sequence_lengths            int16   [N]
Padding past ``seq_len`` is filled with garbage on purpose (the reference leaves it
uninitialised, data_chunks.py:1379-1388): kernels must never read it.
"""
import numpy as np


def synth_chunks(n, chunk_len=100, kmer_context_bases=(4, 4), seed=0, stride=5,
                 frac_n=0.01, frac_edge=0.05, max_seq_len=None, garbage_padding=True, seq_len_range=None):
    rng = np.random.default_rng(seed)
    kmer_len = sum(kmer_context_bases) + 1
    T = int(chunk_len)
    lmax = T // stride if max_seq_len is None else int(max_seq_len)
    n_slots = (T - 1) // stride  # interior multiples of `stride` strictly inside (0, T)
    signal = rng.standard_normal((n, 1, T), dtype=np.float32)
    seq_w = lmax + kmer_len - 1
    if garbage_padding:
        sequence = rng.integers(-128, 127, size=(n, seq_w), dtype=np.int8)
        mapping = rng.integers(-32768, 32767, size=(n, lmax + 1), dtype=np.int16)
    else:
        sequence = np.full((n, seq_w), -1, dtype=np.int8)
        mapping = np.zeros((n, lmax + 1), dtype=np.int16)
    # number of bases per chunk: mean dwell ~ 8-12 samples, at least 1, at most lmax
    lo = max(1, T // 12)
    hi = max(lo, min(lmax, T // 8))
    if seq_len_range is not None:  # explicit range of bases per chunk (short dwells, zero-dwell bases)
        lo, hi = max(1, int(seq_len_range[0])), min(lmax, int(seq_len_range[1]))
    seq_lens = rng.integers(lo, hi + 1, size=n).astype(np.int16)
    for i in range(n):
        L = int(seq_lens[i])
        n_inner = min(L - 1, n_slots)
        inner = np.sort(rng.choice(n_slots, size=n_inner, replace=False) + 1) * stride
        bounds = np.concatenate([[0], inner, [T]])
        if bounds.size < L + 1:  # more bases than stride slots: repeat boundaries (zero-dwell bases)
            extra = rng.choice(bounds, size=L + 1 - bounds.size)
            bounds = np.sort(np.concatenate([bounds, extra]))
            bounds[0], bounds[-1] = 0, T
        mapping[i, : L + 1] = bounds
        bases = rng.integers(0, 4, size=L + kmer_len - 1).astype(np.int8)
        bases[rng.random(bases.size) < frac_n] = -1
        if rng.random() < frac_edge:  # read-edge padding (reference data_chunks.py:394-409)
            run = int(rng.integers(1, max(2, kmer_len)))
            if rng.random() < 0.5:
                bases[:run] = -1
            else:
                bases[-run:] = -1
        sequence[i, : L + kmer_len - 1] = bases
    return {
        "signal": signal,
        "sequence": sequence,
        "sequence_to_signal_mapping": mapping,
        "sequence_lengths": seq_lens,
        "kmer_context_bases": tuple(int(x) for x in kmer_context_bases),
        "chunk_len": T,
    }


def synth_read(n_bases=400, seed=0, mean_dwell=10, frac_n=0.0):
    """Raw pieces of a synthetic read: (dacs int16-valued float array, shift, scale,
    seq_to_sig_map int64, int_seq int64).  Mirrors the argument list of the reference's
    ``RemoraRead`` dataclass (data_chunks.py:151-160)."""
    rng = np.random.default_rng(seed)
    dwells = rng.integers(max(1, mean_dwell // 3), mean_dwell * 2, size=n_bases)
    seq_to_sig_map = np.concatenate([[0], np.cumsum(dwells)]).astype(np.int64)
    int_seq = rng.integers(0, 4, size=n_bases).astype(np.int64)
    if frac_n > 0:
        int_seq[rng.random(n_bases) < frac_n] = -1
    levels = rng.normal(0.0, 1.0, size=n_bases)
    sig = np.repeat(levels, dwells) + rng.normal(0.0, 0.3, size=int(dwells.sum()))
    shift, scale = 431.5, 87.25
    dacs = np.round(sig * scale + shift).astype(np.int16)
    return dacs, shift, scale, seq_to_sig_map, int_seq


def synth_levels_table(kmer_len=6, seed=0):
    """Seeded k-mer level table (float32 [4**kmer_len], index = base-4 number of the k-mer, first
    base most significant, as the reference's ``index_from_int_kmer``,
    refine_signal_map_core.pyx:76-84).  Roughly unit-variance like a gauge-fixed ONT table."""
    rng = np.random.default_rng(seed)
    return rng.normal(0.0, 1.0, size=4 ** kmer_len).astype(np.float32)


def synth_refine_read(n_bases=400, table=None, kmer_len=6, center_idx=2, seed=0, mean_dwell=10,
                      noise=0.25, stride=5, jitter=6, frac_zero_dwell=0.02, frac_stall=0.01,
                      stall_range=(60, 400), shift=431.5, scale=87.25, scale_error=1.08, shift_error=9.0):
    """Raw pieces of a synthetic read whose signal follows a k-mer level table, with a deliberately
    imperfect starting mapping (what a basecaller move table + alignment gives the reference):
    boundaries snapped to multiples of ``stride`` and jittered, some zero-dwell bases (deletions),
    some stalls, and shift/scale estimates that are slightly off so rough re-scaling has work to do.
    Returns (dacs int16, shift, scale, seq_to_sig_map int64, int_seq int64)."""
    rng = np.random.default_rng(seed)
    if table is None:
        table = synth_levels_table(kmer_len, seed=0)
    int_seq = rng.integers(0, 4, size=n_bases).astype(np.int64)
    powers = 4 ** np.arange(kmer_len - 1, -1, -1)
    levels = np.zeros(n_bases, dtype=np.float64)
    if n_bases >= kmer_len:
        windows = np.lib.stride_tricks.sliding_window_view(int_seq, kmer_len)
        levels[center_idx:center_idx + windows.shape[0]] = table[windows @ powers]
    dwells = rng.integers(max(2, mean_dwell // 3), mean_dwell * 2, size=n_bases)
    stall = rng.random(n_bases) < frac_stall
    dwells[stall] = rng.integers(stall_range[0], stall_range[1], size=int(stall.sum()))
    true_map = np.concatenate([[0], np.cumsum(dwells)]).astype(np.int64)
    sig_len = int(true_map[-1])
    sig = np.repeat(levels, dwells) + rng.normal(0.0, noise, size=sig_len)
    dacs = np.clip(np.round(sig * scale + shift), -32768, 32767).astype(np.int16)
    # starting map: jitter, snap to the move stride, keep monotone, pin the ends
    start = true_map + rng.integers(-jitter, jitter + 1, size=true_map.size)
    start = (np.round(start / stride) * stride).astype(np.int64)
    start = np.maximum.accumulate(np.clip(start, 0, sig_len))
    zero = np.nonzero(rng.random(n_bases - 1) < frac_zero_dwell)[0] + 1
    start[zero] = start[zero - 1]
    start = np.maximum.accumulate(start)
    # the last base keeps at least one sample (the reference's rough re-scaling indexes the centre
    # sample of every base, refine_signal_map.py:415)
    start = np.minimum(start, sig_len - 1)
    start[0], start[-1] = 0, sig_len
    return dacs, shift + shift_error, scale * scale_error, start, int_seq


def synth_pod5_bam_run(pod5_path, bam_path, n_reads=12, seed=0, bases=(150, 600), kmer_len=6, center_idx=2):
    """Writes a POD5 + BAM pair (``remora_b200.io`` writers) whose reads follow the seeded k-mer table:
    signal from levels, a stride-5 move table, ``ts`` trimming, ``sm``/``sd`` scaling tags, unmapped
    records.  Returns (pod5_path, bam_path, {read_id: pieces of the RemoraRead the readers must give})."""
    import uuid
    from . import io
    rng = np.random.default_rng(seed)
    table = synth_levels_table(kmer_len, 0)
    ids = [str(uuid.UUID(int=int(rng.integers(1, 2 ** 62)))) for _ in range(n_reads)]
    pod5_reads, bam_recs, truth = [], [], {}
    powers = 4 ** np.arange(kmer_len - 1, -1, -1)
    for rid in ids:
        n = int(rng.integers(bases[0], bases[1]))
        int_seq = rng.integers(0, 4, size=n)
        levels = np.zeros(n)
        win = np.lib.stride_tricks.sliding_window_view(int_seq, kmer_len) @ powers
        levels[center_idx:center_idx + win.size] = table[win]
        dwells = rng.integers(1, 5, size=n) * 5
        ts = int(rng.integers(0, 4)) * 5
        pa = np.repeat(levels, dwells) * 26.0 + 88.0 + rng.normal(0, 6.0, size=int(dwells.sum()))
        cal_off, cal_scale = -240.0, 0.18
        dacs = np.round(pa / cal_scale - cal_off).astype(np.int16)
        full = np.concatenate([rng.integers(400, 900, size=ts).astype(np.int16), dacs])
        mv = np.zeros(dacs.size // 5, dtype=np.int8)
        mv[(np.cumsum(dwells) - dwells) // 5] = 1
        seq = "".join("ACGT"[b] for b in int_seq)
        pod5_reads.append((rid, full, cal_off, cal_scale))
        bam_recs.append(dict(query_name=rid, flag=4, query_sequence=seq,
                             tags=[("mv", "Bc", np.r_[5, mv].astype(np.int8)), ("ts", "i", ts),
                                   ("ns", "i", full.size), ("sm", "f", 88.0), ("sd", "f", 26.0)]))
        truth[rid] = dict(dacs=dacs, seq=seq, ssm=np.concatenate([np.cumsum(dwells) - dwells, [dacs.size]]),
                          shift=-cal_off + (1 / np.float32(cal_scale)) * np.float32(88.0),
                          scale=(1 / np.float32(cal_scale)) * np.float32(26.0))
    io.write_pod5(pod5_path, pod5_reads)
    io.write_bam(bam_path, "@HD\tVN:1.6\tSO:unknown\n", [], bam_recs)
    return pod5_path, bam_path, truth
