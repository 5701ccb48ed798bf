"""Signal-mapping refinement behind the reference's ``remora.refine_signal_map`` surface
(src/remora/refine_signal_map.py), "next" row 4 of SURVEY.md 8f.

``SigMapRefiner`` keeps the reference's dataclass fields and methods (``rough_rescale``, ``rescale``,
``refine_sig_map``, ``extract_levels``, ``load_from_metadata`` ...).  The banded dynamic programme
(``seq_banded_dp``, refine_signal_map_core.pyx:403-473, >95 % of the reference's time per read) runs on
the GPU through the C ABI (``rb200_refine_normalize`` / ``rb200_refine_dp``, one warp per read, a whole
batch of reads per launch: :meth:`SigMapRefiner.refine_reads`); scores, traceback and the returned
mapping are bit-identical to the reference's.  The cheap per-read set-up stays on the host as vectorised
numpy: level lookup (core.pyx:87-100), quantile / least-squares re-scaling (:67-122), band construction
(:634-775, core.pyx:31-69).  There is no CPU fallback for the dynamic programme.
"""
import ctypes
import dataclasses

import numpy as np
import torch

from . import RemoraError, _native

DEFAULT_REFINE_HBW = 5                                   # constants.py:42
DEFAULT_REFINE_SHORT_DWELL_PARAMS = (4, 3, 0.5)          # constants.py:233
REFINE_ALGO_VIT_NAME = "Viterbi"                         # constants.py:234
REFINE_ALGO_DWELL_PEN_NAME = "dwell_penalty"             # constants.py:235
REFINE_ALGOS = (REFINE_ALGO_DWELL_PEN_NAME, REFINE_ALGO_VIT_NAME)
DEFAULT_REFINE_ALGO = REFINE_ALGO_DWELL_PEN_NAME
ROUGH_RESCALE_LEAST_SQUARES = "least_squares"
ROUGH_RESCALE_THEIL_SEN = "theil_sen"
ROUGH_RESCALE_METHODS = (ROUGH_RESCALE_LEAST_SQUARES, ROUGH_RESCALE_THEIL_SEN)
DEFAULT_ROUGH_RESCALE_METHOD = ROUGH_RESCALE_LEAST_SQUARES
MAX_POINTS_FOR_THEIL_SEN = 1000                          # constants.py:245
_ALGO_CODE = {REFINE_ALGO_VIT_NAME: 0, REFINE_ALGO_DWELL_PEN_NAME: 1}
MAX_DWELL_PENALTIES = 16                                 # rb200_refine_dp limit


def compute_dwell_pen_array(target, limit, weight):
    """refine_signal_map.py:34-41"""
    if limit > target:
        limit = target
    return weight * np.square(np.arange(limit, dtype=np.float32) - target)


DEFAULT_REFINE_SHORT_DWELL_PEN = compute_dwell_pen_array(*DEFAULT_REFINE_SHORT_DWELL_PARAMS)


# ------------------------------------------------------------------------------------------------
# re-scaling (host, float64 numpy like the reference)          refine_signal_map.py:54-122
# ------------------------------------------------------------------------------------------------
def quantile_linear(a, q):
    """``np.quantile(a, q)`` (default "linear" method) for a 1-D float array and a float64 array ``q``,
    with the same operations in the same precision (numpy ``_quantile`` / ``_lerp``: virtual index
    (n-1)*q, neighbours from the sorted data, a + (b-a)*t, or b - (b-a)*(1-t) where t >= 0.5) but without
    the generic reduction machinery (6x less overhead on the 20-2000 element arrays of the re-scaling
    step).  Equality with numpy is asserted bit for bit in tests/test_refine.py."""
    a = np.asarray(a)
    q = np.asarray(q, dtype=np.float64)
    n = a.size
    if n == 0 or a.ndim != 1 or a.dtype.kind != "f":
        return np.quantile(a, q)
    srt = np.sort(a)
    if np.isnan(srt[-1]):
        return np.full(q.shape, np.nan)
    vi = (n - 1) * q
    prev = np.floor(vi).astype(np.intp)
    nxt = prev + 1
    above = vi >= n - 1
    prev[above] = n - 1
    nxt[above] = n - 1
    below = vi < 0
    prev[below] = 0
    nxt[below] = 0
    lo, hi = srt[prev], srt[nxt]
    t = vi - prev
    t[above] = vi[above] - (-1)  # numpy takes gamma from the (wrapped) index -1 there; lo == hi anyway
    diff = np.subtract(hi, lo)
    out = np.asanyarray(np.add(lo, diff * t))
    np.subtract(hi, diff * (1 - t), out=out, where=t >= 0.5, casting="unsafe", dtype=type(out.dtype))
    return out


# The re-scaling estimators all do the same thing: fit a line  level ~ b0 + b1 * x  through points
# (x = currently-normalised signal summary, level = expected k-mer level) and fold it into the read's
# (shift, scale).  What differs is (a) which points (quantiles of the whole read for the rough pass,
# per-base means for the refinement rounds) and (b) the line estimator.  The two estimators fold the fit
# in with algebraically equal but differently rounded expressions in the reference
# (refine_signal_map.py:68-80 vs :95-103); both forms are kept verbatim in `_fold_*` because the
# resulting shift / scale are compared bit for bit with the reference's (tests/test_refine.py).
def _line_least_squares(x, y):
    """(b0, b1) minimising |b0 + b1 x - y|^2 through numpy's LAPACK driver - the same call the
    reference makes, so the estimates agree to the last bit."""
    design = np.column_stack([np.ones_like(x), x])
    b0, b1 = np.linalg.lstsq(design, y, rcond=None)[0]
    return b0, b1


def _line_theil_sen(x, y):
    """Median of the pairwise slopes over pairs with increasing x, then the median residual."""
    dx = x[:, np.newaxis] - x
    dy = y[:, np.newaxis] - y
    rising = dx > 0
    b1 = np.median(dy[rising] / dx[rising])
    b0 = np.median(y - (b1 * x))
    return b0, b1


def _fold_least_squares(shift, scale, b0, b1):
    if b1 == 0:  # degenerate fit: keep the read's scaling
        return shift, scale
    return shift - (scale * b0 / b1), scale / b1


def _fold_theil_sen(shift, scale, b0, b1):
    if b1 == 0:
        raise RemoraError("Read failed sequence-based signal re-scaling parameter estimation.")
    return shift + (-b0 / b1 * scale), scale * (1 / b1)


_ESTIMATORS = {
    ROUGH_RESCALE_LEAST_SQUARES: (_line_least_squares, _fold_least_squares),
    ROUGH_RESCALE_THEIL_SEN: (_line_theil_sen, _fold_theil_sen),
}


def recalibrate(points, levels, shift, scale, method, quantiles=None, max_points=None):
    """New (shift, scale) from DAC-domain ``points`` and the ``levels`` they should sit on.
    ``quantiles``: compare the two distributions at these quantiles instead of point by point (rough
    pass, refine_signal_map.py:82-93, 115-122); ``max_points``: random sub-sample of the pairs before a
    Theil-Sen fit, whose cost is quadratic (:105-113)."""
    try:
        line, fold = _ESTIMATORS[method]
    except KeyError:
        raise RemoraError(f"Invalid rough re-scale method: {method}")
    x = (points - shift) / scale
    y = levels
    if quantiles is not None:
        x, y = quantile_linear(x, quantiles), quantile_linear(y, quantiles)
    elif max_points is not None and y.shape[0] > max_points:
        pick = np.random.choice(y.shape[0], max_points, replace=False)
        x, y = x[pick], y[pick]
    return fold(shift, scale, *line(x, y))


def index_from_kmer(kmer, alphabet="ACGT"):
    """refine_signal_map.py:130-146"""
    return sum(alphabet.find(base) * (len(alphabet) ** kmer_pos)
               for kmer_pos, base in enumerate(kmer[::-1]))


# ------------------------------------------------------------------------------------------------
# banding (host, integer numpy; same integers as the reference)
# ------------------------------------------------------------------------------------------------
def base_space_band(seq_to_sig_map, levels, hbw=DEFAULT_REFINE_HBW):
    """Seq band of a read, int32 [2, n_bases]: for every base the half-open range of signal samples it
    may be moved over, before ``adjust_seq_band``.

    Definition (what refine_signal_map.py:634-688 + :743-775 compute through two per-SAMPLE arrays): a
    sample that currently belongs to base b may be re-assigned to bases [b - hbw, b + hbw] (clipped to
    the read); a base without a level (NaN: its k-mer holds an N) keeps exactly its own samples; both
    bounds are made monotone along the signal (running max of the lower, running min from the right of
    the upper bound); base q may then cover the samples whose bounds contain q.  Only bases that own
    at least one sample take part, and the bounds are constant over a base's dwell - so everything is
    computed per BASE here (O(bases log bases), no per-sample array), by two binary searches over the
    monotone bounds.  The integers equal the sample-space construction's (randomised test against the
    oracle's restatement of it, tests/test_refine.py)."""
    n = levels.shape[0]
    rel = np.asarray(seq_to_sig_map, dtype=np.int64)
    if rel.size - 1 != n:
        raise RemoraError("Breakpoints must be one longer than levels.")
    if hbw is None:
        raise RemoraError("Cannot compute band with half width of None.")
    rel = rel - rel[0]
    sig_len = int(rel[-1])
    base = np.arange(n)
    unknown = np.isnan(levels)
    lower = np.where(unknown, base, np.maximum(base - hbw, 0))
    upper = np.where(unknown, base + 1, np.minimum(base + hbw + 1, n))
    owns = np.diff(rel) > 0
    if not owns.any():
        raise RemoraError("Band contains 0-length region")
    first_sample = rel[:-1][owns]
    lower = np.maximum.accumulate(lower[owns])
    upper = np.minimum.accumulate(upper[owns][::-1])[::-1]
    n_q = int(upper[-1])  # bases reachable from the last sample
    q = np.arange(n_q)
    band = np.empty((2, n_q), dtype=np.int32)
    # start: first sample whose upper bound already exceeds q (upper[-1] = n_q > q: always found)
    band[0] = first_sample[np.searchsorted(upper, q, side="right")]
    # end: first sample AFTER sample 0 at which the lower bound steps above q; none -> end of the signal
    step = np.searchsorted(lower, q, side="right")
    later = np.searchsorted(lower, lower[0], side="right")  # first owner whose bound exceeds the first one's
    step = np.where(step == 0, later, step)
    band[1] = np.where(step < first_sample.size, first_sample[np.minimum(step, first_sample.size - 1)], sig_len)
    return band


def adjust_seq_band(seq_band, min_step=2):
    """In-place, vectorised form of the reference's sequential loops (refine_signal_map_core.pyx:31-69):
    every band start (end) at least ``min_step`` below (above) its successor (predecessor), then the
    first starts / last ends repaired so that they stay strictly increasing from the fixed corner."""
    st, en = seq_band[0], seq_band[1]
    n = st.size
    idx = np.arange(n, dtype=np.int64)
    band_min = int(st[0])
    # st[i] = min_{j>=i}(st[j] - min_step*(j-i))
    st[:] = np.minimum.accumulate((st - min_step * idx)[::-1])[::-1] + min_step * idx
    st[0] = band_min
    if n > 1:
        ok = st[1:] > band_min + idx[:-1]        # first position that already exceeds its predecessor
        stop = int(np.argmax(ok)) + 1 if ok.any() else n
        st[1:stop] = band_min + idx[1:stop]
    band_max = int(en[-1])
    en[:] = np.maximum.accumulate(en - min_step * idx) + min_step * idx
    en[-1] = band_max
    if n > 1:
        # walking down from n-2 the reference lowers en[pos] to en[pos+1] - 1 while en[pos] >= en[pos+1];
        # as long as it keeps lowering, en[pos+1] == band_max - dist with dist = n-2-pos
        dist = np.arange(n - 1, dtype=np.int64)    # 0 for pos n-2, 1 for pos n-3, ...
        ok = en[:-1][::-1] < band_max - dist       # first position already below its successor
        stop = int(np.argmax(ok)) if ok.any() else n - 1
        if stop > 0:
            en[n - 2 - np.arange(stop)] = band_max - 1 - np.arange(stop)
    return seq_band


def validate_band(band, sig_len=None, seq_len=None, is_sig_band=True):
    """Raise ``RemoraError`` (the reference's messages, refine_signal_map.py:691-740) unless ``band``
    [2, n] starts at 0, has no empty interval, is monotone in both rows and ends where the other
    coordinate system says it should."""
    starts, ends = band[0], band[1]
    length_of, end_of = ("sig", "seq") if is_sig_band else ("seq", "sig")
    want = {"sig": sig_len, "seq": seq_len}
    checks = [
        (starts[0] != 0, "Band does not start with 0 coordinate."),
        ((ends - starts).min() <= 0, "Band contains 0-length region"),
        (starts.size > 1 and np.diff(starts).min() < 0, "Band start positions are not monotonically increasing"),
        (starts.size > 1 and np.diff(ends).min() < 0, "Band end positions are not monotonically increasing"),
        (want[length_of] is not None and starts.size != want[length_of], "Invalid sig_band length"),
        (want[end_of] is not None and ends[-1] != want[end_of],
         f"Invalid {'sig' if is_sig_band else 'seq'}_band end coordinate"),
    ]
    if not is_sig_band:  # the reference tests the end coordinate before the length for a seq band
        checks[4], checks[5] = checks[5], checks[4]
    for failed, message in checks:
        if failed:
            raise RemoraError(message)


def _check_dp_preconditions(seq_band):
    """What the device programme relies on beyond ``validate_band`` (for the reference these cases
    are out-of-bounds reads): strictly increasing starts and ends, every band reachable from the
    previous one."""
    st, en = seq_band[0], seq_band[1]
    if st.size > 1 and (np.diff(st).min() < 1 or np.diff(en).min() < 1 or (st[1:] > en[:-1]).any()):
        raise RemoraError("Invalid band for signal mapping refinement")


def compute_seq_band(seq_to_sig_map, levels, band_half_width=DEFAULT_REFINE_HBW, adjust_band_min_step=2):
    """seq band of one read, as ``refine_signal_mapping`` builds it (refine_signal_map.py:814-826):
    band in base space, minimum-step adjustment, validation."""
    seq_to_sig_map = np.asarray(seq_to_sig_map)
    seq_band = base_space_band(seq_to_sig_map, levels, band_half_width)
    if seq_band.shape[1] != levels.shape[0]:
        raise RemoraError("Invalid sig_band length")
    seq_band = adjust_seq_band(seq_band, min_step=adjust_band_min_step)
    validate_band(seq_band, sig_len=int(seq_to_sig_map[-1] - seq_to_sig_map[0]), seq_len=levels.shape[0],
                  is_sig_band=False)
    _check_dp_preconditions(seq_band)
    return seq_band


# ------------------------------------------------------------------------------------------------
# the device programme
# ------------------------------------------------------------------------------------------------
def _dacs_code(dacs, shift, scale):
    """dtype code of rb200_refine_normalize reproducing numpy's promotion of (dacs - shift) / scale."""
    if dacs.dtype == np.int16:
        return dacs, 0
    if dacs.dtype == np.float32:
        probe = (dacs[:1] - shift) / scale
        return dacs, (1 if probe.dtype == np.float32 else 3)
    return dacs.astype(np.float64), 2


NEAR_CAPS = (128, 192, 256, 384, 512, 768, 1024)


def choose_near_cap(band_widths, coverage=0.98):
    """Capacity of the per-warp shared-memory rows for a batch: the smallest step that holds the bands
    of ``coverage`` of all DP cells.  The few wider bands (stalls) run from global scratch rows, and the
    smaller footprint lets more reads share an SM (4 x 8 warps at <= 256 samples)."""
    w = np.sort(np.asarray(band_widths, dtype=np.int64))
    cum = np.cumsum(w)
    need = int(w[np.searchsorted(cum, coverage * cum[-1])])
    for cap in NEAR_CAPS:
        if cap >= need:
            return cap
    return NEAR_CAPS[-1]


class DeviceRefineBatch:
    """A batch of reads resident on the GPU for the banded dynamic programme: ``__init__`` builds the
    concatenated arrays and uploads them, :meth:`run` enqueues the two kernels
    (``rb200_refine_normalize`` + ``rb200_refine_dp``) on the current stream, :meth:`paths` reads the
    result back.  ``banded_dp_batch`` is the one-shot form; benchmarks time :meth:`run` alone."""

    def __init__(self, dacs_list, shifts, scales, levels_list, seq_bands, algo=DEFAULT_REFINE_ALGO,
                 short_dwell_pen=DEFAULT_REFINE_SHORT_DWELL_PEN, device=None, near_cap=None):
        if algo not in _ALGO_CODE:
            raise RemoraError(f"Invalid core signal mapping refine method: {algo}")
        self.algo = _ALGO_CODE[algo]
        self.pen = np.ascontiguousarray(short_dwell_pen, dtype=np.float32)
        if algo == REFINE_ALGO_DWELL_PEN_NAME and not 1 <= self.pen.size <= MAX_DWELL_PENALTIES:
            raise RemoraError(f"short dwell penalty array must hold 1..{MAX_DWELL_PENALTIES} values")
        self.lib = _native.load_library()
        if device is None:
            if not torch.cuda.is_available():
                raise RemoraError("signal mapping refinement needs a CUDA device (no CPU fallback)")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device = torch.device(device)
        if device.type != "cuda":
            raise RemoraError("signal mapping refinement needs a CUDA device (no CPU fallback)")
        self.n_reads = n_reads = len(dacs_list)
        codes = set()
        conv = []
        for d, sh, sc in zip(dacs_list, shifts, scales):
            d, code = _dacs_code(np.ascontiguousarray(d), sh, sc)
            conv.append(d)
            codes.add(code)
        if len(codes) > 1:  # mixed sample dtypes: float64 reproduces every promotion except pure float32
            if 1 in codes:
                raise RemoraError("mixed float32 / other sample dtypes in one refinement batch")
            conv = [d.astype(np.float64) for d in conv]
            codes = {2}
        self.code = codes.pop()
        sig_lens = np.array([d.size for d in conv], dtype=np.int64)
        seq_lens = np.array([lv.size for lv in levels_list], dtype=np.int64)
        self.sig_off = sig_off = np.concatenate([[0], np.cumsum(sig_lens)]).astype(np.int64)
        self.seq_off = seq_off = np.concatenate([[0], np.cumsum(seq_lens)]).astype(np.int64)
        widths = [b[1] - b[0] for b in seq_bands]
        tb_lens = np.array([int(w.sum()) for w in widths], dtype=np.int64)
        if tb_lens.max() > np.iinfo(np.uint32).max:
            raise RemoraError("Dynamic programming search space too large. Read likely contains large "
                              "deletions.")
        self.tb_off = tb_off = np.concatenate([[0], np.cumsum(tb_lens)]).astype(np.int64)
        order = np.argsort(-tb_lens, kind="stable").astype(np.int32)
        for r in range(n_reads):
            if seq_bands[r].shape[1] != seq_lens[r] or seq_bands[r][1, -1] != sig_lens[r]:
                raise RemoraError("band does not match read")
        self.max_sig_len = int(sig_lens.max())
        all_w = np.concatenate(widths)
        self.widest = int(all_w.max())
        self.cells = int(tb_off[-1])
        self.near_cap = near_cap if near_cap else choose_near_cap(all_w)
        with torch.cuda.device(device):
            up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=True)  # noqa: E731
            self.d_dacs = up(np.concatenate(conv))
            self.d_sig_off, self.d_seq_off, self.d_tb_off = up(sig_off), up(seq_off), up(tb_off)
            self.d_shift = up(np.asarray(shifts, dtype=np.float64))
            self.d_scale = up(np.asarray(scales, dtype=np.float64))
            self.d_levels = up(np.concatenate([np.asarray(lv, dtype=np.float32) for lv in levels_list]))
            self.d_st = up(np.concatenate([b[0] for b in seq_bands]).astype(np.int32))
            self.d_en = up(np.concatenate([b[1] for b in seq_bands]).astype(np.int32))
            self.d_order = up(order)
            self.d_sig = torch.empty(int(sig_off[-1]), dtype=torch.float32, device=device)
            self.d_tb = torch.empty(int(tb_off[-1]), dtype=torch.int32, device=device)
            self.d_path = torch.empty(int(seq_off[-1]) + n_reads, dtype=torch.int32, device=device)
            self.d_score = torch.empty(n_reads, dtype=torch.float32, device=device)
            self.d_status = torch.empty(n_reads, dtype=torch.int32, device=device)
            self.d_queue = torch.zeros(1, dtype=torch.int32, device=device)
            need = ctypes.c_int64(0)
            _native.check(self.lib.rb200_refine_scratch_bytes(self.near_cap, self.widest, ctypes.byref(need)),
                          "rb200_refine_scratch_bytes")
            self.d_wide = (torch.empty(need.value // 4, dtype=torch.float32, device=device)
                           if need.value else None)

    def run(self):
        """Enqueue normalisation + banded DP on the current stream of the batch's device."""
        ptr = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
        with torch.cuda.device(self.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _native.check(self.lib.rb200_refine_normalize(
                ptr(self.d_dacs), self.code, ptr(self.d_sig_off), ptr(self.d_shift), ptr(self.d_scale),
                self.n_reads, self.max_sig_len, ptr(self.d_sig), stream), "rb200_refine_normalize")
            _native.check(self.lib.rb200_refine_dp(
                ptr(self.d_sig), ptr(self.d_sig_off), ptr(self.d_levels), ptr(self.d_st), ptr(self.d_en),
                ptr(self.d_seq_off), ptr(self.d_tb_off), ptr(self.d_order), self.n_reads,
                self.pen.ctypes.data_as(ctypes.c_void_p), self.pen.size, self.algo, self.near_cap, self.widest,
                ptr(self.d_tb), ptr(self.d_path), ptr(self.d_score), ptr(self.d_status), ptr(self.d_queue),
                ptr(self.d_wide) if self.d_wide is not None else None, stream), "rb200_refine_dp")

    def paths(self, return_scores=False):
        path = self.d_path.cpu().numpy()
        status = self.d_status.cpu().numpy()
        out = []
        for r in range(self.n_reads):
            if status[r] != 0:
                raise RemoraError("signal mapping refinement traceback left the band")
            out.append(path[self.seq_off[r] + r:self.seq_off[r + 1] + r + 1].copy())
        return (out, self.d_score.cpu().numpy()) if return_scores else out


MAX_CELLS_PER_LAUNCH = 1 << 30  # 4 GiB of traceback workspace per launch


def split_by_budget(sizes, budget):
    """Consecutive index ranges [(start, end)] whose summed ``sizes`` stay within ``budget`` (a single
    oversized item gets a range of its own)."""
    spans, start, total = [], 0, 0
    for i, sz in enumerate(sizes):
        if i > start and total + sz > budget:
            spans.append((start, i))
            start, total = i, 0
        total += sz
    if start < len(sizes):
        spans.append((start, len(sizes)))
    return spans


def banded_dp_batch(dacs_list, shifts, scales, levels_list, seq_bands, algo=DEFAULT_REFINE_ALGO,
                    short_dwell_pen=DEFAULT_REFINE_SHORT_DWELL_PEN, device=None, return_scores=False,
                    max_cells=None):
    """Run the banded dynamic programme for a batch of reads on the GPU.

    dacs_list[r]: un-normalised samples of read r already trimmed to its mapped range;
    levels_list[r]: float32 levels (NaN allowed); seq_bands[r]: int32 [2, seq_len] from
    :func:`compute_seq_band`.  Returns the list of paths (int32 [seq_len + 1], path[0] == 0)
    (and the final forward score per read when ``return_scores``)."""
    if len(dacs_list) == 0:
        return ([], np.zeros(0, np.float32)) if return_scores else []
    # the traceback workspace is 4 bytes per band cell: long reads are handled in several launches
    cells = [int((b[1] - b[0]).sum()) for b in seq_bands]
    paths, scores = [], []
    for st, en in split_by_budget(cells, max_cells or MAX_CELLS_PER_LAUNCH):
        batch = DeviceRefineBatch(dacs_list[st:en], shifts[st:en], scales[st:en], levels_list[st:en],
                                  seq_bands[st:en], algo, short_dwell_pen, device)
        batch.run()
        out = batch.paths(return_scores)
        if return_scores:
            paths.extend(out[0])
            scores.append(out[1])
        else:
            paths.extend(out)
    return (paths, np.concatenate(scores)) if return_scores else paths


def refine_signal_mapping(signal, seq_to_sig_map, levels, band_half_width=DEFAULT_REFINE_HBW,
                          refine_algo=DEFAULT_REFINE_ALGO, short_dwell_pen=DEFAULT_REFINE_SHORT_DWELL_PEN,
                          adjust_band_min_step=2, device=None):
    """Single-read form with the reference's arguments (refine_signal_map.py:783-840); ``signal`` is the
    normalised signal.  Returns (path, final score, seq_band); the reference additionally returns the
    full score / traceback arrays, which stay on the device here."""
    seq_to_sig_map = np.asarray(seq_to_sig_map)
    start = int(seq_to_sig_map[0])
    signal = np.asarray(signal)[start:int(seq_to_sig_map[-1])]
    levels = np.asarray(levels, dtype=np.float32)
    seq_band = compute_seq_band(seq_to_sig_map - start, levels, band_half_width, adjust_band_min_step)
    sig32 = np.ascontiguousarray(signal, dtype=np.float32)  # the reference's .astype(np.float32)
    paths, scores = banded_dp_batch([sig32], [0.0], [1.0], [levels], [seq_band], refine_algo,
                                    short_dwell_pen, device, return_scores=True)
    return paths[0] + start, float(scores[0]), seq_band


# ------------------------------------------------------------------------------------------------
# SigMapRefiner
# ------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class SigMapRefiner:
    """Same fields as the reference dataclass (refine_signal_map.py:149-176) plus ``device``."""

    kmer_model_filename: str = None
    do_rough_rescale: bool = False
    scale_iters: int = -1
    algo: str = DEFAULT_REFINE_ALGO
    half_bandwidth: int = DEFAULT_REFINE_HBW
    sd_params: tuple = None
    do_fix_guage: bool = False
    rough_rescale_method: str = DEFAULT_ROUGH_RESCALE_METHOD
    sd_arr: np.ndarray = dataclasses.field(default_factory=lambda: DEFAULT_REFINE_SHORT_DWELL_PEN)
    _levels_array: np.ndarray = None
    str_kmer_levels: dict = None
    kmer_len: int = None
    kmer_idx_stats = None
    center_idx: int = -1
    is_loaded: bool = False
    device: object = None

    def __post_init__(self):
        """Decide where the level table comes from - an array (model metadata), a k-mer level file, or a
        {k-mer: level} dict - derive the k-mer length / dominant position from it, then validate the options
        the way the reference does (refine_signal_map.py:276-322: same error messages)."""
        have_array = self._levels_array is not None and np.ndim(self._levels_array) > 0
        if have_array:
            table = np.ascontiguousarray(self._levels_array, dtype=np.float32)
            k = int(round(np.log2(table.size) / 2)) if table.size else 0
            if table.size == 0 or 4 ** k != table.size:
                raise RemoraError("levels array size is not a power of 4")
            self._levels_array, self.kmer_len, self.is_loaded = table, k, True
        else:
            self._levels_array = None
            if self.kmer_model_filename is not None:
                self.load_kmer_table()
            elif self.str_kmer_levels is not None:
                self.kmer_len = len(next(iter(self.str_kmer_levels)))
            if self.str_kmer_levels is not None:
                self.is_loaded = True
                self.determine_dominant_pos()
                if self.do_fix_guage:
                    self.fix_gauge()
        wants_refinement = self.do_rough_rescale or self.scale_iters >= 0
        if wants_refinement and not self.is_loaded:
            raise RemoraError("Signal re-scaling is requested without levels table. "
                              f"is_loaded: {self.is_loaded} do_rough_rescale: {self.do_rough_rescale} "
                              f"scale_iters: {self.scale_iters}")
        if self.sd_params is not None:
            self.sd_arr = compute_dwell_pen_array(*self.sd_params)
        if self.rough_rescale_method not in ROUGH_RESCALE_METHODS:
            raise RemoraError(f"Invalid rough re-scale method: {self.rough_rescale_method}")

    def __repr__(self):
        if not self.is_loaded:
            return "No Remora signal refine/map settings loaded"
        parts = [f"Loaded {self.kmer_len}-mer table with {self.center_idx + 1} central position."]
        if self.do_rough_rescale:
            parts.append("Rough re-scaling will be executed.")
        if self.scale_iters > 0:
            parts.append(f"{self.scale_iters} rounds of signal mapping refinement followed by precise "
                         "re-scaling will be executed.")
        if self.scale_iters >= 0:
            parts.append("Signal mapping refinement will be executed using the "
                         f"{self.algo} refinement method (band half width: {self.half_bandwidth}).")
            if self.algo == REFINE_ALGO_DWELL_PEN_NAME:
                parts.append(f"Short dwell penalty array set to {self.sd_arr}.")
        return " ".join(parts)

    @property
    def bases_before(self):
        return self.center_idx

    @property
    def bases_after(self):
        return self.kmer_len - 1 - self.center_idx

    @property
    def is_valid(self):
        """A loaded table must be used by at least one step; no table means no step may ask for one."""
        return (self.do_rough_rescale or self.scale_iters >= 0) == bool(self.is_loaded)

    # -- level table ---------------------------------------------------------------------------
    def load_kmer_table(self):
        """Whitespace-separated ``kmer level`` lines (reference format, refine_signal_map.py:217-247):
        every k-mer of one length exactly once; NaN levels become 0."""
        levels = {}
        with open(self.kmer_model_filename) as fh:
            rows = [ln.split() for ln in fh if ln.strip()]
        k = len(rows[0][0]) if rows else 0
        for kmer, text in rows:
            kmer = kmer.upper()
            if kmer in levels:
                raise RemoraError(f"K-mer found twice in levels file '{kmer}'.")
            if len(kmer) != k:
                raise RemoraError(f"K-mer lengths not all equal '{len(kmer)} != {k}' for {kmer}.")
            try:
                value = float(text)
            except ValueError:
                raise RemoraError(f"Could not convert level to float '{text}'") from None
            levels[kmer] = 0 if value != value else value
        if len(levels) != 4 ** k:
            raise RemoraError(f"K-mer table contains fewer entries ({len(levels)}) than expected ({4 ** k})")
        self.kmer_len, self.str_kmer_levels = k, levels

    def determine_dominant_pos(self):
        """The k-mer position whose base says most about the level: per position, the Kruskal-Wallis H
        statistic of the four bases' rank groups in the level-sorted k-mer list (refine_signal_map.py:249-274);
        the largest one becomes ``center_idx``."""
        if self.str_kmer_levels is None:
            return
        from scipy import stats
        by_level = sorted(self.str_kmer_levels, key=lambda km: (self.str_kmer_levels[km], km))
        letters = np.array([list(km) for km in by_level])          # [n_kmers, k], rows in rank order
        ranks = np.arange(len(by_level))
        self.kmer_idx_stats = [stats.kruskal(*(ranks[letters[:, pos] == base] for base in "ACGT"))[0]
                               for pos in range(self.kmer_len)]
        self.center_idx = int(np.argmax(self.kmer_idx_stats))

    def fix_gauge(self):
        """Median / MAD standardisation of the table (MAD scaled to a standard deviation,
        refine_signal_map.py:332-342); the dict view is rebuilt from the array."""
        table = self.levels_array
        centre = np.median(table)
        spread = 1.4826 * np.median(np.abs(table - centre))
        self._levels_array = (table - centre) / spread
        self.str_kmer_levels = {km: self._levels_array[index_from_kmer(km)] for km in self.kmers}

    @property
    def levels_array(self):
        """float32 [4^k] in base-4 k-mer order, built from the dict on first use."""
        if self._levels_array is None and self.str_kmer_levels is not None:
            table = np.empty(4 ** self.kmer_len, dtype=np.float32)
            for km, lv in self.str_kmer_levels.items():
                table[index_from_kmer(km)] = lv
            self._levels_array = table
        return self._levels_array

    @property
    def kmers(self):
        import itertools
        return ("".join(t) for t in itertools.product("ACGT", repeat=self.kmer_len))

    def write_kmer_table(self, fh):
        for kmer in self.kmers:
            fh.write(f"{kmer}\t{self._levels_array[index_from_kmer(kmer)]}\n")

    def extract_levels(self, int_seq):
        """levels[pos + center_idx] = table[k-mer starting at pos], zero at the read edges
        (refine_signal_map_core.pyx:87-100), vectorised.  A k-mer containing N (-1) has no defined
        level in the reference (out-of-range table index); it gets NaN here, which the banding
        treats as "keep this base's samples" (refine_signal_map.py:673-680)."""
        int_seq = np.asarray(int_seq).astype(np.int64)
        table = self.levels_array
        levels = np.zeros(int_seq.size, dtype=np.float32)
        k = self.kmer_len
        if int_seq.size >= k:
            windows = np.lib.stride_tricks.sliding_window_view(int_seq, k)
            idx = windows @ (4 ** np.arange(k - 1, -1, -1, dtype=np.int64))
            has_n = (windows < 0).any(axis=1)
            vals = table[np.where(has_n, 0, idx)]
            levels[self.center_idx:self.center_idx + idx.size] = np.where(has_n, np.float32(np.nan), vals)
        return levels

    # -- re-scaling ----------------------------------------------------------------------------
    def rough_rescale(self, shift, scale, seq_to_sig_map, int_seq, dacs,
                      quants=np.arange(0.05, 1, 0.05), clip_bases=10, use_base_center=True, levels=None):
        """Quantile matching of the read against its expected levels (reference signature,
        refine_signal_map.py:390-430; ``levels``: optional pre-computed ``extract_levels(int_seq)``).
        One sample per base (the one in the middle of its dwell) with ``clip_bases`` dropped at both ends,
        or every mapped sample when ``use_base_center`` is off."""
        if levels is None:
            levels = self.extract_levels(int_seq)
        if not use_base_center:
            picked = dacs[seq_to_sig_map[0]:seq_to_sig_map[-1]]
        else:
            picked = dacs[(seq_to_sig_map[:-1] + seq_to_sig_map[1:]) // 2]
            if clip_bases > 0 and levels.size > 2 * clip_bases:
                inner = slice(clip_bases, -clip_bases)
                picked, levels = picked[inner], levels[inner]
        return recalibrate(picked, levels, shift, scale, self.rough_rescale_method, quantiles=quants)

    @staticmethod
    def _base_means(dacs, seq_to_sig_map):
        """Mean DAC value of every base's dwell from one running sum (NaN for zero-dwell bases)."""
        running = np.empty(dacs.size + 1)
        running[0] = 0
        running[1:] = np.cumsum(dacs)
        dwells = np.diff(seq_to_sig_map)
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.diff(running[seq_to_sig_map]) / dwells, dwells

    def rescale(self, levels, dacs, shift, scale, seq_to_sig_map, dwell_filter_pctls=(10, 90),
                min_abs_level=0.2, edge_filter_bases=10, min_levels=10):
        """Theil-Sen re-fit on per-base means after a refinement round (reference signature,
        refine_signal_map.py:432-472).  Bases are used when their dwell lies strictly inside the given
        dwell percentiles, their level is informative (away from the mean level), they have samples,
        and they are not within ``edge_filter_bases`` of either read end."""
        means, dwells = self._base_means(dacs, seq_to_sig_map)
        shortest, longest = np.percentile(dwells, dwell_filter_pctls)
        keep = (dwells > shortest) & (dwells < longest)
        keep &= np.abs(levels - np.mean(levels)) > min_abs_level
        keep &= ~np.isnan(means)
        if edge_filter_bases > 0:
            keep[:edge_filter_bases] = False
            keep[-edge_filter_bases:] = False
        if np.count_nonzero(keep) < min_levels:
            raise RemoraError("Too few positions")
        return recalibrate(means[keep], levels[keep], shift, scale, ROUGH_RESCALE_THEIL_SEN,
                           max_points=MAX_POINTS_FOR_THEIL_SEN)

    # -- mapping refinement ----------------------------------------------------------------------
    def refine_sig_maps(self, shifts, scales, seq_to_sig_maps, int_seqs, dacs_list, levels=None, errors=None):
        """Batch form of :meth:`refine_sig_map`: one banded-DP launch per refinement round for all
        reads.  Returns lists (seq_to_sig_map, shift, scale) per read.  A read whose band is invalid
        raises ``RemoraError`` like the reference's ``validate_band`` - unless ``errors`` (a list, one
        slot per read) is given: then the message is stored there, the read keeps its mapping and the
        rest of the batch goes on."""
        n = len(dacs_list)
        if levels is None:
            levels = [self.extract_levels(s) for s in int_seqs]
        maps = [np.asarray(m) for m in seq_to_sig_maps]
        sig_st = [int(m[0]) for m in maps]
        trimmed = [np.asarray(d)[m[0]:m[-1]] for d, m in zip(dacs_list, maps)]
        rel = [m - st for m, st in zip(maps, sig_st)]
        shifts, scales = list(shifts), list(scales)
        active = list(range(n))
        for _ in range(max(1, self.scale_iters)):  # 0 = one round without re-scaling (:486-487)
            if not active:
                break
            bands, ok = [], []
            for r in active:
                try:
                    bands.append(compute_seq_band(rel[r], levels[r], self.half_bandwidth))
                    ok.append(r)
                except RemoraError as e:
                    if errors is None:
                        raise
                    errors[r] = str(e)
            active = ok
            if not active:
                break
            paths = banded_dp_batch([trimmed[r] for r in active], [shifts[r] for r in active],
                                    [scales[r] for r in active], [levels[r] for r in active], bands,
                                    self.algo, self.sd_arr, self.device)
            for r, path in zip(active, paths):
                rel[r] = path
            if self.scale_iters > 0:
                still = []
                for r in active:
                    try:
                        shifts[r], scales[r] = self.rescale(levels[r], trimmed[r], shifts[r], scales[r],
                                                            rel[r])
                        still.append(r)
                    except RemoraError:  # the reference stops refining this read (:498-500)
                        pass
                active = still
        return [m + st for m, st in zip(rel, sig_st)], shifts, scales

    def refine_sig_map(self, shift, scale, seq_to_sig_map, int_seq, dacs):
        """Reference signature (refine_signal_map.py:474-502); the dynamic programme runs on the GPU."""
        maps, shifts, scales = self.refine_sig_maps([shift], [scale], [seq_to_sig_map], [int_seq], [dacs])
        return maps[0], shifts[0], scales[0]

    def refine_reads(self, reads):
        """Apply ``RemoraRead.refine_signal_mapping`` (reference data_chunks.py:267-306) to a whole list
        of reads with ONE banded-DP launch per refinement round: rough re-scaling per read on the host,
        then the batch on the GPU.  Reads are updated in place (shift, scale, seq_to_sig_map).
        Returns one entry per read: ``None``, or the message of the error that read alone would have
        raised (index error in the re-scaling of a read whose last base has no sample, invalid band ...);
        such a read is left as it was and does not stop the batch."""
        errors = [None] * len(reads)
        if not self.is_loaded or not reads:
            return errors
        levels = [self.extract_levels(read.int_seq) for read in reads]  # once per read, both steps use them
        if self.do_rough_rescale:
            for i, (read, lv) in enumerate(zip(reads, levels)):
                try:
                    read.shift, read.scale = self.rough_rescale(read.shift, read.scale, read.seq_to_sig_map,
                                                                read.int_seq, read.dacs, levels=lv)
                    read._sig = None
                except (RemoraError, IndexError, ValueError, np.linalg.LinAlgError) as e:
                    errors[i] = f"rough re-scaling failed: {e}"
        if self.scale_iters >= 0:
            live = [i for i in range(len(reads)) if errors[i] is None]
            sub_err = [None] * len(live)
            maps, shifts, scales = self.refine_sig_maps(
                [reads[i].shift for i in live], [reads[i].scale for i in live],
                [reads[i].seq_to_sig_map for i in live], [reads[i].int_seq for i in live],
                [reads[i].dacs for i in live], levels=[levels[i] for i in live], errors=sub_err)
            for k, i in enumerate(live):
                if sub_err[k] is not None:
                    errors[i] = sub_err[k]
                    continue
                reads[i].seq_to_sig_map, reads[i].shift, reads[i].scale = maps[k], shifts[k], scales[k]
                reads[i]._sig = None
        return errors

    # -- (de)serialisation -------------------------------------------------------------------------
    # model metadata key -> field (the key names are the wire format of meta.txt, model_util.py:115-176)
    _METADATA_FIELDS = (
        ("refine_kmer_levels", "_levels_array"), ("refine_kmer_center_idx", "center_idx"),
        ("refine_do_rough_rescale", "do_rough_rescale"), ("refine_scale_iters", "scale_iters"),
        ("refine_algo", "algo"), ("refine_half_bandwidth", "half_bandwidth"), ("refine_sd_arr", "sd_arr"),
        ("rough_rescale_method", "rough_rescale_method"),
    )

    def asdict(self):
        return {key: getattr(self, field) for key, field in self._METADATA_FIELDS}

    @classmethod
    def load_from_metadata(cls, metadata, device=None):
        fields = {field: metadata.get(key) for key, field in cls._METADATA_FIELDS}
        if fields["rough_rescale_method"] is None:  # models exported before the key existed
            fields["rough_rescale_method"] = ROUGH_RESCALE_LEAST_SQUARES
        return cls(device=device, **fields)

    @classmethod
    def load_from_dict(cls, data, do_rough_rescale=True, scale_iters=-1, algo=DEFAULT_REFINE_ALGO,
                       half_bandwidth=DEFAULT_REFINE_HBW, sd_params=None, do_fix_guage=False,
                       sd_arr=DEFAULT_REFINE_SHORT_DWELL_PEN,
                       rough_rescale_method=DEFAULT_ROUGH_RESCALE_METHOD, device=None):
        """A refiner from a ``{kmer: level}`` dictionary (reference keyword surface)."""
        options = dict(locals())
        del options["cls"], options["data"]
        return cls(str_kmer_levels=data, **options)

    def _behaviour_key(self):
        """Everything that decides what this refiner does to a read - and nothing else: settings of a
        stage that is switched off do not count (same notion of equality as refine_signal_map.py:553-580)."""
        def raw(arr):
            return None if arr is None else np.asarray(arr).tobytes()
        key = [self.do_rough_rescale, self.scale_iters]
        if self.do_rough_rescale or self.scale_iters >= 0:
            key += [self.rough_rescale_method, self.center_idx, raw(self._levels_array),
                    None if self._levels_array is None else np.asarray(self._levels_array).shape]
        if self.scale_iters >= 0:
            key += [self.algo, self.half_bandwidth, raw(self.sd_arr)]
        return key

    def __eq__(self, other):
        return isinstance(other, SigMapRefiner) and self._behaviour_key() == other._behaviour_key()
