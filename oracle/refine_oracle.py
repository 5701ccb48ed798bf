"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's signal-mapping refinement, used only as
the checker by tests/, __graft_entry__.smoke() and bench.py's CPU legs.  remora_b200/ never imports
this module.

  numpy part   compute_sig_band          refine_signal_map.py:634-688
               convert_to_seq_band       refine_signal_map.py:743-775
               refine_signal_mapping     refine_signal_map.py:783-840
               rough_rescale (lstsq)     refine_signal_map.py:67-81, 390-430
               SigMapRefiner.refine_sig_map with scale_iters <= 0   refine_signal_map.py:474-499
  C part       oracle_refine.c (adjust_seq_band, extract_levels, banded DP + traceback)

Parity status: PINNED against the reference's own refine_signal_map / refine_signal_map_core run in
the build container (tests/test_oracle.py when /root/reference is present) and the committed vectors
tests/golden/refine_cases.npz (tests/golden/make_golden_refine.py).
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ALGO_VITERBI, ALGO_DWELL_PENALTY = 0, 1
ALGO_CODES = {"Viterbi": ALGO_VITERBI, "dwell_penalty": ALGO_DWELL_PENALTY}
DEFAULT_SD_ARR = (0.5 * np.square(np.arange(3, dtype=np.float32) - 4)).astype(np.float32)

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        from build_ref import build_liboracle
        lib = ctypes.CDLL(build_liboracle())
        vp, i32 = ctypes.c_void_p, ctypes.c_int
        lib.oracle_adjust_seq_band.argtypes = [vp, i32, i32]
        lib.oracle_adjust_seq_band.restype = None
        lib.oracle_extract_levels.argtypes = [vp, i32, vp, i32, i32, vp]
        lib.oracle_extract_levels.restype = None
        lib.oracle_seq_banded_dp.argtypes = [vp, vp, vp, i32, vp, i32, i32, vp, vp, vp, vp]
        lib.oracle_seq_banded_dp.restype = i32
        _LIB = lib
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def extract_levels(int_seq, table, kmer_len, center_idx):
    int_seq = np.ascontiguousarray(int_seq, dtype=np.int32)
    table = np.ascontiguousarray(table, dtype=np.float32)
    levels = np.empty(int_seq.size, dtype=np.float32)
    _lib().oracle_extract_levels(_p(int_seq), int_seq.size, _p(table), kmer_len, center_idx, _p(levels))
    return levels


def compute_sig_band(bps, levels, bhw):
    """refine_signal_map.py:634-688 — per signal sample the range of bases it may belong to."""
    seq_len = levels.size
    sig_len = int(bps[-1] - bps[0])
    seq_indices = np.repeat(np.arange(seq_len), np.diff(bps))
    band = np.empty((2, sig_len), dtype=np.int32)
    band[0] = np.maximum(seq_indices - bhw, 0)
    band[1] = np.minimum(seq_indices + bhw + 1, seq_len)
    nan_mask = np.isin(seq_indices, np.nonzero(np.isnan(levels))[0])
    nan_sig = np.where(nan_mask)[0]
    nan_seq = seq_indices[nan_mask]
    band[0, nan_sig] = nan_seq
    band[1, nan_sig] = nan_seq + 1
    band[0] = np.maximum.accumulate(band[0])
    band[1] = np.minimum.accumulate(band[1, ::-1])[::-1]
    return band


def convert_to_seq_band(sig_band):
    """refine_signal_map.py:743-775 — per base the range of signal samples it may cover."""
    sig_len = sig_band.shape[1]
    seq_len = int(sig_band[1, -1])
    seq_band = np.zeros((2, seq_len), dtype=np.int32)
    seq_band[1, :] = sig_len
    lower_sig_pos = np.nonzero(np.ediff1d(sig_band[1], to_begin=0))[0]
    lower_base_pos = sig_band[1, lower_sig_pos - 1]
    seq_band[0, lower_base_pos] = lower_sig_pos
    seq_band[0] = np.maximum.accumulate(seq_band[0])
    upper_sig_pos = np.nonzero(np.ediff1d(sig_band[0], to_begin=0))[0]
    upper_base_pos = sig_band[0, upper_sig_pos]
    seq_band[1, upper_base_pos - 1] = upper_sig_pos
    seq_band[1] = np.minimum.accumulate(seq_band[1, ::-1])[::-1]
    return seq_band


def adjust_seq_band(seq_band, min_step=2):
    seq_band = np.ascontiguousarray(seq_band, dtype=np.int32)
    _lib().oracle_adjust_seq_band(_p(seq_band), seq_band.shape[1], min_step)
    return seq_band


def seq_banded_dp(signal, levels, seq_band, sd_arr, algo):
    """oracle_refine.c restatement of refine_signal_map_core.pyx:403-473; returns
    (all_scores, path, traceback, base_offsets, rc)."""
    signal = np.ascontiguousarray(signal, dtype=np.float32)
    levels = np.ascontiguousarray(levels, dtype=np.float32)
    seq_band = np.ascontiguousarray(seq_band, dtype=np.int32)
    sd_arr = np.ascontiguousarray(sd_arr, dtype=np.float32)
    n = levels.size
    band_len = int((seq_band[1] - seq_band[0]).sum())
    all_scores = np.empty(band_len, dtype=np.float32)
    traceback = np.empty(band_len, dtype=np.int32)
    base_offsets = np.empty(n + 1, dtype=np.uint32)
    path = np.empty(n + 1, dtype=np.int32)
    rc = _lib().oracle_seq_banded_dp(_p(signal), _p(levels), _p(seq_band), n, _p(sd_arr), sd_arr.size,
                                     ALGO_CODES.get(algo, algo), _p(all_scores), _p(traceback),
                                     _p(base_offsets), _p(path))
    return all_scores, path, traceback, base_offsets, rc


def refine_signal_mapping(signal, seq_to_sig_map, levels, band_half_width=5, algo="dwell_penalty",
                          sd_arr=DEFAULT_SD_ARR, min_step=2):
    """refine_signal_map.py:783-840; returns (path, all_scores, traceback, seq_band, base_offsets)."""
    seq_to_sig_map = np.asarray(seq_to_sig_map)
    signal = signal[seq_to_sig_map[0]:seq_to_sig_map[-1]]
    start = int(seq_to_sig_map[0])
    seq_to_sig_map = seq_to_sig_map - start
    seq_band = adjust_seq_band(convert_to_seq_band(compute_sig_band(seq_to_sig_map, levels,
                                                                    band_half_width)), min_step)
    tmp_levels = np.where(np.isnan(levels), 0, levels).astype(np.float32)
    all_scores, path, traceback, base_offsets, _ = seq_banded_dp(
        signal.astype(np.float32), tmp_levels, seq_band, sd_arr, algo)
    return path + start, all_scores, traceback, seq_band, base_offsets


def rough_rescale_lstsq(dacs, levels, shift, scale, quants):
    """refine_signal_map.py:67-81"""
    norm_sig = (dacs - shift) / scale
    norm_qs = np.quantile(norm_sig, quants)
    shift_est, scale_est = np.linalg.lstsq(
        np.column_stack([np.ones_like(norm_qs), norm_qs]), np.quantile(levels, quants), rcond=None)[0]
    if scale_est == 0:
        return shift, scale
    return shift - (scale * shift_est / scale_est), scale / scale_est


def rough_rescale(levels, shift, scale, seq_to_sig_map, dacs, quants=np.arange(0.05, 1, 0.05),
                  clip_bases=10):
    """SigMapRefiner.rough_rescale with use_base_center=True, least squares
    (refine_signal_map.py:390-430)."""
    optim_dacs = dacs[(seq_to_sig_map[:-1] + seq_to_sig_map[1:]) // 2]
    if clip_bases > 0 and levels.size > clip_bases * 2:
        levels = levels[clip_bases:-clip_bases]
        optim_dacs = optim_dacs[clip_bases:-clip_bases]
    return rough_rescale_lstsq(optim_dacs, levels, shift, scale, quants)


def refine_read(dacs, shift, scale, seq_to_sig_map, int_seq, table, kmer_len, center_idx,
                do_rough_rescale=True, band_half_width=5, algo="dwell_penalty", sd_arr=DEFAULT_SD_ARR):
    """RemoraRead.refine_signal_mapping (data_chunks.py:267-306) for a refiner with
    scale_iters == 0: optional rough rescale, then one round of mapping refinement
    (SigMapRefiner.refine_sig_map, refine_signal_map.py:474-499).  Returns (map, shift, scale)."""
    levels = extract_levels(int_seq, table, kmer_len, center_idx)
    if do_rough_rescale:
        shift, scale = rough_rescale(levels, shift, scale, seq_to_sig_map, dacs)
    sig_st = int(seq_to_sig_map[0])
    trimmed = dacs[seq_to_sig_map[0]:seq_to_sig_map[-1]]
    new_map = refine_signal_mapping((trimmed - shift) / scale, seq_to_sig_map - sig_st, levels,
                                    band_half_width, algo, sd_arr)[0]
    return new_map + sig_st, shift, scale
