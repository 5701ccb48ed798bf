"""Model loading behind the reference's ``remora.model_util`` surface.

``load_model`` keeps the reference signature (src/remora/model_util.py:566-578) and returns
``(model, model_metadata)``; the model is a :class:`B200Model`, an ``nn.Module`` whose
``forward(sigs, enc_kmers)`` runs hand-written sm_100a CUDA kernels through the C-ABI library
(include/remora_b200.h) instead of the TorchScript graph, and which additionally offers
``forward_compact`` (fused encode + forward on the reference's compact chunk arrays).
"""
import ctypes
import json
from os.path import isfile

import numpy as np
import torch
from torch import nn

from . import RemoraError, _native, weights

from .refine_signal_map import SigMapRefiner  # noqa: E402,F401  (same name as the reference exports)


def _make_refiner(levels, center_idx, do_rough_rescale, scale_iters, algo, half_bandwidth, sd_arr):
    return SigMapRefiner(_levels_array=levels, center_idx=center_idx,
                         do_rough_rescale=do_rough_rescale, scale_iters=scale_iters, algo=algo,
                         half_bandwidth=half_bandwidth, sd_arr=sd_arr)


def add_derived_metadata(md):
    """Derive the keys downstream code reads (kmer_len, chunk_len, motifs, can_base, mod_long_names,
    sig_map_refiner ...) from the raw ``meta.txt`` dictionary, in place.  Same resulting
    dictionary as the reference builds at model_util.py:341-448, including support for the older
    ``*_0/_1`` and single-``motif`` key styles (model_util.py:362-379)."""
    md.setdefault("reverse_signal", False)
    md.setdefault("pa_scaling", None)
    if md["mod_bases"] == "None":
        md["mod_bases"] = None
        md["mod_long_names"] = None
    else:
        md["mod_long_names"] = [md[f"mod_long_names_{i}"] for i in range(len(md["mod_bases"]))]
    for key in ("kmer_context_bases", "chunk_context"):
        if key not in md:
            md[key] = (int(md[f"{key}_0"]), int(md[f"{key}_1"]))
    md["kmer_len"] = sum(md["kmer_context_bases"]) + 1
    md["chunk_len"] = sum(md["chunk_context"])
    if "num_motifs" in md:
        md["motifs"] = [(md[f"motif_{i}"], int(md[f"motif_offset_{i}"]))
                        for i in range(int(md["num_motifs"]))]
    else:
        md["motifs"] = [(md["motif"], int(md["motif_offset"]))]
        md["motif_offset"] = int(md["motif_offset"])
    first_motif, first_off = md["motifs"][0]
    md["can_base"] = first_motif[first_off]
    md["motif"] = md["motifs"][0] if len(md["motifs"]) == 1 else (md["can_base"], 0)
    if md["mod_bases"] is not None:
        mods = "; ".join(f"{b}={n}" for b, n in zip(md["mod_bases"], md["mod_long_names"]))
        md["alphabet_str"] = f"loaded modified base model to call (alt to {md['can_base']}): {mods}"
    if md.get("refine_kmer_levels") is not None:
        levels = np.frombuffer(md["refine_kmer_levels"].encode("cp437"), dtype=np.float32)
        sd_arr = np.frombuffer(md["refine_sd_arr"].encode("cp437"), dtype=np.float32)
        md["sig_map_refiner"] = _make_refiner(
            levels, int(md["refine_kmer_center_idx"]), md["refine_do_rough_rescale"],
            int(md["refine_scale_iters"]), md["refine_algo"], int(md["refine_half_bandwidth"]),
            sd_arr)
    else:  # original models without a refiner (model_util.py:439-443)
        md["sig_map_refiner"] = SigMapRefiner()
        md["base_start_justify"] = False
        md["offset"] = 0
    for key in [k for k in md if k.startswith("refine_")]:
        del md[key]
    return md


def _raw_load_torchscript(model_filename):
    """TorchScript zip -> (state_dict on CPU, raw meta.txt dict) (model_util.py:468-481)."""
    extra = {"meta.txt": ""}
    module = torch.jit.load(model_filename, _extra_files=extra, map_location="cpu")
    return module.state_dict(), json.loads(extra["meta.txt"])


def _as_device(device):
    if device is None:
        if not torch.cuda.is_available():
            raise RemoraError("remora_b200 runs on CUDA (sm_100a) only and no GPU is visible; "
                              "there is no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())
    if isinstance(device, int):
        device = torch.device("cuda", device)
    device = torch.device(device)
    if device.type != "cuda":
        raise RemoraError(f"remora_b200 runs on CUDA devices only (got {device}); "
                          "there is no CPU fallback")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


class B200Model(nn.Module):
    """The object ``load_model`` returns in place of the reference's TorchScript module.

    Callers of the reference use: ``model(sigs, enc_kmers)``, ``next(model.parameters()).device``
    (data_chunks.py:528, inference.py:286), ``model.eval()``, ``param.requires_grad``
    (model_util.py:559-562).  All of those work here."""

    def __init__(self, state_dict, device=None):
        super().__init__()
        self._lib = _native.load_library()
        self._pinned_ok = set()   # (address, bytes) of host buffers already verified as pinned
        self._desc, self._blob, self.info = weights.pack_state_dict(state_dict)
        self._handle = ctypes.c_void_p()
        self._device = None
        # one real parameter so that next(model.parameters()).device reports the GPU
        self.anchor = nn.Parameter(torch.zeros(1), requires_grad=False)
        self._source_state = {k: v.detach().cpu() for k, v in state_dict.items()}
        self._create(_as_device(device))

    # -- lifetime ---------------------------------------------------------------------------
    def _create(self, device):
        self._release()
        handle = ctypes.c_void_p()
        rc = self._lib.rb200_create(ctypes.byref(self._desc), self._blob.ctypes.data,
                                    self._blob.size, device.index, ctypes.byref(handle))
        _native.check(rc, "rb200_create")
        self._handle = handle
        self._device = device
        self.anchor.data = self.anchor.data.to(device)

    def _release(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self._lib.rb200_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self._release()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    def to(self, *args, **kwargs):
        device = kwargs.get("device", args[0] if args else None)
        if device is not None and not isinstance(device, torch.dtype):
            device = _as_device(device)
            if device != self._device:
                self._create(device)
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", device if device is not None
                                    else torch.cuda.current_device()))

    def cpu(self):
        raise RemoraError("remora_b200 models run on CUDA only; there is no CPU fallback")

    def reference_state_dict(self):
        """The unfolded fp32 state_dict the model was built from (for export / inspection)."""
        return self._source_state

    @property
    def device(self):
        return self._device

    @property
    def num_out(self):
        return int(self._desc.num_out)

    @property
    def kmer_len(self):
        return int(self._desc.kmer_len)

    # -- controls used by tests / bench ----------------------------------------------------
    def set_impl(self, impl):
        code = {"auto": _native.IMPL_AUTO, "layers": _native.IMPL_LAYERS,
                "fused": _native.IMPL_FUSED, "fused_tc": _native.IMPL_FUSED_TC,
                "tiled": _native.IMPL_TILED, "fused_mega": _native.IMPL_FUSED_MEGA,
                "fused_bf16": _native.IMPL_FUSED_BF16}[impl]
        _native.check(self._lib.rb200_set_impl(self._handle, code), "rb200_set_impl")

    @property
    def last_impl(self):
        return {0: None, 1: "layers", 2: "fused", 3: "fused_tc", 4: "tiled", 5: "fused_mega",
                6: "fused_bf16"}[self._lib.rb200_last_impl(self._handle)]

    def get_flags(self, clear=False):
        """Sticky diagnostics of the single-kernel fp16-split path (rb200_get_flags): bit 0 = an
        intermediate activation left the fp16 range and was saturated.  Synchronises the device."""
        flags = ctypes.c_int32()
        _native.check(self._lib.rb200_get_flags(self._handle, ctypes.byref(flags), int(clear)),
                      "rb200_get_flags")
        return int(flags.value)

    @property
    def launch_count(self):
        return int(self._lib.rb200_launch_count(self._handle))

    def set_profile(self, on=True):
        _native.check(self._lib.rb200_set_profile(self._handle, int(on)), "rb200_set_profile")

    def get_profile(self):
        """(ms of K1, K2, K3 accumulated over the profiled forwards, number of forwards)."""
        ms = (ctypes.c_float * 3)()
        n = ctypes.c_int32()
        _native.check(self._lib.rb200_get_profile(self._handle, ctypes.byref(ms), ctypes.byref(n)),
                      "rb200_get_profile")
        return [float(v) for v in ms], int(n.value)

    def set_debug(self, keep=True):
        _native.check(self._lib.rb200_set_debug(self._handle, int(keep)), "rb200_set_debug")

    def debug_tensor(self, name):
        n, c, t = ctypes.c_int64(), ctypes.c_int32(), ctypes.c_int32()
        stream = ctypes.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)
        _native.check(self._lib.rb200_debug_tensor(self._handle, name.encode(), None, 0,
                                                   ctypes.byref(n), ctypes.byref(c),
                                                   ctypes.byref(t), stream), "rb200_debug_tensor")
        out = torch.empty(n.value, dtype=torch.float32, device=self._device)
        _native.check(self._lib.rb200_debug_tensor(self._handle, name.encode(), _ptr(out), n.value,
                                                   ctypes.byref(n), ctypes.byref(c),
                                                   ctypes.byref(t), stream), "rb200_debug_tensor")
        return out.view(-1, c.value, t.value)

    # -- the hot path ------------------------------------------------------------------------
    def _prep(self, t, dtype, name):
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(t)
        if t.dtype != dtype:
            raise RemoraError(f"{name} must be {dtype} (got {t.dtype})")
        if t.device != self._device:
            t = t.to(self._device, non_blocking=True)
        return t.contiguous()

    def forward(self, sigs, enc_kmers):
        """model(sigs float32 [B,1,T], enc_kmers float32 [B,4k,T]) -> float32 [B,num_out]
        (reference models/ConvLSTM_w_ref.py:39-58, models/Conv_w_ref.py:44-62)."""
        sigs = self._prep(sigs, torch.float32, "sigs")
        enc = self._prep(enc_kmers, torch.float32, "enc_kmers")
        if sigs.dim() != 3 or sigs.shape[1] != 1:
            raise RemoraError(f"sigs must be [B,1,T] (got {tuple(sigs.shape)})")
        B, _, T = sigs.shape
        if tuple(enc.shape) != (B, 4 * self.kmer_len, T):
            raise RemoraError(f"enc_kmers must be [{B},{4 * self.kmer_len},{T}] "
                              f"(got {tuple(enc.shape)})")
        out = torch.empty((B, self.num_out), dtype=torch.float32, device=self._device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)
        rc = self._lib.rb200_forward_dense(self._handle, _ptr(sigs), _ptr(enc), B, T, _ptr(out),
                                           stream)
        _native.check(rc, "rb200_forward_dense")
        return out

    def forward_compact(self, sigs, sequence, seq_to_sig_map, seq_lens, out=None):
        """Fused encode + forward on the reference's compact chunk arrays
        (CoreRemoraDataset._core_dtypes, data_chunks.py:942-948):
        sigs float32 [B,1,T] (or [B,T]), sequence int8 [B,Lmax+k-1], seq_to_sig_map int16 [B,Lmax+1],
        seq_lens int16 [B] -> float32 [B,num_out] (written into ``out`` when given: a contiguous float32
        device tensor of that shape, e.g. a slot of a gather buffer)."""
        sigs = self._prep(sigs, torch.float32, "sigs")
        seqs = self._prep(sequence, torch.int8, "sequence")
        maps = self._prep(seq_to_sig_map, torch.int16, "seq_to_sig_map")
        lens = self._prep(seq_lens, torch.int16, "seq_lens")
        if sigs.dim() == 3:
            if sigs.shape[1] != 1:
                raise RemoraError(f"sigs must be [B,1,T] (got {tuple(sigs.shape)})")
            B, _, T = sigs.shape
        else:
            B, T = sigs.shape
        if seqs.shape[0] != B or maps.shape[0] != B or lens.shape[0] != B:
            raise RemoraError("compact arrays disagree on the batch size")
        if out is None:
            out = torch.empty((B, self.num_out), dtype=torch.float32, device=self._device)
        elif (out.dtype != torch.float32 or not out.is_contiguous() or out.device != self._device
              or tuple(out.shape) != (B, self.num_out)):
            raise RemoraError("out must be a contiguous float32 [B, num_out] tensor on the model's device")
        stream = ctypes.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)
        rc = self._lib.rb200_forward_compact(self._handle, _ptr(sigs), _ptr(seqs), seqs.shape[1],
                                             _ptr(maps), maps.shape[1], _ptr(lens), B, T,
                                             _ptr(out), stream)
        _native.check(rc, "rb200_forward_compact")
        return out

    def forward_compact_gather(self, sigs, sequence, seq_to_sig_map, seq_lens, peer_bases_dev, n_peers,
                               dst_offset, multicast_ptr=0, flag_word=-1, out=None):
        """``forward_compact`` fused with the exchange step of the multi-GPU path
        (``rb200_forward_compact_gather``): the kernel's classifier epilogue stores the [B, num_out]
        logits at float offset ``dst_offset`` of every rank's buffer (``peer_bases_dev``: device pointer
        to the array of ``n_peers`` peer-mapped base pointers, e.g. symmetric memory ``buffer_ptrs_dev``)."""
        sigs = self._prep(sigs, torch.float32, "sigs")
        seqs = self._prep(sequence, torch.int8, "sequence")
        maps = self._prep(seq_to_sig_map, torch.int16, "seq_to_sig_map")
        lens = self._prep(seq_lens, torch.int16, "seq_lens")
        B, T = sigs.shape[0], sigs.shape[-1]
        stream = ctypes.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)
        rc = self._lib.rb200_forward_compact_gather(
            self._handle, _ptr(sigs), _ptr(seqs), seqs.shape[1], _ptr(maps), maps.shape[1], _ptr(lens), B, T,
            _ptr(out) if out is not None else None, ctypes.c_void_p(int(peer_bases_dev)), int(n_peers),
            int(dst_offset), ctypes.c_void_p(int(multicast_ptr)) if multicast_ptr else None, int(flag_word),
            stream)
        _native.check(rc, "rb200_forward_compact_gather")

    def forward_compact_ship(self, arrays, out, peer_bases_dev, n_peers, self_rank, ship_src=None,
                             ship_dst_offset=0, multicast_ptr=0, flag_word=-1, shape_hint=None):
        """``forward_compact`` into ``out`` (this rank's own block) while ONE extra thread block of the same
        launch ships an earlier call's finished block ``ship_src`` (a contiguous float32 device tensor) to
        float offset ``ship_dst_offset`` of every other rank's buffer (``rb200_forward_compact_ship``).
        ``arrays=None`` ships only (the flush after the last step); ``shape_hint`` = (chunk_len, seq_width,
        map_width) of the batches, needed then."""
        stream = ctypes.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)
        n_ship = int(ship_src.numel()) if ship_src is not None else 0
        if arrays is None:
            T, seq_w, map_w = shape_hint
            sigs = seqs = maps = lens = None
            B = 0
        else:
            sigs = self._prep(arrays[0], torch.float32, "sigs")
            seqs = self._prep(arrays[1], torch.int8, "sequence")
            maps = self._prep(arrays[2], torch.int16, "seq_to_sig_map")
            lens = self._prep(arrays[3], torch.int16, "seq_lens")
            B, T, seq_w, map_w = sigs.shape[0], sigs.shape[-1], seqs.shape[1], maps.shape[1]
            if (out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != B * self.num_out):
                raise RemoraError("out must be a contiguous float32 [B, num_out] tensor")
        rc = self._lib.rb200_forward_compact_ship(
            self._handle, _ptr(sigs) if B else None, _ptr(seqs) if B else None, seq_w,
            _ptr(maps) if B else None, map_w, _ptr(lens) if B else None, B, T, _ptr(out) if B else None,
            ctypes.c_void_p(int(peer_bases_dev)), int(n_peers), int(self_rank),
            _ptr(ship_src) if n_ship else None, int(ship_dst_offset), n_ship,
            ctypes.c_void_p(int(multicast_ptr)) if multicast_ptr else None, int(flag_word), stream)
        _native.check(rc, "rb200_forward_compact_ship")

    def softmax_ml(self, logits, want_probs=True):
        """Post-processing on the device (``rb200_softmax_ml``): row softmax of float32 logits
        [N, num_out], class 0 dropped -> (probs float32 [N, num_out-1] or None, ML bytes uint8
        [N, num_out-1] = min(floor(p*256), 255)), both device tensors
        (reference util.py:182-186, 532-535 as used at inference.py:429-459)."""
        logits = self._prep(logits, torch.float32, "logits")
        if logits.dim() != 2 or logits.shape[1] != self.num_out:
            raise RemoraError(f"logits must be [N,{self.num_out}] (got {tuple(logits.shape)})")
        n = logits.shape[0]
        probs = torch.empty((n, self.num_out - 1), dtype=torch.float32, device=self._device) if want_probs else None
        ml = torch.empty((n, self.num_out - 1), dtype=torch.uint8, device=self._device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)
        rc = self._lib.rb200_softmax_ml(_ptr(logits), n, self.num_out,
                                        _ptr(probs) if want_probs else None, _ptr(ml), stream)
        _native.check(rc, "rb200_softmax_ml")
        return probs, ml

    def infer_host(self, sigs, sequence, seq_to_sig_map, seq_lens):
        """numpy in / numpy out through rb200_infer_host (H2D, kernels, D2H, sync)."""
        sigs = np.ascontiguousarray(sigs, dtype=np.float32)
        seqs = np.ascontiguousarray(sequence, dtype=np.int8)
        maps = np.ascontiguousarray(seq_to_sig_map, dtype=np.int16)
        lens = np.ascontiguousarray(seq_lens, dtype=np.int16)
        B, T = sigs.shape[0], sigs.shape[-1]
        out = np.empty((B, self.num_out), dtype=np.float32)
        rc = self._lib.rb200_infer_host(self._handle, sigs.ctypes.data, seqs.ctypes.data,
                                        seqs.shape[1], maps.ctypes.data, maps.shape[1],
                                        lens.ctypes.data, B, T, out.ctypes.data)
        _native.check(rc, "rb200_infer_host")
        return out


    @staticmethod
    def pinned_batch(sigs, sequence, seq_to_sig_map, seq_lens):
        """Copies the four compact arrays (numpy or CPU tensors) into ONE pinned block, each array starting on
        the next 256-byte boundary - the layout ``rb200_infer_host_async`` recognises and moves with a single
        host-to-device copy.  Returns the four pinned views (pass them to :meth:`infer_host_async`)."""
        arrs = [torch.as_tensor(np.ascontiguousarray(a)) for a in (sigs, sequence, seq_to_sig_map, seq_lens)]
        dts = (torch.float32, torch.int8, torch.int16, torch.int16)
        arrs = [a.to(dt).contiguous() for a, dt in zip(arrs, dts)]
        sizes = [a.numel() * a.element_size() for a in arrs]
        offs, pos = [], 0
        for n in sizes:
            offs.append(pos)
            pos += (n + 255) // 256 * 256
        block = torch.empty(pos, dtype=torch.uint8).pin_memory()
        views = []
        for a, off, n in zip(arrs, offs, sizes):
            v = block[off:off + n].view(a.dtype).view(a.shape)
            v.copy_(a)
            views.append(v)
        return tuple(views)

    def infer_host_async(self, sigs, sequence, seq_to_sig_map, seq_lens, out, stream=None):
        """Pipelined host-buffer inference (rb200_infer_host_async): all arguments are PINNED CPU
        tensors (``torch.Tensor.pin_memory()``), ``out`` a pinned float32 [B, num_out] tensor.  Work is
        enqueued on ``stream`` (default: current stream) and the call returns immediately; synchronise
        the stream before reading ``out``."""
        # is_pinned() is a driver query (~1.5 us each): a buffer that passed once is remembered by its storage
        # address and size (pipelined callers reuse a handful of staging buffers)
        seen = self._pinned_ok
        for t, name in ((sigs, "sigs"), (sequence, "sequence"), (seq_to_sig_map, "seq_to_sig_map"),
                        (seq_lens, "seq_lens"), (out, "out")):
            if not (isinstance(t, torch.Tensor) and t.device.type == "cpu" and t.is_contiguous()):
                raise RemoraError(f"{name} must be a contiguous pinned CPU tensor")
            key = (t.data_ptr(), t.numel() * t.element_size())
            if key not in seen:
                if not t.is_pinned():
                    raise RemoraError(f"{name} must be a contiguous pinned CPU tensor")
                if len(seen) > 4096:
                    seen.clear()
                seen.add(key)
        B, T = sigs.shape[0], sigs.shape[-1]
        if stream is None:
            stream = torch.cuda.current_stream(self._device)
        rc = self._lib.rb200_infer_host_async(
            self._handle, _ptr(sigs), _ptr(sequence), sequence.shape[1], _ptr(seq_to_sig_map),
            seq_to_sig_map.shape[1], _ptr(seq_lens), B, T, _ptr(out), ctypes.c_void_p(stream.cuda_stream))
        _native.check(rc, "rb200_infer_host_async")
        return out


def load_torchscript_model(model_filename, device=None, quiet=False, eval_only=False):
    """Same contract as the reference function (model_util.py:532-563)."""
    state_dict, md = _raw_load_torchscript(model_filename)
    add_derived_metadata(md)
    model = B200Model(state_dict, device=device)
    md["sig_map_refiner"].device = next(model.parameters()).device  # the banded DP runs where the model does
    if eval_only:
        model.eval()
        for param in model.parameters():
            param.requires_grad = False
    return model, md


def load_model(model_filename=None, *, pore=None, basecall_model_type=None,
               basecall_model_version=None, modified_bases=None, remora_model_type=None,
               remora_model_version=None, device=None, quiet=True, eval_only=False):
    """Load a Remora model for B200 inference; keyword surface of the reference's ``load_model``
    (model_util.py:566-699).  Only the path form is implemented: the pore/basecaller registry
    lookup needs the reference's download machinery (download.py) and is out of scope here."""
    if model_filename is None:
        raise RemoraError("remora_b200.load_model needs a model path; the pretrained-model registry "
                          "lookup/download of the reference is not part of this package")
    if not isfile(model_filename):
        raise RemoraError(f"Remora model file ({model_filename}) not found.")
    try:
        return load_torchscript_model(model_filename, device, quiet=quiet, eval_only=eval_only)
    except (AttributeError, RuntimeError) as e:
        raise RemoraError(f"Failed loading torchscript model. ({e})")
