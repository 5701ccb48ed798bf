"""Single-read and batched inference entry points behind the reference's ``remora.inference``
surface: ``call_read_mods`` (src/remora/inference.py:661-712) and the device boundary of the
batched CLI pipeline, ``run_model_batched`` (inference.py:277-316)."""
import numpy as np
import torch

from . import constants
from .util import Motif, format_mm_ml_tags, softmax_axis1


def call_read_mods(read, model, model_metadata, batch_size=constants.DEFAULT_BATCH_SIZE,
                   focus_offset=None, return_mm_ml_tags=False, return_mod_probs=False,
                   extract_on_device=False):
    """Call modified bases on a read; same arguments and return forms as the reference
    (inference.py:661-712):
      default               -> (nn_out float32 [N,num_out], labels, positions)
      return_mod_probs      -> (probs float64 [N,num_out-1], labels, positions)
      return_mm_ml_tags     -> (MM string, ML array('B'))
      read without calls    -> three empty arrays
    Positions come back sorted (the reference returns set-iteration order, util.py:419-426).
    ``extract_on_device`` (extension): build the chunk arrays with the GPU extraction kernels
    (``RemoraRead.prepare_batches_gpu``) instead of on the host; results are identical."""
    if focus_offset is None:
        read.set_motif_focus_bases([Motif(*mot) for mot in model_metadata["motifs"]])
    else:
        read.focus_bases = np.array([focus_offset])
    if extract_on_device:
        read.prepare_batches_gpu(model_metadata, batch_size, device=next(model.parameters()).device)
    else:
        read.prepare_batches(model_metadata, batch_size)
    if len(read.batches) == 0:
        return np.array([]), np.array([]), np.array([])
    nn_out, labels, pos = read.run_model(model)
    if not return_mod_probs and not return_mm_ml_tags:
        return nn_out, labels, pos
    probs = softmax_axis1(nn_out)[:, 1:].astype(np.float64)
    if return_mm_ml_tags:
        return format_mm_ml_tags(seq=read.str_seq, poss=pos, probs=probs,
                                 mod_bases=model_metadata["mod_bases"],
                                 can_base=model_metadata["can_base"])
    return probs, labels, pos


def run_model_batched(batches, models, models_metadata, batch_size):
    """Device boundary of ``remora infer`` (reference inference.py:277-316), as a generator
    instead of a queue-to-queue thread body: consumes items
    ``(can_base, b_sigs f32[B,1,T], b_enc_kmers f32[B,4k,T], b_read_pos, b_reads)`` and yields
    ``(can_base, nn_out (device tensor), b_read_pos, b_reads)`` in submission order.  Keeps the
    reference's persistent pinned staging buffers for full batches and the direct path for the
    ragged last batch (inference.py:305-310)."""
    md = {m["can_base"]: m for m in models_metadata}
    devices, sig_bufs, enc_bufs = {}, {}, {}
    for cb, meta in md.items():
        devices[cb] = next(models[cb].parameters()).device
        pin = devices[cb].type == "cuda"
        sig_bufs[cb] = torch.empty((batch_size, 1, meta["chunk_len"]), dtype=torch.float32,
                                   pin_memory=pin)
        enc_bufs[cb] = torch.empty((batch_size, meta["kmer_len"] * 4, meta["chunk_len"]),
                                   dtype=torch.float32, pin_memory=pin)
    for can_base, b_sigs, b_enc_kmers, b_read_pos, b_reads in batches:
        if b_read_pos.size == batch_size:
            sig_bufs[can_base][:] = torch.from_numpy(b_sigs)
            enc_bufs[can_base][:] = torch.from_numpy(b_enc_kmers)
            sigs, enc = sig_bufs[can_base], enc_bufs[can_base]
        else:
            sigs, enc = torch.from_numpy(b_sigs), torch.from_numpy(b_enc_kmers)
        dev = devices[can_base]
        nn_out = models[can_base](sigs.to(dev, non_blocking=True), enc.to(dev, non_blocking=True))
        yield can_base, nn_out, b_read_pos, b_reads
