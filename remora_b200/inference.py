"""Single-read and batched inference entry points behind the reference's ``remora.inference``
surface: ``call_read_mods`` (src/remora/inference.py:661-712) and the device boundary of the
batched CLI pipeline, ``run_model_batched`` (inference.py:277-316)."""
import array

import numpy as np
import torch

from . import RemoraError, constants
from .util import Motif, format_mm_ml_tags, revcomp, softmax_axis1


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


def mods_tags_to_str(mm_tags, ml_arr):
    """SAM tag strings of the per-canonical-base MM strings and the joined ML array (reference
    util.py mods_tags_to_str as used at inference.py:447)."""
    return [f"MM:Z:{''.join(mm_tags)}", "ML:B:C," + ",".join(str(int(v)) for v in ml_arr)]


def _run_merged(model, batches, device, batch_size):
    """Concatenate the compact chunk arrays of several reads (padding the per-read sequence / mapping
    widths to the widest; the kernels never read past ``seq_len``) and run the model over them in
    pieces of ``batch_size`` chunks.  Returns float32 logits [total chunks, num_out] on the host."""
    import torch.nn.functional as F
    from .data_chunks import DeviceChunkBatch
    seq_w = max(b.sequence.shape[1] for b in batches)
    map_w = max(b.seq_to_sig_map.shape[1] for b in batches)
    sigs, seqs, maps, lens = [], [], [], []
    for b in batches:
        if isinstance(b, DeviceChunkBatch):
            sig, seq, mp, ln = b.signal, b.sequence, b.seq_to_sig_map, b.seq_lens
        else:
            sig, seq, mp, ln = (torch.from_numpy(a) for a in (b.signal, b.sequence, b.seq_to_sig_map,
                                                                b.seq_lens))
        sigs.append(sig)
        seqs.append(F.pad(seq, (0, seq_w - seq.shape[1]), value=-1))
        maps.append(F.pad(mp, (0, map_w - mp.shape[1]), value=0))
        lens.append(ln)
    sig, seq, mp, ln = (torch.cat(x).to(device, non_blocking=True) for x in (sigs, seqs, maps, lens))
    outs = []
    for st in range(0, sig.shape[0], batch_size):
        en = st + batch_size
        outs.append(model.forward_compact(sig[st:en], seq[st:en], mp[st:en], ln[st:en]))
    return torch.cat(outs).cpu().numpy()


def infer_from_pod5_and_bam(pod5_path, in_bam_path, models, out_path=None, num_reads=None,
                            batch_size=constants.DEFAULT_BATCH_SIZE, reads_per_batch=256, ref_anchored=False,
                            skip_non_primary=True, extract_on_device=False, return_probs=False,
                            decode_on_device=True, rank=0, world_size=1):
    """``remora infer from_pod5_and_bam`` as one function (reference inference.py:462-660 without its
    process/queue plumbing): POD5 signal + BAM basecalls/move tables -> modified-base calls per read.

    ``models``: ``(model, metadata)`` or ``{can_base: (model, metadata)}`` as ``load_model`` returns them.
    Reads are handled ``reads_per_batch`` at a time: POD5 signal decoded on the GPU
    (``rb200_svb16_decode``, ``decode_on_device``), joined with the BAM records on the host, their
    signal mappings refined in ONE banded-DP launch per model (``SigMapRefiner.refine_reads``), chunk
    arrays built per read (vectorised numpy by default: for one read at a time it beats the per-read
    launches and transfers of ``extract_on_device``, 1.4 k vs 0.8 k reads/s measured) and the network run
    over the chunks of the whole group in ``batch_size`` pieces.  Returns a list of dicts
    ``{read_id, mm, ml (array('B')), error}`` (plus ``calls``: ``{can_base: (positions, probs)}`` when
    ``return_probs``); with ``out_path`` the input records are also written with the MM/ML tags attached
    (previous MM/ML/mv tags dropped) - as BAM when the name ends in ``.bam``, else as SAM text -
    unmapped-style when reference anchored like the reference's output (inference.py:448-456).
    Multi-GPU: one process per GPU calls this with its ``rank`` / ``world_size`` (and its own
    ``out_path``); reads are split by sequence length (``parallel.shard_by_work``), no collective."""
    from . import io as rio
    from .refine_signal_map import SigMapRefiner
    if isinstance(models, tuple):
        models = {models[1]["can_base"]: models}
    rev_sigs = {bool(md["reverse_signal"]) for _, md in models.values()}
    pa_scalings = {None if md["pa_scaling"] is None else tuple(md["pa_scaling"]) for _, md in models.values()}
    if len(rev_sigs) != 1 or len(pa_scalings) != 1:
        raise RemoraError("models disagree on reverse_signal / pa_scaling")
    reverse_signal, pa_scaling = rev_sigs.pop(), pa_scalings.pop()
    bam_idx = rio.ReadIndexedBam(in_bam_path, skip_non_primary=skip_non_primary, req_tags={"mv"})
    if world_size > 1:
        from .parallel import shard_by_work
        ids = bam_idx.read_ids
        keep = shard_by_work([sum(len(r.query_sequence) for r in bam_idx[i]) for i in ids], world_size, rank)
        bam_idx._bam_idx = {ids[i]: bam_idx._bam_idx[ids[i]] for i in keep}
        bam_idx.num_reads = len(bam_idx._bam_idx)
    results = []
    out_fh = None
    out_records = None  # BAM output: records are collected and written at the end
    header_text = (bam_idx.header_text.rstrip("\n") + "\n" if bam_idx.header_text else "") + \
        "@PG\tID:remora_b200\tPN:remora_b200\n"
    if out_path is not None and str(out_path).endswith(".bam"):
        out_records = []
    elif out_path is not None:
        out_fh = open(out_path, "w")
        out_fh.write(header_text)

    def flush(group):
        # group: list of io.Read that converted cleanly; one refinement launch per model
        per_read = [dict(read_id=r.child_read_id, mm=[], ml=array.array("B"), error=None, calls={})
                    for r in group]
        for can_base, (model, md) in models.items():
            rreads = []
            for io_read, res in zip(group, per_read):
                try:
                    rreads.append(io_read.into_remora_read(ref_anchored))
                except RemoraError as e:
                    res["error"] = str(e)
                    rreads.append(None)
            live_idx = [i for i, r in enumerate(rreads) if r is not None]
            refiner = md["sig_map_refiner"]
            if refiner.is_loaded and live_idx:
                for i, msg in zip(live_idx, refiner.refine_reads([rreads[i] for i in live_idx])):
                    if msg is not None:  # this read alone fails, like a failed read of the reference's prep worker
                        per_read[i]["error"] = msg
            md_done = dict(md, sig_map_refiner=SigMapRefiner())  # refinement already applied above
            device = next(model.parameters()).device
            motifs = [Motif(*mot) for mot in md["motifs"]]
            # chunk arrays of every read of the group, then ONE stream of model calls over all of them
            # (the reference batches chunks across reads the same way, inference.py:185-274)
            owners, merged = [], []
            for i, rread in enumerate(rreads):
                if rread is None or per_read[i]["error"] is not None:
                    continue
                rread.set_motif_focus_bases(motifs)
                if extract_on_device:
                    rread.prepare_batches_gpu(md_done, batch_size=1 << 30, device=device)
                else:
                    rread.prepare_batches(md_done, batch_size=1 << 30)
                for b in rread.batches:
                    owners.append(i)
                    merged.append(b)
            if merged:
                nn_out = _run_merged(model, merged, device, batch_size)
                st = 0
                for i, b in zip(owners, merged):
                    out = nn_out[st:st + len(b)]
                    st += len(b)
                    probs = softmax_axis1(out)[:, 1:].astype(np.float64)
                    pos = b.read_focus_bases
                    io_read = group[i]
                    seq = io_read.ref_seq if ref_anchored else io_read.seq
                    mm, ml = format_mm_ml_tags(seq=seq, poss=pos, probs=probs, mod_bases=md["mod_bases"],
                                               can_base=can_base)
                    per_read[i]["mm"].append(mm)
                    per_read[i]["ml"].extend(ml)
                    if return_probs:
                        per_read[i]["calls"][can_base] = (pos, probs)
        for io_read, res in zip(group, per_read):
            res["mm"] = "".join(res["mm"])
            if not return_probs:
                del res["calls"]
            results.append(res)
            rec = io_read.alignment_record
            if (out_fh is not None or out_records is not None) and res["error"] is None and rec is not None:
                extra = mods_tags_to_str([res["mm"]], res["ml"])
                if ref_anchored:
                    import dataclasses
                    seq = io_read.ref_seq if io_read.ref_reg.strand == "+" else revcomp(io_read.ref_seq)
                    rec = dataclasses.replace(rec, cigartuples=[(0, len(io_read.ref_seq))], query_sequence=seq,
                                              query_qualities=np.zeros(0, dtype=np.uint8))
                drop = ("MM", "ML", "Mm", "Ml", "mv")
                if out_records is not None:
                    out_records.append(rec.to_record(drop_tags=drop, extra_tags=[
                        ("MM", "Z", res["mm"]), ("ML", "BC", np.frombuffer(res["ml"], dtype=np.uint8))]))
                else:
                    out_fh.write(rec.to_sam(drop_tags=drop, extra_tags=extra) + "\n")

    group = []
    try:
        decode_dev = next(next(iter(models.values()))[0].parameters()).device if decode_on_device else None
        for io_read, err in rio.iter_io_reads(pod5_path, bam_idx, num_reads=num_reads,
                                              reverse_signal=reverse_signal, pa_scaling=pa_scaling,
                                              device=decode_dev):
            if err is not None:
                results.append(dict(read_id=io_read.read_id, mm="", ml=array.array("B"), error=err))
                continue
            group.append(io_read)
            if len(group) >= reads_per_batch:
                flush(group)
                group = []
        if group:
            flush(group)
    finally:
        if out_fh is not None:
            out_fh.close()
    if out_records is not None:
        rio.write_bam(out_path, header_text, rio.references_from_header(header_text), out_records)
    return results
