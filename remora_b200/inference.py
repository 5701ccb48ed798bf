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
    if not return_mod_probs and not return_mm_ml_tags:
        return read.run_model(model)
    if hasattr(model, "softmax_ml"):
        # post-processing on the device: softmax, class 0 dropped, ML bytes; only
        # the float32 probabilities / uint8 ML bytes come back, the logits never leave the GPU
        nn_dev, labels, pos = read.run_model(model, keep_on_device=True)
        probs_dev, ml_dev = model.softmax_ml(nn_dev, want_probs=not return_mm_ml_tags)
        if return_mm_ml_tags:
            return format_mm_ml_tags(seq=read.str_seq, poss=pos, probs=None, ml_bytes=ml_dev.cpu().numpy(),
                                     mod_bases=model_metadata["mod_bases"],
                                     can_base=model_metadata["can_base"])
        return probs_dev.cpu().numpy().astype(np.float64), labels, pos
    nn_out, labels, pos = read.run_model(model)
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
    pieces of ``batch_size`` chunks.  Returns float32 logits [total chunks, num_out] on the device."""
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
    return torch.cat(outs)


def infer_from_pod5_and_bam(pod5_path, in_bam_path, models, out_path=None, num_reads=None,
                            batch_size=constants.DEFAULT_BATCH_SIZE, reads_per_batch=256, ref_anchored=False,
                            skip_non_primary=True, extract_on_device=False, return_probs=False,
                            decode_on_device=True, rank=0, world_size=1, drop_move_tag=False, out_format=None,
                            bam_in_memory=False):
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
    ``return_probs``); with ``out_path`` EVERY input record of the processed reads is written, like the
    reference's main loop does (inference.py:609-625): called reads with the MM/ML tags attached
    (previous MM/ML dropped), reads that failed (join, refinement, chunking) unchanged without new tags.
    The move table is kept unless ``drop_move_tag`` (the reference keeps it, ``prune(drop_move_tag=False)``).
    Format: ``out_format`` "bam" / "sam", default by the name (``*.bam`` -> BAM, else SAM text);
    reference-anchored calls are written unmapped-style like the reference's (inference.py:448-456).
    The BAM is indexed by file offsets (``bam_in_memory=False``), as the reference does (io.py:255-307).
    Multi-GPU: one process per GPU calls this with its ``rank`` / ``world_size`` (and its own
    ``out_path``); ``num_reads`` selects the reads of the run FIRST, then they are split by sequence
    length (``parallel.shard_by_work``, lengths recorded while indexing), no collective."""
    from . import io as rio
    from .refine_signal_map import SigMapRefiner
    if isinstance(models, tuple):
        models = {models[1]["can_base"]: models}
    rev_sigs = {bool(md["reverse_signal"]) for _, md in models.values()}
    pa_scalings = {None if md["pa_scaling"] is None else tuple(md["pa_scaling"]) for _, md in models.values()}
    if len(rev_sigs) != 1 or len(pa_scalings) != 1:
        raise RemoraError("models disagree on reverse_signal / pa_scaling")
    reverse_signal, pa_scaling = rev_sigs.pop(), pa_scalings.pop()
    bam_idx = rio.ReadIndexedBam(in_bam_path, skip_non_primary=skip_non_primary, req_tags={"mv"},
                                 in_memory=bam_in_memory)
    shard_ids = None
    if world_size > 1:
        from .parallel import shard_by_work
        with rio.Pod5Reader(pod5_path) as reader:  # the run = POD5 order, truncated by num_reads, THEN sharded
            ids = [rid for rid in reader.read_ids if rid in bam_idx]
        if num_reads is not None:
            ids = ids[:num_reads]
        shard_ids = [ids[i] for i in shard_by_work([bam_idx.seq_lens[i] for i in ids], world_size, rank)]
    results = []
    out_fh = None
    out_records = None  # BAM output: records are collected and written at the end
    header_text = (bam_idx.header_text.rstrip("\n") + "\n" if bam_idx.header_text else "") + \
        "@PG\tID:remora_b200\tPN:remora_b200\n"
    if out_format is None:
        out_format = "bam" if str(out_path).endswith(".bam") else "sam"
    if out_format not in ("bam", "sam"):
        raise RemoraError(f"unknown output format {out_format!r} (bam or sam)")
    if out_path is not None and out_format == "bam":
        out_records = []
    elif out_path is not None:
        out_fh = open(out_path, "w")
        out_fh.write(header_text)
    drop = ("MM", "ML", "Mm", "Ml") + (("mv",) if drop_move_tag else ())

    def emit(rec, res, io_read=None):
        """One output record: the input record, plus MM/ML when the read was called."""
        if rec is None or (out_fh is None and out_records is None):
            return
        called = res["error"] is None
        if called and ref_anchored and io_read is not None:
            import dataclasses
            seq = io_read.ref_seq if io_read.ref_reg.strand == "+" else revcomp(io_read.ref_seq)
            rec = dataclasses.replace(rec, cigartuples=[(0, len(io_read.ref_seq))], query_sequence=seq,
                                      query_qualities=np.zeros(0, dtype=np.uint8))
        if out_records is not None:
            extra = [("MM", "Z", res["mm"]), ("ML", "BC", np.frombuffer(res["ml"], dtype=np.uint8))] if called else []
            out_records.append(rec.to_record(drop_tags=drop if called else (), extra_tags=extra))
        else:
            extra = mods_tags_to_str([res["mm"]], res["ml"]) if called else []
            out_fh.write(rec.to_sam(drop_tags=drop if called else (), extra_tags=extra) + "\n")

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
                # softmax + ML-byte quantisation on the device for the chunks of the whole group
                # (rb200_softmax_ml; reference inference.py:429-459 does this per read in numpy)
                probs_dev, ml_dev = model.softmax_ml(_run_merged(model, merged, device, batch_size),
                                                     want_probs=return_probs)
                ml_all = ml_dev.cpu().numpy()
                probs_all = probs_dev.cpu().numpy().astype(np.float64) if return_probs else None
                st = 0
                for i, b in zip(owners, merged):
                    ml_b = ml_all[st:st + len(b)]
                    probs = probs_all[st:st + len(b)] if return_probs else None
                    st += len(b)
                    pos = b.read_focus_bases
                    io_read = group[i]
                    seq = io_read.ref_seq if ref_anchored else io_read.seq
                    mm, ml = format_mm_ml_tags(seq=seq, poss=pos, probs=None, ml_bytes=ml_b,
                                               mod_bases=md["mod_bases"], can_base=can_base)
                    per_read[i]["mm"].append(mm)
                    per_read[i]["ml"].extend(ml)
                    if return_probs:
                        per_read[i]["calls"][can_base] = (pos, probs)
        for io_read, res in zip(group, per_read):
            res["mm"] = "".join(res["mm"])
            if not return_probs:
                del res["calls"]
            results.append(res)
            emit(io_read.alignment_record, res, io_read)

    group = []
    try:
        decode_dev = next(next(iter(models.values()))[0].parameters()).device if decode_on_device else None
        for io_read, err in rio.iter_io_reads(pod5_path, bam_idx, num_reads=num_reads,
                                              reverse_signal=reverse_signal, pa_scaling=pa_scaling,
                                              device=decode_dev, read_ids=shard_ids):
            if err is not None:
                if group:  # keep the output in input order
                    flush(group)
                    group = []
                res = dict(read_id=io_read.child_read_id, mm="", ml=array.array("B"), error=err)
                results.append(res)
                emit(io_read.alignment_record, res)
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
        bam_idx.close()
    if out_records is not None:
        # the binary reference dictionary of the input (present even when the text header has no @SQ lines)
        refs = list(zip(bam_idx.references, bam_idx.ref_lengths)) or rio.references_from_header(header_text)
        rio.write_bam(out_path, header_text, refs, out_records)
    return results
