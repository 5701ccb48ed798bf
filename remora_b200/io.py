"""Minimal POD5 and BAM readers and the read object that joins them ("next" row 3),
behind the parts of the reference's ``remora.io`` surface the inference path touches
(src/remora/io.py:147-180 helpers, :184-360 ``ReadIndexedBam``, :394-411 ``parse_move_tag``,
:415-474 POD5 iteration, :1747-2177 ``Read``).

The reference reads both formats through third-party C libraries (``pysam`` / htslib, ``pod5``); neither
is needed here:
  * BAM  = BGZF (concatenated raw-deflate blocks, each with its size in a gzip extra field) around
           length-prefixed binary records; decoded with ``zlib`` + ``struct`` (SAM/BAM specification
           sections 4.1-4.2).  Records expose the ``pysam.AlignedSegment`` attributes the reference uses.
  * POD5 = three Arrow IPC files (signal, run-info and reads tables) embedded in one file and located
           by a FlatBuffers footer; the tables are read with ``pyarrow``, the signal rows
           ("minknow.vbz": zstd around StreamVByte-16 of zig-zag deltas) are decoded with numpy.
This is host plumbing that feeds the GPU path; it has no device code.
"""
import array
import dataclasses
import mmap
import re
import struct
import uuid
import zlib
from collections import defaultdict

import numpy as np

from . import RemoraError, util
from . import data_chunks as DC

PA_TO_NORM_SCALING_FACTOR = 1.4826  # reference constants.py:243

# ------------------------------------------------------------------------------------------------
# BAM
# ------------------------------------------------------------------------------------------------
_SEQ_CODES = "=ACMGRSVTWYHKDBN"
_SEQ_LUT = np.array([ord(a) for a in _SEQ_CODES for _ in _SEQ_CODES], dtype=np.uint8), \
    np.array([ord(b) for _ in _SEQ_CODES for b in _SEQ_CODES], dtype=np.uint8)
_TAG_SCALARS = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I", "f": "<f"}
_TAG_ARRAYS = {"c": np.int8, "C": np.uint8, "s": "<i2", "S": "<u2", "i": "<i4", "I": "<u4", "f": "<f4"}
_TAG_ARRAY_CODES = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}
FLAG_UNMAPPED, FLAG_REVERSE, FLAG_SECONDARY, FLAG_SUPPLEMENTARY = 0x4, 0x10, 0x100, 0x800
# cigar operations consuming query / reference (reference data_chunks.py:29-35)
MATCH_OPS = np.array([True, False, False, False, False, False, False, True, True])
QUERY_OPS = np.array([True, True, False, False, True, False, False, True, True])
REF_OPS = np.array([True, False, True, True, False, False, False, True, True])


def iter_bgzf_blocks(fh):
    """Yields (file offset, decompressed bytes) of every BGZF block (SAM spec 4.1)."""
    while True:
        start = fh.tell()
        head = fh.read(12)
        if len(head) == 0:
            return
        if len(head) < 12 or head[:4] != b"\x1f\x8b\x08\x04":
            raise RemoraError("not a BGZF block (is this a BAM file?)")
        xlen = struct.unpack("<H", head[10:12])[0]
        extra = fh.read(xlen)
        bsize = None
        pos = 0
        while pos + 4 <= len(extra):
            si1, si2, slen = extra[pos], extra[pos + 1], struct.unpack("<H", extra[pos + 2:pos + 4])[0]
            if si1 == 66 and si2 == 67 and slen == 2:
                bsize = struct.unpack("<H", extra[pos + 4:pos + 6])[0]
            pos += 4 + slen
        if bsize is None:
            raise RemoraError("BGZF block without BC subfield")
        cdata = fh.read(bsize - xlen - 19)
        crc, isize = struct.unpack("<II", fh.read(8))
        data = zlib.decompress(cdata, wbits=-15) if isize else b""
        if len(data) != isize or (zlib.crc32(data) & 0xFFFFFFFF) != crc:
            raise RemoraError("corrupt BGZF block")
        yield start, data


class _BgzfStream:
    """Sequential reader over the decompressed stream that remembers BAM virtual offsets
    (block file offset << 16 | offset within the block)."""

    def __init__(self, fh):
        self._fh = fh
        self._blocks = iter_bgzf_blocks(fh)
        self._buf = b""
        self._pos = 0
        self._block_start = 0

    def seek(self, voffset):
        """Jump to a BAM virtual offset (as returned by :meth:`tell`).  A jump inside the block that is
        already decompressed costs nothing (records fetched through the pointer index mostly sit in the same
        64 KB block as their predecessor: re-inflating it per record was 12 % of the file pipeline)."""
        block = voffset >> 16
        if self._buf and block == self._block_start and (voffset & 0xFFFF) <= len(self._buf):
            self._pos = voffset & 0xFFFF
            return
        self._fh.seek(block)
        self._blocks = iter_bgzf_blocks(self._fh)
        self._buf, self._pos = b"", 0
        if not self._fill():
            raise RemoraError("BAM virtual offset beyond the end of the file")
        self._pos = voffset & 0xFFFF

    def _fill(self):
        for start, data in self._blocks:
            if data:
                self._block_start, self._buf, self._pos = start, data, 0
                return True
        return False

    def tell(self):
        if self._pos >= len(self._buf) and not self._fill():
            return None
        return (self._block_start << 16) | self._pos

    def read(self, n):
        out = []
        while n > 0:
            if self._pos >= len(self._buf) and not self._fill():
                break
            chunk = self._buf[self._pos:self._pos + n]
            self._pos += len(chunk)
            n -= len(chunk)
            out.append(chunk)
        return b"".join(out)


@dataclasses.dataclass
class AlignedSegment:
    """One BAM record with the ``pysam.AlignedSegment`` attributes the reference's io layer reads."""

    query_name: str
    flag: int
    reference_id: int
    reference_name: str
    reference_start: int
    mapping_quality: int
    cigartuples: list
    query_sequence: str
    query_qualities: np.ndarray
    tags: list  # [(tag, value)]
    tag_types: dict = None  # tag -> BAM type character ("B" tags: "B" + element type)
    next_reference_id: int = -1
    next_reference_start: int = -1
    template_length: int = 0
    next_reference_name: str = None

    @property
    def is_unmapped(self):
        return bool(self.flag & FLAG_UNMAPPED)

    @property
    def is_reverse(self):
        return bool(self.flag & FLAG_REVERSE)

    @property
    def is_forward(self):
        return not self.is_reverse

    @property
    def is_secondary(self):
        return bool(self.flag & FLAG_SECONDARY)

    @property
    def is_supplementary(self):
        return bool(self.flag & FLAG_SUPPLEMENTARY)

    def has_tag(self, tag):
        return any(t == tag for t, _ in self.tags)

    def get_tag(self, tag):
        for t, v in self.tags:
            if t == tag:
                return v
        raise KeyError(f"tag '{tag}' not present")

    def get_reference_sequence(self):
        """Reference bases under the alignment, rebuilt from the query, CIGAR and MD tag (what
        ``pysam.AlignedSegment.get_reference_sequence`` does); mismatches come back lower case."""
        try:
            md = self.get_tag("MD")
        except KeyError:
            raise ValueError("MD tag not present")
        aligned = []
        qpos = 0
        for op, ln in self.cigartuples:
            if MATCH_OPS[op]:
                aligned.append(self.query_sequence[qpos:qpos + ln])
            elif op == 3:
                raise ValueError("reference skips (N) are not supported")
            if QUERY_OPS[op]:
                qpos += ln
        aligned = "".join(aligned)
        out = []
        apos = 0
        for num, dele, mis in re.findall(r"(\d+)|\^([A-Za-z]+)|([A-Za-z])", md):
            if num:
                out.append(aligned[apos:apos + int(num)])
                apos += int(num)
            elif dele:
                out.append(dele.upper())
            else:
                out.append(mis.lower())
                apos += 1
        if apos != len(aligned):
            raise ValueError("MD tag discordant with CIGAR")
        return "".join(out)

    def to_sam(self, drop_tags=(), extra_tags=()):
        """SAM text line; ``extra_tags`` are ready-made ``TAG:TYPE:VALUE`` strings."""
        if self.next_reference_name is None:
            rnext = "*"
        else:
            rnext = "=" if self.next_reference_name == self.reference_name else self.next_reference_name
        qual = ("*" if self.query_qualities.size == 0 or self.query_qualities[0] == 0xFF
                else (self.query_qualities + 33).astype(np.uint8).tobytes().decode("ascii"))
        fields = [self.query_name, str(self.flag), self.reference_name or "*", str(self.reference_start + 1),
                  str(self.mapping_quality),
                  "".join(f"{ln}{DC_CIGAR_CODES[op]}" for op, ln in self.cigartuples) or "*", rnext,
                  str(self.next_reference_start + 1), str(self.template_length), self.query_sequence or "*", qual]
        for tag, val in self.tags:
            if tag in drop_tags:
                continue
            typ = (self.tag_types or {}).get(tag, "Z")
            if typ[0] == "B":
                fields.append(f"{tag}:B:{typ[1]}," + ",".join(
                    (repr(float(v)) if typ[1] == "f" else str(int(v))) for v in val))
            elif typ in "cCsSiI":
                fields.append(f"{tag}:i:{int(val)}")
            elif typ == "f":
                fields.append(f"{tag}:f:{float(val):g}")
            else:
                fields.append(f"{tag}:{typ}:{val}")
        fields.extend(extra_tags)
        return "\t".join(fields)

    def to_record(self, drop_tags=(), extra_tags=()):
        """dict form accepted by :func:`write_bam`; ``extra_tags``: [(tag, BAM type, value)]."""
        types = self.tag_types or {}
        tags = [(t, types.get(t, "Z"), v) for t, v in self.tags if t not in drop_tags]
        return dict(query_name=self.query_name, flag=self.flag, reference_id=self.reference_id,
                    reference_start=self.reference_start, mapping_quality=self.mapping_quality,
                    cigartuples=self.cigartuples, query_sequence=self.query_sequence,
                    query_qualities=self.query_qualities, next_reference_id=self.next_reference_id,
                    next_reference_start=self.next_reference_start, template_length=self.template_length,
                    tags=tags + list(extra_tags))

    def to_dict(self):
        return {"name": self.query_name, "flag": str(self.flag), "ref_name": self.reference_name or "*",
                "ref_pos": str(self.reference_start + 1), "map_quality": str(self.mapping_quality),
                "cigar": "".join(f"{ln}{DC_CIGAR_CODES[op]}" for op, ln in self.cigartuples) or "*",
                "seq": self.query_sequence, "tags": [t for t, _ in self.tags]}


DC_CIGAR_CODES = "MIDNSHP=X"


def _parse_tags(buf, pos, end):
    tags, types = [], {}
    while pos < end:
        tag = buf[pos:pos + 2].decode("ascii")
        typ = chr(buf[pos + 2])
        types[tag] = typ if typ != "B" else "B" + chr(buf[pos + 3])
        pos += 3
        if typ in _TAG_SCALARS:
            fmt = _TAG_SCALARS[typ]
            size = struct.calcsize(fmt)
            val = struct.unpack_from(fmt, buf, pos)[0]
            pos += size
        elif typ == "A":
            val = chr(buf[pos])
            pos += 1
        elif typ in "ZH":
            z = buf.index(b"\x00", pos)
            val = buf[pos:z].decode("ascii")
            pos = z + 1
        elif typ == "B":
            sub = chr(buf[pos])
            count = struct.unpack_from("<I", buf, pos + 1)[0]
            dt = np.dtype(_TAG_ARRAYS[sub])
            pos += 5
            # array.array like pysam returns (python-int items, so `mv[0]` is an int, not an int8)
            val = array.array(_TAG_ARRAY_CODES[sub], bytes(buf[pos:pos + count * dt.itemsize]))
            pos += count * dt.itemsize
        else:
            raise RemoraError(f"unknown BAM tag type '{typ}'")
        tags.append((tag, val))
    return tags, types


def _parse_record(buf, ref_names):
    (ref_id, pos, l_name, mapq, _bin, n_cigar, flag, l_seq, next_ref, next_pos,
     tlen) = struct.unpack_from("<iiBBHHHIiii", buf, 0)
    off = 32
    name = buf[off:off + l_name - 1].decode("ascii")
    off += l_name
    cig = np.frombuffer(buf, dtype="<u4", count=n_cigar, offset=off)
    cigartuples = [(int(c & 0xF), int(c >> 4)) for c in cig]
    off += 4 * n_cigar
    packed = np.frombuffer(buf, dtype=np.uint8, count=(l_seq + 1) // 2, offset=off)
    seq = np.empty(packed.size * 2, dtype=np.uint8)
    seq[0::2] = _SEQ_LUT[0][packed]
    seq[1::2] = _SEQ_LUT[1][packed]
    off += (l_seq + 1) // 2
    qual = np.frombuffer(buf, dtype=np.uint8, count=l_seq, offset=off)
    off += l_seq
    tags, types = _parse_tags(buf, off, len(buf))
    return AlignedSegment(
        query_name=name, flag=flag, reference_id=ref_id,
        reference_name=ref_names[ref_id] if 0 <= ref_id < len(ref_names) else None,
        reference_start=pos, mapping_quality=mapq, cigartuples=cigartuples,
        query_sequence=seq[:l_seq].tobytes().decode("ascii"), query_qualities=qual,
        tags=tags, tag_types=types, next_reference_id=next_ref, next_reference_start=next_pos,
        template_length=tlen,
        next_reference_name=ref_names[next_ref] if 0 <= next_ref < len(ref_names) else None)


class BamReader:
    """Sequential BAM reader: ``header_text``, ``references`` / ``lengths`` and iteration over
    :class:`AlignedSegment` records (with their virtual file offsets via :meth:`iter_with_offsets`)."""

    def __init__(self, path):
        self.path = str(path)
        self._fh = open(self.path, "rb")
        self._stream = _BgzfStream(self._fh)
        if self._stream.read(4) != b"BAM\x01":
            raise RemoraError(f"{self.path} is not a BAM file")
        l_text = struct.unpack("<i", self._stream.read(4))[0]
        self.header_text = self._stream.read(l_text).rstrip(b"\x00").decode("utf-8", errors="replace")
        n_ref = struct.unpack("<i", self._stream.read(4))[0]
        self.references, self.lengths = [], []
        for _ in range(n_ref):
            l_name = struct.unpack("<i", self._stream.read(4))[0]
            self.references.append(self._stream.read(l_name)[:-1].decode("ascii"))
            self.lengths.append(struct.unpack("<i", self._stream.read(4))[0])

    def read_at(self, voffset):
        """The record that starts at a virtual offset (random access without an index file)."""
        self._stream.seek(voffset)
        head = self._stream.read(4)
        if len(head) < 4:
            raise RemoraError("truncated BAM record")
        size = struct.unpack("<i", head)[0]
        buf = self._stream.read(size)
        if len(buf) < size:
            raise RemoraError("truncated BAM record")
        return _parse_record(buf, self.references)

    def iter_with_offsets(self):
        while True:
            voff = self._stream.tell()
            if voff is None:
                return
            head = self._stream.read(4)
            if len(head) < 4:
                return
            size = struct.unpack("<i", head)[0]
            buf = self._stream.read(size)
            if len(buf) < size:
                raise RemoraError("truncated BAM record")
            yield voff, _parse_record(buf, self.references)

    def __iter__(self):
        for _, rec in self.iter_with_offsets():
            yield rec

    def close(self):
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def read_is_primary(read):
    """io.py:147-154"""
    return not (read.is_supplementary or read.is_secondary)


def get_parent_id(bam_read):
    """Split reads carry their parent's id in the ``pi`` tag (io.py:174-180)."""
    try:
        return bam_read.get_tag("pi")
    except KeyError:
        return bam_read.query_name


@dataclasses.dataclass
class ReadIndexedBam:
    """BAM records indexed by (parent) read id, same constructor arguments and query methods as the
    reference class (io.py:183-360).  ``in_memory=True`` (default) keeps the parsed records; with
    ``in_memory=False`` only BAM virtual offsets are kept, like the reference's file pointers
    (io.py:255-307), and records are re-read on demand - the form for BAM files that do not fit in RAM."""

    bam_path: str
    skip_non_primary: bool = True
    req_tags: set = None
    read_id_converter: object = None
    parent_read_id_subset: set = None
    child_read_id_subset: set = None
    in_memory: bool = True

    def __post_init__(self):
        self.num_reads = None
        self.num_records = 0
        self.skip_reasons = defaultdict(int)
        self._bam_idx = None
        self.compute_read_index()

    @property
    def filename(self):
        return self.bam_path

    reference_filename = filename

    def compute_read_index(self):
        idx = defaultdict(list)
        self._reader = None
        self.seq_lens = {}  # (extension) total basecall length per indexed read id: sharding work estimate
        self.ref_lengths = []
        with BamReader(self.bam_path) as bam:
            self.header_text = bam.header_text
            self.references = bam.references
            self.ref_lengths = list(bam.lengths)  # binary reference dictionary (may exist without @SQ lines)
            for voff, read in bam.iter_with_offsets():
                index_read_id, reason = self._index_key(read)
                if reason is not None:  # same reasons, same order of the tests as the reference (io.py:255-307)
                    self.skip_reasons[reason] += 1
                    continue
                self.num_records += 1
                idx[index_read_id].append(read if self.in_memory else voff)
                self.seq_lens[index_read_id] = self.seq_lens.get(index_read_id, 0) + len(read.query_sequence)
        self._bam_idx = dict(idx)
        self.num_reads = len(self._bam_idx)

    def _index_key(self, read):
        """(id the record is filed under, None) or (None, why it is skipped)."""
        keep_child, keep_parent = self.child_read_id_subset, self.parent_read_id_subset
        if keep_child is not None and read.query_name not in keep_child:
            return None, "Child read ID filtered"
        key = get_parent_id(read)
        if keep_parent is not None and key not in keep_parent:
            return None, "Parent read ID filtered"
        if self.read_id_converter is not None:
            key = self.read_id_converter(key)
        if self.req_tags is not None and not self.req_tags <= {t for t, _ in read.tags}:
            return None, "Missing BAM tags"
        if self.skip_non_primary and (read.is_supplementary or read.is_secondary):
            return None, "Non-primary alignment"
        return key, None

    def _materialise(self, entry):
        if self.in_memory:
            return entry
        if self._reader is None:
            self._reader = BamReader(self.bam_path)
        return self._reader.read_at(entry)

    def get_alignments(self, read_id):
        try:
            entries = self._bam_idx[read_id]
        except KeyError:
            raise RemoraError(f"Could not find {read_id} in {self.bam_path}")
        for entry in entries:
            yield self._materialise(entry)

    def get_first_alignment(self, read_id):
        return next(self.get_alignments(read_id))

    def __contains__(self, read_id):
        return read_id in self._bam_idx

    def __getitem__(self, read_id):
        return [self._materialise(e) for e in self._bam_idx[read_id]]

    @property
    def read_ids(self):
        return list(self._bam_idx.keys())

    def __iter__(self):
        for entries in self._bam_idx.values():
            for entry in entries:
                yield self._materialise(entry)

    def close(self):
        if self._reader is not None:
            self._reader.close()
            self._reader = None


def parse_move_tag(mv_tag, sig_len, seq_len=None, check=True, reverse_signal=False):
    """Move table (``mv`` tag: stride, then one flag per stride-sized signal block, 1 = a base starts here)
    -> signal index at which every base starts, plus the end of the signal (io.py:394-411)."""
    stride, moves = int(mv_tag[0]), np.asarray(mv_tag[1:])
    starts = np.flatnonzero(moves) * stride
    bounds = np.append(starts, sig_len)
    if reverse_signal:  # the signal was flipped: measure from its other end
        bounds = (sig_len - bounds)[::-1]
    if check:
        if seq_len is not None and bounds.size != seq_len + 1:
            raise RemoraError("Move table discordant with basecalls")
        if moves.size != sig_len // stride:
            raise RemoraError("Move table discordant with signal")
    return bounds, moves, stride


def make_sequence_coordinate_mapping(cigar):
    """Reference position -> (fractional) query position from CIGAR knots (data_chunks.py:77-115)."""
    cigar = list(cigar)
    while len(cigar) > 0 and not MATCH_OPS[cigar[-1][0]]:
        cigar = cigar[:-1]
    if len(cigar) == 0:
        raise RemoraError("No match operations found in alignment cigar")
    ops, lens = map(np.array, zip(*cigar))
    if ops.min() < 0 or ops.max() > 8:
        raise RemoraError("Invalid cigar op(s)")
    if lens.min() < 0:
        raise RemoraError("Cigar lengths may not be negative")
    is_match = MATCH_OPS[ops]
    match_counts = lens[is_match]
    offsets = np.array([match_counts, np.ones_like(match_counts)])
    ref_knots = np.cumsum(np.where(REF_OPS[ops], lens, 0))
    ref_knots = np.concatenate([[0], (ref_knots[is_match] - offsets).T.flatten(), [ref_knots[-1]]])
    query_knots = np.cumsum(np.where(QUERY_OPS[ops], lens, 0))
    query_knots = np.concatenate([[0], (query_knots[is_match] - offsets).T.flatten(), [query_knots[-1]]])
    return np.interp(np.arange(ref_knots[-1] + 1), ref_knots, query_knots)


def compute_ref_to_signal(query_to_signal, cigar):
    """data_chunks.py:60-74, 118-122"""
    knots = make_sequence_coordinate_mapping(cigar)
    return np.floor(np.interp(knots, np.arange(query_to_signal.size), query_to_signal)).astype(int)


# ------------------------------------------------------------------------------------------------
# POD5
# ------------------------------------------------------------------------------------------------
POD5_SIGNATURE = b"\x8bPOD\r\n\x1a\n"


def _fb_table_fields(buf, pos):
    """FlatBuffers table at ``pos`` -> list of absolute field positions (None = absent)."""
    vt = pos - struct.unpack_from("<i", buf, pos)[0]
    vt_size = struct.unpack_from("<H", buf, vt)[0]
    out = []
    for o in range(4, vt_size, 2):
        rel = struct.unpack_from("<H", buf, vt + o)[0]
        out.append(pos + rel if rel else None)
    return out


def parse_pod5_footer(buf):
    """Embedded-file list of a POD5 footer (FlatBuffers ``Footer{file_identifier, software,
    pod5_version, contents:[EmbeddedFile{offset:int64, length:int64, format, content_type}]}``)."""
    root = struct.unpack_from("<I", buf, 0)[0]
    fields = _fb_table_fields(buf, root)
    if len(fields) < 4 or fields[3] is None:
        raise RemoraError("POD5 footer has no embedded-file list")
    vec = fields[3] + struct.unpack_from("<I", buf, fields[3])[0]
    n = struct.unpack_from("<I", buf, vec)[0]
    files = []
    for i in range(n):
        elem = vec + 4 + 4 * i
        tbl = elem + struct.unpack_from("<I", buf, elem)[0]
        f = _fb_table_fields(buf, tbl)
        off = struct.unpack_from("<q", buf, f[0])[0] if len(f) > 0 and f[0] is not None else 0
        length = struct.unpack_from("<q", buf, f[1])[0] if len(f) > 1 and f[1] is not None else 0
        ctype = struct.unpack_from("<h", buf, f[3])[0] if len(f) > 3 and f[3] is not None else 0
        files.append((off, length, ctype))
    return files


def _zstd_frame_content(blob):
    """zstd-decompress one frame whose header carries the content size (POD5 writers always set it)."""
    import pyarrow as pa
    blob = bytes(blob)
    if blob[:4] != b"\x28\xb5\x2f\xfd":
        raise RemoraError("signal chunk is not a zstd frame")
    fhd = blob[4]
    fcs_flag, single, did = fhd >> 6, (fhd >> 5) & 1, fhd & 3
    pos = 5 + (0 if single else 1) + (0, 1, 2, 4)[did]
    fcs_size = (1 if single else 0, 2, 4, 8)[fcs_flag]
    if fcs_size == 0:
        raise RemoraError("zstd frame without content size")
    size = int.from_bytes(blob[pos:pos + fcs_size], "little") + (256 if fcs_size == 2 else 0)
    return pa.Codec("zstd").decompress(blob, decompressed_size=size)


def decode_vbz_rows_gpu(blobs, n_samples, device, read_of_row=None):
    """Decode many "minknow.vbz" rows on the GPU in one launch (``rb200_svb16_decode``): zstd on the host,
    StreamVByte-16 + zig-zag + running sum on the device.  ``read_of_row`` groups consecutive rows into
    reads; returns (int16 device tensor, list of (start, length) per read) - every read starts at a
    multiple of 8 samples so the kernel can use 16-byte stores."""
    import ctypes
    import torch
    from . import _native
    lib = _native.load_library()
    device = torch.device(device)
    if device.type != "cuda":
        raise RemoraError("GPU signal decode needs a CUDA device")
    n_rows = len(blobs)
    if read_of_row is None:
        read_of_row = list(range(n_rows))
    raws = [_zstd_frame_content(b) for b in blobs]
    row_off = np.zeros(n_rows + 1, dtype=np.int64)
    row_off[1:] = np.cumsum([r.size for r in raws])
    packed = np.zeros(int(row_off[-1]) + 16, dtype=np.uint8)
    for r, raw in enumerate(raws):
        packed[row_off[r]:row_off[r + 1]] = np.frombuffer(raw, dtype=np.uint8)
    out_off = np.zeros(n_rows, dtype=np.int64)
    spans, pos, prev = [], 0, None
    for r in range(n_rows):
        if read_of_row[r] != prev:
            pos = (pos + 7) & ~7
            spans.append([pos, 0])
            prev = read_of_row[r]
        out_off[r] = pos
        pos += int(n_samples[r])
        spans[-1][1] += int(n_samples[r])
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    with torch.cuda.device(device):
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        d_packed = torch.from_numpy(packed).to(device, non_blocking=True)
        d_row_off = torch.from_numpy(row_off).to(device, non_blocking=True)
        d_n = torch.from_numpy(np.asarray(n_samples, dtype=np.int32)).to(device, non_blocking=True)
        d_out_off = torch.from_numpy(out_off).to(device, non_blocking=True)
        d_out = torch.zeros(max(pos, 1), dtype=torch.int16, device=device)
        d_status = torch.zeros(max(n_rows, 1), dtype=torch.int32, device=device)
        max_n = int(max(n_samples)) if n_rows else 0
        need = ctypes.c_int64()
        _native.check(lib.rb200_svb16_scratch_bytes(n_rows, max_n, ctypes.byref(need)), "rb200_svb16_scratch_bytes")
        d_scratch = torch.empty(max(int(need.value), 4), dtype=torch.uint8, device=device)
        _native.check(lib.rb200_svb16_decode(ptr(d_packed), ptr(d_row_off), ptr(d_n), ptr(d_out_off), n_rows,
                                             max_n, ptr(d_out), ptr(d_status), ptr(d_scratch), stream),
                      "rb200_svb16_decode")
        if n_rows and int(d_status.max()) != 0:
            raise RemoraError("corrupt svb16 signal chunk")
    return d_out, [tuple(s) for s in spans]


def decode_vbz(blob, n_samples):
    """"minknow.vbz" signal chunk -> int16 samples: zstd frame -> StreamVByte-16 (one key BIT per value,
    0 = one byte, 1 = two bytes; keys first, then the data bytes) -> zig-zag -> running sum."""
    if n_samples == 0:
        return np.zeros(0, dtype=np.int16)
    raw = np.frombuffer(_zstd_frame_content(blob), dtype=np.uint8)
    n_key = (n_samples + 7) // 8
    keys = np.unpackbits(raw[:n_key], bitorder="little")[:n_samples].astype(np.int64)
    data = raw[n_key:]
    if data.size != n_samples + int(keys.sum()):
        raise RemoraError("corrupt svb16 signal chunk")
    off = np.cumsum(1 + keys) - (1 + keys)
    lo = data[off].astype(np.uint16)
    hi = np.where(keys == 1, data[np.minimum(off + 1, data.size - 1)], 0).astype(np.uint16)
    u = lo | (hi << 8)
    delta = (u >> 1).astype(np.int16) ^ -(u & 1).astype(np.int16)
    return np.cumsum(delta, dtype=np.int64).astype(np.int16)


@dataclasses.dataclass
class Calibration:
    offset: float
    scale: float


@dataclasses.dataclass
class Pod5Read:
    """The attributes of ``pod5.ReadRecord`` the reference reads (io.py:455-462, 2103-2110)."""

    read_id: uuid.UUID
    signal: np.ndarray
    calibration: Calibration
    num_samples: int
    read_number: int = 0
    channel: int = 0


class Pod5Reader:
    """Reads one ``.pod5`` file: ``read_ids`` and ``reads(selection=None)`` like ``pod5.Reader``."""

    def __init__(self, path):
        import pyarrow as pa
        import pyarrow.ipc as ipc
        self.path = str(path)
        self._fh = open(self.path, "rb")
        self._mm = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        mm = self._mm
        if mm[:8] != POD5_SIGNATURE or mm[-8:] != POD5_SIGNATURE:
            raise RemoraError(f"{self.path} is not a POD5 file")
        flen = struct.unpack("<q", mm[-32:-24])[0]
        fstart = len(mm) - 32 - flen
        if flen <= 0 or fstart < 8:
            raise RemoraError("corrupt POD5 footer")
        self._tables = {}
        for off, length, _ctype in parse_pod5_footer(mm[fstart:fstart + flen]):
            reader = ipc.open_file(pa.py_buffer(memoryview(mm)[off:off + length]))
            names = set(reader.schema.names)
            if "samples" in names and "signal" in names:
                self._tables["signal"] = reader
            elif "calibration_offset" in names:
                self._tables["reads"] = reader
            elif "acquisition_id" in names:
                self._tables["run_info"] = reader
        if "signal" not in self._tables or "reads" not in self._tables:
            raise RemoraError("POD5 file lacks a signal or reads table")
        self._reads = self._tables["reads"].read_all()
        sig = self._tables["signal"]
        self._sig_batch_rows = np.cumsum([0] + [sig.get_batch(i).num_rows
                                                for i in range(sig.num_record_batches)])
        self._sig_vbz = str(sig.schema.field("signal").type) == "large_binary" or \
            "vbz" in str(sig.schema.field("signal").type)
        ids = self._reads.column("read_id").to_pylist()
        self._ids = [uuid.UUID(bytes=bytes(b)) for b in ids]
        self._row_of = {str(u): i for i, u in enumerate(self._ids)}
        # the per-read fields the join needs, pulled out of Arrow once
        names = set(self._reads.schema.names)
        col = lambda n, default: (self._reads.column(n).to_pylist() if n in names  # noqa: E731
                                  else [default] * len(ids))
        self._rec = {n: col(n, d) for n, d in (("signal", []), ("calibration_offset", 0.0),
                                               ("calibration_scale", 1.0), ("num_samples", 0),
                                               ("read_number", 0), ("channel", 0))}

    def _record(self, row):
        return {n: v[row] for n, v in self._rec.items()}

    @property
    def read_ids(self):
        return [str(u) for u in self._ids]

    @property
    def num_reads(self):
        return len(self._ids)

    def _signal_row(self, row):
        sig = self._tables["signal"]
        b = int(np.searchsorted(self._sig_batch_rows, row, side="right") - 1)
        batch = sig.get_batch(b)
        i = row - int(self._sig_batch_rows[b])
        n = batch.column(batch.schema.get_field_index("samples"))[i].as_py()
        cell = batch.column(batch.schema.get_field_index("signal"))[i]
        if self._sig_vbz:
            return decode_vbz(cell.as_py(), n)
        return np.asarray(cell.as_py(), dtype=np.int16)

    def get_read(self, read_id):
        row = self._row_of.get(str(read_id))
        if row is None:
            raise RemoraError(f"read {read_id} not in {self.path}")
        rec = self._record(row)
        parts = [self._signal_row(int(r)) for r in rec["signal"]]
        signal = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int16)
        return Pod5Read(read_id=self._ids[row], signal=signal,
                        calibration=Calibration(rec["calibration_offset"], rec["calibration_scale"]),
                        num_samples=int(rec["num_samples"]), read_number=int(rec["read_number"]),
                        channel=int(rec["channel"]))

    def get_reads(self, read_ids, device=None):
        """Several reads at once.  With a CUDA ``device`` all their signal rows are decoded by ONE kernel
        launch (``decode_vbz_rows_gpu``) and copied back in one transfer; otherwise row by row in numpy."""
        if device is None or not self._sig_vbz:
            return [self.get_read(rid) for rid in read_ids]
        sig = self._tables["signal"]
        recs, blobs, counts, owner = [], [], [], []
        for k, rid in enumerate(read_ids):
            row = self._row_of.get(str(rid))
            if row is None:
                raise RemoraError(f"read {rid} not in {self.path}")
            rec = self._record(row)
            recs.append((row, rec))
            for r in rec["signal"]:
                b = int(np.searchsorted(self._sig_batch_rows, int(r), side="right") - 1)
                batch = sig.get_batch(b)
                i = int(r) - int(self._sig_batch_rows[b])
                counts.append(batch.column(batch.schema.get_field_index("samples"))[i].as_py())
                blobs.append(batch.column(batch.schema.get_field_index("signal"))[i].as_py())
                owner.append(k)
        d_out, spans = decode_vbz_rows_gpu(blobs, counts, device, owner)
        host = d_out.cpu().numpy()
        span_of = dict(zip(sorted(set(owner)), spans))
        out = []
        for k, (row, rec) in enumerate(recs):
            st, ln = span_of.get(k, (0, 0))
            out.append(Pod5Read(read_id=self._ids[row], signal=host[st:st + ln],
                                calibration=Calibration(rec["calibration_offset"], rec["calibration_scale"]),
                                num_samples=int(rec["num_samples"]), read_number=int(rec["read_number"]),
                                channel=int(rec["channel"])))
        return out

    def reads(self, selection=None):
        ids = self.read_ids if selection is None else [str(s) for s in selection]
        for rid in ids:
            if rid in self._row_of:
                yield self.get_read(rid)

    def close(self):
        self._reads = None
        self._tables = {}
        try:
            self._mm.close()
        except BufferError:  # arrow buffers still alive; the map is released with them
            pass
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def iter_pod5_reads(pod5_path, num_reads=None, read_ids=None):
    """io.py:415-438"""
    with Pod5Reader(pod5_path) as reader:
        for read_num, read in enumerate(reader.reads(selection=read_ids)):
            if num_reads is not None and read_num >= num_reads:
                return
            yield read


# ------------------------------------------------------------------------------------------------
# the joined read
# ------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class RefRegion:
    """io.py:45-56"""

    ctg: str
    strand: str
    start: int
    end: int = None

    @property
    def len(self):
        return self.end - self.start


@dataclasses.dataclass
class Read:
    """Signal + basecalls + move table + scaling + alignment of one read: the fields and the methods
    of the reference's ``io.Read`` (io.py:1747-2177) that lead to a ``RemoraRead``."""

    read_id: str
    dacs: np.ndarray = None
    seq: str = None
    stride: int = None
    mv_table: np.ndarray = None
    query_to_signal: np.ndarray = None
    shift_dacs_to_pa: float = None
    scale_dacs_to_pa: float = None
    shift_pa_to_norm: float = None
    scale_pa_to_norm: float = None
    shift_dacs_to_norm: float = None
    scale_dacs_to_norm: float = None
    shift_pa_to_zc_pa: float = None
    scale_pa_to_zc_pa: float = None
    ref_seq: str = None
    ref_reg: RefRegion = None
    cigar: list = None
    ref_to_signal: np.ndarray = None
    full_align: dict = None
    _child_read_id: str = None
    _sig_len: int = None
    alignment_record: AlignedSegment = None  # (extension) the BAM record this read was joined with

    @property
    def pa_signal(self):
        return self._rescaled(self.shift_dacs_to_pa, self.scale_dacs_to_pa, "pA scaling factors not set")

    @property
    def norm_signal(self):
        return self._rescaled(self.shift_dacs_to_norm, self.scale_dacs_to_norm, "Norm scaling factors not set")

    def _rescaled(self, shift, scale, missing):
        """(dacs - shift) / scale; ``missing`` is the reference's error text when a factor is unset."""
        if shift is None or scale is None:
            raise RemoraError(missing)
        return (self.dacs - shift) / scale

    def compute_pa_to_norm_scaling(self, factor=PA_TO_NORM_SCALING_FACTOR):
        """median / MAD normalisation when the BAM carries no sm/sd tags (io.py:1851-1856)"""
        pa = self.pa_signal
        centre = np.median(pa)
        self.shift_pa_to_norm = centre
        self.scale_pa_to_norm = max(1.0, factor * np.median(np.abs(pa - centre)))

    @property
    def sig_len(self):
        if self._sig_len is None and self.dacs is not None:
            self._sig_len = self.dacs.size
        return self._sig_len

    @property
    def seq_len(self):
        if self.query_to_signal is not None:
            return self.query_to_signal.size - 1
        return len(self.seq) if self.seq is not None else None

    @property
    def child_read_id(self):
        return self.read_id if self._child_read_id is None else self._child_read_id

    # dacs -> zero-centred pA is the composition of dacs -> pA and pA -> zero-centred pA:
    # ((d - s1) / c1 - s2) / c2 = (d - (s1 + c1 s2)) / (c1 c2)
    @property
    def shift_dacs_to_zc_pa(self):
        parts = (self.shift_dacs_to_pa, self.scale_dacs_to_pa, self.shift_pa_to_zc_pa)
        if any(v is None for v in parts):
            raise RemoraError("Zero-centered pA scaling factors not set")
        return parts[0] + parts[1] * parts[2]

    @property
    def scale_dacs_to_zc_pa(self):
        parts = (self.scale_dacs_to_pa, self.scale_pa_to_zc_pa)
        if any(v is None for v in parts):
            raise RemoraError("Zero-centered pA scaling factors not set")
        return parts[0] * parts[1]

    def copy(self):
        return dataclasses.replace(self)

    # -- joining a BAM record onto the signal ---------------------------------------------------------
    # What the basecaller's tags mean (SAM tag conventions of dorado / guppy, as consumed by the reference
    # at io.py:1972-2084):  sp = samples a parent read was split at, ts = samples trimmed from the start
    # before basecalling, ns = number of samples the basecaller saw counted from the split point (so the
    # kept window is [sp + ts, sp + ns)), pi = parent read id of a split read, mv = stride + move table
    # over that window, sm / sd = shift / scale from pA to the normalised signal the basecaller used.
    def _window_signal(self, tags, reverse_signal):
        """Cut the raw samples down to the window the move table describes.  For reversed (RNA) signal
        the tags refer to the acquisition order, so the cut happens in that order."""
        sig = self.dacs[::-1] if reverse_signal else self.dacs
        sig = sig[tags.get("sp", 0):]
        sig = sig[tags.get("ts", 0):tags.get("ns", sig.size)]
        self.dacs = sig[::-1] if reverse_signal else sig
        self._sig_len = None

    def _claim_record(self, record, tags):
        """The record must belong to this read: by name, or - for a split read - through its parent id."""
        parent = tags.get("pi")
        if parent is None and record.query_name != self.read_id:
            raise RemoraError("Read IDs mismatch")
        if parent is not None:
            if parent != self.read_id:
                raise RemoraError("Split read IDs mismatch")
            self._child_read_id = record.query_name

    def _take_basecalls(self, record, tags, reverse_signal):
        """Basecalls in sequencing direction (BAM stores reverse-strand alignments reverse-complemented)
        and the first sample of every base from the move table, when there is one."""
        self.seq = util.revcomp(record.query_sequence) if record.is_reverse else record.query_sequence
        if "mv" in tags:
            self.query_to_signal, self.mv_table, self.stride = parse_move_tag(
                tags["mv"], sig_len=self.sig_len, seq_len=len(self.seq), reverse_signal=reverse_signal)
        else:
            self.query_to_signal = self.mv_table = self.stride = None

    def _take_scaling(self, tags):
        """pA -> normalised: the basecaller's sm / sd when present, else median / MAD of this read;
        composed with the POD5 calibration into the DAC -> normalised pair RemoraRead uses."""
        if "sm" in tags and "sd" in tags:
            self.shift_pa_to_norm, self.scale_pa_to_norm = tags["sm"], tags["sd"]
        else:
            self.compute_pa_to_norm_scaling()
        self.shift_dacs_to_norm = self.shift_dacs_to_pa + (self.scale_dacs_to_pa * self.shift_pa_to_norm)
        self.scale_dacs_to_norm = self.scale_dacs_to_pa * self.scale_pa_to_norm

    def _anchor_to_reference(self, record):
        """Reference coordinates of a mapped record, everything turned into sequencing direction:
        region, reference sequence rebuilt from MD + CIGAR, reversed CIGAR on the minus strand, and the
        first sample of every REFERENCE base by pushing reference positions through the CIGAR onto
        the move table."""
        minus = record.is_reverse
        self.ref_reg = RefRegion(ctg=record.reference_name, strand="-" if minus else "+",
                                 start=record.reference_start)
        try:
            fwd = record.get_reference_sequence().upper()
            self.ref_seq = util.revcomp(fwd) if minus else fwd
        except ValueError:  # no (usable) MD tag
            self.ref_seq = None
        self.cigar = record.cigartuples[::-1] if minus else record.cigartuples
        if self.ref_reg.ctg is None or self.ref_seq is None or self.query_to_signal is None:
            return
        self.ref_to_signal = compute_ref_to_signal(self.query_to_signal, self.cigar)
        if self.ref_to_signal.size != len(self.ref_seq) + 1:
            raise RemoraError("Discordant ref seq lengths")
        self.ref_reg.end = self.ref_reg.start + self.ref_to_signal.size - 1

    def add_alignment(self, alignment_record, parse_ref_align=True, reverse_signal=False, pa_scaling=None):
        """Join a BAM record onto a read that already holds its POD5 samples and calibration
        (same signature and resulting fields as the reference method, io.py:1972-2084)."""
        if self.dacs is None:
            raise RemoraError("Must add signal to io.Read before alignment.")
        if alignment_record.reference_name is None and alignment_record.is_reverse:
            raise RemoraError("Unmapped reads cannot map to reverse strand.")
        if pa_scaling is not None:
            self.shift_pa_to_zc_pa, self.scale_pa_to_zc_pa = pa_scaling
        tags = dict(alignment_record.tags)
        self.full_align = alignment_record.to_dict()
        if isinstance(alignment_record, AlignedSegment):
            self.alignment_record = alignment_record
        self._window_signal(tags, reverse_signal)
        self._claim_record(alignment_record, tags)
        self._take_basecalls(alignment_record, tags, reverse_signal)
        self._take_scaling(tags)
        if parse_ref_align and not alignment_record.is_unmapped:
            self._anchor_to_reference(alignment_record)

    @classmethod
    def from_pod5_and_alignment(cls, pod5_read_record, alignment_record, reverse_signal=False, pa_scaling=None):
        """io.py:2086-2121; POD5 calibration is pA = (dac + offset) * scale, i.e. shift = -offset,
        scale = 1 / scale in this class's (x - shift) / scale convention."""
        dacs = pod5_read_record.signal
        if reverse_signal:
            dacs = dacs[::-1]
        read = cls(read_id=str(pod5_read_record.read_id), dacs=dacs,
                   shift_dacs_to_pa=-pod5_read_record.calibration.offset,
                   scale_dacs_to_pa=1 / pod5_read_record.calibration.scale)
        read.add_alignment(alignment_record, reverse_signal=reverse_signal, pa_scaling=pa_scaling)
        return read

    def into_remora_read(self, use_reference_anchor):
        """The read as the chunking / inference layer wants it (io.py:2123-2177): the samples the sequence is
        mapped to, a mapping that starts at 0, the sequence (basecalls, or the reference it aligns to), and
        the scaling to the model's signal space (zero-centred pA when the model asks for it, else the
        basecaller's normalisation)."""
        if use_reference_anchor:
            if self.ref_to_signal is None:
                if self.cigar is None or self.ref_seq is None:
                    raise RemoraError("Missing reference alignment")
                self.ref_to_signal = compute_ref_to_signal(self.query_to_signal, self.cigar)
                if self.ref_to_signal.size != len(self.ref_seq) + 1:
                    raise RemoraError("Discordant ref seq lengths")
            bounds, seq = self.ref_to_signal, self.ref_seq
        else:
            if self.query_to_signal is None:
                raise RemoraError("Missing query_to_signal (move table)")
            bounds, seq = self.query_to_signal, self.seq
        zero_centred = self.shift_pa_to_zc_pa is not None and self.scale_pa_to_zc_pa is not None
        shift = self.shift_dacs_to_zc_pa if zero_centred else self.shift_dacs_to_norm
        scale = self.scale_dacs_to_zc_pa if zero_centred else self.scale_dacs_to_norm
        out = DC.RemoraRead(dacs=self.dacs[bounds[0]:bounds[-1]], shift=shift, scale=scale,
                            seq_to_sig_map=bounds - bounds[0], str_seq=seq, read_id=self.read_id)
        out.check()
        return out


def iter_io_reads(pod5_path, bam_idx, num_reads=None, reverse_signal=False, pa_scaling=None, device=None,
                  reads_per_decode=64, read_ids=None):
    """(io.Read, error text or None) for every alignment of every POD5 read present in the BAM index:
    the sequential equivalent of the reference's iter_signal -> extract_alignments workers
    (io.py:441-511).  With a CUDA ``device`` the signal of ``reads_per_decode`` reads at a time is decoded
    on the GPU (``Pod5Reader.get_reads``)."""
    with Pod5Reader(pod5_path) as reader:
        wanted = [rid for rid in reader.read_ids if rid in bam_idx]
        if num_reads is not None:
            wanted = wanted[:num_reads]
        if read_ids is not None:  # a rank's shard of the (already truncated) run
            keep = set(read_ids)
            wanted = [rid for rid in wanted if rid in keep]
        for st in range(0, len(wanted), reads_per_decode):
            ids = wanted[st:st + reads_per_decode]
            for rid, pod5_read in zip(ids, reader.get_reads(ids, device=device)):
                for bam_read in bam_idx.get_alignments(rid):
                    try:
                        yield Read.from_pod5_and_alignment(pod5_read, bam_read, reverse_signal=reverse_signal,
                                                           pa_scaling=pa_scaling), None
                    except RemoraError as e:  # the record still reaches the output, without new tags
                        yield Read(read_id=rid, _child_read_id=bam_read.query_name, alignment_record=bam_read), str(e)


# ------------------------------------------------------------------------------------------------
# writers (fixtures, round-trip tests, examples): the same two formats, produced without pysam / pod5
# ------------------------------------------------------------------------------------------------
def encode_vbz(signal):
    """int16 samples -> "minknow.vbz" chunk (inverse of :func:`decode_vbz`)."""
    import pyarrow as pa
    signal = np.ascontiguousarray(signal, dtype=np.int16)
    n = signal.size
    delta = np.diff(signal.astype(np.int32), prepend=np.int32(0)).astype(np.int16)
    u = ((delta.astype(np.int32) << 1) ^ (delta.astype(np.int32) >> 15)).astype(np.uint16)
    keys = (u > 0xFF).astype(np.uint8)
    key_bytes = np.packbits(keys, bitorder="little")
    lens = 1 + keys.astype(np.int64)
    off = np.cumsum(lens) - lens
    data = np.zeros(int(lens.sum()), dtype=np.uint8)
    data[off] = (u & 0xFF).astype(np.uint8)
    two = keys == 1
    data[off[two] + 1] = (u[two] >> 8).astype(np.uint8)
    raw = np.concatenate([key_bytes, data]).tobytes()
    return pa.Codec("zstd").compress(raw, asbytes=True)


def _fb_footer(files):
    """FlatBuffers ``Footer`` with an embedded-file list [(offset, length, content_type)]."""
    def align(buf, a):
        buf.extend(b"\x00" * (-len(buf) % a))

    buf = bytearray(4)  # root uoffset, patched below
    # vtables first (tables refer to them with a positive signed offset)
    vt_root = len(buf)
    buf += struct.pack("<HHHHHH", 12, 20, 4, 8, 12, 16)
    vt_file = len(buf)
    buf += struct.pack("<HHHHHH", 12, 28, 8, 16, 24, 26)
    align(buf, 8)
    root = len(buf)
    struct.pack_into("<I", buf, 0, root)
    buf += struct.pack("<i", root - vt_root) + b"\x00" * 16  # 4 uoffsets patched below
    strings = [b"remora_b200", b"remora_b200", b"0.3.2"]
    for i, text in enumerate(strings):
        align(buf, 4)
        pos = len(buf)
        struct.pack_into("<I", buf, root + 4 + 4 * i, pos - (root + 4 + 4 * i))
        buf += struct.pack("<I", len(text)) + text + b"\x00"
    align(buf, 4)
    vec = len(buf)
    struct.pack_into("<I", buf, root + 16, vec - (root + 16))
    buf += struct.pack("<I", len(files)) + b"\x00" * (4 * len(files))
    for i, (off, length, ctype) in enumerate(files):
        align(buf, 8)
        tbl = len(buf)
        elem = vec + 4 + 4 * i
        struct.pack_into("<I", buf, elem, tbl - elem)
        buf += struct.pack("<i4xqqhh", tbl - vt_file, off, length, 0, ctype)
    align(buf, 8)
    return bytes(buf)


def write_pod5(path, reads, chunk=102400, rows_per_batch=None):
    """Writes a POD5 file holding ``reads``: iterable of (read_id str, int16 signal, calibration offset,
    calibration scale).  Signal rows are VBZ chunks of at most ``chunk`` samples like MinKNOW's;
    ``rows_per_batch`` splits the Arrow tables into several record batches as real files are."""
    import pyarrow as pa
    import pyarrow.ipc as ipc
    sig_ids, sig_blobs, sig_n = [], [], []
    rows = []
    for read_id, signal, cal_off, cal_scale in reads:
        uid = uuid.UUID(str(read_id)).bytes
        signal = np.ascontiguousarray(signal, dtype=np.int16)
        idx = []
        for st in range(0, max(signal.size, 1), chunk):
            part = signal[st:st + chunk]
            idx.append(len(sig_blobs))
            sig_ids.append(uid)
            sig_blobs.append(encode_vbz(part))
            sig_n.append(part.size)
        rows.append((uid, idx, signal.size, float(cal_off), float(cal_scale)))
    signal_tbl = pa.table({
        "read_id": pa.array(sig_ids, type=pa.binary(16)),
        "signal": pa.array(sig_blobs, type=pa.large_binary()),
        "samples": pa.array(sig_n, type=pa.uint32()),
    })
    run_tbl = pa.table({"acquisition_id": pa.array(["remora_b200"]), "sample_rate": pa.array([5000], pa.uint16())})
    reads_tbl = pa.table({
        "read_id": pa.array([r[0] for r in rows], type=pa.binary(16)),
        "signal": pa.array([r[1] for r in rows], type=pa.list_(pa.uint64())),
        "read_number": pa.array(range(len(rows)), type=pa.uint32()),
        "num_samples": pa.array([r[2] for r in rows], type=pa.uint64()),
        "channel": pa.array([1] * len(rows), type=pa.uint16()),
        "calibration_offset": pa.array([r[3] for r in rows], type=pa.float32()),
        "calibration_scale": pa.array([r[4] for r in rows], type=pa.float32()),
    })
    marker = uuid.uuid4().bytes
    out = bytearray(POD5_SIGNATURE + marker)
    files = []
    for tbl, ctype in ((signal_tbl, 1), (run_tbl, 4), (reads_tbl, 0)):
        sink = pa.BufferOutputStream()
        with ipc.new_file(sink, tbl.schema) as writer:
            writer.write_table(tbl, max_chunksize=rows_per_batch)
        blob = sink.getvalue().to_pybytes()
        files.append((len(out), len(blob), ctype))
        out += blob
        out += b"\x00" * (-len(out) % 8)
        out += marker
    footer = _fb_footer(files)
    out += b"FOOTER\x00\x00" + footer + struct.pack("<q", len(footer)) + marker + POD5_SIGNATURE
    with open(path, "wb") as fh:
        fh.write(out)


_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _bgzf_block(data, level=1):
    # deflate level 1: the compression level is not part of the format; at level 6 zlib was 40 % of the host
    # time of the file pipeline's output stage (the move tables and ML arrays compress little anyway)
    comp = zlib.compressobj(level, zlib.DEFLATED, -15)
    cdata = comp.compress(data) + comp.flush()
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00"
            + struct.pack("<H", len(cdata) + 25) + cdata
            + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def _encode_tag(tag, typ, val):
    head = tag.encode("ascii")
    if typ[0] == "B":
        arr = np.asarray(val, dtype=_TAG_ARRAYS[typ[1]])
        return head + b"B" + typ[1].encode() + struct.pack("<I", arr.size) + arr.tobytes()
    if typ in _TAG_SCALARS:
        return head + typ.encode() + struct.pack(_TAG_SCALARS[typ], val)
    if typ == "A":
        return head + b"A" + val.encode("ascii")
    return head + typ.encode() + val.encode("ascii") + b"\x00"


def write_bam(path, header_text, references, records):
    """Writes an (unsorted, unindexed) BAM.  ``references``: [(name, length)]; ``records``: iterable of
    dicts with keys query_name, flag, reference_id, reference_start, mapping_quality, cigartuples,
    query_sequence, tags [(tag, BAM type, value)] (missing keys = unmapped defaults)."""
    text = header_text.encode()
    body = bytearray(b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(references)))
    for name, length in references:
        body += struct.pack("<i", len(name) + 1) + name.encode() + b"\x00" + struct.pack("<i", length)
    lut = np.full(256, 15, dtype=np.uint8)  # 4-bit base codes by ASCII value ("=ACMGRSVTWYHKDBN"), N for the rest
    for i, c in enumerate(_SEQ_CODES):
        lut[ord(c)] = i
    for rec in records:
        seq = rec.get("query_sequence", "")
        name = rec["query_name"].encode() + b"\x00"
        cigar = rec.get("cigartuples", [])
        nib = lut[np.frombuffer(seq.encode("latin-1", "replace"), dtype=np.uint8)]
        if nib.size % 2:
            nib = np.append(nib, np.uint8(0))
        packed = ((nib[0::2] << 4) | nib[1::2]).tobytes()
        tags = b"".join(_encode_tag(t, ty, v) for t, ty, v in rec.get("tags", []))
        core = struct.pack("<iiBBHHHIiii", rec.get("reference_id", -1), rec.get("reference_start", -1),
                           len(name), rec.get("mapping_quality", 0), 4680, len(cigar), rec.get("flag", 4),
                           len(seq), rec.get("next_reference_id", -1), rec.get("next_reference_start", -1),
                           rec.get("template_length", 0))
        qual = rec.get("query_qualities")
        qual = (np.asarray(qual, dtype=np.uint8).tobytes() if qual is not None and len(qual) == len(seq)
                else b"\xff" * len(seq))
        blob = (core + name + b"".join(struct.pack("<I", (ln << 4) | op) for op, ln in cigar) + packed
                + qual + tags)
        body += struct.pack("<i", len(blob)) + blob
    with open(path, "wb") as fh:
        for st in range(0, len(body), 0xFF00):
            fh.write(_bgzf_block(bytes(body[st:st + 0xFF00])))
        fh.write(_BGZF_EOF)


def references_from_header(header_text):
    """[(name, length)] from the @SQ lines of a SAM header."""
    refs = []
    for line in header_text.splitlines():
        if line.startswith("@SQ"):
            fields = dict(f.split(":", 1) for f in line.split("\t")[1:] if ":" in f)
            refs.append((fields.get("SN", "*"), int(fields.get("LN", 0))))
    return refs
