"""Host-side helpers on the hot path's boundary, mirroring the names and behaviour of the
reference's ``remora.util`` (sequence coding, motifs, softmax, MM/ML tag formatting).

Written from the reference's behaviour (file:line cited per function), not copied.
"""
import array
import re
from dataclasses import dataclass

import numpy as np

from . import RemoraError

CAN_ALPHABET = "ACGT"
# IUPAC single-letter codes (reference src/remora/util.py:24-42)
IUPAC = {
    "A": "A", "C": "C", "G": "G", "T": "T", "B": "CGT", "D": "AGT", "H": "ACT", "K": "GT",
    "M": "AC", "N": "ACGT", "R": "AG", "S": "CG", "V": "ACG", "W": "AT", "Y": "CT",
}
_BASE_CODE = np.full(256, -1, dtype=np.int64)
for _i, _b in enumerate(CAN_ALPHABET):
    _BASE_CODE[ord(_b)] = _i


def seq_to_int(seq):
    """String -> int array, A0 C1 G2 T3, anything else -1 (reference util.py:131-142; the
    reference only accepts upper-case A-Z, so do we)."""
    raw = np.frombuffer(seq.encode("ascii"), dtype=np.uint8)
    if raw.size and (raw.min() < ord("A") or raw.max() > ord("Z")):
        raise IndexError("sequence contains characters outside A-Z")
    return _BASE_CODE[raw]


CONV_ALPHABET = "ACGTN"


def int_to_seq(np_seq, alphabet=CONV_ALPHABET):
    """Int array -> string; -1 maps to 'N' because the reference indexes the alphabet string
    'ACGTN' with -1 (util.py:26,145-158)."""
    np_seq = np.asarray(np_seq)
    if np_seq.shape[0] == 0:
        return ""
    if np_seq.max() >= len(alphabet):
        raise RemoraError(f"Invalid value in int sequence ({np_seq.max()})")
    lut = np.frombuffer(alphabet.encode("ascii"), dtype=np.uint8)
    return lut[np_seq].tobytes().decode("ascii")


_COMP = str.maketrans("ACGTBVDHKMRY", "TGCAVBHDMKYR")  # same letters as reference util.py:51


def revcomp(seq):
    """Reverse complement, IUPAC aware, upper-cased first (reference util.py:102-106)."""
    return seq.upper().translate(_COMP)[::-1]


def softmax_axis1(x):
    """Row softmax (reference util.py:182-186), same operation order."""
    shifted = x - np.max(x, axis=1, keepdims=True)
    e_x = np.exp(shifted)
    with np.errstate(divide="ignore"):
        return e_x / e_x.sum(axis=1, keepdims=True)


@dataclass
class Motif:
    """Sequence motif with IUPAC ambiguity codes and a focus position
    (reference util.py:189-312).  Leading/trailing N are clipped like the reference does."""

    raw_motif: str
    focus_pos: int = 0

    def __post_init__(self):
        try:
            self.focus_pos = int(self.focus_pos)
        except ValueError:
            raise RemoraError(f'Motif focus position not an integer: "{self.focus_pos}"')
        if not isinstance(self.raw_motif, str):
            raise RemoraError("Motif sequence must be a string")
        bad = set(self.raw_motif) - set(IUPAC)
        if bad:
            raise RemoraError(f"Motif contains invalid characters: {bad}")
        if self.focus_pos >= len(self.raw_motif):
            raise RemoraError("Motif focus position is past the end of the motif")
        lead = len(self.raw_motif) - len(self.raw_motif.lstrip("N"))
        lead = min(lead, len(self.raw_motif) - 1)
        self.raw_motif = self.raw_motif[lead:]
        self.focus_pos -= lead
        stripped = self.raw_motif.rstrip("N")
        self.raw_motif = stripped if stripped else self.raw_motif[:1]

    def to_tuple(self):
        return self.raw_motif, self.focus_pos

    def __hash__(self):
        return hash(self.to_tuple())

    @property
    def focus_base(self):
        return self.raw_motif[self.focus_pos]

    @property
    def num_bases_after_focus(self):
        return len(self.raw_motif) - self.focus_pos - 1

    @property
    def pattern(self):
        body = "".join(f"[{IUPAC[c]}]" for c in self.raw_motif)
        return re.compile(f"(?=({body}))")

    def _allowed(self):
        """bool [motif_len][4]: which canonical bases each motif position accepts."""
        table = np.zeros((len(self.raw_motif), 4), dtype=bool)
        for i, letter in enumerate(self.raw_motif):
            for b in IUPAC[letter]:
                table[i, CAN_ALPHABET.index(b)] = True
        return table

    def findall(self, int_seq):
        """Start index of every (overlapping) motif hit in an int-coded sequence
        (reference util.py:281-297); -1 (N) never matches."""
        int_seq = np.asarray(int_seq)
        m = len(self.raw_motif)
        n_pos = int_seq.size - m + 1
        if n_pos <= 0:
            return np.zeros(0, dtype=np.int64)
        allowed = self._allowed()
        valid = (int_seq >= 0) & (int_seq < 4)
        safe = np.where(valid, int_seq, 0)
        hit = np.ones(n_pos, dtype=bool)
        for off in range(m):
            window = slice(off, off + n_pos)
            hit &= valid[window] & allowed[off][safe[window]]
        return np.flatnonzero(hit)

    def match(self, int_seq, pos):
        """Does the motif match with its focus base at ``pos``?"""
        start = pos - self.focus_pos
        if start < 0 or start + len(self.raw_motif) > len(int_seq):
            return False
        allowed = self._allowed()
        for off in range(len(self.raw_motif)):
            b = int(int_seq[start + off])
            if b < 0 or b > 3 or not allowed[off, b]:
                return False
        return True


def find_focus_bases_in_int_sequence(int_seq, motifs):
    """Positions (focus base index) of every hit of any motif (reference util.py:413-426).
    The reference returns them in set-iteration order; this returns them sorted, which is the
    order every downstream consumer establishes anyway (format_mm_ml_tags sorts)."""
    hits = [m.findall(int_seq) + m.focus_pos for m in motifs]
    if not hits:
        return np.zeros(0, dtype=int)
    return np.unique(np.concatenate(hits)).astype(int)


def format_mm_ml_tags(seq, poss, probs, mod_bases, can_base, strand="+", ml_bytes=None):
    """MM string + ML uint8 array for SAM/BAM (reference util.py:485-537), vectorised: positions sorted
    (stable, like the reference's sort on position), one MM section per modified base in ``mod_bases``
    order, gaps counted in canonical bases.  ML byte = floor(p*256) with 256 -> 255 (util.py:532-535).
    A ``None`` row of ``probs`` skips that position as in the reference.  ``ml_bytes`` (extension):
    uint8 [N, n_mods] already quantised on the device (``B200Model.softmax_ml``); ``probs`` is then
    not needed."""
    mm_tag, ml_tag = "", array.array("B")
    poss = np.asarray(poss, dtype=np.int64)
    if poss.size == 0:
        return mm_tag, ml_tag
    if ml_bytes is not None:
        probs = np.asarray(ml_bytes).reshape(poss.size, -1)
    elif isinstance(probs, np.ndarray) and probs.dtype != object:
        probs = probs.reshape(poss.size, -1).astype(np.float64)
    else:
        keep = np.array([p is not None for p in probs], dtype=bool)
        poss = poss[keep]
        probs = np.array([p for p in probs if p is not None], dtype=np.float64).reshape(poss.size, -1)
        if poss.size == 0:
            return mm_tag, ml_tag
    order = np.argsort(poss, kind="stable")
    mod_pos, probs = poss[order], probs[order]
    can_count = np.cumsum(np.frombuffer(seq.encode("ascii"), dtype=np.uint8) == ord(can_base))
    can_idx = can_count[mod_pos] - 1
    gaps = np.diff(np.concatenate([[-1], can_idx])) - 1
    gap_txt = ",".join(map(str, gaps.tolist()))
    if ml_bytes is not None:
        scaled = probs.astype(np.uint8)
    else:
        scaled = np.floor(probs * 256)
        scaled[scaled == 256] = 255
        scaled = scaled.astype(np.uint8)
    for col, mb in enumerate(mod_bases):
        if col >= scaled.shape[1]:
            break
        mm_tag += f"{can_base}{strand}{mb}?,{gap_txt};"
        ml_tag.extend(scaled[:, col].tolist())
    return mm_tag, ml_tag
