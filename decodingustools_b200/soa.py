"""Packed structure-of-arrays for decoded BAM records (the "column batch").

This is the input format of the GPU path named in BASELINE.json's north_star: the host decodes
a coordinate-sorted BAM into columns (pos, flag, MAPQ, CIGAR ops, base qualities) and streams
them to the device.  Field meanings follow the BAM record fields the reference reads through
rust-htslib (`record.mapq()`, `record.qual()`, CIGAR via the pileup engine; call sites
/root/reference/src/callable_loci/mod.rs:22-37).

Columns (n = number of records, all little-endian, C-contiguous):
    pos        int32[n]     0-based leftmost reference coordinate
    flag       uint16[n]    BAM FLAG
    mapq       uint8[n]     MAPQ
    cigar_off  uint32[n+1]  prefix offsets into ``cigar``
    cigar      uint32[*]    BAM-encoded ops: len << 4 | op, op in MIDNSHP=X (0..8)
    qual_off   uint64[n+1]  prefix byte offsets into ``qual`` (qual_off[i+1]-qual_off[i] = l_qseq)
    qual       uint8[*]     raw phred bytes
    name_id    uint32[n]    interned QNAME id (host only: never uploaded; mates share an id)
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import Iterable, Optional, Sequence

import numpy as np

CIGAR_OPS = "MIDNSHP=X"
_CIGAR_RE = re.compile(r"(\d+)([MIDNSHP=X])")

# op -> consumes reference / query (SAM spec; htslib bam_cigar_type)
CONSUMES_REF = np.array([1, 0, 1, 1, 0, 0, 0, 1, 1], dtype=np.uint8)
CONSUMES_QUERY = np.array([1, 1, 0, 0, 1, 0, 0, 1, 1], dtype=np.uint8)

FLAG_UNMAP = 0x4


def encode_cigar(text: str) -> np.ndarray:
    """'2M1D2M' -> uint32 BAM ops.  '*' or '' -> empty."""
    if text in ("", "*"):
        return np.zeros(0, dtype=np.uint32)
    ops = _CIGAR_RE.findall(text)
    if "".join(a + b for a, b in ops) != text:
        raise ValueError(f"bad CIGAR {text!r}")
    return np.array([(int(n) << 4) | CIGAR_OPS.index(o) for n, o in ops], dtype=np.uint32)


def decode_cigar(ops: np.ndarray) -> str:
    return "".join(f"{int(v) >> 4}{CIGAR_OPS[int(v) & 15]}" for v in ops) or "*"


@dataclass
class ReadColumns:
    pos: np.ndarray
    flag: np.ndarray
    mapq: np.ndarray
    cigar_off: np.ndarray
    cigar: np.ndarray
    qual_off: np.ndarray
    qual: np.ndarray
    name_id: Optional[np.ndarray] = None

    def __post_init__(self):
        self.pos = np.ascontiguousarray(self.pos, dtype=np.int32)
        self.flag = np.ascontiguousarray(self.flag, dtype=np.uint16)
        self.mapq = np.ascontiguousarray(self.mapq, dtype=np.uint8)
        self.cigar_off = np.ascontiguousarray(self.cigar_off, dtype=np.uint32)
        self.cigar = np.ascontiguousarray(self.cigar, dtype=np.uint32)
        self.qual_off = np.ascontiguousarray(self.qual_off, dtype=np.uint64)
        self.qual = np.ascontiguousarray(self.qual, dtype=np.uint8)
        if self.name_id is not None:
            self.name_id = np.ascontiguousarray(self.name_id, dtype=np.uint32)
        n = self.pos.shape[0]
        if not (self.flag.shape[0] == n and self.mapq.shape[0] == n
                and self.cigar_off.shape[0] == n + 1 and self.qual_off.shape[0] == n + 1):
            raise ValueError("ReadColumns: column lengths disagree")
        if n and (int(self.cigar_off[-1]) != self.cigar.shape[0] or int(self.qual_off[-1]) != self.qual.shape[0]):
            raise ValueError("ReadColumns: offset columns do not end at the payload length")

    # ------------------------------------------------------------------ basics
    @property
    def n(self) -> int:
        return int(self.pos.shape[0])

    @property
    def n_cigar(self) -> int:
        return int(self.cigar.shape[0])

    @property
    def n_qual(self) -> int:
        return int(self.qual.shape[0])

    def nbytes_device(self) -> int:
        """Bytes of the columns that travel to the GPU (SURVEY.md section 8(d) accounting)."""
        return int(self.pos.nbytes + self.flag.nbytes + self.mapq.nbytes + self.cigar_off.nbytes
                   + self.cigar.nbytes + self.qual_off.nbytes + self.qual.nbytes)

    @staticmethod
    def empty() -> "ReadColumns":
        return ReadColumns(np.zeros(0, np.int32), np.zeros(0, np.uint16), np.zeros(0, np.uint8),
                           np.zeros(1, np.uint32), np.zeros(0, np.uint32), np.zeros(1, np.uint64),
                           np.zeros(0, np.uint8), np.zeros(0, np.uint32))

    @staticmethod
    def from_records(records: Iterable[Sequence]) -> "ReadColumns":
        """records: (pos, flag, mapq, cigar_text, quals(list|bytes|int-for-constant), name) tuples.

        ``quals`` may be an int q, meaning "query-length bytes of value q".
        """
        pos, flag, mapq, names = [], [], [], []
        cig, cig_off, quals, q_off = [], [0], [], [0]
        name_ids: dict = {}
        for rec in records:
            p, f, m, c, q = rec[:5]
            nm = rec[5] if len(rec) > 5 else f"r{len(pos)}"
            ops = encode_cigar(c)
            qlen = int(sum(int(v) >> 4 for v in ops if CONSUMES_QUERY[int(v) & 15]))
            if isinstance(q, (int, np.integer)):
                qb = np.full(qlen, int(q), dtype=np.uint8)
            else:
                qb = np.frombuffer(bytes(q), dtype=np.uint8) if isinstance(q, (bytes, bytearray)) else np.asarray(q, dtype=np.uint8)
            pos.append(p); flag.append(f); mapq.append(m)
            cig.append(ops); cig_off.append(cig_off[-1] + len(ops))
            quals.append(qb); q_off.append(q_off[-1] + len(qb))
            names.append(name_ids.setdefault(nm, len(name_ids)))
        return ReadColumns(
            np.array(pos, np.int32), np.array(flag, np.uint16), np.array(mapq, np.uint8),
            np.array(cig_off, np.uint32), np.concatenate(cig) if cig else np.zeros(0, np.uint32),
            np.array(q_off, np.uint64), np.concatenate(quals) if quals else np.zeros(0, np.uint8),
            np.array(names, np.uint32))

    # ------------------------------------------------------------------ derived columns
    def ref_len(self) -> np.ndarray:
        """Reference span of each record (sum over M, D, N, =, X) == htslib bam_cigar2rlen."""
        if self.n == 0:
            return np.zeros(0, np.int64)
        contrib = (self.cigar >> 4).astype(np.int64) * CONSUMES_REF[self.cigar & 15]
        csum = np.concatenate([[0], np.cumsum(contrib)])
        return csum[self.cigar_off[1:].astype(np.int64)] - csum[self.cigar_off[:-1].astype(np.int64)]

    def end(self) -> np.ndarray:
        return self.pos.astype(np.int64) + self.ref_len()

    def max_ref_span(self) -> int:
        rl = self.ref_len()
        return int(rl.max()) if rl.size else 0

    # ------------------------------------------------------------------ slicing / compaction
    def select(self, mask_or_index: np.ndarray) -> "ReadColumns":
        """Compacted copy of the chosen records (boolean mask or sorted index array)."""
        idx = np.asarray(mask_or_index)
        if idx.dtype == np.bool_:
            idx = np.flatnonzero(idx)
        idx = idx.astype(np.int64)
        if idx.size == 0:
            return ReadColumns.empty()
        c0 = self.cigar_off[:-1].astype(np.int64)[idx]; c1 = self.cigar_off[1:].astype(np.int64)[idx]
        q0 = self.qual_off[:-1].astype(np.int64)[idx]; q1 = self.qual_off[1:].astype(np.int64)[idx]
        return ReadColumns(
            self.pos[idx], self.flag[idx], self.mapq[idx],
            np.concatenate([[0], np.cumsum(c1 - c0)]), _gather_ranges(self.cigar, c0, c1),
            np.concatenate([[0], np.cumsum(q1 - q0)]), _gather_ranges(self.qual, q0, q1),
            None if self.name_id is None else self.name_id[idx])

    def slice(self, lo: int, hi: int) -> "ReadColumns":
        """Contiguous record range [lo, hi) with rebased offsets (cheap: payload is a view copy)."""
        lo = max(0, int(lo)); hi = min(self.n, int(hi))
        if hi <= lo:
            return ReadColumns.empty()
        c0, c1 = int(self.cigar_off[lo]), int(self.cigar_off[hi])
        q0, q1 = int(self.qual_off[lo]), int(self.qual_off[hi])
        return ReadColumns(
            self.pos[lo:hi], self.flag[lo:hi], self.mapq[lo:hi],
            self.cigar_off[lo:hi + 1] - np.uint32(c0), self.cigar[c0:c1],
            self.qual_off[lo:hi + 1] - np.uint64(q0), self.qual[q0:q1],
            None if self.name_id is None else self.name_id[lo:hi])


def _gather_ranges(payload: np.ndarray, starts: np.ndarray, ends: np.ndarray) -> np.ndarray:
    lens = ends - starts
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, dtype=payload.dtype)
    out_off = np.concatenate([[0], np.cumsum(lens)])[:-1]
    # index = start[i] + (k - out_off[i]) for k in output range of record i
    rec = np.repeat(np.arange(len(lens)), lens)
    k = np.arange(total, dtype=np.int64)
    return payload[starts[rec] + (k - out_off[rec])]


def n_mask_from_ascii(ref: np.ndarray | bytes) -> np.ndarray:
    """Bit-packed N-mask (bit p&31 of word p>>5 set iff ref[p] in {N,n}), padded to uint32 words.

    REF_N is decided only by these two byte values (callable_profiler.rs:104)."""
    a = np.frombuffer(ref, dtype=np.uint8) if isinstance(ref, (bytes, bytearray)) else np.asarray(ref, dtype=np.uint8)
    isn = (a == ord("N")) | (a == ord("n"))
    nwords = (a.shape[0] + 31) // 32 + 1
    bits = np.zeros(nwords * 32, dtype=np.uint8)
    bits[: a.shape[0]] = isn
    return np.packbits(bits, bitorder="little").view(np.uint32).copy()
