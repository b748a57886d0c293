"""Minimal BGZF/BAM + FASTA/.fai writers for tests (test infrastructure: lets the CLI be driven with real files)."""
from __future__ import annotations

import struct
import zlib
from typing import List, Sequence

import numpy as np

from decodingustools_b200.soa import ReadColumns


def _bgzf_block(data: bytes) -> bytes:
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = comp.compress(data) + comp.flush()
    bsize = len(body) + 25                       # total block length - 1
    hdr = struct.pack("<BBBBIBBHBBHH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6, ord("B"), ord("C"), 2, bsize)
    return hdr + body + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data))


def write_bgzf(path: str, payload: bytes, block: int = 0xFF00):
    with open(path, "wb") as f:
        for o in range(0, len(payload), block):
            f.write(_bgzf_block(payload[o:o + block]))
        f.write(_bgzf_block(b""))                # EOF marker


def default_qname(contig: str, name_id: int) -> str:
    return f"{contig}:q{name_id}"


def write_bam(path: str, contigs: Sequence, header_extra: str = "@PG\tID:bwa\tPN:bwa\n", qname_fn=default_qname, block: int = 0xFF00,
              cg_threshold: int = 65535, unmapped_tail: int = 0):
    """contigs: list of (name, length, ReadColumns) in tid order; QNAMEs are synthesised from name_id (mates share)
    through qname_fn(contig, name_id).  block: uncompressed bytes per BGZF block (small values make records straddle blocks).
    Reads with more than cg_threshold CIGAR ops are stored the way htslib stores them: placeholder CIGAR <l_seq>S<ref_len>N plus
    a CG:B,I tag (an NM:i tag and an RG:Z tag are put in front of it so the aux walk is exercised).  unmapped_tail appends that
    many unmapped records (tid -1, pos -1, flag 4) after the last contig."""
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l, _ in contigs) + header_extra
    out = bytearray(b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(contigs)))
    for n, l, _ in contigs:
        nb = n.encode() + b"\0"
        out += struct.pack("<i", len(nb)) + nb + struct.pack("<i", l)
    for tid, (name, _, rc) in enumerate(contigs):
        for i in range(rc.n):
            qn = qname_fn(name, int(rc.name_id[i]) if rc.name_id is not None else i).encode() + b"\0"
            c0, c1 = int(rc.cigar_off[i]), int(rc.cigar_off[i + 1])
            q0, q1 = int(rc.qual_off[i]), int(rc.qual_off[i + 1])
            lseq = q1 - q0
            cigar = rc.cigar[c0:c1].astype("<u4")
            aux = b""
            if c1 - c0 > cg_threshold:
                ref_len = int(sum(int(v) >> 4 for v in cigar if (int(v) & 15) in (0, 2, 3, 7, 8)))
                aux = b"NMi" + struct.pack("<i", 3) + b"RGZgrp1\0" + b"CGBI" + struct.pack("<I", c1 - c0) + cigar.tobytes()
                cigar = np.array([(lseq << 4) | 4, (ref_len << 4) | 3], dtype="<u4")
            body = struct.pack("<iiBBHHHiiii", tid, int(rc.pos[i]), len(qn), int(rc.mapq[i]), 4680, len(cigar), int(rc.flag[i]),
                               lseq, -1, -1, 0)
            body += qn + cigar.tobytes() + bytes((lseq + 1) // 2) + rc.qual[q0:q1].tobytes() + aux
            out += struct.pack("<i", len(body)) + body
    for i in range(unmapped_tail):
        qn = f"unmapped{i}".encode() + b"\0"
        body = struct.pack("<iiBBHHHiiii", -1, -1, len(qn), 0, 4680, 0, 4, 10, -1, -1, 0) + qn + bytes(5) + bytes([0xFF] * 10)
        out += struct.pack("<i", len(body)) + body
    write_bgzf(path, bytes(out), block)


def write_fasta(path: str, contigs: Sequence, width: int = 60):
    """contigs: list of (name, seq_bytes).  Writes path and path + '.fai'."""
    off = 0
    fai: List[str] = []
    with open(path, "wb") as f:
        for name, seq in contigs:
            seq = bytes(seq)
            hdr = f">{name}\n".encode()
            f.write(hdr); off += len(hdr)
            fai.append(f"{name}\t{len(seq)}\t{off}\t{width}\t{width + 1}\n")
            for o in range(0, len(seq), width):
                line = seq[o:o + width] + b"\n"
                f.write(line); off += len(line)
    with open(path + ".fai", "w") as f:
        f.writelines(fai)
