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


def write_bgzf(path: str, payload: bytes, block: int = 0xFF00) -> List[int]:
    """Returns the file offset of every block (the EOF marker block included)."""
    starts = []
    with open(path, "wb") as f:
        for o in range(0, len(payload), block):
            starts.append(f.tell())
            f.write(_bgzf_block(payload[o:o + block]))
        starts.append(f.tell())
        f.write(_bgzf_block(b""))                # EOF marker
    return starts


def _reg2bin(beg: int, end: int) -> int:
    end -= 1
    for shift, level in ((14, 15), (17, 12), (20, 9), (23, 6), (26, 3)):
        if beg >> shift == end >> shift:
            return ((1 << level) - 1) // 7 + (beg >> shift)
    return 0


def write_bai(path: str, n_ref: int, records, block_starts: List[int], block: int, n_no_coor: int = 0):
    """records: (tid, beg, end, payload_offset_start, payload_offset_end) of every mapped record in file order."""
    voff = lambda o: (block_starts[o // block] << 16) | (o % block)
    out = bytearray(b"BAI\1" + struct.pack("<i", n_ref))
    for tid in range(n_ref):
        bins, linear, first, last, n = {}, {}, None, None, 0
        for t, beg, end, o0, o1 in records:
            if t != tid:
                continue
            v0, v1 = voff(o0), voff(o1)
            chunks = bins.setdefault(_reg2bin(beg, max(end, beg + 1)), [])
            if chunks and chunks[-1][1] == v0:
                chunks[-1][1] = v1                               # consecutive records of one bin share a chunk
            else:
                chunks.append([v0, v1])
            for w in range(beg >> 14, (max(end, beg + 1) - 1 >> 14) + 1):
                linear[w] = min(linear.get(w, v0), v0)
            first = v0 if first is None else first
            last = v1; n += 1
        if n:
            bins[37450] = [[first, last], [n, 0]]                # metadata pseudo-bin
        out += struct.pack("<i", len(bins))
        for b, chunks in bins.items():
            out += struct.pack("<Ii", b, len(chunks)) + b"".join(struct.pack("<QQ", c0, c1) for c0, c1 in chunks)
        n_intv = max(linear) + 1 if linear else 0
        out += struct.pack("<i", n_intv)
        prev = 0
        for w in range(n_intv):
            prev = linear.get(w, prev)
            out += struct.pack("<Q", prev)
    out += struct.pack("<Q", n_no_coor)
    with open(path, "wb") as f:
        f.write(out)


def default_qname(contig: str, name_id: int) -> str:
    return f"{contig}:q{name_id}"


def write_bam(path: str, contigs: Sequence, header_extra: str = "@PG\tID:bwa\tPN:bwa\n", qname_fn=default_qname, block: int = 0xFF00,
              cg_threshold: int = 65535, unmapped_tail: int = 0, index: bool = False):
    """contigs: list of (name, length, ReadColumns) in tid order; QNAMEs are synthesised from name_id (mates share)
    through qname_fn(contig, name_id).  block: uncompressed bytes per BGZF block (small values make records straddle blocks).
    Reads with more than cg_threshold CIGAR ops are stored the way htslib stores them: placeholder CIGAR <l_seq>S<ref_len>N plus
    a CG:B,I tag (an NM:i tag and an RG:Z tag are put in front of it so the aux walk is exercised).  unmapped_tail appends that
    many unmapped records (tid -1, pos -1, flag 4) after the last contig.  index: also write path + ".bai"."""
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l, _ in contigs) + header_extra
    out = bytearray(b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(contigs)))
    for n, l, _ in contigs:
        nb = n.encode() + b"\0"
        out += struct.pack("<i", len(nb)) + nb + struct.pack("<i", l)
    recs = []
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
            span = int(sum(int(v) >> 4 for v in rc.cigar[c0:c1] if (int(v) & 15) in (0, 2, 3, 7, 8)))
            recs.append((tid, int(rc.pos[i]), int(rc.pos[i]) + span, len(out), len(out) + 4 + len(body)))
            out += struct.pack("<i", len(body)) + body
    for i in range(unmapped_tail):
        qn = f"unmapped{i}".encode() + b"\0"
        body = struct.pack("<iiBBHHHiiii", -1, -1, len(qn), 0, 4680, 0, 4, 10, -1, -1, 0) + qn + bytes(5) + bytes([0xFF] * 10)
        out += struct.pack("<i", len(body)) + body
    starts = write_bgzf(path, bytes(out), block)
    if index:
        write_bai(path + ".bai", len(contigs), recs, starts, block, unmapped_tail)


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


def read_bam(path: str):
    """Inverse of write_bam for the test fixtures (no CG tags): returns (header_text, [(name, length)], per-tid
    ReadColumns in file order, per-tid QNAME lists).  Records with tid < 0 are dropped."""
    raw = open(path, "rb").read()
    out = bytearray()
    o = 0
    while o < len(raw):
        assert raw[o:o + 4] == b"\x1f\x8b\x08\x04", "not a BGZF block"
        xlen = struct.unpack_from("<H", raw, o + 10)[0]
        bsize = None
        x = o + 12
        while x < o + 12 + xlen:
            si1, si2, slen = raw[x], raw[x + 1], struct.unpack_from("<H", raw, x + 2)[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack_from("<H", raw, x + 4)[0]
            x += 4 + slen
        assert bsize is not None
        body = raw[o + 12 + xlen:o + bsize + 1 - 8]
        out += zlib.decompress(body, -15)
        o += bsize + 1
    buf = bytes(out)
    assert buf[:4] == b"BAM\1"
    l_text = struct.unpack_from("<i", buf, 4)[0]
    text = buf[8:8 + l_text].decode()
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", buf, p)[0]; p += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", buf, p)[0]; p += 4
        name = buf[p:p + l_name - 1].decode(); p += l_name
        refs.append((name, struct.unpack_from("<i", buf, p)[0])); p += 4
    recs = [[] for _ in range(n_ref)]
    names = [[] for _ in range(n_ref)]
    while p < len(buf):
        bs = struct.unpack_from("<i", buf, p)[0]; p += 4
        tid, pos, l_qn, mapq, _bin, n_cig, flag, l_seq = struct.unpack_from("<iiBBHHHi", buf, p)
        q = p + 32
        qn = buf[q:q + l_qn - 1].decode(); q += l_qn
        cig = np.frombuffer(buf, dtype="<u4", count=n_cig, offset=q); q += 4 * n_cig
        q += (l_seq + 1) // 2
        qual = np.frombuffer(buf, dtype=np.uint8, count=l_seq, offset=q)
        p += bs
        if tid >= 0:
            recs[tid].append((pos, flag, mapq, cig, qual)); names[tid].append(qn)
    cols = []
    for tid in range(n_ref):
        r = recs[tid]
        ids: dict = {}
        nid = np.array([ids.setdefault(n, len(ids)) for n in names[tid]], np.uint32)
        cols.append(ReadColumns(
            np.array([x[0] for x in r], np.int32), np.array([x[1] for x in r], np.uint16), np.array([x[2] for x in r], np.uint8),
            np.concatenate([[0], np.cumsum([len(x[3]) for x in r])]).astype(np.uint32),
            np.concatenate([x[3] for x in r]).astype(np.uint32) if r else np.zeros(0, np.uint32),
            np.concatenate([[0], np.cumsum([len(x[4]) for x in r])]).astype(np.uint64),
            np.concatenate([x[4] for x in r]).astype(np.uint8) if r else np.zeros(0, np.uint8), nid))
    return text, refs, cols, names


def read_fasta(path: str):
    seqs, name, parts = [], None, []
    for ln in open(path, "rb").read().split(b"\n"):
        if ln.startswith(b">"):
            if name is not None:
                seqs.append((name, b"".join(parts)))
            name, parts = ln[1:].split()[0].decode(), []
        elif ln:
            parts.append(ln)
    if name is not None:
        seqs.append((name, b"".join(parts)))
    return seqs


def pack_bam_fast(path: str, name: str, length: int, reads: ReadColumns, workdir: str, threads: int = 0):
    """A single-contig BAM through the C++ packer (decodingustools_b200/clb-pack-bam): columns are dumped as raw files,
    records assembled and BGZF-compressed on several threads.  QNAMEs: A00123:7:HFLOWCELLX:1:1101:<contig>:<name_id>."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "decodingustools_b200", "clb-pack-bam")
    os.makedirs(workdir, exist_ok=True)
    nid = reads.name_id if reads.name_id is not None else np.arange(reads.n, dtype=np.uint32)
    for fn, arr, dt in (("pos.i32", reads.pos, "<i4"), ("flag.u16", reads.flag, "<u2"), ("mapq.u8", reads.mapq, "u1"), ("cigar_off.u32", reads.cigar_off, "<u4"),
                        ("cigar.u32", reads.cigar, "<u4"), ("qual_off.u64", reads.qual_off, "<u8"), ("qual.u8", reads.qual, "u1"), ("name_id.u32", nid, "<u4")):
        np.ascontiguousarray(arr, dtype=dt).tofile(os.path.join(workdir, fn))
    subprocess.check_call([exe, workdir, name, str(int(length)), path] + ([str(threads)] if threads else []))
