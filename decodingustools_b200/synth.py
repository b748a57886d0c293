"""Seeded synthetic inputs for the `coverage` hot path (BASELINE.json configs).

There is no network and the reference ships no fixtures, so every workload is synthetic: a reference
sequence with REF_N blocks and a coordinate-sorted read set emitted directly as packed columns
(:class:`decodingustools_b200.soa.ReadColumns`), i.e. what a BAM decoder would hand to the device.

Short-read mode (configs 1-4): 2x150 bp pairs, fragment starts uniform (Poisson in the limit) thinned by
coverage zones, insert ~N(400,50); CIGAR mix 94 % ``150M``, 3 % soft-clipped, 1.5 % one insertion,
1.5 % one deletion; MAPQ mix 88 % 60 / 6 % 0 / 1 % 1 / 5 % uniform 2..59 plus "repeat" blocks where half
the reads get MAPQ 0; base qualities binned {2,12,23,37} with p = {.01,.03,.06,.90}; flag noise
(dup, secondary, supplementary, QC-fail, placed-unmapped).
Long-read mode (config 5): length ~N(15 kb, 3 kb), an indel op every 20-50 bp, 5 % soft-clipped ends.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

from .soa import ReadColumns

_QUAL_LUT = np.empty(256, dtype=np.uint8)
_QUAL_LUT[:3] = 2        # ~.012
_QUAL_LUT[3:11] = 12     # ~.031
_QUAL_LUT[11:26] = 23    # ~.059
_QUAL_LUT[26:] = 37      # ~.898


@dataclass
class SynthContig:
    name: str
    length: int
    ref: np.ndarray            # uint8 ASCII, len = length
    reads: ReadColumns         # coordinate sorted, unfiltered
    n_blocks: List[Tuple[int, int]]


def _scaled(size: float, length: int, lo: int) -> int:
    return max(lo, int(size * min(1.0, length / 20e6)))


def make_reference(length: int, rng: np.random.Generator):
    ref = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=length, dtype=np.uint8)]
    blocks: List[Tuple[int, int]] = []
    if length >= 64:
        tel = min(_scaled(10_000, length, 4), length // 8)
        blocks += [(0, tel), (length - tel, length)]
        cen = max(4, int(length * 0.012))
        c0 = int(length * 0.49)
        blocks.append((c0, min(length, c0 + cen)))
        for _ in range(20 if length >= 100_000 else 3):
            sz = _scaled(rng.integers(1_000, 50_000), length, 2)
            s = int(rng.integers(0, max(1, length - sz)))
            blocks.append((s, s + sz))
    for s, e in blocks:
        ref[s:e] = ord("N")
    if length >= 64:
        for _ in range(5):                                  # lowercase n runs count as REF_N too
            s = int(rng.integers(0, length - 8)); ref[s:s + int(rng.integers(1, 8))] = ord("n")
        iupac = np.frombuffer(b"RYKMacgt", dtype=np.uint8)
        idx = rng.integers(0, length, size=20)
        ref[idx] = iupac[rng.integers(0, len(iupac), size=20)]
        for s, e in blocks[:2]:                              # keep the telomere blocks pure N
            ref[s:e] = ord("N")
    return ref, sorted(blocks)


def _zones(length: int, frac: float, lo: int, hi: int, rng) -> np.ndarray:
    """Random intervals covering ~frac of the contig; returns int64[k,2] sorted, possibly overlapping."""
    lo_s, hi_s = _scaled(lo, length, 8), _scaled(hi, length, 16)
    mean = (lo_s + hi_s) / 2
    k = max(1, int(length * frac / mean))
    starts = np.sort(rng.integers(0, max(1, length - hi_s), size=k))
    lens = rng.integers(lo_s, hi_s + 1, size=k)
    return np.stack([starts, starts + lens], axis=1).astype(np.int64)


def _in_zones(pos: np.ndarray, zones: np.ndarray) -> np.ndarray:
    if zones.shape[0] == 0:
        return np.zeros(pos.shape[0], dtype=bool)
    i = np.searchsorted(zones[:, 0], pos, side="right") - 1
    ok = i >= 0
    return ok & (pos < zones[np.maximum(i, 0), 1])


def _fill_quals(n_bytes: int, rng, device=None) -> np.ndarray:
    if device is not None:
        import torch
        g = torch.Generator(device=device); g.manual_seed(int(rng.integers(0, 2**31)))
        lut = torch.from_numpy(_QUAL_LUT).to(device)
        out = np.empty(n_bytes, dtype=np.uint8)
        step = 1 << 30
        for s in range(0, n_bytes, step):
            m = min(step, n_bytes - s)
            r = torch.randint(0, 256, (m,), dtype=torch.uint8, device=device, generator=g)
            out[s:s + m] = lut[r.long()].cpu().numpy()
        return out
    out = np.empty(n_bytes, dtype=np.uint8)
    step = 1 << 27
    for s in range(0, n_bytes, step):
        m = min(step, n_bytes - s)
        out[s:s + m] = _QUAL_LUT[rng.integers(0, 256, size=m, dtype=np.uint8)]
    return out


def _mapq_flags(n: int, pos: np.ndarray, repeat_zones: np.ndarray, rng):
    u = rng.random(n)
    mapq = np.full(n, 60, dtype=np.uint8)
    mapq[u < 0.12] = rng.integers(2, 60, size=int((u < 0.12).sum()), dtype=np.uint8)   # 5 % uniform 2..59
    mapq[u < 0.07] = 1
    mapq[u < 0.06] = 0
    rep = _in_zones(pos, repeat_zones) & (rng.random(n) < 0.5)
    mapq[rep] = 0
    f = rng.random(n)
    flag = np.zeros(n, dtype=np.uint16)
    flag[f < 0.012] = 0x4         # placed-unmapped (must be skipped)   0.1 %
    flag[f < 0.011] = 0x200       # QC fail                              0.1 %
    flag[f < 0.010] = 0x800       # supplementary                        0.2 %
    flag[f < 0.008] = 0x100       # secondary                            0.3 %
    flag[f < 0.005] = 0x400       # duplicate                            0.5 %
    return mapq, flag


def synth_short(name: str, length: int, seed: int, depth: float = 30.0, read_len: int = 150,
                qual_device=None) -> SynthContig:
    rng = np.random.default_rng(seed)
    ref, blocks = make_reference(length, rng)
    if length < 2 * read_len + 8:
        return SynthContig(name, length, ref, ReadColumns.empty(), blocks)
    gap_z = _zones(length, 0.005, 500, 3500, rng)
    low_z = _zones(length, 0.01, 1000, 5000, rng)
    rep_z = _zones(length, 0.01, 2000, 20000, rng)
    nblk = np.array([b for b in blocks if b[1] - b[0] >= 2 * read_len] or np.zeros((0, 2)), dtype=np.int64).reshape(-1, 2)

    n_frag = int(depth * length / (2 * read_len))
    start = rng.integers(0, length - read_len, size=n_frag, dtype=np.int64)
    insert = np.maximum(read_len, rng.normal(400, 50, size=n_frag)).astype(np.int64)
    keep = ~_in_zones(start, gap_z) & ~_in_zones(start, nblk)
    lowz = _in_zones(start, low_z)
    keep &= ~lowz | (rng.random(n_frag) < rng.uniform(1 / 30, 3 / 30))
    keep &= (start + insert) <= length
    start, insert = start[keep], insert[keep]
    n_frag = start.shape[0]
    frag_id = np.arange(n_frag, dtype=np.uint32)

    pos = np.concatenate([start, start + insert - read_len])
    name_id = np.concatenate([frag_id, frag_id])
    strand = np.concatenate([np.zeros(n_frag, np.uint16), np.ones(n_frag, np.uint16)])
    order = np.argsort(pos, kind="stable")
    pos, name_id, strand = pos[order], name_id[order], strand[order]
    n = pos.shape[0]

    mapq, flag = _mapq_flags(n, pos, rep_z, rng)
    flag |= np.uint16(0x1) | np.where(strand == 1, np.uint16(0x10 | 0x80), np.uint16(0x20 | 0x40)).astype(np.uint16)

    # CIGARs: kind 0 = 150M, 1 = kS(150-k)M, 2 = (150-k)MkS, 3 = aMkIbM, 4 = aMdDbM
    u = rng.random(n)
    kind = np.zeros(n, dtype=np.int8)
    kind[u < 0.06] = 4
    kind[u < 0.045] = 3
    kind[u < 0.03] = 2
    kind[u < 0.015] = 1
    if read_len >= 71:                                   # (the draws of the standard read lengths must not change: seeded workloads)
        k_hi, ins_hi, a_lo, a_hi = 51, 11, 10, read_len - 20
    else:                                                # short reads: clips and indels scaled so that every op keeps a positive length
        if read_len < 12:
            raise ValueError("synth_short needs read_len >= 12")
        k_hi, ins_hi = read_len // 3 + 1, read_len // 6 + 1
        a_lo = read_len // 4
        a_hi = max(a_lo + 1, read_len - read_len // 4 - ins_hi)
    k = rng.integers(1, k_hi, size=n)                    # clip length
    ins = rng.integers(1, ins_hi, size=n)
    dele = rng.integers(1, 31, size=n)
    a = rng.integers(a_lo, a_hi, size=n)                 # left M length for indel reads
    nops = np.ones(n, dtype=np.int64)
    nops[(kind == 1) | (kind == 2)] = 2
    nops[(kind == 3) | (kind == 4)] = 3
    # deletions must not run past the contig end
    too_far = (kind == 4) & (pos + read_len + dele > length)
    kind[too_far] = 0; nops[too_far] = 1
    cigar_off = np.concatenate([[0], np.cumsum(nops)])
    cigar = np.zeros(int(cigar_off[-1]), dtype=np.uint32)
    o0 = cigar_off[:-1]
    M, I, D, S = 0, 1, 2, 4
    enc = lambda ln, op: (ln.astype(np.uint32) << 4) | np.uint32(op)
    m0 = kind == 0
    cigar[o0[m0]] = (read_len << 4) | M
    m1 = kind == 1
    cigar[o0[m1]] = enc(k[m1], S); cigar[o0[m1] + 1] = enc(read_len - k[m1], M)
    m2 = kind == 2
    cigar[o0[m2]] = enc(read_len - k[m2], M); cigar[o0[m2] + 1] = enc(k[m2], S)
    m3 = kind == 3
    cigar[o0[m3]] = enc(a[m3], M); cigar[o0[m3] + 1] = enc(ins[m3], I); cigar[o0[m3] + 2] = enc(read_len - a[m3] - ins[m3], M)
    m4 = kind == 4
    cigar[o0[m4]] = enc(a[m4], M); cigar[o0[m4] + 1] = enc(dele[m4], D); cigar[o0[m4] + 2] = enc(read_len - a[m4], M)

    qlen = np.full(n, read_len, dtype=np.int64)
    no_seq = (flag & 0x100).astype(bool) & (rng.random(n) < 0.5)      # secondary alignments often carry SEQ '*'
    qlen[no_seq] = 0
    qual_off = np.concatenate([[0], np.cumsum(qlen)]).astype(np.uint64)
    qual = _fill_quals(int(qual_off[-1]), rng, qual_device)
    assert int((cigar >> 4).max(initial=1)) <= read_len + 30 and int((cigar >> 4).min(initial=1)) >= 1, "generator produced an empty or wrapped CIGAR op"
    reads = ReadColumns(pos.astype(np.int32), flag, mapq, cigar_off.astype(np.uint32), cigar, qual_off, qual, name_id)
    return SynthContig(name, length, ref, reads, blocks)


def synth_long(name: str, length: int, seed: int, depth: float = 30.0, mean_len: int = 15_000, sd_len: int = 3_000,
               qual_device=None) -> SynthContig:
    rng = np.random.default_rng(seed)
    ref, blocks = make_reference(length, rng)
    mean_len = min(mean_len, max(200, length // 8)); sd_len = min(sd_len, mean_len // 4)
    n = max(1, int(depth * length / mean_len))
    target = np.clip(rng.normal(mean_len, sd_len, size=n), 200, 4 * mean_len).astype(np.int64)
    nseg = np.maximum(1, target // 35)                        # M runs of 20..50 (mean 35) separated by an indel
    S = int(nseg.sum())
    seg_off = np.concatenate([[0], np.cumsum(nseg)])
    mlen = rng.integers(20, 51, size=S)
    is_ins = rng.random(S) < 0.5
    ilen = rng.integers(1, 4, size=S)
    last = np.zeros(S, dtype=bool); last[seg_off[1:] - 1] = True
    ref_contrib = mlen + np.where(~is_ins & ~last, ilen, 0)
    qry_contrib = mlen + np.where(is_ins & ~last, ilen, 0)
    cs_ref = np.concatenate([[0], np.cumsum(ref_contrib)]); cs_qry = np.concatenate([[0], np.cumsum(qry_contrib)])
    span = cs_ref[seg_off[1:]] - cs_ref[seg_off[:-1]]
    qlen = cs_qry[seg_off[1:]] - cs_qry[seg_off[:-1]]
    clipL = np.where(rng.random(n) < 0.05, rng.integers(10, 500, size=n), 0)
    clipR = np.where(rng.random(n) < 0.05, rng.integers(10, 500, size=n), 0)
    pos = rng.integers(0, max(1, length - int(span.max()) - 1), size=n, dtype=np.int64)
    order = np.argsort(pos, kind="stable")
    # ops per read = [S?] + (2*nseg-1) + [S?]
    nops = 2 * nseg - 1 + (clipL > 0) + (clipR > 0)
    # build in ORIGINAL read order, then permute records with ReadColumns.select
    cigar_off = np.concatenate([[0], np.cumsum(nops)])
    cigar = np.zeros(int(cigar_off[-1]), dtype=np.uint32)
    rid = np.repeat(np.arange(n), nseg)
    j = np.arange(S) - seg_off[rid]                            # segment index within its read
    base = cigar_off[rid] + (clipL[rid] > 0) + 2 * j
    cigar[base] = (mlen.astype(np.uint32) << 4) | 0
    nl = ~last
    cigar[base[nl] + 1] = (ilen[nl].astype(np.uint32) << 4) | np.where(is_ins[nl], 1, 2).astype(np.uint32)
    hasL = clipL > 0
    cigar[cigar_off[:-1][hasL]] = (clipL[hasL].astype(np.uint32) << 4) | 4
    hasR = clipR > 0
    cigar[cigar_off[1:][hasR] - 1] = (clipR[hasR].astype(np.uint32) << 4) | 4
    qtot = qlen + clipL + clipR
    qual_off = np.concatenate([[0], np.cumsum(qtot)]).astype(np.uint64)
    qual = _fill_quals(int(qual_off[-1]), rng, qual_device)
    zones = _zones(length, 0.01, 2000, 20000, rng)
    mapq, flag = _mapq_flags(n, pos, zones, rng)
    flag &= np.uint16(~0x400 & 0xFFFF)
    reads = ReadColumns(pos.astype(np.int32), flag, mapq, cigar_off.astype(np.uint32), cigar, qual_off, qual,
                        np.arange(n, dtype=np.uint32)).select(order)
    return SynthContig(name, length, ref, reads, blocks)


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs (sizes follow hg38; `scale` shrinks lengths for tests)
# ---------------------------------------------------------------------------------------------
HG38 = {"chr1": 248_956_422, "chr2": 242_193_529, "chr3": 198_295_559, "chr4": 190_214_555, "chr5": 181_538_259,
        "chr6": 170_805_979, "chr7": 159_345_973, "chr8": 145_138_636, "chr9": 138_394_717, "chr10": 133_797_422,
        "chr11": 135_086_622, "chr12": 133_275_309, "chr13": 114_364_328, "chr14": 107_043_718, "chr15": 101_991_189,
        "chr16": 90_338_345, "chr17": 83_257_441, "chr18": 80_373_285, "chr19": 58_617_616, "chr20": 64_444_167,
        "chr21": 46_709_983, "chr22": 50_818_468, "chrX": 156_040_895, "chrY": 57_227_415, "chrM": 16_569}
SEED0 = 20261018


def config(config_id: int, scale: float = 1.0, qual_device=None) -> List[SynthContig]:
    """The five BASELINE.json workloads (0-based ids follow `configs`)."""
    L = lambda nm: max(64, int(HG38[nm] * scale)) if nm != "chrM" else HG38[nm]
    seed = SEED0 + config_id
    if config_id == 0:
        return [synth_short("chr22", L("chr22"), seed, qual_device=qual_device)]
    if config_id == 1:
        return [synth_short("chr1", L("chr1"), seed, qual_device=qual_device)]
    if config_id == 2:
        return [synth_short(nm, L(nm), seed + 100 * i, qual_device=qual_device) for i, nm in enumerate(HG38)]
    if config_id == 3:
        return [synth_short("chrY", L("chrY"), seed, depth=2000.0, qual_device=qual_device),
                synth_short("chrM", HG38["chrM"], seed + 1, depth=2000.0, qual_device=qual_device)]
    if config_id == 4:
        return [synth_long("chr1", L("chr1"), seed, qual_device=qual_device)]
    raise ValueError(config_id)
