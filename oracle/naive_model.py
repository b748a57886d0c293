"""Deliberately naive, independently written Python model of the CallableLoci path.

TEST INFRASTRUCTURE ONLY (pure-Python loops: micro-contigs only).  It exists to cross-check
oracle/callable_oracle.c: the C oracle restates htslib's *iterator mechanics* (mempool count,
linked list, column loop); this model instead uses the *derived closed-form rules* of SURVEY.md
section 8 rows A0-A7 and Appendix A, so a disagreement flags a reading error in one of the two.
PARITY UNPINNED, like the C oracle.

Reference citations: /root/reference/src/callable_loci/mod.rs:17-147,
profilers/callable_profiler.rs:39-155, profilers/contig_profiler.rs:47-83,
utils/histogram_plotter.rs:74-102,412-441.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

STATE = ["REF_N", "CALLABLE", "NO_COVERAGE", "LOW_COVERAGE", "EXCESSIVE_COVERAGE", "POOR_MAPPING_QUALITY"]
REF_OPS = {0, 2, 3, 7, 8}
QRY_OPS = {0, 1, 4, 7, 8}


def admitted(reads, maxcnt: int, tid: int = 0) -> List[bool]:
    """drop iff not first at its start position and live >= maxcnt, where
    live = previously admitted reads with end >= pos.  A zero-span read that is not first at its
    position never becomes a live node."""
    keep = []
    live_ends: List[int] = []
    last_pos = None
    first_record = True
    for i in range(reads.n):
        pos = int(reads.pos[i])
        if int(reads.flag[i]) & 4:
            keep.append(False)
            continue
        span = 0
        for c in range(int(reads.cigar_off[i]), int(reads.cigar_off[i + 1])):
            v = int(reads.cigar[c])
            if (v & 15) in REF_OPS:
                span += v >> 4
        end = pos + span
        # the iterator starts at (tid 0, pos 0): the first record of tid 0, if at pos 0, is treated as
        # "not first at its position" (it takes the cap test, and a zero-span one is not retained).
        same = (pos == last_pos) or (first_record and pos == 0 and tid == 0)
        first_record = False
        if same:
            live_ends = [e for e in live_ends if e >= pos]
            if len(live_ends) >= maxcnt:
                keep.append(False)
                continue
            if end > pos:
                live_ends.append(end); keep.append(True)
            else:
                keep.append(False)      # zero-span, not first at its position: node is not retained
        else:
            live_ends = [e for e in live_ends if e >= pos]
            live_ends.append(end); keep.append(True)
        last_pos = pos
    return keep


def per_base(reads, keep, length: int, opt) -> Tuple[List[int], List[int], List[int], Dict[str, int]]:
    raw = [0] * length; qc = [0] * length; low = [0] * length
    sums = dict(summed_baseq=0, summed_mapq=0, quality_bases=0)
    seen = set()
    for i in range(reads.n):
        if not keep[i]:
            continue
        p = int(reads.pos[i]); q = 0
        mq = int(reads.mapq[i])
        q0 = int(reads.qual_off[i]); lq = int(reads.qual_off[i + 1]) - q0
        touched = False
        for c in range(int(reads.cigar_off[i]), int(reads.cigar_off[i + 1])):
            v = int(reads.cigar[c]); op = v & 15; ln = v >> 4
            if op in REF_OPS:
                for j in range(ln):
                    x = p + j
                    if 0 <= x < length:
                        touched = True
                        raw[x] += 1
                        if mq <= opt.max_low_mapq:
                            low[x] += 1
                        if mq >= opt.min_mapping_quality:
                            sums["summed_mapq"] += mq
                            if op in (0, 7, 8) and q + j < lq:
                                b = int(reads.qual[q0 + q + j])
                                if b >= opt.min_base_quality:
                                    qc[x] += 1
                                    sums["summed_baseq"] += b
                                    sums["quality_bases"] += 1
                p += ln
            if op in QRY_OPS:
                q += ln
        if touched and reads.name_id is not None:
            seen.add(int(reads.name_id[i]))
    sums["n_reads"] = len(seen)
    return raw, qc, low, sums


def classify(opt, ref_base: int, raw: int, qc: int, low: int) -> int:
    is_low = raw >= opt.min_depth_for_low_mapq and (float(low) / float(raw)) > opt.max_low_mapq_fraction if raw > 0 else False
    if ref_base in (ord("N"), ord("n")):
        return 0
    if raw == 0:
        return 2
    if is_low:
        return 5
    if qc < opt.min_depth:
        return 3
    if opt.max_depth > 0 and qc > opt.max_depth:
        return 4
    return 1


class NaiveRun:
    """BED text + counters for a sequence of contigs, with quirks Q1 (duplicated boundary line) and
    Q2 (stale range binned into the next contig) produced from run lists rather than a state machine."""

    def __init__(self, opt, largest_contig_length: int):
        self.opt = opt
        self.largest = largest_contig_length
        self.bed_lines: List[str] = []
        self.pending: Optional[Tuple[str, int, int, int]] = None   # last run of the last non-empty contig
        self.results = []

    def process_contig(self, name: str, length: int, ref: bytes, reads, tid: int = 0):
        opt = self.opt
        keep = admitted(reads, opt.max_depth if opt.max_depth > 0 else 500, tid)
        raw, qc, low, sums = per_base(reads, keep, length, opt)
        states = [classify(opt, ref[p] if p < len(ref) else ord("N"), raw[p], qc[p], low[p]) for p in range(length)]
        runs = []
        for p, s in enumerate(states):
            if runs and runs[-1][2] == s:
                runs[-1][1] = p + 1
            else:
                runs.append([p, p + 1, s])
        ranges = []      # coverage_ranges of this contig at finish time
        fmt = lambda r: f"{r[0]}\t{r[1]}\t{r[2]}\t{STATE[r[3]]}\n"
        if length > 0:
            if self.pending is not None:
                self.bed_lines.append(fmt(self.pending))            # Q1
                if self.pending[3] in (1, 5, 0):
                    ranges.append(self.pending[1:])                 # Q2
            for a, b, s in runs:
                self.bed_lines.append(fmt((name, a, b, s)))
                if s in (1, 5, 0):
                    ranges.append((a, b, s))
            self.pending = (name, runs[-1][0], runs[-1][1], runs[-1][2])
        else:
            if self.pending is not None:                            # finish_contig re-writes the stale run
                self.bed_lines.append(fmt(self.pending))
                if self.pending[3] in (1, 5, 0):
                    ranges.append(self.pending[1:])
        bins = None
        stride = 0
        if ranges:
            stride = (16569 + 199) // 200 if name == "chrM" else (self.largest + 1999) // 2000
            nb = length // stride + 1
            bins = [[0] * nb for _ in range(3)]
            for a, b, s in ranges:
                k = {1: 0, 5: 1, 0: 2}[s]
                for p in range(a, b):
                    if p // stride < nb:
                        bins[k][p // stride] += 1
        counts = [0] * 6
        for s in states:
            counts[s] += 1
        res = dict(name=name, length=length, counts=counts, raw=raw, qc=qc, low=low, states=states,
                   n_covered_bases=sum(1 for r in raw if r > 0), summed_coverage=sum(raw),
                   bins=bins, stride=stride, keep=keep, **sums)
        self.results.append(res)
        return res

    def bed(self) -> bytes:
        return "".join(self.bed_lines).encode()
