"""Final floating-point aggregation and contig ordering (host side).

Mirrors /root/reference/src/callable_loci/report.rs:15-134 (build_coverage_export), :337-393 (natural
contig order) and profilers/contig_profiler.rs:123-158 (get_quality_stats).  The numbers are plain IEEE
doubles accumulated in the same natural-contig order, so they are bit-identical to the oracle's.
"""
from __future__ import annotations

import functools
from typing import Dict, List

from .callable_loci import CallableProfiler, ContigProfiler
from .options import CalledState


def split_contig_name(name: str):
    for i, ch in enumerate(name):
        if ch.isascii() and (ch.isdigit() or ch in "XYM"):
            return name[:i], name[i:]
    return name, ""


def _order(s: str):
    t = s[1:] if s.startswith("+") else s          # Rust's u32::from_str accepts a leading '+'
    if t and t.isascii() and t.isdigit() and int(t) <= 0xFFFFFFFF:
        return (0, int(t))
    return ({"X": 1, "Y": 2, "M": 3, "MT": 3}.get(s, 4), 0)


def compare_contig_names(a: str, b: str) -> int:
    (ap, asuf), (bp, bsuf) = split_contig_name(a), split_contig_name(b)
    if ap != bp:
        return -1 if ap.encode() < bp.encode() else 1
    (ac, an), (bc, bn) = _order(asuf), _order(bsuf)
    if ac != bc:
        return -1 if ac < bc else 1
    if ac == 0:
        return (an > bn) - (an < bn)
    ab, bb = asuf.encode(), bsuf.encode()
    return (ab > bb) - (ab < bb)


def quality_stats(s: ContigProfiler) -> Dict[str, float]:
    average_mapq = s.summed_mapq / s.quality_bases if s.quality_bases > 0 else 0.0
    average_baseq = s.summed_baseq / s.quality_bases if s.quality_bases > 0 else 0.0
    if s.quality_bases > 0:
        if average_baseq >= 30.0:
            q30 = 100.0
        elif average_baseq < 20.0:
            q30 = 0.0
        else:
            q30 = ((average_baseq - 20.0) / 10.0) * 100.0
    else:
        q30 = 0.0
    return dict(average_mapq=average_mapq, average_baseq=average_baseq, q30_percentage=q30)


def build_coverage_export(contig_stats: Dict[int, ContigProfiler], counter: CallableProfiler, bam_stats: dict | None = None) -> dict:
    """Returns the `export` object of summary.json."""
    bam_stats = bam_stats or {}
    total_bases = callable_bases = q30_bases = total_qpos = total_unique = 0
    total_depth = total_mapq = total_baseq = 0.0
    contigs: List[dict] = []
    ordered = sorted(contig_stats.values(), key=functools.cmp_to_key(lambda a, b: compare_contig_names(a.name, b.name)))
    for s in ordered:
        counts = counter.get_contig_counts(s.name)
        q = quality_stats(s)
        coverage_percent = (s.n_covered_bases / s.length) * 100.0 if s.length > 0 else 0.0
        average_depth = s.summed_coverage / s.n_covered_bases if s.n_covered_bases > 0 else 0.0
        total_bases += s.length
        callable_bases += int(counts[CalledState.CALLABLE])
        total_depth += average_depth * float(s.length)
        total_mapq += q["average_mapq"] * float(s.length)
        total_baseq += q["average_baseq"] * float(s.length)
        q30_bases += int(q["q30_percentage"] / 100.0 * float(s.length))
        total_qpos += s.length
        total_unique += s.n_reads & 0xFFFFFFFF
        contigs.append(dict(
            name=s.name, length=s.length, unique_reads=s.n_reads, coverage_percent=coverage_percent,
            average_depth=average_depth, covered_bases=s.n_covered_bases, total_bases=s.length, quality_stats=q,
            state_distribution=dict(ref_n=int(counts[0]), callable=int(counts[1]), no_coverage=int(counts[2]),
                                    low_coverage=int(counts[3]), excessive_coverage=int(counts[4]),
                                    poor_mapping_quality=int(counts[5]))))
    summary = dict(
        aligner=bam_stats.get("aligner", "Unknown"), reference_build=bam_stats.get("reference_build", "Unknown"),
        sequencing_platform=bam_stats.get("sequencing_platform", "Unknown"), read_length=bam_stats.get("read_length", 0),
        total_bases=total_bases, callable_bases=callable_bases,
        callable_percentage=(callable_bases / total_bases) * 100.0 if total_bases > 0 else 0.0,
        average_depth=total_depth / total_bases if total_bases > 0 else 0.0, contigs_analyzed=len(contig_stats))
    qm = dict(average_mapq=total_mapq / total_qpos if total_qpos > 0 else 0.0,
              average_baseq=total_baseq / total_qpos if total_qpos > 0 else 0.0,
              q30_percentage=(q30_bases / total_qpos) * 100.0 if total_qpos > 0 else 0.0)
    return dict(summary=summary, contigs=contigs, quality_metrics=qm, total_unique_reads=total_unique)
