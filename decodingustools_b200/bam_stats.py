"""BAM header / read-name sampling that fills the report header.

Host-side mirror of /root/reference/src/callable_loci/profilers/bam_stats.rs:44-141 (the sampler: first
`max_samples` records of the file, primary alignments only) and profilers/platform_inference.rs:16-294 (platform
and instrument model from the read-name format).  Pure host logic, no device work.

Where the reference takes `max_by_key` over a HashMap (primary platform, most common instrument) a tie is
resolved by hash order there, i.e. it is not defined; here ties go to the key that was seen first.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

ILLUMINA, PACBIO, NANOPORE, MGI, UNKNOWN = "Illumina", "PacBio", "Nanopore", "MGI", "Unknown"

_HEX = set("0123456789abcdefABCDEF")


def detect_platform_from_qname(qname: str) -> str:
    """platform_inference.rs:16-94."""
    if len(qname) > 30 and ("-" in qname or "_" in qname):
        parts = qname.split("-")
        if len(parts) == 5:
            is_uuid = len(parts[0]) == 8 and len(parts[1]) == 4 and len(parts[2]) == 4 and len(parts[3]) == 4 and len(parts[4]) >= 12
            if is_uuid and all(all(c in _HEX for c in p) for p in parts):
                return NANOPORE
        if "ch" in qname and "read" in qname:
            return NANOPORE
    if qname.startswith("m") and "/" in qname:
        parts = qname.split("/")
        if len(parts) >= 2 and "_" in parts[0]:
            return PACBIO
    if len(qname) > 15:
        prefix = qname[:5].upper()
        if prefix.startswith(("V300", "E100", "CL100", "G400", "G99")):
            return MGI
        if qname.count(":") >= 6:
            parts = qname.split(":")
            if parts[0].startswith(("V", "E", "CL", "G")) and len(parts) >= 3 and parts[2].startswith("L"):
                return MGI
    if qname.count(":") >= 6:
        return ILLUMINA
    return UNKNOWN


def parse_illumina_read_name(qname: str) -> Optional[Tuple[str, str]]:
    """(instrument, flow cell): platform_inference.rs:99-108."""
    parts = qname.split(":")
    return (parts[0], parts[2]) if len(parts) >= 3 else None


def parse_pacbio_read_name(qname: str) -> Optional[str]:
    """platform_inference.rs:114-126."""
    slash = qname.find("/")
    if slash >= 0:
        movie = qname[:slash]
        if movie.startswith("m"):
            us = movie.find("_")
            if us >= 0:
                return movie[:us]
    return None


def parse_nanopore_read_name(qname: str) -> Optional[str]:
    """platform_inference.rs:132-159."""
    if len(qname) > 30 and "-" in qname and len(qname.split("-")) >= 5:
        return qname.split("_")[0].split("-")[0]
    us = qname.find("_")
    if us >= 0:
        return qname[:us]
    return "nanopore"


def parse_mgi_read_name(qname: str) -> Optional[Tuple[str, str]]:
    """platform_inference.rs:165-191."""
    if qname.count(":") >= 3:
        parts = qname.split(":")
        return parts[0], parts[1]
    if len(qname) > 10:
        l_pos = qname.find("L")
        if l_pos >= 0:
            rest = qname[l_pos:]
            if rest.find("C") >= 0:
                r_pos = rest.find("R")
                end = r_pos if r_pos >= 0 else len(rest)
                return qname[:l_pos], rest[:end]
    return None


def _most_common(counts: Dict[str, int]) -> Optional[str]:
    best = None
    for k, v in counts.items():                     # insertion order: the first key seen wins a tie
        if best is None or v > counts[best]:
            best = k
    return best


def infer_specific_platform(primary: str, instruments: Dict[str, int]) -> str:
    """platform_inference.rs:217-293."""
    top = _most_common(instruments)
    if primary == PACBIO:
        if top is None:
            return "PacBio"
        return ("PacBio Revio" if top.startswith("m84") else "PacBio Sequel II/IIe" if top.startswith("m64")
                else "PacBio Sequel" if top.startswith("m54") else "PacBio")
    if primary == NANOPORE:
        return "Oxford Nanopore"
    if primary == MGI:
        if top is None:
            return "MGI DNBseq"
        return ("MGI DNBSEQ/MGISEQ-2000" if top.startswith("V300") else "MGI MGISEQ-200" if top.startswith("E100")
                else "MGI MGISEQ-T7" if top.startswith("CL100") else "MGI DNBSEQ-G400" if top.startswith("G400")
                else "MGI MGISEQ-T1" if top.startswith("G99") else "MGI DNBseq")
    if primary == ILLUMINA:
        if top is None:
            return "Unknown Illumina"
        return {"a": "NovaSeq", "d": "HiSeq 2500", "j": "HiSeq 3000", "k": "HiSeq 4000", "e": "HiSeq X", "n": "NextSeq",
                "m": "MiSeq", "v": "NovaSeq X", "f": "iSeq"}.get(top[:1].lower() if top[:1].isascii() else "", "Unknown Illumina")
    return "Unknown"


def detect_aligner(header_text: str) -> str:
    """callable_loci/mod.rs:149-177."""
    h = header_text.lower()
    for key, name in (("@pg\tid:bwa-mem2", "BWA-MEM2"), ("@pg\tid:bwa", "BWA"), ("@pg\tid:minimap2", "minimap2"),
                      ("@pg\tid:pbmm2", "pbmm2"), ("@pg\tid:bowtie2", "Bowtie2"), ("@pg\tid:star", "STAR"),
                      ("bwa", "BWA"), ("minimap2", "minimap2"), ("bowtie2", "Bowtie2"), ("star", "STAR")):
        if key in h:
            return name
    return "Unknown"


def reference_build(header_text: str) -> str:
    """types.rs:100-147 (ReferenceGenome::from_header)."""
    t = header_text
    if "AS:GRCh38" in t or "GCA_000001405.15" in t:
        return "GRCh38"
    if "AS:GRCh37" in t or "GCA_000001405.1" in t:
        return "GRCh37"
    if any(k in t for k in ("AS:CHM13", "GCA_009914755.4", "chm13", "CHM13", "t2t", "T2T")):
        return "T2T-CHM13v2.0"
    if "SN:chr1" in t and "LN:248387328" in t and "M5:e469247288ceb332aee524caec92bb22" in t:
        return "T2T-CHM13v2.0"
    if "SN:chr1" in t and "LN:248956422" in t:
        return "GRCh38"
    if "SN:1" in t and "LN:249250621" in t:
        return "GRCh37"
    return "Unknown"


class BamStats:
    """bam_stats.rs:9-141: feed it the records of the file in order; it looks at the first `max_samples`."""

    def __init__(self, max_samples: int = 10000):
        self.max_samples = max_samples
        self.seen = 0
        self.read_count = 0
        self.total_read_length = 0
        self.paired_reads = 0
        self.length_distribution: Dict[int, int] = {}
        self.instruments: Dict[str, int] = {}
        self.flow_cells: Dict[str, int] = {}
        self.platform_counts: Dict[str, int] = {}
        self.aligner = "Unknown"
        self.reference_build = "Unknown"

    def set_header(self, header_text: str):
        self.aligner = detect_aligner(header_text)
        self.reference_build = reference_build(header_text)

    def add_record(self, qname: str, flag: int, seq_len: int) -> bool:
        """Returns False once the sample is full (the caller may stop feeding)."""
        if self.seen >= self.max_samples:
            return False
        self.seen += 1
        if flag & 0x900:                                     # secondary / supplementary
            return True
        self.length_distribution[seq_len] = self.length_distribution.get(seq_len, 0) + 1
        self.read_count += 1
        self.total_read_length += seq_len
        platform = detect_platform_from_qname(qname)
        self.platform_counts[platform] = self.platform_counts.get(platform, 0) + 1
        inst = cell = None
        if platform == ILLUMINA:
            r = parse_illumina_read_name(qname)
            if r:
                inst, cell = r
        elif platform == PACBIO:
            inst = parse_pacbio_read_name(qname)
        elif platform == NANOPORE:
            inst = parse_nanopore_read_name(qname)
        elif platform == MGI:
            r = parse_mgi_read_name(qname)
            if r:
                inst, cell = r
        if inst is not None:
            self.instruments[inst] = self.instruments.get(inst, 0) + 1
        if cell is not None:
            self.flow_cells[cell] = self.flow_cells.get(cell, 0) + 1
        if flag & 0x1:
            self.paired_reads += 1
        return True

    def collect(self, records: Iterable[Tuple[str, int, int]]):
        for qname, flag, seq_len in records:
            if not self.add_record(qname, flag, seq_len):
                break
        return self

    def average_read_length(self) -> int:
        return self.total_read_length // self.read_count if self.read_count else 0

    def primary_platform(self) -> str:
        return _most_common(self.platform_counts) or UNKNOWN

    def infer_platform(self) -> str:
        return infer_specific_platform(self.primary_platform(), self.instruments)

    def as_summary_fields(self) -> dict:
        return dict(aligner=self.aligner, reference_build=self.reference_build, sequencing_platform=self.infer_platform(),
                    read_length=self.average_read_length())
