"""CallableOptions / CalledState — host-side mirror of the reference's option struct and enum.

Reference: /root/reference/src/callable_loci/options.rs:2-38 (fields and constructor argument
order), /root/reference/src/cli.rs:34-60 (defaults), /root/reference/src/callable_loci/types.rs:36-43
(state discriminants; BED text uses the enum identifiers, callable_profiler.rs:42-46).
"""
from __future__ import annotations

import enum
from dataclasses import dataclass
from typing import List, Optional


class CalledState(enum.IntEnum):
    REF_N = 0
    CALLABLE = 1
    NO_COVERAGE = 2
    LOW_COVERAGE = 3
    EXCESSIVE_COVERAGE = 4
    POOR_MAPPING_QUALITY = 5


STATE_NAMES = [s.name for s in CalledState]


@dataclass
class CallableOptions:
    min_depth: int = 4
    max_depth: int = 500
    min_mapping_quality: int = 10
    min_base_quality: int = 20
    min_depth_for_low_mapq: int = 10
    max_low_mapq: int = 1
    max_low_mapq_fraction: float = 0.1
    selected_contigs: Optional[List[str]] = None

    def __post_init__(self):
        for name, hi in (("min_depth", 2**32), ("max_depth", 2**32), ("min_depth_for_low_mapq", 2**32),
                         ("min_mapping_quality", 256), ("min_base_quality", 256), ("max_low_mapq", 256)):
            v = getattr(self, name)
            if not (0 <= int(v) < hi):
                raise ValueError(f"{name}={v} out of range for the reference's field type")

    def with_contigs(self, contigs: Optional[List[str]]) -> "CallableOptions":
        self.selected_contigs = contigs
        return self

    @property
    def pileup_max_depth(self) -> int:
        """htslib maxcnt the reference configures: mod.rs:56-60."""
        return self.max_depth if self.max_depth > 0 else 500
