"""ctypes front-end of the CPU oracle (oracle/callable_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by decodingustools_b200/.
PARITY UNPINNED (see the C file header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "callable_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c11", "-shared", "-o", _LIB_PATH, src])
    return _LIB_PATH


class _Options(C.Structure):
    _fields_ = [("min_depth", C.c_uint32), ("max_depth", C.c_uint32), ("min_depth_for_low_mapq", C.c_uint32),
                ("min_mapping_quality", C.c_uint8), ("min_base_quality", C.c_uint8), ("max_low_mapq", C.c_uint8),
                ("_pad", C.c_uint8), ("max_low_mapq_fraction", C.c_double)]


class _Reads(C.Structure):
    _fields_ = [("n", C.c_uint64), ("pos", C.c_void_p), ("flag", C.c_void_p), ("mapq", C.c_void_p),
                ("cigar_off", C.c_void_p), ("cigar", C.c_void_p), ("qual_off", C.c_void_p), ("qual", C.c_void_p),
                ("name_id", C.c_void_p), ("n_names", C.c_uint32)]


class _ContigResult(C.Structure):
    _fields_ = [("counts", C.c_uint64 * 6), ("n_covered_bases", C.c_uint64), ("summed_coverage", C.c_uint64),
                ("summed_baseq", C.c_uint64), ("summed_mapq", C.c_uint64), ("quality_bases", C.c_uint64),
                ("n_reads", C.c_uint64), ("n_admitted", C.c_uint64), ("length", C.c_uint32),
                ("has_bins", C.c_uint32), ("n_bins", C.c_uint32), ("stride", C.c_uint32)]


class _ContigFloats(C.Structure):
    _fields_ = [("coverage_percent", C.c_double), ("average_depth", C.c_double), ("average_mapq", C.c_double),
                ("average_baseq", C.c_double), ("q30_percentage", C.c_double)]


class _Summary(C.Structure):
    _fields_ = [("total_bases", C.c_uint64), ("callable_bases", C.c_uint64), ("total_unique_reads", C.c_uint64),
                ("contigs_analyzed", C.c_uint64), ("callable_percentage", C.c_double), ("average_depth", C.c_double),
                ("average_mapq", C.c_double), ("average_baseq", C.c_double), ("q30_percentage", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_run_new.restype = C.c_void_p
        L.orc_run_new.argtypes = [C.POINTER(_Options), C.c_uint32]
        L.orc_run_free.argtypes = [C.c_void_p]
        L.orc_bed.restype = C.c_void_p
        L.orc_bed.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_admit.argtypes = [C.POINTER(_Reads), C.c_int32, C.c_uint32, C.c_void_p]
        L.orc_process_contig.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.c_uint32, C.c_void_p, C.c_uint64,
                                         C.POINTER(_Reads), C.POINTER(_ContigResult),
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_get_bins.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_build_export.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_ContigFloats), C.POINTER(_Summary)]
        L.orc_n_contigs.restype = C.c_uint32
        L.orc_n_contigs.argtypes = [C.c_void_p]
        L.orc_compare_contig_names.argtypes = [C.c_char_p, C.c_char_p]
        _lib = L
    return _lib


def _opts(o) -> _Options:
    return _Options(o.min_depth, o.max_depth, o.min_depth_for_low_mapq, o.min_mapping_quality,
                    o.min_base_quality, o.max_low_mapq, 0, float(o.max_low_mapq_fraction))


def _reads(rc) -> _Reads:
    has_names = rc.name_id is not None and rc.n > 0
    n_names = int(rc.name_id.max()) + 1 if has_names else 0
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    return _Reads(rc.n, p(rc.pos), p(rc.flag), p(rc.mapq), p(rc.cigar_off), p(rc.cigar), p(rc.qual_off),
                  p(rc.qual), p(rc.name_id) if has_names else None, n_names)


@dataclass
class OracleContig:
    name: str
    length: int
    counts: List[int]
    n_covered_bases: int
    summed_coverage: int
    summed_baseq: int
    summed_mapq: int
    quality_bases: int
    n_reads: int
    n_admitted: int
    bins: Optional[np.ndarray] = None     # uint32[3, n_bins]: callable, poor_mapq, ref_n; None if not produced
    stride: int = 0
    raw: Optional[np.ndarray] = None
    qc: Optional[np.ndarray] = None
    low: Optional[np.ndarray] = None
    state: Optional[np.ndarray] = None


def admit(rc, maxcnt: int, tid: int = 0) -> np.ndarray:
    """keep[r] for htslib's bam_plp_push admission."""
    keep = np.zeros(rc.n, dtype=np.uint8)
    rd = _reads(rc)
    if lib().orc_admit(C.byref(rd), int(tid), int(maxcnt), keep.ctypes.data_as(C.c_void_p)) != 0:
        raise ValueError("records are not coordinate sorted")
    return keep.astype(bool)


class OracleRun:
    """One `coverage` run: a CallableProfiler that lives across contigs (quirks Q1/Q2 included)."""

    def __init__(self, options, largest_contig_length: int):
        self._o = _opts(options)
        self._h = lib().orc_run_new(C.byref(self._o), int(largest_contig_length))
        self.contigs: List[OracleContig] = []

    def __del__(self):
        try:
            if self._h:
                lib().orc_run_free(self._h)
                self._h = None
        except Exception:
            pass

    def process_contig(self, name: str, tid: int, length: int, ref_ascii, reads, debug: bool = False) -> OracleContig:
        ref = None
        ref_len = 0
        if ref_ascii is not None:
            ref = np.frombuffer(ref_ascii, dtype=np.uint8) if isinstance(ref_ascii, (bytes, bytearray)) else np.ascontiguousarray(ref_ascii, dtype=np.uint8)
            ref_len = ref.shape[0]
        res = _ContigResult()
        rd = _reads(reads)
        dbg = [np.zeros(length, np.uint32) for _ in range(3)] + [np.zeros(length, np.uint8)] if debug else [None] * 4
        ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        rc = lib().orc_process_contig(self._h, name.encode(), int(tid), int(length), ptr(ref), ref_len, C.byref(rd),
                                      C.byref(res), ptr(dbg[0]), ptr(dbg[1]), ptr(dbg[2]), ptr(dbg[3]))
        if rc != 0:
            raise RuntimeError(lib().orc_last_error(self._h).decode())
        oc = OracleContig(name, length, list(res.counts), res.n_covered_bases, res.summed_coverage, res.summed_baseq,
                          res.summed_mapq, res.quality_bases, res.n_reads, res.n_admitted, None, res.stride,
                          dbg[0], dbg[1], dbg[2], dbg[3])
        if res.has_bins:
            b = np.zeros((3, res.n_bins), dtype=np.uint32)
            lib().orc_get_bins(self._h, len(self.contigs), ptr(b[0]), ptr(b[1]), ptr(b[2]))
            oc.bins = b
        self.contigs.append(oc)
        return oc

    def bed(self) -> bytes:
        n = C.c_uint64(0)
        p = lib().orc_bed(self._h, C.byref(n))
        return C.string_at(p, n.value)

    def export(self):
        """(order, per-contig floats in natural order, summary) == report.rs:15-134."""
        n = len(self.contigs)
        order = np.zeros(max(n, 1), dtype=np.uint32)
        floats = (_ContigFloats * max(n, 1))()
        summ = _Summary()
        lib().orc_build_export(self._h, order.ctypes.data_as(C.c_void_p), floats, C.byref(summ))
        fl = [dict(coverage_percent=f.coverage_percent, average_depth=f.average_depth, average_mapq=f.average_mapq,
                   average_baseq=f.average_baseq, q30_percentage=f.q30_percentage) for f in floats[:n]]
        sm = {k: getattr(summ, k) for k, _ in _Summary._fields_}
        return [int(i) for i in order[:n]], fl, sm


def compare_contig_names(a: str, b: str) -> int:
    return int(lib().orc_compare_contig_names(a.encode(), b.encode()))
