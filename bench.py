#!/usr/bin/env python
"""bench.py -- aligned Gbases/s of the CallableLoci pileup + classify + BED-segmentation path on B200.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` (torchrun for N > 1) prints ONE JSON line.
A "step" is one pass of the hot path over one contig-sized batch of synthetic reads:
    N = 1 : BASELINE.json configs[1] -- chr1-size contig (248.96 Mbp), synthetic 30x 2x150 bp, one B200.
    N > 1 : every rank owns one chr1-size region shard of an N x chr1 genome (weak scaling); the only
            collective is the all-reduce of the additive counters/bins (NCCL via torch.distributed).
`value`  : pileup cells (aligned bases = sum of raw depth) per second, inputs resident in HBM.
`e2e`    : same metric through the public C-ABI call sequence with HOST (pinned) column buffers: reference upload,
           column-batch H2D copies, kernels, D2H of intervals + counters, all inside the timed region.
`--impl reference`: the CPU oracle (restatement of the reference's single-threaded loop; the Rust reference itself
           cannot be built in this image) timed on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "aligned Gbases/s CallableLoci pileup"
UNIT = "Gbases/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the contig (1.0 = chr1, 248.96 Mbp)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample-mbp", type=float, default=48.0, help="oracle sample for cpu_baseline (Mbp of the contig; about 12 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch-reads", type=int, default=4_000_000, help="column-batch size of the e2e leg")
    return ap.parse_args()


def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML, the library behind nvidia-smi's
    clocks.sm / clocks_event_reasons.* query of B200_PROFILING.md; falls back to nvidia-smi itself)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int, period_s: float = 0.004):
        import threading
        self.idx, self.period = gpu_index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        import threading
        if self._h is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set(); self._thr.join(timeout=2)
            return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(self.reasons), "samples": len(self.samples), "source": "nvml"}
        try:   # one-shot fallback
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]),
                    "reasons": [n for n, v in zip(names, out[2:6]) if v.strip().lower().startswith("active")], "samples": 1,
                    "source": "nvidia-smi after the timed region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}


def make_workload(scale: float, rank: int, device):
    """chr1-size synthetic contig for this rank (seed differs per rank), admission-filtered and packed."""
    from decodingustools_b200 import synth
    from decodingustools_b200.callable_loci import admit_reads, compact_reads
    from decodingustools_b200.options import CallableOptions
    opt = CallableOptions()
    length = max(100_000, int(synth.HG38["chr1"] * scale))
    t0 = time.time()
    c = synth.synth_short("chr1", length, synth.SEED0 + 1 + 1000 * rank, qual_device=device)
    t_gen = time.time() - t0
    t0 = time.time()
    keep = admit_reads(c.reads, opt.pileup_max_depth, 0)
    cells = int(c.reads.ref_len()[keep].sum())
    if np.array_equal(keep, (c.reads.flag & 4) == 0):
        reads = c.reads          # only placed-unmapped records were refused: the kernel skips FLAG 0x4 itself, no repacking needed
    else:
        reads = compact_reads(c.reads, keep)
    t_admit = time.time() - t0
    return opt, c, reads, cells, {"synth_s": round(t_gen, 2), "host_admission_s": round(t_admit, 2)}


def oracle_sample(c, reads_unfiltered, opt, sample_bp: int):
    """Bounded CPU sample: the first sample_bp bases of the contig with every read that starts inside."""
    from oracle import oracle
    sample_bp = min(sample_bp, c.length)
    hi = int(np.searchsorted(reads_unfiltered.pos, sample_bp - 200, side="left"))
    sub = reads_unfiltered.slice(0, hi)
    t0 = time.perf_counter()
    run = oracle.OracleRun(opt, sample_bp)
    oc = run.process_contig(c.name, 0, sample_bp, c.ref[:sample_bp], sub)
    dt = time.perf_counter() - t0
    return oc, run, dt, sub


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of the reference loop (single-threaded, like the reference)."""
    if rank != 0:
        return
    from decodingustools_b200 import synth
    from decodingustools_b200.options import CallableOptions
    opt = CallableOptions()
    sample_bp = int(args.cpu_sample_mbp * 1e6 * min(1.0, args.scale * 4))
    length = max(sample_bp, 100_000)
    c = synth.synth_short("chr1", length, synth.SEED0 + 1)
    times, cells = [], 0
    for i in range(args.warmup + args.steps):
        oc, _, dt, _ = oracle_sample(c, c.reads, opt, sample_bp)
        if i >= args.warmup:
            times.append(dt)
        cells = oc.summed_coverage
    t = sum(times)
    val = cells * len(times) / t / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/u32 integer", "data": "synthetic",
        "config": {"workload": "chr1-size synthetic 30x 2x150bp PE (BASELINE configs[1]); CPU arm runs a bounded sample",
                   "sample_bp": sample_bp, "sample_cells": int(cells)},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"first {sample_bp} bp of the contig ({int(cells)} cells) per step; oracle/callable_oracle.c, "
                                   "single thread like the reference (src/api/coverage.rs:232-234); omits the reference's per-base "
                                   "faidx call and per-cell qname hashing, so it is faster than the real binary"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from decodingustools_b200.callable_loci import CallableLociContext

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    opt, c, reads, cells_expected, prep = make_workload(args.scale, rank, dev)
    alg_bytes = reads.nbytes_device() + c.length // 8          # packed columns + 1 bit per reference base
    span = reads.max_ref_span()

    # page-lock the column buffers in place (the e2e leg streams from them; no second host copy)
    rt = torch.cuda.cudart()
    class _Pinned:
        def __init__(self, arr):
            self.arr = arr
            rc = rt.cudaHostRegister(arr.ctypes.data, arr.nbytes, 0)
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister failed: {rc}")
        def data_ptr(self):
            return self.arr.ctypes.data
    cols = {k: _Pinned(getattr(reads, k)) for k in ("pos", "flag", "mapq", "cigar", "qual")}
    ref_pinned = _Pinned(c.ref)

    ctx = CallableLociContext(opt, device=local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    cig_off = reads.cigar_off.astype(np.int64); q_off = reads.qual_off.astype(np.int64)
    # column batches as a decoder thread would emit them: batch-relative offset columns, page-locked
    batches = []
    for lo in range(0, reads.n, args.batch_reads):
        hi = min(reads.n, lo + args.batch_reads)
        co = _Pinned(np.ascontiguousarray((cig_off[lo:hi + 1] - cig_off[lo]).astype(np.uint32)))
        qo = _Pinned(np.ascontiguousarray((q_off[lo:hi + 1] - q_off[lo]).astype(np.uint64)))
        batches.append((lo, hi, co, qo))

    def e2e_pass():
        """Public call sequence with host buffers: begin (reference upload) -> column batches -> finish (D2H)."""
        ctx._check(ctx._L.clb_begin_contig(ctx._h, 0, b"chr1", c.length, ref_pinned.data_ptr(), c.length, 0, c.length, 0, c.length, span))
        ctx.reserve(reads.n, reads.n_cigar, reads.n_qual)
        for lo, hi, co, qo in batches:
            ctx.push_raw(hi - lo, int(cig_off[hi] - cig_off[lo]), int(q_off[hi] - q_off[lo]),
                         cols["pos"].data_ptr() + 4 * lo, cols["flag"].data_ptr() + 2 * lo, cols["mapq"].data_ptr() + lo,
                         co.data_ptr(), cols["cigar"].data_ptr() + 4 * int(cig_off[lo]), qo.data_ptr(),
                         cols["qual"].data_ptr() + int(q_off[lo]))
        return ctx.finish_contig(copy_intervals=False)

    keepalive: list = []
    first = e2e_pass()                      # also leaves the contig resident for the HBM-resident leg
    # size-independent properties of the full-size result (the oracle only sees a bounded sample, below)
    assert first.summed_coverage == cells_expected, (first.summed_coverage, cells_expected)
    assert int(first.state_counts.sum()) == c.length
    _, chk = ctx.rerun_resident(fetch=True, copy_intervals=True)
    iv = chk.intervals
    assert iv["start"][0] == 0 and iv["end"][-1] == c.length and np.array_equal(iv["start"][1:], iv["end"][:-1])
    assert np.all(iv["state"][1:] != iv["state"][:-1])
    assert np.array_equal(np.bincount(iv["state"], weights=(iv["end"] - iv["start"]).astype(np.float64), minlength=6).astype(np.int64),
                          chk.state_counts.astype(np.int64))
    assert int(chk.bins.sum()) == int(chk.state_counts[0] + chk.state_counts[1] + chk.state_counts[5])
    n_intervals = int(iv.shape[0]); del iv, chk

    def allreduce_counters():
        if world == 1:
            return
        ptr, n = ctx.counters_device()
        class _Wrap:     # torch tensor aliasing the library's counter buffer (uint64 sums == int64 sums bitwise)
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}
        t = torch.as_tensor(_Wrap(), device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    # ---------------- HBM-resident leg: W warm-up + K timed steps, barrier + sync on both sides
    for _ in range(args.warmup):
        ctx.rerun_resident(fetch=False)
        allreduce_counters()
    sampler = ClockSampler(local_rank); sampler.start()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    step_ms = []
    for _ in range(args.steps):
        ms, _ = ctx.rerun_resident(fetch=False)
        allreduce_counters()
        step_ms.append(ms)
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    _, res = ctx.rerun_resident(fetch=True)
    pileup_ms = res.pileup_ms
    launches_per_step = res.gpu_launches

    # ---------------- end-to-end leg through the public API with host buffers
    e2e_ms = []
    r = first
    for i in range(args.e2e_steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        r = e2e_pass()
        allreduce_counters()
        torch.cuda.synchronize()
        e2e_ms.append(1e3 * (time.perf_counter() - t0))
    h2d_bytes, d2h_bytes = r.h2d_bytes, r.d2h_bytes
    e2e_best = min(e2e_ms) if e2e_ms else float("nan")

    # ---------------- max over ranks
    t_total = torch.tensor([total_ms, e2e_best, float(cells_expected)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t_total.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t_total.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, e2e_best, cells_all = float(tmax[0]), float(tmax[1]), float(tsum[2])
    else:
        cells_all = float(cells_expected)

    if rank == 0:
        peak, peak_src = load_peak()
        ms_per_step = total_ms / args.steps
        value = cells_all / (ms_per_step * 1e-3) / 1e9
        achieved = alg_bytes / (pileup_ms * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of ONE resident launch on the default workload, from the ncu --set full
        # capture summarised in profiles/r01_k_pileup_classify_resident_chr1.txt (same seed, same launch)
        traffic = 8_922_642_000 if (args.scale == 1.0 and world == 1) else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/u32 integer", "data": "synthetic",
            "config": {"workload": "chr1-size synthetic 30x 2x150bp PE, pileup+classify+BED on one B200 per rank (BASELINE configs[1])",
                       "contig_bp": c.length, "reads": reads.n, "cells_per_rank": cells_expected, "bed_intervals": n_intervals, "scale": args.scale,
                       "l2_policy": f"inputs ({alg_bytes / 1e9:.2f} GB per step) are far larger than the 126 MB L2; no flush needed",
                       "parallelism": f"region shards x{world}, counters all-reduced" if world > 1 else "single GPU", **prep},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "k_pileup_classify", "kernel_ms": pileup_ms,
                         "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_cell": alg_bytes / cells_expected, "peak_source": peak_src},
            "e2e": {"value": cells_all / (e2e_best * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": e2e_best, "h2d_device_ms": round(float(r.h2d_ms), 2),
                    "note": "clb_begin_contig + clb_push_reads batches from pinned host columns + clb_finish_contig (D2H); PCIe-bound"},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks,
            "kernel_step_ms": {"min": min(step_ms), "median": statistics.median(step_ms), "max": max(step_ms)},
        }
        if not args.no_cpu_baseline and world == 1:
            from decodingustools_b200 import synth
            sample_bp = int(min(args.cpu_sample_mbp * 1e6, c.length))
            oc, orun, dt, sub = oracle_sample(c, c.reads, opt, sample_bp)
            line["cpu_baseline"] = {
                "value": oc.summed_coverage / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"first {sample_bp} bp of the same contig ({oc.summed_coverage} cells, {dt:.1f} s); oracle/callable_oracle.c "
                          f"single-threaded like the reference; host has {os.cpu_count()} cores"}
            # size-independent sanity of the full-size GPU result + exact parity on the sample prefix
            ctx.begin_contig(0, "chr1", sample_bp, c.ref[:sample_bp], sample_bp, max_ref_span=span)
            from decodingustools_b200.callable_loci import admit_reads
            from decodingustools_b200.callable_loci import compact_reads
            ctx.push_reads(compact_reads(sub, admit_reads(sub, opt.pileup_max_depth, 0)))
            g = ctx.finish_contig()
            ok = (g.state_counts.tolist() == oc.counts and g.summed_coverage == oc.summed_coverage and g.summed_baseq == oc.summed_baseq
                  and g.summed_mapq == oc.summed_mapq and g.quality_bases == oc.quality_bases)
            line["parity_on_cpu_sample"] = bool(ok)
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
