"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): region planning, halo read selection, counter
all-reduce, interval gathering + stitching.  The per-shard compute is stood in by slices of the oracle's per-base
output (tests may use the oracle as the checker); the GPU tests cover the kernel's own region handling."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from decodingustools_b200 import sharding, synth
from decodingustools_b200.callable_loci import INTERVAL_DTYPE, ContigDeviceResult, bin_geometry
from decodingustools_b200.options import CallableOptions
from oracle import oracle

WINDOW = 2047


def test_plan_regions_tiles_the_genome():
    for lens, ws in ([[100_000, 50_000, 16_569], 2], [[248_956_422, 242_193_529, 16_569], 8], [[5], 4], [[0, 7000, 0], 3], [[2047 * 5], 5],
                     [[49_791_284], 2], [[248_956_422], 8], [[248_956_422], 4]):
        plan = sharding.plan_regions(lens, ws, WINDOW)
        assert len(plan) == ws
        cover = {tid: [] for tid in range(len(lens))}
        for shards in plan:
            for s in shards:
                assert s.end > s.start
                cover[s.tid].append((s.start, s.end))
        for tid, L in enumerate(lens):
            pos = 0
            for a, b in sorted(cover[tid]):
                assert a == pos
                assert b == L or b % WINDOW == 0          # cuts fall on window multiples
                pos = b
            assert pos == L
        for shards in plan:                                   # pieces of one rank inside one contig are merged
            for a, b in zip(shards[:-1], shards[1:]):
                assert not (a.tid == b.tid and a.end == b.start)
        sizes = [sum(s.end - s.start for s in shards) for shards in plan]
        if sum(lens) > 50 * ws * WINDOW:
            assert max(sizes) - min(sizes) <= 2 * WINDOW + max(sizes) * 0.01


def test_reads_for_region_includes_every_overlapping_read():
    c = synth.synth_short("chr22", 200_000, seed=3)
    end = c.reads.end()
    for a, b in ((0, 50_000), (49_999, 50_001), (123_456, 200_000)):
        lo, hi = sharding.reads_for_region(c.reads, a, b)
        need = np.flatnonzero((c.reads.pos < b) & (end > a - 1))
        assert need.size == 0 or (need.min() >= lo and need.max() < hi)


def _shard_result(oc, name, length, largest, a, b) -> ContigDeviceResult:
    st = oc.state[a:b]
    starts = np.flatnonzero(np.concatenate([[True], st[1:] != st[:-1]]))
    iv = np.zeros(starts.size, INTERVAL_DTYPE)
    iv["start"] = starts + a; iv["end"] = np.concatenate([starts[1:], [b - a]]) + a; iv["state"] = st[starts]
    if a > 0 and starts.size and oc.state[a - 1] == st[0]:
        iv["soft_start"][0] = 1
    stride, nb = bin_geometry(name, length, largest)
    bins = np.zeros((3, nb), np.uint32)
    pos = np.arange(a, b)
    for row, s in enumerate((1, 5, 0)):
        np.add.at(bins[row], pos[st == s] // stride, 1)
    raw = oc.raw[a:b].astype(np.int64)
    return ContigDeviceResult(np.bincount(st, minlength=6).astype(np.uint64), int((raw > 0).sum()), int(raw.sum()),
                              0, 0, int(oc.qc[a:b].astype(np.int64).sum()), iv, bins, stride, a, b)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        opt = CallableOptions()
        c = synth.synth_short("chr22", 60_000, seed=21)
        o = oracle.OracleRun(opt, c.length)
        oc = o.process_contig(c.name, 0, c.length, c.ref, c.reads, debug=True)
        plan = sharding.plan_regions([c.length], world, WINDOW)
        mine = plan[rank]
        assert len(mine) == 1
        sh = mine[0]
        res = _shard_result(oc, c.name, c.length, c.length, sh.start, sh.end)
        res = sharding.allreduce_counters(res)
        iv = sharding.gather_and_stitch(res.intervals, sh.start, dst=0)
        if rank == 0:
            whole = _shard_result(oc, c.name, c.length, c.length, 0, c.length)
            ok = (np.array_equal(res.state_counts, whole.state_counts) and res.summed_coverage == whole.summed_coverage
                  and res.n_covered_bases == whole.n_covered_bases and res.quality_bases == whole.quality_bases
                  and np.array_equal(res.bins, whole.bins) and np.array_equal(iv, whole.intervals)
                  and res.state_counts.tolist() == oc.counts)
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_two_and_three_rank_reduction_and_stitching(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
