#!/usr/bin/env python
"""Write tests/golden/: small real input files (BAM + BAI + FASTA + FAI) and the outputs the `coverage` command is
expected to produce for them, so that anyone with a Rust toolchain can run the reference binary on the same files and
diff -- the one thing that would pin this repository's parity to the real reference.

    python scripts/make_golden_bundle.py            # regenerates tests/golden/ (deterministic: seeded)

Every case directory holds
    in.bam, in.bam.bai, ref.fa, ref.fa.fai        inputs
    expected.callable_regions.bed                  what callable_regions.bed must be, byte for byte
    expected.summary.json                          what summary.json must be (bed_file "callable_regions.bed")
    COMMAND.txt                                    the exact reference command line (/root/reference/src/cli.rs:14-61)
and tests/golden/manifest.json lists flags and sha256 sums.

PROVENANCE: the expected.* files are produced by this repository's CPU ORACLE (oracle/callable_oracle.c, a restatement
of the reference loop and of htslib's pileup iterator), NOT by the reference binary, which cannot be built in the
image this repository is developed in (Rust + rust-htslib, no toolchain, no network).  Parity stays "unpinned" until
the commands in COMMAND.txt have been run with the real `decodingus-tools` and the outputs compared (or swapped in).
tests/test_golden.py checks the oracle and the CUDA path against these files.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import bamio                                                       # noqa: E402
from tests.golden_cases import FLAG_NAMES, cases, expected_outputs, qname, sha  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    manifest = {"generated_by": "scripts/make_golden_bundle.py", "expected_outputs_from": "oracle/callable_oracle.c (NOT the reference binary)", "cases": []}
    for name, contigs, opt in cases():
        d = os.path.join(OUT, name)
        os.makedirs(d)
        bam, fa = os.path.join(d, "in.bam"), os.path.join(d, "ref.fa")
        bamio.write_bam(bam, [(n, l, r) for n, l, _, r in contigs], qname_fn=qname, index=True, block=0x8000)
        bamio.write_fasta(fa, [(n, ref) for n, _, ref, _ in contigs])
        header_text, refs, cols, names = bamio.read_bam(bam)            # what a decoder sees (round trip of the writer)
        bed, summary = expected_outputs([(n, l, ref, cols[tid]) for tid, (n, l, ref, _) in enumerate(contigs)], opt, header_text, names)
        open(os.path.join(d, "expected.callable_regions.bed"), "wb").write(bed)
        open(os.path.join(d, "expected.summary.json"), "w").write(summary)
        flags = []
        for attr, flag, default in FLAG_NAMES:
            v = getattr(opt, attr)
            if v != default:
                flags += [flag, repr(v) if isinstance(v, float) else str(v)]
        cmd = "decodingus-tools coverage in.bam -r ref.fa -o callable_regions.bed" + ("" if not flags else " " + " ".join(flags))
        open(os.path.join(d, "COMMAND.txt"), "w").write(
            "# run inside this directory with the reference binary (cargo build --release in JamesKane/DecodingUsTools):\n"
            f"{cmd}\n"
            "cmp callable_regions.bed expected.callable_regions.bed\n"
            "cmp summary.json expected.summary.json      # integers must match; floats to 1e-9 relative\n")
        manifest["cases"].append({"name": name, "flags": flags, "contigs": [{"name": n, "length": l, "reads": int(r.n)} for n, l, _, r in contigs],
                                  "sha256": {f: sha(os.path.join(d, f)) for f in ("in.bam", "in.bam.bai", "ref.fa", "ref.fa.fai",
                                                                                  "expected.callable_regions.bed", "expected.summary.json")}})
    json.dump(manifest, open(os.path.join(OUT, "manifest.json"), "w"), indent=1)
    open(os.path.join(OUT, "README.md"), "w").write("# tests/golden\n\n" + __doc__.split("\n", 1)[1].strip() + "\n")
    total = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(OUT) for f in fs)
    print(f"wrote {len(manifest['cases'])} cases, {total / 1e6:.2f} MB, to {OUT}")


if __name__ == "__main__":
    main()
