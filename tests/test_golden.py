"""tests/golden/: real BAM/BAI/FASTA/FAI files with the outputs `coverage` must produce for them (see tests/golden/README.md
for the provenance: expected by the oracle, to be confirmed with the reference binary).  CPU: the fixtures are intact and
the oracle reproduces them from the files; GPU: the C++ `coverage` command over the C ABI reproduces them byte for byte."""
import json
import os
import shutil
import subprocess

import pytest

from tests import bamio
from tests.golden_cases import expected_outputs, sha

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
MANIFEST = json.load(open(os.path.join(GOLDEN, "manifest.json")))
CASES = [c["name"] for c in MANIFEST["cases"]]
CLI = os.path.join(os.path.dirname(HERE), "decodingustools_b200", "decodingus-tools-b200")


def _case(name):
    return next(c for c in MANIFEST["cases"] if c["name"] == name)


def _options(flags):
    from decodingustools_b200.options import CallableOptions
    kw = {}
    for k, v in zip(flags[::2], flags[1::2]):
        key = k[2:].replace("-", "_")
        kw[key] = float(v) if key == "max_low_mapq_fraction" else int(v)
    return CallableOptions(**kw)


def test_bundle_covers_the_known_answers_and_config_miniatures():
    assert len(CASES) >= 13 and {"ka1_no_reads", "ka6_depth_cap_3", "depth_cap_pile_default_500", "mini_config1_30x",
                                 "mini_config4_deep_cap500", "mini_config5_long_reads"} <= set(CASES)


@pytest.mark.parametrize("name", CASES)
def test_fixture_files_are_intact_and_the_oracle_reproduces_them(name):
    case = _case(name)
    d = os.path.join(GOLDEN, name)
    for f, digest in case["sha256"].items():
        assert sha(os.path.join(d, f)) == digest, f
    text, refs, cols, names = bamio.read_bam(os.path.join(d, "in.bam"))
    seqs = dict(bamio.read_fasta(os.path.join(d, "ref.fa")))
    assert [(n, l) for n, l in refs] == [(c["name"], c["length"]) for c in case["contigs"]]
    contigs = [(n, l, seqs[n], cols[tid]) for tid, (n, l) in enumerate(refs)]
    bed, summary = expected_outputs(contigs, _options(case["flags"]), text, names)
    assert bed == open(os.path.join(d, "expected.callable_regions.bed"), "rb").read()
    assert summary == open(os.path.join(d, "expected.summary.json")).read()
    json.loads(summary)                               # well-formed


def test_known_answer_bed_files_are_the_hand_derived_ones():
    """these few are small enough to state literally."""
    rd = lambda n: open(os.path.join(GOLDEN, n, "expected.callable_regions.bed")).read()
    assert rd("ka1_no_reads") == "c1\t0\t2\tREF_N\nc1\t2\t10\tNO_COVERAGE\n"
    assert rd("ka2_four_reads") == "c1\t0\t2\tREF_N\nc1\t2\t7\tCALLABLE\nc1\t7\t10\tNO_COVERAGE\n"
    assert rd("ka3_deletion") == "c1\t0\t2\tREF_N\nc1\t2\t4\tCALLABLE\nc1\t4\t5\tLOW_COVERAGE\nc1\t5\t7\tCALLABLE\nc1\t7\t10\tNO_COVERAGE\n"
    assert rd("ka4_low_mapq_two_of_ten") == "c1\t0\t2\tREF_N\nc1\t2\t7\tPOOR_MAPPING_QUALITY\nc1\t7\t10\tNO_COVERAGE\n"
    assert rd("ka4_low_mapq_one_of_ten") == "c1\t0\t2\tREF_N\nc1\t2\t7\tCALLABLE\nc1\t7\t10\tNO_COVERAGE\n"
    assert rd("ka5_contig_boundary_quirk") == "a\t0\t3\tNO_COVERAGE\n" * 2 + "b\t0\t2\tNO_COVERAGE\n"
    assert rd("ka6_depth_cap_3") == "c\t0\t1\tCALLABLE\nc\t1\t5\tEXCESSIVE_COVERAGE\nc\t5\t6\tCALLABLE\n"
    assert rd("ka7_ref_n_lowercase_iupac") == "c1\t0\t2\tREF_N\nc1\t2\t7\tCALLABLE\nc1\t7\t10\tNO_COVERAGE\n"
    ka3 = json.load(open(os.path.join(GOLDEN, "ka3_deletion", "expected.summary.json")))["export"]["contigs"][0]
    assert ka3["quality_stats"]["average_mapq"] == 75.0          # quirk Q5


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_coverage_command_reproduces_the_golden_outputs(name, tmp_path):
    case = _case(name)
    for f in ("in.bam", "in.bam.bai", "ref.fa", "ref.fa.fai"):
        shutil.copy(os.path.join(GOLDEN, name, f), tmp_path / f)
    p = subprocess.run([CLI, "coverage", "in.bam", "-r", "ref.fa", "-o", "callable_regions.bed", *case["flags"]], cwd=tmp_path,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    assert (tmp_path / "callable_regions.bed").read_bytes() == open(os.path.join(GOLDEN, name, "expected.callable_regions.bed"), "rb").read()
    assert (tmp_path / "summary.json").read_text() == open(os.path.join(GOLDEN, name, "expected.summary.json")).read()
