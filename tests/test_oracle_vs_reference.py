"""Pins the oracle: (1) against the md5 goldens the survey recorded from the unmodified reference on its own
fixtures ('Golden vectors'), (2) against the reference binary built by oracle/Makefile,
run here on the same inputs.  The reference's only shipped known-answer is the comment at C:56-70."""
import hashlib
from pathlib import Path

import pytest

from tests import cli_cases as cc
from tests import oracle_lib as ol

from tests.fixtures import GOLDEN, cat_md5, fixture_dir

FIX = fixture_dir()   # the reference's own test/*.fastq, committed (gzipped) under tests/golden/fixtures
needs_fixtures = pytest.mark.skipif(not FIX.exists(), reason="fixtures missing")


@needs_fixtures
@pytest.mark.parametrize("g", GOLDEN, ids=lambda g: " ".join(g[0][4:]))
def test_oracle_reproduces_survey_goldens(tmp_path, g):
    argv, counters, k, norm, parts, md5f, md5r, md5d = g
    ol.build_oracle()
    argv = [str(FIX / a) if a.endswith(".fastq") else a for a in argv]
    res = cc.run_cli(ol.ORACLE_CLI, argv, tmp_path)
    assert res["rc"] == 0
    assert res["counters"][-1] == counters
    assert cat_md5(tmp_path, "output_forward", k, norm, parts) == md5f
    assert cat_md5(tmp_path, "output_reverse", k, norm, parts) == md5r
    if md5d:
        assert cat_md5(tmp_path, "output_kmer", k, norm, parts, "tsv") == md5d


@needs_fixtures
def test_oracle_reproduces_survey_counts_k31_fasta_out(tmp_path):
    """SURVEY 8 golden list: a1,b1 -k 31 -d 4 -p 2 -o fa -> 5000 / 3784 / 1216, max used 399,565 (the survey recorded no
    md5 for this one; the files are compared with the reference binary in test_oracle_matches_reference_binary)."""
    ol.build_oracle()
    res = cc.run_cli(ol.ORACLE_CLI, ["-f", FIX / "a1.fastq", "-r", FIX / "b1.fastq", "-k", 31, "-d", 4, "-p", 2, "-o", "fa"], tmp_path)
    assert res["rc"] == 0 and res["counters"][-1] == (5000, 3784, 1216, 399565)
    first = (tmp_path / "output_forward.k31_norm2_thread0.fastq").read_bytes().split(b"\n", 2)
    assert first[0].startswith(b">") and first[0].endswith(b"/1")   # fq -> fa appends /1, /2 (SURVEY known-answers)


@needs_fixtures
def test_oracle_known_answer_from_reference_comment(tmp_path):
    """C:56-70: 2seq.fastq single-end k=15 depth 2 -> high/total 0/73, 73/73, 70/73, 58/73."""
    recs = (FIX / "2seq.fastq").read_bytes().split(b"\n")
    tab = ol.OracleTable(67108879)
    seqs = [recs[i] for i in range(1, len(recs), 4) if recs[i]]
    for s in seqs:   # seeding stores every k-mer with count 0 first (C:1322-1373)
        tab.seed(s, 15, False)
    got = [tab.score(s, 15, False, 2) for s in seqs]
    assert got == [(0, 73), (73, 73), (70, 73), (58, 73)]
    res = cc.run_cli(ol.ORACLE_CLI, ["-f", FIX / "2seq.fastq", "-s", "-k", 15, "-d", 2], tmp_path)
    assert res["counters"][-1] == (4, 2, 2, 91)
    assert hashlib.md5((tmp_path / "output_forward.k15_norm2_thread0.fastq").read_bytes()).hexdigest() == \
        "cdbda297d5a4b5aa995748e4b5f0b6b0"


@needs_fixtures
@pytest.mark.parametrize("argv,ref", [
    (["-f", "a1.fastq", "a2.fastq", "-r", "b1.fastq", "-s", "-k", "15", "-d", "4", "-p", "2", "-m", "1"], "nkml"),
    (["-f", "a1.fastq", "-r", "b1.fastq", "-k", "31", "-d", "4", "-p", "2", "-o", "fa", "-m", "1"], "nkml"),
    (["-f", "a2.fastq", "-r", "b2.fastq", "-k", "25", "-c", "-p", "3", "-d", "9", "-m", "1", "-P"], "nkml_tls"),
])
def test_oracle_matches_reference_binary(tmp_path, argv, ref):
    binary = ol.ORACLE_DIR / "_ref" / ref
    if not binary.exists():
        pytest.skip("oracle/_ref not built")
    ol.build_oracle()
    argv = [str(FIX / a) if a.endswith(".fastq") else a for a in argv]
    want = cc.run_cli(binary, argv, tmp_path / "ref")
    got = cc.run_cli(ol.ORACLE_CLI, argv, tmp_path / "oracle")
    cc.assert_same(got, want)


def test_oracle_matches_reference_on_synthetic(tmp_path):
    """Same check on generated data, so it also runs where the reference's fixtures are absent but the binary travelled."""
    if not ol.REF_BIN.exists():
        pytest.skip("oracle/_ref not built")
    ol.build_oracle()
    f, r = cc.synth(tmp_path, "s", 3000, seed=9)
    argv = ["-f", f, "-r", r, "-k", 25, "-p", 2, "-d", 8, "-m", 1, "-P"]
    cc.assert_same(cc.run_cli(ol.ORACLE_CLI, argv, tmp_path / "oracle"), cc.run_cli(ol.REF_BIN, argv, tmp_path / "ref"))
