"""Pins the oracle: (1) against the md5 goldens the survey recorded from the unmodified reference on its own
fixtures ('Golden vectors'), (2) against the reference binary built by oracle/Makefile,
run here on the same inputs.  The reference's only shipped known-answer is the comment at C:56-70."""
import hashlib
from pathlib import Path

import pytest

from tests import cli_cases as cc
from tests import oracle_lib as ol

FIX = Path("/root/reference/test")
needs_fixtures = pytest.mark.skipif(not FIX.exists(), reason="reference checkout (fixtures) not present on this box")


def cat_md5(cwd: Path, stem: str, k: int, norm: int, parts: int, ext="fastq"):
    h = hashlib.md5()
    for t in range(parts):
        h.update((cwd / f"{stem}.k{k}_norm{norm}_thread{t}.{ext}").read_bytes())
    return h.hexdigest()


GOLDEN = [
    # argv, (processed, printed, skipped, max used), k, norm, parts, md5 fwd, md5 rev, md5 dump
    (["-f", "a1.fastq", "-r", "b1.fastq", "-k", "15", "-d", "8", "-p", "1"], (5000, 4297, 703, 494772), 15, 8, 1,
     "3bf335a853dacaee40d90a79267e7e1a", "d747e2ba166555e4f324f04455ae5ad5", None),
    (["-f", "a1.fastq", "-r", "b1.fastq", "-k", "15", "-d", "8", "-p", "2"], (5000, 4212, 788, 494772), 15, 4, 2,
     "ba4e1e06aedc2d072aadc5907cef5a5c", "e395140bf92fc4fa8c0c930ad0b7450c", None),
    (["-f", "a1.fastq", "-r", "b1.fastq", "-k", "15", "-d", "8", "-p", "4"], (5000, 3931, 1069, 494772), 15, 2, 4,
     "553a1f70275dfb2224031192ae3cbbbc", "bdb2b6d384284d6b9ae556c1ccdc368e", None),
    (["-f", "a2.fastq", "-r", "b2.fastq", "-k", "15", "-m", "1", "-p", "8", "-d", "16", "-P"], (10000, 8013, 1987, 879127), 15, 2, 8,
     "bf47086702b9053d33db33d8c2d17b60", "79316c15e66f6c894ef28347a0d1b55d", "e040344a4e100aad2de95a36bb600ff9"),
    (["-f", "a2.fastq", "-r", "b2.fastq", "-k", "15", "-m", "1", "-p", "64", "-d", "128", "-P"], (10000, 9577, 423, 587516), 15, 2, 64,
     "dd9e4e47a02518106081e6e69076f0a3", "8bcd568c0d6a4cfc61d8e4d7611396f1", "97f1cb4e72ea80f8ac5c59381f5aef9e"),
    (["-f", "a2.fastq", "-r", "b2.fastq", "-k", "15", "-m", "1", "-p", "200", "-d", "400", "-P"], (10000, 9870, 130, 402794), 15, 2, 200,
     "2d3c561aa6c9a7cec2848423f41bce4b", "ea76b874d6f1d3e166df72e7df3bc1d6", "e07adb5ca7541798f901ac7024105711"),
    (["-f", "a2.fastq", "-r", "b2.fastq", "-k", "21", "-c", "-m", "1", "-p", "8", "-d", "16", "-P"], (10000, 6673, 3327, 685418), 21, 2, 8,
     "b4ad25dfb3f1249abe8f3d8bdfbab7df", "eef52816ce26f8bfea442ace601b8a51", "ecebb87b6ee2ded0b8b74236076fa460"),
    (["-f", "a1.fastq", "-r", "a1.fastq", "-k", "15", "-d", "8", "-p", "2"], (7445, 5608, 1837, 248238), 15, 4, 2,
     "9454638bd82210f33b380ca86235cb07", "9454638bd82210f33b380ca86235cb07", None),
    # mixed paired + single-end input lists (the second forward file has no mate)
    (["-f", "a1.fastq", "a2.fastq", "-r", "b1.fastq", "-s", "-k", "15", "-d", "4", "-p", "2"], (19948, 13957, 5991, 924786), 15, 2, 2,
     "de5059c8d6101832b462a81580232861", "090d6555b16001e594af630d7ef45212", None),
]


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
