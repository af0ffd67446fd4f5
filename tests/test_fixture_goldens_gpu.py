"""The reference's own fixtures through the CUDA path: the survey's golden list (md5s recorded from the unmodified
reference) and BASELINE.json configs[0] exactly as SURVEY 8(d) makes it concrete
(-k 15 -d 100 -g 0.9 -p 1, default capacity of 67,108,879 slots) on a1/b1 and a2/b2, against the oracle and the
reference binary."""
import hashlib

import pytest

from nomalise_kmers_multi_large_b200 import capi
from tests import cli_cases as cc
from tests import oracle_lib as ol
from tests.fixtures import GOLDEN, cat_md5, fixture_dir

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("g", GOLDEN, ids=lambda g: " ".join(g[0][4:]))
def test_cuda_cli_reproduces_survey_goldens(tmp_path, g):
    argv, counters, k, norm, parts, md5f, md5r, md5d = g
    fix = fixture_dir()
    argv = [str(fix / a) if a.endswith(".fastq") else a for a in argv]
    res = cc.run_cli(capi.CLI_PATH, argv, tmp_path)
    assert res["rc"] == 0, res["stderr"][-400:]
    assert res["counters"][-1] == counters
    assert cat_md5(tmp_path, "output_forward", k, norm, parts) == md5f
    assert cat_md5(tmp_path, "output_reverse", k, norm, parts) == md5r
    if md5d:
        assert cat_md5(tmp_path, "output_kmer", k, norm, parts, "tsv") == md5d


def test_cuda_cli_known_answer_from_reference_comment(tmp_path):
    """C:56-70: 2seq.fastq, single-end, k = 15, depth 2"""
    res = cc.run_cli(capi.CLI_PATH, ["-f", fixture_dir() / "2seq.fastq", "-s", "-k", 15, "-d", 2], tmp_path)
    assert res["rc"] == 0 and res["counters"][-1] == (4, 2, 2, 91)
    assert hashlib.md5((tmp_path / "output_forward.k15_norm2_thread0.fastq").read_bytes()).hexdigest() == \
        "cdbda297d5a4b5aa995748e4b5f0b6b0"


@pytest.mark.parametrize("pair", [("a1", "b1"), ("a2", "b2")])
def test_config_c1_default_capacity(tmp_path, pair):
    """BASELINE configs[0]: -k 15 -d 100 -g 0.9 -p 1 at the default capacity (1 GiB table), byte-exact"""
    ol.build_oracle()
    fix = fixture_dir()
    argv = ["-f", fix / f"{pair[0]}.fastq", "-r", fix / f"{pair[1]}.fastq", "-k", 15, "-d", 100, "-g", 0.9, "-p", 1]
    want = cc.run_cli(ol.ORACLE_CLI, argv, tmp_path / "oracle")
    got = cc.run_cli(capi.CLI_PATH, argv, tmp_path / "b200")
    cc.assert_same(got, want, "C1 " + pair[0])
    if ol.REF_BIN.exists():
        cc.assert_same(got, cc.run_cli(ol.REF_BIN, argv, tmp_path / "reference"), "C1 vs reference " + pair[0])


def test_default_capacity_table_grows_like_the_reference(tmp_path):
    """1 GiB table -> 1.5 GiB (67,108,879 -> 100,663,318 slots, C:1055-1108) during seeding and scoring, checked against
    what the unmodified reference wrote for the same input (tests/golden/growth_default_capacity.json)."""
    import json
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    gold = json.loads((root / "tests" / "golden" / "growth_default_capacity.json").read_text())
    subprocess.run(["make", "-C", str(root / "tools")], check=True, capture_output=True)
    subprocess.run([str(root / "tools" / "nk_synth")] + gold["synth"].split() + ["-o", str(tmp_path / "g")], check=True,
                   capture_output=True)
    res = cc.run_cli(capi.CLI_PATH, ["-f", tmp_path / "g_1.fastq", "-r", tmp_path / "g_2.fastq"] + gold["flags"].split() + ["-e"],
                     tmp_path / "out")
    assert res["rc"] == 0, res["stderr"][-400:]
    assert res["final"] == gold["final"]
    assert res["files"] == gold["files_md5"]
    assert gold["final"]["Cumulative Max unique kmers in any thread"] > 0.8 * gold["capacity_sequence"][0]
