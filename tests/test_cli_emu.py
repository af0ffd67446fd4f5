"""Host logic on CPU: the product's C host pipeline (CLI, partitioners, indexer, staging, writer, seeding)
linked against the test-only emulated engine must reproduce the oracle's files and counters byte for byte."""
import subprocess
from pathlib import Path

import pytest

from tests import cli_cases as cc
from tests import oracle_lib as ol

ROOT = Path(__file__).resolve().parent.parent
EMU_CLI = ROOT / "tests" / "emu" / "nk_emu_cli"


@pytest.fixture(scope="module")
def cases(tmp_path_factory):
    subprocess.run(["make", "-C", str(ROOT / "tests" / "emu")], check=True, capture_output=True)
    ol.build_oracle()
    tmp = tmp_path_factory.mktemp("cli_emu")
    return tmp, dict(cc.standard_cases(tmp, 1200))


NAMES = ["canonical_p8", "stranded_k31_fa_growth", "dump_p4_k15", "equal_sizes_F6", "single_end",
         "single_end_fq2fa_empty", "fasta_in_out_mixed", "multi_file", "ragged_lengths", "tiny_k5",
         "one_partition_default_depth", "p64_canonical_config3_shape", "depth_coverage_sweep_point"]


@pytest.mark.parametrize("name", NAMES)
def test_cli_matches_oracle(cases, name):
    tmp, table = cases
    args = table[name]
    want = cc.run_cli(ol.ORACLE_CLI, args, tmp / name / "oracle")
    got = cc.run_cli(EMU_CLI, args, tmp / name / "emu", env={"NKB200_STEP_PAIRS": "64"})
    cc.assert_same(got, want, name)


@pytest.mark.parametrize("name", ["canonical_p8", "single_end", "p64_canonical_config3_shape"])
def test_cli_merged_table_and_output(cases, name):
    tmp, table = cases
    args = [a for a in table[name] if a != "-P"]
    n = cc.check_merged_extras(EMU_CLI, ol.ORACLE_CLI, args, tmp / ("merged_" + name), env={"NKB200_STEP_PAIRS": "64"})
    assert n > 1000


def test_cli_multi_device_placement(cases):
    """partition t on GPU t mod G (SURVEY 8.B row e): same bytes for any G, including the merged extras whose
    k-mers cross from the other GPUs through nkd_compact / nkd_merge_add"""
    tmp, table = cases
    args = table["canonical_p8"]
    env = {"NKB200_STEP_PAIRS": "64", "NK_EMU_DEVICES": "3", "NKB200_GPUS": "3"}
    want = cc.run_cli(ol.ORACLE_CLI, args, tmp / "multidev" / "oracle")
    got = cc.run_cli(EMU_CLI, args, tmp / "multidev" / "emu", env=env)
    verbose = cc.run_cli(EMU_CLI, args + ["-e"], tmp / "multidev" / "emu_verbose", env=env)
    assert "on 3 GPU(s)" in verbose["stdout"], verbose["stdout"][-400:]
    cc.assert_same(got, want, "3 emulated devices")
    cc.check_merged_extras(EMU_CLI, ol.ORACLE_CLI, args, tmp / "multidev_merged", env=env)


@pytest.mark.parametrize("argv,needle", [
    (["-k", "32"], "Only kmer sizes (32) of 5 to 31 are supported"),
    (["-d", "100", "-p", "64"], "must be at least 2 x number of CPUs"),
    (["-t", "fa"], "cannot request an output format of FASTQ when input is FASTA"),
    (["-g", "1.5"], "Coverage"),
])
def test_cli_rejects_what_the_reference_rejects(cases, argv, needle):
    """usage + exit 1 (C:704-743)"""
    tmp, table = cases
    f, r = table["canonical_p8"][1], table["canonical_p8"][3]
    for binary in (ol.ORACLE_CLI, EMU_CLI):
        res = cc.run_cli(binary, ["-f", f, "-r", r] + argv, tmp / "reject" / Path(binary).name)
        assert res["rc"] == 1 and needle in res["stderr"], (binary, res["stderr"][-300:])


def test_cli_fatal_on_non_dna(cases):
    """lowercase / IUPAC -> FATAL + exit 1 (C:1445-1454)"""
    tmp, table = cases
    f, r = Path(table["canonical_p8"][1]), Path(table["canonical_p8"][3])
    lines = f.read_bytes().split(b"\n")
    lines[4 * 700 + 1] = lines[4 * 700 + 1][:30] + b"r" + lines[4 * 700 + 1][31:]
    bad = tmp / "bad_1.fastq"
    bad.write_bytes(b"\n".join(lines))
    for binary in (ol.ORACLE_CLI, EMU_CLI):
        res = cc.run_cli(binary, ["-f", bad, "-r", r, "-k", 15, "-m", 1], tmp / "fatal" / Path(binary).name)
        assert res["rc"] == 1 and "FATAL: FWD sequence does not appear to be a DNA sequence" in res["stderr"]


@pytest.mark.parametrize("name", ["dump_p4_k15", "equal_sizes_F6", "multi_file", "single_end_fq2fa_empty", "fasta_in_out_mixed"])
def test_cli_stdout_matches_reference_binary(cases, name):
    """the stdout lines of SURVEY 8.B row (b), compared line by line with the reference binary's (timing fields masked)"""
    if not ol.REF_BIN.exists():
        pytest.skip("oracle/_ref not built (no reference checkout on this box)")
    tmp, table = cases
    args = table[name]
    binary = ol.REF_BIN_TLS if "-c" in args else ol.REF_BIN
    want = cc.run_cli(binary, args, tmp / name / "reference_stdout")
    got = cc.run_cli(EMU_CLI, args, tmp / name / "emu_stdout", env={"NKB200_STEP_PAIRS": "64"})
    cc.assert_same(got, want, name)
    cc.assert_same_stdout(got, want, name)
