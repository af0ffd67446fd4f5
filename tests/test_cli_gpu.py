"""B200 parity of the whole drop-in program: normalise_kmers_multi_large_b200 (C host + sm_100a kernels)
against the oracle CLI and, where it travelled with the snapshot, the reference binary itself
(oracle/_ref/nkml, or nkml_tls for --canonical with -p > 1)."""
from pathlib import Path

import pytest

from nomalise_kmers_multi_large_b200 import capi
from tests import cli_cases as cc
from tests import oracle_lib as ol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cases(tmp_path_factory):
    ol.build_oracle()
    assert capi.CLI_PATH.exists(), "build the CLI first (__graft_entry__.build)"
    tmp = tmp_path_factory.mktemp("cli_gpu")
    return tmp, dict(cc.standard_cases(tmp, 20000))


NAMES = ["canonical_p8", "stranded_k31_fa_growth", "dump_p4_k15", "equal_sizes_F6", "single_end",
         "single_end_fq2fa_empty", "fasta_in_out_mixed", "multi_file", "ragged_lengths", "tiny_k5",
         "one_partition_default_depth", "p64_canonical_config3_shape", "depth_coverage_sweep_point"]


@pytest.mark.parametrize("name", NAMES)
def test_cli_matches_oracle(cases, name):
    tmp, table = cases
    args = table[name]
    want = cc.run_cli(ol.ORACLE_CLI, args, tmp / name / "oracle")
    got = cc.run_cli(capi.CLI_PATH, args, tmp / name / "b200")
    cc.assert_same(got, want, name)


@pytest.mark.parametrize("step_pairs", ["64", "1000"])
def test_cli_result_independent_of_step_size(cases, step_pairs):
    tmp, table = cases
    args = table["stranded_k31_fa_growth"]
    want = cc.run_cli(ol.ORACLE_CLI, args, tmp / "growth_oracle")
    got = cc.run_cli(capi.CLI_PATH, args, tmp / f"growth_b200_{step_pairs}", env={"NKB200_STEP_PAIRS": step_pairs})
    cc.assert_same(got, want, f"step_pairs={step_pairs}")


@pytest.mark.parametrize("name,ref", [("dump_p4_k15", ol.REF_BIN), ("canonical_p8", ol.REF_BIN_TLS),
                                      ("equal_sizes_F6", ol.REF_BIN)])
def test_cli_matches_reference_binary(cases, name, ref):
    if not Path(ref).exists():
        pytest.skip("reference binary did not travel with this snapshot")
    tmp, table = cases
    args = table[name]
    want = cc.run_cli(ref, args, tmp / name / "reference")   # sleeps 1 s per partition (C:1879)
    got = cc.run_cli(capi.CLI_PATH, args, tmp / name / "b200_vs_ref")
    cc.assert_same(got, want, name)
    cc.assert_same_stdout(got, want, name)   # every stdout line of the reference, timing fields masked


@pytest.mark.parametrize("name", ["canonical_p8", "p64_canonical_config3_shape"])
def test_cli_merged_table_and_output(cases, name):
    """device-side dump formatting, compaction, sort + reduce for the merged table"""
    tmp, table = cases
    n = cc.check_merged_extras(capi.CLI_PATH, ol.ORACLE_CLI, table[name], tmp / ("merged_" + name))
    assert n > 1000


def test_cli_fatal_on_non_dna(cases):
    tmp, table = cases
    f, r = Path(table["canonical_p8"][1]), Path(table["canonical_p8"][3])
    lines = f.read_bytes().split(b"\n")
    lines[4 * 7000 + 1] = lines[4 * 7000 + 1][:30] + b"r" + lines[4 * 7000 + 1][31:]
    bad = tmp / "bad_1.fastq"
    bad.write_bytes(b"\n".join(lines))
    res = cc.run_cli(capi.CLI_PATH, ["-f", bad, "-r", r, "-k", 15, "-m", 1], tmp / "fatal")
    assert res["rc"] == 1 and "FATAL: FWD sequence does not appear to be a DNA sequence" in res["stderr"]


def test_results_do_not_depend_on_gpu_count(cases):
    """Partitions are independent (README:68): spreading -p over more GPUs must not change a byte."""
    lib = capi.load_library()
    if lib.nkd_device_count() < 2:
        pytest.skip("one GPU visible")
    tmp, table = cases
    args = table["canonical_p8"]
    one = cc.run_cli(capi.CLI_PATH, args, tmp / "g1", env={"NKB200_GPUS": "1"})
    two = cc.run_cli(capi.CLI_PATH, args, tmp / "g2", env={"NKB200_GPUS": "2"})
    cc.assert_same(two, one, "2 GPUs vs 1")
    cc.check_merged_extras(capi.CLI_PATH, ol.ORACLE_CLI, args, tmp / "g2_merged", env={"NKB200_GPUS": "2"})
