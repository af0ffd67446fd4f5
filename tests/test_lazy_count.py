"""Partition ranges that come without a count of the files (files of equal size are split by size, C:1807-1813; -p 1,
C:1796-1803) are counted by the engines' step builders as they go, a step ahead of the device, instead of before
the pipelines start.  Files of different size go through the record-count partitioner (C:1815-1828): both files are counted
first, or (NKB200_ROLLING_COUNT=1) the forward file first and the reverse file while the first partitions are already
being worked on.  The files here span many 256-KB index chunks and the steps are small, so that a partition's
index is extended many times, becomes exact near the end of its range, and hands over to the host parser where the
text is not regular.  Files and counters equal the oracle's whatever the route."""
import re
import subprocess
from pathlib import Path

import pytest

from nomalise_kmers_multi_large_b200 import capi
from tests import cli_cases as cc
from tests import oracle_lib as ol

ROOT = Path(__file__).resolve().parent.parent
EMU_CLI = ROOT / "tests" / "emu" / "nk_emu_cli"
STEPS_RE = re.compile(r"B200: (\d+) device steps on raw record text, (\d+) on host-parsed records")
ROUTE_RE = re.compile(r"B200: line ends counted before the first step for (\d+) input\(s\), reverse file alongside the first steps "
                      r"for (\d+), by the step builders for (\d+)")
N_PAIRS = 7500


@pytest.fixture(scope="module")
def inputs(tmp_path_factory):
    subprocess.run(["make", "-C", str(ROOT / "tests" / "emu")], check=True, capture_output=True)
    ol.build_oracle()
    tmp = tmp_path_factory.mktemp("lazy")
    f, r = cc.synth(tmp, "s", N_PAIRS, seed=77, read_len=100, equal=True)
    bf, br = f.read_bytes(), r.read_bytes()
    assert len(bf) == len(br) and len(bf) > 6 * (256 << 10)
    files = {"regular": (f, r)}

    def put(name, a, b, ext="fastq"):
        assert len(a) == len(b)
        pa, pb = tmp / f"{name}_1.{ext}", tmp / f"{name}_2.{ext}"
        pa.write_bytes(a)
        pb.write_bytes(b)
        files[name] = (pa, pb)

    put("no_trailing_newline", bf[:-1], br[:-1])
    lines = br.split(b"\n")
    q = 4 * 6700 + 3
    lines[q] = lines[q][:10] + b"\0" + lines[q][11:]            # NUL in a quality line of the second partition
    put("nul_in_quality", bf, b"\n".join(lines))
    # a sequence line of 1024+ chars in both files, sizes still equal: read_line cuts it (C:397) and what is left of it
    # becomes the next line of the record frame
    lf, lr = bf.split(b"\n"), br.split(b"\n")
    for ls in (lf, lr):
        ls[4 * 4000 + 1] = ls[4 * 4000 + 1] * 12
        ls[4 * 4000 + 3] = ls[4 * 4000 + 3] * 12
    put("long_sequence_line", b"\n".join(lf), b"\n".join(lr))
    fa_f, fa_r = cc.synth(tmp, "fa", N_PAIRS, seed=78, read_len=100, equal=True, fasta=True)
    files["fasta"] = (fa_f, fa_r)
    # files of different size: the record-count partitioner (C:1815-1828) needs the forward file counted to its end, the
    # reverse file is counted while the first partitions are worked on
    files["unequal"] = cc.synth(tmp, "u", N_PAIRS, seed=79, read_len=100)
    uf, ur = (x.read_bytes() for x in files["unequal"])
    assert len(uf) != len(ur)
    files["unequal_fewer_reverse_records"] = (files["unequal"][0], tmp / "u_short_2.fastq")
    (tmp / "u_short_2.fastq").write_bytes(b"\n".join(ur.split(b"\n")[:4 * 6000]) + b"\n")
    lines = ur.split(b"\n")
    lines[4 * 6100 + 1] = lines[4 * 6100 + 1][:30] + b"\0" + lines[4 * 6100 + 1][31:]      # NUL in a late sequence line
    files["unequal_nul"] = (files["unequal"][0], tmp / "u_nul_2.fastq")
    (tmp / "u_nul_2.fastq").write_bytes(b"\n".join(lines))
    return tmp, files


# name, extra arguments, expect host-parsed steps as well
CASES = [
    ("regular", ["-p", 3], False),
    ("regular", ["-p", 1], False),
    ("regular", ["-p", 4, "-c", "-o", "fa"], False),
    ("no_trailing_newline", ["-p", 1], True),     # (split by size, the last bytes of a file belong to no partition)
    ("nul_in_quality", ["-p", 3], True),
    ("long_sequence_line", ["-p", 3], None),
    ("fasta", ["-p", 2, "-t", "fa", "-o", "fa"], False),
    ("unequal", ["-p", 3, "-c"], False),
    ("unequal", ["-p", 8], False),
    ("unequal_fewer_reverse_records", ["-p", 4], False),
    ("unequal_nul", ["-p", 4], True),
]


def run(binary, inputs, name, extra, mixed, tag, env):
    tmp, files = inputs
    f, r = files[name]
    depth = max(6, 2 * extra[extra.index("-p") + 1])    # the reference insists on depth >= 2 x partitions
    args = ["-f", f, "-r", r, "-k", 21, "-d", depth, "-m", 1] + extra
    tag = f"{name}_{'_'.join(str(x) for x in extra)}_{tag}".replace("-", "")
    want = cc.run_cli(ol.ORACLE_CLI, args, tmp / tag / "oracle")
    got = cc.run_cli(binary, args + ["-e"], tmp / tag / "got", env=env)
    if want["rc"] != 0:
        # exit status and a message are the contract after a fatal exit (a seed record with a line of 1024+ chars is
        # refused with a message of its own)
        assert got["rc"] == want["rc"] and "FATAL" in want["stderr"]
        assert "FATAL" in got["stderr"] or "sequence line of" in got["stderr"]
        return None
    cc.assert_same(got, want, name)
    m = STEPS_RE.search(got["stdout"])
    assert m, got["stdout"][-400:]
    raw_steps, parsed_steps = int(m.group(1)), int(m.group(2))
    if mixed is not None:
        assert raw_steps > 0 and (parsed_steps > 0) == mixed, (raw_steps, parsed_steps)
    route = tuple(int(x) for x in ROUTE_RE.search(got["stdout"]).groups())
    by_records = name.startswith("unequal") and extra[extra.index("-p") + 1] > 1
    if env.get("NKB200_EAGER_COUNT") or (by_records and not env.get("NKB200_ROLLING_COUNT")):
        assert route == (1, 0, 0), route
    else:
        assert route == ((0, 1, 0) if by_records else (0, 0, 1)), route
    return got


def env_of(name):
    """the reverse file counted alongside the first steps is opt-in (NKB200_ROLLING_COUNT)"""
    env = {"NKB200_STEP_PAIRS": "256"}
    if name.startswith("unequal"):
        env["NKB200_ROLLING_COUNT"] = "1"
    return env


@pytest.mark.parametrize("name,extra,mixed", CASES, ids=[f"{c[0]}{i}" for i, c in enumerate(CASES)])
def test_ranges_counted_step_by_step_emu(inputs, name, extra, mixed):
    run(EMU_CLI, inputs, name, extra, mixed, "emu", env_of(name))


@pytest.mark.parametrize("name", ["regular", "unequal"])
def test_counting_first_gives_the_same_emu(inputs, name):
    a = run(EMU_CLI, inputs, name, ["-p", 3, "-c"], False, "lazy", env_of(name))
    b = run(EMU_CLI, inputs, name, ["-p", 3, "-c"], False, "eager", {"NKB200_STEP_PAIRS": "256", "NKB200_EAGER_COUNT": "1"})
    if name == "unequal":   # and the default: both files counted first
        c = run(EMU_CLI, inputs, name, ["-p", 3, "-c"], False, "default", {"NKB200_STEP_PAIRS": "256"})
        cc.assert_same(a, c, "alongside vs default")
    cc.assert_same(a, b, "step-by-step vs up-front count")
    if name == "regular":   # (while partitions are still being released, steps are cut differently)
        assert STEPS_RE.search(a["stdout"]).groups() == STEPS_RE.search(b["stdout"]).groups()


def test_single_end_split_by_size_emu(inputs):
    """pure single-end at -p > 1 is an uninitialised FILE* in the reference (oracle header D4): oracle only"""
    tmp, files = inputs
    f, _ = files["regular"]
    args = ["-f", f, "-s", "-k", 21, "-d", 6, "-m", 1, "-p", 3]
    want = cc.run_cli(ol.ORACLE_CLI, args, tmp / "single" / "oracle")
    got = cc.run_cli(EMU_CLI, args + ["-e"], tmp / "single" / "got", env={"NKB200_STEP_PAIRS": "256"})
    cc.assert_same(got, want, "single-end")
    assert int(STEPS_RE.search(got["stdout"]).group(1)) > 3


@pytest.mark.gpu
@pytest.mark.parametrize("name,extra,mixed", CASES, ids=[f"{c[0]}{i}" for i, c in enumerate(CASES)])
def test_ranges_counted_step_by_step_gpu(inputs, name, extra, mixed):
    run(capi.CLI_PATH, inputs, name, extra, mixed, "gpu", env_of(name))


@pytest.mark.gpu
def test_ranges_counted_step_by_step_default_steps_gpu(inputs):
    run(capi.CLI_PATH, inputs, "regular", ["-p", 8, "-c"], False, "gpu_default", {})
