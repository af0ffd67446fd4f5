"""The host pipeline hands steps to the device as raw record text and keeps the byte-exact host parser for what
the device declines (NUL bytes, lines of 1024+ chars) or cannot be cut into whole records (a last record that the
end of the file cuts short).  Whatever the mix, files and counters equal the oracle's (and the reference's)."""
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
SEEDS_RE = re.compile(r"B200: (\d+) seed records taken from raw text on the device, (\d+) parsed by the host")


@pytest.fixture(scope="module")
def inputs(tmp_path_factory):
    subprocess.run(["make", "-C", str(ROOT / "tests" / "emu")], check=True, capture_output=True)
    ol.build_oracle()
    tmp = tmp_path_factory.mktemp("fallback")
    f, r = cc.synth(tmp, "s", 3000, seed=31, read_len=80)
    bf, br = f.read_bytes(), r.read_bytes()
    files = {"regular": (f, r)}

    def put(name, a, b):
        pa, pb = tmp / f"{name}_1.fastq", tmp / f"{name}_2.fastq"
        pa.write_bytes(a)
        pb.write_bytes(b)
        files[name] = (pa, pb)

    put("no_trailing_newline", bf[:-1], br[:-1])
    lines = bf.split(b"\n")
    lines[4 * 1700] = b"@" + b"h" * 1500                      # a header line read_line cuts at 1023 chars (C:397)
    put("long_header_mid_file", b"\n".join(lines), br)
    lines = br.split(b"\n")
    lines[4 * 2100 + 3] = lines[4 * 2100 + 3][:10] + b"\0" + lines[4 * 2100 + 3][11:]   # NUL in a quality line
    put("nul_in_quality", bf, b"\n".join(lines))
    put("reverse_has_fewer_records", bf, b"\n".join(br.split(b"\n")[:4 * 2500]) + b"\n")
    return tmp, files


CASES = [("regular", False), ("no_trailing_newline", True), ("long_header_mid_file", True), ("nul_in_quality", True),
         ("reverse_has_fewer_records", False)]


def run(binary, env, inputs, name, mixed, gpu):
    tmp, files = inputs
    f, r = files[name]
    args = ["-f", f, "-r", r, "-k", 21, "-d", 6, "-m", 1, "-p", 3, "-c"]
    tag = name + ("_gpu" if gpu else "_emu")
    want = cc.run_cli(ol.ORACLE_CLI, args, tmp / tag / "oracle")
    got = cc.run_cli(binary, args + ["-e"], tmp / tag / "got", env=env)
    if want["rc"] != 0:
        # a line of 1024+ chars shifts the reference's record frame (C:397) until something that is not DNA lands in a
        # sequence line: FATAL, exit 1; what is on disk by then is unspecified
        assert got["rc"] == want["rc"] and "FATAL" in want["stderr"] and "FATAL" in got["stderr"]
        return got
    cc.assert_same(got, want, name)
    m = STEPS_RE.search(got["stdout"])
    assert m, got["stdout"][-400:]
    raw_steps, parsed_steps = int(m.group(1)), int(m.group(2))
    seeds_raw, seeds_parsed = (int(x) for x in SEEDS_RE.search(got["stdout"]).groups())
    # seeding (C:1322-1373) reads the same files: whole records go to the device as raw text, the rest to the host parser
    assert seeds_raw > 0 and (seeds_parsed > 0) == (name == "nul_in_quality"), (seeds_raw, seeds_parsed)
    # an engine whose step holds text the device declines hands all of its partitions to the host parser from there on,
    # so a run on one engine may see no raw-text step at all
    assert raw_steps > 0 or mixed
    assert (parsed_steps > 0) == mixed, (name, raw_steps, parsed_steps)
    return got


@pytest.mark.parametrize("name,mixed", CASES)
def test_mix_of_raw_and_parsed_steps_emu(inputs, name, mixed):
    run(EMU_CLI, {"NKB200_STEP_PAIRS": "128"}, inputs, name, mixed, False)


def test_host_parse_switch_emu(inputs):
    tmp, files = inputs
    f, r = files["regular"]
    args = ["-f", f, "-r", r, "-k", 21, "-d", 6, "-m", 1, "-p", 3, "-e"]
    a = cc.run_cli(EMU_CLI, args, tmp / "sw_raw", env={"NKB200_STEP_PAIRS": "128"})
    b = cc.run_cli(EMU_CLI, args, tmp / "sw_parsed", env={"NKB200_STEP_PAIRS": "128", "NKB200_HOST_PARSE": "1"})
    c = cc.run_cli(EMU_CLI, args, tmp / "sw_zero", env={"NKB200_STEP_PAIRS": "128", "NKB200_HOST_PARSE": "0"})
    cc.assert_same(a, b, "raw vs host-parsed")
    assert STEPS_RE.search(b["stdout"]).group(1) == "0" and STEPS_RE.search(a["stdout"]).group(2) == "0"
    assert STEPS_RE.search(c["stdout"]).group(2) == "0", "NKB200_HOST_PARSE=0 must leave the switch off"


@pytest.mark.gpu
@pytest.mark.parametrize("name,mixed", CASES)
def test_mix_of_raw_and_parsed_steps_gpu(inputs, name, mixed):
    run(capi.CLI_PATH, {"NKB200_STEP_PAIRS": "128"}, inputs, name, mixed, True)


WAVES_RE = re.compile(r"B200: (\d+) wave\(s\) of partitions, (\d+) tables parked in host memory, (\d+) brought in")


def check_waves(binary, inputs, env, tag):
    """Tables that do not fit the GPU together: partitions are worked on in waves and the tables of the others wait in
    host memory (SURVEY 8.B row e, H6).  A small artificial budget forces that here; two input file pairs make every
    table go out and come back, and -P dumps them all at the end.  Nothing may change in any file."""
    tmp, files = inputs
    f, r = files["regular"]
    f2, r2 = files["reverse_has_fewer_records"]
    args = ["-f", f, f2, "-r", r, r2, "-k", 21, "-d", 16, "-m", 1, "-p", 8, "-c", "-P"]
    want = cc.run_cli(ol.ORACLE_CLI, args, tmp / tag / "oracle")
    got = cc.run_cli(binary, args + ["-e"], tmp / tag / "got", env=dict(env, NKB200_TABLE_BUDGET_MB="450"))
    cc.assert_same(got, want, "waves")
    m = WAVES_RE.search(got["stdout"])
    assert m, got["stdout"][-600:]
    waves, parked, loaded = (int(x) for x in m.groups())
    assert waves >= 4 and parked > 0 and loaded >= 8 + parked, (waves, parked, loaded)   # 2 files x >= 2 waves
    free = cc.run_cli(binary, args + ["-e"], tmp / tag / "free", env=env)
    assert WAVES_RE.search(free["stdout"]).groups() == ("2", "0", "0")   # everything fits: one wave per file, as before
    cc.assert_same(free, want, "no budget")


def test_waves_when_tables_do_not_fit_emu(inputs):
    check_waves(EMU_CLI, inputs, {"NKB200_STEP_PAIRS": "256"}, "waves_emu")


@pytest.mark.gpu
def test_waves_when_tables_do_not_fit_gpu(inputs):
    check_waves(capi.CLI_PATH, inputs, {"NKB200_STEP_PAIRS": "256"}, "waves_gpu")
