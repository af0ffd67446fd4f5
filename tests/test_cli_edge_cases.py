"""Edge inputs the reference handles in a defined way: the oracle is checked against the reference
binary when it is present, and the product's host pipeline (over the emulated engine on CPU) against the oracle."""
import subprocess
from pathlib import Path

import pytest

from tests import cli_cases as cc
from tests import oracle_lib as ol

ROOT = Path(__file__).resolve().parent.parent
EMU_CLI = ROOT / "tests" / "emu" / "nk_emu_cli"


def fq(records):
    return b"".join(b"@" + n + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n" for n, s in records)


@pytest.fixture(scope="module")
def inputs(tmp_path_factory):
    subprocess.run(["make", "-C", str(ROOT / "tests" / "emu")], check=True, capture_output=True)
    ol.build_oracle()
    tmp = tmp_path_factory.mktemp("edge")
    f, r = cc.synth(tmp, "s", 400, seed=21, read_len=60)
    base_f, base_r = f.read_bytes(), r.read_bytes()
    files = {}

    def put(name, a, b):
        pa, pb = tmp / f"{name}_1.fastq", tmp / f"{name}_2.fastq"
        pa.write_bytes(a)
        pb.write_bytes(b)
        files[name] = (pa, pb)

    put("no_trailing_newline", base_f[:-1], base_r[:-1])
    put("trailing_blank_line", base_f + b"\n", base_r + b"\n")
    # mates of length exactly K, K-1 and K+1 in the middle of the file (K = 21)
    recs_f = [(b"a%d/1" % i, b"ACGTTGCA" * 8) for i in range(30)]
    recs_r = [(b"a%d/2" % i, b"TTGACCAG" * 8) for i in range(30)]
    recs_f[7] = (b"k/1", b"ACGTACGTACGTACGTACGTA")          # 21 = K
    recs_f[9] = (b"km1/1", b"ACGTACGTACGTACGTACGT")         # 20 < K: pair vanishes (C:1430-1443)
    recs_r[11] = (b"km1/2", b"ACGTACGTACGTACGTACG")         # rev shorter than K
    recs_f[13] = (b"kp1/1", b"ACGTACGTACGTACGTACGTAC")      # 22
    recs_f[15] = (b"polyA/1", b"A" * 64)                     # every window has key 0: total = 0, ratio 0
    recs_f[17] = (b"withN/1", b"ACGTTGCANACGTTGCA" * 3)      # N -> A in the printed sequence (C:1426)
    put("length_gate", fq(recs_f), fq(recs_r))
    put("unequal_counts", fq(recs_f), fq(recs_r[:22]))
    put("crlf", base_f.replace(b"\n", b"\r\n"), base_r.replace(b"\n", b"\r\n"))
    long_f = list(recs_f)
    long_f[5] = (b"long/1", b"ACGT" * 400)                    # 1600 > 1023: read_line cuts the line (C:397)
    put("long_line", fq(long_f), fq(recs_r))
    return tmp, files


CASES = [
    ("no_trailing_newline", ["-k", 21, "-d", 4, "-m", 1], True),
    ("trailing_blank_line", ["-k", 21, "-d", 4, "-m", 1], True),
    ("length_gate", ["-k", 21, "-d", 2, "-m", 1], True),
    ("length_gate", ["-k", 21, "-d", 2, "-m", 1, "-c", "-o", "fa"], True),
    ("unequal_counts", ["-k", 21, "-d", 2, "-m", 1], True),
    ("crlf", ["-k", 21, "-d", 4, "-m", 1], True),                 # '\r' is not DNA -> FATAL, exit 1 (C:1445-1454)
    ("long_line", ["-k", 21, "-d", 2, "-m", 1], False),            # reference behaviour after the cut depends on stack reuse
]


@pytest.mark.parametrize("name,extra,vs_reference", CASES, ids=[f"{c[0]}-{i}" for i, c in enumerate(CASES)])
def test_edge_input(inputs, name, extra, vs_reference):
    tmp, files = inputs
    f, r = files[name]
    args = ["-f", f, "-r", r] + extra
    tag = name + "_" + "_".join(str(x) for x in extra)
    want = cc.run_cli(ol.ORACLE_CLI, args, tmp / tag / "oracle")
    got = cc.run_cli(EMU_CLI, args, tmp / tag / "emu", env={"NKB200_STEP_PAIRS": "16"})
    assert got["rc"] == want["rc"], (got["stderr"][-300:], want["stderr"][-300:])
    if want["rc"] == 0:
        assert got["counters"] == want["counters"] and got["files"] == want["files"]
    else:   # what is on disk after a fatal exit is unspecified; the exit status and a message are the contract
        assert "FATAL" in want["stderr"]
        assert "FATAL" in got["stderr"] or "sequence line of" in got["stderr"]
    if vs_reference and ol.REF_BIN.exists():
        ref = cc.run_cli(ol.REF_BIN_TLS if "-c" in extra else ol.REF_BIN, args, tmp / tag / "ref")
        assert ref["rc"] == want["rc"]
        if want["rc"] == 0:
            assert ref["counters"] == want["counters"] and ref["files"] == want["files"]


def test_n_is_printed_as_a(inputs):
    """the emitted sequence line carries N -> A; header and quality are verbatim."""
    tmp, files = inputs
    f, r = files["length_gate"]
    out = tmp / "n_to_a"
    res = cc.run_cli(EMU_CLI, ["-f", f, "-r", r, "-k", 21, "-d", 1000000, "-m", 1, "-g", 1.0], out)
    assert res["rc"] == 0
    text = (out / "output_forward.k21_norm1000000_thread0.fastq").read_bytes()
    assert b"@withN/1\nACGTTGCAAACGTTGCA" in text and b"N" not in text.split(b"@withN/1\n")[1].split(b"\n")[0]


def test_record_cut_short_by_a_nul_is_still_scored(inputs):
    """C:1622-1631, C:1733: when read_line gives up inside a record (a NUL byte here), the reference still scores and
    counts what it has read and only then leaves the loop.  Counters and the -P table must agree with the reference
    binary; the record's own printed lines are undefined there (stale stack bytes), so read files are compared between
    the oracle and the product only."""
    tmp, files = inputs
    f, r = files["trailing_blank_line"]
    lines = f.read_bytes().split(b"\n")
    lines[4 * 250 + 1] = lines[4 * 250 + 1][:40] + b"\0" + lines[4 * 250 + 1][41:]   # NUL inside a forward sequence line
    cut = tmp / "nul_1.fastq"
    cut.write_bytes(b"\n".join(lines))
    args = ["-f", cut, "-r", r, "-k", 21, "-d", 4, "-m", 1, "-P"]
    want = cc.run_cli(ol.ORACLE_CLI, args, tmp / "nul" / "oracle")
    got = cc.run_cli(EMU_CLI, args, tmp / "nul" / "emu", env={"NKB200_STEP_PAIRS": "16"})
    cc.assert_same(got, want, "NUL in a sequence line")
    assert want["counters"][-1][0] == 251   # 250 whole records and the cut one
    if ol.REF_BIN.exists():
        ref = cc.run_cli(ol.REF_BIN, args, tmp / "nul" / "ref")
        assert ref["rc"] == want["rc"] and ref["counters"] == want["counters"] and ref["final"] == want["final"]
        dumps = [n for n in want["files"] if n.startswith("output_kmer.")]
        assert dumps and all(ref["files"][n] == want["files"][n] for n in dumps)
