"""N>1 on CPU: two gloo ranks, each owning half of the fixed -p partitions (the placement bench.py --gpus 2 uses),
must together produce exactly the files and counters of one process / of the oracle -- partitions share nothing."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

from tests import cli_cases as cc
from tests import oracle_lib as ol

ROOT = Path(__file__).resolve().parent.parent
EMU_LIB = ROOT / "tests" / "emu" / "libnk_emu.so"


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


import pytest


@pytest.mark.parametrize("mode", ["planned", "indexed"])
def test_two_ranks_equal_one_process(tmp_path, mode):
    subprocess.run(["make", "-C", str(ROOT / "tests" / "emu")], check=True, capture_output=True)
    ol.build_oracle()
    f, r = cc.synth(tmp_path, "s", 4000, seed=11)   # a few 256-KB chunks per file, so that both ranks count some
    parts, k, depth = 4, 21, 16
    want = cc.run_cli(ol.ORACLE_CLI, ["-f", f, "-r", r, "-k", k, "-c", "-p", parts, "-d", depth, "-m", 1], tmp_path / "oracle")
    out = tmp_path / "ranks"
    out.mkdir()
    port = free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, str(ROOT / "tests" / "_rank_worker.py"), str(f), str(r), str(out),
                                       str(parts), str(k), str(depth), str(EMU_LIB), mode], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    for p in procs:
        o, e = p.communicate(timeout=600)
        assert p.returncode == 0, e[-2000:]
    import hashlib
    got = {p.name: hashlib.md5(p.read_bytes()).hexdigest() for p in sorted(out.glob("output_*"))}
    assert got == want["files"]
    totals = json.loads((out / "totals.json").read_text())
    assert (totals["processed"], totals["printed"], totals["skipped"], totals["max_used"]) == want["counters"][-1]


def test_merged_extras_need_every_partition(tmp_path):
    """--merged-table / --merged-output are whole-run outputs: a context that owns a slice of the partitions
    (one rank of a multi-process launch) must refuse them instead of writing a partial file.  Runs in its own
    process: the emulation library exports the product's symbol names and must not share a process with it."""
    subprocess.run(["make", "-C", str(ROOT / "tests" / "emu")], check=True, capture_output=True)
    code = f"""
import ctypes, sys
sys.path.insert(0, {str(ROOT)!r})
from nomalise_kmers_multi_large_b200 import capi
from nomalise_kmers_multi_large_b200.pipeline import Pipeline
lib = ctypes.CDLL({str(EMU_LIB)!r})
capi._declare_engine(lib)
capi._declare_pipeline(lib)
for extra in ({{"merged_table": True}}, {{"merged_output": True}}):
    try:
        Pipeline(k=21, depth=16, partitions=4, memory_gb=1, out_dir={str(tmp_path)!r}, part_first=0, part_count=2, lib=lib, **extra)
    except capi.NkError as e:
        assert "owns all 4 partitions" in str(e), str(e)
    else:
        raise SystemExit("a partial context accepted " + str(extra))
with Pipeline(k=21, depth=16, partitions=4, memory_gb=1, out_dir={str(tmp_path)!r}, merged_table=True, merged_output=True,
              lib=lib) as p:
    assert p.totals()["engines"] >= 1
print("ok")
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == "ok", (r.stdout[-500:], r.stderr[-1500:])
