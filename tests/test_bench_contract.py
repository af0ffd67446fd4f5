"""bench.py's reference arm end to end on CPU (SURVEY 8.B row d): the unmodified reference binary (oracle/_ref) timed
on a small sample must print one JSON line with the contract's keys.  The CUDA arm needs a GPU and is exercised by the
driver; its JSON is built by the same code path (run_b200) from the library's counters."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

from tests import oracle_lib as ol

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_the_contract_line():
    if not ol.REF_BIN_TLS.exists():
        pytest.skip("oracle/_ref not built (no reference checkout on this box)")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--pairs", "2000"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-500:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["ms_per_step"] > 0
    assert "pairs/sec" in d["metric"] and d["config"]["workload"] and d["config"]["pairs"] == 2000
    # both arms describe the workload with the same function: the driver compares the two config objects
    sys.path.insert(0, str(ROOT))
    import bench
    args = type("A", (), {"workload": "c2", "pairs": 2000, "memory": 0})()
    assert d["config"] == bench.config_of(bench.workload_of(args))
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
