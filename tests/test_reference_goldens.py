"""Golden vectors produced by the reference binary itself (tests/golden/make_golden.py, committed with its output):
the oracle must reproduce them on CPU, the CUDA program on the B200 -- also on boxes where the reference is absent."""
import json
import sys
from pathlib import Path

import pytest

from tests import cli_cases as cc
from tests import oracle_lib as ol

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "golden"))
import make_golden  # noqa: E402

GOLD = json.loads((ROOT / "tests" / "golden" / "reference_goldens.json").read_text())
CASES = {name: (kw, argv) for name, kw, argv in make_golden.CASES}


def check(binary, name, tmp_path, env=None):
    kw, argv = CASES[name]
    f, r = make_golden.inputs_for(tmp_path, name, kw)
    res = cc.run_cli(binary, ["-f", f, "-r", r] + argv, tmp_path / "run", env=env)
    want = GOLD[name]
    assert res["rc"] == 0, res["stderr"][-500:]
    assert [list(c) for c in res["counters"]] == want["counters"]
    assert res["final"] == want["final"]
    assert res["files"] == want["files"]


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_reference_goldens(name, tmp_path):
    ol.build_oracle()
    check(ol.ORACLE_CLI, name, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_b200_reproduces_reference_goldens(name, tmp_path):
    from nomalise_kmers_multi_large_b200 import capi
    check(capi.CLI_PATH, name, tmp_path)
