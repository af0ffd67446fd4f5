"""Generates tests/golden/reference_goldens.json by running the REFERENCE binary (oracle/_ref/nkml, or nkml_tls for
--canonical with -p > 1) on seeded synthetic inputs (tools/nk_synth.c is deterministic, so the inputs are
re-created bit for bit wherever the tests run).  Run it in the container that has /root/reference:

    make -C oracle ref && python tests/golden/make_golden.py
"""
import json
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from tests import cli_cases as cc  # noqa: E402
from tests import oracle_lib as ol  # noqa: E402

# name, synth kwargs, argv after -f/-r
CASES = [
    ("canonical_k25_p8", dict(n_pairs=6000, seed=101), ["-k", 25, "-c", "-p", 8, "-d", 100, "-m", 1]),
    ("stranded_k31_fa_p4", dict(n_pairs=5000, seed=102), ["-k", 31, "-g", 0.96, "-o", "fa", "-m", 1, "-p", 4, "-d", 16]),
    ("dump_k15_p2", dict(n_pairs=4000, seed=103, read_len=100), ["-k", 15, "-p", 2, "-d", 8, "-m", 1, "-P"]),
    ("equal_sizes_p3", dict(n_pairs=5000, seed=104, equal=True), ["-k", 21, "-p", 3, "-d", 12, "-m", 1]),
    ("growth_p64", dict(n_pairs=8000, seed=105), ["-k", 25, "-p", 64, "-d", 256, "-m", 1]),
    ("one_partition", dict(n_pairs=3000, seed=106, read_len=75), ["-k", 20, "-m", 1, "-d", 6]),
]


def inputs_for(tmp: Path, name, kw):
    return cc.synth(tmp, name, kw["n_pairs"], seed=kw["seed"], read_len=kw.get("read_len", 150), equal=kw.get("equal", False))


def main():
    assert ol.REF_BIN.exists() and ol.REF_BIN_TLS.exists(), "build the reference first: make -C oracle ref"
    out = {}
    with tempfile.TemporaryDirectory() as d:
        tmp = Path(d)
        for name, kw, argv in CASES:
            f, r = inputs_for(tmp, name, kw)
            binary = ol.REF_BIN_TLS if "-c" in argv else ol.REF_BIN
            res = cc.run_cli(binary, ["-f", f, "-r", r] + argv, tmp / name)
            assert res["rc"] == 0, res["stderr"]
            out[name] = {"binary": binary.name, "counters": res["counters"], "final": res["final"], "files": res["files"]}
            print(name, res["counters"][-1], len(res["files"]), "files")
    (Path(__file__).parent / "reference_goldens.json").write_text(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
