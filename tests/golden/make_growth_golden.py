"""Generates tests/golden/growth_default_capacity.json: the unmodified REFERENCE binary (oracle/_ref/nkml) at the
default capacity (67,108,879 slots = 1 GiB) on an input with more distinct k-mers than 0.8 x that, so that the table
grows 67,108,879 -> 100,663,318 slots (C:1055-1108) -- the path the small -m 1 cases never reach.  1 M synthetic pairs
(tools/nk_synth -n 1000000 -s 7 -t 300000), -k 25 -p 1 -d 100; takes ~3 minutes:

    make -C oracle ref && python tests/golden/make_growth_golden.py
"""
import hashlib
import json
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
SYNTH = ["-n", "1000000", "-s", "7", "-t", "300000", "-L", "150"]
FLAGS = ["-k", "25", "-p", "1", "-d", "100"]


def main():
    ref = ROOT / "oracle" / "_ref" / "nkml"
    assert ref.exists(), "build the reference first: make -C oracle ref"
    subprocess.run(["make", "-C", str(ROOT / "tools")], check=True, capture_output=True)
    with tempfile.TemporaryDirectory(dir="/dev/shm") as d:
        d = Path(d)
        subprocess.run([str(ROOT / "tools" / "nk_synth")] + SYNTH + ["-o", str(d / "g")], check=True, capture_output=True)
        out = d / "ref"
        out.mkdir()
        p = subprocess.run([str(ref), "-f", str(d / "g_1.fastq"), "-r", str(d / "g_2.fastq")] + FLAGS, cwd=out,
                           capture_output=True, text=True, check=True)
        final = {k: int(re.search(k + r": ([\d,]+)", p.stdout).group(1).replace(",", ""))
                 for k in ("Processed Records", "Printed Records", "Skipped Records",
                           "Cumulative Max unique kmers in any thread")}
        files = {f.name: hashlib.md5(f.read_bytes()).hexdigest() for f in sorted(out.glob("output_*"))}
    gold = {"synth": " ".join(SYNTH), "flags": " ".join(FLAGS), "final": final, "files_md5": files,
            "capacity_sequence": [67108879, 100663318]}
    (Path(__file__).parent / "growth_default_capacity.json").write_text(json.dumps(gold, indent=1) + "\n")
    print(gold)


if __name__ == "__main__":
    sys.exit(main())
