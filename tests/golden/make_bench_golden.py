"""Generates tests/golden/bench_c2_10M.json: the REFERENCE binary (oracle/_ref/nkml_tls, the thread-local variant
needed for --canonical with -p > 1) on the benchmark workload itself -- BASELINE.json configs[1],
10,000,000 synthetic pairs from tools/nk_synth (deterministic, so bench.py re-creates the same bytes on the GPU
box), -k 25 -c -p 8 -d 100, default capacity.  bench.py hashes its own outputs in one untimed pass and refuses to
print a value when they differ from this file.  Takes ~6 minutes and ~25 GB of /dev/shm:

    make -C oracle ref && python tests/golden/make_bench_golden.py [pairs]
"""
import hashlib
import json
import re
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
PAIRS = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
FLAGS = ["-k", "25", "-c", "-p", "8", "-d", "100"]


def md5_of(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 24), b""):
            h.update(chunk)
    return h.hexdigest()


def main():
    ref = ROOT / "oracle" / "_ref" / "nkml_tls"
    assert ref.exists(), "build the reference first: make -C oracle ref"
    subprocess.run(["make", "-C", str(ROOT / "tools")], check=True, capture_output=True)
    with tempfile.TemporaryDirectory(dir="/dev/shm") as d:
        d = Path(d)
        subprocess.run([str(ROOT / "tools" / "nk_synth"), "-n", str(PAIRS), "-s", "1", "-t", "20000", "-L", "150",
                        "-o", str(d / "bench")], check=True, capture_output=True)
        out = d / "ref"
        out.mkdir()
        t0 = time.time()
        p = subprocess.run([str(ref), "-f", str(d / "bench_1.fastq"), "-r", str(d / "bench_2.fastq")] + FLAGS, cwd=out,
                           capture_output=True, text=True, check=True)
        wall = time.time() - t0
        final = {k: int(re.search(k + r": ([\d,]+)", p.stdout).group(1).replace(",", ""))
                 for k in ("Processed Records", "Printed Records", "Skipped Records",
                           "Cumulative Max unique kmers in any thread")}
        files = {f.name: md5_of(f) for f in sorted(out.glob("output_*"))}
        inputs = {f.name: md5_of(f) for f in (d / "bench_1.fastq", d / "bench_2.fastq")}
    gold = {"pairs": PAIRS, "flags": " ".join(FLAGS), "generator": "tools/nk_synth -n %d -s 1 -t 20000 -L 150" % PAIRS,
            "reference": "oracle/_ref/nkml_tls (gcc -O2 of /root/reference/normalise_kmers_multi_large.c, rev_comp thread-local)",
            "reference_wall_s": round(wall, 1), "inputs_md5": inputs, "final": final, "files_md5": files}
    name = "bench_c2_10M.json" if PAIRS == 10_000_000 else f"bench_c2_{PAIRS}.json"
    (Path(__file__).parent / name).write_text(json.dumps(gold, indent=1) + "\n")
    print(json.dumps(gold["final"]), "->", name)


if __name__ == "__main__":
    main()
