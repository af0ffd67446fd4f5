#!/bin/bash
# run on a GPU box: sampling profile of the host pipeline over the real engine, resolved to functions/modules
set -u
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
W=$(mktemp -d /dev/shm/nk_prof.XXXXXX); trap 'rm -rf "$W"' EXIT
PAIRS=${1:-10000000}
"$ROOT/tools/nk_synth" -n "$PAIRS" -o "$W/s" -s 1 | tail -1
nproc
NK_HOSTBENCH_PROF=$W/prof.txt "$ROOT/tools/hostbench/nk_hostlib_gpu" "$W/s_1.fastq" "$W/s_2.fastq" 8 3
python3 - "$W/prof.txt" "$ROOT/tools/hostbench/nk_hostlib_gpu" <<'PY'
import subprocess, sys, collections
addrs = [int(x, 16) for x in open(sys.argv[1]).read().split()]
maps = []
for line in open(sys.argv[1] + ".maps"):
    f = line.split()
    lo, hi = (int(x, 16) for x in f[0].split("-"))
    maps.append((lo, hi, f[5] if len(f) > 5 else "[anon]"))
exe = sys.argv[2]
mod = collections.Counter(); inexe = []
for a in addrs:
    m = next((n for lo, hi, n in maps if lo <= a < hi), "?")
    mod[m.split("/")[-1]] += 1
    if m.endswith("nk_hostlib_gpu"):
        inexe.append(a)
print("samples", len(addrs)); print("by module:", mod.most_common(8))
out = subprocess.run(["addr2line", "-f", "-e", exe] + [hex(a) for a in inexe], capture_output=True, text=True).stdout.split("\n")
fn = collections.Counter(out[0::2]); print("in nk_host.c:", fn.most_common(15))
ln = collections.Counter(l.split("/")[-1] for l in out[1::2]); print("hot lines:", ln.most_common(25))
PY
