#!/bin/bash
# Mid-scale parity + capacity check on a GPU box (BASELINE configs[2] flags: -k 25 -c -p 64 -d 256):
#  1. 2 M synthetic pairs, -m 16 (256 MiB table per partition): product CLI vs the CPU oracle CLI, md5 of every output;
#  2. the same input at the DEFAULT capacity (67,108,879 slots = 1 GiB per partition, 64 GiB of tables on one GPU):
#     product only (the oracle would need 64 GiB of host RAM), to show the default-capacity -p 64 case fits and runs.
# usage: tools/scale_check.sh [pairs] > gpurun_out/scale_check.txt
set -u
PAIRS=${1:-2000000}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d /dev/shm/nk_scale.XXXXXX)
trap 'rm -rf "$W"' EXIT
"$ROOT/tools/nk_synth" -n "$PAIRS" -o "$W/s" -s 7 | tail -1
ARGS="-f $W/s_1.fastq -r $W/s_2.fastq -k 25 -c -p 64 -d 256"
mkdir -p "$W/oracle" "$W/b200" "$W/b200_default"
t0=$(date +%s.%N)
(cd "$W/oracle" && "$ROOT/oracle/nk_oracle" $ARGS -m 16 > log.txt 2>&1); echo "oracle rc $?"
t1=$(date +%s.%N)
(cd "$W/b200" && "$ROOT/nomalise_kmers_multi_large_b200/csrc/normalise_kmers_multi_large_b200" $ARGS -m 16 -e > log.txt 2>&1); echo "b200 rc $?"
t2=$(date +%s.%N)
python3 - "$t0" "$t1" "$t2" <<'PY'
import sys
t0, t1, t2 = map(float, sys.argv[1:])
print("wall: oracle %.1f s, b200 %.1f s" % (t1 - t0, t2 - t1))
PY
(cd "$W/oracle" && md5sum output_* | sort -k2) > "$W/o.md5"
(cd "$W/b200" && md5sum output_* | sort -k2) > "$W/b.md5"
echo "files: $(wc -l < "$W/o.md5") oracle, $(wc -l < "$W/b.md5") b200"
if cmp -s "$W/o.md5" "$W/b.md5"; then echo "PARITY OK: all outputs byte-identical"; else echo "PARITY FAILED"; diff "$W/o.md5" "$W/b.md5" | head; fi
grep -h "Final Report" -A6 "$W/oracle/log.txt" | head -8
grep -h "Final Report" -A6 "$W/b200/log.txt" | head -8
grep -h "^B200:" "$W/b200/log.txt"
(cd "$W/b200_default" && "$ROOT/nomalise_kmers_multi_large_b200/csrc/normalise_kmers_multi_large_b200" $ARGS -e > log.txt 2>&1); echo "b200 default-capacity rc $?"
grep -h "Initial hash\|Printed Records\|^B200:" "$W/b200_default/log.txt"
nvidia-smi --query-gpu=memory.used,memory.total --format=csv,noheader
