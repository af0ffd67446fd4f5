#!/bin/bash
# One large run of the drop-in binary with BASELINE configs[2] flags (-k 25 -c -p 64 -d 256, default capacity:
# 64 x 1 GiB tables on one GPU) on N synthetic pairs; prints the program's own report and the B200 stage times.
# usage: tools/scale_run.sh [pairs] [extra args...]     (NK_FLAGS overrides "-k 25 -c -p 64 -d 256")
set -u
PAIRS=${1:-20000000}; shift || true
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d /dev/shm/nk_scale.XXXXXX)
trap 'rm -rf "$W"' EXIT
free -g | head -2
"$ROOT/tools/nk_synth" -n "$PAIRS" -o "$W/s" -s 11 -t $((PAIRS / 500)) | tail -1
mkdir -p "$W/out"
t0=$(date +%s.%N)
(cd "$W/out" && "$ROOT/nomalise_kmers_multi_large_b200/csrc/normalise_kmers_multi_large_b200" -f "$W/s_1.fastq" -r "$W/s_2.fastq" ${NK_FLAGS:--k 25 -c -p 64 -d 256} -e "$@" > log.txt 2>&1); echo "rc $?"
t1=$(date +%s.%N)
python3 -c "print('wall %.2f s' % ($t1 - $t0))"
grep -h "Initial hash\|Seeding took\|Final Report\|Records:\|Cumulative Max\|Total runtime\|Overall processing\|^B200:" "$W/out/log.txt"
grep -h "^\[nk" "$W/out/log.txt" | grep -v "seed flush" | head -${NK_DEBUG_LINES:-0}
ls -la "$W/out" | grep "output_kmer" | awk '{s+=$5} END {printf "k-mer table text: %.2f GB in %d files\n", s/1e9, NR}'
ls "$W/out" | wc -l; du -sh "$W/out" | cut -f1
nvidia-smi --query-gpu=memory.used,memory.total --format=csv,noheader
