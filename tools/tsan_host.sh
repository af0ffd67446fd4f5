#!/bin/bash
# ThreadSanitizer over the C host pipeline (nk_host.c: worker pool, three-stage pipelines, upload claiming, waves)
# linked against the CPU emulation of the engine (tests/emu): SURVEY section 5, "TSAN on host C".
# usage: tools/tsan_host.sh [report-file]
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
S=$ROOT/nomalise_kmers_multi_large_b200/csrc
W=$(mktemp -d /tmp/nk_tsan.XXXXXX); trap 'rm -rf "$W"' EXIT
OUT=${1:-/dev/stdout}
gcc -O1 -g -std=gnu11 -fsanitize=thread -fPIC -pthread -Wno-format -DNK_LI_CHUNK_BYTES=4096 -DNK_ROLL_CHUNKS=8 -c -o $W/host.o $S/nk_host.c || exit 1
g++ -O1 -g -std=c++17 -fsanitize=thread -fPIC -pthread -c -o $W/emu.o $ROOT/tests/emu/nk_emu.cpp || exit 1
gcc -O1 -g -fsanitize=thread -c -o $W/main.o $S/nk_main.c || exit 1
g++ -fsanitize=thread -pthread -o $W/cli $W/main.o $W/host.o $W/emu.o || exit 1
$ROOT/tools/nk_synth -n 6000 -s 5 -o $W/s > /dev/null
$ROOT/tools/nk_synth -n 6000 -s 6 --equal -o $W/q > /dev/null   # files of equal size: split by size, ranges counted by the step builders
{
    for cfg in "NKB200_GPUS=2 NK_EMU_DEVICES=2" "NKB200_ROLLING_COUNT=1 NKB200_GPUS=2 NK_EMU_DEVICES=2" "NKB200_ROLLING_COUNT=1 NKB200_ENGINES_PER_GPU=4" "NKB200_HOST_PARSE=1" "NKB200_TABLE_BUDGET_MB=450" "NKB200_ENGINES_PER_GPU=4" \
               "IN=q NKB200_ENGINES_PER_GPU=4" "IN=q NKB200_GPUS=2 NK_EMU_DEVICES=2"; do
        mkdir -p $W/out && cd $W/out && rm -f output_*
        in=s; case "$cfg" in IN=q*) in=q;; esac
        env $cfg NKB200_STEP_PAIRS=128 NKB200_THREADS=6 TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" \
            timeout 1200 $W/cli -f $W/${in}_1.fastq -r $W/${in}_2.fastq -k 21 -c -p 8 -d 32 -m 1 -P -e > $W/stdout.txt 2> $W/stderr.txt
        echo "== $cfg: exit $?, $(grep -c 'WARNING: ThreadSanitizer' $W/stderr.txt) ThreadSanitizer warnings, $(grep -h 'Printed Records' $W/stdout.txt)"
        grep -A14 "WARNING: ThreadSanitizer" $W/stderr.txt | head -80
    done
} > "$OUT" 2>&1
