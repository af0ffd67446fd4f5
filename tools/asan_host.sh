#!/bin/bash
# AddressSanitizer + UndefinedBehaviorSanitizer over the C host pipeline linked against the CPU emulation of the engine
# (the companion of tsan_host.sh).  usage: tools/asan_host.sh [report-file]
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
S=$ROOT/nomalise_kmers_multi_large_b200/csrc
W=$(mktemp -d /tmp/nk_asan.XXXXXX); trap 'rm -rf "$W"' EXIT
OUT=${1:-/dev/stdout}
SAN="-fsanitize=address,undefined -fno-omit-frame-pointer"
gcc -O1 -g -std=gnu11 $SAN -fPIC -pthread -Wno-format -DNK_LI_CHUNK_BYTES=4096 -DNK_ROLL_CHUNKS=8 -c -o $W/host.o $S/nk_host.c || exit 1
g++ -O1 -g -std=c++17 $SAN -fPIC -pthread -c -o $W/emu.o $ROOT/tests/emu/nk_emu.cpp || exit 1
gcc -O1 -g $SAN -c -o $W/main.o $S/nk_main.c || exit 1
g++ $SAN -pthread -o $W/cli $W/main.o $W/host.o $W/emu.o || exit 1
$ROOT/tools/nk_synth -n 5000 -s 5 -o $W/s > /dev/null 2>&1
$ROOT/tools/nk_synth -n 5000 -s 6 --equal -o $W/q > /dev/null 2>&1
head -c -1 $W/s_1.fastq > $W/t_1.fastq; head -c -1 $W/s_2.fastq > $W/t_2.fastq   # last record cut by the end of the file
{
    for cfg in "NKB200_GPUS=2 NK_EMU_DEVICES=2" "NKB200_ROLLING_COUNT=1 NKB200_ENGINES_PER_GPU=4" "NKB200_HOST_PARSE=1 NKB200_HOST_SEED=1" \
               "NKB200_TABLE_BUDGET_MB=450" "IN=q NKB200_ENGINES_PER_GPU=2" "IN=t NKB200_ROLLING_COUNT=1" "IN=q P=1"; do
        mkdir -p $W/out && cd $W/out && rm -f output_*
        in=s; p=8; case "$cfg" in IN=q*) in=q;; IN=t*) in=t;; esac; case "$cfg" in *P=1*) p=1;; esac
        env ${cfg/P=1/X=1} NKB200_STEP_PAIRS=128 NKB200_THREADS=6 ASAN_OPTIONS="detect_leaks=1" UBSAN_OPTIONS="print_stacktrace=1" \
            timeout 1200 $W/cli -f $W/${in}_1.fastq -r $W/${in}_2.fastq -k 21 -c -p $p -d 32 -m 1 -P -e > $W/stdout.txt 2> $W/stderr.txt
        echo "== $cfg: exit $?, $(grep -c 'ERROR: AddressSanitizer\|runtime error\|ERROR: LeakSanitizer' $W/stderr.txt) sanitizer reports, $(grep -h 'Printed Records' $W/stdout.txt)"
        grep -A12 "ERROR: AddressSanitizer\|runtime error\|ERROR: LeakSanitizer" $W/stderr.txt | head -60
    done
} > "$OUT" 2>&1
