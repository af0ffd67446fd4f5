#!/bin/bash
# A/B runs of the default bench under different environment settings: tools/ab_bench.sh "VAR=1 OTHER=2" "VAR=3" ...
# AB_GPUS=N runs under torchrun on N GPUs.  Each argument is one run; results are appended to gpurun_out/ab.txt (device ms, e2e ms, host stage seconds).
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/gpurun_out"
for cfg in "$@"; do
    if [ "${AB_GPUS:-1}" -gt 1 ]; then
        LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${AB_GPUS} --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400))"
    else
        LAUNCH="python"
    fi
    env $cfg timeout 400 $LAUNCH "$ROOT/bench.py" --gpus ${AB_GPUS:-1} --steps ${AB_STEPS:-3} --warmup 1 --no-cpu-baseline --no-isolated-probe ${AB_ARGS:-} \
        > "$ROOT/gpurun_out/ab_last.json" 2> "$ROOT/gpurun_out/ab_last.err" || tail -3 "$ROOT/gpurun_out/ab_last.err"
    python - "$cfg" "$ROOT/gpurun_out/ab_last.json" >> "$ROOT/gpurun_out/ab.txt" <<'PY'
import json, sys
try:
    a = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "| device ms", round(a["ms_per_step"], 1), "| e2e ms", round(a["e2e"]["ms_per_step"], 1), "|",
          {k: round(v, 3) for k, v in a["e2e"]["host_s_per_step"].items()}, "|",
          {k: round(v, 1) for k, v in a["kernel_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[1], "| failed:", e)
PY
done
cat "$ROOT/gpurun_out/ab.txt"
