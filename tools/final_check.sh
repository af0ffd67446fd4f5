#!/bin/bash
# what the driver runs at round end, in one go on a GPU box: smoke, the GPU tests, both bench arms
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --impl reference > gpurun_out/bench_reference.json 2>gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench_default.json 2>gpurun_out/bench_default.err
tail -1 gpurun_out/bench_default.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['isolated']['frac'], d['roofline']['share_of_step'], d['cpu_baseline']['value'], d['clocks'])"
tail -1 gpurun_out/bench_reference.json | cut -c1-300
