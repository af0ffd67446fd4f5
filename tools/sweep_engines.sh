# bench.py with several engines-per-GPU / step-size settings
for cfg in "4 32768" "8 32768" "8 65536" "4 65536" "8 16384" "4 16384"; do
  set -- $cfg
  export NKB200_ENGINES_PER_GPU=$1
  if [ "$2" = "0" ]; then unset NKB200_STEP_PAIRS; else export NKB200_STEP_PAIRS=$2; fi
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/sw.json
  python -c "
import json; d=json.load(open('gpurun_out/sw.json'))
print('engines/gpu, step_pairs: $cfg', 'value %.1fM' % (d['value']/1e6), 'ms %.1f' % d['ms_per_step'], 'e2e %.1fM' % (d['e2e']['value']/1e6), 'e2e_ms %.1f' % d['e2e']['ms_per_step'], {k: round(v,1) for k,v in d.get('kernel_ms_per_step',{}).items()}, d['counters']['printed'], d['e2e']['host_s_per_step'])"
done
