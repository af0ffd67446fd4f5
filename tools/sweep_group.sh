for cfg in "0 0" "4 65536" "2 131072" "1 262144" "2 65536" "1 131072" "4 32768"; do
  set -- $cfg
  if [ "$1" = "0" ]; then unset NKB200_GROUP NKB200_STEP_PAIRS; else export NKB200_GROUP=$1 NKB200_STEP_PAIRS=$2; fi
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/sw.json
  python -c "
import json; d=json.load(open('gpurun_out/sw.json'))
print('group/sp $cfg', 'value %.1fM' % (d['value']/1e6), 'ms %.1f' % d['ms_per_step'], 'e2e %.1fM' % (d['e2e']['value']/1e6), {k: round(v,1) for k,v in d.get('kernel_ms_per_step',{}).items()}, d['counters']['printed'])"
done
