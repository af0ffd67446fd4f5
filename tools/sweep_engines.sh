# bench.py with 1, 2, 4 engines (streams + pipeline threads) per GPU
for epg in 1 2 4; do
  export NKB200_ENGINES_PER_GPU=$epg
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/sw.json
  python -c "
import json; d=json.load(open('gpurun_out/sw.json'))
print('engines/gpu $epg', 'value %.1fM' % (d['value']/1e6), 'ms %.1f' % d['ms_per_step'], 'e2e %.1fM' % (d['e2e']['value']/1e6), 'e2e_ms %.1f' % d['e2e']['ms_per_step'], {k: round(v,1) for k,v in d.get('kernel_ms_per_step',{}).items()}, d['counters']['printed'], d.get('host_stage_s'))"
done
