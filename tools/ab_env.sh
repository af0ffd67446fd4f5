# A/B of one environment switch on the default bench: tools/ab_env.sh VAR VALUE [repeats]
VAR=$1; VAL=$2; REP=${3:-2}
show() { python -c "
import json; d=json.load(open('gpurun_out/sw.json'))
print('$1', 'value %.1fM' % (d['value']/1e6), 'ms %.1f' % d['ms_per_step'], 'e2e %.1fM' % (d['e2e']['value']/1e6), 'e2e_ms %.1f' % d['e2e']['ms_per_step'], d['counters']['printed'], {k: round(v,3) for k,v in d['e2e']['host_s_per_step'].items()})"; }
for i in $(seq $REP); do
  unset $VAR
  python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-isolated-probe 2>/dev/null | tail -1 > gpurun_out/sw.json; show "default      "
  export $VAR=$VAL
  python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-isolated-probe 2>/dev/null | tail -1 > gpurun_out/sw.json; show "$VAR=$VAL"
done
