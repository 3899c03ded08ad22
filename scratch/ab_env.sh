# usage: ab_env.sh VAR v1 v2 ...
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/ab_env.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$var=$v: value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['single_frame_latency_ms'],3), 'prim', round(r['ms_per_launch'],4), 'bounce', round(r['bounce_launch']['ms_per_launch'],4))" || tail -5 gpurun_out/ab_env.err
done
