for v in "$@"; do
  BEAM_B200_LIB=$PWD/scratch/libs/libbeam_$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/ab_$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$v: value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'prim', round(r['ms_per_launch'],4), 'bounce', round(r['bounce_launch']['ms_per_launch'],4), 'build', round(d['build']['ms'],4))" || tail -5 gpurun_out/ab_$v.err
done
