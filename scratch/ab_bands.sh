for b in 1 2 3 4 6; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --bands $b 2>gpurun_out/bands_$b.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('bands $b: value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3), 'prim', round(r['ms_per_launch'],3))" || tail -5 gpurun_out/bands_$b.err
done
