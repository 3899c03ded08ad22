for w in "$@"; do
  BM_LEAF_WAIT=$w python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('wait $w c4', round(d['value'],1), round(d['ms_per_step'],2))"
  BM_LEAF_WAIT=$w python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('wait $w c2', round(d['value'],1), 'bounce', round(r['bounce_launch']['ms_per_launch'],4))"
  BM_LEAF_WAIT=$w python scratch/sweep_c5.py 5 4 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('wait $w soup5 ref', round(d['reference']['mrays_per_s'],2), 'closest', round(d['closest']['mrays_per_s'],2))"
done
