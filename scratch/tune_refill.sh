for k in 1 4 8 16 24 32; do
  BM_REFILL_MIN=$k python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('refill>=$k: c2', round(d['value'],1), round(d['ms_per_step'],3), round(r['ms_per_launch'],3), round(r['bounce_launch']['ms_per_launch'],3))"
done
