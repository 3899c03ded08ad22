for k in 8 16 32; do
  BM_REFILL_MIN=$k python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('new-unit fetch at >=$k idle: c2', round(d['value'],1), round(d['ms_per_step'],3), round(r['ms_per_launch'],3), round(r['bounce_launch']['ms_per_launch'],3))"
done
python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('default: c2', round(d['value'],1), round(d['ms_per_step'],3), round(r['ms_per_launch'],3), round(r['bounce_launch']['ms_per_launch'],3))"
python scratch/sweep_c5.py 10 8 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('soup ref', round(d['reference']['mrays_per_s'],1), 'closest', round(d['closest']['mrays_per_s'],1))"
timeout 300 python -m pytest tests/test_gpu_trace.py tests/test_gpu_host_api.py -m gpu -x -q 2>&1 | tail -2
