#!/usr/bin/env bash
# scratch: the round's ncu captures (run under gpurun; reports land in gpurun_out/)
set -x
N="ncu --set full --clock-control none -f"
raw() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_rawpage.csv 2>/dev/null; [ "$2" = keep ] || rm -f gpurun_out/$1.ncu-rep; }
for w in c2 c3 c4; do
  $N -k regex:trace_kernel --launch-skip 4 --launch-count 2 -o gpurun_out/r3c_$w python scratch/c2_probe.py --workload $w --reps 3 > gpurun_out/r3c_$w.log 2>&1
  raw r3c_$w $([ $w = c2 ] && echo keep)
done
$N -k regex:trace_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/r3c_hbm_direct python scratch/c5_probe.py --mtris 20 --mrays 8 --modes 0 --bricks 0 --reps 1 > gpurun_out/r3c_hbm_direct.log 2>&1; raw r3c_hbm_direct
$N -k "regex:trace_kernel|brick_segments|brick_resolve|onesweep" --launch-skip 7 --launch-count 7 -o gpurun_out/r3c_hbm_bricked python scratch/c5_probe.py --mtris 20 --mrays 8 --modes 0 --bricks 1 --reps 1 > gpurun_out/r3c_hbm_bricked.log 2>&1; raw r3c_hbm_bricked
$N -k regex:trace_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/r3c_c5 python scratch/c5_probe.py --mtris 100 --mrays 16 --modes 0 --reps 1 > gpurun_out/r3c_c5.log 2>&1; raw r3c_c5
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r3c_launches_bench_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-hbm-leg > gpurun_out/r3c_ncu_bench.log 2>&1
du -sh gpurun_out; ls -la gpurun_out/r3c_*
