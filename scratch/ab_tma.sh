# A/B of the bulk-staged cooperative leaf (scratch/libs/libbeam_tma.so, -DBM_TMA_LEAF=1) against the default library
for v in default tma; do
  if [ $v = tma ]; then export BEAM_B200_LIB=$PWD/scratch/libs/libbeam_tma.so; else unset BEAM_B200_LIB; fi
  echo "== $v"; python scratch/c2_probe.py --reps 5 | tail -3
  python scratch/c2_probe.py --workload c4 --reps 4 | tail -2
done
BEAM_B200_LIB=$PWD/scratch/libs/libbeam_tma.so timeout 900 python -m pytest tests/test_gpu_trace.py tests/test_gpu_host_api.py -m gpu -x -q 2>&1 | tail -2
