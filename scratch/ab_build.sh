for v in "$@"; do
  echo "== $v"
  BEAM_B200_LIB=$PWD/scratch/libs/libbeam_$v.so BM_BUILD_TIMING=1 NB=4 python scratch/prof_build.py c2 2>&1 | tail -2 | cut -c1-260
  BEAM_B200_LIB=$PWD/scratch/libs/libbeam_$v.so NB=3 python scratch/prof_build.py soup20 2>&1 | tail -1 | sed 's/.*build/build/'
done
