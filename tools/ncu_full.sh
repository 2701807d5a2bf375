#!/bin/bash
# ncu --set full captures of the hot kernels (one GPU).  usage: bash tools/ncu_full.sh <tag> [kernels...]
TAG=${1:-r01}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for what in "$@"; do
  case $what in
    hash)   K=hash_fused_kernel ;;
    scan)   K=match_tiled_kernel ;;
    tensor) K=match_tensor4 ;;
    tensor32k) K=match_tensor4x2 ;;
    tensor128) K=match_tensor4x2 ;;
    rows8) K=match_rows_m ;;
    one) K=hash_one_kernel ;;
    mailbox1) K=match_rows_m ;;
    *)      K=$what ;;
  esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o $OUT/ncu_$what \
      python tools/profile_kernels.py $what > $OUT/ncu_$what.log 2>&1
  echo "ncu $what exit $?"
  ncu -i $OUT/ncu_$what.ncu-rep --page raw --csv > $OUT/ncu_${what}_raw.csv 2>/dev/null
  ncu -i $OUT/ncu_$what.ncu-rep --page source --csv > $OUT/ncu_${what}_src.csv 2>/dev/null
done
