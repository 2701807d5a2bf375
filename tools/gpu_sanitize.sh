#!/bin/bash
# compute-sanitizer over the GPU tests of this round's kernels.  usage: gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh r02san'
TAG=${1:-r02san}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_matcher.py tests/test_gpu_hasher.py tests/test_gpu_screen.py tests/test_gpu_pixels.py -m gpu -x -q -k "not headline and not full_size and not large_needle" > $OUT/memcheck.log 2>&1; echo "memcheck exit $?"; tail -4 $OUT/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_matcher.py -m gpu -x -q -k "rows_in_m or planted or auto" > $OUT/racecheck.log 2>&1; echo "racecheck exit $?"; tail -4 $OUT/racecheck.log
