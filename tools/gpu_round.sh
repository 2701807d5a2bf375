#!/bin/bash
# One gpurun call: GPU parity tests, contract bench (both arms), developer micro-benches, ncu launch list.
# usage (from repo root): gpurun --timeout 1800 -- 'bash tools/gpu_round.sh [tag] [variant libs...]'
TAG=${1:-r01}
shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
lscpu | head -20 >> $OUT/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cat $OUT/bench.json; tail -3 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref exit $?"
cat $OUT/bench_ref.json
timeout 600 python tools/quick_bench.py scan hash hashpx tc4p > $OUT/quick.log 2>&1
cat $OUT/quick.log
for v in "$@"; do
    YUME_PDQ_CUDA_LIB=$PWD/yume_pdq_b200/libyume_pdq_cuda_$v.so timeout 300 python tools/quick_bench.py tc4p > $OUT/quick_$v.log 2>&1
    cat $OUT/quick_$v.log
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_under_ncu.log 2>&1
echo "ncu launches exit $?"
