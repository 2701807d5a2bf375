#!/bin/bash
# One multi-GPU gpurun call: N-GPU contract bench, BASELINE config 5 (1B-row DB sharded over N GPUs, 1024 images per
# batch), the NCCL parity test.  usage: gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_multi.sh r01m 8'
TAG=${1:-r01m}; N=${2:-8}; MODE=${3:-full}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $OUT/gpus.txt 2>&1
nvidia-smi topo -m >> $OUT/gpus.txt 2>&1
run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps 10 --warmup 3 > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err; echo "bench N=$N exit $?"; cat $OUT/bench_${N}gpu.json
[ "$MODE" = light ] && exit 0
run --steps 5 --warmup 3 --db-rows 1000000000 --frames $((1024 / N)) --no-e2e > $OUT/bench_${N}gpu_1b.json 2> $OUT/bench_${N}gpu_1b.err; echo "config5 exit $?"; cat $OUT/bench_${N}gpu_1b.json
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q > $OUT/pytest_sharded.log 2>&1; tail -3 $OUT/pytest_sharded.log
for f in $OUT/*.err; do tail -n 3 $f; done; true
