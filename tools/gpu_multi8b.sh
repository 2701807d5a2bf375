#!/bin/bash
# 8-GPU bench + 4-GPU bench + NCCL / peer-memory parity test.  usage: gpurun --gpus 8 --timeout 1200 -- 'bash tools/gpu_multi8b.sh r03s8'
TAG=${1:-r03s8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > $OUT/bench_8gpu.json 2> $OUT/bench_8gpu.err; echo "bench N=8 exit $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 --no-worst-case > $OUT/bench_4gpu.json 2> $OUT/bench_4gpu.err; echo "bench N=4 exit $?"
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q > $OUT/pytest_sharded.log 2>&1; tail -3 $OUT/pytest_sharded.log
for f in $OUT/*.err; do echo "== $f"; tail -n 3 $f; done; true
