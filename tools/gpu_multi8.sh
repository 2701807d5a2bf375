#!/bin/bash
# One 8-GPU gpurun call: host->device ceiling probe at 1/2/4/8 ranks, the 8-GPU contract bench (with config 4 and 5),
# the NCCL parity test.  usage: gpurun --gpus 8 --timeout 1200 -- 'bash tools/gpu_multi8.sh r02s8'
TAG=${1:-r02s8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $OUT/gpus.txt 2>&1
nvidia-smi topo -m >> $OUT/gpus.txt 2>&1
lscpu | head -25 >> $OUT/gpus.txt; numactl -H >> $OUT/gpus.txt 2>&1; free -g >> $OUT/gpus.txt
for N in 1 2 4 8; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/h2d_probe.py > $OUT/h2d_${N}.json 2> $OUT/h2d_${N}.err; echo "h2d N=$N exit $?"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > $OUT/bench_8gpu.json 2> $OUT/bench_8gpu.err; echo "bench N=8 exit $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 --no-worst-case > $OUT/bench_4gpu.json 2> $OUT/bench_4gpu.err; echo "bench N=4 exit $?"
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q > $OUT/pytest_sharded.log 2>&1; tail -3 $OUT/pytest_sharded.log
for f in $OUT/*.err; do echo "== $f"; tail -n 3 $f; done; true
