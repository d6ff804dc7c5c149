#!/bin/bash
# usage: gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_round_multi.sh r04 N'
TAG=${1:-r04}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
echo "== multi-gpu parity check (N=$N)"; timeout 600 $RUN scripts/multi_gpu_check.py > $OUT/multi_check_$N.log 2>&1; echo "rc=$?"; grep -E "multi-gpu check|MULTI_GPU_CHECK_OK|Error|error" $OUT/multi_check_$N.log | head -20
echo "== bench N=1"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_1.json 2> $OUT/bench_1.err; cat $OUT/bench_1.json | cut -c1-400
for G in 2 4 8; do
  if [ $G -le $N ]; then
    echo "== bench N=$G"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $G --steps 20 --warmup 3 > $OUT/bench_$G.json 2> $OUT/bench_$G.err; echo "rc=$?"; cat $OUT/bench_$G.json | cut -c1-1200; tail -5 $OUT/bench_$G.err
  fi
done
ls -la $OUT
