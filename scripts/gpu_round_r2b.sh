#!/bin/bash
# Round-2 multi-GPU call (N = 2 by default): full GPU test suite incl. the one-process multi-device tests, bench at N.
# usage: gpurun --gpus 2 --timeout 1500 -- 'bash scripts/gpu_round_r2b.sh r41 2'
TAG=${1:-r41}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt; free -g | head -2 >> $OUT/gpu.txt
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -s > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log; grep "multi-gpu check\|full solve" $OUT/pytest.log; tail -8 $OUT/pytest.log
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench rc=$?"; cut -c1-3000 $OUT/bench_n$N.json; tail -5 $OUT/bench_n$N.err
ls -la $OUT
