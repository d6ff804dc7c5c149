#!/bin/bash
# usage: gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_round_multi.sh r05 N'
TAG=${1:-r05}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
echo "== multi-gpu parity check (N=$N)"; timeout 240 $RUN scripts/multi_gpu_check.py > $OUT/multi_check_$N.log 2>&1; echo "rc=$?"; grep -E "multi-gpu check|MULTI_GPU_CHECK|Error|error" $OUT/multi_check_$N.log | head -20
echo "== bench N=1"; timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_1.json 2> $OUT/bench_1.err; cut -c1-300 $OUT/bench_1.json
for G in 2 4 8; do
  if [ $G -le $N ]; then
    for P2P in "" "--no-p2p"; do
      echo "== bench N=$G $P2P"
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $G --steps 20 --warmup 3 $P2P > $OUT/bench_$G$P2P.json 2> $OUT/bench_$G$P2P.err
      echo "rc=$?"; cut -c1-1300 $OUT/bench_$G$P2P.json; grep -v "^\*\|OMP_NUM\|^$" $OUT/bench_$G$P2P.err | tail -4
    done
  fi
done
ls -la $OUT
