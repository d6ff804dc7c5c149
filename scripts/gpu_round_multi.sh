#!/bin/bash
# usage: gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_round_multi.sh r09 N [skip_lower]'
TAG=${1:-r09}
N=${2:-2}
ONLY=${3:-0}   # 1 = bench only at N (skip the smaller world sizes)
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
echo "== multi-gpu parity check (N=$N)"; timeout 300 $RUN tests/multi_gpu_check.py > $OUT/multi_check_$N.log 2>&1; echo "rc=$?"; grep -E "multi-gpu check|MULTI_GPU_CHECK|Error|error" $OUT/multi_check_$N.log | head -20
if [ $ONLY -eq 0 ]; then
  echo "== bench N=1"; timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_1.json 2> $OUT/bench_1.err; cut -c1-300 $OUT/bench_1.json
fi
for G in 2 4 8; do
  if [ $G -le $N ] && { [ $ONLY -eq 0 ] || [ $G -eq $N ]; }; then
    for P2P in "" "--no-p2p"; do
      echo "== bench N=$G $P2P"
      ONSAS_BENCH_TWO_LEVEL_MULTI=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $G --steps 20 --warmup 3 $P2P > $OUT/bench_$G$P2P.json 2> $OUT/bench_$G$P2P.err
      echo "rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$G$P2P.json").read().strip().splitlines()[-1]); ns = d["newton_step"]
    print("N=%d value %.3f Gtets/s  newton %.1f ms  cg_iters %d  us/iter %.2f  e2e %.3f Gtets/s" % (d["n_gpus"], d["value"] / 1e9, d["newton_step_ms"], ns["cg_iters"], 1e3 * ns["ms_solve"] / ns["cg_iters"], d["e2e"]["value"] / 1e9))
    print("   two-level:", d.get("newton_step_two_level"))
except Exception as ex:
    print("no JSON:", ex)
PY
      grep -v "^\*\|OMP_NUM\|^$" $OUT/bench_$G$P2P.err | tail -4
    done
  fi
done
ls -la $OUT
