#!/bin/bash
# Weak-scaling bench lines on one box: usage: gpurun --gpus 4 --timeout 600 -- 'bash scripts/gpu_round_scale.sh r26 "2 4"'
TAG=${1:-r26}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for G in ${2:-2 4}; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $G --steps 20 --warmup 3 > $OUT/bench_$G.json 2> $OUT/bench_$G.err
  echo "N=$G rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$G.json").read().strip().splitlines()[-1]); ns = d["newton_step"]; t = d.get("newton_step_two_level") or {}
    print("N=%d value %.3f Gtets/s  newton %.1f ms  cg_iters %d  us/iter %.2f  e2e %.3f Gtets/s | two-level %.1f ms %s its" % (d["n_gpus"], d["value"] / 1e9, d["newton_step_ms"], ns["cg_iters"], 1e3 * ns["ms_solve"] / ns["cg_iters"], d["e2e"]["value"] / 1e9, t.get("ms", float("nan")), t.get("cg_iters")))
except Exception as ex:
    print("no JSON:", ex)
PY
  grep -v "^\*\|OMP_NUM\|^$" $OUT/bench_$G.err | tail -3
done
