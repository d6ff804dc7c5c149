#!/bin/bash
# N-GPU call: parity check under torchrun at N ranks, then bench at N (weak leg + configs[3] strong leg).
# usage: gpurun --gpus 8 --timeout 1200 -- 'bash scripts/gpu_round_r2d.sh r43 8'
TAG=${1:-r43}
N=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt; free -g | head -2 >> $OUT/gpu.txt
if [ "$N" -gt 1 ]; then
echo "== multi_gpu_check N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py > $OUT/multi_check_$N.log 2>&1; echo "rc=$?"; grep "multi-gpu check\|MULTI_GPU" $OUT/multi_check_$N.log
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench rc=$?"; tail -5 $OUT/bench_n$N.err
else
echo "== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$?"; tail -5 $OUT/bench_n1.err
fi
python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench_n$N.json") if l.startswith("{")][0])
keep={k:d[k] for k in ("value","ms_per_step","newton_step_ms","newton_step","newton_step_two_level","parity","full_solve","strong_c4","clocks")}
keep["e2e"]={k:d["e2e"][k] for k in ("value","ms_per_step")}
keep["roofline_frac"]=d["roofline"]["frac"]; keep["pcg_us"]=d["roofline_pcg"]["us_per_iteration"]
print(json.dumps(keep,indent=1))
PY
ls -la $OUT
