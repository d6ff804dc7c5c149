#!/bin/bash
# N-GPU measurement campaign: bench at N, cg phase profile at N, configs[2] / configs[4] at N.
# usage: gpurun --gpus 8 --timeout 1500 -- 'bash scripts/gpu_round_r2f.sh r47 8 "bench prof c5"'
TAG=${1:-r47}; N=${2:-8}; WHAT=${3:-"bench prof"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for w in $WHAT; do
case $w in
bench) echo "== bench N=$N"; if [ "$N" -gt 1 ]; then timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; else timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; fi; echo "rc=$?"; tail -3 $OUT/bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench_n$N.json") if l.startswith("{")][0])
keep={k:d.get(k) for k in ("value","ms_per_step","newton_step_ms","newton_step","newton_step_two_level","parity","full_solve","strong_c4","clocks","cpu_baseline")}
keep["e2e"]={k:d["e2e"][k] for k in ("value","ms_per_step")}
keep["roofline_frac"]=d["roofline"]["frac"]; keep["pcg_us"]=d["roofline_pcg"]["us_per_iteration"]
print(json.dumps(keep,indent=1)[:6000])
PY
;;
prof) echo "== cg profile N=$N"; if [ "$N" -gt 1 ]; then timeout 600 $TR --master-port 29513 scripts/cg_profile_multi.py > $OUT/cg_profile_n$N.jsonl 2> $OUT/cg_profile_n$N.err; else timeout 600 python scripts/cg_profile_multi.py > $OUT/cg_profile_n$N.jsonl 2> $OUT/cg_profile_n$N.err; fi; echo "rc=$?"; grep "^{" $OUT/cg_profile_n$N.jsonl | cut -c1-1500; tail -3 $OUT/cg_profile_n$N.err;;
c3|c5) echo "== $w N=$N"; if [ "$N" -gt 1 ]; then timeout 900 $TR --master-port 29512 scripts/config_multi.py $w > $OUT/${w}_n$N.json 2> $OUT/${w}_n$N.err; else timeout 900 python scripts/config_multi.py $w > $OUT/${w}_n$N.json 2> $OUT/${w}_n$N.err; fi; echo "rc=$?"; grep "^{" $OUT/${w}_n$N.json | cut -c1-3000; tail -3 $OUT/${w}_n$N.err;;
esac
done
ls -la $OUT
