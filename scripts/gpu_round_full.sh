#!/bin/bash
# One gpurun call: tests, bench (ours + reference arm), launch list, full ncu captures (assembly, streamed CG with the
# Jacobi and the two-level preconditioner), optional config sweep.
# usage (from the repo root): gpurun --timeout 1500 -- 'bash scripts/gpu_round_full.sh r13 [sweep]'
TAG=${1:-r25}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -2 $OUT/smoke.log
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log; tail -5 $OUT/pytest.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2>> $OUT/bench.err; cat $OUT/bench_ref.json
echo "== cg probe"; ONSAS_PROF_VERBOSE=1 ONSAS_VERBOSE=1 timeout 300 python scripts/cg_stream_probe.py > $OUT/cg_stream_probe.log 2>&1; grep "prof=\|two-level" $OUT/cg_stream_probe.log
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full: streamed CG, two-level"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cg_stream|k_gj_invert|k_coarse_assemble" -c 3 -o $OUT/prof_cg_two_level python scripts/profile_target.py 55 neo 0 0 1 2 > $OUT/ncu_cg2.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu_cg2.log
echo "== ncu full: streamed CG, Jacobi"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:cg_stream -c 1 -o $OUT/prof_cg_stream python scripts/profile_target.py 55 neo 0 0 1 1 > $OUT/ncu_cg.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu_cg.log
echo "== ncu full: assembly"; ONSAS_ASM_MINB=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble -s 2 -c 1 -o $OUT/prof_assemble python scripts/profile_target.py 55 neo 4 0 0 > $OUT/ncu_asm.log 2>&1; echo "rc=$?"
if [ "$2" = "sweep" ]; then
echo "== config sweep"; timeout 1200 python scripts/config_sweep.py c2 c3 c5 > $OUT/config_sweep.jsonl 2> $OUT/config_sweep.err; echo "rc=$?"; cut -c1-400 $OUT/config_sweep.jsonl; tail -3 $OUT/config_sweep.err
fi
ls -la $OUT
