#!/bin/bash
# One gpurun call: tests, bench, launch list, full ncu captures of the two dominant kernels, sanitizer.
# usage (from the repo root): gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh r01'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
python -c "import torch; print(torch.cuda.get_device_name(0))" >> $OUT/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log; tail -5 $OUT/pytest.log
echo "== pytest (continue past first failure)"; grep -q " failed" $OUT/pytest.log && (timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_all.log 2>&1; tail -30 $OUT/pytest_all.log)
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2>> $OUT/bench.err; cat $OUT/bench_ref.json
for MB in 1 3; do echo "== variant minb=$MB"; ONSAS_ASM_MINB=$MB timeout 300 python scripts/time_variants.py $MB >> $OUT/variants.log 2>&1; done; cat $OUT/variants.log
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full: assembly"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble -s 2 -c 2 -o $OUT/prof_assemble python scripts/profile_target.py 55 neo 4 0 0 > $OUT/ncu_asm.log 2>&1; echo "rc=$?"
echo "== ncu full: spmv"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_dot -s 1 -c 2 -o $OUT/prof_spmv python scripts/profile_target.py 55 neo 1 3 0 > $OUT/ncu_spmv.log 2>&1; echo "rc=$?"
echo "== sanitizer"; timeout 600 compute-sanitizer --tool memcheck python scripts/profile_target.py 6 neo 2 2 1 > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/memcheck.log
timeout 600 compute-sanitizer --tool racecheck python scripts/profile_target.py 6 svk 2 0 0 > $OUT/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 $OUT/racecheck.log
ls -la $OUT
