#!/bin/bash
# Round-2 quick call: GPU tests, bench, assembly register variants, ncu full of the assembly kernel.
# usage: gpurun --timeout 1200 -- 'bash scripts/gpu_round_r2a.sh r40'
TAG=${1:-r40}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -s > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log; grep "full solve" $OUT/pytest.log; tail -8 $OUT/pytest.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-1500 $OUT/bench.json; tail -3 $OUT/bench.err
echo "== variants"; for m in 1 2 3; do timeout 200 python scripts/time_variants.py $m 55; done > $OUT/time_variants.log 2>&1; cat $OUT/time_variants.log
echo "== ncu full: assembly"; ONSAS_ASM_MINB=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble -s 2 -c 1 -o $OUT/prof_assemble python scripts/profile_target.py 55 neo 4 0 0 > $OUT/ncu_asm.log 2>&1; echo "rc=$?"
ls -la $OUT
