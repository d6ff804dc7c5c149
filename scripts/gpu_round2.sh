#!/bin/bash
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log; tail -15 $OUT/pytest.log
echo "== variants"; for MB in 1 2 3; do timeout 300 python scripts/time_variants.py $MB >> $OUT/variants.log 2>&1; done; cat $OUT/variants.log
echo "== cg variants"; timeout 600 python scripts/cg_variants.py > $OUT/cg_variants.log 2>&1; cat $OUT/cg_variants.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
echo "== ncu launch list (multi-launch CG phases)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file $OUT/launches_cgmode1.csv python scripts/profile_target.py 55 neo 1 1 1 > $OUT/ncu_cg.log 2>&1; echo "rc=$?"
echo "== ncu full: assembly"; ONSAS_ASM_MINB=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble -s 2 -c 1 -o $OUT/prof_assemble python scripts/profile_target.py 55 neo 4 0 0 > $OUT/ncu_asm.log 2>&1; echo "rc=$?"
ls -la $OUT
