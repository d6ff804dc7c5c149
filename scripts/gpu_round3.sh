#!/bin/bash
TAG=${1:-r03}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log; tail -5 $OUT/pytest.log
echo "== variants"; for MB in 1 2 3; do timeout 300 python scripts/time_variants.py $MB >> $OUT/variants.log 2>&1; done; cat $OUT/variants.log
echo "== ncu full: assembly"; ONSAS_ASM_MINB=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble -s 2 -c 1 -o $OUT/prof_assemble python scripts/profile_target.py 55 neo 4 0 0 > $OUT/ncu_asm.log 2>&1; echo "rc=$?"
ls -la $OUT
