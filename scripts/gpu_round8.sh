#!/bin/bash
# assembly kernel iteration: tests + timing of the register variants x materials
TAG=${1:-r08}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log; tail -5 $OUT/pytest.log
echo "== variants"; for MB in 1 2 3; do timeout 300 python scripts/time_variants.py $MB >> $OUT/variants.log 2>&1; done; cat $OUT/variants.log
