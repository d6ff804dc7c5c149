#!/bin/bash
# fused CG: tests + sweep
TAG=${1:-r06}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log; tail -5 $OUT/pytest.log
echo "== sweep"; timeout 600 python scripts/cg_sweep.py 55 > $OUT/cg_sweep.log 2>&1; echo "rc=$?"; cat $OUT/cg_sweep.log
ls -la $OUT
