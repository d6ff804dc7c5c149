#!/bin/bash
# GPU tests only (any N): usage: gpurun --gpus 2 --timeout 1200 -- 'bash scripts/gpu_round_r2c.sh r42 [pytest args]'
TAG=${1:-r42}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -s "$@" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log; grep "multi-gpu check\|full solve\|^FAILED\|^ERROR" $OUT/pytest.log; tail -30 $OUT/pytest.log
