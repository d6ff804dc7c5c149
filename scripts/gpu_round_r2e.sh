#!/bin/bash
# 1-GPU call: assembly variants after the phase-B batching, truss register budgets on the 10 M-bar lattice, configs[2] on 1 GPU,
# GPU tests.  usage: gpurun --timeout 1500 -- 'bash scripts/gpu_round_r2e.sh r46'
TAG=${1:-r46}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== variants"; timeout 200 python scripts/time_variants.py 3 55 > $OUT/time_variants.log 2>&1; cat $OUT/time_variants.log
echo "== trusses"; for m in 2 3 4; do ONSAS_TRUSS_MINB=$m timeout 300 python scripts/config_sweep.py c5 2>&1 | cut -c1-330; done | tee $OUT/c5_minb.jsonl
echo "== c3 1 GPU"; timeout 400 python scripts/config_multi.py c3 > $OUT/c3_n1.json 2> $OUT/c3_n1.err; cat $OUT/c3_n1.json; tail -3 $OUT/c3_n1.err
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log; tail -5 $OUT/pytest.log
