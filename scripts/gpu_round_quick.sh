#!/bin/bash
# A short gpurun call: GPU tests, bench, the streamed-CG probe (two-level set-up cost) and the e2e chunk sweep.
# usage: gpurun --timeout 900 -- 'bash scripts/gpu_round_quick.sh r14'
TAG=${1:-r14}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log; tail -15 $OUT/pytest.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
echo "== cg probe"; ONSAS_PROF_VERBOSE=1 ONSAS_VERBOSE=1 timeout 300 python scripts/cg_stream_probe.py > $OUT/cg_stream_probe.log 2>&1; grep "prof=\|two-level\|onsas prof" $OUT/cg_stream_probe.log
echo "== e2e sweep"; timeout 300 python scripts/e2e_sweep.py > $OUT/e2e_sweep.log 2>&1; cat $OUT/e2e_sweep.log
ls -la $OUT
