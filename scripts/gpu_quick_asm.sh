#!/bin/bash
# quick A/B call for assembly-kernel changes: the bench's assembly / e2e legs and the parity + determinism tests of the assembly
TAG=${1:-r78}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-c4 --no-full-solve > $OUT/bench_quick.json 2> $OUT/bench_quick.err
python -c "
import json; d=json.loads(open('$OUT/bench_quick.json').read().strip().splitlines()[-1]); print('asm ms', d['ms_per_step'], 'Gtets/s', d['value']/1e9, 'frac', d['roofline']['frac'], 'e2e ms', d['e2e']['ms_per_step'], 'newton', d['newton_step_ms'], d['newton_step_two_level']['ms'])"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
