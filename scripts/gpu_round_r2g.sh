#!/bin/bash
# 1-GPU evidence call: GPU tests, launch list of the bench, ncu --set full of the assembly kernels (NeoHookean, SVK, truss) and of
# the streamed CG (single-reduction Jacobi, two-level).  usage: gpurun --timeout 1800 -- 'bash scripts/gpu_round_r2g.sh r49'
TAG=${1:-r49}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log; tail -4 $OUT/pytest.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c4 --no-full-solve > $OUT/ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full: assembly neo"; ONSAS_ASM_MINB=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble -s 2 -c 1 -o $OUT/prof_assemble_neo python scripts/profile_target.py 55 neo 4 0 0 > $OUT/ncu_asm.log 2>&1; echo "rc=$?"
echo "== ncu full: assembly svk"; ONSAS_ASM_MINB=3 timeout 600 ncu --set full --clock-control none -k regex:k_assemble -s 2 -c 1 -o $OUT/prof_assemble_svk python scripts/profile_target.py 55 svk 4 0 0 >> $OUT/ncu_asm.log 2>&1; echo "rc=$?"
echo "== ncu full: truss"; timeout 600 ncu --set full --clock-control none -k regex:k_assemble -s 2 -c 1 -o $OUT/prof_assemble_truss python scripts/config_sweep.py c5 > $OUT/ncu_truss.log 2>&1; echo "rc=$?"
echo "== ncu full: streamed CG, Jacobi single-reduction"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:cg_stream -c 1 -o $OUT/prof_cg_stream python scripts/profile_target.py 55 neo 0 0 1 1 > $OUT/ncu_cg.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu_cg.log
echo "== ncu full: streamed CG, two-level"; timeout 600 ncu --set full --clock-control none -k regex:"cg_stream|k_gj_invert|k_coarse_assemble" -c 3 -o $OUT/prof_cg_two_level python scripts/profile_target.py 55 neo 0 0 1 2 > $OUT/ncu_cg2.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu_cg2.log
ls -la $OUT
