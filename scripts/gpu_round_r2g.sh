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
echo "== summaries (on the box: the .ncu-rep files are too large to travel)"
for r in assemble_neo assemble_svk assemble_truss cg_stream cg_two_level; do
  python scripts/ncu_summary.py $OUT/prof_$r.ncu-rep $OUT/${r}_ncu.md > /dev/null 2>&1
  ncu -i $OUT/prof_$r.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys,json
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    g=lambda k: r[h.index(k)] if k in h else None
    print(json.dumps({'report':'$r','kernel':g('Kernel Name'),'us':g('gpu__time_duration.sum'),'dram_read_MB':g('dram__bytes_read.sum'),'dram_write_MB':g('dram__bytes_write.sum'),'lsu_pct':g('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'),'fp64_pct':g('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'),'dram_pct':g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),'warps_pct':g('sm__warps_active.avg.pct_of_peak_sustained_active'),'regs':g('launch__registers_per_thread')}))
" >> $OUT/ncu_kernels.jsonl
done
cat $OUT/ncu_kernels.jsonl
rm -f $OUT/*.ncu-rep
ls -la $OUT
