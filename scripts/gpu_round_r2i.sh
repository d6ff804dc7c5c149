#!/bin/bash
# 1-GPU call after the two-level solver changes: full bench (all legs), ncu --set full of the two-level streamed CG + its coarse set-up kernels.
# usage: gpurun --timeout 900 -- 'bash scripts/gpu_round_r2i.sh r74'
TAG=${1:-r74}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$?"; tail -3 $OUT/bench_n1.err
python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench_n1.json") if l.startswith("{")][0])
keep={k:d.get(k) for k in ("value","ms_per_step","newton_step_ms","newton_step","newton_step_two_level","parity","full_solve","clocks")}
keep["e2e"]={k:d["e2e"][k] for k in ("value","ms_per_step")}
keep["roofline_frac"]=d["roofline"]["frac"]; keep["pcg_us"]=d["roofline_pcg"]["us_per_iteration"]
c4=d.get("strong_c4") or {}
keep["c4"]={k:c4.get(k) for k in ("assembly_ms","newton_step_jacobi","newton_step_two_level","error")}
print(json.dumps(keep,indent=1)[:5000])
PY
echo "== ncu full: streamed CG, two-level"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cg_stream|k_gj_invert|k_coarse_assemble" -c 3 -o $OUT/prof_cg_two_level python scripts/profile_target.py 55 neo 0 0 1 2 > $OUT/ncu_cg2.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu_cg2.log
python scripts/ncu_summary.py $OUT/prof_cg_two_level.ncu-rep $OUT/cg_two_level_ncu.md > /dev/null 2>&1
ncu -i $OUT/prof_cg_two_level.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys,json
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    g=lambda k: r[h.index(k)] if k in h else None
    print(json.dumps({'report':'cg_two_level','kernel':g('Kernel Name'),'us':g('gpu__time_duration.sum'),'dram_read_MB':g('dram__bytes_read.sum'),'dram_write_MB':g('dram__bytes_write.sum'),'lsu_pct':g('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'),'dram_pct':g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),'lts_pct':g('lts__throughput.avg.pct_of_peak_sustained_elapsed'),'warps_pct':g('sm__warps_active.avg.pct_of_peak_sustained_active'),'regs':g('launch__registers_per_thread')}))
" > $OUT/ncu_kernels.jsonl
cat $OUT/ncu_kernels.jsonl
rm -f $OUT/*.ncu-rep
ls -la $OUT
