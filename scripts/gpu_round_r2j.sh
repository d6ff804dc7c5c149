#!/bin/bash
# coarse set-up kernels of the two-level preconditioner: ncu timings of both forms of k_coarse_assemble and of the blocked inverse,
# the cross-check of E, the Newton step with each.  usage: gpurun --timeout 400 -- 'bash scripts/gpu_round_r2j.sh r75'
TAG=${1:-r75}
OUT=gpurun_out/$TAG
mkdir -p $OUT
M=gpu__time_duration.sum,smsp__cycles_active.avg,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__inst_executed.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:"k_coarse_assemble|k_gj" -c 2 --csv --log-file $OUT/ncu_coarse_new.csv python scripts/profile_target.py 55 neo 0 0 1 2 > $OUT/ncu_coarse_new.log 2>&1
grep -v "^==" $OUT/ncu_coarse_new.csv | cut -d, -f5,13-
ONSAS_COARSE_SERIAL=1 timeout 300 ncu --metrics $M --clock-control none -k regex:"k_coarse_assemble|k_gj" -c 2 --csv --log-file $OUT/ncu_coarse_serial.csv python scripts/profile_target.py 55 neo 0 0 1 2 > $OUT/ncu_coarse_serial.log 2>&1
grep -v "^==" $OUT/ncu_coarse_serial.csv | cut -d, -f5,13- | grep "coarse"
(ONSAS_COARSE_CHECK=1 timeout 120 python scripts/l2pf_sweep.py 55 0 2>&1 | grep "coarse operator" | sort | uniq -c | sort -rn | head -3
echo "-- new"; timeout 120 python scripts/l2pf_sweep.py 55 0 2>&1 | grep "two_level"
echo "-- serial"; ONSAS_COARSE_SERIAL=1 timeout 120 python scripts/l2pf_sweep.py 55 0 2>&1 | grep "two_level") | tee $OUT/coarse_assemble2.log
