#!/bin/bash
# L2 prefetch of the streamed CG: sweep (timing, bitwise check) and ncu counters (DRAM bytes, L2 hit rate): are the prefetched lines used?
TAG=${1:-r65}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python scripts/l2pf_sweep.py 55 0 2 4 6 8 12 0 > $OUT/l2pf_sweep_hint${ONSAS_STREAM_L2_PF_HINT:-1}.log 2>&1; cat $OUT/l2pf_sweep_hint${ONSAS_STREAM_L2_PF_HINT:-1}.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
for pf in ${PFS:-0 4}; do
  ONSAS_STREAM_L2_PREFETCH=$pf timeout 300 ncu --metrics $M --clock-control none -k regex:cg_stream -c 1 --csv --log-file $OUT/ncu_pf$pf.csv python scripts/profile_target.py 55 neo 1 0 1 1 > $OUT/ncu_pf$pf.log 2>&1
  echo "pf=$pf rc=$?"; grep -v "^==" $OUT/ncu_pf$pf.csv | cut -d, -f13- | tail -7
done
