"""Turns an .ncu-rep (ncu --set full --import-source on) into the small text summary committed under profiles/.
usage: python scripts/ncu_summary.py gpurun_out/r03/prof_assemble.ncu-rep profiles/r03/assembly_ncu.md"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg", "smsp__average_warp_latency_per_inst_issued.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
lines = [f"# ncu summary of `{rep}`", "", "Captured with `ncu --set full --clock-control none --import-source on` on a B200 (one launch, cold caches).", ""]
for r in rows[2:]:
    lines += [f"## {r[hdr.index('Kernel Name')]}", "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in hdr:
            lines.append(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
    lines += ["", "Warp stall reasons (warps stalled per issued instruction):", "", "| reason | value |", "|---|---|"]
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            v = float(r[i] or 0)
            if v >= 0.05:
                lines.append(f"| {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} | {v:.2f} |")
    lines.append("")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    lines += ["Source page (first kernel of the report):", ""]
    hdr = rows[1]
    iS, iSrc, iEx = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
    data = []
    for r in rows[2:]:
        if len(r) != len(hdr) or r[0] == "Address":
            break
        data.append(r)
    tot = sum(int(r[iS]) for r in data)
    bars = [i for i, r in enumerate(data) if "BAR.SYNC" in r[iSrc]]
    lines += [f"SASS instructions: {len(data)}, stall samples: {tot}, executed warp instructions: {sum(int(r[iEx]) for r in data)}", ""]
    if bars:
        b = bars[0]
        lines.append(f"Before the first barrier (phase A): {sum(int(r[iS]) for r in data[:b])} samples, "
                     f"{sum(int(r[iEx]) for r in data[:b])} warp instructions; after it (phase B): "
                     f"{sum(int(r[iS]) for r in data[b:])} samples, {sum(int(r[iEx]) for r in data[b:])} warp instructions.")
        lines.append("")
    lines += ["Hottest instructions by stall samples:", "", "| # | samples | executed | SASS |", "|---|---|---|---|"]
    for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:20]):
        lines.append(f"| {i} | {data[i][iS]} | {data[i][iEx]} | `{data[i][iSrc].strip()[:90]}` |")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
