# usage: ncu_digest.sh <report.ncu-rep> <out-prefix> [phase specs for ncu_lines.py ...]
# Turns one ncu report into small text files (raw metrics, stall / opcode summary, hottest source lines) and deletes it.
rep=$1; out=$2; shift 2
ncu -i $rep --page raw --csv > ${out}_raw.csv 2>/dev/null
ncu -i $rep --page source --csv > /tmp/_src.csv 2>/dev/null
python scripts/ncu_summary.py /tmp/_src.csv > ${out}_summary.txt 2>&1
python scripts/ncu_lines.py $rep 40 "$@" >> ${out}_summary.txt 2>&1
python - "$out" <<'PY' >> ${out}_summary.txt
import csv, sys
rows = list(csv.reader(open(sys.argv[1] + "_raw.csv")))
hdr, units, d = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_per_inst_issued.ratio", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
print("key metrics (ncu --set full, --clock-control none):")
for i, h in enumerate(hdr):
    if h in want:
        print(f"  {h} [{units[i]}] = {d[i]}")
PY
rm -f $rep /tmp/_src.csv
