"""Condense an `ncu -i report.ncu-rep --page raw --csv` export into the one-row-per-kernel table kept under profiles/
(same columns as profiles/r2_ncu_full_summary.csv).  usage: python tools/ncu_raw_extract.py raw.csv > summary.csv"""
import csv
import sys

COLS = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic"]
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
idx = [hdr.index(c) if c in hdr else None for c in COLS]
out = csv.writer(sys.stdout)
for r in rows[:2] + rows[2:]:
    out.writerow([(r[i] if i is not None and i < len(r) else "") for i in idx])
