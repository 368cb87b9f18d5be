"""Prints the metrics quoted in DESIGN.md / profiles/*_summary.txt from an ncu report (ncu -i REP --page raw --csv).
   python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep"""
import csv, io, subprocess, sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, val = rows[0], rows[1], rows[2]
print("Kernel Name =", val[hdr.index("Kernel Name")])
for i, h in enumerate(hdr):
    stall = "issue_stalled" in h and h.endswith("per_issue_active.ratio")
    if h in WANT or (stall and float(val[i] or 0) >= 0.05):
        print(f"{h} = {val[i]} {units[i]}")
