"""Summarise a .ncu-rep (ncu --set full) into the handful of counters DESIGN.md / bench.py cite."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
        "sm__cycles_elapsed.max", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
for r in rows[2:]:
    print("----")
    for w, i in idx:
        print(f"{w} [{units[i]}] = {r[i]}")
