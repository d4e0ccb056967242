"""Extract the judged metrics of an ncu report into CSV: python profiles/summarize.py <rep> > out.csv"""
import csv, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
keys = ("Kernel Name", "Block Size", "Grid Size", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "launch__shared_mem_per_block_dynamic")
keep = [i for i, h in enumerate(hdr) if h in keys or "warp_issue_stalled" in h and h.endswith("per_warp_active.pct")]
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + ["launch %d" % i for i in range(len(rows) - 2)])
for i in keep:
    w.writerow([hdr[i], units[i]] + [r[i] for r in rows[2:]])
