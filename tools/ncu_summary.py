"""Print the judged metrics of an .ncu-rep (every profiled launch): duration, tensor / XU pipe, DRAM bytes, stalls."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
keys = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for n, vals in enumerate(rows[2:]):
    d = dict(zip(hdr, vals))
    print(f"==== launch {n}")
    for k in keys:
        for h in hdr:
            if h == k or h.endswith("." + k):
                print(f"{k} = {d[h]}")
                break
    print("-- stall reasons (warps per issue-active cycle)")
    st = [(float(v), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")) for h, v in d.items()
          if "smsp__average_warps_issue_stalled_" in h and h.endswith("_per_issue_active.ratio") and v not in ("", "n/a")]
    for v, h in sorted(st, reverse=True)[:8]:
        print(f"  {h:28s} {v:.3f}")
