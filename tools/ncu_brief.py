"""Print a short list of metrics from an .ncu-rep (ncu -i REP --page raw --csv piped through this filter).

    python tools/ncu_brief.py gpurun_out/x.ncu-rep [substring ...]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct",
        "dram__throughput.avg.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_issue_stalled", "l1tex__data_bank_conflicts",
        "sm__warps_active.avg.pct_of_peak_sustained_active"] + extra
for r in rows[2:]:
    print("KERNEL", r[hdr.index("Kernel Name")][:70])
    for i, h in enumerate(hdr):
        if any(k in h for k in want):
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if "stalled" in h and v < 0.05:
                continue
            print(f"   {h:110s} {units[i]:12s} {r[i]}")
