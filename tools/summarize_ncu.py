"""Text summary of an .ncu-rep (ncu --set full): one block per captured launch with the metrics the profiles/ files
quote.  Usage: summarize_ncu.py <report.ncu-rep> [header line ...]   (needs `ncu` on PATH; runs without a GPU)"""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")]
for line in sys.argv[2:]:
    print("# " + line)
print(f"# source: {rep} (ncu --set full --clock-control none), one block per captured launch\n")
for d in data:
    for w in WANT:
        if w in idx:
            v = d[idx[w]][:100] if w == "Kernel Name" else d[idx[w]]
            print(f"{w:90s} {v} {units[idx[w]]}")
    top = sorted(((float(d[idx[h]].replace(",", "")), h) for h in stalls if d[idx[h]] not in ("", "n/a")), reverse=True)[:5]
    for x, h in top:
        print(f"{h:90s} {x:.3f} inst")
    print()
