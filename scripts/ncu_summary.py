"""Writes the few ncu metrics the design cites from a .ncu-rep into a CSV under profiles/ (one block per captured launch)."""
import csv
import subprocess
import sys

rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
with open(out, "w") as f:
    f.write(f"# {note}\nkernel,metric,unit,value\n")
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        name = d.get("Kernel Name", "?").split("(")[0].replace("void ", "")
        for h, u in zip(hdr, units):
            if h in keep or ("warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(d[h] or 0) > 0.2):
                f.write(f'"{name}",{h},{u},{d[h]}\n')
print(open(out).read()[:1500])
