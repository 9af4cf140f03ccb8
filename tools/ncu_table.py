"""Compact table of an ncu --set full report: one row per captured launch with the metrics the roofline discussion uses.
usage: python tools/ncu_table.py <report.ncu-rep> <out.md> "<title>" """
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum"]
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
ki = hdr.index("Kernel Name")
with open(out, "w") as f:
    f.write(f"# {title}\n\n(`ncu --set full --clock-control none`, report {rep}; one row per captured launch)\n\n")
    f.write("| kernel | " + " | ".join(w.split(".")[0].replace("__", ".") for w, _ in idx) + " |\n|---|" + "---|" * len(idx) + "\n")
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("cofi::", "")
        f.write("| `" + name[:52] + "` | " + " | ".join((r[i][:9] + " " + units[i]).strip() for _, i in idx) + " |\n")
print(open(out).read()[:3000])
