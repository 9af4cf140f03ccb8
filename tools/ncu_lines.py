#!/usr/bin/env python
"""Per-source-line share of warp-stall samples and executed instructions of one kernel in an .ncu-rep
(needs -lineinfo and --import-source on).  usage: python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [min pct]"""
import csv
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{rx}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = next(r for r in rows if r and r[0] == "Line No")
    ix = {}
    for i, n in enumerate(hdr):
        ix.setdefault(n, i)

    def f(x):
        try:
            return float(x.replace(",", ""))
        except ValueError:
            return 0.0
    path = next(r[1] for r in rows if r and r[0] == "File Path")
    src = open(path).read().split("\n")
    lines = [r for r in rows if len(r) == len(hdr) and r[0].isdigit()]
    S, I = ix["Warp Stall Sampling (All Samples)"], ix["Instructions Executed"]
    ts, ti = sum(f(r[S]) for r in lines), sum(f(r[I]) for r in lines)
    print(f"samples {ts:.0f}  warp instructions {ti:.0f}")
    for r in lines:
        s, i = f(r[S]) / ts * 100, f(r[I]) / ti * 100
        if s >= thr or i >= thr:
            print(f"{r[0]:>5} {s:5.1f}% smp {i:5.1f}% inst | {src[int(r[0]) - 1].strip()[:110]}")


if __name__ == "__main__":
    main()
