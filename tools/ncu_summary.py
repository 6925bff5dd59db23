#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one row per kernel launch with the metrics the roofline lines quote.
usage: python tools/ncu_summary.py gpurun_out/r1_raw.csv [--md]"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "us", 1.0), ("dram__bytes_read.sum", "rd MB", 1.0), ("dram__bytes_write.sum", "wr MB", 1.0),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %", 1.0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %", 1.0),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1 %", 1.0), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 %", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1.0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %", 1.0), ("launch__registers_per_thread", "regs", 1.0),
        ("launch__grid_size", "grid", 1.0)]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(m, t) for m, t, _ in COLS if m in idx]
    unit = {m: units[idx[m]] for m, _ in cols}
    md = "--md" in sys.argv
    head = ["kernel"] + [t for _, t in cols]
    print(("| " + " | ".join(head) + " |") if md else "\t".join(head))
    if md:
        print("|" + "---|" * len(head))
    for d in data:
        name = d[idx["Kernel Name"]].split("(")[0]
        name = name.replace("srlz::", "")
        vals = []
        for m, _ in cols:
            v = d[idx[m]].replace(",", "")
            try:
                f = float(v)
                u = unit[m].lower()
                if u == "byte":
                    f /= 1e6
                elif u == "kbyte":
                    f /= 1e3
                elif u == "gbyte":
                    f *= 1e3
                elif u in ("ns", "nsecond"):
                    f /= 1e3
                elif u in ("ms", "msecond"):
                    f *= 1e3
                vals.append("%.1f" % f if abs(f) < 1e6 else "%.0f" % f)
            except ValueError:
                vals.append(v)
        print(("| `" + name + "` | " + " | ".join(vals) + " |") if md else name + "\t" + "\t".join(vals))


if __name__ == "__main__":
    main()
