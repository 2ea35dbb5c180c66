"""Condenses an `ncu --csv --page raw` log (one row per profiled launch, one column per metric) to the handful of columns
that decide what bounds a memory-bound kernel. Usage: python tools/ncu_raw_summary.py file.csv > summary.csv"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("smsp__inst_executed.sum", "inst"),
]
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = list(csv.reader(lines))
hdr, units, rows = rd[0], rd[1], rd[2:]
idx = {h: i for i, h in enumerate(hdr)}
print(",".join(["kernel"] + ["%s[%s]" % (short, units[idx[name]]) if name in idx else short for name, short in COLS]))
for r in rows:
    name = r[idx["Kernel Name"]].split("(")[0]
    print(",".join([name[:48]] + [r[idx[n]] if n in idx else "" for n, _ in COLS]))
