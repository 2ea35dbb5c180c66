"""Splits the SASS-level `ncu --page source --csv` export of ONE kernel at its BAR.SYNC instructions and prints, per
segment (= phase between two __syncthreads()), the share of warp-stall samples and of executed instructions, plus the
instructions with the most samples. Usage: ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > k.csv;
python tools/ncu_source_segments.py k.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[h]
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
tot = sum(int(r[iS]) for r in data)
tex = sum(int(r[iEx]) for r in data)
print("total samples %d, %d SASS instructions, %d warp instructions executed" % (tot, len(data), tex))
acc = ninst = exe = start = seg = 0
for k, r in enumerate(data):
    acc += int(r[iS]); ninst += 1; exe += int(r[iEx])
    if "BAR.SYNC" in r[iSrc] or k == len(data) - 1:
        print("seg %2d sass[%5d..%5d] samples %6d (%5.1f%%)  executed %9d (%5.1f%%)  static %5d" %
              (seg, start, k, acc, 100.0 * acc / max(tot, 1), exe, 100.0 * exe / max(tex, 1), ninst))
        seg += 1; acc = ninst = exe = 0; start = k + 1
for r in sorted(data, key=lambda r: -int(r[iS]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print("%6s samples  %8s exec  %s" % (r[iS], r[iEx], r[iSrc][:100]))
