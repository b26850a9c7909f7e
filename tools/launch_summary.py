"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals/shares and the ordered list."""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
tot = OrderedDict()
seq = []
for r in rows:
    # columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section, Metric Name, Metric Unit, Metric Value
    name, grid, unit, val = r[4], r[8], r[-2], float(r[-1].replace(",", ""))
    ms = val / 1e6 if unit in ("ns", "nsecond") else val / 1e3 if unit in ("us", "usecond") else val
    short = name.split("(")[0].replace("g4c::", "")
    t = tot.setdefault(short, [0, 0.0])
    t[0] += 1
    t[1] += ms
    seq.append((short, grid, ms))
total = sum(v[1] for v in tot.values())
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:45s} n={n:4d} total={ms:10.3f} ms share={ms / total:.3f}")
print(f"# total {total:.3f} ms over {len(seq)} launches")
if len(sys.argv) > 2:
    print("# launches in order (kernel, grid, ms)")
    for s in seq[: int(sys.argv[2])]:
        print(f"{s[0]:45s} {s[1]:>18s} {s[2]:9.3f}")
