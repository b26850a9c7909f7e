"""Top stall-sample SASS instructions of an .ncu-rep (source page), with a little context."""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = rows[hdr_i + 1:]
si = hdr.index("# Samples") if "# Samples" in hdr else hdr.index("Warp Stall Sampling (All Samples)")
ei = hdr.index("Instructions Executed")
total = sum(int(r[si] or 0) for r in body)
top = sorted(range(len(body)), key=lambda i: -int(body[i][si] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]
print(f"# total samples {total}")
for i in sorted(top):
    r = body[i]
    print(f"{i:6d} {int(r[si]):7d} {100.0 * int(r[si]) / total:5.1f}%  exec={r[ei]:>10s}  {r[1].strip()[:110]}")
