"""Top stall sites of a kernel from `ncu -i X.ncu-rep --page source --csv` (SASS view)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
k = ci["Warp Stall Sampling (All Samples)"]
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot = sum(int(r[k] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
for idx, r in enumerate(body):
    r.append(idx)
top = sorted(body, key=lambda r: -int(r[k] or 0))[:n]
for r in top:
    print("%6s %5.1f%%  #%-5d exec=%-9s %s" % (r[k], 100.0 * int(r[k] or 0) / max(tot, 1), r[-1], r[ci["Instructions Executed"]], r[ci["Source"]].strip()[:90]))
