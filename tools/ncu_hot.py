#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel launch of an ncu report (source page, needs -lineinfo + --import-source).
usage: python tools/ncu_hot.py report.ncu-rep [launch_index] [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else "0"
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = heads[int(kid)]
end = heads[int(kid) + 1] - 1 if int(kid) + 1 < len(heads) else len(rows)
hdr = rows[hi]
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
samp = "# Samples"
tot = sum(int(r[ci[samp]]) for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ci[h]]) for r in data) for h in stalls}
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True))
for idx, r in sorted(enumerate(data), key=lambda t: -int(t[1][ci[samp]]))[:topn]:
    s = int(r[ci[samp]])
    st = sorted([(int(r[ci[h]]), h[6:]) for h in stalls], reverse=True)[:2]
    print(f"{idx:5d} {s:6d} {s / tot * 100:5.1f}%  {r[ci['Source']].strip()[:70]:70s} {st}")
