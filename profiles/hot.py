"""Print the hottest instructions of an `ncu --page source --csv` dump (stall samples).

usage: python profiles/hot.py dump.csv [top=40] [section=0]
The dump holds one section per (kernel launch, view); sections are listed when `section` is -1.
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if want < 0:
    for k, i in enumerate(starts):
        print(k, rows[i][1][:100], "first col:", rows[i + 2][1][:60] if i + 2 < len(rows) else "")
    sys.exit(0)
lo = starts[want]
hi = starts[want + 1] if want + 1 < len(starts) else len(rows)
print(rows[lo][1][:120])
hdr = rows[lo + 1]
ci = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[lo + 2:hi]:
    try:
        data.append((int(r[ci["# Samples"]]), int(r[ci["Instructions Executed"]]), r[ci["Source"]].strip(), r))
    except Exception:
        pass
tot = sum(d[0] for d in data) or 1
print("total samples", tot, "instructions", len(data))
idx = sorted(range(len(data)), key=lambda i: -data[i][0])[:top]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for i in sorted(idx):
    s, n, src, r = data[i]
    st = sorted(((int(r[ci[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {s:6d} ({100*s/tot:4.1f}%) exec={n:8d} {src[:80]:80s} {st}")
