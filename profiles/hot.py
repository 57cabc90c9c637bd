"""Print the hottest SASS instructions of an `ncu --page source --csv` dump (stall samples)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr) - 2 or r[0] == "Address" or r[0] == "Kernel Name":
        continue
    try:
        data.append((int(r[ci["# Samples"]]), int(r[ci["Instructions Executed"]]), r[ci["Source"]].strip(), r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
idx = sorted(range(len(data)), key=lambda i: -data[i][0])[:top]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for i in sorted(idx):
    s, n, src, r = data[i]
    st = sorted(((int(r[ci[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {s:6d} ({100*s/tot:4.1f}%) exec={n:8d} {src[:70]:70s} {st}")
