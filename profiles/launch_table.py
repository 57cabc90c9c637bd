"""ncu launch list (csv from `--metrics gpu__time_duration.sum --csv`) -> markdown table by kernel.
usage: python profiles/launch_table.py gpurun_out/r1_launches_align.csv align > profiles/r1_launches_align.md"""
import collections
import csv
import re
import sys

path, name = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ci = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows:
    if r is hdr or r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    k = re.sub(r"\(.*$", "", r[ci["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
    k = k.replace("rz::gemm::", "rz::")
    v = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
n = sum(a[0] for a in agg.values())
print(f"# ncu launch list, `bench.py` workload `{name}` (--metrics gpu__time_duration.sum --clock-control none)\n")
print(f"Per-launch times are serialised and cold-cache: the SHARE of the step is what counts. "
      f"Total {tot / 1e3:.2f} ms over {n} launches.\n")
print("| launches | total us | share | avg us | kernel |\n|---|---|---|---|---|")
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:24]:
    print(f"| {c} | {t:.1f} | {100 * t / tot:.1f}% | {t / c:.2f} | `{k[:90]}` |")
