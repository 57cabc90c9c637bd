"""Pivot an `ncu --metrics ... --csv` log (one row per metric per launch) into one line per launch."""
import csv, io, re, sys
for path in sys.argv[1:]:
    lines = [l for l in open(path).read().splitlines() if l.startswith('"')]
    rows = list(csv.reader(io.StringIO("\n".join(lines))))
    hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
    out = {}
    for r in rows[1:]:
        k = (int(r[ci["ID"]]), re.sub(r"\(.*", "", r[ci["Kernel Name"]]).replace("void ", "")[-40:])
        out.setdefault(k, {})[r[ci["Metric Name"]]] = (float(r[ci["Metric Value"]].replace(",", "")), r[ci["Metric Unit"]])
    print("==", path)
    for (i, name), m in sorted(out.items()):
        t, u = m["gpu__time_duration.sum"]
        t_us = t / 1e3 if u in ("ns", "nsecond") else t * 1e3 if u in ("ms", "msecond") else t
        def gb(key):
            v, u = m[key]
            return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1e-9)
        print(f"{name:42s} {t_us:10.1f} us  tensor {m['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'][0]:5.1f}%"
              f"  dram rd {gb('dram__bytes_read.sum'):7.2f} GB  wr {gb('dram__bytes_write.sum'):7.2f} GB")
