"""Summarise an .ncu-rep (one row per profiled launch) into a markdown table.
usage: python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/rNN_name.md"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("lts__t_bytes.sum", "L2 bytes"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("launch__block_size", "block"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %")]
cols = [(h, n) for h, n in want if h in ci]
print("| " + " | ".join(n for _, n in cols) + " |")
print("|" + "---|" * len(cols))
for r in data:
    out = []
    for h, n in cols:
        v = r[ci[h]]
        if h == "Kernel Name":
            v = v.replace("void ", "").replace("<unnamed>::", "")[:60]
        else:
            v = f"{v} {units[ci[h]]}".strip()
        out.append(v)
    print("| " + " | ".join(out) + " |")
