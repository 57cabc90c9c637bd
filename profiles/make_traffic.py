"""profiles/traffic.json <- DRAM bytes (read + write) per launch from the `ncu --set full` captures.
usage: python profiles/make_traffic.py gpurun_out [r2]   (directory holding rN_prof_{cls,seg,c5,c4,pp}.ncu-rep;
a capture missing for round N falls back to the round-1 file)"""
import csv, io, json, os, subprocess, sys

def rows_of(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr, units = r[0], r[1]
    ci = {h: i for i, h in enumerate(hdr)}
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = []
    for x in r[2:]:
        b = sum(float(x[ci[k]]) * mul[units[ci[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        out.append((x[ci["Kernel Name"]], b))
    return out

d = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
rnd = sys.argv[2] if len(sys.argv) > 2 else "r1"


def rep(name):
    p = os.path.join(d, f"{rnd}_prof_{name}.ncu-rep")
    return p if os.path.exists(p) else os.path.join(d, f"r1_prof_{name}.ncu-rep")


t = {}
cls = rows_of(rep("cls"))
t["cls"] = {"sim_small_kernel": next(b for n, b in cls if "sim_small" in n)}
seg = rows_of(rep("seg"))
t["seg"] = {"upsample_kernel": next(b for n, b in seg if "upsample" in n),
            "sim_small_kernel": next(b for n, b in seg if "sim_small" in n)}
c5 = rows_of(rep("c5"))
t["openvocab"] = {"rz_sim_fwd_large": sum(b for n, b in c5)}
c4 = rows_of(rep("c4"))
seen, tot = set(), 0.0
for n, b in c4:                       # one step: the first launch of every distinct kernel
    key = n.split("(")[0]
    if key in seen and "prep_rows_bwd" not in key:
        continue
    seen.add(key); tot += b
t["contrastive"] = {"step": tot}
if os.path.exists(os.path.join(d, f"{rnd}_prof_pp.ncu-rep")):
    t["preprocess"] = {"step": sum(b for n, b in rows_of(os.path.join(d, f"{rnd}_prof_pp.ncu-rep")))}
json.dump(t, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json"), "w"), indent=1)
print(json.dumps(t, indent=1))
