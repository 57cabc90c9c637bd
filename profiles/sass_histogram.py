"""Opcode evidence that the shipped library is tcgen05 / TMEM / TMA code.
usage: python profiles/sass_histogram.py > profiles/r2_sass_histogram.md   (needs cuobjdump; no GPU)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "radzero_b200", "_build")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "SYNCS", "UTCBAR", "FFMA2", "FMUL2", "FADD2",
        "MUFU", "LDS", "STS", "LDG", "STG", "REDUX", "HMMA", "IMMA"]
print("# SASS opcode histogram of librz_b200.so (sm_100a), per object file\n")
print("`cuobjdump -sass radzero_b200/_build/*.o`, counted by `profiles/sass_histogram.py`.  UTCHMMA = tcgen05.mma "
      "(kind::f16), `.2CTA` = cta_group::2; UTMALDG / UTMASTG = TMA tensor load / store; LDTM / STTM = tcgen05.ld / "
      "tcgen05.st (TMEM); SYNCS = mbarrier; FFMA2 / FMUL2 / FADD2 = packed fp32 pairs.  HMMA / IMMA (legacy "
      "mma.sync) are 0 everywhere except rz_align_bwd.o: the two attention-backward kernels of the AlignTransformer "
      "training step are the one place still on the register-operand tensor path (profiles/r2_full_attn_bwd.md).\n")
print("| object | " + " | ".join(KEYS) + " | kernels |")
print("|---|" + "---|" * (len(KEYS) + 1))
tot = collections.Counter()
for f in sorted(os.listdir(BUILD)):
    if not f.endswith(".o"):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, f)], capture_output=True, text=True).stdout
    c = collections.Counter()
    kernels = 0
    for line in sass.splitlines():
        if "Function :" in line:
            kernels += 1
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        base = op.split(".")[0]
        if base in KEYS:
            c[base] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            c["UTCHMMA.2CTA"] += 1
    tot.update(c)
    print(f"| {f} | " + " | ".join(str(c.get(k, 0)) for k in KEYS) + f" | {kernels} |")
print("| **total** | " + " | ".join(str(tot.get(k, 0)) for k in KEYS) + " | |")
