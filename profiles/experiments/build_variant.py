"""Build a variant of librz_b200.so with extra -D flags for ONE source file (kernel ablations).

    python profiles/experiments/build_variant.py NAME rz_sim_bwd.cu -DRZ_EXP_D2=1
    -> radzero_b200/_build/variants/librz_NAME.so   (load it with RZ_B200_LIB=<path>)
The other objects are the ones of the regular build (radzero_b200/_build/*.o).
"""
import os, subprocess, sys
sys.path.insert(0, os.getcwd())
from radzero_b200 import build as B

name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
out_dir = os.path.join(B.BUILD, "variants")
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, f"{src[:-3]}_{name}.o")
subprocess.run([B.NVCC, *B.ARCH_FLAGS, *B.CFLAGS, *flags, "-c", os.path.join(B.CSRC, src), "-o", obj],
               check=True, capture_output=True)
objs = [obj if f == src else os.path.join(B.BUILD, f[:-3] + ".o") for f in B._sources()]
lib = os.path.join(out_dir, f"librz_{name}.so")
subprocess.run([B.NVCC, *B.ARCH_FLAGS, "-shared", "-o", lib, *objs, "-cudart", "static"], check=True)
print(lib)
