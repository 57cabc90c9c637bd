import sys, os
sys.path.insert(0, os.getcwd())
exec(open("profiles/experiments/r2_stress_shapes.py").read().split("run(256, 1370)\n")[0])
run(256, 1370, iters=300)
run(64, 1370, iters=300)
