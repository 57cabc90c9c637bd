import torch, sys, os
sys.path.insert(0, os.getcwd())
from radzero_b200 import ops
B, L, D = 256, 1370, 768
qkv = (torch.randn(B, L, 3 * D, device="cuda") * 0.5).half()
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("attention ms", timeit(lambda: ops.attention(qkv, 12)))
