import torch, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from radzero_b200 import inference, ops, _lib
import oracle
dev = torch.device("cuda:0")
M, size = 512, (518, 518)
g = torch.Generator(device=dev).manual_seed(1)
scores = torch.randn(M, 1369, generator=g, device=dev) * 3
masks = (torch.rand(M, *size, generator=g, device=dev) < 0.05).to(torch.uint8)
def timeit(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t_stats = timeit(lambda: inference.dice_sweep_stats(scores, masks, size, "blip"))
t_map = timeit(lambda: inference.interpolate_similarity_scores(scores, size, "blip", mode="sigmoid"))
print(f"fused stats (101 thresholds, {M} maps {size}): {t_stats:.3f} ms = {M / t_stats * 1e3:.0f} maps/s;  "
      f"sigmoid map only: {t_map:.3f} ms")
# the reference's way on the host, 8 maps
s8, m8 = scores[:8].cpu(), masks[:8].cpu()
t0 = time.perf_counter(); oracle.dice_sweep_stats(s8, m8, size, "blip"); dt = time.perf_counter() - t0
print(f"reference-style CPU loop: {dt / 8 * 1e3:.1f} ms per map ({os.cpu_count()} cores) -> {8 / dt:.1f} maps/s")
