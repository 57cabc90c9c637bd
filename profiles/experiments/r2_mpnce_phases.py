"""Phase timing inside mpnce_partials_kernel (build variant -DRZ_EXP_MPNCE_TIMING prints %globaltimer deltas
of the last CTA) + event timing of the two launches at C4 size."""
import torch, sys, os
sys.path.insert(0, os.getcwd())
from radzero_b200 import ops
dev = "cuda"
torch.manual_seed(0)
for n, b in ((6084, 1024), (6084, 256)):
    z = (torch.rand(n, b, device=dev) * 2 - 1)
    counts = torch.randint(3, 10, (1024,))
    gm = torch.repeat_interleave(torch.arange(1024), counts)[:n].to(dev)
    lt = torch.full((1,), -2.659, device=dev)
    for _ in range(3):
        rs, ps, cn, cp = ops.mpnce_partials(z, gm, 0, log_tau=lt, b_global=1024)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda._sleep(20_000_000)
    e[0].record()
    for _ in range(10): ops.mpnce_partials(z, gm, 0, log_tau=lt, b_global=1024)
    e[1].record(); torch.cuda.synchronize()
    print("n", n, "b_local", b, "partials", round(e[0].elapsed_time(e[1]) * 100, 1), "us")
