"""sim_small_kernel: time against the number of images -- fixed cost vs streaming rate."""
import torch, sys, os
sys.path.insert(0, os.getcwd())
from radzero_b200 import ops, synthetic
dev = "cuda"
N = 14
lt = torch.full((1,), -2.659, device=dev)
res = []
for B in (37, 74, 148, 256, 296, 512, 592, 1024, 1184):
    tok, text, gamma, beta, _ = synthetic.make_inputs(B, N, seed=42, device=dev)
    q16, _, _ = ops.prep_rows(text, gamma, beta)
    f = lambda: ops.sim_fwd_tokens(tok, gamma, beta, q16, 1.0, want_scores=False, z_sigmoid=True,
                                   z_image_major=True, log_tau_z=lt, log_tau_scale=lt)
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(20_000_000)
    e0.record()
    for _ in range(30): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    by = tok.numel() * 4
    res.append((B, ms))
    print(B, round(ms, 4), "ms", round(by / ms / 1e6, 1), "GB/s", flush=True)
    del tok
(b0, t0), (b1, t1) = res[3], res[-2]
slope = (t1 - t0) / (b1 - b0)
print("slope ms/image", slope, "-> streaming GB/s", 1370 * 768 * 4 / slope / 1e6, "fixed ms", t0 - slope * b0)
