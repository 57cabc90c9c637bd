"""What does the second launch (merge_partials_kernel) cost inside a steady loop?  want_z=False skips it."""
import torch, sys, os
sys.path.insert(0, os.getcwd())
from radzero_b200 import ops, synthetic
dev = "cuda"; N = 14
lt = torch.full((1,), -2.659, device=dev)
for B in (64, 256):
    tok, text, gamma, beta, _ = synthetic.make_inputs(B, N, seed=42, device=dev)
    q16, _, _ = ops.prep_rows(text, gamma, beta)
    for wz in (True, False):
        f = lambda: ops.sim_fwd_tokens(tok, gamma, beta, q16, 1.0, want_scores=False, want_z=wz, z_sigmoid=True,
                                       z_image_major=True, log_tau_z=lt, log_tau_scale=lt)
        for _ in range(5): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(20_000_000)
        e0.record()
        for _ in range(50): f()
        e1.record(); torch.cuda.synchronize()
        print("B", B, "want_z", wz, round(e0.elapsed_time(e1) / 50 * 1000, 1), "us")
