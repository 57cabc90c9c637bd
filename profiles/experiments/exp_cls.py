"""C2 device time of the fused small-N kernel alone (256 images x 14 prompts), median of 5 x 20 launches."""
import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from radzero_b200 import ops, synthetic
dev = torch.device("cuda:0")
tok, text, gamma, beta, log_tau = synthetic.make_inputs(256, 14, seed=42, device=dev)
q16, _, _ = ops.prep_rows(text, gamma, beta)
fn = lambda: ops.sim_fwd_tokens(tok, gamma, beta, q16, 1.0, want_scores=False, z_sigmoid=True, z_image_major=True, log_tau_z=log_tau, log_tau_scale=log_tau)
for _ in range(5): fn()
ts = []
for _ in range(5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 20)
ts.sort()
print(f"sim_small C2: median {ts[2]*1e3:.1f} us  min {ts[0]*1e3:.1f} us  -> {256*1370*768*4/ts[2]/1e6:.0f} GB/s")
