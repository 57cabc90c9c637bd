"""sim_small_kernel on 16-bit tokens only (variants whose fp32 configuration does not fit shared memory)."""
import torch, sys, os
sys.path.insert(0, os.getcwd())
from radzero_b200 import ops, synthetic
dev = "cuda"; B, N = 256, 14
tok, text, gamma, beta, _ = synthetic.make_inputs(B, N, seed=42, device=dev)
q16, _, _ = ops.prep_rows(text, gamma, beta)
lt = torch.full((1,), -2.659, device=dev)
ref = None
for dt in (torch.bfloat16, torch.float16):
    t = tok.to(dt)
    f = lambda: ops.sim_fwd_tokens(t, gamma, beta, q16, 1.0, want_scores=False, z_sigmoid=True, z_image_major=True,
                                   log_tau_z=lt, log_tau_scale=lt)
    for _ in range(5): out = f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(dt, round(ms, 4), "ms", round(t.numel() * 2 / ms / 1e6, 1), "GB/s", "checksum", float(out["z"].double().sum()))
