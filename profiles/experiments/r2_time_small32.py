import torch, sys, os
sys.path.insert(0, os.getcwd())
from radzero_b200 import ops, synthetic
dev="cuda"; B,N=256,14
tok, text, gamma, beta, _ = synthetic.make_inputs(B, N, seed=42, device=dev)
q16,_,_ = ops.prep_rows(text, gamma, beta)
lt = torch.full((1,), -2.659, device=dev)
for ws in (False, True):
    f = lambda: ops.sim_fwd_tokens(tok, gamma, beta, q16, 1.0, want_scores=ws, z_sigmoid=True, z_image_major=True, log_tau_z=lt, log_tau_scale=lt)
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): f()
    e1.record(); torch.cuda.synchronize()
    print("fp32 scores=%s"%ws, round(e0.elapsed_time(e1)/50,4), "ms")
