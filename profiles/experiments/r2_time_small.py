import torch, sys, os
sys.path.insert(0, os.getcwd())
from radzero_b200 import ops, synthetic
dev="cuda"
B,N=256,14
tok, text, gamma, beta, _ = synthetic.make_inputs(B, N, seed=42, device=dev)
q16,_,_ = ops.prep_rows(text, gamma, beta)
lt = torch.full((1,), -2.659, device=dev)
for dt in (torch.float32, torch.bfloat16, torch.float16):
    t = tok.to(dt)
    f = lambda: ops.sim_fwd_tokens(t, gamma, beta, q16, 1.0, want_scores=False, z_sigmoid=True, z_image_major=True, log_tau_z=lt, log_tau_scale=lt)
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): f()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/50
    by=t.numel()*t.element_size()
    print(dt, round(ms,4), "ms", round(by/ms/1e6,1), "GB/s")
# the prepped-operand kernel for comparison
Lp=ops.padded_tokens(1370)
k16,_,_=ops.prep_rows(tok, gamma, beta, rows_per_group=1370, rows_per_group_padded=Lp)
f=lambda: ops.sim_fwd(k16.view(B,Lp,768), q16, 1370, 1.0, want_scores=False, log_tau_scale=lt)
for _ in range(5): f()
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): f()
e1.record(); torch.cuda.synchronize()
print("sim_fwd (prepped fp16)", e0.elapsed_time(e1)/50)
