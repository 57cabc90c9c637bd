import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from radzero_b200 import ops, synthetic
dev = torch.device("cuda:0")
B, N, L = 128, 1024, 1370
tok, text, gamma, beta, log_tau = synthetic.make_inputs(B, N, seed=1, device=dev)
Lp = ops.padded_tokens(L)
k16, _, _ = ops.prep_rows(tok, gamma, beta, rows_per_group=L, rows_per_group_padded=Lp)
k16 = k16.view(B, Lp, 768)
q16, _, _ = ops.prep_rows(text, gamma, beta)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for ws in (False, True):
    for wz in (False, True):
        t = timeit(lambda: ops.sim_fwd(k16, q16, L, 1 / 0.07, want_scores=ws, want_z=wz, drop_cls=True))
        print(f"want_scores={ws} want_z={wz}: {t*1e3:.1f} us")
