import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from radzero_b200 import ops, synthetic
dev = torch.device("cuda:0")
B, N = 256, 14
tok, text, gamma, beta, log_tau = synthetic.make_inputs(B, N, seed=3, device=dev)
q16, _, _ = ops.prep_rows(text, gamma, beta)
for _ in range(3):
    ops.sim_fwd_tokens(tok, gamma, beta, q16, 1 / 0.07, want_scores=False, drop_cls=True)
torch.cuda.synchronize()
