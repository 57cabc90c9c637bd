import torch, sys, os, collections
sys.path.insert(0, os.getcwd())
from radzero_b200 import losses, ops, synthetic
dev = "cuda"
S = int(sys.argv[1])
B, L, N = 64, 1370, 14
T = (L + 15) // 16
tok, text, gamma, beta, log_tau = synthetic.make_inputs(B, N, tokens_per_image=L, seed=42, device=dev)
lt_ = torch.full((1,), -2.659, device=dev)
q16, _, _ = ops.prep_rows(text, gamma, beta)
def f():
    o = ops.sim_fwd_tokens(tok, gamma, beta, q16, 1.0, want_scores=True, drop_cls=False, log_tau_scale=lt_)
    return o["scores"]
ref = f().clone()
n_ctas = 148
total = B * T
def rb(c): return (total * c) // n_ctas
for it in range(400):
    s = f(); torch.cuda.synchronize()
    d = (s - ref).abs().amax(1)
    for b, l in torch.nonzero(d > 0).tolist():
        g = b * T + l // 16
        c = (g * n_ctas) // total
        while rb(c + 1) <= g: c += 1
        while rb(c) > g: c -= 1
        lt = g - rb(c)
        rg = lt * 4 + (l % 16) // 4
        # image start within CTA
        img_lt0 = b * T - rb(c)
        print(f"it {it} img {b} tok {l} cta {c} range {rb(c)}-{rb(c+1)} lt {lt} (img starts at lt {img_lt0}) stage {lt % S} use {lt // S} rg {rg} team {rg % 6} ringuse {rg // 6} w {l % 4} q4 {(l % 16)//4}")
