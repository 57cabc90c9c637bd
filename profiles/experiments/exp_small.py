import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from radzero_b200 import ops, synthetic
dev = torch.device("cuda:0")
def run(B, N, L=1370, dtype=torch.float32, scores=True):
    tok, text, gamma, beta, log_tau = synthetic.make_inputs(B, N, tokens_per_image=L, seed=3, device=dev)
    tok = tok.to(dtype)
    Lp = ops.padded_tokens(L)
    q16, _, _ = ops.prep_rows(text, gamma, beta)
    k16, _, _ = ops.prep_rows(tok, gamma, beta, rows_per_group=L, rows_per_group_padded=Lp)
    ref = ops.sim_fwd(k16.view(B, Lp, 768), q16, L, 1 / 0.07, want_scores=scores, drop_cls=True)
    out = ops.sim_fwd_tokens(tok, gamma, beta, q16, 1 / 0.07, want_scores=scores, drop_cls=True)
    torch.cuda.synchronize()
    dz = (out["z"] - ref["z"]).abs().max().item()
    ds = (out["scores"] - ref["scores"]).abs().max().item() if scores else 0.0
    print(f"B={B} N={N} L={L} {dtype}: max|dz|={dz:.2e} max|ds|={ds:.2e} nan={torch.isnan(out['z']).any().item()}")
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for cfg in [(1, 14), (2, 5), (3, 16), (7, 14), (64, 8), (256, 14)]:
    run(*cfg)
run(5, 14, L=50); run(4, 3, L=200); run(9, 14, dtype=torch.bfloat16); run(9, 14, dtype=torch.float16)
for B, N, sc in [(256, 14, False), (64, 8, True), (1, 14, True)]:
    tok, text, gamma, beta, log_tau = synthetic.make_inputs(B, N, seed=3, device=dev)
    q16, _, _ = ops.prep_rows(text, gamma, beta)
    t = timeit(lambda: ops.sim_fwd_tokens(tok, gamma, beta, q16, 1 / 0.07, want_scores=sc, drop_cls=True))
    gb = tok.numel() * 4 / 1e9
    print(f"B={B} N={N} scores={sc}: {t*1e3:.1f} us  ({gb / (t * 1e-3):.0f} GB/s)")
