import torch, sys, os, collections
sys.path.insert(0, os.getcwd())
from radzero_b200 import losses, ops, synthetic
dev = "cuda"
def run(B, L, N=14, dtype=torch.float32, iters=150, scores=True):
    tok, text, gamma, beta, log_tau = synthetic.make_inputs(B, N, tokens_per_image=L, seed=42, device=dev)
    tok = tok.to(dtype)
    fn = losses.RadZeroLoss(sim_op="cos").to(dev)
    with torch.no_grad():
        fn.layer_norm.weight.copy_(gamma); fn.layer_norm.bias.copy_(beta)
    q16, _, _ = ops.prep_rows(text, gamma, beta)
    def f():
        o = ops.sim_fwd_tokens(tok, gamma, beta, q16, 1.0, want_scores=scores, drop_cls=False, log_tau_scale=fn.loss_temperature)
        return o["scores"], o["z"]
    ref = [t.clone() if t is not None else None for t in f()]
    torch.cuda.synchronize()
    where = collections.Counter(); zbad = 0
    for it in range(iters):
        s, z = f()
        torch.cuda.synchronize()
        if scores:
            d = (s - ref[0]).abs().amax(1)
            for b, l in torch.nonzero(d > 0).tolist(): where[l] += 1
        if float((z - ref[1]).abs().max()) > 0: zbad += 1
    print(f"B={B} L={L} dtype={dtype} scores={scores}: bad tokens {dict(where)} z-nondeterministic iters {zbad}/{iters}", flush=True)
run(256, 1370)
run(64, 1370)
run(256, 1000)
run(256, 1370, dtype=torch.bfloat16)
run(256, 1370, scores=False)
run(8, 1370, iters=300)
run(256, 1370, N=8)
