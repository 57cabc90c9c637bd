import torch, sys, os
sys.path.insert(0, os.getcwd())
from radzero_b200 import losses, ops, synthetic
dev = "cuda"
B, N = 256, 14
tok, text, gamma, beta, log_tau = synthetic.make_inputs(B, N, seed=42, device=dev)
fn = losses.RadZeroLoss(sim_op="cos").to(dev)
with torch.no_grad():
    fn.layer_norm.weight.copy_(gamma); fn.layer_norm.bias.copy_(beta)
ids = torch.zeros((N, 1), dtype=torch.int64, device=dev)
kp = [{"input_ids": ids, "attention_mask": torch.ones_like(ids)}]
def tm(enc): return {"text_features_wo_l2_norm": text, "text_features": text}
def A():
    with torch.no_grad():
        o = fn(kp, tok, tm, ddp_gather=False, need_attn_weights=True, compute_loss=False)
    return o["t2i_attn_weights"][0], o["t2i_logits"]
def Bf():
    lg, sc, z = fn.similarity(text, tok, want_scores=True)
    return sc, z
refA = [t.clone() for t in A()]
refB = [t.clone() for t in Bf()]
torch.cuda.synchronize()
bad = 0
for it in range(60):
    junk = [torch.full((int(torch.randint(1, 40, (1,)).item()) * 1000003,), float("nan"), device=dev) for _ in range(3)]
    del junk
    a = A(); b = Bf(); p = fn.similarity_prob(text, tok)
    torch.cuda.synchronize()
    for name, got, ref in (("A.scores", a[0], refA[0]), ("A.z", a[1], refA[1]), ("B.scores", b[0], refB[0]), ("B.z", b[1], refB[1])):
        d = (got - ref).abs()
        d = torch.nan_to_num(d, nan=1e9)
        if float(d.max()) > 0:
            bad += 1
            idx = torch.nonzero(d > 0)
            print(it, name, "max", float(d.max()), "count", idx.shape[0], "first", idx[:3].tolist(), "last", idx[-1].tolist())
print("bad", bad)

# ---- identify what the wrong column holds
print("--- search")
found = 0
for it in range(200):
    a = A()
    torch.cuda.synchronize()
    d = (a[0] - refA[0]).abs()
    if float(d.max()) > 0:
        idx = torch.nonzero(d.amax(1) > 0)          # (b, l)
        for b, l in idx.tolist()[:2]:
            wrong = a[0][b, :, l]                   # 14 scores
            dist = (refA[0] - wrong.view(1, N, 1)).pow(2).sum(1)      # (B, L)
            best = int(dist.argmin())
            bb, ll = divmod(best, dist.shape[1])
            print("iter", it, "wrong at", (b, l), "closest ref column", (bb, ll), "dist", float(dist.min()),
                  "dist to own", float(dist[b, l]), "wrong[:4]", wrong[:4].tolist(), "ref[:4]", refA[0][b, :4, l].tolist())
        found += 1
        if found >= 6: break
