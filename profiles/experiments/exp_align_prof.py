"""One launch of each AlignTransformer kernel at B images (for ncu captures)."""
import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from radzero_b200 import ops, synthetic
from radzero_b200.align import pack_layer
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L, D = 1370, 768
enc = synthetic.build_align_encoder(seed=1, device=dev)
tok = synthetic.make_inputs(B, 1, seed=1, device=dev)[0]
w = pack_layer(enc.layer[0], dev)
x2 = tok.view(B * L, D).clone()
for _ in range(2):
    h = ops.ln_rows(x2, w["g1"], w["b1"], 1e-6)
    qkv = ops.linear(h, w["wqkv"], w["bqkv"], "bias")
    a = ops.attention(qkv.view(B, L, 3 * D), 12)
    ops.linear(a.view(B * L, D), w["wo"], w["bo"], "residual", scale=w["ls1"], residual=x2, out=x2)
    g = ops.linear(h, w["w1"], w["bf1"], "gelu")
    ops.linear(g, w["w2"], w["bf2"], "residual", scale=w["ls2"], residual=x2, out=x2)
torch.cuda.synchronize()
print("ok")
