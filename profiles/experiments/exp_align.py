"""Per-kernel timing of the AlignTransformer forward at the C2 batch (256 images x 1370 tokens)
next to the stock transformers Dinov2Encoder (fp32 = the reference's inference dtype,
inference/utils.py:37; bf16 autocast = its training dtype)."""
import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from radzero_b200 import ops, synthetic
from radzero_b200.align import AlignTransformer, pack_layer
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L, D = 1370, 768
enc = synthetic.build_align_encoder(seed=1, device=dev)
tok = synthetic.make_inputs(B, 1, seed=1, device=dev)[0]
w = pack_layer(enc.layer[0], dev)
M = B * L
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
x2 = tok.view(M, D).clone()
h = ops.ln_rows(x2, w["g1"], w["b1"], 1e-6)
qkv = ops.linear(h, w["wqkv"], w["bqkv"], "bias")
a = ops.attention(qkv.view(B, L, 3 * D), 12)
g = ops.linear(h, w["w1"], w["bf1"], "gelu")
rows = []
def rep(name, ms, flops=None, bytes_=None):
    s = f"{name:28s} {ms:8.3f} ms"
    if flops: s += f"  {flops / ms / 1e9:8.1f} TFLOP/s"
    if bytes_: s += f"  {bytes_ / ms / 1e6:8.1f} GB/s"
    print(s)
rep("ln_rows", timeit(lambda: ops.ln_rows(x2, w["g1"], w["b1"], 1e-6)), bytes_=M * D * 6)
rep("linear qkv (768->2304)", timeit(lambda: ops.linear(h, w["wqkv"], w["bqkv"], "bias")), 2.0 * M * D * 3 * D)
rep("attention", timeit(lambda: ops.attention(qkv.view(B, L, 3 * D), 12)), 4.0 * B * 12 * L * L * 64)
rep("linear proj+res (768->768)", timeit(lambda: ops.linear(a.view(M, D), w["wo"], w["bo"], "residual", scale=w["ls1"], residual=x2, out=x2)), 2.0 * M * D * D)
rep("linear fc1+gelu (768->3072)", timeit(lambda: ops.linear(h, w["w1"], w["bf1"], "gelu")), 2.0 * M * D * 4 * D)
rep("linear fc2+res (3072->768)", timeit(lambda: ops.linear(g, w["w2"], w["bf2"], "residual", scale=w["ls2"], residual=x2, out=x2)), 2.0 * M * D * 4 * D)
mod = AlignTransformer(enc).eval()
flops2 = 2 * (2.0 * M * D * D * 12 + 4.0 * B * 12 * L * L * 64)
rep("AlignTransformer (2 layers)", timeit(lambda: mod(tok), 3), flops2)
with torch.no_grad():
    t32 = timeit(lambda: enc(tok), 2)
    rep("HF Dinov2Encoder fp32", t32, flops2)
    torch.backends.cuda.matmul.allow_tf32 = True
    rep("HF Dinov2Encoder tf32", timeit(lambda: enc(tok), 2), flops2)
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.autocast("cuda", dtype=torch.bfloat16):
        rep("HF Dinov2Encoder bf16 autocast", timeit(lambda: enc(tok), 3), flops2)
    encb = synthetic.build_align_encoder(seed=1, device=dev).to(torch.bfloat16)
    tb = tok.to(torch.bfloat16)
    rep("HF Dinov2Encoder bf16 weights", timeit(lambda: encb(tb), 3), flops2)
