"""Two training steps of the AlignTransformer (64 images) for the ncu launch list (capture_r2_align_train.sh)."""
import torch
from radzero_b200 import synthetic
from radzero_b200.align import AlignTransformer

enc = synthetic.build_align_encoder(seed=42, device="cuda")
mod = AlignTransformer(enc).train()
tok = synthetic.make_inputs(64, 1, seed=42, device="cuda")[0]
for _ in range(2):
    for p in mod.parameters():
        p.grad = None
    (mod(tok) * 1e-3).sum().backward()
torch.cuda.synchronize()
