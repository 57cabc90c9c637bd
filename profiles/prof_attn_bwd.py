import torch
from radzero_b200 import ops
B, L, H = 64, 1370, 12
qkv = (torch.randn(B, L, 3 * H * 64, device='cuda') * 0.5).half()
out = ops.attention(qkv, H)
dout = torch.randn(B, L, H * 64, device='cuda').half()
for _ in range(2):
    ops.attention_bwd(qkv, out, dout, H, 0.125)
torch.cuda.synchronize()
