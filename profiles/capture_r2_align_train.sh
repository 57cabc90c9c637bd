# AlignTransformer training step (forward + backward on the kernels): plain run first, then the ncu launch list
# and one full capture of the two attention-backward kernels.  Run on the GPU box through gpurun.
set -x
NCU="ncu --clock-control none"
python bench.py --workload align_train --steps 5 > gpurun_out/r2_align_train.json 2> gpurun_out/r2_align_train.err || exit 1
$NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file gpurun_out/r2_launches_align_train.csv python -c "
import torch
from radzero_b200 import synthetic
from radzero_b200.align import AlignTransformer
enc = synthetic.build_align_encoder(seed=42, device='cuda')
mod = AlignTransformer(enc).train()
tok = synthetic.make_inputs(64, 1, seed=42, device='cuda')[0]
for _ in range(2):
    for p in mod.parameters(): p.grad = None
    (mod(tok) * 1e-3).sum().backward()
torch.cuda.synchronize()
" > gpurun_out/r2_ncu_align_train.log 2>&1
