"""World-size-2 `gloo` test: the AlignTransformer under autograd (radzero_b200.align._AlignFn, kernels replaced
by the fp64 test double tests/cpu_align_ops.py) inside torch DistributedDataParallel.  The Function receives the
parameters as inputs and returns their gradients to autograd, so DDP's reducer must see every one of them and
average them over ranks exactly as it does for the stock module the reference trains
(`module_to_update: [align_transformer, ...]` under the HF Trainer's DDP).  Checker: the stock Dinov2Encoder
under single-process autograd on the concatenated batch."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from radzero_b200 import synthetic

SEED, L, B_LOCAL = 81, 13, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Wrapper(torch.nn.Module):
    """AlignTransformer.train_forward without its CUDA-only guard (the kernels are doubled on CPU here)."""

    def __init__(self, mod):
        super().__init__()
        self.mod = mod

    def forward(self, x):
        from radzero_b200 import align
        params = [p for l in self.mod.transformer_layers.layer for p in align.layer_params(l)]
        return align._AlignFn.apply(x, self.mod, *params)


def _worker(rank, world, port, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from radzero_b200 import align
        from tests import cpu_align_ops
        align.ops = cpu_align_ops
        align.OVERLAP_WEIGHT_GRADS = False
        enc = synthetic.build_align_encoder(seed=SEED, layers=1)
        ddp = torch.nn.parallel.DistributedDataParallel(_Wrapper(align.AlignTransformer(enc).train()))
        tok = synthetic.make_inputs(B_LOCAL * world, 1, tokens_per_image=L, seed=SEED)[0]
        up = torch.randn(tok.shape, generator=torch.Generator().manual_seed(SEED))
        sl = slice(rank * B_LOCAL, (rank + 1) * B_LOCAL)
        (ddp(tok[sl]) * up[sl]).sum().backward()
        out_q.put((rank, {n: p.grad.clone().numpy() for n, p in enc.named_parameters()}))
    except Exception:  # surface worker failures instead of a queue timeout
        import traceback
        out_q.put((rank, "error", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_ddp_averages_the_gradients_the_function_returns():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for g in got:
        assert g[1] != "error", g[2]
    grads = {r: g for r, g in got}
    # checker: the stock encoder on the whole batch; DDP averages, so compare with the total gradient / world
    ref = synthetic.build_align_encoder(seed=SEED, layers=1).double().train()
    tok = synthetic.make_inputs(B_LOCAL * world, 1, tokens_per_image=L, seed=SEED)[0].double()
    up = torch.randn(tok.shape, generator=torch.Generator().manual_seed(SEED)).double()
    (ref(tok)["last_hidden_state"] * up).sum().backward()
    for name, p in ref.named_parameters():
        a, b = torch.from_numpy(grads[0][name]).double(), torch.from_numpy(grads[1][name]).double()
        assert torch.equal(a, b), name                                   # all-reduced: identical on both ranks
        want = p.grad / world
        if name.endswith("key.bias"):
            continue                                                     # zero in exact arithmetic
        rel = ((a - want).norm() / want.norm()).item()
        assert rel <= 5e-3, (name, rel)
