"""World-size-2 (and 3) `gloo` tests of the HOST logic of the image-sharded contrastive step
(radzero_b200/training.py): ragged sentence gather, column offsets, row-sum all-reduce, loss
assembly, dL/dq all-reduce + slicing and the DDP gradient scale.  The kernels are replaced by
the torch-CPU test double tests/cpu_ops.py; the checker is the oracle's single-process autograd
(oracle.contrastive_step_reference), which restates the reference's redundant all-gather form
(losses.py:87-88, 156-161)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from radzero_b200 import synthetic


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, counts_all, L, ddp_compatible, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from radzero_b200 import losses, training
        from tests import cpu_ops
        torch.manual_seed(0)
        b_local = len(counts_all) // world
        B = len(counts_all)
        tok, text, gamma, beta, log_tau = synthetic.make_inputs(B, sum(counts_all), tokens_per_image=L, seed=21)
        offs = [0]
        for c in counts_all:
            offs.append(offs[-1] + c)
        i0, i1 = rank * b_local, (rank + 1) * b_local
        tok_l = tok[i0:i1].double().clone().requires_grad_(True)
        text_l = text[offs[i0]:offs[i1]].double().clone().requires_grad_(True)
        fn = losses.RadZeroLoss(sim_op="cos").double()
        with torch.no_grad():
            fn.layer_norm.weight.copy_(gamma)
            fn.layer_norm.bias.copy_(beta)
        gm = oracle.build_group_map(counts_all[i0:i1], rank=rank)      # i + rank * B_local
        res = training.contrastive_step(fn, text_l, gm, tok_l, distributed=True,
                                        ddp_compatible=ddp_compatible, kernel_ops=cpu_ops)
        res["loss"].backward()
        out_q.put((rank, res["loss"].item(), tuple(res["z"].shape), text_l.grad.numpy(), tok_l.grad.numpy(),
                   fn.layer_norm.weight.grad.numpy(), fn.layer_norm.bias.grad.numpy(),
                   fn.loss_temperature.grad.numpy()))
    except Exception as e:  # surface worker failures instead of a queue timeout
        import traceback
        out_q.put((rank, "error", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,counts,ddp", [(2, [2, 1, 3, 2], False), (2, [1, 4, 2, 2], True),
                                              (3, [2, 2, 1, 3, 1, 2], False)])
def test_sharded_step_matches_single_process(world, counts, ddp):
    L = 40
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, counts, L, ddp, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for r in results:
        assert r[1] != "error", r[2]
    results = sorted(results, key=lambda t: t[0])
    results = [tuple(torch.from_numpy(v) if hasattr(v, "dtype") and not torch.is_tensor(v) else v
                     for v in r) for r in results]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    B = len(counts)
    tok, text, gamma, beta, log_tau = synthetic.make_inputs(B, sum(counts), tokens_per_image=L, seed=21)
    gm = oracle.build_group_map(counts)
    loss, grads = oracle.contrastive_step_reference(text.double(), gm, tok.double(), gamma.double(),
                                                    beta.double(), log_tau.double())
    scale = float(world) if ddp else 1.0
    b_local = B // world
    text_grad = torch.cat([r[3] for r in results])
    tok_grad = torch.cat([r[4] for r in results])
    for r in results:
        assert abs(r[1] - loss.item()) < 1e-9 * abs(loss.item()) + 1e-12     # same global loss on every rank
        assert r[2] == (sum(counts), b_local)                               # local column block of Z
    assert (text_grad - scale * grads["text"]).abs().max() < 1e-8 * scale * grads["text"].abs().max() + 1e-14
    assert (tok_grad - scale * grads["vision_tokens"]).abs().max() < \
        1e-8 * scale * grads["vision_tokens"].abs().max() + 1e-14
    # shared parameters: each rank holds a partial; their SUM is the full gradient (times scale),
    # i.e. DDP's average over W ranks of (W * partial) is the reference gradient
    for idx, key in ((5, "gamma"), (6, "beta"), (7, "log_tau")):
        tot = sum(r[idx] for r in results)
        ref = scale * grads[key]
        assert (tot - ref).abs().max() < 1e-7 * ref.abs().max() + 1e-13, key
