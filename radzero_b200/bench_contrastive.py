"""Benchmark of the contrastive training step (BASELINE.json configs[3]: 1024 images x ~6
sentences each, image batch sharded over the GPUs with the text all-gather).

One step = forward + backward of the fused node (radzero_b200/training.py) from
``vision_tokens`` / ``text_features`` to their gradients (+ d gamma, d beta, d log tau);
encoders and optimiser are outside the path.  STRONG scaling: the
global batch is fixed, each of the W ranks owns 1024 / W images and their sentences.
"""
from __future__ import annotations

import os
import time

import torch

L, D = 1370, 768
B_GLOBAL = 1024
UNIT_FLOP = 2.0 * L * D            # one GEMM unit per (image, sentence) pair
GEMM_UNITS = 6                     # fwd: S, P.K ; bwd: T, dQ, dK (x2)   (recompute of S not counted)


def _traffic():
    """DRAM bytes of one step from the committed ncu capture (profiles/traffic.json), 1 GPU."""
    import json
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    try:
        return json.load(open(p))["contrastive"]["step"]
    except Exception:
        return None


CHUNKS = 8                         # the global batch is generated in 8 fixed chunks of 128 images


def _inputs(rank, world, dev, dtype, b_global=B_GLOBAL):
    """This rank's slice of ONE fixed global problem: the same 1024 images / 6084 sentences whatever
    the number of ranks (chunk c = images [128c, 128c+128) from seed 1000 + c), so the loss printed
    at 1, 2, 4 and 8 GPUs is the same number."""
    from radzero_b200 import synthetic
    assert CHUNKS % world == 0, "world size must divide 8"
    counts_all = synthetic.sentence_counts(b_global, seed=42)
    per = b_global // CHUNKS
    toks, texts, counts = [], [], []
    gamma = beta = None
    for c in range(rank * CHUNKS // world, (rank + 1) * CHUNKS // world):
        cc = counts_all[c * per:(c + 1) * per]
        tok, text, gamma, beta, _ = synthetic.make_inputs(per, sum(cc), seed=1000 + c, device=dev)
        toks.append(tok.to(dtype)); texts.append(text.to(dtype)); counts += cc
    b_local = b_global // world
    gm = synthetic.group_map_from_counts(counts, first_image=rank * b_local, device=dev)
    # one LayerNorm for the whole job: gamma / beta from a fixed seed so that all ranks agree
    _, _, gamma, beta, _ = synthetic.make_inputs(1, 1, seed=999, device=dev)
    return torch.cat(toks), torch.cat(texts), gamma, beta, gm, sum(counts_all)


def run(args, world, rank, local, pk, steps=None, warmup=None, quiet=False):
    import torch.distributed as dist
    from radzero_b200 import _lib, losses, training
    dev = torch.device("cuda", local)
    steps = steps or max(2, min(args.steps, 10))
    warmup = warmup if warmup is not None else 3
    dtype = torch.bfloat16
    tok, text, gamma, beta, gm, n_total = _inputs(rank, world, dev, dtype)
    fn = losses.RadZeroLoss(sim_op="cos").to(dev)
    with torch.no_grad():
        fn.layer_norm.weight.copy_(gamma)
        fn.layer_norm.bias.copy_(beta)
    distributed = world > 1

    def step(tk, tx):
        tk = tk.detach().requires_grad_(True)
        tx = tx.detach().requires_grad_(True)
        fn.zero_grad(set_to_none=True)
        res = training.contrastive_step(fn, tx, gm, tk, distributed=distributed)
        res["loss"].backward()
        return res["loss"].detach(), tk.grad

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        loss, _ = step(tok, text)
    barrier()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss, _ = step(tok, text)
    e1.record()
    barrier()
    launches = _lib.launch_count() - n0
    ms = e0.elapsed_time(e1)
    if distributed:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_step = ms / steps
    flops = GEMM_UNITS * UNIT_FLOP * B_GLOBAL * n_total
    ach = flops / (ms_step * 1e-3) / 1e12 / world            # per GPU
    # end to end: host-resident (pinned) inputs, loss read back every step.  As in a real training loop
    # the NEXT step's inputs are copied on a side stream while the current step computes (double
    # buffering); every step's H2D copy and D2H read still happen inside the timed region.
    h_tok, h_txt = tok.cpu().pin_memory(), text.cpu().pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)

    # two static device buffers (ping-pong): no allocator traffic inside the timed region
    d_tok = [torch.empty(h_tok.shape, dtype=h_tok.dtype, device=dev) for _ in range(2)]
    d_txt = [torch.empty(h_txt.shape, dtype=h_txt.dtype, device=dev) for _ in range(2)]
    used = [None, None]            # event: the step that last read buffer k has been enqueued and finished

    def upload(k):
        with torch.cuda.stream(copy_stream):
            if used[k] is not None:
                copy_stream.wait_event(used[k])
            d_tok[k].copy_(h_tok, non_blocking=True)
            d_txt[k].copy_(h_txt, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def e2e_loop(n):
        last = None
        ev = upload(0)
        for i in range(n):
            k = i & 1
            main.wait_event(ev)
            if i + 1 < n:
                ev = upload(k ^ 1)
            a, b = d_tok[k], d_txt[k]
            l, _ = step(a, b)
            last = l.item()                     # D2H read of the step's loss
            used[k] = torch.cuda.Event()
            used[k].record(main)
        return last

    e2e_loop(2)
    barrier()
    ks = 4
    e0.record()
    l_host = e2e_loop(ks)
    e1.record()
    barrier()
    ms2 = e0.elapsed_time(e1)
    if distributed:
        t = torch.tensor([ms2], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms2 = float(t.item())
    # the MP-NCE loss kernels on their own (K10): algorithmic bytes = read Z + write dZ
    from radzero_b200 import ops
    b_local = B_GLOBAL // world
    zt = torch.randn(n_total, b_local, device=dev) * 0.3
    gm_all = gm if not distributed else torch.arange(n_total, device=dev) % B_GLOBAL
    def nce():
        rs, ps, cn, cp = ops.mpnce_partials(zt, gm_all, rank * b_local, 1.0 / 0.07)
        return ops.mpnce_finish(zt, gm_all, rank * b_local, B_GLOBAL, 1.0 / 0.07, rs, ps, cn, cp)
    for _ in range(3):
        nce()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        nce()
    e1.record()
    torch.cuda.synchronize()
    nce_ms = e0.elapsed_time(e1) / 10
    nce_bytes = 2 * n_total * b_local * 4
    mpnce = {"ms": nce_ms, "algorithmic_bytes": nce_bytes, "achieved_gbs": nce_bytes / (nce_ms * 1e-3) / 1e9,
             "frac_of_hbm": nce_bytes / (nce_ms * 1e-3) / 1e9 / pk["hbm"],
             "note": "7 small kernels (two phases + coefficient vectors); Z is 25 MB at C4 and stays in L2, "
                     "so this is launch/latency-bound, not HBM-bound"}
    out = {
        "metric": "contrastive steps/sec", "value": 1e3 / ms_step, "unit": "steps/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms_step, "scaling": "strong", "dtype": "f16",
        "loss": float(loss.item()), "gpu_launches": int(launches),
        "config": {"workload": f"C4 contrastive step, {B_GLOBAL} images x {n_total} sentences (n_i ~ U{{3..9}}), "
                               f"image batch sharded x{world} with text all-gather",
                   "input_dtype": "bf16", "tokens": L, "hidden": D},
        "roofline": {"bound": "tensor", "kernel": "whole step (sim_fwd + rz_sim_bwd GEMM passes)",
                     "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"],
                     "algorithmic_flops_per_step": flops, "traffic": _traffic(),
                     "peak_source": pk["src"] + " sustained (kernel timed inside a long step)"},
        "e2e": {"value": 1e3 / (ms2 / ks), "unit": "steps/s",
                "h2d_bytes_per_step": (h_tok.numel() + h_txt.numel()) * 2, "d2h_bytes_per_step": 4,
                "api": "RadZeroLoss.forward + backward", "steps": ks,
                "note": "inputs of step i+1 are uploaded on a copy stream while step i computes"},
        "mpnce": mpnce,
    }
    return out
