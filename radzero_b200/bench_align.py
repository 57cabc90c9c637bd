"""Benchmark of the widened path: AlignTransformer (two DINOv2 layers,
exp/cxr_pt/model/align_transformers.py:37-45) -> VL-CABS ``similarity_prob`` at BASELINE.json
configs[1] (256 images x 14 prompts), i.e. the zero-shot classification step starting from the
vision encoder's output instead of from ``vision_tokens``.

One step = ``AlignTransformer.forward`` + ``RadZeroLoss.similarity_prob``.  The step is tensor-bound:
its algorithmic work is the twelve 768-wide GEMM units of the two layers plus the attention products,
2 * (24 * M * 768^2 + 4 * B * 12 * L^2 * 64) FLOP with M = B * L rows.
"""
from __future__ import annotations

import os
import time

import torch

L, D, HEADS, LAYERS = 1370, 768, 12, 2
B, N = 256, 14
DESC = ("C2 upstream: AlignTransformer (2 DINOv2 layers, 12 heads) + similarity_prob, "
        "256 images x 14 prompts from the vision encoder's output tokens")


def align_flops(b: int) -> float:
    m = b * L
    return LAYERS * (24.0 * m * D * D + 4.0 * b * HEADS * L * L * 64)


def _model(dev, seed=42):
    from radzero_b200 import losses, synthetic
    from radzero_b200.align import AlignTransformer
    enc = synthetic.build_align_encoder(seed=seed, device=dev)
    _, text, gamma, beta, _ = synthetic.make_inputs(1, N, seed=seed, device=dev)
    fn = losses.RadZeroLoss(sim_op="cos").to(dev)
    with torch.no_grad():
        fn.layer_norm.weight.copy_(gamma)
        fn.layer_norm.bias.copy_(beta)
    return AlignTransformer(enc).eval(), fn, text


def run(args, world, rank, local, pk, steps=None, warmup=None):
    import torch.distributed as dist
    from radzero_b200 import _lib, ops, synthetic
    from radzero_b200.align import pack_layer
    dev = torch.device("cuda", local)
    steps = steps or max(3, min(args.steps, 20))
    warmup = max(warmup or args.warmup, 3)
    align, fn, text = _model(dev)
    tok = synthetic.make_inputs(B, 1, seed=42 + rank, device=dev)[0]

    def step(t):
        # the last layer's tokens reach the similarity kernel in fp16 (RZ_LIN_RESIDUAL_F16): half the bytes
        return fn.similarity_prob(text, align(t, handoff_f16=True))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        res = step(tok)
    barrier()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        res = step(tok)
    e1.record()
    barrier()
    launches = _lib.launch_count() - n0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / steps
    value = B * N * world / (ms_per_step * 1e-3)

    # per-kernel device times (each launched `reps` times back to back between two events)
    w = pack_layer(align.transformer_layers.layer[0], dev)
    m = B * L
    x2 = tok.view(m, D).clone()
    h = ops.ln_rows(x2, w["g1"], w["b1"], w["eps1"])
    qkv = ops.linear(h, w["wqkv"], w["bqkv"], "bias")
    a = ops.attention(qkv.view(B, L, 3 * D), HEADS)
    g = ops.linear(h, w["w1"], w["bf1"], "gelu")
    stages = [
        ("prep_rows_kernel (LayerNorm -> fp16, x2 per layer)", lambda: ops.ln_rows(x2, w["g1"], w["b1"], w["eps1"]),
         None, m * D * 6.0),
        ("gemm_kernel<Lin BIAS> qkv 768->2304", lambda: ops.linear(h, w["wqkv"], w["bqkv"], "bias"), 2.0 * m * D * 3 * D, None),
        ("attn_kernel (tcgen05 flash attention, 12 heads x 64)", lambda: ops.attention(qkv.view(B, L, 3 * D), HEADS),
         4.0 * B * HEADS * L * L * 64, None),
        ("gemm_kernel<Lin RESIDUAL> proj 768->768", lambda: ops.linear(a.view(m, D), w["wo"], w["bo"], "residual",
                                                                        scale=w["ls1"], residual=x2, out=x2), 2.0 * m * D * D, None),
        ("gemm_kernel<Lin GELU> fc1 768->3072", lambda: ops.linear(h, w["w1"], w["bf1"], "gelu"), 2.0 * m * D * 4 * D, None),
        ("gemm_kernel<Lin RESIDUAL> fc2 3072->768", lambda: ops.linear(g, w["w2"], w["bf2"], "residual", scale=w["ls2"],
                                                                        residual=x2, out=x2), 2.0 * m * D * 4 * D, None),
    ]
    kms, detail = {}, {}
    for name, fnk, fl, by in stages:
        for _ in range(2):
            fnk()
        torch.cuda.synchronize()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        for _ in range(5):
            fnk()
        eb.record()
        torch.cuda.synchronize()
        t = ea.elapsed_time(eb) / 5
        kms[name] = round(t, 4)
        detail[name] = ({"tflops": round(fl / t / 1e9, 1)} if fl else {"gbs": round(by / t / 1e6, 1)})
    del x2, h, qkv, a, g
    # the similarity stage behind the hand-off: rz_sim_fwd_tokens on fp16 tokens (0.54 GB read once)
    x16 = align(tok, handoff_f16=True)
    for _ in range(3):
        fn.similarity_prob(text, x16)
    torch.cuda.synchronize()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    for _ in range(20):
        fn.similarity_prob(text, x16)
    eb.record()
    torch.cuda.synchronize()
    sim_ms = ea.elapsed_time(eb) / 20
    sim_bytes = B * L * D * 2 + N * D * 4 + B * N * 4
    sim_stage = {"ms": round(sim_ms, 4), "bytes_read": sim_bytes, "input_dtype": "fp16 (fc2 epilogue hand-off)",
                 "frac_of_hbm": round(sim_bytes / (sim_ms * 1e-3) / 1e9 / pk["hbm"], 3)}
    del x16
    flops = align_flops(B)
    ach = flops / (ms_per_step * 1e-3) / 1e12
    roof = {"bound": "tensor", "kernel": "whole step (6 GEMM-shaped kernels per layer x 2 layers + similarity)",
            "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"],
            "traffic": None, "peak_source": pk["src"] + " sustained (kernels timed inside a long step)",
            "algorithmic_flops_per_step": flops, "kernel_ms": kms, "kernel_rates": detail}

    # end to end: pinned host tokens -> H2D -> AlignTransformer (in place on the upload) -> prob -> D2H
    h_tok = tok.cpu().pin_memory()
    d_tok = [torch.empty_like(tok) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    r_host = torch.empty((B, N), dtype=torch.float32, pin_memory=True)
    used = [None, None]

    def upload(k):
        with torch.cuda.stream(copy_stream):
            if used[k] is not None:
                copy_stream.wait_event(used[k])
            d_tok[k].copy_(h_tok, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def e2e_loop(n):
        ev = upload(0)
        for i in range(n):
            k = i & 1
            main.wait_event(ev)
            if i + 1 < n:
                ev = upload(k ^ 1)
            r_host.copy_(fn.similarity_prob(text, align(d_tok[k], inplace=True, handoff_f16=True)), non_blocking=True)
            used[k] = torch.cuda.Event()
            used[k].record(main)
        main.synchronize()

    ksteps = max(2, min(steps, 5))
    e2e_loop(2)
    barrier()
    e0.record()
    e2e_loop(ksteps)
    e1.record()
    barrier()
    ms2 = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms2], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms2 = float(t.item())
    e2e = {"value": B * N * world / (ms2 / ksteps * 1e-3), "unit": "maps/s",
           "h2d_bytes_per_step": h_tok.numel() * 4, "d2h_bytes_per_step": r_host.numel() * 4, "steps": ksteps,
           "api": "AlignTransformer.forward + RadZeroLoss.similarity_prob"}
    return {"metric": "similarity maps/sec", "value": value, "unit": "maps/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "scaling": "weak", "dtype": "f16",
            "gpu_launches": int(launches), "e2e": e2e, "roofline": roof,
            "config": {"workload": DESC, "images_per_gpu": B, "prompts": N, "tokens": L, "hidden": D,
                       "input_dtype": "fp32", "l2": "activations (>= 0.5 GB per kernel) larger than L2; no flush",
                       "parallelism": f"images sharded x{world}, no collective"},
            "sim_stage": sim_stage, "prob_checksum": float(res.double().sum().item())}


# ----------------------------------------------------------------------------- training step (forward + backward)
def train_flops(b: int) -> float:
    """Forward + backward of the two layers: the twelve linear products per layer cost 2 M K N each way
    (forward, dX, dW) = 3x the forward's, the attention backward five L x L x 64 products per head (the
    recomputation of the scores counted, as FlashAttention does) next to the forward's two."""
    m = b * L
    return LAYERS * (3 * 24.0 * m * D * D + (4.0 + 10.0) * b * HEADS * L * L * 64)


def run_train(args, world, rank, local, pk, images: int = 64, steps: int = 5, warmup: int = 3, stock: bool = True):
    """One step = ``AlignTransformer.forward`` under autograd + ``backward`` of a linear loss on ``images``
    images per GPU (config 4 trains the module on 128 images a GPU; 64 keeps the stock arm's materialised
    attention matrices small).  Reports the step on the kernels, the same step through the stock HF modules
    (``kernel_backward = False``: torch eager + cuBLAS, fp32 and under bf16 autocast as the reference trains,
    radzero.yaml ``bf16: true``) and the device time of every backward kernel."""
    from radzero_b200 import _lib, ops, synthetic
    from radzero_b200.align import AlignTransformer, layer_forward_train, pack_layer, pack_layer_bwd
    dev = torch.device("cuda", local)
    enc = synthetic.build_align_encoder(seed=42, device=dev)
    mod = AlignTransformer(enc).train()
    tok = synthetic.make_inputs(images, 1, seed=42 + rank, device=dev)[0]
    up = torch.randn(tok.shape, device=dev, generator=torch.Generator(dev).manual_seed(7)) * 1e-3

    def step():
        for p in mod.parameters():
            p.grad = None
        (mod(tok) * up).sum().backward()

    def timed(fn, n, w):
        for _ in range(w):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    n0 = _lib.launch_count()
    ms = timed(step, steps, warmup)
    launches = (_lib.launch_count() - n0) // (steps + warmup)
    flops = train_flops(images)
    out = {"workload": f"AlignTransformer forward + backward, {images} images x {L} tokens, 2 layers",
           "ms_per_step": round(ms, 3), "images_per_s": round(images / (ms * 1e-3), 1), "gpu_launches_per_step": int(launches),
           "tflops": round(flops / ms / 1e9, 1), "frac_of_tensor_sustained": round(flops / ms / 1e9 / pk["tf_sust"], 3)}
    if stock:
        mod.kernel_backward = False
        try:
            out["stock_fp32_ms"] = round(timed(step, 2, 1), 3)

            def step_bf16():
                for p in mod.parameters():
                    p.grad = None
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    y = mod(tok)
                (y.float() * up).sum().backward()
            out["stock_bf16_autocast_ms"] = round(timed(step_bf16, 3, 2), 3)
        except Exception as e:      # the stock arm materialises B x 12 x L x L matrices
            out["stock_error"] = repr(e)[:160]
        mod.kernel_backward = True
        for p in mod.parameters():
            p.grad = None
        torch.cuda.empty_cache()
    # per-kernel device times of one layer's backward
    layer = enc.layer[0]
    w, wb = pack_layer(layer, dev), pack_layer_bwd(layer, dev)
    m = images * L
    with torch.no_grad():
        x2 = tok.view(m, D).clone()
        _, (x2, h1, qkv, a, y, h2, g) = layer_forward_train(x2, images, L, w)
        dz = up.view(m, D).contiguous()
        sc = ops.grad_scale(dz)
        do2 = ops.ls_cast_bwd(dz, w["ls2"], None, sc, None)
        doT, gT = ops.transpose_pad(do2), ops.transpose_pad(g)
        dg = ops.linear(do2, wb["w2_t"], None, "bias")
        u = ops.linear(h2, w["w1"], w["bf1"], "bias")
        da = ops.linear(do2, wb["wo_t"], None, "bias")
        acc = torch.zeros(D, 4 * D, device=dev)
        zero = torch.zeros(4 * D, device=dev)
        att_fl = 10.0 * images * HEADS * L * L * 64
        stages = [
            ("attn_bwd_dq_kernel + attn_bwd_dkv_kernel (mma.sync)", lambda: ops.attention_bwd(
                qkv.view(images, L, 3 * D), a.view(images, L, D), da.view(images, L, D), HEADS, 0.125), att_fl, None),
            ("gemm_kernel<Lin BIAS> dX 768->3072 (fc2^T)", lambda: ops.linear(do2, wb["w2_t"], None, "bias"), 2.0 * m * D * 4 * D, None),
            ("gemm_kernel<Lin RESIDUAL> dW 768 x 3072, K = rows", lambda: ops.linear(doT, gT, None, "residual", scale=sc[4:4 + 4 * D],
                                                                                    residual=acc, out=acc), 2.0 * m * D * 4 * D, None),
            ("transpose_kernel 3072 columns (+ bias gradient)", lambda: ops.transpose_pad(g, sc, zero), None, m * 4 * D * 4.0),
            ("gelu_bwd_kernel", lambda: ops.gelu_bwd(dg, u), None, m * 4 * D * 6.0),
            ("ln_bwd_kernel", lambda: ops.ln_rows_bwd(y, do2, w["g2"], w["eps2"], dz, sc, zero[:D], zero[D:2 * D]), None, m * D * 14.0),
            ("ls_cast_kernel", lambda: ops.ls_cast_bwd(dz, w["ls2"], do2, sc, zero[:D]), None, m * D * 8.0),
        ]
        kms = {}
        for name, fnk, fl, by in stages:
            t = timed(fnk, 3, 1)
            kms[name] = {"ms": round(t, 4), **({"tflops": round(fl / t / 1e9, 1)} if fl else {"gbs": round(by / t / 1e6, 1)})}
    out["backward_kernels_one_layer"] = kms
    return out
