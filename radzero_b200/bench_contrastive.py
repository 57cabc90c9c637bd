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


def config(n_total: int):
    """The `config` object of the contrastive workload -- identical in both arms of bench.py."""
    return {"workload": f"C4 contrastive step, {B_GLOBAL} images x {n_total} sentences (n_i ~ U{{3..9}}), "
                        "image batch sharded over the GPUs with text all-gather",
            "input_dtype": "bf16", "tokens": L, "hidden": D}


def _one_rank_record():
    import json
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "contrastive_1rank.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def _key_phrases(counts, dev):
    """Per-image tokenised sentence batches as the reference's dataset hands them over
    (dataset.py:172-181); the ids index the rows of the precomputed sentence embeddings."""
    out, o = [], 0
    ids_all = torch.arange(sum(counts), device=dev).view(-1, 1)
    ones = torch.ones_like(ids_all)
    for c in counts:
        out.append({"input_ids": ids_all[o:o + c], "attention_mask": ones[o:o + c]})
        o += c
    return out


def run(args, world, rank, local, pk, steps=None, warmup=None, quiet=False):
    import torch.distributed as dist
    from radzero_b200 import _lib, losses, ops, training
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
    b_local = B_GLOBAL // world
    counts = torch.bincount(gm - rank * b_local, minlength=b_local).tolist()

    def step(tk, tx):
        tk = tk.detach().requires_grad_(True)
        tx = tx.detach().requires_grad_(True)
        fn.zero_grad(set_to_none=True)
        res = training.contrastive_step(fn, tx, gm, tk, distributed=distributed)
        res["loss"].backward()
        return res["loss"].detach(), tk.grad, tx.grad

    # the same step through the reference surface: RadZeroLoss.forward(key_phrases, vision_tokens,
    # forward_text_model) (losses.py:71-124) -- per-image sentence batches, the text model as a callback
    key_phrases = _key_phrases(counts, dev)
    holder = {}

    def text_model(enc):
        f = holder["text"][enc["input_ids"][:, 0]] if enc["input_ids"].shape[0] != holder["text"].shape[0] \
            else holder["text"]
        return {"text_features_wo_l2_norm": f, "text_features": f}

    def surface_step(tk, tx):
        tk = tk.detach().requires_grad_(True)
        holder["text"] = tx.detach().requires_grad_(True)
        fn.zero_grad(set_to_none=True)
        out = fn(key_phrases, tk, text_model)            # ddp_gather=True: sharded when torch.distributed is up
        out["losses"]["loss"].backward()
        return out["losses"]["loss"].detach()

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        loss, dtok, dtxt = step(tok, text)
    # loss / gradient fingerprint of THIS global problem: the same numbers at every rank count
    chk = torch.stack([dtok.double().abs().sum(), dtxt.double().abs().sum()])
    if distributed:
        dist.all_reduce(chk)
    # under torch.distributed the node multiplies gradients by W (DDP averages them afterwards): undo it
    grad_checksum = float(chk.sum().item()) / (world if distributed else 1)
    del dtok, dtxt
    barrier()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss, _, _ = step(tok, text)
    e1.record()
    barrier()
    launches = _lib.launch_count() - n0
    ms = e0.elapsed_time(e1)
    if distributed:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_step = ms / steps
    # the same step through RadZeroLoss.forward (per-image sentence batches merged on the host), device-resident
    for _ in range(2):
        surface_step(tok, text)
    barrier()
    e0.record()
    ns = max(3, steps // 2)
    for _ in range(ns):
        surface_step(tok, text)
    e1.record()
    barrier()
    ms_surface = e0.elapsed_time(e1) / ns
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

    # every step's loss is copied to pinned host memory inside the timed region; as a training loop that logs
    # its loss does, the host reads step i's value while step i + 1 is already queued (no pipeline drain)
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [None, None]

    def e2e_loop(n):
        last = None
        ev = upload(0)
        for i in range(n):
            k = i & 1
            main.wait_event(ev)
            if i + 1 < n:
                ev = upload(k ^ 1)
            l = surface_step(d_tok[k], d_txt[k])
            loss_host[k].copy_(l, non_blocking=True)          # D2H read of the step's loss
            loss_ready[k] = torch.cuda.Event()
            loss_ready[k].record(main)
            used[k] = loss_ready[k]
            if i > 0:                                        # the previous step's loss, now on the host
                loss_ready[k ^ 1].synchronize()
                last = float(loss_host[k ^ 1])
        loss_ready[(n - 1) & 1].synchronize()
        return float(loss_host[(n - 1) & 1])

    e2e_loop(2)
    barrier()
    ks = max(4, min(steps, 8))
    e0.record()
    l_host = e2e_loop(ks)
    e1.record()
    barrier()
    ms2 = e0.elapsed_time(e1)
    if distributed:
        t = torch.tensor([ms2], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms2 = float(t.item())
    del d_tok, d_txt

    def dev_time(f, reps=10):
        """Device time per call of a few-microsecond operation: the launches are queued behind a ~20 ms
        spin kernel so that the host's launch overhead (Python + ctypes, tens of us) is not what is timed."""
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        torch.cuda._sleep(40_000_000)
        e0.record()
        for _ in range(reps):
            f()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # the MP-NCE loss kernels on their own (K10): algorithmic bytes = read Z + write dZ
    zt = torch.randn(n_total, b_local, device=dev) * 0.3
    gm_all = gm if not distributed else torch.arange(n_total, device=dev) % B_GLOBAL
    lt = fn.loss_temperature.detach()

    def nce():
        rs, ps, cn, cp = ops.mpnce_partials(zt, gm_all, rank * b_local, log_tau=lt, b_global=B_GLOBAL)
        return ops.mpnce_finish(zt, gm_all, rank * b_local, B_GLOBAL, 1.0, rs, ps, cn, cp, log_tau=lt)
    n1 = _lib.launch_count()
    nce()
    nce_launches = _lib.launch_count() - n1
    nce_ms = dev_time(nce)
    rs_, ps_, cn_, cp_ = ops.mpnce_partials(zt, gm_all, rank * b_local, log_tau=lt, b_global=B_GLOBAL)
    nce_parts = {"partials_us": round(dev_time(lambda: ops.mpnce_partials(zt, gm_all, rank * b_local, log_tau=lt, b_global=B_GLOBAL)) * 1e3, 1),
                 "finish_us": round(dev_time(lambda: ops.mpnce_finish(zt, gm_all, rank * b_local, B_GLOBAL, 1.0, rs_, ps_,
                                                                      cn_, cp_, log_tau=lt)) * 1e3, 1)}
    nce_bytes = 2 * n_total * b_local * 4
    mpnce = {"ms": nce_ms, "launches": int(nce_launches), **nce_parts, "algorithmic_bytes": nce_bytes,
             "achieved_gbs": nce_bytes / (nce_ms * 1e-3) / 1e9,
             "frac_of_hbm": nce_bytes / (nce_ms * 1e-3) / 1e9 / pk["hbm"],
             "note": "two persistent cooperative launches (partials | coefficients + dZ + terms); Z is 25 MB at C4 "
                     "and is read from L2; the fraction is of the HBM copy peak on read-Z + write-dZ bytes"}
    # the step's collectives on their own (same message sizes, back to back): device time per step
    comm_us = None
    if distributed:
        cap = -(-n_total // world) + 64
        g_send = torch.empty(cap * (D * 2 + 8), dtype=torch.uint8, device=dev)
        g_recv = torch.empty(world * g_send.numel(), dtype=torch.uint8, device=dev)
        rowpos = torch.empty(2 * n_total, device=dev)
        scal = torch.empty(1, device=dev)
        dq_send = torch.empty(world * cap, D, device=dev)
        dq_out = torch.empty(cap, D, device=dev)
        comm_us = {
            "all_gather_text": round(dev_time(lambda: dist.all_gather_into_tensor(g_recv, g_send)) * 1e3, 1),
            "all_reduce_rows": round(dev_time(lambda: dist.all_reduce(rowpos)) * 1e3, 1),
            "all_reduce_loss": round(dev_time(lambda: dist.all_reduce(scal)) * 1e3, 1),
            "reduce_scatter_dq": round(dev_time(lambda: dist.reduce_scatter_tensor(dq_out, dq_send)) * 1e3, 1),
        }
        comm_us["total"] = round(sum(comm_us.values()), 1)
        comm_us["frac_of_step"] = round(comm_us["total"] * 1e-3 / ms_step, 4)
    one = _one_rank_record()
    match = None
    if one is not None:
        match = bool(abs(float(loss.item()) - one["loss"]) <= 1e-5 * abs(one["loss"])
                     and abs(grad_checksum - one["grad_checksum"]) <= 2e-3 * abs(one["grad_checksum"]))
    out = {
        "metric": "contrastive steps/sec", "value": 1e3 / ms_step, "unit": "steps/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms_step, "scaling": "strong", "dtype": "f16",
        "loss": float(loss.item()), "grad_checksum": grad_checksum, "matches_1rank": match,
        "one_rank_record": one, "gpu_launches": int(launches),
        "config": config(n_total),
        "notes": {"parallelism": f"images sharded x{world}, text all-gather, dq reduce-scatter"},
        "roofline": {"bound": "tensor", "kernel": "whole step (sim_fwd + rz_sim_bwd GEMM passes)",
                     "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"],
                     "algorithmic_flops_per_step": flops, "traffic": _traffic(),
                     "peak_source": pk["src"] + " sustained (kernel timed inside a long step)"},
        "e2e": {"value": 1e3 / (ms2 / ks), "unit": "steps/s",
                "h2d_bytes_per_step": (h_tok.numel() + h_txt.numel()) * 2, "d2h_bytes_per_step": 4,
                "api": "RadZeroLoss.forward(key_phrases, vision_tokens, forward_text_model) + loss.backward()",
                "steps": ks, "loss": l_host,
                "note": "inputs of step i+1 are uploaded on a copy stream while step i computes; each step's loss is "
                        "copied to pinned host memory and read one step later"},
        "surface": {"api": "RadZeroLoss.forward + backward, device-resident", "ms_per_step": ms_surface},
        "mpnce": mpnce, "comm_us": comm_us,
    }
    return out
