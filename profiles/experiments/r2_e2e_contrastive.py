"""Where does the contrastive e2e step lose its 30 ms?  Variants of the e2e loop of bench_contrastive."""
import sys, os, time, torch
sys.path.insert(0, os.getcwd())
from radzero_b200 import bench_contrastive as bc, losses, training
dev = torch.device("cuda", 0)
tok, text, gamma, beta, gm, n_total = bc._inputs(0, 1, dev, torch.bfloat16)
fn = losses.RadZeroLoss(sim_op="cos").to(dev)
with torch.no_grad():
    fn.layer_norm.weight.copy_(gamma); fn.layer_norm.bias.copy_(beta)
def step(tk, tx):
    tk = tk.detach().requires_grad_(True); tx = tx.detach().requires_grad_(True)
    fn.zero_grad(set_to_none=True)
    res = training.contrastive_step(fn, tx, gm, tk, distributed=False)
    res["loss"].backward()
    return res["loss"].detach()
h_tok, h_txt = tok.cpu().pin_memory(), text.cpu().pin_memory()
d_tok = [torch.empty_like(tok) for _ in range(2)]; d_txt = [torch.empty_like(text) for _ in range(2)]
copy_stream = torch.cuda.Stream(device=dev); main = torch.cuda.current_stream(dev)
for _ in range(3): step(tok, text)
torch.cuda.synchronize()
def timeit(name, f, n=6):
    f(2); torch.cuda.synchronize(); t0 = time.perf_counter(); f(n); torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / n * 1e3:.1f} ms/step", flush=True)
def resident(n):
    for i in range(n): step(tok, text)
def copy_only(n):
    for i in range(n):
        with torch.cuda.stream(copy_stream):
            d_tok[i & 1].copy_(h_tok, non_blocking=True); d_txt[i & 1].copy_(h_txt, non_blocking=True)
    copy_stream.synchronize()
def overlapped(n):
    used = [None, None]
    def upload(k):
        with torch.cuda.stream(copy_stream):
            if used[k] is not None: copy_stream.wait_event(used[k])
            d_tok[k].copy_(h_tok, non_blocking=True); d_txt[k].copy_(h_txt, non_blocking=True)
            ev = torch.cuda.Event(); ev.record(copy_stream)
        return ev
    ev = upload(0)
    for i in range(n):
        k = i & 1
        main.wait_event(ev)
        if i + 1 < n: ev = upload(k ^ 1)
        step(d_tok[k], d_txt[k])
        used[k] = torch.cuda.Event(); used[k].record(main)
def serial(n):
    for i in range(n):
        d_tok[0].copy_(h_tok, non_blocking=True); d_txt[0].copy_(h_txt, non_blocking=True)
        step(d_tok[0], d_txt[0])
timeit("resident", resident); timeit("copy only (2.16 GB)", copy_only); timeit("serial copy+step", serial); timeit("overlapped", overlapped)

# ---- the same loop through the reference surface (RadZeroLoss.forward with per-image key phrases)
b_local = bc.B_GLOBAL
counts = torch.bincount(gm, minlength=b_local).tolist()
key_phrases = bc._key_phrases(counts, dev)
holder = {}
def text_model(enc):
    return {"text_features_wo_l2_norm": holder["text"], "text_features": holder["text"]}
def surface_step(tk, tx):
    tk = tk.detach().requires_grad_(True)
    holder["text"] = tx.detach().requires_grad_(True)
    fn.zero_grad(set_to_none=True)
    out = fn(key_phrases, tk, text_model)
    out["losses"]["loss"].backward()
    return out["losses"]["loss"].detach()
def make_loop(stepfn, lagged_sync):
    def loop(n):
        used = [None, None]
        def upload(k):
            with torch.cuda.stream(copy_stream):
                if used[k] is not None: copy_stream.wait_event(used[k])
                d_tok[k].copy_(h_tok, non_blocking=True); d_txt[k].copy_(h_txt, non_blocking=True)
                ev = torch.cuda.Event(); ev.record(copy_stream)
            return ev
        ev = upload(0)
        for i in range(n):
            k = i & 1
            main.wait_event(ev)
            if i + 1 < n: ev = upload(k ^ 1)
            t0 = time.perf_counter()
            stepfn(d_tok[k], d_txt[k])
            cpu_ms.append((time.perf_counter() - t0) * 1e3)
            used[k] = torch.cuda.Event(); used[k].record(main)
            if lagged_sync and i > 0: used[k ^ 1].synchronize()
    return loop
cpu_ms = []
for _ in range(2): surface_step(tok, text)
torch.cuda.synchronize()
timeit("overlapped, surface", make_loop(surface_step, False)); print("   host ms per surface_step:", [round(x, 1) for x in cpu_ms[-6:]])
cpu_ms = []
timeit("overlapped, surface, lagged sync", make_loop(surface_step, True)); print("   host ms:", [round(x, 1) for x in cpu_ms[-6:]])
cpu_ms = []
timeit("overlapped, direct, lagged sync", make_loop(step, True)); print("   host ms:", [round(x, 1) for x in cpu_ms[-6:]])
