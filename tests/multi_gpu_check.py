"""Multi-GPU equivalence check of the image-sharded contrastive step (run under torchrun):

    python -m torch.distributed.run --nproc-per-node W --master-addr 127.0.0.1 tests/multi_gpu_check.py

Every rank builds the SAME small global problem from one seed, keeps its shard of the images
(and their sentences), and runs the distributed fused step over NCCL.  Rank 0 also runs the
single-GPU step on the whole problem; loss and gradients must agree (SURVEY.md section 8e:
same loss and gradients up to summation order)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from radzero_b200 import losses, synthetic, training  # noqa: E402


def main():
    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, L = 4 * world, 300
    counts = synthetic.sentence_counts(B, seed=5, lo=20, hi=40)
    tok, text, gamma, beta, log_tau = synthetic.make_inputs(B, sum(counts), tokens_per_image=L, seed=77)
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    bl = B // world
    i0, i1 = rank * bl, (rank + 1) * bl

    def run_mode(sim_op):
        def make_fn():
            fn = losses.RadZeroLoss(sim_op=sim_op).to(dev)
            with torch.no_grad():
                fn.layer_norm.weight.copy_(gamma)
                fn.layer_norm.bias.copy_(beta)
            return fn

        fn = make_fn()
        tk = tok[i0:i1].to(dev).requires_grad_(True)
        tx = text[offs[i0]:offs[i1]].to(dev).requires_grad_(True)
        gm = synthetic.group_map_from_counts(counts[i0:i1], first_image=i0, device=dev)
        res = training.contrastive_step(fn, tx, gm, tk, distributed=True, ddp_compatible=False)
        res["loss"].backward()
        # gather the sharded results on rank 0
        tok_g = [torch.empty_like(tk.grad) for _ in range(world)]
        dist.all_gather(tok_g, tk.grad.contiguous())
        pg = torch.cat([fn.layer_norm.weight.grad, fn.layer_norm.bias.grad, fn.loss_temperature.grad])
        dist.all_reduce(pg)
        sizes = [offs[(r + 1) * bl] - offs[r * bl] for r in range(world)]
        nmax = max(sizes)
        pad = torch.zeros(nmax, 768, device=dev)
        pad[: tx.grad.shape[0]] = tx.grad
        txt_g = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(txt_g, pad)
        ok = True
        if rank == 0:
            fn1 = make_fn()
            tk1 = tok.to(dev).requires_grad_(True)
            tx1 = text.to(dev).requires_grad_(True)
            gm1 = synthetic.group_map_from_counts(counts, device=dev)
            r1 = training.contrastive_step(fn1, tx1, gm1, tk1, distributed=False)
            r1["loss"].backward()
            tg = torch.cat(tok_g)
            xg = torch.cat([txt_g[r][: sizes[r]] for r in range(world)])
            pg1 = torch.cat([fn1.layer_norm.weight.grad, fn1.layer_norm.bias.grad, fn1.loss_temperature.grad])

            def rel(a, b):
                return float((a - b).abs().max() / b.abs().max())
            e_loss = abs(res["loss"].item() - r1["loss"].item()) / abs(r1["loss"].item())
            e_tok, e_txt, e_par = rel(tg, tk1.grad), rel(xg, tx1.grad), rel(pg, pg1)
            print(f"sim_op={sim_op} world={world} loss {res['loss'].item():.6f} vs {r1['loss'].item():.6f} rel {e_loss:.2e}; "
                  f"grad rel err tokens {e_tok:.2e} text {e_txt:.2e} params {e_par:.2e}")
            ok = e_loss < 1e-5 and e_tok < 2e-3 and e_txt < 2e-3 and e_par < 2e-3
        return ok

    # both similarity operators: "cos" (radzero.yaml) and "dot" (the constructor default, losses.py:45)
    ok = all([run_mode("cos"), run_mode("dot")])
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
