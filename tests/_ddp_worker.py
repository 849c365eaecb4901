"""Worker of tests/test_ddp_gpu.py, launched by `python -m torch.distributed.run --nproc-per-node 2` (one rank per GPU over
NCCL when the box has >= 2 GPUs; otherwise both ranks share cuda:0 and DDP all-reduces over gloo, which accepts CUDA tensors).

It wraps the drop-in module exactly like the reference (train_ddp_v3mod2.py:816-822: `torch.compile(model)` then
`DDP(model, device_ids=[local_rank], find_unused_parameters=False)`) and checks
  * DDP's start-up broadcast reaches the packed device copies (the ranks start from DIFFERENT random weights on purpose);
  * the all-reduced gradient of step 0 equals the single-process gradient on the concatenated batch;
  * after 3 optimizer steps every parameter is bit-identical on all ranks, and so are eval-mode outputs (packed weights in sync).
Modes:  reference  -- the literal reference sequence: fp16 autocast + GradScaler + clip_grad_norm_ + torch.optim.AdamW
                      (train_ddp_v3mod2.py:745, 854, 922-929), default gradient hand-off (copies)
        view       -- what bench.py --mode train runs: grad_handoff='view' + gradient_as_bucket_view + jat_b200.FusedAdamW
        view_bf16  -- the same with the bf16 gradient exchange (jat_b200.ddp.register_bf16_allreduce): the all-reduced gradient
                      equals the single-process one to bf16 rounding (< 5e-3 rel-L2), ranks stay bit-identical
        view_bf16_fused -- bf16 exchange consumed directly by FusedAdamW (no re-expansion; .grad keeps the local gradient):
                      after 3 steps the parameters equal those of the view_bf16 run to f32 rounding, ranks stay bit-identical
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist
import torch.nn.functional as F

CFG = dict(input_channels=32, cond_channels=32, patch_len=4, hidden_size=256, depth=4, num_q_heads=4, num_kv_heads=2,
           bottleneck_dim=128, mlp_ratio=4.0, dropout=0.0, drop_path_rate=0.0)


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def build(cls, seed, dev):
    import jat_b200
    torch.manual_seed(seed)
    m = getattr(jat_b200, cls)(**CFG)
    g = torch.Generator().manual_seed(seed + 100)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "adaLN_modulation.1" in n or n.startswith("final_layer.1"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return m.to(dev).train()


def main():
    mode, cls = sys.argv[1], sys.argv[2]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    shared = torch.cuda.device_count() < world
    local = 0 if shared else rank
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("gloo" if shared else "nccl")
    import jat_b200

    b, C, T = 3, 32, 170
    g = torch.Generator(device=dev).manual_seed(7)          # the same global batch on every rank; rank r trains on its slice
    hr, lr, eps = (torch.randn(world * b, C, T, generator=g, device=dev) for _ in range(3))
    t = torch.rand(world * b, generator=g, device=dev)
    z_t = t.view(-1, 1, 1) * hr + (1 - t.view(-1, 1, 1)) * eps
    mine = slice(rank * b, (rank + 1) * b)

    model = build(cls, seed=rank, dev=dev)                   # different weights per rank until DDP broadcasts rank 0's
    amp = mode == "reference"
    if mode.startswith("view"):
        model.grad_handoff = "view"
    net = torch.compile(model, mode="default", backend="inductor")                                    # :816
    net = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], find_unused_parameters=False,   # :822
                                                    gradient_as_bucket_view=mode.startswith("view"), bucket_cap_mb=1)
    if mode == "view_bf16":
        jat_b200.ddp.register_bf16_allreduce(net)
    fused_consumer = mode == "view_bf16_fused"
    # single-process truth on the concatenated batch, from the weights DDP has just broadcast
    ref = build(cls, seed=0, dev=dev)
    ref.load_state_dict(model.state_dict())
    assert all(torch.equal(p, q) for p, q in zip(model.parameters(), ref.parameters()))
    if rank != 0:
        m0 = build(cls, seed=0, dev=dev)
        assert all(torch.equal(p, q) for p, q in zip(model.parameters(), m0.parameters())), "start-up broadcast did not land"
        del m0

    def loss_of(module, sl):
        if amp:
            with torch.amp.autocast("cuda"):
                pred = module(z_t[sl], t[sl], lr[sl])
                assert pred.dtype == torch.float16
                return F.mse_loss(pred, hr[sl])
        return F.mse_loss(module(z_t[sl], t[sl], lr[sl]), hr[sl])

    scale = 65536.0 if amp else 1.0
    (loss_of(ref, slice(None)) * scale).backward()
    want = [p.grad.detach().clone() / scale for p in ref.parameters()]
    del ref

    if mode.startswith("view"):
        opt = jat_b200.FusedAdamW(model.parameters(), lr=2e-3, weight_decay=0.1, max_grad_norm=1.0, model=model)
        if fused_consumer:
            jat_b200.ddp.register_bf16_allreduce(net, optimizer=opt)
            # twin run on the SAME ranks with the expanding hook: what the parameters must come out as
            twin = build(cls, seed=0, dev=dev)
            twin.load_state_dict(model.state_dict())
            twin.grad_handoff = "view"
            twin_net = torch.nn.parallel.DistributedDataParallel(twin, device_ids=[local], find_unused_parameters=False,
                                                                 gradient_as_bucket_view=True, bucket_cap_mb=1)
            jat_b200.ddp.register_bf16_allreduce(twin_net)
            twin_opt = jat_b200.FusedAdamW(twin.parameters(), lr=2e-3, weight_decay=0.1, max_grad_norm=1.0, model=twin)
    else:
        opt = torch.optim.AdamW(model.parameters(), lr=2e-3, weight_decay=0.1)
    scaler = torch.amp.GradScaler("cuda", enabled=amp)
    losses, grad_err = [], None
    for step in range(3):
        opt.zero_grad(set_to_none=True)
        loss = loss_of(net, mine)
        scaler.scale(loss).backward()                        # :922 (+ DDP bucketed all-reduce, averaged over ranks)
        scaler.unscale_(opt)                                 # :925
        if step == 0 and not fused_consumer:
            num = sum((p.grad.double() - w.double()).pow(2).sum().item() for p, w in zip(model.parameters(), want))
            den = sum(w.double().pow(2).sum().item() for w in want)
            grad_err = (num / den) ** 0.5
            worst = max(rel_l2(p.grad, w) for p, w in zip(model.parameters(), want))
            tol = (5e-3, 2e-2) if mode == "view_bf16" else (2e-4, 5e-3)
            assert grad_err < tol[0] and worst < tol[1], (grad_err, worst)
            if mode == "view_bf16":
                assert grad_err > 1e-4, "the bf16 exchange did not take place"
        if not mode.startswith("view"):
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)   # :926 (FusedAdamW clips inside step())
        scaler.step(opt)                                     # :928
        scaler.update()
        losses.append(float(loss))
        if fused_consumer:
            twin_opt.zero_grad(set_to_none=True)
            F.mse_loss(twin_net(z_t[mine], t[mine], lr[mine]), hr[mine]).backward()
            twin_opt.step()
            # the two runs reduce the same gradients up to the run-to-run order of the f32 reduce-adds in the backward; an
            # element whose bf16 payload rounds the other way moves by one Adam step, so the bound loosens with the step count
            err = max(rel_l2(p, q) for p, q in zip(model.parameters(), twin.parameters()))
            assert err < (2e-4 if step == 0 else 5e-3), (step, err)
            assert abs(float(opt.grad_norm) - float(twin_opt.grad_norm)) <= 1e-3 * float(twin_opt.grad_norm)
            if step == 0:
                grad_err = err
    assert all(l == l for l in losses)

    def gathered(x):   # gloo gathers host tensors, NCCL device tensors
        x = x.cpu() if shared else x.contiguous()
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x)
        return parts

    both = gathered(torch.cat([p.detach().flatten() for p in model.parameters()]))
    assert all(torch.equal(both[0], x) for x in both[1:]), "parameters diverged across ranks"
    model.eval()
    with torch.no_grad():
        outs = gathered(model(z_t[:b], t[:b], lr[:b]))
    assert all(torch.equal(outs[0], x) for x in outs[1:]), "eval outputs differ across ranks (stale packed weights)"
    if rank == 0:
        print("DDP_WORKER_OK " + json.dumps({"mode": mode, "cls": cls, "backend": dist.get_backend(), "shared_gpu": shared,
                                               "grad_rel_l2_vs_single_process": grad_err, "losses": losses,
                                               "launches": int(jat_b200._lib.load().jat_launch_count(jat_b200._lib.context(local)))}),
              flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
