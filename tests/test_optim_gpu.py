"""GPU parity of the fused parameter update (`jat_b200.FusedAdamW` = `jat_grad_sumsq` + `jat_adamw_step`) against what
the reference's training loop runs (train_ddp_v3mod2.py:709, 926-928): `clip_grad_norm_(params, 1.0)` followed by
`torch.optim.AdamW.step()`.  Stated tolerance: fp32 round-off of one update (the kernels use ATen's operand types; the
only reordering is the summation order inside the gradient norm): |dp| <= 1e-6 + 2e-6 |p| after every step."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda", 0)


SHAPES = [(1280, 1280), (5120,), (7,), (3, 5), (4099,), (257, 33), (1,), (8192, 64)]


def _clone_params(params):
    return [torch.nn.Parameter(p.detach().clone()) for p in params]


def _close(a, b, what):
    err = (a - b).abs()
    tol = 1e-6 + 2e-6 * b.abs()
    assert bool((err <= tol).all()), (what, float(err.max()))


@pytest.mark.parametrize("max_norm", [None, 1.0, 1e4])
@pytest.mark.parametrize("weight_decay", [0.0, 0.1])
def test_fused_adamw_matches_torch_clip_plus_adamw(max_norm, weight_decay):
    import jat_b200
    g = torch.Generator(device=dev()).manual_seed(5)
    mine = [torch.nn.Parameter(torch.randn(s, generator=g, device=dev())) for s in SHAPES]
    # an unaligned parameter (4-byte offset into its storage): exercises the scalar path
    base = torch.randn(1001, generator=g, device=dev())
    mine.append(torch.nn.Parameter(base[1:]))
    ref = _clone_params(mine)
    kw = dict(lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=weight_decay)
    opt = jat_b200.FusedAdamW(mine, max_grad_norm=max_norm, **kw)
    ropt = torch.optim.AdamW(ref, fused=True, **kw)
    for step in range(6):
        for p, q in zip(mine, ref):
            gr = torch.randn(p.shape, generator=g, device=dev()) * (10.0 if step % 2 else 0.01)
            p.grad, q.grad = gr.clone(), gr.clone()
        v0 = [p._version for p in mine]
        g_before = [p.grad.clone() for p in mine]
        if max_norm is not None:
            want_norm = torch.nn.utils.clip_grad_norm_(ref, max_norm)
        ropt.step()
        opt.step()
        assert all(p._version > v for p, v in zip(mine, v0))           # autograd sees the in-place update
        if max_norm is not None:
            assert abs(float(opt.grad_norm) - float(want_norm)) <= 1e-5 * float(want_norm)
            assert all(torch.equal(p.grad, gb) for p, gb in zip(mine, g_before))  # p.grad is left unscaled
        for i, (p, q) in enumerate(zip(mine, ref)):
            _close(p.detach(), q.detach(), ("param", step, i))
            _close(opt.state[p]["exp_avg"], ropt.state[q]["exp_avg"], ("exp_avg", step, i))
            _close(opt.state[p]["exp_avg_sq"], ropt.state[q]["exp_avg_sq"], ("exp_avg_sq", step, i))
            assert float(opt.state[p]["step"]) == step + 1


def test_fused_adamw_state_dict_interchanges_with_torch_adamw():
    """Same state layout as torch.optim.AdamW: checkpoints move both ways mid-run and the runs stay together."""
    import jat_b200
    g = torch.Generator(device=dev()).manual_seed(6)
    a = [torch.nn.Parameter(torch.randn(s, generator=g, device=dev())) for s in SHAPES[:4]]
    b = _clone_params(a)
    kw = dict(lr=1e-2, weight_decay=0.05)
    fused, stock = jat_b200.FusedAdamW(a, **kw), torch.optim.AdamW(b, **kw)

    def both_step(oa, ob):
        for p, q in zip(a, b):
            gr = torch.randn(p.shape, generator=g, device=dev())
            p.grad, q.grad = gr.clone(), gr.clone()
        oa.step()
        ob.step()
    for _ in range(3):
        both_step(fused, stock)
    # swap the optimizers' states through state_dict(), continue, compare
    fused2, stock2 = jat_b200.FusedAdamW(a, **kw), torch.optim.AdamW(b, **kw)
    fused2.load_state_dict(stock.state_dict())
    stock2.load_state_dict(fused.state_dict())
    for _ in range(3):
        both_step(fused2, stock2)
    for p, q in zip(a, b):
        _close(p.detach(), q.detach(), "param after swap")
        assert float(fused2.state[p]["step"]) == float(stock2.state[q]["step"]) == 6.0


def test_fused_adamw_refreshes_the_packed_weights_of_the_drop_in_module():
    """With `model=` the same pass writes the bf16 / f32 copies the GEMMs read: after a step they equal a fresh cast of
    the parameters bit for bit, the engine does not re-pack, and training tracks the stock clip + AdamW loop."""
    import jat_b200
    from jat_b200.engine import PackedWeights
    cfg = dict(input_channels=32, cond_channels=32, patch_len=4, hidden_size=128, depth=2, num_q_heads=2, num_kv_heads=1,
               bottleneck_dim=128, mlp_ratio=2.0, dropout=0.0, drop_path_rate=0.0)
    torch.manual_seed(0)
    model = jat_b200.JaT_AudioSR_V3(**cfg).to(dev()).train()
    g = torch.Generator(device=dev()).manual_seed(3)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "adaLN_modulation.1" in n or n.startswith("final_layer.1"):
                p.copy_(torch.randn(p.shape, generator=g, device=dev()) * 0.02)
    twin = jat_b200.JaT_AudioSR_V3(**cfg).to(dev()).train()
    twin.load_state_dict(model.state_dict())
    B, T = 4, 86
    hr, lr, eps = (torch.randn(B, 32, T, generator=g, device=dev()) for _ in range(3))
    t = torch.rand(B, generator=g, device=dev())
    z_t = t.view(B, 1, 1) * hr + (1 - t.view(B, 1, 1)) * eps
    kw = dict(lr=2e-3, weight_decay=0.1)
    opt = jat_b200.FusedAdamW(model.parameters(), max_grad_norm=1.0, model=model, **kw)
    ropt = torch.optim.AdamW(twin.parameters(), **kw)
    losses = []
    for step in range(6):
        opt.zero_grad(set_to_none=True)
        ropt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(model(z_t, t, lr), hr)
        loss.backward()
        rloss = torch.nn.functional.mse_loss(twin(z_t, t, lr), hr)
        rloss.backward()
        losses.append((loss.item(), rloss.item()))
        torch.nn.utils.clip_grad_norm_(twin.parameters(), 1.0)
        ropt.step()
        opt.step()
        pk = model._engine.packed
        assert pk is not None and not pk.stale(model, dev())           # nothing left to re-cast
        for dst, src in zip(pk._dst, pk._src):
            assert torch.equal(dst, src.detach().to(dst.dtype))        # the copy IS the updated parameter
        fresh = PackedWeights(model, dev())
        for k in ("wqkv", "w1", "w2", "wo"):
            for x, y in zip(pk.keep[k], fresh.keep[k]):
                assert torch.equal(x, y)
        assert model._engine.weights(dev()) is pk
    assert losses[-1][0] < 0.95 * losses[0][0], losses
    for l_mine, l_ref in losses:
        assert abs(l_mine - l_ref) <= 2e-3 * max(1.0, l_ref), losses
    num = sum(float(((p.detach() - q.detach()) ** 2).sum()) for p, q in zip(model.parameters(), twin.parameters()))
    den = sum(float((q.detach() ** 2).sum()) for q in twin.parameters())
    assert (num / den) ** 0.5 < 2e-3   # two bf16-forward training runs, 6 steps apart from round-off in the gradients


def test_fused_adamw_param_groups_share_one_clip_norm():
    """Two parameter groups (no weight decay on the 1-D tensors, another lr): one norm over both, per-group hyper-parameters,
    lr changed between steps as a scheduler would."""
    import jat_b200
    g = torch.Generator(device=dev()).manual_seed(7)
    mine = [torch.nn.Parameter(torch.randn(s, generator=g, device=dev())) for s in SHAPES]
    ref = _clone_params(mine)

    def groups(ps):
        return [dict(params=[p for p in ps if p.dim() > 1], weight_decay=0.1),
                dict(params=[p for p in ps if p.dim() <= 1], weight_decay=0.0, lr=5e-3)]
    opt = jat_b200.FusedAdamW(groups(mine), lr=1e-3, max_grad_norm=0.5)
    ropt = torch.optim.AdamW(groups(ref), lr=1e-3)
    for step in range(4):
        for o in (opt, ropt):
            o.param_groups[0]["lr"] = 1e-3 * (step + 1) / 4            # warm-up, train_ddp_v3mod2.py:712-717
        for p, q in zip(mine, ref):
            gr = torch.randn(p.shape, generator=g, device=dev())
            p.grad, q.grad = gr.clone(), gr.clone()
        want_norm = torch.nn.utils.clip_grad_norm_(ref, 0.5)
        ropt.step()
        opt.step()
        assert abs(float(opt.grad_norm) - float(want_norm)) <= 1e-5 * float(want_norm)
        for i, (p, q) in enumerate(zip(mine, ref)):
            _close(p.detach(), q.detach(), ("param", step, i))


def test_fused_adamw_skips_parameters_without_gradient():
    import jat_b200
    g = torch.Generator(device=dev()).manual_seed(8)
    mine = [torch.nn.Parameter(torch.randn(s, generator=g, device=dev())) for s in SHAPES[:4]]
    ref = _clone_params(mine)
    opt, ropt = jat_b200.FusedAdamW(mine, lr=1e-2), torch.optim.AdamW(ref, lr=1e-2)
    for step in range(3):
        for i, (p, q) in enumerate(zip(mine, ref)):
            if i == 1 and step < 2:          # frozen for the first two steps
                p.grad = q.grad = None
                continue
            gr = torch.randn(p.shape, generator=g, device=dev())
            p.grad, q.grad = gr.clone(), gr.clone()
        ropt.step()
        opt.step()
        for p, q in zip(mine, ref):
            _close(p.detach(), q.detach(), ("param", step))
    assert float(opt.state[mine[1]]["step"]) == 1.0 and float(opt.state[mine[0]]["step"]) == 3.0


@pytest.mark.parametrize("handoff", ["copy", "view"])
def test_grad_scaler_loop_as_in_the_reference(handoff):
    """The reference's step (train_ddp_v3mod2.py:922-930): scaler.scale(loss).backward(); scaler.unscale_(optimizer);
    clip_grad_norm_; scaler.step(optimizer); scaler.update() -- with FusedAdamW in place of AdamW + clip, against the stock
    sequence on a twin model.  An overflowing step (inf in the loss scale path) must be skipped by both."""
    import jat_b200
    cfg = dict(input_channels=32, cond_channels=32, patch_len=4, hidden_size=128, depth=2, num_q_heads=2, num_kv_heads=1,
               bottleneck_dim=128, mlp_ratio=2.0, dropout=0.0, drop_path_rate=0.0)
    torch.manual_seed(0)
    model = jat_b200.JaT_AudioSR_V2(**cfg).to(dev()).train()
    g = torch.Generator(device=dev()).manual_seed(3)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "adaLN_modulation.1" in n or n.startswith("final_layer.1"):
                p.copy_(torch.randn(p.shape, generator=g, device=dev()) * 0.02)
    twin = jat_b200.JaT_AudioSR_V2(**cfg).to(dev()).train()
    twin.load_state_dict(model.state_dict())
    model.grad_handoff = handoff
    B, T = 4, 86
    hr, lr, eps = (torch.randn(B, 32, T, generator=g, device=dev()) for _ in range(3))
    t = torch.rand(B, generator=g, device=dev())
    z_t = t.view(B, 1, 1) * hr + (1 - t.view(B, 1, 1)) * eps
    kw = dict(lr=2e-3, weight_decay=0.1)
    opt = jat_b200.FusedAdamW(model.parameters(), max_grad_norm=1.0, model=model, **kw)
    ropt = torch.optim.AdamW(twin.parameters(), **kw)
    sc, rsc = torch.amp.GradScaler("cuda", init_scale=1024.0), torch.amp.GradScaler("cuda", init_scale=1024.0)
    for step in range(5):
        poison = float("inf") if step == 2 else 1.0          # one overflowing step: both scalers must skip it
        opt.zero_grad(set_to_none=True)
        ropt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(model(z_t, t, lr), hr) * poison
        sc.scale(loss).backward()
        sc.unscale_(opt)
        sc.step(opt)
        sc.update()
        rloss = torch.nn.functional.mse_loss(twin(z_t, t, lr), hr) * poison
        rsc.scale(rloss).backward()
        rsc.unscale_(ropt)
        torch.nn.utils.clip_grad_norm_(twin.parameters(), 1.0)
        rsc.step(ropt)
        rsc.update()
        assert sc.get_scale() == rsc.get_scale()
        if step != 2:
            assert abs(loss.item() - rloss.item()) <= 2e-3 * max(1.0, rloss.item())
    assert sc.get_scale() == 512.0                            # halved once by the skipped step
    num = sum(float(((p.detach() - q.detach()) ** 2).sum()) for p, q in zip(model.parameters(), twin.parameters()))
    den = sum(float((q.detach() ** 2).sum()) for q in twin.parameters())
    assert (num / den) ** 0.5 < 2e-3
    assert float(opt.state[next(iter(model.parameters()))]["step"]) == 4.0   # 5 iterations, one skipped
