"""The module protocol of SURVEY.md 8b on the GPU: `.to(rank)` -> `AdamW(model.parameters())` -> `torch.compile(model)`
(train_ddp_v3mod2.py:816) -> `model.train()/eval()` -> fp16 `autocast` + `GradScaler.scale(loss).backward()` (:745, :854, :922)
-> `unscale_` + `clip_grad_norm_` (:925-926) -> `scaler.step` -- single process here, the DDP wrap is tests/test_ddp_gpu.py.
Plus the two host-side guards of the training state (single-slot saved activations)."""
import pytest
import torch
import torch.nn.functional as F

from tests._util import rerandomise_zero_init

pytestmark = pytest.mark.gpu
CFG = dict(input_channels=32, cond_channels=32, patch_len=4, hidden_size=256, depth=3, num_q_heads=4, num_kv_heads=2,
           bottleneck_dim=128, mlp_ratio=4.0, dropout=0.0, drop_path_rate=0.0)


def dev():
    return torch.device("cuda", 0)


def _data(B=3, T=170, seed=3):
    g = torch.Generator(device=dev()).manual_seed(seed)
    hr, lr, eps = (torch.randn(B, 32, T, generator=g, device=dev()) for _ in range(3))
    t = torch.rand(B, generator=g, device=dev())
    return t.view(B, 1, 1) * hr + (1 - t.view(B, 1, 1)) * eps, t, lr, hr


def _model(cls="JaT_AudioSR_V2"):
    import jat_b200
    torch.manual_seed(0)
    return rerandomise_zero_init(getattr(jat_b200, cls)(**CFG)).to(dev())


def test_reference_training_sequence_compile_autocast_gradscaler():
    model = _model().train()
    opt = torch.optim.AdamW(model.parameters(), lr=2e-3, weight_decay=0.1)
    scaler = torch.amp.GradScaler("cuda")
    net = torch.compile(model, mode="default", backend="inductor")
    assert all(k.startswith("_orig_mod.") for k in net.state_dict())     # what the reference's loaders strip
    z_t, t, lr, hr = _data()
    plain = _model().train()
    F.mse_loss(plain(z_t, t, lr), hr).backward()
    losses = []
    for step in range(6):
        opt.zero_grad(set_to_none=True)
        with torch.amp.autocast("cuda"):
            pred = net(z_t, t, lr)
            assert pred.dtype == torch.float16 and pred.requires_grad
            loss = F.mse_loss(pred, hr)
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        if step == 0:   # the unscaled gradients are the plain-module gradients up to the fp16 rounding of the prediction
            for (k, p), q in zip(model.named_parameters(), plain.parameters()):
                e = ((p.grad - q.grad).norm() / q.grad.norm().clamp_min(1e-20)).item()
                assert e < 2e-2, (k, e)
        gn = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        assert torch.isfinite(gn)
        scaler.step(opt)
        scaler.update()
        losses.append(loss.item())
    assert scaler.get_scale() == 65536.0          # no step was skipped for inf / nan gradients
    assert losses[-1] < 0.9 * losses[0], losses
    net.eval()
    with torch.no_grad():
        out = net(z_t, t, lr)
    assert out.dtype == torch.float32 and abs(F.mse_loss(out, hr).item() - losses[-1]) < 0.2 * losses[-1]


def test_second_train_forward_before_backward_raises_instead_of_wrong_gradients():
    from jat_b200 import _lib as L
    model = _model("JaT_AudioSR_V3").train()
    z_t, t, lr, hr = _data()
    first = F.mse_loss(model(z_t, t, lr), hr)
    second = F.mse_loss(model(z_t * 0.5, t, lr), hr)       # overwrites the single-slot saved activations
    with pytest.raises(L.JatError, match="stale forward"):
        first.backward()
    model.zero_grad(set_to_none=True)
    second.backward()                                      # the latest forward is still consistent
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


def test_eval_forward_between_train_forward_and_backward_is_harmless():
    """A no-grad pass of the SAME shape between forward and backward (EMA / teacher / self-conditioning) used to share the
    workspace the backward reads (mod, x, h, patches, ...); the training state now owns its workspace."""
    model = _model().train()
    z_t, t, lr, hr = _data()
    F.mse_loss(model(z_t, t, lr), hr).backward()
    want = [p.grad.clone() for p in model.parameters()]
    model.zero_grad(set_to_none=True)
    loss = F.mse_loss(model(z_t, t, lr), hr)
    with torch.no_grad():
        model.eval()
        model(torch.randn_like(z_t), 1.0 - t, torch.randn_like(lr))
        model.train()
        model(torch.randn_like(z_t), 1.0 - t, torch.randn_like(lr))    # train mode without grad = plain forward, too
    loss.backward()
    for p, w in zip(model.parameters(), want):
        # (run-to-run differences: f32 reduce-add order of the split-K weight gradients and the column sums; a clobbered
        #  workspace gives errors of order one)
        assert ((p.grad - w).norm() / w.norm().clamp_min(1e-20)).item() < 2e-4


def test_sampler_reuses_plan_and_graph_with_unindexed_device():
    import jat_b200
    model = _model().eval()
    g = torch.Generator(device=dev()).manual_seed(1)
    lr = torch.randn(2, 32, 170, generator=g, device=dev())
    z0 = torch.randn(2, 32, 170, generator=g, device=dev())
    a = jat_b200.flow_matching_sample(model, lr, num_steps=4, cfg_scale=3.0, device="cuda", verbose=False, z0=z0)
    packed = model._engine.packed
    plan = next(iter(model.__dict__["_sampler_plans"].values()))
    graph = plan.graph
    assert graph is not None
    b = jat_b200.flow_matching_sample(model, lr, num_steps=4, cfg_scale=3.0, device="cuda", verbose=False, z0=z0)
    c = jat_b200.flow_matching_sample(model, lr, num_steps=4, cfg_scale=3.0, device=torch.device("cuda", 0), verbose=False, z0=z0)
    assert torch.equal(a, b) and torch.equal(a, c)
    assert model._engine.packed is packed
    assert next(iter(model.__dict__["_sampler_plans"].values())) is plan and plan.graph is graph
    with torch.no_grad():   # the plain forward path resolves 'cuda:0' tensors to the same packed copies
        model(z0, torch.rand(2, device=dev()), lr)
    assert model._engine.packed is packed
