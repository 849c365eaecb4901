"""GPU parity of the training step (BASELINE config C4): forward in train mode + backward
through the C-ABI (`jat_dit_forward_train` / `jat_dit_backward`) against parameter gradients of the UNMODIFIED reference
(fixtures generated on CPU in fp32 by tests/golden/make_golden_grads.py; dropout = drop_path = 0).

Stated tolerance (bf16 operands, fp32 accumulation / statistics / residual stream and its gradient): global gradient
rel-L2 and every parameter's rel-L2 <= 2x what torch's own bf16 autocast makes on the reference for the same step
(tests/golden/bf16_autocast_yardstick.json["grads"]: global 0.54 %, worst parameter 1.2 %)."""
import ast
import json
import os

import numpy as np
import pytest
import torch

from tests._util import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda", 0)


def load(tag):
    d = np.load(os.path.join(GOLDEN, f"micro_{tag}_grads.npz"), allow_pickle=False)
    cfg = ast.literal_eval(str(d["cfg_json"]))
    return d, cfg


@pytest.mark.parametrize("tag,cls", [("v2_layernorm", "JaT_AudioSR_V2"), ("v3_rmsnorm", "JaT_AudioSR_V3")])
def test_training_step_gradients_match_reference(tag, cls):
    import jat_b200
    d, cfg = load(tag)
    yard = json.load(open(os.path.join(GOLDEN, "bf16_autocast_yardstick.json")))["grads"][tag]
    model = getattr(jat_b200, cls)(**cfg)
    model.load_state_dict({k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("w::")}, strict=False)
    model = model.to(dev()).train()
    z_t, t, lr, hr = [torch.from_numpy(d[k]).to(dev()) for k in ("z_t", "t", "lr", "hr")]
    pred = model(z_t, t, lr)
    assert pred.requires_grad
    loss = torch.nn.functional.mse_loss(pred, hr)            # the loss itself stays on torch (train_ddp_v3mod2.py:889)
    loss.backward()
    assert rel_l2(pred.detach().cpu().numpy(), d["pred"]) <= 2 * yard["pred_rel_l2"]
    assert abs(loss.item() - float(d["loss"])) < 2e-3
    num = den = 0.0
    worst = ("", 0.0)
    for k, p in model.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape, k
        want = d["g::" + k]
        got = p.grad.float().cpu().numpy()
        e = rel_l2(got, want)
        if e > worst[1]:
            worst = (k, e)
        num += float(((got - want) ** 2).sum())
        den += float((want ** 2).sum())
    glob = (num / den) ** 0.5
    assert glob <= 2 * yard["global_rel_l2"], (glob, worst)
    assert worst[1] <= 2 * yard["max_param_rel_l2"], worst


@pytest.mark.parametrize("opt_kind", ["torch_default", "torch_fused", "jat_fused"])
def test_training_step_runs_under_optimizer_and_is_repeatable(opt_kind):
    """AdamW + clip_grad_norm_ on the drop-in module (train_ddp_v3mod2.py:709,926-929): two identical steps from the same
    state give identical gradients (the backward is deterministic apart from f32 atomics in the column reductions),
    the loss goes down over a few steps, and eval-mode outputs track the updated parameters (packed weights refreshed).
    torch's fused AdamW updates the parameters WITHOUT bumping their version counters: the packed copies must follow
    all the same (they are re-cast after every backward unless jat_b200.FusedAdamW has already written them)."""
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
    B, T = 4, 86
    hr, lr, eps = (torch.randn(B, 32, T, generator=g, device=dev()) for _ in range(3))
    t = torch.rand(B, generator=g, device=dev())
    z_t = t.view(B, 1, 1) * hr + (1 - t.view(B, 1, 1)) * eps
    if opt_kind == "jat_fused":
        opt = jat_b200.FusedAdamW(model.parameters(), lr=2e-3, weight_decay=0.1, max_grad_norm=1.0, model=model)
    else:
        opt = torch.optim.AdamW(model.parameters(), lr=2e-3, weight_decay=0.1, fused=(opt_kind == "torch_fused"))

    def grads():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(model(z_t, t, lr), hr)
        loss.backward()
        return loss.item(), torch.cat([p.grad.flatten() for p in model.parameters()])
    l0, g0 = grads()
    l0b, g0b = grads()
    assert l0 == l0b and rel_l2(g0.cpu().numpy(), g0b.cpu().numpy()) < 1e-5
    losses = [l0]
    for _ in range(8):
        if opt_kind != "jat_fused":
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        losses.append(grads()[0])
    assert losses[-1] < 0.9 * losses[0], losses
    model.eval()
    with torch.no_grad():
        e = torch.nn.functional.mse_loss(model(z_t, t, lr), hr).item()
    assert abs(e - losses[-1]) < 1e-3 * max(1.0, losses[-1])


def test_gradient_accumulation_over_two_backward_passes():
    """Two micro-batches without zero_grad in between: p.grad holds the sum (the packed gradient buffers are re-used by every
    backward, so what autograd accumulates must be a copy), and the forward in between sees unchanged weights."""
    import jat_b200
    cfg = dict(input_channels=32, cond_channels=32, patch_len=4, hidden_size=128, depth=2, num_q_heads=2, num_kv_heads=1,
               bottleneck_dim=128, mlp_ratio=2.0, dropout=0.0, drop_path_rate=0.0)
    torch.manual_seed(0)
    model = jat_b200.JaT_AudioSR_V2(**cfg).to(dev()).train()
    g = torch.Generator(device=dev()).manual_seed(4)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "adaLN_modulation.1" in n or n.startswith("final_layer.1"):
                p.copy_(torch.randn(p.shape, generator=g, device=dev()) * 0.02)
    B, T = 3, 86
    batches = []
    for _ in range(2):
        hr, lr, eps = (torch.randn(B, 32, T, generator=g, device=dev()) for _ in range(3))
        t = torch.rand(B, generator=g, device=dev())
        batches.append((t.view(B, 1, 1) * hr + (1 - t.view(B, 1, 1)) * eps, t, lr, hr))

    def grad_of(batch):
        model.zero_grad(set_to_none=True)
        z, t, lr, hr = batch
        torch.nn.functional.mse_loss(model(z, t, lr), hr).backward()
        return [p.grad.clone() for p in model.parameters()]
    g0, g1 = grad_of(batches[0]), grad_of(batches[1])
    model.zero_grad(set_to_none=True)
    for z, t, lr, hr in batches:
        torch.nn.functional.mse_loss(model(z, t, lr), hr).backward()
    for p, a, b in zip(model.parameters(), g0, g1):
        assert rel_l2((a + b).cpu().numpy(), p.grad.cpu().numpy()) < 1e-5


def test_zero_copy_gradient_handoff_matches_copy_and_guards_against_aliasing():
    """grad_handoff='view': .grad tensors are views of the packed gradient buffers -- same values as the default copy mode,
    an optimizer loop with zero_grad(set_to_none=True) trains identically, and a backward pass that would overwrite a live
    .grad (gradient accumulation / set_to_none=False) raises instead of corrupting it."""
    import jat_b200
    from jat_b200 import _lib as L
    cfg = dict(input_channels=32, cond_channels=32, patch_len=4, hidden_size=128, depth=2, num_q_heads=2, num_kv_heads=1,
               bottleneck_dim=128, mlp_ratio=2.0, dropout=0.0, drop_path_rate=0.0)
    g = torch.Generator(device=dev()).manual_seed(5)
    B, T = 3, 86
    hr, lr, eps = (torch.randn(B, 32, T, generator=g, device=dev()) for _ in range(3))
    t = torch.rand(B, generator=g, device=dev())
    z_t = t.view(B, 1, 1) * hr + (1 - t.view(B, 1, 1)) * eps
    models = []
    for mode in ("copy", "view"):
        torch.manual_seed(0)
        m = jat_b200.JaT_AudioSR_V3(**cfg).to(dev()).train()
        gi = torch.Generator(device=dev()).manual_seed(9)
        with torch.no_grad():
            for n, p in m.named_parameters():
                if "adaLN_modulation.1" in n or n.startswith("final_layer.1"):
                    p.copy_(torch.randn(p.shape, generator=gi, device=dev()) * 0.02)
        m.grad_handoff = mode
        models.append(m)
    opts = [jat_b200.FusedAdamW(m.parameters(), lr=2e-3, weight_decay=0.1, max_grad_norm=1.0, model=m) for m in models]
    for step in range(4):
        losses = []
        for m, o in zip(models, opts):
            o.zero_grad(set_to_none=True)
            loss = torch.nn.functional.mse_loss(m(z_t, t, lr), hr)
            loss.backward()
            losses.append(loss.item())
        for (name, p), q in zip(models[0].named_parameters(), models[1].parameters()):
            err = rel_l2(p.grad.cpu().numpy(), q.grad.cpu().numpy())
            # step 0: same parameters, the only difference is the order of the f32 atomics in the column sums.  Later steps:
            # Adam's g / sqrt(v) turns a last-bit difference of a near-zero gradient into a finite parameter difference, so the
            # two trajectories drift apart at the rate any two runs of ONE model do (seen: 2e-4 at step 2)
            assert err < (1e-5 if step == 0 else 2e-3), (step, name, err)
        packed = models[1]._engine.grads(dev()).by_param
        assert all(q.grad.data_ptr() == packed[q].data_ptr() for q in models[1].parameters())   # really zero-copy
        assert abs(losses[0] - losses[1]) <= 1e-4 * max(1.0, losses[0])
        for o in opts:
            o.step()
    with pytest.raises(L.JatError):   # .grad still alive: the next backward would overwrite it
        torch.nn.functional.mse_loss(models[1](z_t, t, lr), hr).backward()
    models[1].zero_grad(set_to_none=True)
    torch.nn.functional.mse_loss(models[1](z_t, t, lr), hr).backward()       # fine again


# ------------------------------------------------------------------------------------------------ f2: fused step glue
@pytest.mark.parametrize("shape", [(3, 64, 1378), (2, 8, 87)])
@pytest.mark.parametrize("variant", ["v3mod2", "v3m2_cfg_dropout", "adaptive", "plain"])
def test_prepare_inputs_bit_exact(shape, variant):
    """`training.prepare_inputs` against the reference's torch expressions (train_ddp_v3mod2.py:856-883,
    train_ddp_v3m2.py:547-580), evaluated on the same device in fp32: bit-identical."""
    from jat_b200 import training
    B, C, T = shape
    g = torch.Generator(device=dev()).manual_seed(5)
    hr, lr, noise, cn = (torch.randn(B, C, T, generator=g, device=dev()) * s for s in (3.0, 2.0, 1.0, 1.0))
    hr_mean, lr_mean = (torch.randn(1, C, 1, generator=g, device=dev()) for _ in range(2))
    hr_std, lr_std = (torch.rand(1, C, 1, generator=g, device=dev()) + 0.5 for _ in range(2))
    t = torch.rand(B, generator=g, device=dev())
    ratio = 0.05
    hr_norm = (hr - hr_mean) / hr_std
    lr_norm = (lr - lr_mean) / lr_std
    kw = {}
    if variant in ("v3mod2", "v3m2_cfg_dropout"):
        lr_norm = lr_norm + cn * (ratio * 1.0)
        kw = dict(cond_noise=cn, cond_scale=ratio)
    elif variant == "adaptive":
        std = lr_norm.std().clamp(0.5, 2.0)
        lr_norm = lr_norm + cn * (ratio * std)
        kw = dict(cond_noise=cn, cond_scale=ratio, cond_scale_dev=std)
    if variant == "v3m2_cfg_dropout":
        mask = torch.tensor([True, False, True][:B], device=dev()).view(B, 1, 1)
        lr_norm = lr_norm * (~mask).float()
        kw["keep"] = (~mask).float()
    tv = t.view(-1, 1, 1)
    z_t = tv * hr_norm + (1 - tv) * noise
    got = training.prepare_inputs(hr, lr, hr_mean, hr_std, lr_mean, lr_std, t, noise, **kw)
    for name, a, b in zip(("hr_norm", "lr_cond", "z_t"), got, (hr_norm, lr_norm, z_t)):
        assert torch.equal(a, b), (name, (a - b).abs().max().item())


@pytest.mark.parametrize("n_shape", [(4, 64, 1378), (1, 3, 7)])
def test_fused_mse_loss_matches_torch(n_shape):
    from jat_b200 import training
    g = torch.Generator(device=dev()).manual_seed(6)
    pred = torch.randn(*n_shape, generator=g, device=dev()).requires_grad_(True)
    target = torch.randn(*n_shape, generator=g, device=dev())
    loss, stats = training.mse_loss(pred, target, return_stats=True)
    (loss * 3.0).backward()
    p2 = pred.detach().clone().requires_grad_(True)
    want = torch.nn.functional.mse_loss(p2, target)
    (want * 3.0).backward()
    assert abs(loss.item() - want.item()) <= 2e-6 * abs(want.item())
    assert rel_l2(pred.grad.cpu().numpy(), p2.grad.cpu().numpy()) < 1e-6
    n = pred.numel()
    s = stats.cpu().numpy()
    assert abs(s[1] / n - pred.mean().item()) < 1e-5                                   # pred mean (:902)
    assert abs((s[2] / n - (s[1] / n) ** 2) ** 0.5 - pred.std(unbiased=False).item()) < 1e-4
    snr = 10 * np.log10(s[3] / s[0])                                                   # :905-908
    want_snr = (10 * torch.log10((target ** 2).mean() / (((pred - target) ** 2).mean() + 1e-8))).item()
    assert abs(snr - want_snr) < 1e-3


@pytest.mark.parametrize("n_shape", [(4, 64, 1378), (1, 3, 7)])
def test_fused_charbonnier_loss_matches_reference_expression(n_shape):
    """train_ddp_v3mod3.py:57-85: loss = sqrt((pred - target)^2 + eps).mean(); value and gradient against torch autograd."""
    from jat_b200 import training
    g = torch.Generator(device=dev()).manual_seed(8)
    pred = torch.randn(*n_shape, generator=g, device=dev()).requires_grad_(True)
    target = torch.randn(*n_shape, generator=g, device=dev())
    target[..., :2] = pred.detach()[..., :2]                 # exact zeros of the residual: the eps branch
    loss, stats = training.charbonnier_loss(pred, target, eps=1e-6, return_stats=True)
    (loss * 2.0).backward()
    p2 = pred.detach().clone().requires_grad_(True)
    want = torch.sqrt((p2 - target) ** 2 + 1e-6).mean()
    (want * 2.0).backward()
    assert abs(loss.item() - want.item()) <= 3e-6 * abs(want.item())
    assert rel_l2(pred.grad.cpu().numpy(), p2.grad.cpu().numpy()) < 1e-6
    s = stats.cpu().numpy()
    assert abs(s[4] / pred.numel() - ((pred - target) ** 2).mean().item()) < 1e-5
