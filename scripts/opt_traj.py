"""Loss / gradient-norm trajectory of the configs[3] training step under torch's clip + AdamW and under FusedAdamW."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import jat_b200  # noqa: E402
from jat_b200 import training  # noqa: E402

dev = torch.device("cuda", 0)
B, C, T = bench.B, bench.C, bench.T
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8


def run(kind):
    torch.manual_seed(0)
    with torch.device(dev):
        model = jat_b200.JaT_AudioSR_V2(**bench.CFG)
    g = torch.Generator(device=dev).manual_seed(1)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "adaLN_modulation.1" in name or name.startswith("final_layer.1"):
                p.copy_(torch.randn(p.shape, generator=g, device=dev) * 0.02)
    model.train()
    if kind == "fused":
        opt = jat_b200.FusedAdamW(model.parameters(), lr=5e-5, weight_decay=0.1, max_grad_norm=1.0, model=model)
    else:
        opt = torch.optim.AdamW(model.parameters(), lr=5e-5, weight_decay=0.1, fused=True)
    gd = torch.Generator(device=dev).manual_seed(100)
    hr = torch.randn(B, C, T, generator=gd, device=dev) * 2.0 + 0.3
    lr = torch.randn(B, C, T, generator=gd, device=dev) * 2.0 + 0.3
    mean, std = torch.full((1, C, 1), 0.3, device=dev), torch.full((1, C, 1), 2.0, device=dev)
    out = []
    for _ in range(steps):
        u = torch.rand(B, generator=gd, device=dev)
        t = torch.where(u < 0.5, (2 * u).sqrt() / 2, 1 - (2 * (1 - u)).sqrt() / 2)
        noise = torch.randn(B, C, T, generator=gd, device=dev)
        cn = torch.randn(B, C, T, generator=gd, device=dev)
        hr_n, lr_c, z_t = training.prepare_inputs(hr, lr, mean, std, mean, std, t, noise, cond_noise=cn, cond_scale=0.05)
        opt.zero_grad(set_to_none=True)
        loss = training.mse_loss(model(z_t, t, lr_c), hr_n)
        loss.backward()
        if kind == "fused":
            opt.step()
            gn = float(opt.grad_norm)
        else:
            gn = float(torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0))
            opt.step()
        psum = float(sum(p.detach().double().abs().sum() for p in model.parameters()))
        out.append((round(loss.item(), 5), round(gn, 5), psum))
    return out


a, b = run("torch"), run("fused")
for i, (x, y) in enumerate(zip(a, b)):
    print(i, "torch loss %.5f gnorm %.5f |p| %.6e   fused loss %.5f gnorm %.5f |p| %.6e" % (x + y))
