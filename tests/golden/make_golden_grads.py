"""Generates tests/golden/micro_*_grads.npz: parameter gradients of the UNMODIFIED reference modules (CPU, fp32) for one
training-style step -- x-prediction MSE loss on z_t = t x + (1 - t) eps (train_ddp_v3mod2.py:881-889) -- at a micro
configuration with dropout = drop_path = 0 (gradient parity is only defined without the RNG-driven masks).

Run in the build container only:  python tests/golden/make_golden_grads.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import MICRO, build, import_reference  # noqa: E402


def main():
    V2, V3, _ = import_reference()
    cfg = dict(MICRO, dropout=0.0, drop_path_rate=0.0)
    for tag, cls in (("v2_layernorm", V2), ("v3_rmsnorm", V3)):
        model = build(cls, cfg, seed=20).train()
        g = torch.Generator().manual_seed(6)
        B, C, T = 3, cfg["input_channels"], 86
        hr = torch.randn(B, C, T, generator=g)
        lr = torch.randn(B, C, T, generator=g)
        noise = torch.randn(B, C, T, generator=g)
        t = torch.tensor([0.07, 0.55, 0.93])
        z_t = t.view(B, 1, 1) * hr + (1 - t.view(B, 1, 1)) * noise          # train_ddp_v3mod2.py:881-883
        pred = model(z_t, t, lr)
        loss = torch.nn.functional.mse_loss(pred, hr)                        # :889
        loss.backward()
        data = {"w::" + k: v.detach().numpy() for k, v in model.state_dict().items()
                if not k.endswith(("cos_cached", "sin_cached"))}
        data.update({"g::" + k: p.grad.numpy() for k, p in model.named_parameters()})
        data.update(z_t=z_t.numpy(), t=t.numpy(), lr=lr.numpy(), hr=hr.numpy(), pred=pred.detach().numpy(),
                    loss=np.float32(loss.item()), cfg_json=np.array(repr(cfg)))
        out = os.path.join(HERE, f"micro_{tag}_grads.npz")
        np.savez_compressed(out, **data)
        gn = float(torch.sqrt(sum((p.grad ** 2).sum() for p in model.parameters())))
        print(f"{tag}: loss {loss.item():.5f} grad-norm {gn:.4f} -> {out} ({os.path.getsize(out) // 1024} KiB)")


if __name__ == "__main__":
    main()
