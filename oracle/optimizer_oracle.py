"""CPU oracle for the parameter update of the training step -- TEST INFRASTRUCTURE ONLY.

A plain numpy restatement of what the reference's loop does with the gradients (train_ddp_v3mod2.py:926-928):
  * `torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)`: one 2-norm over ALL gradients, then every gradient is
    scaled by min(1, max_norm / (norm + 1e-6));
  * `optimizer.step()` with `AdamW(lr, weight_decay)` (:709): decoupled weight decay, Adam moments, bias corrections from
    each parameter's own step count (torch/optim/adamw.py, single-tensor path).
The reference's implementation IS torch, which is present wherever the tests run: `tests/test_oracle.py` pins this file
against `clip_grad_norm_` + `torch.optim.AdamW` on CPU, and the CUDA path (`jat_grad_sumsq` + `jat_adamw_step` behind
`jat_b200.FusedAdamW`) is tested against torch directly in `tests/test_optim_gpu.py`.  Only tests may import this module.
"""
from __future__ import annotations

import numpy as np


def grad_norm(grads):
    """2-norm over all gradient tensors (clip_grad_norm_: norm of the per-tensor norms, accumulated in f32 by torch;
    f64 here, the difference is far below the f32 rounding of the result)."""
    return float(np.sqrt(sum(float(np.sum(g.astype(np.float64) ** 2)) for g in grads)))


def clip_coefficient(total_norm, max_norm):
    """clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to at most 1."""
    return min(1.0, float(max_norm) / (float(total_norm) + 1e-6))


def adamw_step(params, grads, exp_avgs, exp_avg_sqs, steps, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2,
               max_norm=None):
    """One update, in place on the f32 numpy arrays; `steps` is a list of per-parameter step counts BEFORE the update
    (incremented here).  Returns the total gradient norm (None without clipping)."""
    b1, b2 = betas
    total = clip = None
    if max_norm is not None:
        total = grad_norm(grads)
        clip = np.float32(clip_coefficient(total, max_norm))
    for i, (p, g, m, v) in enumerate(zip(params, grads, exp_avgs, exp_avg_sqs)):
        steps[i] += 1
        g = g.astype(np.float32)
        if clip is not None:
            g = g * clip
        p *= np.float32(1.0 - lr * weight_decay)
        m += (g - m) * np.float32(1.0 - b1)                       # exp_avg.lerp_(grad, 1 - beta1)
        v *= np.float32(b2)
        v += np.float32(1.0 - b2) * g * g                         # exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        bc1 = 1.0 - b1 ** steps[i]
        bc2_sqrt = (1.0 - b2 ** steps[i]) ** 0.5
        denom = np.sqrt(v) / np.float32(bc2_sqrt) + np.float32(eps)
        p -= np.float32(lr / bc1) * (m / denom)                   # param.addcdiv_(exp_avg, denom, value=-step_size)
    return total
