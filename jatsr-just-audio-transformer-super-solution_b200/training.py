"""Fused elementwise pieces of the training step around the model call.

  prepare_inputs  <->  train_ddp_v3mod2.py:856-883 / train_ddp_v3m2.py:547-580: normalise, conditional-noise
                       augmentation, CFG condition dropout, flow-matching mix -- one kernel (`jat_train_inputs`)
  mse_loss        <->  F.mse_loss(pred_x0, hr_norm) (:889) + the monitoring sums of :900-911 -- one kernel
                       (`jat_mse_loss`) that also writes the gradient seed, so `loss.backward()` costs one scaling pass
  charbonnier_loss <-> charbonnier_loss(pred_x0, hr_norm, eps) of the MOD3 script (train_ddp_v3mod3.py:57-85, :957)

The random draws stay torch's (`torch.rand`, `torch.randn_like` from the global generators, like the reference), so a
training script keeps its RNG stream; only the arithmetic moves into the library.
"""
from __future__ import annotations

import torch

from . import _lib as L
from .ops import _chk, _ctx, _p, _stream


def u_shaped_timestep_sampling(batch_size, device, alpha=0.5):
    """train_ddp_v3mod2.py:449-457 (B scalars: left on torch)."""
    u = torch.rand(batch_size, device=device)
    return torch.where(u < 0.5, (2 * u) ** alpha / 2, 1 - ((2 * (1 - u)) ** alpha) / 2)


def prepare_inputs(hr, lr, hr_mean, hr_std, lr_mean, lr_std, t, noise, cond_noise=None, cond_scale=0.0, cond_scale_dev=None,
                   keep=None):
    """-> (hr_norm, lr_cond, z_t), all f32 [B, C, T].  hr_mean/... are the reference's [1, C, 1] statistics tensors
    (any shape with C elements).  cond_scale_dev: optional device scalar multiplied onto cond_scale (adaptive noise:
    `lr_norm.std().clamp(0.5, 2.0)`).  keep: f32 [B], 0 where the CFG dropout masks the condition (`(~cfg_mask).float()`)."""
    for x, n in ((hr, "hr"), (lr, "lr"), (noise, "noise"), (t, "t")):
        _chk(x, torch.float32, n)
    B, C, T = hr.shape
    stats = [s.reshape(-1).contiguous().float() for s in (hr_mean, hr_std, lr_mean, lr_std)]
    assert all(s.numel() == C for s in stats) and t.shape == (B,)
    if cond_noise is not None:
        _chk(cond_noise, torch.float32, "cond_noise")
    if keep is not None:
        keep = keep.reshape(B).float().contiguous()
    if cond_scale_dev is not None:
        cond_scale_dev = cond_scale_dev.reshape(1).float().contiguous()
    hr_norm, lr_cond, z_t = torch.empty_like(hr), torch.empty_like(hr), torch.empty_like(hr)
    L.check(L.load().jat_train_inputs(_ctx(hr), hr.data_ptr(), lr.data_ptr(), *[s.data_ptr() for s in stats], noise.data_ptr(),
                                      _p(cond_noise), _p(cond_scale_dev), float(cond_scale), _p(keep), t.data_ptr(),
                                      hr_norm.data_ptr(), lr_cond.data_ptr(), z_t.data_ptr(), B, C, T, _stream(hr.device)))
    return hr_norm, lr_cond, z_t


class _MseLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, charbonnier_eps=None):
        p = pred.float().contiguous()
        q = target.float().contiguous()
        need = pred.requires_grad
        d_pred = torch.empty_like(p) if need else None
        if charbonnier_eps is None:
            stats = torch.empty(4, dtype=torch.float64, device=p.device)
            L.check(L.load().jat_mse_loss(_ctx(p), p.data_ptr(), q.data_ptr(), _p(d_pred), stats.data_ptr(), p.numel(),
                                          _stream(p.device)))
        else:
            stats = torch.empty(5, dtype=torch.float64, device=p.device)
            L.check(L.load().jat_charbonnier_loss(_ctx(p), p.data_ptr(), q.data_ptr(), _p(d_pred), stats.data_ptr(), p.numel(),
                                                  float(charbonnier_eps), _stream(p.device)))
        ctx.d_pred, ctx.in_dtype = d_pred, pred.dtype
        loss = (stats[0] / p.numel()).float()
        ctx.mark_non_differentiable(stats)
        return loss, stats

    @staticmethod
    def backward(ctx, g_loss, _g_stats):
        d = ctx.d_pred
        ctx.d_pred = None
        return (d * g_loss).to(ctx.in_dtype), None, None


def mse_loss(pred, target, return_stats=False):
    """mean((pred - target)^2) as a device scalar.  With return_stats also the device tensor (float64 [4])
    [sum sq err, sum pred, sum pred^2, sum target^2] from which the reference's monitoring values follow without extra
    passes or syncs: pred mean / std, SNR = 10 log10(sum target^2 / sum sq err) (train_ddp_v3mod2.py:900-911)."""
    loss, stats = _MseLoss.apply(pred, target)
    return (loss, stats) if return_stats else loss


def charbonnier_loss(pred, target, eps=1e-6, return_stats=False):
    """mean(sqrt((pred - target)^2 + eps)) -- `charbonnier_loss` of the MOD3 training script (train_ddp_v3mod3.py:57-85) in
    one pass with its gradient seed.  With return_stats also the device tensor (float64 [5]) [sum sqrt(d^2 + eps), sum pred,
    sum pred^2, sum target^2, sum d^2] behind the script's monitoring values (:1031ff)."""
    loss, stats = _MseLoss.apply(pred, target, float(eps))
    return (loss, stats) if return_stats else loss
