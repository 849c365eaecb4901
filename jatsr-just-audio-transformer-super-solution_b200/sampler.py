"""Flow-matching Euler sampler with classifier-free guidance -- drop-in for
`flow_matching_sample` of the reference (infer_test_v3m2.py:108-185).

Differences in HOW (not in WHAT) it computes:
  * the conditional and unconditional halves run as ONE batch of 2B token rows, but nothing is
    concatenated: the patchify kernel reads z for both halves, the condition for the first half and
    zeros for the second (reference: three torch.cat per step, :154-156);
  * t is the same for every row of a step, so t_embedder and all blocks' adaLN modulations are
    computed ONCE for all steps (one [steps, D] x [D, depth*6D] GEMM) before the loop;
  * CFG combine + x-prediction->velocity + Euler update is a single fused kernel reading t/dt from
    device memory: no host synchronisation inside the loop (reference: 3 syncs per step, :150,173,183);
  * the whole loop can be captured into a CUDA graph and replayed (`use_graph`).
The initial noise is drawn exactly like the reference (`torch.randn(B, C, T, device=device)` from the
global generator, :133), so under the same seed both samplers start from the same z0.
"""
from __future__ import annotations

import os

import torch

from . import ops
from .models import _JaTBase


_MAX_PLANS = 4


class _Plan:
    """Buffers (+ optional captured graph) for one (model, B, C, T, steps, cfg) sampling problem."""

    def __init__(self, model, B, Cc, T, num_steps, cfg_scale, device):
        """Cc = channels of the condition latent; the generated latent has model.input_channels (the reference draws z0
        with the condition's shape, infer_test_v3m2.py:133 -- its configs have equal channel counts)."""
        self.key = (B, Cc, T, num_steps, float(cfg_scale), device)
        Cz = model.input_channels
        self.use_cfg = cfg_scale != 1.0
        self.Beff = 2 * B if self.use_cfg else B
        eng = model._engine
        self.ws = eng.workspace(self.Beff, T, num_steps, device)
        self.z = torch.empty(B, Cz, T, dtype=torch.float32, device=device)
        self.lr = torch.empty(B, Cc, T, dtype=torch.float32, device=device)
        self.x_pred = torch.empty(self.Beff, Cz, T, dtype=torch.float32, device=device)
        # timesteps exactly as the reference builds them (:136), on the same device
        ts = torch.linspace(0.0, 1.0, num_steps + 1, device=device)
        self.t_curr = ts[:-1].contiguous()
        self.t_dt = torch.stack([ts[:-1], ts[1:] - ts[:-1]], dim=1).contiguous()
        self.graph = None


def _run_steps(model, plan, B, num_steps, cfg_scale):
    eng = model._engine
    mod = eng.modulation(plan.ws, plan.t_curr)  # [steps, depth*6D], one row per Euler step
    for i in range(num_steps):
        eng.forward_tokens(plan.ws, plan.z, plan.lr, plan.Beff, mod[i], 0, plan.x_pred, cond_batch=B)
        x_c = plan.x_pred[:B]
        x_u = plan.x_pred[B:] if plan.use_cfg else None
        ops.cfg_euler_update(plan.z, x_c, x_u, cfg_scale, plan.t_dt, i)


@torch.no_grad()
def flow_matching_sample(model, lr_latent, num_steps=50, cfg_scale=1.0, device="cuda", verbose=True,
                         use_graph=None, z0=None):
    """Same signature and return value as the reference (extra keyword-only-style options last).

    lr_latent: [B, C, T] normalised LR condition latent.  Returns the generated HR latent [B, C, T] f32.
    use_graph: capture the step loop into a CUDA graph (default: env JAT_B200_GRAPH, on by default).
    z0: optional initial noise (testing); by default drawn with torch.randn like the reference.
    """
    if not isinstance(model, _JaTBase):
        raise TypeError("flow_matching_sample needs a jat_b200 JaT_AudioSR_V2/V3 model")
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("jat_b200 sampler runs on CUDA only; there is no CPU fallback")
    if device.index is None:
        # the reference's default is device='cuda' (no index); torch.device('cuda') != torch.device('cuda:0'), so resolve
        # it once: the plan key, the packed weights and the tensors all carry the same indexed device and the caches hit
        device = torch.device("cuda", torch.cuda.current_device())
    if verbose:
        print(f"  Flow Matching sampling ({num_steps} steps, CFG scale={cfg_scale})...")
    if model.training:
        raise RuntimeError("flow_matching_sample expects model.eval()")
    B, Cc, T = lr_latent.shape
    if Cc != model.cond_channels:
        raise ValueError(f"expected lr_latent [B, {model.cond_channels}, T], got {tuple(lr_latent.shape)}")
    z_init = torch.randn(B, model.input_channels, T, device=device) if z0 is None else z0.to(device=device, dtype=torch.float32)
    if use_graph is None:
        use_graph = os.environ.get("JAT_B200_GRAPH", "1") != "0"

    # Plans (buffers + captured graph) are kept per problem shape, a few at a time: the chunk loop of a long track alternates
    # between the full-chunk batch and the short last chunk, and re-capturing 10 k launches (plus freeing the previous graph's
    # memory) on every call cost more than the launches themselves.  A new PackedWeights object (weights re-allocated, e.g.
    # after `.to()`) invalidates every captured graph; an in-place refresh keeps the pointers and the plans.
    cache = model.__dict__.setdefault("_sampler_plans", {})
    key = (B, Cc, T, num_steps, float(cfg_scale), device)
    packed = model._engine.weights(device)
    if any(pl.packed is not packed for pl in cache.values()):
        cache.clear()
    plan = cache.pop(key, None)
    if plan is None:
        while len(cache) >= _MAX_PLANS:
            cache.pop(next(iter(cache)))          # least recently used
        plan = _Plan(model, B, Cc, T, num_steps, cfg_scale, device)
        plan.packed = packed
    cache[key] = plan                             # most recently used last
    plan.z.copy_(z_init)
    plan.lr.copy_(lr_latent.to(device=device, dtype=torch.float32))

    if use_graph:
        if plan.graph is None:
            # warm up once outside capture (lazy cudaFuncSetAttribute etc.), then capture the whole loop
            z_save = plan.z.clone()
            _run_steps(model, plan, B, min(num_steps, 1), cfg_scale)
            torch.cuda.synchronize(device)
            plan.z.copy_(z_save)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                _run_steps(model, plan, B, num_steps, cfg_scale)
            plan.graph = g
            plan.z.copy_(z_save)
        plan.graph.replay()
    else:
        _run_steps(model, plan, B, num_steps, cfg_scale)

    if verbose:
        ts = torch.linspace(0.0, 1.0, num_steps + 1)
        for i in range(num_steps):
            if (i + 1) % 10 == 0 or i == num_steps - 1:
                print(f"    Step {i+1}/{num_steps}, t={ts[i]:.3f} → {ts[i+1]:.3f}")
    return plan.z.clone()
