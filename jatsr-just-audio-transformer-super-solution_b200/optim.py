"""The parameter update of the training step as two multi-tensor CUDA passes (optimizer side).

  FusedAdamW  <->  train_ddp_v3mod2.py:709 `optim.AdamW(model.parameters(), lr, weight_decay)` together with
                   :926 `clip_grad_norm_(model.parameters(), 1.0)` and the re-cast of the updated f32 parameters into
                   the packed bf16 buffers the GEMMs read (engine.PackedWeights.refresh).

It IS a `torch.optim.AdamW`: same constructor, same `param_groups` (lr schedulers work), same `state` / `state_dict()`
layout (`step`, `exp_avg`, `exp_avg_sq` per parameter), so the reference's checkpoints load into it and its checkpoints
load into the stock optimizer.  Only `step()` differs: `jat_grad_sumsq` (one read of the gradients -> ||g||^2 on the
device) and `jat_adamw_step` (one pass: clip coefficient applied on the fly, decoupled decay, Adam update in ATen's
operation order, packed copy written in the same pass).  There is no host synchronisation and no CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib as L
from .ops import _stream


class _Group:
    """Device table (jat_adamw_tensor [n]) + chunk index of one parameter group."""

    def __init__(self, params, states, packed_of, device, grad_source=None):
        self.params = params
        # reduced bf16 gradients living in the payload buffers of the bf16 gradient exchange (jat_b200.ddp, fused consumer):
        # (payload tensor, element offset) per parameter, or None -> the parameter's f32 .grad
        where = grad_source.where if grad_source is not None else {}
        self.bf16_src = [where.get(id(p)) for p in params]
        self.ids = [id(p) for p in params]
        n = len(params)
        chunk = L.load().jat_adamw_chunk_elems()
        first, total = [], 0
        for p in params:
            first.append(total)
            total += (p.numel() + chunk - 1) // chunk
        self.n, self.total_chunks = n, total
        self.chunk_first = torch.tensor(first, dtype=torch.int32).to(device)
        self.table = torch.zeros(n, 8, dtype=torch.int64, device=device)
        self.partials = torch.empty(total, dtype=torch.float32, device=device)
        dsts = [packed_of.get(id(p)) for p in params]
        rows = [[p.data_ptr(), 0, states[p]["exp_avg"].data_ptr(), states[p]["exp_avg_sq"].data_ptr(),
                 d.data_ptr() if d is not None else 0, p.numel(), 0, 0] for p, d in zip(params, dsts)]
        self.static = torch.tensor(rows, dtype=torch.int64)
        self.static_ok = torch.tensor([all(r[j] % 16 == 0 for j in (0, 2, 3, 4)) and r[5] % 4 == 0 for r in rows])
        self.dtype_bits = torch.tensor([1 if (d is not None and d.dtype == torch.bfloat16) else 0 for d in dsts],
                                       dtype=torch.int64)
        self.grad_bits = torch.tensor([0 if w is None else 2 for w in self.bf16_src], dtype=torch.int64)    # JAT_ADAMW_GRAD_BF16
        self.grad_align = torch.tensor([16 if w is None else 8 for w in self.bf16_src], dtype=torch.int64)
        self.key = self._key(states)
        # every tensor at the same step count (the usual case): kept as one python int, see upload()
        st = [float(states[p]["step"]) for p in params]
        self.uniform_step = int(st[0]) if all(x == st[0] for x in st) else None
        # two pinned staging buffers used alternately, each guarded by an event: the host may run a step ahead of the stream
        self.host = [torch.empty(n, 8, dtype=torch.int64).pin_memory() for _ in range(2)]
        self.done = [None, None]
        self.turn = 0

    def _key(self, states):
        return tuple((p.data_ptr(), states[p]["exp_avg"].data_ptr(), states[p]["exp_avg_sq"].data_ptr()) for p in self.params)

    def upload(self, steps, beta1, beta2):
        """Gradient pointers change from step to step (autograd hands out fresh tensors): column 1 + the vec_ok flag; the
        bias corrections follow each tensor's own step count (torch/optim/adam.py keeps `step` per parameter).  Host cost
        matters here: under DDP the host is NOT ahead of the GPU at the end of the backward, so every 100 us spent before the
        update kernels are launched is a stall of the device.  The common case (every tensor at the same step) therefore
        avoids the per-tensor work: one python-side step count, two scalars broadcast into the table."""
        g = torch.tensor([p.grad.data_ptr() if w is None else w[0].data_ptr() + 2 * w[1] for p, w in zip(self.params, self.bf16_src)],
                         dtype=torch.int64)
        k = self.turn
        self.turn ^= 1
        if self.done[k] is not None:
            self.done[k].synchronize()      # the copy issued two steps ago has long finished; wait if it has not
        h = self.host[k]
        h.copy_(self.static)
        h[:, 1] = g
        h[:, 6] = self.dtype_bits | self.grad_bits | ((self.static_ok & (g % self.grad_align == 0)).to(torch.int64) << 32)
        if self.uniform_step is not None:
            self.uniform_step += 1
            st = float(self.uniform_step)
            bc = torch.tensor([1.0 - beta1 ** st, (1.0 - beta2 ** st) ** 0.5], dtype=torch.float64).to(torch.float32)
            h[:, 7] = bc.view(torch.int64)  # two packed f32: bias_corr1 | bias_corr2_sqrt, the same for every tensor
        else:
            st = torch.stack(steps).to(torch.float64)
            bc = torch.stack([1.0 - beta1 ** st, torch.sqrt(1.0 - beta2 ** st)], 1).to(torch.float32).contiguous()
            h[:, 7] = bc.view(torch.int64).reshape(-1)
        self.table.copy_(h, non_blocking=True)
        self.done[k] = torch.cuda.Event()
        self.done[k].record(torch.cuda.current_stream(self.table.device))


class FusedAdamW(torch.optim.AdamW):
    """`torch.optim.AdamW` whose `step()` runs on the library's multi-tensor kernels.

    max_grad_norm: if set, `step()` also does what `clip_grad_norm_(params, max_grad_norm)` does before the update (one
        norm over ALL parameters of the optimizer), except that `p.grad` is left unscaled; the total norm of the last
        step is in `self.grad_norm` (device scalar, no sync).
    model: a jat_b200 drop-in module; its packed bf16 / f32 weight copies are then refreshed by the same pass, so the next
        forward does not re-cast 766 M parameters.
    Every parameter must be f32, contiguous, on one CUDA device and have a dense gradient at `step()` time."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, *, max_grad_norm=None, model=None):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, foreach=False, fused=False)
        self.max_grad_norm = None if max_grad_norm is None else float(max_grad_norm)
        self._model = model
        self._groups = None
        self._packed_seen = None
        self._sumsq = None
        self._covers_model = False
        self._grad_source = None   # set by jat_b200.ddp.register_bf16_allreduce(..., optimizer=self)
        self.grad_norm = None

    # state in torch.optim.AdamW's layout (torch/optim/adam.py:_init_group): step (f32 scalar), exp_avg, exp_avg_sq
    def _ensure_state(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _packed_map(self):
        eng = getattr(self._model, "_engine", None) if self._model is not None else None
        pk = eng.packed if eng is not None else None
        if pk is None:
            return None, {}
        return pk, {id(src): dst for dst, src in zip(pk._dst, pk._src)}

    def _build(self, device):
        pk, packed_of = self._packed_map()
        self._packed_seen = pk
        self._groups = []
        for group in self.param_groups:
            params = [p for p in group["params"] if p.grad is not None and p.numel() > 0]
            for p in params:
                if p.dtype != torch.float32 or not p.is_contiguous() or p.device != device or p.grad.is_sparse:
                    raise L.JatError(L.ERR_BAD_ARG, "FusedAdamW: parameters must be f32, contiguous, dense and on one CUDA device")
                self._ensure_state(p)
                if self.state[p]["step"].device.type != "cpu":   # e.g. a checkpoint written by torch's fused / capturable AdamW
                    self.state[p]["step"] = self.state[p]["step"].detach().to("cpu", torch.float32)
                if self.state[p]["exp_avg"].device != device:  # state loaded from a checkpoint on another device
                    for k in ("exp_avg", "exp_avg_sq"):
                        self.state[p][k] = self.state[p][k].to(device)
            self._groups.append(_Group(params, self.state, packed_of, device, self._grad_source) if params else None)
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=device)
        # the packed copies may be declared fresh after a step only if this optimizer updates every parameter they mirror
        mine = {id(p) for g in self._groups if g is not None for p in g.params}
        self._covers_model = pk is not None and all(id(src) in mine for src in pk._src) and \
            all(id(p) in packed_of for p in self._model.parameters())

    def _current(self):
        if self._groups is None:
            return False
        pk, _ = self._packed_map()
        if pk is not self._packed_seen:
            return False
        for group, g in zip(self.param_groups, self._groups):
            params = [p for p in group["params"] if p.grad is not None and p.numel() > 0]
            if (g is None) != (not params) or (g is not None and ([id(p) for p in params] != g.ids or g.key != g._key(self.state))):
                return False
        return True

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        first = next((p for group in self.param_groups for p in group["params"] if p.grad is not None), None)
        if first is None:
            return loss
        device = first.device
        if device.type != "cuda":
            raise L.JatError(L.ERR_BAD_ARG, "FusedAdamW runs on CUDA parameters only (no CPU path)")
        if not self._current():
            self._build(device)
        lib, ctx, stream = L.load(), L.context(device.index if device.index is not None else torch.cuda.current_device()), \
            _stream(device)
        live = [g for g in self._groups if g is not None]
        for group, g in zip(self.param_groups, self._groups):
            if g is None:
                continue
            if group.get("amsgrad") or group.get("maximize"):
                raise L.JatError(L.ERR_BAD_ARG, "FusedAdamW: amsgrad / maximize are not supported")
            steps = [self.state[p]["step"] for p in g.params]
            torch._foreach_add_(steps, 1)          # host tensors: no device sync
            g.upload(steps, *group["betas"])
        clip = self.max_grad_norm is not None
        if clip:
            for j, g in enumerate(live):
                L.check(lib.jat_grad_sumsq(ctx, g.table.data_ptr(), g.chunk_first.data_ptr(), g.n, g.total_chunks,
                                           g.partials.data_ptr(), self._sumsq.data_ptr(), int(j > 0), stream))
            self.grad_norm = self._sumsq.sqrt().float().squeeze(0)
        for group, g in zip(self.param_groups, self._groups):
            if g is None:
                continue
            b1, b2 = group["betas"]
            lr = group["lr"]
            L.check(lib.jat_adamw_step(ctx, g.table.data_ptr(), g.chunk_first.data_ptr(), g.n, g.total_chunks,
                                       float(lr), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                                       float(self.max_grad_norm) if clip else 0.0,
                                       self._sumsq.data_ptr() if clip else None, stream))
        # the kernels wrote through raw pointers: tell autograd / the engine that the parameters changed ...
        torch.autograd.graph.increment_version([p for g in live for p in g.params])
        pk = self._packed_seen
        if pk is not None and self._covers_model:
            pk.versions = pk._versions(self._model)   # ... and that the packed copies already hold the new values
            pk.dirty = False
        return loss
