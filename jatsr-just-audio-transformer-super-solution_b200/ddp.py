"""Gradient exchange of the DDP training step (training row; reference train_ddp_v3mod2.py:822, 922).

The reference wraps the model in `torch.nn.parallel.DistributedDataParallel` and lets it all-reduce 766 M f32 gradients
(3.06 GB per step and rank) in buckets while the backward runs.  The drop-in module works with exactly that call (its staged
autograd chain hands gradients to DDP block by block, `models._Stage`).  This module adds the one thing worth changing on
NVLink: the PAYLOAD.  `register_bf16_allreduce(ddp_model)` installs a DDP communication hook that

  1. casts a ready bucket to bf16, pre-scaled by 1 / world_size, with one library kernel (`jat_grad_compress`: read 4 B,
     write 2 B per element) into a persistent per-bucket payload buffer,
  2. all-reduces the bf16 payload (NCCL SUM over NVLink / NVSwitch) -- half the bytes, half the time NCCL's ring kernels
     share SMs and HBM with the weight-gradient GEMMs of the blocks still in the backward,
  3. expands the mean back into the f32 bucket (`jat_grad_decompress`), so `.grad`, `clip_grad_norm_`, `GradScaler.unscale_`
     and every optimizer see ordinary f32 gradients (rounded to bf16 precision).

torch ships a `bf16_compress_hook` with the same dataflow built from `Tensor.to` / `div_` / `copy_` (4 elementwise passes and
two allocations per bucket); at 8 GPUs it was SLOWER than the plain f32 all-reduce.  The two fused passes
here cost 12 B per parameter per step in total.

With `optimizer=` a `jat_b200.FusedAdamW`, step 3 is dropped as well ("fused consumer"): the update and the gradient-norm pass
read the all-reduced bf16 payload directly (`JAT_ADAMW_GRAD_BF16` table entries), the f32 buckets are never re-expanded and
`.grad` keeps the rank's LOCAL f32 gradient -- so this mode is for loops whose only consumer of the gradients is that
optimizer (its `max_grad_norm` clipping included); anything else that reads `.grad` should use the default mode.

Numerics: each rank's gradient is rounded once to bf16 (relative 2^-9) before the sum and the sum once more: the exchanged
gradient differs from the f32 all-reduce by rel-L2 ~3e-3 (tests/_ddp_worker.py asserts < 5e-3; all ranks hold bit-identical
results, so parameters stay in lock-step).  Keep the default f32 exchange when that matters more than step time.
"""

import torch
import torch.distributed as dist

from . import _lib as L
from .ops import _ctx, _stream


class Bf16AllreduceState:
    def __init__(self, process_group=None, optimizer=None):
        self.group = process_group if process_group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.payload = {}   # bucket index -> persistent bf16 buffer
        self.optimizer = optimizer   # FusedAdamW consuming the payload directly (fused consumer), or None
        self.where = {}     # id(parameter) -> (payload buffer, element offset): where its reduced bf16 gradient lives
        self.layout = {}    # bucket index -> (buffer data_ptr, numel) the mapping above was built for
        # gloo (CPU-tested path / single-GPU boxes) has no bf16 reduction: the bf16-rounded values travel as f32 there
        self.wire_f32 = dist.get_backend(self.group) != "nccl"

    def buffer(self, index, n, device):
        b = self.payload.get(index)
        if b is None or b.numel() != n or b.device != device:
            b = self.payload[index] = torch.empty(n, dtype=torch.bfloat16, device=device)
        return b


def bf16_allreduce_hook(state: Bf16AllreduceState, bucket: dist.GradBucket):
    """DDP communication hook (`ddp.register_comm_hook(state, bf16_allreduce_hook)`); see the module docstring."""
    grads = bucket.buffer()
    if grads.dtype != torch.float32 or not grads.is_cuda:
        raise L.JatError(L.ERR_BAD_ARG, "bf16_allreduce_hook expects f32 CUDA gradient buckets")
    n = grads.numel()
    pay = state.buffer(bucket.index(), n, grads.device)
    lib = L.load()
    L.check(lib.jat_grad_compress(_ctx(grads), grads.data_ptr(), pay.data_ptr(), n, 1.0 / state.world, _stream(grads.device)))
    wire = pay.float() if state.wire_f32 else pay
    fut = dist.all_reduce(wire, group=state.group, async_op=True).get_future()

    if state.optimizer is not None:
        # fused consumer: remember where each parameter's reduced gradient sits in the payload (DDP rebuilds its buckets once,
        # after the first backward pass), leave the f32 bucket alone
        key = (grads.data_ptr(), n, pay.data_ptr())
        if state.layout.get(bucket.index()) != key:
            base = grads.data_ptr()
            for p_, g_ in zip(bucket.parameters(), bucket.gradients()):
                state.where[id(p_)] = (pay, (g_.data_ptr() - base) // 4)
            state.layout[bucket.index()] = key
            state.optimizer._groups = None     # rebuild the pointer table at the next step

        def keep(f):
            if state.wire_f32:
                pay.copy_(f.value()[0])
            return grads
        return fut.then(keep)

    def expand(f):
        if state.wire_f32:
            pay.copy_(f.value()[0])
        L.check(lib.jat_grad_decompress(_ctx(grads), pay.data_ptr(), grads.data_ptr(), n, _stream(grads.device)))
        return grads

    return fut.then(expand)


def register_bf16_allreduce(ddp_model, process_group=None, optimizer=None):
    """Install the bf16 gradient exchange on a `DistributedDataParallel` instance; returns the hook state.
    optimizer: a `jat_b200.FusedAdamW` that will consume the reduced bf16 payload directly (see the module docstring);
    needs `gradient_as_bucket_view=True` (the parameter -> payload map is read off the bucket views)."""
    state = Bf16AllreduceState(process_group, optimizer)
    if optimizer is not None:
        if not hasattr(optimizer, "_grad_source"):
            raise L.JatError(L.ERR_BAD_ARG, "register_bf16_allreduce(optimizer=...) needs a jat_b200.FusedAdamW")
        if not getattr(ddp_model, "gradient_as_bucket_view", False):
            raise L.JatError(L.ERR_BAD_ARG, "the fused consumer needs DistributedDataParallel(..., gradient_as_bucket_view=True)")
        optimizer._grad_source = state
    ddp_model.register_comm_hook(state, bf16_allreduce_hook)
    return state
