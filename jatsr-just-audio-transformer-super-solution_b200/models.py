"""Drop-in mirrors of the reference DiT modules, executing on the sm_100a C-ABI library.

  JaT_AudioSR_V2  <->  reference src/models/jat_audiosr_v2.py:292 (LayerNorm, no affine)
  JaT_AudioSR_V3  <->  reference src/models/jat_audiosr_v3.py:311 (RMSNorm with weight)

Same constructor keywords (jat_audiosr_v2.py:297-308), same ``forward(x_t, t, x_cond)`` contract
(:399-448), same ``state_dict`` keys / shapes / persistent RoPE buffers, and -- because parameters are
created in the reference's order with the same initialisers -- the same random initial weights under
the same ``torch.manual_seed``.  The nn.Module tree below is a parameter container only: the forward
pass is one call into `Engine`, which enqueues hand-written CUDA kernels.  There is no PyTorch or CPU
execution path; calling the model on CPU tensors raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from .engine import Engine


class _Stage(torch.autograd.Function):
    """The forward pass is ONE C-ABI call (`jat_dit_forward_train`, issued by the first stage); the autograd graph,
    however, is a chain of stages  embed -> block 0 -> ... -> block depth-1 -> final  linked by a dummy token, each
    stage taking its own parameters as inputs.  Its backward runs the matching `jat_dit_backward_{end,block,begin}`
    and returns that stage's parameter gradients right away -- so DDP's bucketed all-reduce (train_ddp_v3mod2.py:822)
    overlaps with the backward of the earlier blocks exactly as it does for the reference module, and AdamW /
    GradScaler / clip_grad_norm_ see ordinary .grad tensors.  x_t, t and x_cond get no gradient (the reference never
    differentiates them)."""

    @staticmethod
    def forward(ctx, model, kind, index, token, payload, *params):
        ctx.model, ctx.kind, ctx.index, ctx.params = model, kind, index, params
        if kind == "embed":
            x_t, t, x_cond = payload
            ctx.shape = model._train_shape = (x_t.shape[0], x_t.shape[2], x_t.device)
            # one draw from torch's global generator per step seeds every Dropout / DropPath mask of the step
            # (the reference draws its masks from the same generator, so torch.manual_seed controls both)
            p_drop = float(model.dropout_p)
            stochastic = p_drop > 0.0 or model.drop_path_rate > 0.0
            seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item()) if stochastic else 0
            model._train_seed = seed  # parity tests rebuild the step's masks from it
            model._train_out = model._engine.forward_train(x_t, t, x_cond, dropout_p=p_drop, seed=seed)
            ctx.generation = model._train_generation = model._engine.generation
            return torch.zeros(1, device=x_t.device)
        ctx.shape, ctx.generation = model._train_shape, model._train_generation
        if kind == "final":
            out, model._train_out = model._train_out, None
            return out
        return token.clone()

    @staticmethod
    def backward(ctx, grad):
        model, (B, T, dev) = ctx.model, ctx.shape
        eng = model._engine
        need = [p for p in ctx.params if p.requires_grad]
        view_mode = model.grad_handoff == "view"
        if view_mode:
            packed = eng.grads(dev).by_param
            for p in need:  # a live .grad that IS the packed buffer would be overwritten (and then added to itself)
                if p.grad is not None and p.grad.data_ptr() == packed[p].data_ptr():
                    raise L.JatError(L.ERR_BAD_ARG, "grad_handoff='view': a parameter's .grad from the previous backward is still "
                                     "alive; call zero_grad(set_to_none=True) before every backward (no gradient accumulation), "
                                     "or use grad_handoff='copy'")
        if ctx.kind == "final":
            by_param = eng.backward_begin(grad.float().contiguous(), B, T, ctx.generation)
        elif ctx.kind == "block":
            by_param = eng.backward_block(ctx.index, B, T, dev, ctx.generation)
        else:
            by_param = eng.backward_end(B, T, dev, ctx.generation)
        if view_mode:
            # zero-copy hand-off: fresh view objects of the packed gradient buffers (autograd adopts them as .grad; DDP copies
            # them into its buckets).  Valid until the next backward pass re-uses the buffers.
            outs = [by_param[p].view(p.shape) for p in need]
        else:
            # fresh tensors: autograd / DDP keep (or accumulate into) what is returned, the packed buffers are reused.
            # One flat allocation + one multi-tensor copy per stage instead of a clone per parameter.
            flat = torch.empty(sum(p.numel() for p in need), dtype=torch.float32, device=dev)
            outs = [v.view(p.shape) for v, p in zip(flat.split([p.numel() for p in need]), need)]
            if outs:
                torch._foreach_copy_(outs, [by_param[p] for p in need])
        it = iter(outs)
        grads = tuple(next(it) if p.requires_grad else None for p in ctx.params)
        tok = None if ctx.kind == "embed" else torch.zeros(1, device=dev)
        return (None, None, None, tok, None) + grads


def _forward_train(model, x_t, t, x_cond):
    embed = list(model.patch_embed.parameters()) + list(model.t_embedder.parameters())
    tok = _Stage.apply(model, "embed", -1, None, (x_t, t, x_cond), *embed)
    for i, blk in enumerate(model.blocks):
        tok = _Stage.apply(model, "block", i, tok, None, *blk.parameters())
    return _Stage.apply(model, "final", -1, tok, None, *model.final_layer.parameters())


class TimeEmbedding(nn.Module):
    """Parameter-free placeholder at t_embedder.0 (jat_audiosr_v2.py:170-190); computed by
    `jat_timestep_features`."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim


class RoPE(nn.Module):
    """Persistent RoPE buffers with the reference's names and values (jat_audiosr_v2.py:50-68)."""

    def __init__(self, dim, max_seq_len=4096, base=10000):
        super().__init__()
        self.dim, self.max_seq_len, self.base = dim, max_seq_len, base
        inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim))
        self.register_buffer("inv_freq", inv_freq)
        freqs = torch.outer(torch.arange(max_seq_len).float(), inv_freq)
        emb = torch.cat([freqs, freqs], dim=-1)
        self.register_buffer("cos_cached", emb.cos())
        self.register_buffer("sin_cached", emb.sin())


class GroupedQueryAttention(nn.Module):
    """q/k/v/out projections (no bias) + RoPE buffers (jat_audiosr_v2.py:94-125)."""

    def __init__(self, hidden_size, num_q_heads, num_kv_heads, dropout=0.0):
        super().__init__()
        assert hidden_size % num_q_heads == 0, "hidden_size must be divisible by num_q_heads"
        assert num_q_heads % num_kv_heads == 0, "num_q_heads must be divisible by num_kv_heads"
        self.hidden_size, self.num_q_heads, self.num_kv_heads = hidden_size, num_q_heads, num_kv_heads
        self.num_groups = num_q_heads // num_kv_heads
        self.head_dim = hidden_size // num_q_heads
        self.q_proj = nn.Linear(hidden_size, hidden_size, bias=False)
        kv = num_kv_heads * self.head_dim
        self.k_proj = nn.Linear(hidden_size, kv, bias=False)
        self.v_proj = nn.Linear(hidden_size, kv, bias=False)
        self.out_proj = nn.Linear(hidden_size, hidden_size, bias=False)
        self.dropout = nn.Dropout(dropout)
        self.rope = RoPE(self.head_dim)


class BottleneckPatchEmbed1D(nn.Module):
    def __init__(self, patch_len, in_chans, embed_dim, bottleneck_dim):
        super().__init__()
        self.patch_len = patch_len
        self.flatten_dim = patch_len * in_chans
        self.proj = nn.Sequential(nn.Linear(self.flatten_dim, bottleneck_dim), nn.GELU(),
                                  nn.Linear(bottleneck_dim, embed_dim))


class DropPath(nn.Module):
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob


class DiTBlock_GQA(nn.Module):
    """adaLN-Zero block parameters (jat_audiosr_v2.py:234-263 / jat_audiosr_v3.py:252-282)."""

    def __init__(self, hidden_size, num_q_heads, num_kv_heads, mlp_ratio=4.0, dropout=0.1, drop_path=0.0,
                 rms_norm=False):
        super().__init__()
        mk = (lambda: nn.RMSNorm(hidden_size, eps=1e-6)) if rms_norm else \
            (lambda: nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6))
        self.norm1 = mk()
        self.attn = GroupedQueryAttention(hidden_size, num_q_heads, num_kv_heads, dropout=dropout)
        self.norm2 = mk()
        hid = int(hidden_size * mlp_ratio)
        self.mlp = nn.Sequential(nn.Linear(hidden_size, hid), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hid, hidden_size), nn.Dropout(dropout))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 6 * hidden_size, bias=True))
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()


class _JaTBase(nn.Module):
    norm_kind = L.NORM_LAYERNORM

    def __init__(self, input_channels=1024, cond_channels=1024, patch_len=4, hidden_size=1024, depth=16,
                 num_q_heads=16, num_kv_heads=4, bottleneck_dim=512, mlp_ratio=4.0, dropout=0.1,
                 drop_path_rate=0.0):
        super().__init__()
        self.input_channels, self.cond_channels = input_channels, cond_channels
        self.patch_len, self.hidden_size = patch_len, hidden_size
        self.dropout_p, self.drop_path_rate = dropout, drop_path_rate
        rms = self.norm_kind == L.NORM_RMSNORM
        self.patch_embed = BottleneckPatchEmbed1D(patch_len, input_channels + cond_channels, hidden_size,
                                                  bottleneck_dim)
        self.max_len = 2048
        self.t_embedder = nn.Sequential(TimeEmbedding(hidden_size), nn.Linear(hidden_size, hidden_size), nn.SiLU(),
                                        nn.Linear(hidden_size, hidden_size))
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            DiTBlock_GQA(hidden_size, num_q_heads, num_kv_heads, mlp_ratio, dropout=dropout, drop_path=dpr[i],
                         rms_norm=rms) for i in range(depth)])
        fin_norm = nn.RMSNorm(hidden_size, eps=1e-6) if rms else \
            nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.final_layer = nn.Sequential(fin_norm, nn.Linear(hidden_size, patch_len * input_channels))
        self.initialize_weights()
        # DistributedDataParallel (train_ddp_v3mod2.py:822, default broadcast_buffers=True) re-broadcasts every module buffer
        # from rank 0 at the start of EVERY forward: here that is 84 RoPE tables (58 MB: cat + NCCL broadcast + copy back,
        # measured 2.5 ms of a 53 ms step with the compute stream idle).  The tables are pure functions of (head_dim, base),
        # identical on every rank by construction, so they are put on DDP's ignore list (the attribute DDP itself reads);
        # both spellings, because DDP sees `_orig_mod.`-prefixed names when the module went through torch.compile first.
        rope = [f"blocks.{i}.attn.rope.{n}" for i in range(depth) for n in ("inv_freq", "cos_cached", "sin_cached")]
        self._ddp_params_and_buffers_to_ignore = rope + ["_orig_mod." + n for n in rope]
        object.__setattr__(self, "_engine", Engine(self))  # not a submodule / not in state_dict
        # the engine caches the parameter list (see PackedWeights._versions); load_state_dict(assign=True) replaces parameters
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.__dict__.pop("_param_list", None))
        # How parameter gradients leave the backward pass.  "copy" (default): fresh tensors, ordinary autograd semantics
        # (accumulation over several backward passes, zero_grad(set_to_none=False), .grad kept across steps).  "view": the
        # .grad tensors are views of the library's packed gradient buffers (no 3 GB copy per step); they are valid until the
        # next backward pass, which requires zero_grad(set_to_none=True) before every backward (checked) -- or DDP with
        # gradient_as_bucket_view=True, which moves them into its buckets right away.
        object.__setattr__(self, "grad_handoff", "copy")

    def initialize_weights(self):
        """adaLN-Zero: zero the modulation and final projections (jat_audiosr_v2.py:372-381)."""
        for block in self.blocks:
            nn.init.constant_(block.adaLN_modulation[-1].weight, 0)
            nn.init.constant_(block.adaLN_modulation[-1].bias, 0)
        nn.init.constant_(self.final_layer[-1].weight, 0)
        nn.init.constant_(self.final_layer[-1].bias, 0)

    # ------------------------------------------------------------------------------------ forward
    def _check_inputs(self, x_t, t, x_cond):
        if not (x_t.is_cuda and x_cond.is_cuda and t.is_cuda):
            raise RuntimeError("jat_b200 models run on CUDA (sm_100a) tensors only; there is no CPU fallback")
        if (x_t.dim() != 3 or x_cond.dim() != 3 or x_t.shape[0] != x_cond.shape[0] or x_t.shape[2] != x_cond.shape[2]
                or x_t.shape[1] != self.input_channels or x_cond.shape[1] != self.cond_channels):
            raise ValueError(f"expected x_t [B, {self.input_channels}, T] and x_cond [B, {self.cond_channels}, T], got "
                             f"{tuple(x_t.shape)} / {tuple(x_cond.shape)}")
        if t.dim() != 1 or t.shape[0] != x_t.shape[0]:
            raise ValueError("t must be [B]")
        if not 0.0 <= self.dropout_p < 1.0:
            raise ValueError(f"dropout probability has to be in [0, 1), got {self.dropout_p}")
        N = (x_t.shape[-1] + self.patch_len - 1) // self.patch_len
        if N > self.max_len:
            raise ValueError(f"Sequence length {N} exceeds max_len {self.max_len}")

    @torch.compiler.disable
    def forward(self, x_t, t, x_cond):
        """x_t, x_cond [B, C, T]; t [B] in [0, 1] -> x_pred [B, C, T] (jat_audiosr_v2.py:399-448).

        `torch.compile(model)` (train_ddp_v3mod2.py:816) is part of the reference's module protocol: the forward is
        excluded from dynamo tracing (`torch.compiler.disable`) -- there is nothing for a tracing compiler to fuse, the
        whole pass is already one C-ABI call enqueueing hand-written kernels -- so the compiled wrapper is a transparent
        no-op that keeps `_orig_mod.`-prefixed state dicts, DDP wrapping and autograd working as in the reference."""
        self._check_inputs(x_t, t, x_cond)
        xt, tt, xc = x_t.float().contiguous(), t.float().contiguous(), x_cond.float().contiguous()
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            out = _forward_train(self, xt, tt, xc)
        else:
            out = self._engine.forward(xt, tt, xc)
        if torch.is_autocast_enabled():
            out = out.to(torch.get_autocast_dtype("cuda"))
        return out

    @torch.no_grad()
    def forward_with_blocks(self, x_t, t, x_cond):
        """Parity-test helper: (x_pred, per-block residual stream f32 [depth, B*N, D])."""
        self._check_inputs(x_t, t, x_cond)
        return self._engine.forward(x_t.float().contiguous(), t.float().contiguous(), x_cond.float().contiguous(),
                                    keep_blocks=True)

    def refresh_packed_weights(self):
        """Re-cast the parameters into the packed device copies before the next forward.  Needed only after in-place
        writes that autograd's version counters do not see and that were not followed by a backward pass (e.g.
        `p.data.copy_(...)` from an EMA shadow); optimizer steps, `load_state_dict` and `.to()` are detected."""
        if self._engine.packed is not None:
            self._engine.packed.dirty = True

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self.__dict__.pop("_param_list", None)
        if getattr(self, "_engine", None) is not None:
            self._engine.packed = None  # .to()/.half()/... invalidates packed device copies
        return r


class JaT_AudioSR_V2(_JaTBase):
    """LayerNorm variant (the class `train_ddp_v3mod2.py` trains)."""
    norm_kind = L.NORM_LAYERNORM


class JaT_AudioSR_V3(_JaTBase):
    """RMSNorm variant (the class `infer_test_v3m2.py` / `train_ddp_v3m2.py` use)."""
    norm_kind = L.NORM_RMSNORM
