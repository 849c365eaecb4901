"""Host-side engine behind the nn.Module facade: weight packing, workspaces and the launch plan.

All device computation is done by libjat_b200.so through `jat_dit_modulation` /
`jat_dit_forward_tokens` (see include/jat_b200.h); torch only owns the memory.

HBM layout (M = B * N token rows, N = ceil(T / patch_len)):
  weights   bf16, nn.Linear [out, in] row-major = K-major B operand of the tcgen05 GEMM;
            q/k/v projections concatenated row-wise into one [(Hq+2Hkv)*64, D] matrix per block;
            every block's adaLN_modulation.1 stacked into one [depth*6D, D] matrix (one GEMM for all
            blocks' shift/scale/gate); biases / norm weights / RoPE tables f32.
  patches   bf16 [M, 2*C*P]   pe_hid bf16 [M, bottleneck]   x f32 [M, D] (residual stream)
  h         bf16 [M, D]       qkv   bf16 [M, (Hq+2Hkv)*64]  attn bf16 [M, D]   mlp_hid bf16 [M, 4D]
  mod       f32 [Bt, depth*6D]  (Bt = batch for a plain forward, = num_steps for the sampler)
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr


class PackedWeights:
    """bf16/f32 device copies of a model's parameters in the layout the C-ABI expects."""

    def __init__(self, model, device):
        self.device = device
        self.versions = self._versions(model)
        # True once a backward pass has produced gradients from these copies: an optimizer step probably follows, and not
        # every in-place update bumps `Tensor._version` (torch's fused AdamW / `p.data` writes do not), so the next
        # forward re-casts unless the updater (jat_b200.FusedAdamW) has written the copies itself and cleared the flag
        self.dirty = False
        self._build(model, device)
        self._map_sources(model)

    def _map_sources(self, model):
        """(bf16 destination view, parameter) pairs and (f32 destination, parameter) pairs: `refresh` re-casts every
        parameter into the SAME device buffers with two multi-tensor copies (pointers in `struct` stay valid), instead of
        re-allocating ~400 tensors after every optimizer step."""
        k = self.keep
        D = model.hidden_size
        attn0 = model.blocks[0].attn
        qd, kd = attn0.q_proj.out_features, attn0.k_proj.out_features
        pe, te = model.patch_embed.proj, model.t_embedder
        pairs = [(k["pe_w1"], pe[0].weight), (k["pe_b1"], pe[0].bias), (k["pe_w2"], pe[2].weight), (k["pe_b2"], pe[2].bias),
                 (k["te_w1"], te[1].weight), (k["te_b1"], te[1].bias), (k["te_w2"], te[3].weight), (k["te_b2"], te[3].bias),
                 (k["final_w"], model.final_layer[1].weight), (k["final_b"], model.final_layer[1].bias)]
        for i, b in enumerate(model.blocks):
            pairs += [(k["ada_w"][i * 6 * D:(i + 1) * 6 * D], b.adaLN_modulation[1].weight),
                      (k["ada_b"][i * 6 * D:(i + 1) * 6 * D], b.adaLN_modulation[1].bias),
                      (k["wqkv"][i][:qd], b.attn.q_proj.weight), (k["wqkv"][i][qd:qd + kd], b.attn.k_proj.weight),
                      (k["wqkv"][i][qd + kd:], b.attn.v_proj.weight), (k["wo"][i], b.attn.out_proj.weight),
                      (k["w1"][i], b.mlp[0].weight), (k["b1"][i], b.mlp[0].bias), (k["w2"][i], b.mlp[3].weight),
                      (k["b2"][i], b.mlp[3].bias)]
            if "n1" in k:
                pairs += [(k["n1"][i], b.norm1.weight), (k["n2"][i], b.norm2.weight)]
        if "nf" in k:
            pairs.append((k["nf"], model.final_layer[0].weight))
        self._dst = [d for d, _ in pairs]
        self._src = [p_ for _, p_ in pairs]

    @torch.no_grad()
    def refresh(self, model):
        """Re-cast the (updated) parameters into the existing packed buffers."""
        torch._foreach_copy_(self._dst, [p_.detach() for p_ in self._src])
        self.versions = self._versions(model)
        self.dirty = False

    def _build(self, model, device):
        bf = lambda t: t.detach().to(device=device, dtype=torch.bfloat16).contiguous()
        f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
        D, depth = model.hidden_size, len(model.blocks)
        keep = self.keep = {}
        keep["pe_w1"], keep["pe_b1"] = bf(model.patch_embed.proj[0].weight), f32(model.patch_embed.proj[0].bias)
        keep["pe_w2"], keep["pe_b2"] = bf(model.patch_embed.proj[2].weight), f32(model.patch_embed.proj[2].bias)
        keep["te_w1"], keep["te_b1"] = bf(model.t_embedder[1].weight), f32(model.t_embedder[1].bias)
        keep["te_w2"], keep["te_b2"] = bf(model.t_embedder[3].weight), f32(model.t_embedder[3].bias)
        keep["ada_w"] = bf(torch.cat([b.adaLN_modulation[1].weight.detach() for b in model.blocks], 0))
        keep["ada_b"] = f32(torch.cat([b.adaLN_modulation[1].bias.detach() for b in model.blocks], 0))
        keep["wqkv"] = [bf(torch.cat([b.attn.q_proj.weight.detach(), b.attn.k_proj.weight.detach(),
                                      b.attn.v_proj.weight.detach()], 0)) for b in model.blocks]
        keep["wo"] = [bf(b.attn.out_proj.weight) for b in model.blocks]
        keep["w1"] = [bf(b.mlp[0].weight) for b in model.blocks]
        keep["b1"] = [f32(b.mlp[0].bias) for b in model.blocks]
        keep["w2"] = [bf(b.mlp[3].weight) for b in model.blocks]
        keep["b2"] = [f32(b.mlp[3].bias) for b in model.blocks]
        rms = model.norm_kind == L.NORM_RMSNORM
        if rms:
            keep["n1"] = [f32(b.norm1.weight) for b in model.blocks]
            keep["n2"] = [f32(b.norm2.weight) for b in model.blocks]
            keep["nf"] = f32(model.final_layer[0].weight)
        keep["final_w"], keep["final_b"] = bf(model.final_layer[1].weight), f32(model.final_layer[1].bias)
        rope = model.blocks[0].attn.rope
        keep["cos"], keep["sin"] = f32(rope.cos_cached), f32(rope.sin_cached)

        self.arrays = {k: _ptr_array(keep[k]) for k in ("wqkv", "wo", "w1", "b1", "w2", "b2")}
        if rms:
            self.arrays["n1"], self.arrays["n2"] = _ptr_array(keep["n1"]), _ptr_array(keep["n2"])
        w = self.struct = L.DitWeights()
        attn0 = model.blocks[0].attn
        w.hidden, w.depth = D, depth
        w.n_q_heads, w.n_kv_heads, w.head_dim = attn0.num_q_heads, attn0.num_kv_heads, attn0.head_dim
        w.mlp_hidden = model.blocks[0].mlp[0].out_features
        w.bottleneck = model.patch_embed.proj[0].out_features
        w.channels, w.patch_len = model.input_channels, model.patch_len
        w.cond_channels = model.cond_channels
        w.norm_kind, w.max_len, w.rope_max_pos = model.norm_kind, model.max_len, rope.max_seq_len
        w.norm_eps = 1e-6
        p = lambda t: t.data_ptr()
        w.pe_w1, w.pe_b1, w.pe_w2, w.pe_b2 = p(keep["pe_w1"]), p(keep["pe_b1"]), p(keep["pe_w2"]), p(keep["pe_b2"])
        w.te_w1, w.te_b1, w.te_w2, w.te_b2 = p(keep["te_w1"]), p(keep["te_b1"]), p(keep["te_w2"]), p(keep["te_b2"])
        w.ada_w, w.ada_b = p(keep["ada_w"]), p(keep["ada_b"])
        cast = lambda a: C.cast(a, C.POINTER(C.c_void_p))
        w.wqkv, w.wo = cast(self.arrays["wqkv"]), cast(self.arrays["wo"])
        w.w1, w.b1, w.w2, w.b2 = (cast(self.arrays[k]) for k in ("w1", "b1", "w2", "b2"))
        if rms:
            w.norm1_w, w.norm2_w, w.final_norm_w = cast(self.arrays["n1"]), cast(self.arrays["n2"]), p(keep["nf"])
        w.final_w, w.final_b = p(keep["final_w"]), p(keep["final_b"])
        w.rope_cos, w.rope_sin = p(keep["cos"]), p(keep["sin"])

    @staticmethod
    def _versions(model):
        # (the module-tree walk of model.parameters() costs more than the 400 version reads: the list is cached on the model
        #  and dropped by _apply / load_state_dict(assign=True)-style re-registration through the id check below)
        plist = model.__dict__.get("_param_list")
        if plist is None or len(plist) != model.__dict__.get("_param_count", -1):
            plist = list(model.parameters())
            model.__dict__["_param_list"], model.__dict__["_param_count"] = plist, len(plist)
        return tuple((id(p), p._version, p.device) for p in plist)

    def stale(self, model, device):
        return device != self.device or self._versions(model) != self.versions


class PackedGrads:
    """f32 gradient buffers in the packed layout of PackedWeights (q|k|v rows concatenated, adaLN stacked) plus the
    map back to the model's parameters: every parameter's gradient is a contiguous row-slice VIEW of a packed buffer,
    so `jat_dit_backward` writes straight into what autograd hands to the optimizer / DDP."""

    def __init__(self, model, device):
        D, depth = model.hidden_size, len(model.blocks)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=device)
        like = lambda p: z(*p.shape)
        attn0 = model.blocks[0].attn
        qd, kd = attn0.q_proj.out_features, attn0.k_proj.out_features
        k = self.keep = {}
        pe, te = model.patch_embed.proj, model.t_embedder
        k["pe_w1"], k["pe_b1"], k["pe_w2"], k["pe_b2"] = like(pe[0].weight), like(pe[0].bias), like(pe[2].weight), like(pe[2].bias)
        k["te_w1"], k["te_b1"], k["te_w2"], k["te_b2"] = like(te[1].weight), like(te[1].bias), like(te[3].weight), like(te[3].bias)
        k["ada_w"], k["ada_b"] = z(depth * 6 * D, D), z(depth * 6 * D)
        k["wqkv"] = [z(qd + 2 * kd, D) for _ in range(depth)]
        k["wo"] = [like(b.attn.out_proj.weight) for b in model.blocks]
        k["w1"] = [like(b.mlp[0].weight) for b in model.blocks]
        k["b1"] = [like(b.mlp[0].bias) for b in model.blocks]
        k["w2"] = [like(b.mlp[3].weight) for b in model.blocks]
        k["b2"] = [like(b.mlp[3].bias) for b in model.blocks]
        rms = model.norm_kind == L.NORM_RMSNORM
        if rms:
            k["n1"] = [z(D) for _ in range(depth)]
            k["n2"] = [z(D) for _ in range(depth)]
            k["nf"] = z(D)
        k["final_w"], k["final_b"] = like(model.final_layer[1].weight), like(model.final_layer[1].bias)
        # parameter -> gradient view
        g = self.by_param = {}
        g[pe[0].weight], g[pe[0].bias], g[pe[2].weight], g[pe[2].bias] = k["pe_w1"], k["pe_b1"], k["pe_w2"], k["pe_b2"]
        g[te[1].weight], g[te[1].bias], g[te[3].weight], g[te[3].bias] = k["te_w1"], k["te_b1"], k["te_w2"], k["te_b2"]
        for i, b in enumerate(model.blocks):
            g[b.adaLN_modulation[1].weight] = k["ada_w"][i * 6 * D:(i + 1) * 6 * D]
            g[b.adaLN_modulation[1].bias] = k["ada_b"][i * 6 * D:(i + 1) * 6 * D]
            g[b.attn.q_proj.weight] = k["wqkv"][i][:qd]
            g[b.attn.k_proj.weight] = k["wqkv"][i][qd:qd + kd]
            g[b.attn.v_proj.weight] = k["wqkv"][i][qd + kd:]
            g[b.attn.out_proj.weight] = k["wo"][i]
            g[b.mlp[0].weight], g[b.mlp[0].bias], g[b.mlp[3].weight], g[b.mlp[3].bias] = k["w1"][i], k["b1"][i], k["w2"][i], k["b2"][i]
            if rms:
                g[b.norm1.weight], g[b.norm2.weight] = k["n1"][i], k["n2"][i]
        if rms:
            g[model.final_layer[0].weight] = k["nf"]
        g[model.final_layer[1].weight], g[model.final_layer[1].bias] = k["final_w"], k["final_b"]
        self.arrays = {n: _ptr_array(k[n]) for n in ("wqkv", "wo", "w1", "b1", "w2", "b2")}
        if rms:
            self.arrays["n1"], self.arrays["n2"] = _ptr_array(k["n1"]), _ptr_array(k["n2"])
        w = self.struct = L.DitWeights()
        p = lambda t: t.data_ptr()
        w.pe_w1, w.pe_b1, w.pe_w2, w.pe_b2 = p(k["pe_w1"]), p(k["pe_b1"]), p(k["pe_w2"]), p(k["pe_b2"])
        w.te_w1, w.te_b1, w.te_w2, w.te_b2 = p(k["te_w1"]), p(k["te_b1"]), p(k["te_w2"]), p(k["te_b2"])
        w.ada_w, w.ada_b = p(k["ada_w"]), p(k["ada_b"])
        cast = lambda a: C.cast(a, C.POINTER(C.c_void_p))
        w.wqkv, w.wo = cast(self.arrays["wqkv"]), cast(self.arrays["wo"])
        w.w1, w.b1, w.w2, w.b2 = (cast(self.arrays[n]) for n in ("w1", "b1", "w2", "b2"))
        if rms:
            w.norm1_w, w.norm2_w, w.final_norm_w = cast(self.arrays["n1"]), cast(self.arrays["n2"]), p(k["nf"])
        w.final_w, w.final_b = p(k["final_w"]), p(k["final_b"])

    def zero_(self):
        if getattr(self, "_all", None) is None:
            self._all = [t for v in self.keep.values() for t in (v if isinstance(v, list) else [v])]
        torch._foreach_zero_(self._all)


class TrainBuffers:
    """Saved activations + backward scratch + the activation workspace of one (B, T) training problem.  The workspace is
    its own (not the eval path's): the backward reads mod / x / h / patches / pe_hid / t_* from it, so a no-grad forward of
    the same shape between forward and backward (EMA / teacher / self-conditioning pass) must not touch it.  `generation`
    identifies the forward pass whose activations the buffers currently hold."""

    def __init__(self, model, B, T, device):
        self.ws = Workspace(model, B, T, B, device)
        self.generation = 0
        D, depth = model.hidden_size, len(model.blocks)
        P, Cc = model.patch_len, model.input_channels
        N = (T + P - 1) // P
        M = B * N
        attn0 = model.blocks[0].attn
        Hq, Hkv = attn0.num_q_heads, attn0.num_kv_heads
        qkv = (Hq + 2 * Hkv) * attn0.head_dim
        F = model.blocks[0].mlp[0].out_features
        BD = model.patch_embed.proj[0].out_features
        bf = lambda *s: torch.empty(*s, dtype=torch.bfloat16, device=device)
        f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=device)
        self.saved = dict(x_in=f32(depth, M, D), x_mid=f32(depth, M, D), h1=bf(depth, M, D), qkv=bf(depth, M, qkv),
                          attn=bf(depth, M, D), lse=f32(depth, B, Hq, N), y1=bf(depth, M, D), h2=bf(depth, M, D),
                          u=bf(depth, M, F), mact=bf(depth, M, F), y2=bf(depth, M, D), pe_u=bf(M, BD), t_u1=bf(B, D),
                          t_u2=bf(B, D), rs1=f32(depth, M, 2), rs2=f32(depth, M, 2))
        NM = depth * 6 * D
        self.scratch = dict(dx=f32(M, D), dy=bf(M, D), dh=bf(M, D), da=bf(M, D), du=bf(M, F), dqkv=bf(M, qkv),
                            dsum=f32(B, Hq, N), dq_acc=f32(M, Hq * attn0.head_dim), dmod=f32(depth, B, 6 * D),
                            dmod_bf16=bf(depth, B, 6 * D), dt_acc=f32(B, D), rowstats=f32(M, 2),
                            dxsum=f32(B, D), dout_p=bf(M, Cc * P), dpe=bf(M, BD), dt_a=bf(B, D), dt_b=bf(B, D))
        # train-mode regularisers: per-block DropPath rates (the reference's linspace, jat_audiosr_v2.py:351) and factors
        rates = [float(getattr(b.drop_path, "drop_prob", 0.0)) for b in model.blocks]
        self.dp_rates = torch.tensor(rates, dtype=torch.float32, device=device) if any(r > 0 for r in rates) else None
        self.dp_scale = f32(depth, 2, B) if self.dp_rates is not None else None
        self.sv, self.sc = L.DitSaved(), L.DitBwdScratch()
        for k_, v in self.saved.items():
            setattr(self.sv, k_, v.data_ptr())
        self.sv.drop_path_rates = self.dp_rates.data_ptr() if self.dp_rates is not None else None
        self.sv.dp_scale = self.dp_scale.data_ptr() if self.dp_rates is not None else None
        for k_, v in self.scratch.items():
            setattr(self.sc, k_, v.data_ptr())


class Workspace:
    """Activation buffers for one (B, T, Bt) problem size."""

    def __init__(self, model, B, T, Bt, device, keep_blocks=False):
        w = model
        D = w.hidden_size
        P, Cc = w.patch_len, w.input_channels
        N = (T + P - 1) // P
        M = B * N
        attn0 = w.blocks[0].attn
        qkv = (attn0.num_q_heads + 2 * attn0.num_kv_heads) * attn0.head_dim
        F = w.blocks[0].mlp[0].out_features
        depth = len(w.blocks)
        bf = lambda *s: torch.empty(*s, dtype=torch.bfloat16, device=device)
        f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=device)
        self.B, self.T, self.Bt, self.N, self.M = B, T, Bt, N, M
        self.buf = dict(
            patches=bf(M, (Cc + w.cond_channels) * P), pe_hid=bf(M, w.patch_embed.proj[0].out_features), x=f32(M, D), h=bf(M, D),
            qkv=bf(M, qkv), attn=bf(M, D), mlp_hid=bf(M, F),
            t_feat=bf(Bt, D), t_hid=bf(Bt, D), t_act=bf(Bt, D), mod=f32(Bt, depth * 6 * D),
        )
        self.block_out = f32(depth, M, D) if keep_blocks else None
        # sequences of more than 352 tokens: the attention runs one pass per 352-key chunk and merges the partial results
        passes = L.load().jat_attention_passes(N)
        self.attn_part = bf(passes, M, D) if passes > 1 else None
        self.lse_part = f32(passes, B, attn0.num_q_heads, N) if passes > 1 else None
        s = self.struct = L.DitWorkspace()
        for k, v in self.buf.items():
            setattr(s, k, v.data_ptr())
        s.block_out = self.block_out.data_ptr() if keep_blocks else None
        s.attn_part = self.attn_part.data_ptr() if passes > 1 else None
        s.lse_part = self.lse_part.data_ptr() if passes > 1 else None
        self.mod = self.buf["mod"]


class Engine:
    """Per-model engine: caches packed weights and workspaces, issues the C-ABI calls."""

    def __init__(self, model):
        self.model = model
        self.packed = None
        self.workspaces = {}

    def weights(self, device):
        device = self._resolve(device)
        if self.packed is None or self.packed.device != device:
            self.packed = PackedWeights(self.model, device)
        elif self.packed.dirty or self.packed.stale(self.model, device):
            same_objects = [id(p) for p in self.model.parameters()] == [v[0] for v in self.packed.versions]
            if same_objects and all(p.device == device for p in self.model.parameters()):
                self.packed.refresh(self.model)   # same buffers, same pointers: plans and captured graphs stay valid
            else:
                self.packed = PackedWeights(self.model, device)
        return self.packed

    def workspace(self, B, T, Bt, device, keep_blocks=False):
        device = self._resolve(device)
        key = (B, T, Bt, device, keep_blocks)
        ws = self.workspaces.get(key)
        if ws is None:
            if len(self.workspaces) >= 4:
                self.workspaces.clear()
            ws = self.workspaces[key] = Workspace(self.model, B, T, Bt, device, keep_blocks)
        return ws

    @staticmethod
    def _dev_index(device):
        return device.index if device.index is not None else torch.cuda.current_device()

    @staticmethod
    def _resolve(device):
        """torch.device('cuda') -> torch.device('cuda', current): cache keys compare indexed devices only."""
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        return device

    def modulation(self, ws, t):
        """t f32 [Bt] (device) -> ws.mod [Bt, depth*6D]."""
        dev = t.device
        lib, ctx = L.load(), L.context(self._dev_index(dev))
        pw = self.weights(dev)
        L.check(lib.jat_dit_modulation(ctx, C.byref(pw.struct), C.byref(ws.struct), t.data_ptr(), t.shape[0],
                                       torch.cuda.current_stream(dev).cuda_stream))
        return ws.mod

    _det_mode = {}   # device index -> deterministic flag last pushed into that device's context

    @classmethod
    def _sync_determinism(cls, idx):
        """torch.use_deterministic_algorithms(True) selects the bit-reproducible GEMM schedule (no tail split); switching it
        off again restores the library default (JAT_GEMM_TAIL or 2, see jat_create)."""
        det = torch.are_deterministic_algorithms_enabled()
        if cls._det_mode.get(idx, False) != det:
            import os
            mode = 0 if det else int(os.environ.get("JAT_GEMM_TAIL", "2"))
            L.check(L.load().jat_set_gemm_tail_split(L.context(idx), mode))
            cls._det_mode[idx] = det

    def forward_tokens(self, ws, x_t, x_cond, B, mod, mod_batch_stride, out, cond_batch=None):
        dev = x_t.device
        self._sync_determinism(self._dev_index(dev))
        lib, ctx = L.load(), L.context(self._dev_index(dev))
        pw = self.weights(dev)
        T = x_t.shape[-1]
        cb = 0 if x_cond is None else (x_cond.shape[0] if cond_batch is None else cond_batch)
        code = lib.jat_dit_forward_tokens(ctx, C.byref(pw.struct), C.byref(ws.struct), x_t.data_ptr(), x_t.shape[0],
                                          0 if x_cond is None else x_cond.data_ptr(), cb, mod.data_ptr(),
                                          mod_batch_stride, out.data_ptr(), B, T,
                                          torch.cuda.current_stream(dev).cuda_stream)
        if code == L.ERR_SEQ_TOO_LONG:  # reference raises ValueError (jat_audiosr_v2.py:428-429)
            raise ValueError(lib.jat_last_error().decode())
        L.check(code)
        return out

    # ------------------------------------------------------------------------------------ training step
    def train_buffers(self, B, T, device):
        key = (B, T, self._resolve(device))
        tb = getattr(self, "_train", {}).get(key)
        if tb is None:
            self._train = {key: TrainBuffers(self.model, B, T, device)}
            tb = self._train[key]
        return tb

    def grads(self, device):
        if getattr(self, "_grads", None) is None or self._grads.keep["pe_b1"].device != device:
            self._grads = PackedGrads(self.model, device)
        return self._grads

    def forward_train(self, x_t, t, x_cond, dropout_p=0.0, seed=0):
        """dropout_p / seed: this step's train-mode Dropout probability and mask seed (DropPath rates come from the
        blocks); they are stored next to the saved activations so that the backward regenerates the same masks."""
        B, Cc, T = x_t.shape
        dev = x_t.device
        self._sync_determinism(self._dev_index(dev))
        lib, ctx = L.load(), L.context(self._dev_index(dev))
        pw = self.weights(dev)
        tb = self.train_buffers(B, T, dev)
        ws = tb.ws
        self._generation = getattr(self, "_generation", 0) + 1
        tb.generation = self._generation   # the saved activations now belong to THIS forward (checked by every backward stage)
        tb.sv.dropout_p, tb.sv.seed = float(dropout_p), int(seed)
        out = torch.empty(B, Cc, T, dtype=torch.float32, device=dev)
        code = lib.jat_dit_forward_train(ctx, C.byref(pw.struct), C.byref(ws.struct), C.byref(tb.sv), x_t.data_ptr(),
                                         x_cond.data_ptr(), t.data_ptr(), out.data_ptr(), B, T,
                                         torch.cuda.current_stream(dev).cuda_stream)
        if code == L.ERR_SEQ_TOO_LONG:
            raise ValueError(lib.jat_last_error().decode())
        L.check(code)
        return out

    @property
    def generation(self):
        """Id of the latest forward_train (0 = none yet); a backward stage passes the id of ITS forward."""
        return getattr(self, "_generation", 0)

    def _bwd_args(self, dev, B, T, generation=None):
        pw = self.packed
        tb = getattr(self, "_train", {}).get((B, T, self._resolve(dev)))
        if pw is None or pw.device != self._resolve(dev) or tb is None or tb.generation == 0:
            raise L.JatError(L.ERR_BAD_ARG, "backward without a matching forward_train on this device")
        if generation is not None and tb.generation != generation:
            # single-slot training state: the activations this backward needs were overwritten by a later train-mode forward
            raise L.JatError(L.ERR_BAD_ARG, "backward of a stale forward: another train-mode forward ran on this model before "
                             "this backward (the saved activations are single-slot); run forward -> backward pairs in order, "
                             "or wrap additional forwards in torch.no_grad() / model.eval()")
        pw.dirty = True
        gr = self.grads(dev)
        return (L.context(self._dev_index(dev)), C.byref(pw.struct), C.byref(tb.ws.struct), C.byref(tb.sv), C.byref(tb.sc),
                C.byref(gr.struct)), gr

    def backward_begin(self, d_out, B, T, generation=None):
        """Stage 1 of the backward pass (final layer); zeroes the packed gradient buffers first."""
        dev = d_out.device
        args, gr = self._bwd_args(dev, B, T, generation)
        gr.zero_()
        L.check(L.load().jat_dit_backward_begin(*args, d_out.data_ptr(), B, T, torch.cuda.current_stream(dev).cuda_stream))
        return gr.by_param

    def backward_block(self, i, B, T, dev, generation=None):
        args, gr = self._bwd_args(dev, B, T, generation)
        L.check(L.load().jat_dit_backward_block(*args, i, B, T, torch.cuda.current_stream(dev).cuda_stream))
        return gr.by_param

    def backward_end(self, B, T, dev, generation=None):
        args, gr = self._bwd_args(dev, B, T, generation)
        L.check(L.load().jat_dit_backward_end(*args, B, T, torch.cuda.current_stream(dev).cuda_stream))
        return gr.by_param

    def backward(self, d_out, B, T):
        """Whole backward in one call: d_out f32 [B, C, T] -> {parameter: f32 gradient view}; consumes the saved
        activations of the last forward_train with the same (B, T)."""
        dev = d_out.device
        args, gr = self._bwd_args(dev, B, T)
        gr.zero_()
        L.check(L.load().jat_dit_backward(*args, d_out.data_ptr(), B, T, torch.cuda.current_stream(dev).cuda_stream))
        return gr.by_param

    def forward(self, x_t, t, x_cond, keep_blocks=False):
        """Plain model forward: per-sample t (jat_audiosr_v2.py:399-448)."""
        B, Cc, T = x_t.shape
        dev = x_t.device
        ws = self.workspace(B, T, B, dev, keep_blocks)
        self.modulation(ws, t)
        out = torch.empty(B, Cc, T, dtype=torch.float32, device=dev)
        self.forward_tokens(ws, x_t, x_cond, B, ws.mod, ws.mod.shape[1], out)
        return (out, ws.block_out) if keep_blocks else out
