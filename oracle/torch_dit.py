"""Differentiable torch restatement of the reference DiT forward WITH injectable train-mode masks.

TEST INFRASTRUCTURE (an oracle, like dit_oracle.py): imported only by tests/ (as tests/_torch_dit.py), and by bench.py's
baseline legs (`cpu_baseline` / `--impl reference` on the host cores through ATen -- the library the reference's own CPU
path executes in -- and `gpu_baseline`, the same torch code on the B200 as the stock-PyTorch comparator).  Never imported by
the product package.  Pinned against the unmodified reference modules and sampler by tests/test_oracle.py.

Follows src/models/jat_audiosr_v2.py:399-448 (forward), :265-289 (block), :127-167 (GQA attention), :50-91 (RoPE),
:177-190 (time embedding), :212-231 (patch embed), :21-34 (drop_path) of the reference; `rms=True` switches to the
jat_audiosr_v3.py norms.  `masks` replaces the reference's random draws by given multiplier tensors so that the CUDA
path's counter-based masks can be replayed:
    masks["attn"][i]  [B, Hq, N, N]   multiplier on the softmax probabilities of block i     (nn.Dropout, :158)
    masks["hid"][i]   [B*N, F]        multiplier after the MLP's GELU                        (nn.Dropout, :250)
    masks["out"][i]   [B*N, D]        multiplier after mlp.3                                 (nn.Dropout, :252)
    masks["path"]     [depth, 2, B]   DropPath factor on gate * branch (0 = attention, 1 = MLP)  (:281, :287)
tests/test_oracle.py pins this file against the unmodified reference (eval mode, and train mode with the reference's
Dropout / drop_path replaced by the same masks)."""
import math

import torch
import torch.nn.functional as F


def _norm(x, w, rms, eps=1e-6):
    if rms:
        return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w
    return F.layer_norm(x, (x.shape[-1],), eps=eps)


def _rope(x, cos, sin):
    # x [B, N, H, 64]; cos/sin [N, 64]
    c, s = cos[None, :, None, :], sin[None, :, None, :]
    x1, x2 = x[..., :32], x[..., 32:]
    return x * c + torch.cat([-x2, x1], -1) * s


def block_forward(x, t_emb, w, cos, sin, Hq, Hkv, rms, m_attn=None, m_hid=None, m_out=None, m_path=None):
    """One DiTBlock_GQA (jat_audiosr_v2.py:265-289).  w = (adaLN w, adaLN b, q, k, v, out_proj, mlp.0 w, mlp.0 b, mlp.3 w,
    mlp.3 b, norm1 w | None, norm2 w | None); tensors only, so one `torch.compile` of this function serves every block."""
    ada_w, ada_b, wq, wk, wv, wo, w1, b1, w2, b2, n1, n2 = w
    B, N, D = x.shape
    G = Hq // Hkv
    mod = F.silu(t_emb) @ ada_w.T + ada_b
    sh1, sc1, g1, sh2, sc2, g2 = mod.chunk(6, 1)
    h = _norm(x, n1, rms) * (1 + sc1[:, None]) + sh1[:, None]
    q = (h @ wq.T).view(B, N, Hq, 64)
    kk = (h @ wk.T).view(B, N, Hkv, 64)
    v = (h @ wv.T).view(B, N, Hkv, 64)
    q, kk = _rope(q, cos, sin), _rope(kk, cos, sin)
    kk, v = kk.repeat_interleave(G, 2), v.repeat_interleave(G, 2)
    q, kk, v = q.transpose(1, 2), kk.transpose(1, 2), v.transpose(1, 2)
    a = torch.softmax(q @ kk.transpose(-2, -1) / 8.0, -1)
    if m_attn is not None:
        a = a * m_attn
    a = (a @ v).transpose(1, 2).reshape(B, N, D) @ wo.T
    br = g1[:, None] * a
    if m_path is not None:
        br = br * m_path[0][:, None, None]
    x = x + br
    h = _norm(x, n2, rms) * (1 + sc2[:, None]) + sh2[:, None]
    u = F.gelu(h @ w1.T + b1)
    if m_hid is not None:
        u = u * m_hid.view(B, N, -1)
    y = u @ w2.T + b2
    if m_out is not None:
        y = y * m_out.view(B, N, -1)
    br = g2[:, None] * y
    if m_path is not None:
        br = br * m_path[1][:, None, None]
    return x + br


def block_weights(p, i):
    k = f"blocks.{i}."
    return (p[k + "adaLN_modulation.1.weight"], p[k + "adaLN_modulation.1.bias"], p[k + "attn.q_proj.weight"],
            p[k + "attn.k_proj.weight"], p[k + "attn.v_proj.weight"], p[k + "attn.out_proj.weight"], p[k + "mlp.0.weight"],
            p[k + "mlp.0.bias"], p[k + "mlp.3.weight"], p[k + "mlp.3.bias"], p.get(k + "norm1.weight"), p.get(k + "norm2.weight"))


def dit_forward(p, cfg, x_t, t, x_cond, rms=False, masks=None, blocks_out=None, block_fn=None):
    """p: {state_dict key: tensor (requires_grad ok)}; returns x_pred [B, C, T].  blocks_out: optional list that receives
    the residual stream [B*N, D] after every block (what forward hooks on model.blocks[i] see in the reference).
    block_fn: replacement for `block_forward` with the same signature (bench.py passes `torch.compile(block_forward)`)."""
    D, depth, P = cfg["hidden_size"], cfg["depth"], cfg["patch_len"]
    Hq, Hkv = cfg["num_q_heads"], cfg["num_kv_heads"]
    B, C, T = x_t.shape
    pad = (P - T % P) % P
    x = torch.cat([F.pad(x_t, (0, pad)), F.pad(x_cond, (0, pad))], 1)
    N = x.shape[-1] // P
    Cin = x.shape[1]   # input_channels + cond_channels
    x = x.reshape(B, Cin, N, P).permute(0, 2, 1, 3).reshape(B, N, Cin * P)
    x = F.gelu(x @ p["patch_embed.proj.0.weight"].T + p["patch_embed.proj.0.bias"])
    x = x @ p["patch_embed.proj.2.weight"].T + p["patch_embed.proj.2.bias"]
    half = D // 2
    freq = torch.exp(torch.arange(half, device=t.device, dtype=torch.float32) * -(math.log(10000.0) / (half - 1)))
    e = t[:, None] * freq[None]
    e = torch.cat([e.sin(), e.cos()], -1).to(x.dtype)
    e = F.silu(e @ p["t_embedder.1.weight"].T + p["t_embedder.1.bias"])
    t_emb = e @ p["t_embedder.3.weight"].T + p["t_embedder.3.bias"]
    fn = block_fn or block_forward
    has = lambda name: masks is not None and masks.get(name) is not None
    for i in range(depth):
        k = f"blocks.{i}."
        cos, sin = p[k + "attn.rope.cos_cached"][:N].to(x.dtype), p[k + "attn.rope.sin_cached"][:N].to(x.dtype)
        x = fn(x, t_emb, block_weights(p, i), cos, sin, Hq, Hkv, rms,
               masks["attn"][i] if has("attn") else None, masks["hid"][i] if has("hid") else None,
               masks["out"][i] if has("out") else None, masks["path"][i] if has("path") else None)
        if blocks_out is not None:
            blocks_out.append(x.detach().reshape(B * N, D))
    x = _norm(x, p.get("final_layer.0.weight"), rms)
    x = x @ p["final_layer.1.weight"].T + p["final_layer.1.bias"]
    x = x.view(B, N, C, P).permute(0, 2, 1, 3).reshape(B, C, N * P)
    return x[:, :, :T]


@torch.no_grad()
def flow_matching_sample(p, cfg, lr_latent, z0, num_steps=50, cfg_scale=3.0, rms=False):
    """The reference sampler (infer_test_v3m2.py:108-185) on the restatement above, z0 injected."""
    B = lr_latent.shape[0]
    z = z0.clone()
    ts = torch.linspace(0.0, 1.0, num_steps + 1, device=z.device)
    for i in range(num_steps):
        t_curr, t_next = ts[i], ts[i + 1]
        dt = t_next - t_curr
        tb = torch.full((B,), float(t_curr), device=z.device)
        if cfg_scale != 1.0:
            out = dit_forward(p, cfg, torch.cat([z, z]), torch.cat([tb, tb]), torch.cat([lr_latent, torch.zeros_like(lr_latent)]),
                              rms=rms)
            x_c, x_u = out[:B], out[B:]
            x = x_u + cfg_scale * (x_c - x_u)
        else:
            x = dit_forward(p, cfg, z, tb, lr_latent, rms=rms)
        if float(t_curr) < 0.999:
            z = z + (x - z) / (1 - t_curr + 1e-5) * dt
        else:
            z = x
    return z
