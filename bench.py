#!/usr/bin/env python
"""Headline benchmark: v3mod2 DiT denoise steps/sec (BASELINE.json metric).

One "step" = one Euler step of the flow-matching sampler with CFG on a batch of B=28 latent chunks
[28, 1024, 1378]: one DiT forward over B_eff = 56 (M = 19320 token rows, 766 M params, depth 28) plus
the fused CFG/velocity/Euler update.  Weights random-init (zero-init layers re-randomised), synthetic
unit-variance latents.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--sub auto|none|train,long,v2,gpu_baseline]

N > 1 is launched by torchrun (one rank per GPU); the sampling path shards by batch (every rank denoises its
own B=28 batch, no data-path collective) -> weak scaling, value = N*K / max-over-ranks time.

JSON keys beyond the base contract:
  e2e          same metric through the public API `flow_matching_sample` with HOST (pinned) buffers:
               H2D of the condition latent + K steps + D2H of the result inside the timed region
  roofline     dominant kernel class of the step, timed live with CUDA events (jat_profile_*)
  kernels      the same for every kernel class (share of the step, achieved TFLOP/s or GB/s)
  cpu_baseline the torch (ATen) fp32 restatement of the reference timed on the host cores on a bounded sample
  train        (every N) BASELINE configs[3]: the DDP training step, batch 28 per GPU -- ms/step, fraction of the tensor
               roofline, per-class kernel ms, the all-reduce cost that is not hidden, a cross-rank parameter checksum
  long         (every N) BASELINE configs[4]: the 10-minute track, its 43 chunks dealt over the N GPUs (strong scaling)
  v2           (N = 1) BASELINE configs[1] (v2 288 M, B = 1, 25 steps)
  gpu_baseline (N = 1) stock PyTorch on the same B200: the torch restatement of the reference (oracle/torch_dit.py) for the
               same denoise step under eager fp32, eager bf16 autocast and torch.compile + bf16 autocast, and for the training
               step under eager bf16 autocast; `speedup_vs_*` = ours / torch on the same GPU
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(input_channels=1024, cond_channels=1024, patch_len=4, hidden_size=1280, depth=28, num_q_heads=20,
           num_kv_heads=4, bottleneck_dim=512, mlp_ratio=4.0, dropout=0.1, drop_path_rate=0.05)  # train_ddp_v3mod2.py:343-355
CFG_V2 = dict(input_channels=1024, cond_channels=1024, patch_len=4, hidden_size=1024, depth=16, num_q_heads=16,
              num_kv_heads=4, bottleneck_dim=512, mlp_ratio=4.0, dropout=0.1, drop_path_rate=0.0)   # train_ddp_v2.py:64-76
B, C, T, CFG_SCALE = 28, 1024, 1378, 3.0
METRIC = "v3mod2 DiT denoise steps/sec (B=28, CFG=3.0, 766M params)"


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return dict(hbm=p["hbm_gbs"], tf=p["bf16_tflops_sustained"], tf_burst=p["bf16_tflops"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, src="fallback")


def model_flops_per_token(cfg, N):
    """Algorithmic forward FLOPs per token: projections + attention + patch embed + final layer."""
    D, depth, BD = cfg["hidden_size"], cfg["depth"], cfg["bottleneck_dim"]
    F_ = int(D * cfg["mlp_ratio"])
    qkv = (cfg["num_q_heads"] + 2 * cfg["num_kv_heads"]) * 64
    per_block = 2.0 * D * qkv + 2.0 * D * D + 4.0 * D * F_ + 4.0 * N * D
    KIN = 2 * cfg["input_channels"] * cfg["patch_len"]
    return depth * per_block + 2.0 * KIN * BD + 2.0 * BD * D + 2.0 * D * cfg["input_channels"] * cfg["patch_len"]


def step_work(Beff=2 * B):
    """Algorithmic FLOPs / bytes of one CFG step per kernel class (padding not counted)."""
    D, F, H, Hkv, depth, BD = 1280, 5120, 20, 4, 28, 512
    N = (T + 3) // 4
    M = Beff * N
    qkv = (H + 2 * Hkv) * 64
    w = {
        "gemm_bias_act": ("flops", 2.0 * M * (8192 * BD + BD * D) + depth * 2.0 * M * D * F),
        "gemm_qkv_rope": ("flops", depth * 2.0 * M * D * qkv),
        "gemm_gate_residual": ("flops", depth * 2.0 * M * (D * D + F * D)),
        "gemm_unpatchify": ("flops", 2.0 * M * D * 4096),
        "gqa_attention_fwd": ("flops", depth * 4.0 * Beff * H * N * N * 64),
        "adaln_norm_modulate": ("bytes", (2 * depth + 1) * M * D * 6.0),
        "patchify_cast": ("bytes", 2.0 * B * C * T * 4 + M * 8192 * 2.0),
        "cfg_euler_update": ("bytes", 4.0 * B * C * T * 4),
    }
    total_flops = sum(v for k, (u, v) in w.items() if u == "flops")
    return w, total_flops


def ncu_traffic(kernel_class):
    """DRAM bytes per launch of a kernel class from the committed ncu `--set full` capture (None if absent)."""
    import re
    epi = {"gemm_bias_act": 0, "gemm_qkv_rope": 1, "gemm_gate_residual": 2, "gemm_unpatchify": 3}
    d = None
    for tag in ("r2", "r1c", "r1b"):   # latest committed capture first
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.json")))
            ncu_traffic.source = f"profiles/{tag}_ncu_full_summary.json"
            break
        except Exception:
            continue
    if d is None:
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = []
    for items in d.values():
        for it in items:
            name = it.get("Kernel Name", "")
            m = re.search(r"gemm_tcgen05_kernel<(\d+), (\d+), (\d+)", name)
            hit = (m and epi.get(kernel_class) == int(m.group(3))) or (not m and kernel_class.replace("_fwd", "") in name)
            if not hit:
                continue
            try:
                tot = 0.0
                for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    v, u = it[k].split()
                    tot += float(v) * unit[u]
                vals.append(tot)
            except Exception:
                pass
    return round(sum(vals) / len(vals)) if vals else None


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            out = ""
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "samples": len(sm),
                "reasons": sorted(reasons)}


class Dist:
    """torch.distributed plumbing for the bench: one NCCL group per process (high-priority stream: its kernels get SMs at
    the next kernel boundary instead of queueing behind the persistent GEMM grids), barrier + max-over-ranks helpers."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        import torch
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            # DDP training (configs[3]): NCCL's all-reduce kernels and the library's persistent GEMM grids must not fight for
            # SMs -- a GEMM CTA that waits for an SM held by a long-running NCCL CTA delays its whole statically scheduled
            # grid (weight-gradient GEMMs +15 %, GELU' dgrad +35 % measured).  NCCL is capped at 8 CTAs and the GEMM grids
            # leave 8 SMs free (train_record: jat_set_gemm_sm_reserve); measured at 2 / 8 GPUs in DESIGN.md section 6.
            os.environ.setdefault("NCCL_MAX_CTAS", os.environ.get("JAT_NCCL_MAX_CTAS", "8"))
            if os.environ.get("JAT_NCCL_HIPRI", "1") == "1":
                dist.init_process_group("nccl", init_method="env://",
                                        pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
            else:
                dist.init_process_group("nccl", init_method="env://")
            self.dist = dist

    def barrier(self):
        import torch
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize(self.dev)

    def max(self, x):
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def build_model(dev, norm, cfg=None, seed=0):
    import torch
    import jat_b200
    cls = jat_b200.JaT_AudioSR_V2 if norm == "layernorm" else jat_b200.JaT_AudioSR_V3
    torch.manual_seed(seed)
    with torch.device(dev):
        model = cls(**(cfg or CFG))
    g = torch.Generator(device=dev).manual_seed(1)
    with torch.no_grad():
        for name, p in model.named_parameters():  # adaLN-Zero / final layers are zero-initialised
            if "adaLN_modulation.1" in name or name.startswith("final_layer.1"):
                p.copy_(torch.randn(p.shape, generator=g, device=dev) * 0.02)
    return model.eval()


def run_steps(model, plan, first, count):
    from jat_b200 import ops
    eng = model._engine
    for i in range(first, first + count):
        eng.forward_tokens(plan.ws, plan.z, plan.lr, plan.Beff, plan.mod[i], 0, plan.x_pred, cond_batch=B)
        ops.cfg_euler_update(plan.z, plan.x_pred[:B], plan.x_pred[B:], CFG_SCALE, plan.t_dt, i)


# ------------------------------------------------------------------------------------------------ CPU baseline (ATen on the host cores)
def host_threads():
    """All host cores, set explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers, which halved this arm in round 1."""
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0)) or n
    except Exception:
        pass
    torch.set_num_threads(n)
    return n


def cpu_weights(norm, seed=0):
    """Random fp32 weights of the v3mod2 architecture on the host (state_dict layout of the reference module)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    D, depth, F = 1280, 28, 5120
    w = {}

    def lin(name, o, i, bias=True, std=None):
        s = std if std is not None else 1.0 / (i ** 0.5)
        w[name + ".weight"] = torch.randn(o, i, generator=g) * s
        if bias:
            w[name + ".bias"] = torch.randn(o, generator=g) * 0.02
    lin("patch_embed.proj.0", 512, 8192)
    lin("patch_embed.proj.2", D, 512)
    lin("t_embedder.1", D, D)
    lin("t_embedder.3", D, D)
    inv = 1.0 / (10000 ** (torch.arange(0, 64, 2).float() / 64))
    fr = torch.outer(torch.arange(4096).float(), inv)
    emb = torch.cat([fr, fr], -1)
    cos, sin = emb.cos(), emb.sin()
    for i in range(depth):
        p = f"blocks.{i}."
        lin(p + "attn.q_proj", D, D, False)
        lin(p + "attn.k_proj", 256, D, False)
        lin(p + "attn.v_proj", 256, D, False)
        lin(p + "attn.out_proj", D, D, False)
        lin(p + "mlp.0", F, D)
        lin(p + "mlp.3", D, F)
        lin(p + "adaLN_modulation.1", 6 * D, D, True, 0.02)
        w[p + "attn.rope.cos_cached"], w[p + "attn.rope.sin_cached"] = cos, sin
        if norm == "rmsnorm":
            w[p + "norm1.weight"] = torch.ones(D)
            w[p + "norm2.weight"] = torch.ones(D)
    lin("final_layer.1", 4096, D, True, 0.02)
    if norm == "rmsnorm":
        w["final_layer.0.weight"] = torch.ones(D)
    return w


CPU_SAMPLE_B = 1


def cpu_baseline_sample(weights, norm):
    """One CFG denoise step of the reference's algorithm in fp32 through ATen at batch CPU_SAMPLE_B (B_eff = 2, 690 token
    rows, full depth 28): the reference's per-step work (infer_test_v3m2.py:145-179) on oracle/torch_dit.py."""
    import torch
    from oracle.torch_dit import dit_forward   # baseline leg only -- never on the product path
    g = torch.Generator().manual_seed(0)
    b = CPU_SAMPLE_B
    z = torch.randn(b, C, T, generator=g)
    lr = torch.randn(b, C, T, generator=g)
    t0 = time.perf_counter()
    with torch.no_grad():
        tb = torch.full((2 * b,), 0.37)
        both = dit_forward(weights, CFG, torch.cat([z, z]), tb, torch.cat([lr, torch.zeros_like(lr)]), rms=(norm == "rmsnorm"))
        x = both[b:] + CFG_SCALE * (both[:b] - both[b:])
        z = z + (x - z) / (1 - 0.37 + 1e-5) * 0.02
    return time.perf_counter() - t0


def reference_arm(a, K, W, rank, world, config):
    """The reference's CPU implementation of the path: the reference is pure PyTorch (not present on the GPU box), so its
    CPU path IS ATen -- timed here through the torch restatement oracle/torch_dit.py (pinned to the unmodified reference by
    tests/test_oracle.py), fp32, all host threads, each step a bounded sample (batch 1 of the 28)."""
    if rank != 0:
        return
    threads = host_threads()
    w = cpu_weights(a.norm)
    budget = 200.0
    t_first = cpu_baseline_sample(w, a.norm)  # warm-up 1 (also sizes the run)
    k_eff = K
    if (K + W) * t_first > budget:
        k_eff = max(1, int(budget / t_first) - W)
    for _ in range(max(W - 1, 0)):
        cpu_baseline_sample(w, a.norm)
    times = [cpu_baseline_sample(w, a.norm) for _ in range(k_eff)]
    ts = sum(times) / len(times)
    value = 1.0 / (ts * B / CPU_SAMPLE_B)
    sample = (f"torch (ATen) fp32 restatement of the reference forward + CFG/Euler update on {threads} host threads "
              f"(torch.set_num_threads, independent of OMP_NUM_THREADS), each step = 1 CFG denoise step at batch {CPU_SAMPLE_B} "
              f"(B_eff={2 * CPU_SAMPLE_B}, {2 * CPU_SAMPLE_B * 345} token rows, depth 28), mean {ts:.2f} s over {k_eff} steps, "
              f"scaled x{B // CPU_SAMPLE_B} to the batch-28 step")
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 6), "unit": "steps/s", "n_gpus": world,
            "steps": k_eff, "warmup": W, "ms_per_step": round(ts * B / CPU_SAMPLE_B * 1e3, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": round(value, 6), "unit": "steps/s", "cores": threads, "kind": "port",
                             "sample": sample},
            "e2e": {"value": round(value, 6), "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ configs[3]: training step
def train_record(a, K, W, D_, model=None):
    """BASELINE configs[3]: v3mod2 training step, batch 28 per GPU, x-prediction MSE flow-matching loss
    (train_ddp_v3mod2.py:842-930 without the perceptual losses / logging), DDP over NCCL for N > 1.
    One step = zero_grad + input prep + forward + loss + backward (+ gradient all-reduce) + clip_grad_norm_ + AdamW + weight re-pack."""
    import torch
    import jat_b200
    from jat_b200 import _lib as L
    from jat_b200 import training
    dev, world, rank, local = D_.dev, D_.world, D_.rank, D_.local
    if model is None:
        model = build_model(dev, a.norm)
    model.train()
    model.grad_handoff = os.environ.get("JAT_GRAD_HANDOFF", "view")   # zero-copy .grad (the step calls zero_grad(set_to_none=True))
    net = model
    # gradient exchange: f32 (the reference's) up to 2 GPUs, bf16 payload (jat_b200.ddp, two fused passes) from 4 GPUs on, where
    # the ring moves 1.75-1.9 x the gradient bytes per rank and 8 NCCL CTAs cannot hide 3.06 GB of f32 behind the backward
    grad_wire = os.environ.get("JAT_DDP_GRAD_DTYPE", "bf16" if world >= 4 else "f32")
    sm_reserve = int(os.environ.get("JAT_SM_RESERVE", "8")) if world > 1 else 0
    if world > 1:
        from jat_b200 import ops as _ops
        _ops.set_gemm_sm_reserve(dev, sm_reserve)
        net = torch.nn.parallel.DistributedDataParallel(
            model, device_ids=[local], find_unused_parameters=False,
            # the reference's call (train_ddp_v3mod2.py:822) + gradients as views of the all-reduce buckets (no copy back);
            # torch's default 25 MB buckets: the all-reduce of the LAST bucket cannot overlap anything, so it should be small
            bucket_cap_mb=int(os.environ.get("JAT_DDP_BUCKET_MB", "25")),
            gradient_as_bucket_view=os.environ.get("JAT_DDP_BUCKET_VIEW", "1") == "1")
        # (registered below, once the optimizer exists: FusedAdamW consumes the bf16 payload directly)
    fused_opt = os.environ.get("JAT_BENCH_TORCH_OPT", "0") == "0"
    if fused_opt:   # clip_grad_norm_(1.0) + AdamW + bf16 re-pack in two multi-tensor passes (jat_b200.FusedAdamW)
        opt = jat_b200.FusedAdamW(model.parameters(), lr=5e-5, weight_decay=0.1, max_grad_norm=1.0, model=model)
    else:           # the reference's own calls, train_ddp_v3mod2.py:709, 926-928
        opt = torch.optim.AdamW(model.parameters(), lr=5e-5, weight_decay=0.1, fused=True)
    if world > 1 and grad_wire == "bf16":   # bf16 payload (jat_b200.ddp): one fused compress pass, consumed by FusedAdamW as is
        # (the fused consumer measured the same as the expanding hook at 8 GPUs -- 52.6 vs 52.4 ms, the expansion pass overlaps the
        #  backward -- so the default keeps `.grad` = the averaged gradient; JAT_DDP_FUSED_CONSUMER=1 selects it)
        fused_consumer = fused_opt and os.environ.get("JAT_DDP_FUSED_CONSUMER", "0") == "1" and \
            os.environ.get("JAT_DDP_BUCKET_VIEW", "1") == "1"
        jat_b200.ddp.register_bf16_allreduce(net, optimizer=opt if fused_consumer else None)
    gd = torch.Generator(device=dev).manual_seed(100 + rank)
    hr = torch.randn(B, C, T, generator=gd, device=dev) * 2.0 + 0.3                   # raw (un-normalised) DAC latents
    lr = torch.randn(B, C, T, generator=gd, device=dev) * 2.0 + 0.3
    hr_mean = torch.full((1, C, 1), 0.3, device=dev)
    hr_std = torch.full((1, C, 1), 2.0, device=dev)

    def step():
        u = torch.rand(B, generator=gd, device=dev)                                   # U-shaped t, :449-457
        t = torch.where(u < 0.5, (2 * u).sqrt() / 2, 1 - (2 * (1 - u)).sqrt() / 2)
        noise = torch.randn(B, C, T, generator=gd, device=dev)
        cond_noise = torch.randn(B, C, T, generator=gd, device=dev)
        # normalise + 5 % conditional noise + z_t = t x + (1 - t) eps in one kernel (:856-883)
        hr_norm, lr_cond, z_t = training.prepare_inputs(hr, lr, hr_mean, hr_std, hr_mean, hr_std, t, noise,
                                                        cond_noise=cond_noise, cond_scale=0.05)
        opt.zero_grad(set_to_none=True)
        loss = training.mse_loss(net(z_t, t, lr_cond), hr_norm)                       # :886-889, fused with its gradient seed
        loss.backward()                                                               # :922 (+ DDP all-reduce)
        if not fused_opt:
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)                   # :926
        opt.step()                                                                    # :928
        return loss

    def timed(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D_.barrier()
        e0.record()
        h0 = time.perf_counter()
        for _ in range(n):
            out = step()
        timed.host_ms = (time.perf_counter() - h0) * 1e3 / n   # time the host needs to ISSUE a step (no sync inside)
        e1.record()
        D_.barrier()
        return D_.max(e0.elapsed_time(e1)), out

    for _ in range(W):
        step()
    D_.barrier()
    clocks = ClockSampler(local)
    ctx, lib = L.context(local), L.load()
    l0 = lib.jat_launch_count(ctx)
    ms, loss = timed(K)
    host_issue_ms = timed.host_ms
    clk = clocks.stop()
    launches = lib.jat_launch_count(ctx) - l0
    assert torch.isfinite(loss).all()
    # every rank must hold the same parameters after the K + W optimizer steps (DDP correctness; tests/test_ddp_gpu.py
    # checks it bit by bit on a small model): compare an order-independent f64 checksum across ranks
    chk = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
    chk_abs = torch.stack([p.detach().double().abs().sum() for p in model.parameters()]).sum().reshape(1)
    checksum_ok = True
    if world > 1:
        both = torch.cat([chk, chk_abs])
        lo, hi = both.clone(), both.clone()
        D_.dist.all_reduce(lo, op=D_.dist.ReduceOp.MIN)
        D_.dist.all_reduce(hi, op=D_.dist.ReduceOp.MAX)
        checksum_ok = bool(torch.equal(lo, hi))
        if not checksum_ok:
            raise SystemExit(f"bench.py: parameters diverged across ranks after DDP training: {lo.tolist()} vs {hi.tolist()}")
    # per-kernel-class timing (CUDA event pair around every launch of the library)
    L.profile_begin(local)
    pk = min(K, 3)
    for _ in range(pk):
        step()
    prof = L.profile_end(local)
    # the same step WITHOUT the gradient exchange (DDP.no_sync): what the all-reduce costs the step, hidden part excluded
    nosync_ms = None
    if world > 1:
        with net.no_sync():
            step()
            kn = min(K, 5)
            nosync_total, _ = timed(kn)
        nosync_ms = nosync_total / kn
    pkz = peaks()
    tot = sum(v[0] for v in prof.values())
    kernels = {n: {"ms_per_step": round(v[0] / pk, 3), "launches_per_step": v[1] // pk, "share_of_kernel_time": round(v[0] / tot, 4)}
               for n, v in prof.items()}
    fw, _ = step_work(B)
    fwd_gemm = sum(fw[k][1] for k in ("gemm_bias_act", "gemm_qkv_rope", "gemm_gate_residual", "gemm_unpatchify"))
    n_params = sum(p.numel() for p in model.parameters())
    N = (T + 3) // 4
    roofs = {}
    if "gemm_accum" in kernels:     # the weight-gradient GEMMs do exactly the forward GEMMs' FLOPs
        tfs = fwd_gemm / (kernels["gemm_accum"]["ms_per_step"] * 1e-3) / 1e12
        roofs["gemm_accum"] = {"bound": "tensor", "achieved": round(tfs, 1), "peak": pkz["tf"], "unit": "TFLOP/s",
                               "frac": round(tfs / pkz["tf"], 4)}
    if "attention_bwd" in kernels:  # 2.5 x the forward attention FLOPs (5 MMAs of the forward's 2 shapes)
        tfs = 2.5 * fw["gqa_attention_fwd"][1] / (kernels["attention_bwd"]["ms_per_step"] * 1e-3) / 1e12
        roofs["attention_bwd"] = {"bound": "tensor", "achieved": round(tfs, 1), "peak": pkz["tf"], "unit": "TFLOP/s",
                                  "frac": round(tfs / pkz["tf"], 4)}
    if "optimizer" in kernels:      # 30 B per parameter for the update + 4 B for the norm pass
        gbs = n_params * 34.0 / (kernels["optimizer"]["ms_per_step"] * 1e-3) / 1e9
        roofs["optimizer"] = {"bound": "hbm", "achieved": round(gbs, 1), "peak": pkz["hbm"], "unit": "GB/s",
                              "frac": round(gbs / pkz["hbm"], 4)}
    step_flops = 3 * model_flops_per_token(CFG, N) * B * N   # forward + backward = 3 x forward, no recompute counted
    rec = {"metric": "v3mod2 DDP training steps/sec (batch 28 per GPU, x-prediction MSE flow-matching loss)",
           "value": round(world * K / (ms / 1e3), 3), "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
           "ms_per_step": round(ms / K, 3), "scaling": "weak",
           "config": {"workload": "configs[3]: v3mod2 DiT 1280/28/20Q/4KV training step, batch 28 x [1024,1378] per GPU "
                                  "(9660 token rows): normalise + cond-noise + flow-matching mix, forward (Dropout 0.1, DropPath 0.05), "
                                  "MSE x-prediction loss, backward (grad_handoff=" + model.grad_handoff + "), clip_grad_norm_(1.0) + AdamW + bf16 weight re-pack "
                                  + ("(jat_b200.FusedAdamW: 2 multi-tensor passes)" if fused_opt else "(torch: foreach clip, fused AdamW, re-cast)"),
                      "norm": a.norm, "dropout": CFG["dropout"], "drop_path": CFG["drop_path_rate"], "cond_noise_ratio": 0.05,
                      "parallelism": (f"DDP x{world} (torch DistributedDataParallel, {grad_wire} NCCL gradient all-reduce on a high-priority stream, "
                                      f"{os.environ.get('JAT_DDP_BUCKET_MB', '25')} MB buckets, gradient_as_bucket_view, NCCL_MAX_CTAS={os.environ.get('NCCL_MAX_CTAS')}, "
                                      f"{sm_reserve} SMs kept out of the GEMM grids)") if world > 1 else "single GPU"},
           "clocks": clk, "gpu_launches": int(launches), "loss": round(float(loss.item()), 5),
           "host_issue_ms_per_step": round(host_issue_ms, 3),
           "step_tflops_per_gpu": round(step_flops / (ms / K) / 1e9, 1),
           "step_tensor_frac_sustained": round(step_flops / (ms / K) / 1e9 / pkz["tf"], 4),
           "rooflines": roofs, "kernels": kernels, "kernel_ms_per_step": round(tot / pk, 2),
           "params_identical_across_ranks": checksum_ok if world > 1 else None,
           "param_checksum": float(chk.item())}
    if world > 1:
        _ops.set_gemm_sm_reserve(dev, 0)
    if nosync_ms is not None:
        rec["no_sync_ms_per_step"] = round(nosync_ms, 3)
        rec["allreduce_exposed_ms"] = round(ms / K - nosync_ms, 3)
        rec["allreduce_note"] = ("no_sync = the same DDP-wrapped step with the gradient exchange switched off (DDP.no_sync), timed on "
                                 "the same ranks right after; allreduce_exposed_ms = ms_per_step - no_sync: all-reduce time that is "
                                 "not hidden behind the backward plus the slow-down of kernels sharing SMs / HBM with NCCL")
    return rec, model


def train_main(a, K, W):
    D_ = Dist()
    rec, _ = train_record(a, K, W, D_)
    if D_.rank == 0:
        rec.update(higher_is_better=True, vs_baseline=None, dtype="bf16", data="synthetic")
        print(json.dumps(rec), flush=True)
    D_.close()


# ------------------------------------------------------------------------------------------------ configs[4]: long audio
def long_record(a, K, W, D_, model=None):
    """BASELINE configs[4]: a 10-minute track (latent [1024, 51679] at 86.13 frames/s) cut into 43 chunks of 1378 frames
    (overlap 172, infer_test_v3m2.py:340-348), the chunks dealt round-robin over the GPUs, each GPU denoising its chunks as
    ONE batch with 50 CFG = 3.0 steps, all-gather of the finished chunk latents, crossfade + de-normalise
    (`jat_b200.chunked.sample_long`).  One "step" = the whole track; value = audio seconds per wall second."""
    import torch
    from jat_b200 import chunked
    dev, world, rank, local = D_.dev, D_.world, D_.rank, D_.local
    if model is None:
        model = build_model(dev, a.norm)
    model.eval()
    seconds, frames = 600.0, 51679
    g = torch.Generator().manual_seed(5)
    track = (torch.randn(C, frames, generator=g) * 2.0 + 0.3).pin_memory()
    mean, std = torch.full((C,), 0.3), torch.full((C,), 2.0)

    def run():
        out = chunked.sample_long(model, track.to(dev, non_blocking=True), mean, std, mean, std, num_steps=50, cfg_scale=CFG_SCALE,
                                  device=dev)
        return out.cpu() if rank == 0 else out

    for _ in range(max(W, 1)):
        run()
    D_.barrier()
    clocks = ClockSampler(local)
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        out = run()
    e1.record()
    D_.barrier()
    wall = time.perf_counter() - t0
    clk = clocks.stop()
    ms = D_.max(e0.elapsed_time(e1))
    if rank == 0:
        assert out.shape == (1, C, frames) and torch.isfinite(out).all()
    n_chunks = len(chunked.plan_chunks(frames))
    N = (T + 3) // 4
    flops = 50 * 2 * n_chunks * N * model_flops_per_token(CFG, N)
    return {"metric": "long-audio chunked inference: audio seconds per second (10-minute track, 50-step CFG=3.0)",
            "value": round(K * seconds / (ms / 1e3), 1), "unit": "audio s/s", "n_gpus": world, "steps": K, "warmup": max(W, 1),
            "ms_per_step": round(ms / K, 1), "scaling": "strong",
            "tensor_frac_sustained": round(flops / (ms / K) / 1e9 / peaks()["tf"] / world, 4),
            "config": {"workload": f"configs[4]: 10-minute track = latent [1024, {frames}] -> {n_chunks} chunks of 1378 frames "
                                   f"(overlap 172), round-robin over {world} GPU(s), 50 Euler steps CFG=3.0 per chunk batch, "
                                   "all-gather + crossfade + de-normalise; host track in, host latent out",
                       "norm": a.norm, "chunks": n_chunks, "chunks_per_gpu": -(-n_chunks // world)},
            "clocks": clk, "wall_s": round(wall, 3)}


def long_main(a, K, W):
    D_ = Dist()
    rec = long_record(a, K, W, D_)
    if D_.rank == 0:
        rec.update(higher_is_better=True, vs_baseline=None, dtype="bf16", data="synthetic")
        print(json.dumps(rec), flush=True)
    D_.close()


# ------------------------------------------------------------------------------------------------ configs[1]: v2, B = 1
def v2_record(a, D_):
    """BASELINE configs[1]: v2 DiT 1024 / 16 blocks / 16Q / 4KV (288 M), 25-step CFG = 3.0 sampling at the
    reference's own inference batch B = 1 (B_eff = 2, 690 token rows: the small-M regime), through the public sampler
    (device-resident condition latent; the 25-step loop is one CUDA graph replay) and, for comparison, launch by launch."""
    import torch
    import jat_b200
    dev = D_.dev
    model = build_model(dev, a.norm, CFG_V2, seed=2)
    g = torch.Generator(device=dev).manual_seed(3)
    lr = torch.randn(1, C, T, generator=g, device=dev)
    steps = 25
    out = {}
    for tag, use_graph in (("graph", True), ("launches", False)):
        for _ in range(2):
            jat_b200.flow_matching_sample(model, lr, num_steps=steps, cfg_scale=CFG_SCALE, device=dev, verbose=False,
                                          use_graph=use_graph)
        torch.cuda.synchronize(dev)
        reps = 4
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            z = jat_b200.flow_matching_sample(model, lr, num_steps=steps, cfg_scale=CFG_SCALE, device=dev, verbose=False,
                                              use_graph=use_graph)
        e1.record()
        torch.cuda.synchronize(dev)
        out[tag] = e0.elapsed_time(e1) / (reps * steps)
    assert torch.isfinite(z).all()
    from jat_b200 import _lib as L
    L.profile_begin(D_.local)   # per-class kernel time of one launch-by-launch run (CUDA event pair around every launch)
    jat_b200.flow_matching_sample(model, lr, num_steps=steps, cfg_scale=CFG_SCALE, device=dev, verbose=False, use_graph=False)
    prof = L.profile_end(D_.local)
    kernels = {n: {"us_per_launch": round(v[0] / v[1] * 1e3, 2), "launches_per_step": round(v[1] / steps, 1),
                   "ms_per_step": round(v[0] / steps, 4)} for n, v in prof.items()}
    N = (T + 3) // 4
    flops = 2 * N * model_flops_per_token(CFG_V2, N)
    best = min(out.values())
    pk = peaks()
    return {"metric": "v2 DiT denoise steps/sec (B=1, CFG=3.0, 288M params, 25-step sampler)", "value": round(1e3 / best, 1),
            "unit": "steps/s", "ms_per_step": round(best, 4), "ms_per_step_graph": round(out["graph"], 4),
            "ms_per_step_launch_by_launch": round(out["launches"], 4),
            "step_tflops": round(flops / best / 1e9, 1), "step_tensor_frac_sustained": round(flops / best / 1e9 / pk["tf"], 4),
            "ideal_ms_per_step_at_sustained_peak": round(flops / pk["tf"] / 1e9, 4), "kernels": kernels,
            "kernel_sum_ms_per_step": round(sum(k["ms_per_step"] for k in kernels.values()), 4),
            "config": {"workload": "configs[1]: v2 DiT 1024/16/16Q/4KV, CFG=3.0, 25 Euler steps, batch 1 x [1024,1378] "
                                   "(B_eff 2, 690 token rows)", "norm": a.norm}}


# ------------------------------------------------------------------------------------------------ stock PyTorch on the same GPU
def gpu_baseline_main(a, K, W):
    """Stock PyTorch on this GPU (SURVEY.md 0.1 / 8d i-iii: "the kernel to beat"): the torch restatement of the reference
    (oracle/torch_dit.py, pinned to the unmodified reference modules by tests/test_oracle.py -- the reference itself is not
    on the GPU box), same weights shapes / inputs / step definition as the headline metric, CUDA events, W warm-up + K timed.
    Runs in its own process (spawned by the default bench run) so that a slow or failing compile cannot take the bench down."""
    import torch
    import torch.nn.functional as F
    from oracle.torch_dit import block_forward, dit_forward   # baseline leg only -- never on the product path
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = False   # torch's default: true fp32 matmuls (what the reference's fp32 inference runs)
    model = build_model(dev, a.norm)
    prm = {k: v.detach() for k, v in model.state_dict().items()}
    rms = a.norm == "rmsnorm"
    g = torch.Generator(device=dev).manual_seed(123)
    z0 = torch.randn(B, C, T, generator=g, device=dev)
    lr = torch.randn(B, C, T, generator=g, device=dev)
    ts = torch.linspace(0.0, 1.0, 51, device=dev)
    res = {"torch": torch.__version__, "steps": K, "warmup": W,
           "what": "oracle/torch_dit.py (torch restatement of the reference forward, materialised softmax attention as in "
                   "jat_audiosr_v2.py:147-164) + the reference's per-step sampler math (infer_test_v3m2.py:145-179, host syncs "
                   "removed), batch 28 (B_eff 56), per-sample t-embedding / adaLN recomputed every step as the reference does"}

    def sample_step(z, i, fwd):
        tb = ts[i].expand(2 * B)
        out = fwd(prm, CFG, torch.cat([z, z]), tb, torch.cat([lr, torch.zeros_like(lr)]), rms=rms)
        x = out[B:] + CFG_SCALE * (out[:B] - out[B:])
        return z + (x - z) / (1 - ts[i] + 1e-5) * (ts[i + 1] - ts[i])

    def time_sampling(fwd, autocast):
        z = z0.clone()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            for i in range(W):
                z = sample_step(z, i, fwd)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(W, W + K):
                z = sample_step(z, i, fwd)
            e1.record()
            torch.cuda.synchronize(dev)
        assert torch.isfinite(z).all()
        ms = e0.elapsed_time(e1) / K
        return {"ms_per_step": round(ms, 3), "steps_per_s": round(1e3 / ms, 3)}

    def attempt(name, fn):
        try:
            res[name] = fn()
        except Exception as e:   # an OOM / compile failure of the BASELINE is reported, not fatal
            res[name] = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
            torch.cuda.empty_cache()
        print("GPU_BASELINE_PARTIAL " + json.dumps(res), flush=True)

    attempt("sample_eager_bf16_autocast", lambda: time_sampling(dit_forward, True))
    attempt("sample_eager_fp32", lambda: time_sampling(dit_forward, False))

    def compiled():
        t0 = time.perf_counter()
        blk = torch.compile(block_forward, mode="default", backend="inductor", dynamic=False)   # train_ddp_v3mod2.py:816's call,
        fwd = lambda *x, **k: dit_forward(*x, block_fn=blk, **k)                               # applied per block (one graph, 28 uses)
        r = time_sampling(fwd, True)
        r["compile_plus_run_s"] = round(time.perf_counter() - t0, 1)
        r["note"] = "torch.compile(mode='default', backend='inductor') of the DiT block function (regional compilation: one " \
                    "compile serves the 28 blocks), embeddings / final layer / sampler math eager"
        return r
    attempt("sample_compile_bf16_autocast", compiled)

    def train_eager():
        p = {k: (v.clone().requires_grad_(True) if (v.dtype.is_floating_point and "rope" not in k) else v) for k, v in prm.items()}
        params = [v for v in p.values() if v.requires_grad]
        opt = torch.optim.AdamW(params, lr=5e-5, weight_decay=0.1, fused=True)
        hr = torch.randn(B, C, T, generator=g, device=dev)
        cfg0 = dict(CFG)

        def step():
            t = torch.rand(B, generator=g, device=dev)
            noise = torch.randn(B, C, T, generator=g, device=dev)
            z_t = t.view(-1, 1, 1) * hr + (1 - t.view(-1, 1, 1)) * noise
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = F.mse_loss(dit_forward(p, cfg0, z_t, t, lr, rms=rms), hr)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            return loss
        kk = max(2, min(K, 5))
        for _ in range(2):
            step()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(kk):
            loss = step()
        e1.record()
        torch.cuda.synchronize(dev)
        assert torch.isfinite(loss)
        ms = e0.elapsed_time(e1) / kk
        return {"ms_per_step": round(ms, 3), "steps_per_s": round(1e3 / ms, 3), "steps": kk,
                "peak_mem_gb": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 1),
                "note": "forward + backward + clip_grad_norm_ + torch fused AdamW, batch 28, bf16 autocast, Dropout / DropPath OFF "
                        "(the restatement injects masks instead of drawing them), f32 master weights"}
    attempt("train_eager_bf16_autocast", train_eager)
    print("GPU_BASELINE " + json.dumps(res), flush=True)


def gpu_baseline_record(a, local):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[local]
               if os.environ.get("CUDA_VISIBLE_DEVICES") else str(local))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    budget = int(os.environ.get("JAT_GPU_BASELINE_TIMEOUT", "420"))
    cmd = [sys.executable, os.path.abspath(__file__), "--mode", "gpu_baseline", "--steps", "5", "--warmup", "3", "--norm", a.norm]
    out, note = "", None
    try:
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=budget)
        out = r.stdout
        if r.returncode != 0:
            note = f"exit code {r.returncode}: {r.stderr[-300:]}"
    except subprocess.TimeoutExpired as e:
        out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
        note = f"stopped after the {budget} s budget (results so far kept)"
    rec = None
    for line in out.splitlines():
        if line.startswith("GPU_BASELINE ") or line.startswith("GPU_BASELINE_PARTIAL "):
            rec = json.loads(line.split(" ", 1)[1])
    if rec is None:
        return {"unavailable": note or "no output"}
    if note:
        rec["note"] = note
    return rec


# ------------------------------------------------------------------------------------------------ headline
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--norm", default="layernorm", choices=["layernorm", "rmsnorm"])
    ap.add_argument("--graph", action="store_true", help="replay the K timed steps from one captured CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--mode", default="sample", choices=["sample", "train", "long", "gpu_baseline"],
                    help="sample = headline CFG denoise step (configs[2]) + the sub-records of --sub; train = only the DDP training "
                         "step (configs[3]); long = only the 10-minute track (configs[4]); gpu_baseline = only stock PyTorch on this GPU")
    ap.add_argument("--sub", default="auto", help="sub-records added to the headline line: auto (train and long at every N; v2, "
                                                  "gpu_baseline at N = 1), none, or a comma list of train,long,v2,gpu_baseline")
    a = ap.parse_args()
    K, W = a.steps, max(a.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": "configs[2]: v3mod2 DiT 1280/28/20Q/4KV, CFG=3.0 sampler step, batch 28 x [1024,1378] per GPU "
                          "(B_eff 56, 19320 token rows)", "norm": a.norm, "cfg_scale": CFG_SCALE, "batch_per_gpu": B,
              "latent": [C, T], "parallelism": f"batch-sharded x{world} (no collective)",
              "l2": "per-step working set (1.5 GB bf16 weights + >0.5 GB activations) exceeds the 126 MB L2; no flush needed"}

    if a.impl == "reference":
        return reference_arm(a, K, W, rank, world, config)
    if a.mode == "train":
        return train_main(a, K, W)
    if a.mode == "long":
        return long_main(a, K, W)
    if a.mode == "gpu_baseline":
        return gpu_baseline_main(a, K, W)

    import torch
    import jat_b200
    from jat_b200 import _lib as L
    from jat_b200.sampler import _Plan
    D_ = Dist()
    dev, local = D_.dev, D_.local
    model = build_model(dev, a.norm)
    total = W + K
    plan = _Plan(model, B, C, T, total, CFG_SCALE, dev)
    g = torch.Generator(device=dev).manual_seed(123 + rank)
    plan.z.copy_(torch.randn(B, C, T, generator=g, device=dev))
    plan.lr.copy_(torch.randn(B, C, T, generator=g, device=dev))
    plan.mod = model._engine.modulation(plan.ws, plan.t_curr)
    ctx = L.context(local)
    lib = L.load()

    run_steps(model, plan, 0, W)
    graph = None
    if a.graph:
        torch.cuda.synchronize(dev)
        z_keep = plan.z.clone()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            run_steps(model, plan, W, K)
        plan.z.copy_(z_keep)
    D_.barrier()
    clocks = ClockSampler(local)
    l0 = lib.jat_launch_count(ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if graph is not None:
        graph.replay()
    else:
        run_steps(model, plan, W, K)
    e1.record()
    D_.barrier()
    clk = clocks.stop()
    launches = lib.jat_launch_count(ctx) - l0
    if graph is not None:
        launches = None
    ms = D_.max(e0.elapsed_time(e1))
    assert torch.isfinite(plan.z).all(), "non-finite latents"

    # ---- per-kernel-class timing inside a real step (CUDA events on the launching stream)
    L.profile_begin(local)
    pk = min(K, 5)
    run_steps(model, plan, W, pk)
    prof = L.profile_end(local)
    if launches is None:
        launches = sum(c for _, c in prof.values()) // pk * K
    pkz = peaks()
    work, step_flops = step_work()
    kernels, tot_ms = {}, sum(v[0] for v in prof.values())
    for name, (tms, cnt) in prof.items():
        per_step = tms / pk
        ent = {"ms_per_step": round(per_step, 4), "launches_per_step": cnt // pk, "share": round(tms / tot_ms, 4)}
        if name in work:
            unit, amount = work[name]
            if unit == "flops":
                ent.update(bound="tensor", achieved=round(amount / per_step / 1e9, 1), peak=pkz["tf"], unit="TFLOP/s")
            else:
                ent.update(bound="hbm", achieved=round(amount / per_step / 1e6, 1), peak=pkz["hbm"], unit="GB/s")
            ent["frac"] = round(ent["achieved"] / ent["peak"], 4)
        kernels[name] = ent
    top = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    roof = {k: kernels[top][k] for k in ("bound", "achieved", "peak", "unit", "frac")}
    roof.update(kernel=top, peak_source=f"MEASURED_PEAKS.json sustained ({pkz['src']})", traffic=ncu_traffic(top),
                share_of_step=kernels[top]["share"])
    roof["traffic_source"] = (getattr(ncu_traffic, "source", "profiles/") + ": dram__bytes_read.sum + dram__bytes_write.sum per launch "
                              "(ncu --set full, cold cache), mean over the launches of this class in the capture")
    kernel_sum_ms = tot_ms / pk

    # ---- e2e through the public sampler API with host buffers
    e2e = None
    if not a.no_e2e:
        lr_host = torch.randn(B, C, T).pin_memory()
        out_host = torch.empty(B, C, T).pin_memory()
        del plan
        jat_b200.flow_matching_sample(model, lr_host.to(dev, non_blocking=True), num_steps=K, cfg_scale=CFG_SCALE,
                                      device=dev, verbose=False)  # untimed: builds the plan / graph
        D_.barrier()
        t0 = time.perf_counter()
        zf = jat_b200.flow_matching_sample(model, lr_host.to(dev, non_blocking=True), num_steps=K, cfg_scale=CFG_SCALE,
                                           device=dev, verbose=False)
        out_host.copy_(zf, non_blocking=True)
        torch.cuda.synchronize(dev)
        dt = D_.max(time.perf_counter() - t0)
        nbytes = B * C * T * 4
        e2e = {"value": round(world * K / dt, 3), "unit": "steps/s", "h2d_bytes_per_step": nbytes // K,
               "d2h_bytes_per_step": nbytes // K, "api": "flow_matching_sample(model, lr_latent[host pinned]) -> host, "
               f"{K} steps per call (one CUDA graph replay); bytes amortised over the call's steps"}
        # `value` times the engine's launch-by-launch loop, `e2e` the public API: they must tell the same story
        value_now = world * K / (ms / 1e3)
        e2e["agrees_with_value"] = bool(abs(e2e["value"] - value_now) <= 0.05 * value_now)
        model.__dict__.pop("_sampler_plans", None)

    # ---- CPU baseline (rank 0, N == 1)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        threads = host_threads()
        weights = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
        cpu_baseline_sample(weights, a.norm)
        ts = cpu_baseline_sample(weights, a.norm)
        del weights
        cpu = {"value": round(1.0 / (ts * B / CPU_SAMPLE_B), 6), "unit": "steps/s", "cores": threads, "kind": "port",
               "sample": f"torch (ATen) fp32 restatement of the reference, 1 CFG denoise step at batch {CPU_SAMPLE_B} (B_eff=2, 690 token "
                         f"rows, depth 28) took {ts:.2f} s on {threads} threads; scaled x{B // CPU_SAMPLE_B} to the batch-28 step"}

    # ---- sub-records: the other BASELINE configs and the stock-PyTorch comparator
    subs = a.sub.split(",") if a.sub not in ("auto", "none") else (["train", "long"] + (["v2", "gpu_baseline"] if world == 1 else [])
                                                                    if a.sub == "auto" else [])
    extra = {}

    def sub(name, fn):
        if name not in subs:
            return
        try:
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            extra[name] = fn()
            if isinstance(extra[name], dict):
                extra[name]["bench_wall_s"] = round(time.perf_counter() - t0, 1)
        except SystemExit:
            raise
        except Exception as e:   # a sub-record must not lose the headline; under DDP every rank takes the same path
            if world > 1:
                raise
            extra[name] = {"unavailable": f"{type(e).__name__}: {str(e)[:300]}"}

    sub("long", lambda: long_record(a, 1, 1, D_, model))
    sub("v2", lambda: v2_record(a, D_))
    model._engine.workspaces.clear()
    model.__dict__.pop("_sampler_plans", None)
    torch.cuda.empty_cache()
    sub("train", lambda: train_record(a, min(K, 20), max(W, 3), D_, model)[0])
    if "gpu_baseline" in subs and rank == 0:
        del model
        torch.cuda.empty_cache()
        extra["gpu_baseline"] = gpu_baseline_record(a, local)
        gb = extra["gpu_baseline"]
        ours_ms = ms / K
        for k, v in list(gb.items()):
            if isinstance(v, dict) and "ms_per_step" in v:
                mine = extra.get("train", {}).get("ms_per_step") if k.startswith("train") else ours_ms
                if mine:
                    v["speedup_ours_vs_this"] = round(v["ms_per_step"] / mine, 2)

    if rank == 0:
        value = world * K / (ms / 1e3)
        line = {"metric": METRIC, "value": round(value, 3), "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic", "config": config, "clocks": clk, "e2e": e2e,
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "kernels": kernels,
                "kernel_sum_ms_per_step": round(kernel_sum_ms, 4),
                "kernel_time_share_of_step": round(kernel_sum_ms / (ms / K), 4),
                "step_tflops": round(step_flops / (ms / K) / 1e9, 1),
                "step_tensor_frac_sustained": round(step_flops / (ms / K) / 1e9 / pkz["tf"], 4),
                "graph": bool(a.graph)}
        line.update(extra)
        print(json.dumps(line), flush=True)
    D_.close()


if __name__ == "__main__":
    main()
