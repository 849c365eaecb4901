#!/usr/bin/env python
"""Headline benchmark: v3mod2 DiT denoise steps/sec (BASELINE.json metric).

One "step" = one Euler step of the flow-matching sampler with CFG on a batch of B=28 latent chunks
[28, 1024, 1378]: one DiT forward over B_eff = 56 (M = 19320 token rows, 766 M params, depth 28) plus
the fused CFG/velocity/Euler update.  Weights random-init (zero-init layers re-randomised), synthetic
unit-variance latents.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU); the path shards by batch (every rank denoises its
own B=28 batch, no data-path collective) -> weak scaling, value = N*K / max-over-ranks time.

JSON keys beyond the base contract:
  e2e          same metric through the public API `flow_matching_sample` with HOST (pinned) buffers:
               H2D of the condition latent + K steps + D2H of the result inside the timed region
  roofline     dominant kernel class of the step, timed live with CUDA events (jat_profile_*)
  kernels      the same for every kernel class (share of the step, achieved TFLOP/s or GB/s)
  cpu_baseline the numpy oracle port timed on the host cores on a bounded sample
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
B, C, T, CFG_SCALE = 28, 1024, 1378, 3.0
METRIC = "v3mod2 DiT denoise steps/sec (B=28, CFG=3.0, 766M params)"


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return dict(hbm=p["hbm_gbs"], tf=p["bf16_tflops_sustained"], tf_burst=p["bf16_tflops"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, src="fallback")


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
    for tag in ("r1c", "r1b"):   # latest committed capture first
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


def build_model(dev, norm):
    import torch
    import jat_b200
    cls = jat_b200.JaT_AudioSR_V2 if norm == "layernorm" else jat_b200.JaT_AudioSR_V3
    torch.manual_seed(0)
    with torch.device(dev):
        model = cls(**CFG)
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


def cpu_baseline_sample(weights, threads):
    """One CFG denoise step of the numpy oracle at batch 1 (B_eff = 2, 690 token rows, full depth)."""
    import numpy as np
    from oracle import dit_oracle as O
    rng = np.random.default_rng(0)
    z = rng.standard_normal((1, C, T), dtype=np.float32)
    lr = rng.standard_normal((1, C, T), dtype=np.float32)
    t0 = time.perf_counter()
    tb = np.full((2,), 0.37, np.float32)
    both = O.dit_forward(weights, np.concatenate([z, z]), tb, np.concatenate([lr, np.zeros_like(lr)]),
                         num_q_heads=20, num_kv_heads=4)
    O.euler_cfg_update(z, both[:1], both[1:], CFG_SCALE, np.float32(0.37), np.float32(0.02))
    return time.perf_counter() - t0


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
    ap.add_argument("--mode", default="sample", choices=["sample", "train", "long"],
                    help="sample = headline CFG denoise step (configs[2]); train = DDP training step (configs[3]); "
                         "long = 10-minute track, chunked 50-step CFG inference sharded over the GPUs (configs[4])")
    a = ap.parse_args()
    K, W = a.steps, max(a.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "configs[2]: v3mod2 DiT 1280/28/20Q/4KV, CFG=3.0 sampler step, batch 28 x [1024,1378] per GPU "
                          "(B_eff 56, 19320 token rows)", "norm": a.norm, "cfg_scale": CFG_SCALE, "batch_per_gpu": B,
              "latent": [C, T], "parallelism": f"batch-sharded x{world} (no collective)",
              "l2": "per-step working set (1.5 GB bf16 weights + >0.5 GB activations) exceeds the 126 MB L2; no flush needed"}

    if a.impl == "reference":
        return reference_arm(a, K, W, rank, world, config)
    if a.mode == "train":
        return train_main(a, K, W, rank, world, local)
    if a.mode == "long":
        return long_main(a, K, W, rank, world, local)

    import torch
    import jat_b200
    from jat_b200 import _lib as L
    from jat_b200.sampler import _Plan
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", init_method="env://")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    model = build_model(dev, a.norm)
    total = W + K
    plan = _Plan(model, B, C, T, total, CFG_SCALE, dev)
    g = torch.Generator(device=dev).manual_seed(123 + rank)
    plan.z.copy_(torch.randn(B, C, T, generator=g, device=dev))
    plan.lr.copy_(torch.randn(B, C, T, generator=g, device=dev))
    plan.mod = model._engine.modulation(plan.ws, plan.t_curr)
    ctx = L.context(local)
    lib = L.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    run_steps(model, plan, 0, W)
    graph = None
    if a.graph:
        torch.cuda.synchronize(dev)
        z_keep = plan.z.clone()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            run_steps(model, plan, W, K)
        plan.z.copy_(z_keep)
    barrier()
    clocks = ClockSampler(local)
    l0 = lib.jat_launch_count(ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if graph is not None:
        graph.replay()
    else:
        run_steps(model, plan, W, K)
    e1.record()
    barrier()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    launches = lib.jat_launch_count(ctx) - l0
    if graph is not None:
        launches = None
    if world > 1:
        tmax = torch.tensor([ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
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

    # ---- e2e through the public sampler API with host buffers
    e2e = None
    if not a.no_e2e:
        lr_host = torch.randn(B, C, T).pin_memory()
        out_host = torch.empty(B, C, T).pin_memory()
        del plan
        jat_b200.flow_matching_sample(model, lr_host.to(dev, non_blocking=True), num_steps=K, cfg_scale=CFG_SCALE,
                                      device=dev, verbose=False)  # untimed: builds the plan / graph
        barrier()
        t0 = time.perf_counter()
        zf = jat_b200.flow_matching_sample(model, lr_host.to(dev, non_blocking=True), num_steps=K, cfg_scale=CFG_SCALE,
                                           device=dev, verbose=False)
        out_host.copy_(zf, non_blocking=True)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if world > 1:
            tm = torch.tensor([dt], device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dt = float(tm.item())
        nbytes = B * C * T * 4
        e2e = {"value": round(world * K / dt, 3), "unit": "steps/s", "h2d_bytes_per_step": nbytes // K,
               "d2h_bytes_per_step": nbytes // K, "api": "flow_matching_sample(model, lr_latent[host pinned]) -> host, "
               f"{K} steps per call; bytes amortised over the call's steps"}

    # ---- CPU baseline (rank 0, N == 1)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        weights = {k: v.detach().float().cpu().numpy() for k, v in model.state_dict().items()
                   if not k.endswith(("cos_cached", "sin_cached"))}
        ts = cpu_baseline_sample(weights, os.cpu_count())
        cpu = {"value": round(1.0 / (ts * B), 6), "unit": "steps/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"numpy fp32 oracle, 1 CFG denoise step at batch 1 (B_eff=2, 690 token rows, depth 28) took "
                         f"{ts:.2f} s; scaled x{B} to the batch-28 step"}

    if rank == 0:
        value = world * K / (ms / 1e3)
        line = {"metric": METRIC, "value": round(value, 3), "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic", "config": config, "clocks": clk, "e2e": e2e,
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "kernels": kernels,
                "step_tflops": round(step_flops / (ms / K) / 1e9, 1),
                "step_tensor_frac_sustained": round(step_flops / (ms / K) / 1e9 / pkz["tf"], 4),
                "graph": bool(a.graph)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def train_main(a, K, W, rank, world, local):
    """BASELINE configs[3]: v3mod2 training step, batch 28 per GPU, x-prediction MSE flow-matching loss
    (train_ddp_v3mod2.py:842-930 without the perceptual losses / logging), DDP over NCCL for N > 1.
    One step = zero_grad + forward + loss + backward (+ gradient all-reduce) + clip_grad_norm_ + AdamW + weight re-pack."""
    import torch
    import jat_b200
    from jat_b200 import _lib as L
    if world > 1:
        import torch.distributed as dist
        # NCCL kernels on a high-priority stream: they get SMs at the next kernel boundary instead of queueing behind the
        # library's persistent GEMM grids (4 GPUs: 57.4 -> 56.1 ms / step); JAT_NCCL_HIPRI=0 = torch's default
        if os.environ.get("JAT_NCCL_HIPRI", "1") == "1":
            dist.init_process_group("nccl", init_method="env://",
                                    pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
        else:
            dist.init_process_group("nccl", init_method="env://")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cfg = dict(CFG)  # dropout 0.1, drop_path 0.05: the reference's training configuration (train_ddp_v3mod2.py:343-355)
    from jat_b200 import training
    cls = jat_b200.JaT_AudioSR_V2 if a.norm == "layernorm" else jat_b200.JaT_AudioSR_V3
    torch.manual_seed(0)
    with torch.device(dev):
        model = cls(**cfg)
    g = torch.Generator(device=dev).manual_seed(1)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "adaLN_modulation.1" in name or name.startswith("final_layer.1"):
                p.copy_(torch.randn(p.shape, generator=g, device=dev) * 0.02)
    model.train()
    model.grad_handoff = os.environ.get("JAT_GRAD_HANDOFF", "view")   # zero-copy .grad (the step calls zero_grad(set_to_none=True))
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(
            model, device_ids=[local], find_unused_parameters=False,
            # the reference's call (train_ddp_v3mod2.py:822) + two DDP options measured at 2 GPUs: gradients as views of
            # the all-reduce buckets (no copy back: 59.6 -> 57.3 ms / step) and 200 MB buckets (56.9 ms)
            bucket_cap_mb=int(os.environ.get("JAT_DDP_BUCKET_MB", "200")),
            gradient_as_bucket_view=os.environ.get("JAT_DDP_BUCKET_VIEW", "1") == "1")
        # (torch's bf16_compress_hook was measured too: 74.0 ms / step at 8 GPUs against 58.5 ms with the plain f32 all-reduce)
    fused_opt = os.environ.get("JAT_BENCH_TORCH_OPT", "0") == "0"
    if fused_opt:   # clip_grad_norm_(1.0) + AdamW + bf16 re-pack in two multi-tensor passes (jat_b200.FusedAdamW)
        opt = jat_b200.FusedAdamW(model.parameters(), lr=5e-5, weight_decay=0.1, max_grad_norm=1.0, model=model)
    else:           # the reference's own calls, train_ddp_v3mod2.py:709, 926-928
        opt = torch.optim.AdamW(model.parameters(), lr=5e-5, weight_decay=0.1, fused=True)
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(W):
        step()
    barrier()
    clocks = ClockSampler(local)
    ctx, lib = L.context(local), L.load()
    l0 = lib.jat_launch_count(ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        loss = step()
    e1.record()
    barrier()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    launches = lib.jat_launch_count(ctx) - l0
    if world > 1:
        tmax = torch.tensor([ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    assert torch.isfinite(loss).all()
    L.profile_begin(local)
    pk = min(K, 3)
    for _ in range(pk):
        step()
    prof = L.profile_end(local)
    pkz = peaks()
    tot = sum(v[0] for v in prof.values())
    kernels = {n: {"ms_per_step": round(v[0] / pk, 3), "launches_per_step": v[1] // pk, "share_of_kernel_time": round(v[0] / tot, 4)}
               for n, v in prof.items()}
    # roofline of the training step's dominant kernel classes (algorithmic work of step_work() at B_eff = B; the weight-gradient
    # GEMMs do exactly the forward GEMMs' FLOPs; the parameter update moves 30 B and the norm pass 4 B per parameter)
    fw, _ = step_work(B)
    fwd_gemm = sum(fw[k][1] for k in ("gemm_bias_act", "gemm_qkv_rope", "gemm_gate_residual", "gemm_unpatchify"))
    n_params = sum(p.numel() for p in model.parameters())
    roofs = {}
    if "gemm_accum" in kernels:
        tfs = fwd_gemm / (kernels["gemm_accum"]["ms_per_step"] * 1e-3) / 1e12
        roofs["gemm_accum"] = {"bound": "tensor", "achieved": round(tfs, 1), "peak": pkz["tf"], "unit": "TFLOP/s",
                               "frac": round(tfs / pkz["tf"], 4)}
    if "optimizer" in kernels:
        gbs = n_params * 34.0 / (kernels["optimizer"]["ms_per_step"] * 1e-3) / 1e9
        roofs["optimizer"] = {"bound": "hbm", "achieved": round(gbs, 1), "peak": pkz["hbm"], "unit": "GB/s",
                              "frac": round(gbs / pkz["hbm"], 4)}
    top = max(kernels, key=lambda k: kernels[k]["ms_per_step"]) if kernels else None
    roofline = dict(roofs.get(top, {}), kernel=top, share_of_step=kernels[top]["share_of_kernel_time"],
                    peak_source=f"MEASURED_PEAKS.json sustained ({pkz['src']})") if top in roofs else None
    # algorithmic FLOPs: forward 1023.85 MFLOP/token, backward = 2x forward, no recompute counted
    N = (T + 3) // 4
    step_flops = 3 * 1023.85e6 * B * N
    if rank == 0:
        value = world * K / (ms / 1e3)
        line = {"metric": "v3mod2 DDP training steps/sec (batch 28 per GPU, x-prediction MSE flow-matching loss)",
                "value": round(value, 3), "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": round(ms / K, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "configs[3]: v3mod2 DiT 1280/28/20Q/4KV training step, batch 28 x [1024,1378] per GPU "
                                       "(9660 token rows): normalise + cond-noise + flow-matching mix, forward (Dropout 0.1, DropPath 0.05), "
                                       "MSE x-prediction loss, backward (grad_handoff=" + model.grad_handoff + "), clip_grad_norm_(1.0) + AdamW + bf16 weight re-pack "
                                       + ("(jat_b200.FusedAdamW: 2 multi-tensor passes)" if fused_opt else "(torch: foreach clip, fused AdamW, re-cast)"),
                           "norm": a.norm, "dropout": cfg["dropout"], "drop_path": cfg["drop_path_rate"],
                           "cond_noise_ratio": 0.05,
                           "parallelism": f"DDP x{world} (NCCL gradient all-reduce on a high-priority stream, 200 MB buckets, gradient_as_bucket_view)" if world > 1 else "single GPU"},
                "clocks": clk, "gpu_launches": int(launches), "loss": round(float(loss.item()), 5),
                "step_tflops_per_gpu": round(step_flops / (ms / K) / 1e9, 1),
                "step_tensor_frac_sustained": round(step_flops / (ms / K) / 1e9 / pkz["tf"], 4),
                "roofline": roofline, "rooflines": roofs,
                "kernels": kernels, "kernel_ms_per_step": round(tot / pk, 2)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def long_main(a, K, W, rank, world, local):
    """BASELINE configs[4]: a 10-minute track (latent [1024, 51679] at 86.13 frames/s) cut into 43 chunks of 1378 frames
    (overlap 172, infer_test_v3m2.py:340-348), the chunks dealt round-robin over the GPUs, each GPU denoising its chunks as
    ONE batch with 50 CFG = 3.0 steps, all-gather of the finished chunk latents, crossfade + de-normalise
    (`jat_b200.chunked.sample_long`).  One "step" = the whole track; value = audio seconds per wall second."""
    import torch
    import jat_b200
    from jat_b200 import chunked
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", init_method="env://")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    model = build_model(dev, a.norm)
    seconds, frames = 600.0, 51679
    g = torch.Generator().manual_seed(5)
    track = (torch.randn(C, frames, generator=g) * 2.0 + 0.3).pin_memory()
    mean, std = torch.full((C,), 0.3), torch.full((C,), 2.0)

    def run():
        out = chunked.sample_long(model, track.to(dev, non_blocking=True), mean, std, mean, std, num_steps=50, cfg_scale=CFG_SCALE,
                                  device=dev)
        return out.cpu() if rank == 0 else out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
    for _ in range(max(W, 1)):
        run()
    barrier()
    clocks = ClockSampler(local)
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        out = run()
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tmax = torch.tensor([ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    if rank == 0:
        assert out.shape == (1, C, frames) and torch.isfinite(out).all()
        n_chunks = len(chunked.plan_chunks(frames))
        line = {"metric": "long-audio chunked inference: audio seconds per second (10-minute track, 50-step CFG=3.0)",
                "value": round(K * seconds / (ms / 1e3), 1), "unit": "audio s/s", "n_gpus": world, "steps": K, "warmup": max(W, 1),
                "ms_per_step": round(ms / K, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": f"configs[4]: 10-minute track = latent [1024, {frames}] -> {n_chunks} chunks of 1378 frames "
                                       f"(overlap 172), round-robin over {world} GPU(s), 50 Euler steps CFG=3.0 per chunk batch, "
                                       "all-gather + crossfade + de-normalise; host track in, host latent out",
                           "norm": a.norm, "chunks": n_chunks, "chunks_per_gpu": -(-n_chunks // world)},
                "clocks": clk, "wall_s": round(wall, 3)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def reference_arm(a, K, W, rank, world, config):
    """The reference's CPU implementation of the path = the numpy oracle port (the reference is pure
    PyTorch and is not present on the GPU box), all host threads, each step a bounded sample."""
    if rank != 0:
        return
    import numpy as np
    sys.path.insert(0, ROOT)
    rng = np.random.default_rng(0)
    D, depth, F = 1280, 28, 5120
    w = {}

    def lin(name, o, i, bias=True, std=None):
        s = std if std is not None else 1.0 / np.sqrt(i)
        w[name + ".weight"] = (rng.standard_normal((o, i), dtype=np.float32) * s)
        if bias:
            w[name + ".bias"] = rng.standard_normal((o,), dtype=np.float32) * 0.02
    lin("patch_embed.proj.0", 512, 8192)
    lin("patch_embed.proj.2", D, 512)
    lin("t_embedder.1", D, D)
    lin("t_embedder.3", D, D)
    for i in range(depth):
        p = f"blocks.{i}."
        lin(p + "attn.q_proj", D, D, False)
        lin(p + "attn.k_proj", 256, D, False)
        lin(p + "attn.v_proj", 256, D, False)
        lin(p + "attn.out_proj", D, D, False)
        lin(p + "mlp.0", F, D)
        lin(p + "mlp.3", D, F)
        lin(p + "adaLN_modulation.1", 6 * D, D, True, 0.02)
        if a.norm == "rmsnorm":
            w[p + "norm1.weight"] = np.ones(D, np.float32)
            w[p + "norm2.weight"] = np.ones(D, np.float32)
    lin("final_layer.1", 4096, D, True, 0.02)
    if a.norm == "rmsnorm":
        w["final_layer.0.weight"] = np.ones(D, np.float32)
    times = []
    budget = 240.0
    t_first = cpu_baseline_sample(w, os.cpu_count())  # warm-up 1 (also sizes the run)
    k_eff = K
    if (K + W) * t_first > budget:
        k_eff = max(1, int(budget / t_first) - W)
    for _ in range(max(W - 1, 0)):
        cpu_baseline_sample(w, os.cpu_count())
    for _ in range(k_eff):
        times.append(cpu_baseline_sample(w, os.cpu_count()))
    ts = sum(times) / len(times)
    value = 1.0 / (ts * B)
    sample = (f"numpy fp32 oracle port of the reference forward+update, each step = 1 CFG denoise step at batch 1 "
              f"(B_eff=2, 690 token rows, depth 28), mean {ts:.2f} s over {k_eff} steps, scaled x{B} to the batch-28 step")
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 6), "unit": "steps/s", "n_gpus": world,
            "steps": k_eff, "warmup": W, "ms_per_step": round(ts * B * 1e3, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": round(value, 6), "unit": "steps/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": round(value, 6), "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
