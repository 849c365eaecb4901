"""Checkpoint compatibility with the reference scripts -- the on-disk format at the boundary.

  load_model       <->  infer_test_v3m2.py:33-94 (`load_model`) + the legacy key remap of infer_test_v3.py:58-69
  save_checkpoint  <->  train_ddp_v3mod2.py:1120-1148 (`save_checkpoint`, "Perfect Resume" dict layout)
  resume           <->  train_ddp_v3mod2.py:752-810 (strict state-dict load + optimizer / scaler / RNG restore)

A checkpoint written here loads into the unmodified reference model with `strict=True` (same keys, shapes and
persistent RoPE buffers), and a checkpoint written by the reference -- raw, torch.compile'd (`_orig_mod.`) or DDP-wrapped
(`module.`), with or without the pre-Dropout MLP indexing -- loads into the drop-in modules.  Host-side logic only: no
device code is involved until the loaded model is called.
"""
from __future__ import annotations

import random

import numpy as np
import torch

from .models import JaT_AudioSR_V2, JaT_AudioSR_V3

# infer_test_v3m2.py:42-54: the configuration assumed when a checkpoint carries no 'config'
DEFAULT_CONFIG = dict(input_channels=1024, cond_channels=1024, patch_len=4, hidden_size=1280, depth=28, num_q_heads=20,
                      num_kv_heads=4, bottleneck_dim=512, mlp_ratio=4.0, dropout=0.1, drop_path_rate=0.05)


def clean_state_dict(state_dict, verbose=False):
    """Key normalisation the reference loaders apply: strip the torch.compile (`_orig_mod.`) and DDP (`module.`) prefixes
    (infer_test_v3m2.py:63-71; substring replace, like the reference) and remap the pre-Dropout MLP layout
    `blocks.*.mlp.2.{weight,bias}` -> `mlp.3.*` (infer_test_v3.py:58-69)."""
    sd = state_dict
    for _ in range(2):  # either nesting order: DDP(compile(model)) or compile(DDP(model))
        if any(k.startswith("_orig_mod.") for k in sd):
            sd = {k.replace("_orig_mod.", ""): v for k, v in sd.items()}
            if verbose:
                print("  Removed torch.compile prefix (_orig_mod.)")
        if any(k.startswith("module.") for k in sd):
            sd = {k.replace("module.", ""): v for k, v in sd.items()}
            if verbose:
                print("  Removed DDP prefix (module.)")
    if any(".mlp.2.weight" in k or ".mlp.2.bias" in k for k in sd) and not any(".mlp.3." in k for k in sd):
        sd = {k.replace(".mlp.2.", ".mlp.3."): v for k, v in sd.items()}
        if verbose:
            print("  Remapping: mlp.2 -> mlp.3 (pre-Dropout checkpoint)")
    return dict(sd)


def model_class_for(state_dict):
    """RMSNorm weights present -> the V3 class (jat_audiosr_v3.py), otherwise the LayerNorm V2 class."""
    return JaT_AudioSR_V3 if any(k.endswith("norm1.weight") for k in state_dict) else JaT_AudioSR_V2


def load_model(checkpoint_path, device="cuda", model_class=None, verbose=True):
    """Drop-in for the reference `load_model(checkpoint_path, device)`: returns the model in eval mode on `device`.
    `model_class=None` picks V2 / V3 from the checkpoint's keys (the reference script hard-codes its own class)."""
    if verbose:
        print(f"Loading checkpoint from: {checkpoint_path}")
    checkpoint = torch.load(checkpoint_path, map_location=device, weights_only=False)
    config = dict(checkpoint.get("config", DEFAULT_CONFIG))
    state_dict = clean_state_dict(checkpoint["model_state_dict"], verbose)
    cls = model_class or model_class_for(state_dict)
    model = cls(**config).to(device)
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    if verbose:
        for name, keys in (("Missing", missing), ("Unexpected", unexpected)):
            if keys:
                print(f"  {name} keys: {len(keys)}" + ("".join(f"\n    - {k}" for k in keys) if len(keys) <= 5 else ""))
        print(f"Model loaded (Epoch {checkpoint.get('epoch', 0)}, Step {checkpoint.get('global_step', 0)})")
    model.eval()
    return model


def model_config(model):
    """The constructor keywords of a drop-in model (what the reference stores as `TrainConfig.model_params`)."""
    blk = model.blocks[0]
    return dict(input_channels=model.input_channels, cond_channels=model.cond_channels, patch_len=model.patch_len,
                hidden_size=model.hidden_size, depth=len(model.blocks), num_q_heads=blk.attn.num_q_heads,
                num_kv_heads=blk.attn.num_kv_heads, bottleneck_dim=model.patch_embed.proj[0].out_features,
                mlp_ratio=blk.mlp[0].out_features / model.hidden_size, dropout=model.dropout_p,
                drop_path_rate=model.drop_path_rate)


def save_checkpoint(model, optimizer, scaler, epoch, step, best_loss, path, config=None):
    """Same dict layout as the reference's `save_checkpoint` (train_ddp_v3mod2.py:1120-1148)."""
    net = model.module if hasattr(model, "module") else model
    model_state = net.state_dict()
    if any(k.startswith("_orig_mod.") for k in model_state):
        model_state = {k.replace("_orig_mod.", ""): v for k, v in model_state.items()}
    rng_state = {"python": random.getstate(), "numpy": np.random.get_state(), "torch": torch.get_rng_state(),
                 "cuda_all": torch.cuda.get_rng_state_all() if torch.cuda.is_available() else []}
    base = getattr(net, "_orig_mod", net)
    state = {"epoch": epoch, "global_step": step, "best_val_loss": best_loss, "model_state_dict": model_state,
             "optimizer_state_dict": optimizer.state_dict() if optimizer is not None else None,
             "scaler_state_dict": scaler.state_dict() if scaler is not None else None, "rng_state": rng_state,
             "config": dict(config) if config is not None else model_config(base)}
    torch.save(state, path)
    return state


def resume(model, optimizer, scaler, path, device="cuda", restore_rng=True):
    """Perfect resume (train_ddp_v3mod2.py:752-810): strict model load, optimizer / scaler state, RNG streams.
    Returns (start_epoch, global_step, best_val_loss)."""
    checkpoint = torch.load(path, map_location=device, weights_only=False)
    state_dict = checkpoint["model_state_dict"]
    if any(k.startswith("_orig_mod.") for k in state_dict):
        state_dict = {k.replace("_orig_mod.", ""): v for k, v in state_dict.items()}
    net = model.module if hasattr(model, "module") else model
    net.load_state_dict(state_dict)  # strict, like the reference
    if optimizer is not None and checkpoint.get("optimizer_state_dict") is not None:
        optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
    if scaler is not None and checkpoint.get("scaler_state_dict") is not None:
        scaler.load_state_dict(checkpoint["scaler_state_dict"])
    if restore_rng and "rng_state" in checkpoint:
        rs = checkpoint["rng_state"]
        try:
            random.setstate(rs["python"])
            np.random.set_state(rs["numpy"])
            torch.set_rng_state(rs["torch"].to(torch.uint8).cpu())
            if rs.get("cuda_all") and torch.cuda.is_available():
                torch.cuda.set_rng_state_all([s.to(torch.uint8).cpu() for s in rs["cuda_all"]])
            elif "cuda" in rs and torch.cuda.is_available():
                torch.cuda.set_rng_state(rs["cuda"].to(torch.uint8).cpu())
        except Exception as e:  # the reference also carries on (train_ddp_v3mod2.py:803-806)
            print(f"Warning: could not restore RNG states: {e}")
    return checkpoint["epoch"] + 1, checkpoint["global_step"], checkpoint.get("best_val_loss", float("inf"))
