"""Checkpoint compatibility at the boundary, CPU only: our `save_checkpoint` -> the unmodified
reference model loads it strictly; reference-style checkpoints (DDP / torch.compile prefixes, pre-Dropout MLP keys,
missing 'config') -> our `load_model`; resume restores optimizer state and RNG streams."""
import contextlib
import io

import pytest
import torch

from tests._util import have_reference, import_reference, rerandomise_zero_init

CFG = dict(input_channels=16, cond_channels=16, patch_len=4, hidden_size=128, depth=2, num_q_heads=2, num_kv_heads=1,
           bottleneck_dim=64, mlp_ratio=2.0, dropout=0.1, drop_path_rate=0.05)


def _ours(cls_name, seed=0):
    import jat_b200
    torch.manual_seed(seed)
    return rerandomise_zero_init(getattr(jat_b200, cls_name)(**CFG), bf16_exact=False)


@pytest.mark.parametrize("cls_name", ["JaT_AudioSR_V2", "JaT_AudioSR_V3"])
def test_save_load_roundtrip_and_key_cleaning(cls_name, tmp_path):
    import jat_b200
    from jat_b200 import checkpoint as ck
    model = _ours(cls_name)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    path = tmp_path / "ck.pt"
    state = jat_b200.save_checkpoint(model, opt, None, epoch=3, step=1234, best_loss=0.5, path=str(path))
    assert set(state) == {"epoch", "global_step", "best_val_loss", "model_state_dict", "optimizer_state_dict",
                          "scaler_state_dict", "rng_state", "config"}
    assert state["config"] == CFG
    loaded = jat_b200.load_model(str(path), device="cpu", verbose=False)
    assert type(loaded).__name__ == cls_name and not loaded.training
    for (k, a), (k2, b) in zip(model.state_dict().items(), loaded.state_dict().items()):
        assert k == k2 and torch.equal(a, b), k
    # DDP + torch.compile prefixes and the pre-Dropout MLP indexing, no 'config' -> defaults would not fit, so pass one
    sd = {("module._orig_mod." + k).replace(".mlp.3.", ".mlp.2."): v for k, v in model.state_dict().items()}
    clean = ck.clean_state_dict(sd)
    assert set(clean) == set(model.state_dict())
    torch.save({"model_state_dict": sd, "config": CFG, "epoch": 1, "global_step": 2}, str(path))
    again = jat_b200.load_model(str(path), device="cpu", verbose=False)
    for (k, a), (_, b) in zip(model.state_dict().items(), again.state_dict().items()):
        assert torch.equal(a, b), k
    assert ck.DEFAULT_CONFIG["hidden_size"] == 1280 and ck.DEFAULT_CONFIG["depth"] == 28  # infer_test_v3m2.py:42-54


def test_resume_restores_optimizer_and_rng(tmp_path):
    from jat_b200 import checkpoint as ck
    model = _ours("JaT_AudioSR_V2")
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    for p in model.parameters():
        p.grad = torch.randn_like(p)
    opt.step()
    path = str(tmp_path / "resume.pt")
    ck.save_checkpoint(model, opt, None, epoch=7, step=99, best_loss=0.25, path=path)
    want_next = torch.rand(3)
    model2 = _ours("JaT_AudioSR_V2", seed=5)
    opt2 = torch.optim.AdamW(model2.parameters(), lr=1e-3)
    torch.manual_seed(12345)
    start_epoch, step, best = ck.resume(model2, opt2, None, path, device="cpu")
    assert (start_epoch, step, best) == (8, 99, 0.25)
    assert torch.equal(torch.rand(3), want_next)            # torch RNG stream continues where the checkpoint left it
    for a, b in zip(model.parameters(), model2.parameters()):
        assert torch.equal(a, b)
    s1, s2 = opt.state_dict()["state"], opt2.state_dict()["state"]
    assert all(torch.equal(s1[i]["exp_avg"], s2[i]["exp_avg"]) for i in s1)
    with pytest.raises(RuntimeError):                      # strict: a V3 model must not swallow a V2 checkpoint
        ck.resume(_ours("JaT_AudioSR_V3"), None, None, path, device="cpu")


@pytest.mark.skipif(not have_reference(), reason="reference not mounted")
@pytest.mark.parametrize("idx,cls_name", [(0, "JaT_AudioSR_V2"), (1, "JaT_AudioSR_V3")])
def test_interchange_with_reference_model(idx, cls_name, tmp_path):
    """ours -> reference (strict load into the unmodified module) and reference -> ours; forward of the reference on the
    loaded weights equals its forward on the source weights (the tensors really are the same)."""
    import jat_b200
    ref_cls = import_reference()[idx]
    ours = _ours(cls_name)
    path = str(tmp_path / "x.pt")
    jat_b200.save_checkpoint(ours, None, None, 0, 0, 1.0, path)
    ck = torch.load(path, weights_only=False)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = ref_cls(**ck["config"])
    ref.load_state_dict(ck["model_state_dict"])             # strict=True
    # reference -> ours, through a DDP-style prefix
    torch.save({"model_state_dict": {"module." + k: v for k, v in ref.state_dict().items()}, "config": ck["config"]}, path)
    back = jat_b200.load_model(path, device="cpu", verbose=False)
    assert type(back).__name__ == cls_name
    assert list(back.state_dict()) == list(ref.state_dict())
    for (k, a), (_, b) in zip(ref.state_dict().items(), back.state_dict().items()):
        assert a.shape == b.shape and torch.equal(a, b), k
