"""Host-side engine logic that needs no GPU: the packed weight copies follow the parameters (version counters, the dirty
flag a backward pass sets, explicit refresh), they are re-cast IN PLACE (pointers handed to the C ABI / captured CUDA graphs
stay valid), and the packed gradient buffers map back to the reference's parameter layout.  torch owns the memory, so all of
this runs on CPU tensors; no kernel is launched."""
import pytest
import torch

CFG = dict(input_channels=32, cond_channels=32, patch_len=4, hidden_size=128, depth=2, num_q_heads=4, num_kv_heads=2,
           bottleneck_dim=64, mlp_ratio=2.0, dropout=0.0, drop_path_rate=0.0)
CPU = torch.device("cpu")


@pytest.fixture(params=["JaT_AudioSR_V2", "JaT_AudioSR_V3"])
def model(request):
    import jat_b200
    torch.manual_seed(0)
    m = getattr(jat_b200, request.param)(**CFG)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn_like(p) * 0.02)
    return m


def _fresh(model, pk):
    return all(torch.equal(d, s.detach().to(d.dtype)) for d, s in zip(pk._dst, pk._src))


def test_packed_copies_mirror_every_parameter(model):
    pk = model._engine.weights(CPU)
    mirrored = {id(s) for s in pk._src}
    assert all(id(p) in mirrored for p in model.parameters()) and len(pk._src) == len(list(model.parameters()))
    assert _fresh(model, pk) and not pk.dirty
    blk = model.blocks[1]
    qd, kd = blk.attn.q_proj.out_features, blk.attn.k_proj.out_features
    wqkv = pk.keep["wqkv"][1]                                    # q | k | v rows of one block in one matrix
    assert torch.equal(wqkv[:qd], blk.attn.q_proj.weight.detach().bfloat16())
    assert torch.equal(wqkv[qd:qd + kd], blk.attn.k_proj.weight.detach().bfloat16())
    assert torch.equal(wqkv[qd + kd:], blk.attn.v_proj.weight.detach().bfloat16())
    D = model.hidden_size
    assert torch.equal(pk.keep["ada_w"][6 * D:], blk.adaLN_modulation[1].weight.detach().bfloat16())   # stacked over blocks


def test_version_bump_recasts_in_place(model):
    eng = model._engine
    pk = eng.weights(CPU)
    ptrs = [d.data_ptr() for d in pk._dst]
    with torch.no_grad():
        model.blocks[0].mlp[0].weight.mul_(1.5)                  # what torch.optim.AdamW (foreach) / load_state_dict do
    assert pk.stale(model, CPU)
    assert eng.weights(CPU) is pk and [d.data_ptr() for d in pk._dst] == ptrs and _fresh(model, pk)


def test_updates_that_bypass_version_counters_need_the_dirty_flag(model):
    """`p.data` writes (and torch.optim.AdamW(fused=True) on CUDA) do not bump Tensor._version: only the dirty flag -- set by
    every backward pass, or by model.refresh_packed_weights() -- makes the next forward re-cast."""
    eng = model._engine
    pk = eng.weights(CPU)
    model.final_layer[1].weight.data.mul_(2.0)
    assert not pk.stale(model, CPU) and not _fresh(model, pk)    # invisible ...
    assert eng.weights(CPU) is pk and not _fresh(model, pk)
    model.refresh_packed_weights()                               # ... until told
    assert pk.dirty
    assert eng.weights(CPU) is pk and _fresh(model, pk) and not pk.dirty
    model.final_layer[1].weight.data.mul_(0.5)
    pk.dirty = True                                              # what Engine._bwd_args does on every backward stage
    eng.weights(CPU)
    assert _fresh(model, pk)


def test_module_apply_drops_the_packed_copies(model):
    eng = model._engine
    pk = eng.weights(CPU)
    model.double()
    assert eng.packed is None
    model.float()
    assert eng.weights(CPU) is not pk


def test_packed_gradients_are_views_in_the_parameter_layout(model):
    from jat_b200.engine import PackedGrads
    g = PackedGrads(model, CPU)
    for p in model.parameters():
        v = g.by_param[p]
        assert v.shape == p.shape and v.dtype == torch.float32 and v.is_contiguous()
    blk = model.blocks[0]
    qd = blk.attn.q_proj.out_features
    packed = g.keep["wqkv"][0]
    assert g.by_param[blk.attn.q_proj.weight].data_ptr() == packed.data_ptr()
    assert g.by_param[blk.attn.k_proj.weight].data_ptr() == packed[qd:].data_ptr()
    g.by_param[blk.attn.v_proj.weight].fill_(3.0)
    g.zero_()
    assert float(packed.abs().sum()) == 0.0


def test_backward_without_forward_fails_loudly(model):
    from jat_b200 import _lib as L
    with pytest.raises(L.JatError):
        model._engine.backward(torch.zeros(1, 32, 8), 1, 8)
