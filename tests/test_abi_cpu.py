"""CPU-only checks of the drop-in boundary: the C-ABI library exports exactly what include/jat_b200.h declares,
the ctypes binding covers every declaration, and the nn.Module mirror has the reference's parameter layout.
No compute kernels are launched here."""
import ctypes
import os
import re

import pytest
import torch

from tests._util import have_reference, import_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "jat_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|void|int64_t|uint32_t|const char\*)\s+(jat_\w+)\s*\(([^;{]*)\)\s*;", src):
        args = m.group(2).strip()
        decls[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return decls


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from jat_b200 import _lib
    return _lib


def test_header_declares_what_binding_binds(lib):
    decls = header_functions()
    assert set(decls) == set(lib.SIGNATURES), set(decls) ^ set(lib.SIGNATURES)
    for name, nargs in decls.items():
        assert len(lib.SIGNATURES[name][1]) == nargs, name


def test_library_exports_every_symbol(lib):
    dll = ctypes.CDLL(lib.LIB_PATH)
    for name in header_functions():
        assert getattr(dll, name) is not None
    assert dll.jat_abi_version() == 1


def test_structs_match_header_layout(lib, tmp_path):
    """sizeof / offsetof as a C compiler sees include/jat_b200.h == the ctypes mirrors in _lib.py."""
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "jat_b200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(jat_gemm_epilogue),'
        ' offsetof(jat_gemm_epilogue, gate),'
        ' offsetof(jat_gemm_epilogue, t_out), sizeof(jat_dit_weights), offsetof(jat_dit_weights, pe_w1),'
        ' offsetof(jat_dit_weights, rope_sin), sizeof(jat_dit_workspace), offsetof(jat_dit_workspace, block_out),'
        ' offsetof(jat_gemm_epilogue, drop_seed), offsetof(jat_gemm_epilogue, gate_rowscale), sizeof(jat_dit_saved),'
        ' offsetof(jat_dit_saved, seed), offsetof(jat_dit_saved, dp_scale), sizeof(jat_dit_bwd_scratch));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    E, W, S, V = lib.GemmEpilogue, lib.DitWeights, lib.DitWorkspace, lib.DitSaved
    want = [ctypes.sizeof(E), E.gate.offset, E.t_out.offset, ctypes.sizeof(W), W.pe_w1.offset, W.rope_sin.offset,
            ctypes.sizeof(S), S.block_out.offset, E.drop_seed.offset, E.gate_rowscale.offset, ctypes.sizeof(V),
            V.seed.offset, V.dp_scale.offset, ctypes.sizeof(lib.DitBwdScratch)]
    assert got == want


def test_adamw_table_entry_is_eight_int64_words(tmp_path):
    """optim.FusedAdamW writes the device table of jat_adamw_step as an int64 [n, 8] tensor: word k of a row must be field k
    of jat_adamw_tensor (pointers, numel, packed_dtype | vec_ok << 32, the two f32 bias corrections packed into word 7)."""
    import subprocess
    src = tmp_path / "opt_layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "jat_b200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(jat_adamw_tensor),'
        ' offsetof(jat_adamw_tensor, param), offsetof(jat_adamw_tensor, grad), offsetof(jat_adamw_tensor, exp_avg),'
        ' offsetof(jat_adamw_tensor, exp_avg_sq), offsetof(jat_adamw_tensor, packed), offsetof(jat_adamw_tensor, numel),'
        ' offsetof(jat_adamw_tensor, packed_dtype), offsetof(jat_adamw_tensor, vec_ok),'
        ' offsetof(jat_adamw_tensor, bias_corr1), offsetof(jat_adamw_tensor, bias_corr2_sqrt));return 0;}\n')
    exe = tmp_path / "opt_layout"
    subprocess.run(["gcc", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [64, 0, 8, 16, 24, 32, 40, 48, 52, 56, 60]
    # the packing of word 7 as optim._Group.upload does it: low half = bias_corr1, high half = bias_corr2_sqrt
    bc = torch.tensor([[0.25, 0.5]], dtype=torch.float32)
    word = int(bc.view(torch.int64).reshape(-1)[0])
    lo = torch.tensor([word & 0xffffffff], dtype=torch.int64).to(torch.int32).view(torch.float32)
    hi = torch.tensor([word >> 32], dtype=torch.int64).to(torch.int32).view(torch.float32)
    assert float(lo) == 0.25 and float(hi) == 0.5


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly(lib):
    out = ctypes.c_void_p()
    rc = lib.load().jat_create(0, ctypes.byref(out))
    assert rc != 0 and not out.value
    assert len(lib.load().jat_last_error()) > 0
    import jat_b200
    m = jat_b200.JaT_AudioSR_V2(input_channels=32, cond_channels=32, hidden_size=128, depth=1, num_q_heads=2,
                                num_kv_heads=1, bottleneck_dim=128).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 32, 8), torch.zeros(1), torch.zeros(1, 32, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        jat_b200.flow_matching_sample(m, torch.zeros(1, 32, 8), device="cpu", verbose=False)


@pytest.mark.skipif(not have_reference(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("cls_name", ["JaT_AudioSR_V2", "JaT_AudioSR_V3"])
def test_state_dict_is_interchangeable_with_reference(cls_name):
    """Same keys, shapes, dtypes and -- under the same seed -- the same initial values as the reference module;
    strict load works in both directions (train_ddp_v3mod2.py:767 resumes with strict=True)."""
    import contextlib
    import io
    import jat_b200
    V2, V3, _ = import_reference()
    ref_cls = V2 if cls_name.endswith("V2") else V3
    cfg = dict(input_channels=32, cond_channels=32, patch_len=4, hidden_size=128, depth=3, num_q_heads=4, num_kv_heads=2,
               bottleneck_dim=64, mlp_ratio=2.0, dropout=0.1, drop_path_rate=0.05)
    torch.manual_seed(7)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = ref_cls(**cfg)
    torch.manual_seed(7)
    ours = getattr(jat_b200, cls_name)(**cfg)
    a, b = ref.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k
        assert torch.equal(a[k], b[k]), k
    ours.load_state_dict(a, strict=True)
    ref.load_state_dict(b, strict=True)
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in ours.named_parameters()]


def test_fused_adamw_is_a_torch_adamw_and_has_no_cpu_path(lib):
    """Same constructor / param_groups / state_dict layout as torch.optim.AdamW (train_ddp_v3mod2.py:709); stepping CPU
    parameters raises instead of falling back."""
    import jat_b200
    m = torch.nn.Linear(4, 3)
    stock = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=0.1)
    m(torch.randn(2, 4)).sum().backward()
    stock.step()
    opt = jat_b200.FusedAdamW(m.parameters(), lr=1e-3, weight_decay=0.1, max_grad_norm=1.0, model=None)
    assert isinstance(opt, torch.optim.AdamW)
    opt.load_state_dict(stock.state_dict())
    sd = opt.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 1.0
    assert set(sd["param_groups"][0]) == set(stock.state_dict()["param_groups"][0])
    with pytest.raises(lib.JatError):
        opt.step()
