"""jat_b200 -- B200-native (sm_100a) implementation of the JaT-AudioSR DiT denoiser hot path.

Import name: ``jat_b200`` (the on-disk directory carries the reference's full name and is aliased by
the top-level ``jat_b200.py`` shim, because hyphens are not importable).

Public surface (mirrors the reference's):
    JaT_AudioSR_V2, JaT_AudioSR_V3     -- nn.Module drop-ins (src/models/jat_audiosr_v2.py / _v3.py)
    flow_matching_sample               -- Euler/CFG sampler drop-in (infer_test_v3m2.py:108)
    chunked.sample_long / crossfade_chunks -- long-audio chunk loop + crossfade (infer_test_v3m2.py:340-406, 188-233)
    FusedAdamW                         -- torch.optim.AdamW whose step() = clip_grad_norm_ + AdamW + bf16 re-pack in two
                                          multi-tensor passes (train_ddp_v3mod2.py:709, 926-928)
    ddp.register_bf16_allreduce        -- optional bf16 payload for DDP's bucketed gradient all-reduce (train_ddp_v3mod2.py:822)
    load_model / save_checkpoint / checkpoint.resume -- the reference's checkpoint format (infer_test_v3m2.py:33-94,
                                          train_ddp_v3mod2.py:1120-1148, 752-810)
"""
from . import _lib  # noqa: F401
from .models import JaT_AudioSR_V2, JaT_AudioSR_V3  # noqa: F401
from .sampler import flow_matching_sample  # noqa: F401
from . import chunked  # noqa: F401
from . import checkpoint  # noqa: F401
from . import training  # noqa: F401
from . import ddp  # noqa: F401
from .optim import FusedAdamW  # noqa: F401
from .checkpoint import load_model, save_checkpoint  # noqa: F401

__all__ = ["JaT_AudioSR_V2", "JaT_AudioSR_V3", "flow_matching_sample", "FusedAdamW", "chunked", "checkpoint", "training", "ddp", "load_model",
           "save_checkpoint", "_lib"]
__version__ = "0.1.0"
