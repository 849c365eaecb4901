"""Re-export of oracle/torch_dit.py (the differentiable torch restatement of the reference forward / sampler) under the
name the parity tests have always imported."""
from oracle.torch_dit import *  # noqa: F401,F403
from oracle.torch_dit import _norm, _rope, dit_forward, flow_matching_sample  # noqa: F401
