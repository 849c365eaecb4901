"""DDP correctness of the drop-in module (SURVEY.md 8a row a14, 8e training row; reference train_ddp_v3mod2.py:816-822,
922-929): two ranks launched with torch.distributed.run wrap the module like the reference does (`torch.compile` then
`DistributedDataParallel(..., find_unused_parameters=False)`); see tests/_ddp_worker.py for what each rank asserts
(start-up broadcast reaches the packed weights, all-reduced gradient == single-process gradient on the concatenated batch,
bit-identical parameters and eval outputs on all ranks after 3 optimizer steps).

One rank per GPU over NCCL when the box has two GPUs; on a single-GPU box both ranks share cuda:0 and DDP runs over gloo
(same DDP code path: buckets, hooks, gradient views -- only the transport differs)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("mode,cls", [("reference", "JaT_AudioSR_V2"), ("view", "JaT_AudioSR_V2"), ("view", "JaT_AudioSR_V3"),
                                      ("view_bf16", "JaT_AudioSR_V2"),
                                      ("view_bf16_fused", "JaT_AudioSR_V3")])
def test_two_rank_ddp_matches_single_process(mode, cls):
    env = dict(os.environ, OMP_NUM_THREADS="4")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_ddp_worker.py"), mode, cls]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    tail = (r.stdout[-3000:] + "\n" + r.stderr[-3000:])
    assert r.returncode == 0, tail
    assert "DDP_WORKER_OK" in r.stdout, tail
    print([l for l in r.stdout.splitlines() if l.startswith("DDP_WORKER_OK")][0])
