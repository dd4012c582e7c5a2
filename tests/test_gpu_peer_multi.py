"""The peer-memory step on REAL ranks (one process per GPU, symmetric memory over NVLink): scripts/check_peer_step.py under
torchrun.  Needs two GPUs; on a one-GPU box the emulated-fabric test in test_gpu_scan.py covers the same kernels."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_memory_step_on_two_gpus():
    """Collective step, step_peer, graph-replayed advance() and step_peer under stream skew agree bit for bit; n', u' of the
    first step equal the undivided solver's; every field of the classical rollout certified."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "check_peer_step.py")]
    run = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-4000:]
    rec = json.loads([ln for ln in run.stdout.splitlines() if ln.startswith("{")][-1])
    for kind in ("baseline", "hybrid_fp16x3"):
        assert rec[kind]["bit_identical"] is True
    assert all(rec["baseline"]["all_fields_certified"].values())
