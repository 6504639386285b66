"""GPU, N > 1: the sharded device run (tests/mgpu_device_run.py) under torchrun on 2 GPUs of the box.  Skipped on a
single-GPU box; the host-side rules it relies on are covered on CPU by tests/test_sharding.py (gloo, world size 2)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_device_run_on_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "mgpu_device_run.py"), "--steps",
                        "300"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads(line[-1])
    assert out["ok"] and all(out["checks"].values()), out
