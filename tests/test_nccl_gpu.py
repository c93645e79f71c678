"""Multi-GPU parity on real hardware: torchrun with one rank per GPU over NCCL (tests/nccl_parity_worker.py).
Skipped on boxes with fewer than two GPUs; the log of a 2-GPU run is committed under profiles/."""
import json
import os
import subprocess
import sys

import pytest
import torch

from parity_common import ROOT, ref_available

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not ref_available(), reason="reference build (oracle/_ref) not present")
def test_reduced_gradients_equal_sum_of_reference_single_view_gradients():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ)
    env.setdefault("NCCL_DEBUG", "WARN")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(ROOT, "tests", "nccl_parity_worker.py")],
                       capture_output=True, text=True, timeout=900, env=env)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, (r.returncode, r.stdout[-2000:], r.stderr[-4000:])
    rep = json.loads(lines[-1])
    print(rep)
    assert rep["ok"] and rep["world"] == world
