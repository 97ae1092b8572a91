"""GPU: the data-parallel training path on hardware — see tests/dp_worker.py for the checks.
Two ranks: NCCL over two GPUs when the box has them, gloo over one shared GPU otherwise."""
import json
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_rank_gradient_exchange_and_parameter_sync():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dp_worker.py")]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, cwd=ROOT)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("DP_RESULT ")]
    assert p.returncode == 0 and lines, p.stdout[-4000:]
    res = json.loads(lines[-1][len("DP_RESULT "):])
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, f"dp_parity_{res['backend']}.json"), "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    assert res["ok"], res
