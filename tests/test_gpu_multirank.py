"""GPU parity at N > 1: two ranks over NCCL give the single-rank log-likelihood, MCNR sums, mvn_ll (single and
batched) and a 3-iteration mcml_full.  Needs two GPUs (`gpurun --gpus 2`); skipped on a one-GPU box.  The same check at full size
runs inside bench.py whenever it is launched with more than one rank (`step_parity`, `configs.*.parity`)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_over_nccl_match_one_rank():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "nccl_parity_worker.py")], capture_output=True, text=True, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("NCCL_PARITY ")]
    assert line, (r.stdout[-2000:], r.stderr[-2000:])
    res = json.loads(line[-1][len("NCCL_PARITY "):])
    assert res["ok"] and r.returncode == 0, res
