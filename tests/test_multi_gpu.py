"""N > 1 on real devices (NCCL, one process per GPU, launched like the bench contract launches
bench.py): batch sharding of the op, query-range sharded inference and the DDP step against the
single-GPU results.  Needs >= 2 GPUs in the box (`gpurun --gpus 2`); the CPU/gloo twins of these
checks are tests/test_sharding.py."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(900)
def test_two_rank_nccl_sharding_and_ddp_match_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multi", "worker_nccl.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=850, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == 2 and out["query_sharding_max_abs_diff_vs_unsharded"] <= 2e-5
