"""GPU suite, multi-GPU part: `dist.ShardedCRFLoss` over NCCL (BASELINE configs[4]: 256 frames sharded over the GPUs
of one box, scalar loss all-reduce).  Needs >= 2 GPUs; skipped otherwise (the world-size-2 gloo tests in
tests/test_dist_gloo.py cover the host logic on CPU)."""
import json
import os
import re
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("n_total", [256, 37])
def test_sharded_crf_loss_over_nccl(n_total):
    import torch
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_nccl_worker.py"), str(n_total)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    # the ranks share one stdout: two lines may arrive glued together
    dec = json.JSONDecoder()
    lines = [dec.raw_decode(res.stdout, m.end())[0] for m in re.finditer(r"NCCL_WORKER ", res.stdout)]
    assert len(lines) == world
    assert sorted((d["lo"], d["hi"]) for d in lines)[0][0] == 0 and max(d["hi"] for d in lines) == n_total
    for d in lines:
        for reduction in ("global", "global_async", "local"):
            if reduction == "local" and n_total % world:
                continue
            assert d[reduction]["loss_rel"] < 1e-5 and d[reduction]["grad_rel"] < 1e-5, d
