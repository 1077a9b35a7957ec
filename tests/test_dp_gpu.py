"""Data-parallel numerics on real GPUs (needs >= 2): `torchrun --nproc-per-node 2 tools/dp_check.py` — every rank takes one
step on its own batch through GraphGPTEngine (per-segment NCCL all-reduce overlapped with backward); the all-reduced
gradient must equal the single-process sum of the per-rank gradients and the updated parameters the single-process
accumulated step, for the bf16 wire format (default, as DeepSpeed bf16) and for fp32."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on the box")]


@pytest.mark.parametrize("wire", ["bf16", "fp32"])
def test_dp2_gradients_equal_the_single_rank_sum(wire):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611" if wire == "bf16" else "29612", os.path.join(ROOT, "tools", "dp_check.py")]
    if wire == "fp32":
        cmd.append("--fp32-wire")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    out = p.stdout + p.stderr
    rep = os.path.join(ROOT, "gpurun_out", f"dp_check_{wire}.txt")
    os.makedirs(os.path.dirname(rep), exist_ok=True)
    with open(rep, "w") as f:
        f.write("\n".join(l for l in p.stdout.splitlines() if l.startswith("dp_check")) + "\n")
    assert p.returncode == 0 and "dp_check OK" in p.stdout, out[-3000:]
