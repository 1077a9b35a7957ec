"""Host-side enqueue time of one training step vs its device time (is the launch thread the bottleneck?).
usage: python tools/host_overhead.py  -> prints ms of host enqueue per step, device ms per step, with / without GEMM events"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from graphgpt_b200 import GraphGPTConfig, GraphGPTPretrainBase, synth  # noqa: E402
from graphgpt_b200.dp import GraphGPTEngine  # noqa: E402
from graphgpt_b200.lib import GEMM_FUNCS, KernelTimer, lib  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = GraphGPTPretrainBase(GraphGPTConfig(**bench.MODEL)).to(dev).train()
engine = GraphGPTEngine(model, lr=3e-4)
b = synth.make_batch(64, 1024, layout="packed", seed=1)
batch = {k: torch.from_numpy(b[k]).to(dev) for k in ("input_ids", "attention_mask", "labels")}


def step():
    out = engine(**batch)
    engine.backward(out.head1_loss)
    engine.step()


for _ in range(3):
    step()
for label, timer in (("no events", None), ("GEMM events", "gemm"), ("all events", "all"), ("no events", None)):
    lib.timer = None if timer is None else KernelTimer(only=GEMM_FUNCS if timer == "gemm" else None)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    t0 = time.perf_counter()
    for _ in range(8):
        step()
    host = (time.perf_counter() - t0) / 8 * 1e3
    e.record()
    torch.cuda.synchronize()
    print(f"{label:12s}: host enqueue {host:6.2f} ms/step, device {a.elapsed_time(e) / 8:6.2f} ms/step")
    lib.timer = None
