#!/usr/bin/env python
"""Kernel timeline of the data-parallel training step (there is no nsys in the image: CUPTI through torch.profiler).

  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_trace.py <tag> [--no-overlap]
  python tools/dp_trace.py <tag>                      (single GPU, same table — the comparison row)

Writes gpurun_out/<tag>_trace_rank0.csv.gz (name, stream, start_us, dur_us of every kernel / memcpy of ONE timed step)
and prints: step span, busy time of the compute stream, per-kernel-family time, the NCCL kernels with what ran beside
them, and the idle gaps of the compute stream > 20 us."""
import gzip
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from torch.profiler import ProfilerActivity, profile

    import bench
    from graphgpt_b200 import GraphGPTConfig, GraphGPTPretrainBase
    from graphgpt_b200.dp import GraphGPTEngine
    from graphgpt_b200.lib import lib

    tag = sys.argv[1] if len(sys.argv) > 1 else "dp"
    overlap = "--no-overlap" not in sys.argv
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()
    c = bench.CONFIGS["c2"]
    devb = []
    for i in range(2):
        b = bench.make_host_batch("c2", "packed", 64, 1234 + 1000 * rank + i)
        devb.append({k: torch.from_numpy(v).to(dev) for k, v in bench.model_inputs("c2", b).items()})
    torch.manual_seed(0)
    model = GraphGPTPretrainBase(GraphGPTConfig(**c["model"])).to(dev).train()
    engine = GraphGPTEngine(model, max_grad_norm=1.0, overlap_comm=overlap, **c["opt"])

    def step(b):
        out = engine(**b)
        engine.backward(out.head1_loss)
        engine.step()

    for i in range(4):
        step(devb[i % 2])
    engine.freeze_gc()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(3):
            step(devb[i % 2])
        torch.cuda.synchronize()
    if rank != 0:
        if world > 1:
            dist.barrier()
        return
    evs = []
    for e in prof.events():
        if e.device_type is not None and str(e.device_type).endswith("CUDA"):
            tr = e.time_range
            evs.append((e.name, getattr(e, "device_index", 0), tr.start, tr.end - tr.start))
    evs.sort(key=lambda x: x[2])

    def short(n):
        n = re.sub(r"^void ", "", n)
        n = re.sub(r"\(.*$", "", n)
        return n.replace("ggpt::", "")

    starts = [i for i, e in enumerate(evs) if short(e[0]).startswith("embed_fwd_kernel")]
    a, b = starts[1], starts[2]
    one = evs[a:b]
    t0 = one[0][2]
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with gzip.open(os.path.join(out, f"{tag}_trace_rank0.csv.gz"), "wt") as f:
        f.write("name,start_us,dur_us\n")
        for n, _, s, d in one:
            f.write(f"\"{short(n)}\",{s - t0:.1f},{d:.1f}\n")
    span = evs[b][2] - t0
    fam = {}
    for n, _, s, d in one:
        k = short(n)
        k = "nccl" if "nccl" in k.lower() else re.sub(r"<.*", "", k)
        fam[k] = fam.get(k, 0.0) + d
    print(f"world {world} overlap {overlap}: one step = {span / 1e3:.3f} ms, {len(one)} device activities")
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1])[:24]:
        print(f"  {k:40s} {v / 1e3:8.3f} ms")
    comp = [(s - t0, s - t0 + d, short(n)) for n, _, s, d in one if "nccl" not in n.lower() and "Memcpy" not in n and "Memset" not in n]
    nccl = [(s - t0, s - t0 + d) for n, _, s, d in one if "nccl" in n.lower()]
    busy = sum(e - s for s, e, _ in comp)
    print(f"  compute-kernel time {busy / 1e3:.3f} ms; NCCL kernels {len(nccl)}: total {sum(e - s for s, e in nccl) / 1e3:.3f} ms")
    for s, e in nccl:
        beside = [(n, min(e, ce) - max(s, cs)) for cs, ce, n in comp if ce > s and cs < e]
        print(f"    nccl {s / 1e3:8.3f} .. {e / 1e3:8.3f} ms ({(e - s):7.1f} us) beside: " +
              ", ".join(f"{re.sub(r'<.*', '', n)}:{d:.0f}" for n, d in beside[:6]))
    prev_end, gaps = 0.0, []
    for s, e, n in comp:
        if s - prev_end > 20.0:
            gaps.append((prev_end, s - prev_end, n))
        prev_end = max(prev_end, e)
    print(f"  idle gaps > 20 us on the compute stream: {len(gaps)}, total {sum(g[1] for g in gaps) / 1e3:.3f} ms")
    for at, g, n in gaps[:40]:
        print(f"    at {at / 1e3:8.3f} ms: {g:7.1f} us before {n[:60]}")
    if world > 1:
        dist.barrier()


if __name__ == "__main__":
    main()
