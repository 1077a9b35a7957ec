#!/usr/bin/env python
"""bench.py — tokens/sec of the GraphGPT SMTP pre-training step on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (CPU arm: the oracle port on the host cores)

Workload (config.workload): PCQM4M-v2 SMTP pre-training, GraphGPT-B12 = 12L / 768d / 12 heads x 64 / GeGLU 3072,
F=13 stacked tokens, V=756, sequences of 1024 tokens; synthetic PCQM4M-shaped Eulerian-path samples (mean 23.3 rows)
packed to 1024 with a block-diagonal attention mask, per-sequence SMTP masking (see graph-gpt_b200/synth.py).
A step = forward + backward + gradient all-reduce + AdamW over one batch of `--seqs` x 1024 tokens per GPU (weak
scaling).  `value` = non-pad tokens of all ranks / max-over-ranks device time with inputs resident in HBM;
`e2e` = same step driven from pinned host batches (H2D copies of ids / labels / mask inside the timed region, loss
read back every step).  Distinct batches (> L2) are cycled, so no flush is needed between iterations.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = dict(vocab_size=756, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
             num_key_value_heads=12, head_dim=64, hidden_act="gelu", max_position_embeddings=1024, rms_norm_eps=1e-6,
             rope_theta=10000.0, pad_token_id=0, bos_token_id=20, eos_token_id=19, causal_attention=False,
             stacked_feat=13, stack_method="short", stacked_feat_agg_method="sum", next_n_token=13, use_cache=False,
             attention_dropout=0.1)   # examples/graph_lvl/pcqm4m_v2_pretrain.sh:20 — active in the training step
SEQ = 1024
METRIC = "tokens/sec (device-timed) PCQM4M-v2 SMTP pretrain"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seqs", type=int, default=64, help="sequences of 1024 tokens per GPU per step")
    ap.add_argument("--layout", default="packed", choices=["packed", "dense"])
    ap.add_argument("--batches", type=int, default=4, help="distinct synthetic batches cycled through")
    ap.add_argument("--mask-format", default="reference", choices=["reference", "segments"],
                    help="packed layout only: 'reference' = the collator's int64 [N,S,S] block-diagonal mask (default, the "
                         "reference-facing contract); 'segments' = the [N,S] segment-id mask of graphgpt_b200.packing")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--fwd-only", action="store_true", help="also report the forward-only pass (inference mode)")
    return ap.parse_args()


def flops_per_token(layout, batch_stats):
    """Forward model FLOPs per token: L*(8d^2 + 6dI + 4*S_vis*d) + head."""
    d, I, L, F, V = 768, 3072, 12, 13, 756
    s_vis = batch_stats["s_vis"]
    backbone = L * (8 * d * d + 6 * d * I + 4 * s_vis * d)
    head = batch_stats["row_frac"] * 2 * d * d * F + batch_stats["entry_per_tok"] * 2 * d * V
    return backbone + head


def batch_stats(b, layout):
    import numpy as np
    lab = b["labels"]
    N, S = lab.shape[:2]
    mask = lab != -100
    st = {"row_frac": float(mask.any(-1).mean()), "entry_per_tok": float(mask.sum() / (N * S))}
    if layout == "packed":
        segs = b["segment_lens"]
        st["s_vis"] = float(sum(l * l for row in segs for l in row) / (N * S))
    else:
        st["s_vis"] = float(S)
    return st


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port timed on the host cores (reported baseline; the reference is pure Python/torch and
# cannot travel to the GPU box, so kind = "port")
# ------------------------------------------------------------------------------------------------------------------
def cpu_step_tokens_per_s(n_seq, steps, warmup, layout):
    import torch

    from graphgpt_b200 import synth
    from oracle import graphgpt_oracle as oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b = synth.make_batch(n_seq, SEQ, layout=layout, seed=1234)
    ids, am, labels = (torch.from_numpy(b[k]) for k in ("input_ids", "attention_mask", "labels"))
    sd = {k: v.requires_grad_(True) for k, v in oracle.init_state_dict(MODEL, seed=0).items()}
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = oracle.pretrain_forward(sd, MODEL, ids, am, labels)
        out["loss"].backward()
        with torch.no_grad():                      # plain SGD-free timing: AdamW cost on CPU is negligible vs fwd/bwd
            for v in sd.values():
                v.grad = None
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    tok = n_seq * SEQ
    ms = statistics.median(times) * 1e3
    return tok / (ms / 1e3), ms, cores, f"{n_seq} x {SEQ} tokens ({layout}), fwd+bwd, fp32, {cores} threads, median of {steps}"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_seq = 1
    steps = max(1, min(args.steps, 3))
    warm = 1
    tps, ms, cores, sample = cpu_step_tokens_per_s(n_seq, steps, warm, args.layout)
    line = {"impl": "reference", "metric": METRIC, "value": tps, "unit": "tokens/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "seq_len": SEQ, "tokens_per_step": n_seq * SEQ,
                       "note": "CPU oracle port of the reference path (oracle/graphgpt_oracle.py), bounded sample"},
            "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_name(args):
    suffix = "-segment-id-mask" if (args.layout == "packed" and getattr(args, "mask_format", "reference") == "segments") else ""
    return f"pcqm4m-v2-smtp-pretrain-B12(12L/768d/F13/V756)-seq1024-{args.layout}{suffix}"


# ------------------------------------------------------------------------------------------------------------------
def clocks_sampler_start(path):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        return subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                stdout=open(path, "w"), stderr=subprocess.DEVNULL)
    except Exception:
        return None


def clocks_summary(path, dev_index):
    sm, mx, reasons = [], 0.0, set()
    try:
        for ln in open(path):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9 or f[0] != str(dev_index):
                continue
            sm.append(float(f[1]))
            mx = max(mx, float(f[2]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
    except Exception:
        pass
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    busy = sorted(sm)[len(sm) // 2:]          # upper half = samples under load
    return {"sm_mhz": statistics.median(busy), "sm_max_mhz": mx, "reasons": sorted(reasons)}


def run_b200(args):
    import torch
    import torch.distributed as dist

    from graphgpt_b200 import GraphGPTConfig, GraphGPTPretrainBase, synth
    from graphgpt_b200.dp import GraphGPTEngine
    from graphgpt_b200.lib import GEMM_FUNCS, KernelTimer, lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()

    # ---- synthetic batches (host, pinned) -------------------------------------------------------------------
    host, stats = [], []
    for i in range(args.batches):
        b = synth.make_batch(args.seqs, SEQ, layout=args.layout, seed=1234 + 1000 * rank + i, return_segments=True)
        stats.append(batch_stats(b, args.layout))
        if args.layout == "packed" and args.mask_format == "segments":
            import numpy as np
            seg = np.zeros((args.seqs, SEQ), np.int64)
            for n_, lens in enumerate(b["segment_lens"]):
                seg[n_] = np.repeat(np.arange(1, len(lens) + 1), lens)[:SEQ]
            b["attention_mask"] = seg
        host.append({k: torch.from_numpy(b[k]).pin_memory() for k in ("input_ids", "attention_mask", "labels")})
    tok_per_step = args.seqs * SEQ                                   # packed / dense layouts have no pad tokens
    st = {k: sum(s[k] for s in stats) / len(stats) for k in stats[0]}
    fpt = flops_per_token(args.layout, st)
    devb = [{k: v.to(dev, non_blocking=True) for k, v in hb.items()} for hb in host]

    torch.manual_seed(0)
    model = GraphGPTPretrainBase(GraphGPTConfig(**MODEL)).to(dev).train()
    engine = GraphGPTEngine(model, lr=3e-4, betas=(0.9, 0.95), eps=1e-6, weight_decay=0.1, max_grad_norm=1.0)

    def step(batch):
        out = engine(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"], labels=batch["labels"])
        engine.backward(out.head1_loss)
        engine.step()
        return out.head1_loss

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, n):
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        marks = []
        for i in range(n):
            fn(i)
            if os.environ.get("GGPT_BENCH_DEBUG"):
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append(ev)
        b.record()
        sync_all()
        if marks:
            prev, per = a, []
            for ev in marks:
                per.append(round(prev.elapsed_time(ev), 2))
                prev = ev
            sys.stderr.write(f"[bench debug] per-step device ms: {per}\n")
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- warm-up, then the timed region (inputs resident in HBM) -------------------------------------------------
    for i in range(args.warmup):
        step(devb[i % len(devb)])
    # Everything built so far (model, batches, numpy scaffolding of the synthetic data) is long-lived: move it out of the
    # garbage collector's reach, otherwise a full collection walks it in the middle of a step and stalls the launch
    # thread for tens of milliseconds (the queue runs only ~5 ms ahead of the device).
    engine.freeze_gc()
    clk_path = os.path.join(tempfile.gettempdir(), f"ggpt_clocks_{os.getpid()}.csv")
    sampler = clocks_sampler_start(clk_path) if (rank == 0 and not os.environ.get("GGPT_BENCH_NO_SAMPLER")) else None
    launches0 = lib.launch_count
    # inside the timed region only the dominant kernel (the tcgen05 GEMM, every launch of it) is bracketed by CUDA
    # events: those durations give roofline.achieved.  The other entry points are timed in a separate pass below.
    lib.timer = KernelTimer(only=GEMM_FUNCS)
    total_ms = timed(lambda i: step(devb[i % len(devb)]), args.steps)
    gtimes = lib.timer.summary()
    lib.timer = None
    launches = lib.launch_count - launches0
    if sampler is not None:
        sampler.terminate()
    ms_per_step = total_ms / args.steps
    value = world * tok_per_step / (ms_per_step / 1e3)
    # per-entry-point breakdown (informational `kernel_ms_per_step`): every call bracketed, outside the timed region
    n_prof = min(args.steps, 4)
    lib.timer = KernelTimer()
    timed(lambda i: step(devb[i % len(devb)]), n_prof)
    ktimes = lib.timer.summary()
    lib.timer = None

    # ---- end-to-end: pinned host batches -> H2D -> step -> loss.item() ---------------------------------------------
    e2e = None
    if not args.no_e2e:
        # what a training loop does: pinned host batch -> non-blocking H2D on a copy stream (the next batch is in
        # flight while the current step computes) -> forward/backward/step -> loss read back every step
        copy_stream = torch.cuda.Stream()

        def prefetch(i):
            hb = host[i % len(host)]
            with torch.cuda.stream(copy_stream):
                batch = {k: v.to(dev, non_blocking=True) for k, v in hb.items()}
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return batch, ev

        state = {"next": prefetch(0)}

        def e2e_step(i):
            batch, ev = state["next"]
            state["next"] = prefetch(i + 1)
            torch.cuda.current_stream().wait_event(ev)
            for t in batch.values():
                t.record_stream(torch.cuda.current_stream())
            return step(batch).item()
        e2e_step(0)
        n_e2e = max(3, args.steps // 2)
        ms = timed(lambda i: e2e_step(i + 1), n_e2e) / n_e2e
        h2d = sum(v.numel() * v.element_size() for v in host[0].values())
        e2e = {"value": world * tok_per_step / (ms / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 4, "ms_per_step": ms}

    # ---- forward-only (the north-star "fused attention+MLP forward" figure) ---------------------------------------------
    fwd = None
    if args.fwd_only:
        model.eval()
        with torch.no_grad():
            def fstep(i):
                b = devb[i % len(devb)]
                return model(input_ids=b["input_ids"], attention_mask=b["attention_mask"], labels=b["labels"]).head1_loss
            for i in range(2):
                fstep(i)
            ms = timed(fstep, args.steps) / args.steps
        model.train()
        tf = fpt * tok_per_step / (ms / 1e3) / 1e12
        fwd = {"ms_per_step": ms, "tokens_per_s_per_gpu": tok_per_step / (ms / 1e3), "model_tflops_per_gpu": tf}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback (B200_PROFILING.md ~1.4 PF sustained)"
    # dominant kernel = the tcgen05 GEMM (all instantiations of gemm_kernel<>): FLOPs of every launch / summed time
    gemm = [(k, v) for k, v in gtimes.items() if k.startswith("ggpt_gemm")]
    g_ms = sum(v["ms"] for _, v in gemm)
    g_fl = sum(v["flops"] for _, v in gemm)
    all_ms = total_ms
    top = max(gemm, key=lambda kv: kv[1]["ms"]) if gemm else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_gemm_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "ggpt::gemm_kernel<> (tcgen05 GEMM, all instantiations)",
                "achieved": g_fl / (g_ms / 1e3) / 1e12 if g_ms > 0 else None, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": (g_fl / (g_ms / 1e3) / 1e12) / peak_tf if g_ms > 0 else None, "traffic": traffic,
                "peak_source": peak_src, "share_of_step_time": g_ms / all_ms if all_ms > 0 else None,
                "launches": sum(v["calls"] for _, v in gemm),
                "top_instance": None if top is None else {
                    "name": top[0], "calls": top[1]["calls"], "ms": top[1]["ms"],
                    "tflops": top[1]["flops"] / (top[1]["ms"] / 1e3) / 1e12},
                # per entry point: the plain GEMMs (forward / dgrad / wgrad majors) next to the fused-epilogue variants,
                # whose launch time also contains the HBM-bound epilogue work folded into them (GeGLU factors, GeGLU
                # backward, RoPE) and is charged to the GEMM FLOPs only
                "instances": [{"name": k, "calls": v["calls"], "ms_per_step": v["ms"] / args.steps,
                               "tflops": v["flops"] / (v["ms"] / 1e3) / 1e12,
                               "frac": v["flops"] / (v["ms"] / 1e3) / 1e12 / peak_tf}
                              for k, v in sorted(gemm, key=lambda kv: -kv[1]["ms"])]}
    step_tf = 3 * fpt * tok_per_step / (ms_per_step / 1e3) / 1e12
    line = {"metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(args), "seq_len": SEQ, "seqs_per_gpu": args.seqs,
                       "tokens_per_step_per_gpu": tok_per_step, "parallelism": f"dp{world}",
                       "l2_policy": f"{args.batches} distinct batches cycled; activations per step >> 126 MB L2",
                       "mean_visible_keys": st["s_vis"], "fwd_mflop_per_token": fpt / 1e6},
            "model_tflops_per_gpu": step_tf, "model_tflops_frac_of_peak": step_tf / peak_tf,
            "roofline": roofline, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks_summary(clk_path, local),
            "kernel_ms_per_step": {k: round(v["ms"] / n_prof, 4) for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1]["ms"])}}
    if fwd is not None:
        fwd["frac_of_peak"] = fwd["model_tflops_per_gpu"] / peak_tf
        line["forward_only"] = fwd
    if not args.no_cpu_baseline and world == 1:
        tps, ms, cores, sample = cpu_step_tokens_per_s(1, 2, 1, args.layout)
        line["cpu_baseline"] = {"value": tps, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    # stdout must carry exactly ONE JSON line (rank 0): anything libraries print there (e.g. NCCL's version banner)
    # is diverted to stderr, and the result line is written to the original stdout at the end.
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    _buf = []
    _print = print

    def print(*args, **kw):  # noqa: A001 - the only print() calls below emit the result line
        _buf.append(" ".join(str(x) for x in args))

    import builtins
    builtins.print = print
    try:
        if a.impl == "reference":
            run_reference(a)
        else:
            run_b200(a)
    finally:
        builtins.print = _print
        sys.stdout.flush()
        os.dup2(_real_stdout, 1)
        for line in _buf:
            os.write(1, (line + "\n").encode())
