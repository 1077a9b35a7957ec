"""Time the attention kernels alone on C2-shaped inputs (N sequences x S x 12 heads x 64) for dense / causal / padded /
packed masks, with and without attention dropout.   Usage: python tools/attn_bench.py [tag] [kinds] [--once]
(--once: a single fwd + bwd per case, for running under ncu).  Output -> gpurun_out/attn_bench_<tag>.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from graphgpt_b200 import ops, synth  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    tag = args[0] if args else "x"
    kinds = args[1].split(",") if len(args) > 1 else ["dense", "causal", "padded", "packed"]
    once = "--once" in sys.argv
    dev = "cuda"
    N, S, H = 64, 1024, 12
    d = H * 64
    torch.manual_seed(0)
    qkv = (torch.randn(N * S, 3 * d, device=dev) * 0.7).to(torch.bfloat16)
    dout = (torch.randn(N * S, d, device=dev) * 0.1).to(torch.bfloat16)
    pos = torch.arange(S, device=dev, dtype=torch.int32).repeat(N)
    inv = 1.0 / (10000.0 ** (torch.arange(0, 64, 2).float() / 64))
    fr = torch.arange(S).float()[:, None] * inv[None]
    cos, sin = fr.cos().to(dev).contiguous(), fr.sin().to(dev).contiguous()
    lines = []
    for kind in kinds:
        causal = kind == "causal"
        if kind in ("dense", "causal"):
            am = None
            vis = S if kind == "dense" else (S + 1) / 2
        elif kind == "padded":
            g = torch.Generator().manual_seed(1)
            lens = torch.randint(S // 4, S + 1, (N,), generator=g)
            am = (torch.arange(S)[None, :] < lens[:, None]).long().to(dev)
            vis = float((lens.float() ** 2).sum() / lens.sum())
        else:
            b = synth.make_batch(N, S, layout="packed", seed=3)
            am = torch.from_numpy(b["attention_mask"]).to(dev)
            vis = float(am.sum() / max(1, (am.sum(-1) > 0).sum()))
        mask = ops.attn_mask_build(am, N, S, causal, dev)
        ntok = N * S if am is None else int((am.sum(-1) > 0).sum()) if am.dim() == 3 else int(am.sum())
        flops_f = 4.0 * ntok * vis * d
        for p_drop in (0.0, 0.1):
            def fwd():
                return ops.attn_fwd(qkv, mask, H, want_lo=True, dropout_p=p_drop, seed=7)

            out, lse, out_lo = fwd()

            def bwd():
                return ops.attn_bwd(dout, qkv, out, lse, mask, H, pos, cos, sin, out_lo=out_lo, dropout_p=p_drop, seed=7)

            bwd()
            torch.cuda.synchronize()
            if once:
                continue
            res = []
            for fn in (fwd, bwd):
                for _ in range(2):
                    fn()
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(10):
                    fn()
                b_.record()
                torch.cuda.synchronize()
                res.append(a.elapsed_time(b_) / 10)
            lines.append(f"{kind:7s} p={p_drop:.1f} vis={vis:7.1f}  fwd {res[0]:7.3f} ms {flops_f / res[0] / 1e9:7.1f} TF/s   "
                         f"bwd {res[1]:7.3f} ms {2.5 * flops_f / res[1] / 1e9:7.1f} TF/s")
            print(lines[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    if not once:
        with open(os.path.join(ROOT, "gpurun_out", f"attn_bench_{tag}.txt"), "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
