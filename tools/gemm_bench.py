"""Time the tcgen05 GEMM on the shapes of one C2 training step (T = 65536 tokens, d 768, I 3072) in each launch mode
(GGPT_GEMM_MODE = single | multicast | pair, read once per process -> one subprocess per mode) and check each result
against torch.matmul on the same bf16 operands.    Usage: python tools/gemm_bench.py [tag]  -> gpurun_out/gemm_bench_<tag>.txt
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker():
    import torch
    from graphgpt_b200 import ops

    torch.manual_seed(0)
    dev = "cuda"
    T, d, I = 65536, 768, 3072
    BF16, F32 = torch.bfloat16, torch.float32
    x = (torch.randn(T, d, device=dev) * 0.5).to(BF16)
    xi = (torch.randn(T, I, device=dev) * 0.5).to(BF16)
    x3 = (torch.randn(T, 3 * d, device=dev) * 0.5).to(BF16)
    x6 = (torch.randn(T, 2 * I, device=dev) * 0.5).to(BF16)
    w_qkv = (torch.randn(3 * d, d, device=dev) * 0.05).to(BF16)
    w_o = (torch.randn(d, d, device=dev) * 0.05).to(BF16)
    w_gu = (torch.randn(2 * I, d, device=dev) * 0.05).to(BF16)
    w_d = (torch.randn(d, I, device=dev) * 0.05).to(BF16)
    pos = torch.arange(T, device=dev, dtype=torch.int32) % 1024
    inv = 1.0 / (10000.0 ** (torch.arange(0, 64, 2).float() / 64))
    fr = torch.arange(1024).float()[:, None] * inv[None]
    cos, sin = fr.cos().to(dev).contiguous(), fr.sin().to(dev).contiguous()
    g_qkv = torch.zeros(3 * d, d, device=dev, dtype=F32)
    g_o = torch.zeros(d, d, device=dev, dtype=F32)
    g_gu = torch.zeros(2 * I, d, device=dev, dtype=F32)
    g_d = torch.zeros(d, I, device=dev, dtype=F32)
    wg = dict(a_mn_major=True, b_mn_major=True, out_dtype=F32, accumulate=True)

    def relerr(a, b):
        return ((a.float() - b.float()).norm() / b.float().norm()).item()

    cases = {
        # name: (fn, flops, checker or None)
        "fwd o_proj   [T,768]x[768,768]": (lambda: ops.gemm(x, w_o), 2 * T * d * d, lambda y: relerr(y, x.float() @ w_o.float().t())),
        "fwd down     [T,3072]x[768,3072]": (lambda: ops.gemm(xi, w_d), 2 * T * d * I, lambda y: relerr(y, xi.float() @ w_d.float().t())),
        "fwd qkv+rope [T,768]x[2304,768]": (lambda: ops.gemm_qkv_rope(x, w_qkv, pos, cos, sin, 2 * d), 2 * T * 3 * d * d, None),
        "fwd geglu    [T,768]x[6144,768]": (lambda: ops.gemm_geglu(x, w_gu, want_gu=True)[0], 2 * T * 2 * I * d, None),
        "fwd geglu (inference epilogue)": (lambda: ops.gemm_geglu(x, w_gu, want_gu=False)[1], 2 * T * 2 * I * d, None),
        "dgrad down + GeGLU bwd (fused epilogue)": (lambda: ops.gemm_dgeglu(x, w_d, x6), 2 * T * d * I, None),
        "geglu_bwd stand-alone kernel": (lambda: ops.geglu_bwd(xi, x6), 0, None),
        "dgrad d->d   dy[T,768] W[768,768]": (lambda: ops.gemm(x, w_o, b_mn_major=True), 2 * T * d * d, lambda y: relerr(y, x.float() @ w_o.float())),
        "dgrad qkv    dy[T,2304] W[2304,768]": (lambda: ops.gemm(x3, w_qkv, b_mn_major=True), 2 * T * 3 * d * d, lambda y: relerr(y, x3.float() @ w_qkv.float())),
        "dgrad down   dy[T,768] W[768,3072]": (lambda: ops.gemm(x, w_d, b_mn_major=True), 2 * T * d * I, lambda y: relerr(y, x.float() @ w_d.float())),
        "dgrad gu     dy[T,6144] W[6144,768]": (lambda: ops.gemm(x6, w_gu, b_mn_major=True), 2 * T * 2 * I * d, lambda y: relerr(y, x6.float() @ w_gu.float())),
        "wgrad o      dy[T,768]^T x[T,768]": (lambda: ops.gemm(x, x, out=g_o, **wg), 2 * T * d * d, None),
        "wgrad qkv    dy[T,2304]^T x[T,768]": (lambda: ops.gemm(x3, x, out=g_qkv, **wg), 2 * T * 3 * d * d, None),
        "wgrad down   dy[T,768]^T a[T,3072]": (lambda: ops.gemm(x, xi, out=g_d, **wg), 2 * T * d * I, None),
        "wgrad gu     dy[T,6144]^T x[T,768]": (lambda: ops.gemm(x6, x, out=g_gu, **wg), 2 * T * 2 * I * d, None),
    }
    res = {}
    for name, (fn, flops, chk) in cases.items():
        y = fn()
        err = chk(y) if chk is not None else None
        del y
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        n = 20
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        res[name] = dict(ms=ms, tflops=flops / ms / 1e9, err=err)
    # wgrad accumulation check (one call into a zeroed buffer)
    g_o.zero_()
    ops.gemm(x, x, out=g_o, **wg)
    res["wgrad o      dy[T,768]^T x[T,768]"]["err"] = relerr(g_o, x.float().t() @ x.float())
    g_d.zero_()
    ops.gemm(x, xi, out=g_d, **wg)
    res["wgrad down   dy[T,768]^T a[T,3072]"]["err"] = relerr(g_d, x.float().t() @ xi.float())
    print("RESULT " + json.dumps(res))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "x"
    modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["multicast", "pair"]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = {}
    for mode in modes:
        env = dict(os.environ, GGPT_GEMM_MODE=mode)
        try:
            p = subprocess.run([sys.executable, __file__, "--worker"], capture_output=True, text=True, env=env, timeout=300)
        except subprocess.TimeoutExpired:
            out[mode] = "TIMEOUT"
            continue
        line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
        out[mode] = json.loads(line[0][7:]) if line else ("FAILED rc=%d\n" % p.returncode) + (p.stdout + p.stderr)[-3000:]
    with open(os.path.join(ROOT, "gpurun_out", f"gemm_bench_{tag}.txt"), "w") as f:
        names = next((list(v.keys()) for v in out.values() if isinstance(v, dict)), [])
        for mode, v in out.items():
            if not isinstance(v, dict):
                f.write(f"{mode}: {v}\n")
        f.write(f"{'case':44s}" + "".join(f"{m:>30s}" for m in modes) + "\n")
        for nme in names:
            row = f"{nme:44s}"
            for m in modes:
                v = out[m]
                if isinstance(v, dict):
                    r = v[nme]
                    e = "   -   " if r["err"] is None else f"{r['err']:.1e}"
                    row += f"  {r['ms']:7.3f} ms {r['tflops']:7.1f} TF {e}"
                else:
                    row += " " * 30
            f.write(row + "\n")
    print(open(os.path.join(ROOT, "gpurun_out", f"gemm_bench_{tag}.txt")).read())


if __name__ == "__main__":
    if "--worker" in sys.argv:
        worker()
    else:
        main()
