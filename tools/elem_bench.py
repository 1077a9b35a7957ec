"""Time the HBM-bound kernels of one C2 layer (T = 65536 tokens, d 768, I 3072) against their algorithmic bytes.
Usage: python tools/elem_bench.py [tag]   -> gpurun_out/elem_bench_<tag>.txt   (GGPT_RMSNORM_BWD_NO_BULK=1 for the register version)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from graphgpt_b200 import ops  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "x"
    dev = "cuda"
    T, d, I = 65536, 768, 3072
    torch.manual_seed(0)
    BF16 = torch.bfloat16
    # several distinct buffers so consecutive calls do not hit in the 126 MB L2
    xs = [torch.randn(T, d, device=dev) for _ in range(3)]
    ys = [torch.randn(T, d, device=dev).to(BF16) for _ in range(3)]
    rs = [torch.randn(T, d, device=dev) for _ in range(3)]
    w = torch.ones(d, device=dev)
    rstd = torch.rand(T, device=dev) + 0.5
    dw = torch.zeros(d, device=dev)
    gu = [torch.randn(T, 2 * I, device=dev).to(BF16) for _ in range(2)]
    da = [torch.randn(T, I, device=dev).to(BF16) for _ in range(2)]
    k = [0]

    def f_rms_bwd():
        i = k[0] = (k[0] + 1) % 3
        ops.rmsnorm_bwd(ys[i], xs[i], rstd, w, rs[i], dw)

    def f_add_rms():
        i = k[0] = (k[0] + 1) % 3
        ops.add_rmsnorm_fwd(xs[i], ys[i], w, 1e-6)

    def f_rms_fwd():
        i = k[0] = (k[0] + 1) % 3
        ops.rmsnorm_fwd(xs[i], w, 1e-6)

    def f_geglu_bwd():
        i = k[0] = (k[0] + 1) % 2
        ops.geglu_bwd(da[i], gu[i])

    cases = [("rmsnorm_bwd", f_rms_bwd, T * d * 16), ("add_rmsnorm_fwd", f_add_rms, T * d * 12),
             ("rmsnorm_fwd", f_rms_fwd, T * d * 6), ("geglu_bwd", f_geglu_bwd, T * I * 10)]
    lines = []
    for name, fn, nbytes in cases:
        ms = timeit(fn)
        lines.append(f"{name:18s} {ms:7.3f} ms  {nbytes / ms / 1e6:7.1f} GB/s  ({nbytes / 1e6:.0f} MB algorithmic)")
        print(lines[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"elem_bench_{tag}.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
