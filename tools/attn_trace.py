#!/usr/bin/env python
"""Per-phase timeline of the general attention forward kernel (one softmax warp of 64 CTAs), from clock64 stamps compiled
in with -DGGPT_ATTN_TRACE:   python tools/attn_trace.py        (builds a scratch copy of the library under /tmp)
slots: 0 loop top, 6 s_full seen, 1 scores in registers, 2 max + vote done, 3 exp done, 4 pv_done(it-1) seen, 5 P stored."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
csrc = os.path.join(ROOT, "graph-gpt_b200", "csrc")
so = "/tmp/libggpt_trace.so"
srcs = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc)) if f.endswith(".cu")]
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-DGGPT_ATTN_TRACE"] + os.environ.get("GGPT_TRACE_DEFS", "").split() + [
                       "-I", os.path.join(ROOT, "include"), "-shared", "-o", so] + srcs + ["-lcudart"])
import torch  # noqa: E402

import graphgpt_b200.lib as L  # noqa: E402
from graphgpt_b200 import ops  # noqa: E402

L.LIB_PATH = so
L.lib.load()
N, S, H = 64, 1024, 12
d = H * 64
qkv = (torch.randn(N * S, 3 * d, device="cuda") * 0.7).to(torch.bfloat16)
mask = ops.attn_mask_build(None, N, S, False, "cuda")
for _ in range(3):
    ops.attn_fwd(qkv, mask, H, want_lo=True)
torch.cuda.synchronize()
dll = ctypes.CDLL(so)
buf = (ctypes.c_longlong * (64 * 16 * 8))()
assert dll.ggpt_debug_attn_trace(buf, 64 * 16 * 8) == 0
import numpy as np  # noqa: E402
full = np.array(buf[:], dtype=np.int64).reshape(64, 16, 8)
t = full[:, :8]
hh = full[:, 8:11]          # per head (3 heads per CTA): 0 head start, 1 loop end, 2 last PV seen, 3 l exchanged, 4 staged, 5 pair barrier, 6 flushed
print("per head (cycles): tile loop %.0f | wait last PV %.0f | l exchange %.0f | O ld + staging %.0f | pair barrier %.0f | flush %.0f | head period %.0f" % (
    (hh[:, :, 1] - hh[:, :, 0]).mean(), (hh[:, :, 2] - hh[:, :, 1]).mean(), (hh[:, :, 3] - hh[:, :, 2]).mean(),
    (hh[:, :, 4] - hh[:, :, 3]).mean(), (hh[:, :, 5] - hh[:, :, 4]).mean(), (hh[:, :, 6] - hh[:, :, 5]).mean(),
    (hh[:, 1:, 0] - hh[:, :-1, 0]).mean()))
names = {"wait s_full": (0, 6), "tmem ld": (6, 1), "max+vote": (1, 2), "exp": (2, 3), "wait pv_done": (3, 4), "store P+fence": (4, 5)}
print("mean cycles per phase over 64 CTAs x tiles 1..7 (tile 0 separately)")
for k, (a, b) in names.items():
    dlt = t[:, :, b] - t[:, :, a]
    print(f"  {k:16s} tile0 {dlt[:, 0].mean():8.0f}   tiles1-7 mean {dlt[:, 1:].mean():8.0f}  p90 {np.percentile(dlt[:, 1:], 90):8.0f}")
per_tile = t[:, 1:, 0] - t[:, :-1, 0]
print(f"  loop period      mean {per_tile.mean():8.0f}  p90 {np.percentile(per_tile, 90):8.0f}")
print(f"  CTA span (tile loop only) mean {(t[:, 7, 5] - t[:, 0, 0]).mean():8.0f}")

# ---- packed (isolated-tile) backward: compute warp 2 of 32 CTAs, items 8..23
if "diag" in sys.argv:
    from graphgpt_b200 import synth  # noqa: E402
    b = synth.make_batch(N, S, layout="packed", seed=3)
    am = torch.from_numpy(b["attention_mask"]).cuda()
    pmask = ops.attn_mask_build(am, N, S, False, "cuda")
    pos = torch.arange(S, device="cuda", dtype=torch.int32).repeat(N)
    inv = 1.0 / (10000.0 ** (torch.arange(0, 64, 2).float() / 64))
    fr = torch.arange(S).float()[:, None] * inv[None]
    cos, sin = fr.cos().cuda().contiguous(), fr.sin().cuda().contiguous()
    dout = (torch.randn(N * S, d, device="cuda") * 0.1).to(torch.bfloat16)
    out, lse = ops.attn_fwd(qkv, pmask, H, dropout_p=0.1, seed=7)
    for _ in range(3):
        ops.attn_bwd(dout, qkv, out, lse, pmask, H, pos, cos, sin, dropout_p=0.1, seed=7)
    torch.cuda.synchronize()
    buf2 = (ctypes.c_longlong * (32 * 16 * 8))()
    assert dll.ggpt_debug_diag_trace(buf2, 32 * 16 * 8) == 0
    t2 = np.array(buf2[:], dtype=np.int64).reshape(32, 16, 8)[:, 1:15]
    ph = {"prefetch": (0, 1), "wait S/dP": (1, 2), "wait P/dS buffer": (2, 3), "softmax-gradient pass": (3, 4),
          "MMA thread: gradient MMAs issued -> accumulators ready": (5, 6), "warp 2 published P/dS -> MMA thread issues": (4, 5), "epilogue warp: drain + rotate + store": (6, 7)}
    print("packed backward, cycles per item (compute warp 2, 32 CTAs x 14 items):")
    for k, (a, b_) in ph.items():
        dlt = t2[:, :, b_] - t2[:, :, a]
        print(f"  {k:40s} mean {dlt.mean():8.0f}  p90 {np.percentile(dlt, 90):8.0f}")
    lat = t2[:, :, 6] - t2[:, :, 4]
    print(f"  P/dS published -> accumulators ready (gradient MMAs) mean {lat.mean():8.0f}  p90 {np.percentile(lat, 90):8.0f}")
    lat2 = t2[:, 2:, 2] - t2[:, :-2, 7]
    print(f"  epilogue(i-2) done -> scores(i) ready               mean {lat2.mean():8.0f}  p90 {np.percentile(lat2, 90):8.0f}")
    per = t2[:, 1:, 0] - t2[:, :-1, 0]
    print(f"  item period                              mean {per.mean():8.0f}  p90 {np.percentile(per, 90):8.0f}")

# ---- general (tile-loop) backward, dense: softmax-gradient warp 2 of 64 CTAs
if "bwd" in sys.argv:
    pos = torch.arange(S, device="cuda", dtype=torch.int32).repeat(N)
    inv = 1.0 / (10000.0 ** (torch.arange(0, 64, 2).float() / 64))
    fr = torch.arange(S).float()[:, None] * inv[None]
    cos, sin = fr.cos().cuda().contiguous(), fr.sin().cuda().contiguous()
    dout = (torch.randn(N * S, d, device="cuda") * 0.1).to(torch.bfloat16)
    out, lse, out_lo = ops.attn_fwd(qkv, mask, H, want_lo=True)
    for _ in range(3):
        ops.attn_bwd(dout, qkv, out, lse, mask, H, pos, cos, sin, out_lo=out_lo)
    torch.cuda.synchronize()
    buf3 = (ctypes.c_longlong * (64 * 12 * 8))()
    assert dll.ggpt_debug_bwd_trace(buf3, 64 * 12 * 8) == 0
    t3 = np.array(buf3[:], dtype=np.int64).reshape(64, 12, 8)
    tl, cta = t3[:, :8], t3[:, 8]
    ph = {"prefetch + wait S/dP": (0, 1), "TMEM drain": (1, 2), "P / dS arithmetic": (2, 3), "wait gradient MMAs(it-1) + dQ drain": (3, 4),
          "P / dS store + fence": (4, 5), "dQ staging + TMA reduce issue": (5, 6)}
    print("general backward (dense), cycles per query tile, tiles 1..7:")
    for k, (a, b_) in ph.items():
        dlt = tl[:, 1:, b_] - tl[:, 1:, a]
        print(f"  {k:40s} mean {dlt.mean():8.0f}  p90 {np.percentile(dlt, 90):8.0f}")
    per = tl[:, 1:, 0] - tl[:, :-1, 0]
    print(f"  tile period                              mean {per.mean():8.0f}")
    print(f"  CTA: kernel entry -> loop start {(cta[:, 1] - cta[:, 0]).mean():8.0f} | tile loop {(cta[:, 2] - cta[:, 1]).mean():8.0f} | "
          f"last dQ + epilogue {(cta[:, 3] - cta[:, 2]).mean():8.0f}")

# ---- tcgen05 GEMM epilogues: epilogue warp 2 and the MMA thread of 32 leader CTAs, first 24 tiles
if "gemm" in sys.argv:
    T, dm, I = 65536, 768, 3072
    x = (torch.randn(T, dm, device="cuda") * 0.5).to(torch.bfloat16)
    w_qkv = (torch.randn(3 * dm, dm, device="cuda") * 0.05).to(torch.bfloat16)
    w_gu = (torch.randn(2 * I, dm, device="cuda") * 0.05).to(torch.bfloat16)
    w_o = (torch.randn(dm, dm, device="cuda") * 0.05).to(torch.bfloat16)
    posg = torch.arange(T, device="cuda", dtype=torch.int32) % 1024
    inv = 1.0 / (10000.0 ** (torch.arange(0, 64, 2).float() / 64))
    fr = torch.arange(1024).float()[:, None] * inv[None]
    cosg, sing = fr.cos().cuda().contiguous(), fr.sin().cuda().contiguous()
    cases = {"plain fwd [T,768]x[768,768]": lambda: ops.gemm(x, w_o),
             "qkv+rope [T,768]x[2304,768]": lambda: ops.gemm_qkv_rope(x, w_qkv, posg, cosg, sing, 2 * dm),
             "geglu (training) [T,768]x[6144,768]": lambda: ops.gemm_geglu(x, w_gu, want_gu=True)}
    for name, fn in cases.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        bufg = (ctypes.c_longlong * (32 * 24 * 8))()
        assert dll.ggpt_debug_gemm_trace(bufg, 32 * 24 * 8) == 0
        tg = np.array(bufg[:], dtype=np.int64).reshape(32, 24, 8)[:, 2:8]
        print(f"{name}: cycles per tile (leader CTAs, tiles 2..7)")
        print(f"   epilogue warp: prologue (gathers) {np.mean(tg[:, :, 1] - tg[:, :, 0]):7.0f} | wait accumulator {np.mean(tg[:, :, 2] - tg[:, :, 1]):7.0f} | "
              f"drain + math + stores {np.mean(tg[:, :, 3] - tg[:, :, 2]):7.0f} | tile period {np.mean(tg[:, 1:, 0] - tg[:, :-1, 0]):7.0f}")
        print(f"   MMA thread   : wait TMEM buffer {np.mean(tg[:, :, 6] - tg[:, :, 5]):7.0f} | mainloop issue {np.mean(tg[:, :, 7] - tg[:, :, 6]):7.0f} | "
              f"tile period {np.mean(tg[:, 1:, 5] - tg[:, :-1, 5]):7.0f}")
