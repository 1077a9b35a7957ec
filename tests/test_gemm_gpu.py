"""GPU parity of the tcgen05 GEMM (all operand majors and fused epilogues) against torch fp32 matmul of the
same bf16-rounded operands.  Tolerances: fp32 accumulation order only -> 2e-3 of the output scale for bf16
outputs (one bf16 rounding, eps 2^-8 relative), 1e-4 for fp32 outputs."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).cuda()


def _report(name, got, ref, tol):
    got = got.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-12
    rel = err / scale
    if not rel <= tol:
        bad = ((got - ref).abs() > tol * scale).nonzero()
        msg = f"{name}: max rel err {rel:.3e} > {tol:.1e}; {bad.shape[0]} bad elements; first {bad[:8].tolist()}"
        rows = torch.unique(bad[:, 0])[:16].tolist()
        cols = torch.unique(bad[:, 1])[:16].tolist()
        msg += f"; bad rows {rows}; bad cols {cols}"
        raise AssertionError(msg)
    return rel


SHAPES = [
    (128, 128, 64),
    (128, 256, 64),
    (256, 256, 128),
    (128, 128, 768),
    (1000, 760, 768),     # M tail, N tail (lm_head-like, ld padded)
    (4096, 2304, 768),    # QKV
    (4096, 768, 3072),    # down_proj
    (300, 136, 200),      # everything ragged (K tail through TMA zero fill)
]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_tn(M, N, K):
    from graphgpt_b200 import ops
    Kp = (K + 7) // 8 * 8
    a = _rand((M, Kp), 1)[:, :K]
    b = _rand((N, Kp), 2)[:, :K]
    ref = a.float() @ b.float().t()
    out = ops.gemm(a, b)
    torch.cuda.synchronize()
    _report(f"tn bf16 {M}x{N}x{K}", out, ref, 6e-3)
    out32 = ops.gemm(a, b, out_dtype=torch.float32)
    torch.cuda.synchronize()
    _report(f"tn f32 {M}x{N}x{K}", out32, ref, 1e-4)


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_dgrad_b_mn_major(M, N, K):
    """dx[M,N] = dy[M,K] @ W[K,N]  (W read in place, MN-major B)"""
    from graphgpt_b200 import ops
    Kp = (K + 7) // 8 * 8
    Np = (N + 7) // 8 * 8
    a = _rand((M, Kp), 3)[:, :K]
    w = _rand((K, Np), 4)[:, :N]
    ref = a.float() @ w.float()
    out = ops.gemm(a, w, b_mn_major=True)
    torch.cuda.synchronize()
    _report(f"dgrad {M}x{N}x{K}", out, ref, 6e-3)


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_wgrad_both_mn_major(M, N, K):
    """dW[M,N] = dy[K,M]^T @ x[K,N]"""
    from graphgpt_b200 import ops
    Mp = (M + 7) // 8 * 8
    Np = (N + 7) // 8 * 8
    dy = _rand((K, Mp), 5)[:, :M]
    x = _rand((K, Np), 6)[:, :N]
    ref = dy.float().t() @ x.float()
    out = ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, out_dtype=torch.float32)
    torch.cuda.synchronize()
    _report(f"wgrad {M}x{N}x{K}", out, ref, 1e-4)
    out2 = ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, out=out.clone(), out_dtype=torch.float32, accumulate=True)
    torch.cuda.synchronize()
    _report(f"wgrad acc {M}x{N}x{K}", out2, 2 * ref, 1e-4)


def test_gemm_resid():
    from graphgpt_b200 import ops
    M, N, K = 1000, 768, 3072
    a = _rand((M, K), 7, 0.5)
    w = _rand((N, K), 8, 0.05)
    g = torch.Generator().manual_seed(9)
    resid = torch.randn((M, N), generator=g).cuda()
    cs = torch.rand((N,), generator=g).cuda()
    rs = torch.rand((M,), generator=g).cuda()
    y = a.float() @ w.float().t()
    out = ops.gemm_resid(a, w, resid)
    torch.cuda.synchronize()
    _report("resid", out, resid + y, 1e-4)
    out = ops.gemm_resid(a, w, resid, colscale=cs, rowscale=rs)
    torch.cuda.synchronize()
    _report("resid scaled", out, resid + rs[:, None] * cs[None, :] * y, 1e-4)


@pytest.mark.parametrize("M,I,K", [(1000, 3072, 768), (130, 256, 64), (77, 512, 128)])
def test_gemm_geglu(M, I, K):
    from graphgpt_b200 import ops
    a = _rand((M, K), 10)
    wgu = _rand((2 * I, K), 11, 1.0 / math.sqrt(K))
    y = a.float() @ wgu.float().t()
    gf, act = ops.gemm_geglu(a, wgu)
    torch.cuda.synchronize()
    # training mode stores the backward factors gf = [u * gelu'(g) | gelu(g)] instead of [g | u]
    g_, u_ = y[:, :I].clone().requires_grad_(True), y[:, I:]
    gelu_g = torch.nn.functional.gelu(g_)
    gelu_g.sum().backward()
    _report("geglu factor u*gelu'(g)", gf[:, :I], u_ * g_.grad, 6e-3)
    _report("geglu factor gelu(g)", gf[:, I:], gelu_g.detach(), 6e-3)
    ref_act = gelu_g.detach() * u_
    _report("geglu act", act, ref_act, 6e-3)
    _, act2 = ops.gemm_geglu(a, wgu, want_gu=False)          # inference path: 1-MUFU GELU, act only
    torch.cuda.synchronize()
    _report("geglu act (inference epilogue)", act2, ref_act, 6e-3)
    assert (act.float() - act2.float()).abs().max().item() <= 2e-2 * ref_act.abs().max().item()


@pytest.mark.parametrize("M,d,K", [(1000, 768, 768), (200, 64, 64), (300, 128, 128)])
def test_gemm_qkv_rope(M, d, K):
    from graphgpt_b200 import ops
    a = _rand((M, K), 12)
    w = _rand((3 * d, K), 13, 1.0 / math.sqrt(K))
    max_pos = 1024
    g = torch.Generator().manual_seed(14)
    pos = torch.randint(0, max_pos, (M,), generator=g, dtype=torch.int32).cuda()
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, 64, 2, dtype=torch.int64).float() / 64))
    freqs = torch.arange(max_pos).float()[:, None] * inv_freq[None, :]
    cos_tab, sin_tab = freqs.cos().cuda(), freqs.sin().cuda()
    out = ops.gemm_qkv_rope(a, w, pos, cos_tab, sin_tab, 2 * d)
    torch.cuda.synchronize()
    y = a.float() @ w.float().t()
    H = d // 64
    qk = y[:, : 2 * d].reshape(M, 2 * H, 64)
    cos = torch.cat([cos_tab, cos_tab], -1)[pos.long()][:, None, :]
    sin = torch.cat([sin_tab, sin_tab], -1)[pos.long()][:, None, :]
    rot = torch.cat([-qk[..., 32:], qk[..., :32]], -1)
    qk = (qk * cos + rot * sin).reshape(M, 2 * d)
    ref = torch.cat([qk, y[:, 2 * d:]], -1)
    _report("qkv_rope", out, ref, 6e-3)


@pytest.mark.parametrize("M,N,K", [(768, 768, 20000), (2304, 768, 8192), (756, 768, 5000)])
def test_gemm_wgrad_split_k(M, N, K):
    """Long-K weight gradients take the split-K path (vector red.global.add into the fp32 gradient buffer)."""
    from graphgpt_b200 import ops
    Mp = (M + 7) // 8 * 8
    dy = _rand((K, Mp), 21, 0.5)[:, :M]
    x = _rand((K, N), 22, 0.5)
    ref = dy.float().t() @ x.float()
    base = torch.randn((M, N), generator=torch.Generator().manual_seed(23)).cuda()
    out = ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, out=base.clone(), out_dtype=torch.float32, accumulate=True)
    torch.cuda.synchronize()
    _report(f"wgrad split-k {M}x{N}x{K}", out, base + ref, 1e-4)


@pytest.mark.parametrize("M,I,K", [(1000, 3072, 768), (130, 256, 64), (300, 320, 128), (9000, 1024, 256), (77, 512, 768)])
def test_gemm_dgeglu(M, I, K):
    """Fused down_proj dgrad + GeGLU backward against torch autograd of gelu(g)*u."""
    from graphgpt_b200 import ops
    dy = _rand((M, K), 31)
    wd = _rand((K, I), 32, 1.0 / math.sqrt(K))
    gu = _rand((M, 2 * I), 33)
    dact = dy.float() @ wd.float()
    gg = gu.float().clone().requires_grad_(True)
    (torch.nn.functional.gelu(gg[:, :I]) * gg[:, I:]).backward(dact)
    # the factors the forward epilogue would have stored for these gate/up values
    g_ = gu.float()[:, :I].clone().requires_grad_(True)
    gelu_g = torch.nn.functional.gelu(g_)
    gelu_g.sum().backward()
    gf = torch.cat([gu.float()[:, I:] * g_.grad, gelu_g.detach()], dim=1).to(torch.bfloat16)
    dgu = ops.gemm_dgeglu(dy, wd, gf)
    torch.cuda.synchronize()
    _report("dgeglu", dgu, gg.grad, 8e-3)
