"""Thin tensor-level wrappers over the C ABI (include/ggpt_b200.h).

Each function validates dtypes/shapes, hands raw device pointers + the current CUDA stream to the library and
returns the output tensor.  No arithmetic happens in Python and nothing here falls back to torch ops.
"""
import torch

from .lib import lib

BF16 = torch.bfloat16
F32 = torch.float32


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _check(t, dtype, name, dims=None):
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the GraphGPT hot path has no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if dims is not None and t.dim() != dims:
        raise RuntimeError(f"{name}: expected {dims} dims, got shape {tuple(t.shape)}")
    if t.dim() >= 1 and t.stride(-1) != 1:
        raise RuntimeError(f"{name}: innermost dimension must be contiguous")


def _ld(t):
    return t.stride(0) if t.dim() == 2 else t.shape[-1]


# ------------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------------
def gemm(a, b, *, a_mn_major=False, b_mn_major=False, out=None, out_dtype=BF16, accumulate=False):
    """C[M,N] = A * B^T with A logical [M,K], B logical [N,K].

    a_mn_major: `a` is stored [K, M] (so C = a^T @ ...);  b_mn_major: `b` is stored [K, N] (so C = ... @ b).
      forward  y  = x @ W.T   -> gemm(x, W)
      dgrad    dx = dy @ W    -> gemm(dy, W, b_mn_major=True)
      wgrad    dW = dy.T @ x  -> gemm(dy, x, a_mn_major=True, b_mn_major=True, out_dtype=F32)
    """
    _check(a, BF16, "gemm a", 2)
    _check(b, BF16, "gemm b", 2)
    M, K = (a.shape[1], a.shape[0]) if a_mn_major else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn_major else (b.shape[0], b.shape[1])
    if K != Kb:
        raise RuntimeError(f"gemm: K mismatch {K} vs {Kb}")
    if out is None:
        ldc = (N + 7) // 8 * 8
        buf = torch.empty((M, ldc), device=a.device, dtype=out_dtype)
        out = buf[:, :N]
    else:
        _check(out, out_dtype, "gemm out", 2)
    lib.ggpt_gemm_bf16(a.data_ptr(), a.stride(0), int(a_mn_major), b.data_ptr(), b.stride(0), int(b_mn_major),
                       out.data_ptr(), out.stride(0), int(out.dtype == F32), int(accumulate), M, N, K, _stream())
    return out


def gemm_resid(a, w, resid, *, colscale=None, rowscale=None, out=None):
    """out = resid + rowscale[:,None] * colscale[None,:] * (a @ w.T); fp32 residual stream."""
    _check(a, BF16, "gemm_resid a", 2)
    _check(w, BF16, "gemm_resid w", 2)
    _check(resid, F32, "gemm_resid resid", 2)
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=F32)
    lib.ggpt_gemm_bf16_resid(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), resid.data_ptr(), resid.stride(0),
                             _ptr(colscale), _ptr(rowscale), out.data_ptr(), out.stride(0), M, N, K, _stream())
    return out


def gemm_geglu(a, wgu, *, want_gu=True):
    """wgu = [gate_proj.weight ; up_proj.weight] ([2I, K]).  Returns (gu [M,2I] or None, act [M,I])."""
    _check(a, BF16, "gemm_geglu a", 2)
    _check(wgu, BF16, "gemm_geglu wgu", 2)
    M, K = a.shape
    N2 = wgu.shape[0]
    gu = torch.empty((M, N2), device=a.device, dtype=BF16) if want_gu else None
    act = torch.empty((M, N2 // 2), device=a.device, dtype=BF16)
    lib.ggpt_gemm_bf16_geglu(a.data_ptr(), a.stride(0), wgu.data_ptr(), wgu.stride(0), _ptr(gu),
                             N2, act.data_ptr(), act.stride(0), M, N2, K, _stream())
    return gu, act


def gemm_qkv_rope(a, wqkv, pos, cos_tab, sin_tab, rope_cols):
    """qkv[M,3d] = a @ wqkv.T with rotary embedding applied to the first rope_cols columns (heads of 64)."""
    _check(a, BF16, "gemm_qkv_rope a", 2)
    _check(wqkv, BF16, "gemm_qkv_rope wqkv", 2)
    _check(pos, torch.int32, "gemm_qkv_rope pos", 1)
    _check(cos_tab, F32, "cos_tab", 2)
    _check(sin_tab, F32, "sin_tab", 2)
    M, K = a.shape
    N = wqkv.shape[0]
    out = torch.empty((M, N), device=a.device, dtype=BF16)
    lib.ggpt_gemm_bf16_qkv_rope(a.data_ptr(), a.stride(0), wqkv.data_ptr(), wqkv.stride(0), out.data_ptr(), N,
                                pos.data_ptr(), cos_tab.data_ptr(), sin_tab.data_ptr(), rope_cols, M, N, K, _stream())
    return out
