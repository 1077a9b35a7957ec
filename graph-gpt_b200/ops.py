"""Thin tensor-level wrappers over the C ABI (include/ggpt_b200.h).

Each function validates dtypes/shapes, hands raw device pointers + the current CUDA stream to the library and
returns the output tensor.  No arithmetic happens in Python and nothing here falls back to torch ops.
"""
import os

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
    """wgu = [gate_proj.weight ; up_proj.weight] ([2I, K]).  Returns (gf [M,2I] or None, act [M,I]) with
    act = gelu(g) * u and gf = [u * gelu'(g) | gelu(g)], the factors GeGLU backward multiplies dact with."""
    _check(a, BF16, "gemm_geglu a", 2)
    _check(wgu, BF16, "gemm_geglu wgu", 2)
    M, K = a.shape
    N2 = wgu.shape[0]
    gu = torch.empty((M, N2), device=a.device, dtype=BF16) if want_gu else None
    act = torch.empty((M, N2 // 2), device=a.device, dtype=BF16)
    lib.ggpt_gemm_bf16_geglu(a.data_ptr(), a.stride(0), wgu.data_ptr(), wgu.stride(0), _ptr(gu),
                             N2, act.data_ptr(), act.stride(0), M, N2, K, _stream())
    return gu, act


def gemm_dgeglu(dy, wd, gu):
    """dgu[M,2I] = [dact * gf[:, :I] | dact * gf[:, I:]] with dact = dy @ wd (never stored); gf = the factors saved by
    gemm_geglu, wd = down_proj.weight [d, I]."""
    _check(dy, BF16, "gemm_dgeglu dy", 2)
    _check(wd, BF16, "gemm_dgeglu wd", 2)
    _check(gu, BF16, "gemm_dgeglu gu", 2)
    M, K = dy.shape
    I = wd.shape[1]
    dgu = torch.empty_like(gu)
    lib.ggpt_gemm_bf16_dgeglu(dy.data_ptr(), dy.stride(0), wd.data_ptr(), wd.stride(0), gu.data_ptr(), gu.stride(0),
                              dgu.data_ptr(), dgu.stride(0), M, I, K, _stream())
    return dgu


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


# ------------------------------------------------------------------------------------------------
# Attention
# ------------------------------------------------------------------------------------------------
class AttnMask:
    """Bit-matrix form of the reference's additive attention mask (built once per step, shared by all layers)."""

    def __init__(self, bits, tile_start, n_tiles, cls, iso_flags, iso_list, iso_count, N, S):
        self.bits, self.tile_start, self.n_tiles, self.cls, self.N, self.S = bits, tile_start, n_tiles, cls, N, S
        self.iso_flags, self.iso_list, self.iso_count = iso_flags, iso_list, iso_count
        self.use_diag = os.environ.get("GGPT_ATTN_NO_DIAG") is None   # debug switch: force the general kernels
        # [#isolated tiles, #tiles] copied to pinned host memory right behind the mask build; read lazily so the
        # host only waits for the mask kernels, never for the layers already queued behind them
        self._counts_host = torch.empty((2,), dtype=torch.int32, pin_memory=True)
        self._counts_host.copy_(iso_count, non_blocking=True)
        self._counts_event = torch.cuda.Event()
        self._counts_event.record()
        self._run_general = None

    def _iso_args(self):
        """(iso_flags, iso_list, iso_count, run_general) for ggpt_attn_fwd / ggpt_attn_bwd."""
        if not self.use_diag:
            return 0, 0, 0, 1
        if self._run_general is None:
            self._counts_event.synchronize()
            iso, total = self._counts_host.tolist()
            self._run_general = int(iso != total)     # packed batches: every tile isolated -> general kernels skipped
        return self.iso_flags.data_ptr(), self.iso_list.data_ptr(), self.iso_count.data_ptr(), self._run_general


def attn_mask_build(attention_mask, N, S, causal, device):
    """attention_mask: None, int64 [N,S] or int64 [N,S,S] (modeling_helpers.py:38-64)."""
    words = lib.ggpt_attn_mask_words(S)
    mt = lib.ggpt_attn_max_tiles(S)
    bits = torch.empty((N, S, words), device=device, dtype=torch.int32)
    tile_start = torch.empty((N, mt + 1), device=device, dtype=torch.int32)
    n_tiles = torch.empty((N,), device=device, dtype=torch.int32)
    cls = torch.empty((N, mt, mt), device=device, dtype=torch.uint8)
    iso_flags = torch.empty((N, mt), device=device, dtype=torch.uint8)
    iso_list = torch.empty((N * mt, 4), device=device, dtype=torch.int32)    # {sequence, first row, rows, class} per tile
    iso_count = torch.empty((2,), device=device, dtype=torch.int32)
    dims = 0
    if attention_mask is not None:
        if attention_mask.dim() not in (2, 3):
            raise NotImplementedError(f"attention_mask of shape {tuple(attention_mask.shape)} is not Implemented")
        if attention_mask.dtype != torch.int64:
            attention_mask = attention_mask.to(torch.int64)
        attention_mask = attention_mask.contiguous()
        _check(attention_mask, torch.int64, "attention_mask")
        dims = attention_mask.dim()
        if attention_mask.shape[0] != N or attention_mask.shape[-1] != S:
            raise RuntimeError(f"attention_mask shape {tuple(attention_mask.shape)} does not match N={N}, S={S}")
    lib.ggpt_attn_mask_build(_ptr(attention_mask), dims, N, S, int(bool(causal)), bits.data_ptr(), tile_start.data_ptr(),
                             n_tiles.data_ptr(), cls.data_ptr(), iso_flags.data_ptr(), iso_list.data_ptr(),
                             iso_count.data_ptr(), _stream())
    return AttnMask(bits, tile_start, n_tiles, cls, iso_flags, iso_list, iso_count, N, S)


def attn_fwd(qkv, mask: AttnMask, H, *, want_lse=True, want_lo=False, dropout_p=0.0, seed=0):
    """qkv bf16 [N*S, 3*H*64] (q | k | v).  Returns (out bf16 [N*S, H*64], lse f32 [N,H,S] or None) and, with want_lo, a
    third element: the bf16 rounding residual of `out` that attn_bwd needs on the general (tile-loop) path — None when
    every tile is isolated (packed batches) and the diagonal kernels do all the work, or when want_lse is off (inference)."""
    _check(qkv, BF16, "attn_fwd qkv", 2)
    N, S = mask.N, mask.S
    d = H * 64
    if qkv.shape[0] != N * S or qkv.shape[1] != 3 * d:
        raise RuntimeError(f"attn_fwd: qkv shape {tuple(qkv.shape)} does not match N={N} S={S} H={H}")
    out = torch.empty((N * S, d), device=qkv.device, dtype=BF16)
    lse = torch.empty((N, H, S), device=qkv.device, dtype=F32) if want_lse else None
    iso = mask._iso_args()
    out_lo = torch.empty((N * S, d), device=qkv.device, dtype=BF16) if (want_lo and want_lse and iso[3]) else None
    lib.ggpt_attn_fwd(qkv.data_ptr(), qkv.stride(0), 0, d, 2 * d, mask.bits.data_ptr(), mask.tile_start.data_ptr(),
                      mask.n_tiles.data_ptr(), mask.cls.data_ptr(), *iso, float(dropout_p), int(seed),
                      out.data_ptr(), _ptr(out_lo), out.stride(0), _ptr(lse), N, S, H, _stream())
    if want_lo:
        return out, lse, out_lo
    return out, lse


_DQ_ACC = {}


def _dq_accumulator(device, numel):
    """fp32 scratch of the general attention backward: zero on entry, zeroed again by the kernel that consumes it, so one
    buffer per (device, size) is allocated and zero-filled exactly once."""
    key = (device.index, numel)
    buf = _DQ_ACC.get(key)
    if buf is None:
        buf = _DQ_ACC[key] = torch.zeros((numel,), device=device, dtype=F32)
    return buf


def attn_bwd(dout, qkv, out, lse, mask: AttnMask, H, pos, cos_tab, sin_tab, *, out_lo=None, dropout_p=0.0, seed=0):
    """Returns dqkv bf16 [N*S, 3*H*64] (gradient of the fused projection output, RoPE already undone).  out_lo: third
    result of attn_fwd(want_lo=True) (without it D = rowsum(dO * O) carries the bf16 rounding error of O)."""
    _check(dout, BF16, "attn_bwd dout", 2)
    _check(qkv, BF16, "attn_bwd qkv", 2)
    _check(out, BF16, "attn_bwd out", 2)
    if out_lo is not None:
        _check(out_lo, BF16, "attn_bwd out_lo", 2)
        if out_lo.stride(0) != out.stride(0):
            raise RuntimeError("attn_bwd: out_lo must have the layout of out")
    N, S = mask.N, mask.S
    d = H * 64
    dqkv = torch.empty_like(qkv)
    dsum = torch.empty((N, H, S), device=qkv.device, dtype=F32)
    iso = mask._iso_args()
    dq_acc = _dq_accumulator(qkv.device, N * S * d) if iso[3] else None
    lib.ggpt_attn_bwd(qkv.data_ptr(), qkv.stride(0), 0, d, 2 * d, out.data_ptr(), _ptr(out_lo), out.stride(0), dout.data_ptr(),
                      dout.stride(0), lse.data_ptr(), mask.bits.data_ptr(), mask.tile_start.data_ptr(),
                      mask.n_tiles.data_ptr(), mask.cls.data_ptr(), *iso, float(dropout_p), int(seed),
                      pos.data_ptr(), cos_tab.data_ptr(), sin_tab.data_ptr(), dsum.data_ptr(), _ptr(dq_acc), dqkv.data_ptr(),
                      dqkv.stride(0), N, S, H, _stream())
    return dqkv


# ------------------------------------------------------------------------------------------------
# HBM-bound kernels
# ------------------------------------------------------------------------------------------------
def embed_fwd(ids, table, gate=None, long_scale=False, err_flag=None, *, drop_p=0.0, drop_seed=0):
    """ids int64 [T,F]; table f32 [V,d] -> x f32 [T,d].  drop_p > 0: embed_dropout on the gathered rows."""
    _check(ids, torch.int64, "embed ids", 2)
    _check(table, F32, "embed table", 2)
    T, F_ = ids.shape
    V, d = table.shape
    out = torch.empty((T, d), device=table.device, dtype=F32)
    lib.ggpt_embed_fwd(ids.data_ptr(), table.data_ptr(), _ptr(gate), out.data_ptr(), T, F_, d, V, int(long_scale),
                       _ptr(err_flag), float(drop_p), int(drop_seed), _stream())
    return out


def embed_bwd(ids, dx, table, gate, dtable, dgate, padding_idx=0, long_scale=False, *, drop_p=0.0, drop_seed=0):
    _check(ids, torch.int64, "embed ids", 2)
    _check(dx, F32, "embed dx", 2)
    T, F_ = ids.shape
    V, d = dtable.shape
    lib.ggpt_embed_bwd(ids.data_ptr(), dx.data_ptr(), _ptr(table), _ptr(gate), dtable.data_ptr(), _ptr(dgate), T, F_, d, V,
                       padding_idx, int(long_scale), float(drop_p), int(drop_seed), _stream())


def embed_count(ids, V, padding_idx):
    """bf16 [T, ceil8(V)] count matrix for the GEMM form of the embedding gradient."""
    _check(ids, torch.int64, "embed ids", 2)
    T, F_ = ids.shape
    ldc = (V + 7) // 8 * 8
    cnt = torch.empty((T, ldc), device=ids.device, dtype=BF16)
    lib.ggpt_embed_count(ids.data_ptr(), cnt.data_ptr(), ldc, T, F_, V, padding_idx, _stream())
    return cnt[:, :V]


def smtp_mask_2d(ids, F, mr, u_node, power, *, mask_token=1, label_pad=-100, err_flag=None):
    """In-model SMTP masking.  ids int64 [N,S,Ftot] (column F+2 = node index), mr f32 [N], u_node f32 [N,S,F].
    Returns (input_ids int64 [N,S,F], labels int64 [N,S,F])."""
    if ids.dtype != torch.int64 or ids.dim() != 3 or not ids.is_cuda or not ids.is_contiguous():
        raise RuntimeError("smtp_mask_2d: ids must be a contiguous CUDA int64 [N,S,Ftot] tensor")
    N, S, Ftot = ids.shape
    if Ftot < F + 3:
        raise RuntimeError(f"smtp_mask_2d: ids has {Ftot} columns, expected stacked_feat + 4 = {F + 4} (node index at {F + 2})")
    _check(mr, F32, "smtp_mask_2d mr")
    _check(u_node, F32, "smtp_mask_2d u_node")
    if mr.numel() != N or u_node.numel() != N * S * F:
        raise RuntimeError("smtp_mask_2d: mr / u_node sizes do not match ids")
    out = torch.empty((N, S, F), device=ids.device, dtype=torch.int64)
    labels = torch.empty((N, S, F), device=ids.device, dtype=torch.int64)
    lib.ggpt_smtp_mask_2d(ids.data_ptr(), Ftot, F + 2, mr.data_ptr(), u_node.data_ptr(), float(power), out.data_ptr(),
                          labels.data_ptr(), N, S, F, int(mask_token), int(label_pad), _ptr(err_flag), _stream())
    return out, labels


def raw_embed_norm_fwd(raw, w, eps, *, labels=None, fchk=0, mask_tok=None, want_stash=True):
    """raw f32 [T,E]; labels int64 [T,F] or None.  Returns (h bf16 [T,E], rstd f32 [T] | None, keep u8 [T] | None):
    rows whose first `fchk` labels are all labelled take `mask_tok` instead of their raw features, then RMSNorm."""
    _check(raw, F32, "raw_embed raw", 2)
    _check(w, F32, "raw_embed w", 1)
    if not raw.is_contiguous():
        raise RuntimeError("raw_embed raw: expected a contiguous [T,E] tensor")
    T, E = raw.shape
    ldl = 0
    if labels is not None:
        _check(labels, torch.int64, "raw_embed labels", 2)
        if labels.shape[0] != T or mask_tok is None:
            raise RuntimeError("raw_embed: labels must have T rows and come with mask_tok")
        _check(mask_tok, F32, "raw_embed mask_tok")
        ldl = labels.stride(0)
    h = torch.empty((T, E), device=raw.device, dtype=BF16)
    rstd = torch.empty((T,), device=raw.device, dtype=F32) if want_stash else None
    keep = torch.empty((T,), device=raw.device, dtype=torch.uint8) if (want_stash and labels is not None) else None
    lib.ggpt_raw_embed_norm_fwd(raw.data_ptr(), _ptr(labels), ldl, int(fchk), _ptr(mask_tok), w.data_ptr(), h.data_ptr(), E,
                                _ptr(rstd), _ptr(keep), T, E, float(eps), _stream())
    return h, rstd, keep


def raw_embed_norm_bwd(dh, raw, keep, mask_tok, rstd, w, dw, dmask_tok):
    """Accumulates dw [E] (and dmask_tok [E] when given) from dh bf16 [T,E]."""
    _check(dh, BF16, "raw_embed_bwd dh", 2)
    _check(raw, F32, "raw_embed_bwd raw", 2)
    T, E = raw.shape
    lib.ggpt_raw_embed_norm_bwd(dh.data_ptr(), dh.stride(0), raw.data_ptr(), _ptr(keep), _ptr(mask_tok), rstd.data_ptr(),
                                w.data_ptr(), dw.data_ptr(), _ptr(dmask_tok), T, E, _stream())


def raw_embed_norm_sum_fwd(raw3, w, eps, *, drop_p=0.0, drop_seed=0):
    """raw3 f32 [T,S2,E] (edge embeddings of token t towards every position) -> bf16 [T,E] = w * sum_j rmsnorm-hat rows."""
    _check(raw3, F32, "raw_embed_sum raw", 3)
    if not raw3.is_contiguous():
        raise RuntimeError("raw_embed_sum raw: expected a contiguous [T,S2,E] tensor")
    T, S2, E = raw3.shape
    h = torch.empty((T, E), device=raw3.device, dtype=BF16)
    lib.ggpt_raw_embed_norm_sum(raw3.data_ptr(), w.data_ptr(), h.data_ptr(), E, 0, 0, 0, T, S2, E, float(eps), float(drop_p),
                                int(drop_seed), _stream())
    return h


def raw_embed_norm_sum_bwd(dh, raw3, dw, eps, *, drop_p=0.0, drop_seed=0):
    _check(dh, BF16, "raw_embed_sum dh", 2)
    T, S2, E = raw3.shape
    lib.ggpt_raw_embed_norm_sum(raw3.data_ptr(), 0, 0, 0, dh.data_ptr(), dh.stride(0), dw.data_ptr(), T, S2, E, float(eps),
                                float(drop_p), int(drop_seed), _stream())


def layerscale_bwd(dx, x_out, x_in, lam, rowscale, dlam):
    """dy bf16 [T,d] = dx * lam * rowscale; accumulates dlam [d] += sum_t dx * (x_out - x_in) / lam when dlam is given."""
    _check(dx, F32, "layerscale_bwd dx", 2)
    T, d = dx.shape
    dy = torch.empty((T, d), device=dx.device, dtype=BF16)
    lib.ggpt_layerscale_bwd(dx.data_ptr(), _ptr(x_out), _ptr(x_in), _ptr(lam), _ptr(rowscale), dy.data_ptr(), _ptr(dlam),
                            T, d, _stream())
    return dy


def dropout_(x, p, seed):
    """In-place element dropout on a contiguous bf16 tensor (mask = pure function of (seed, linear index))."""
    _check(x, BF16, "dropout x")
    if not x.is_contiguous():
        raise RuntimeError("dropout: expected a contiguous tensor")
    if p > 0:
        lib.ggpt_dropout_bf16(x.data_ptr(), x.numel(), float(p), int(seed), _stream())
    return x


def dropout_scale(n, p, seed, device):
    """f32 [n] of keep(e)/(1-p): the factors ggpt_dropout_bf16 / ggpt_embed_fwd apply for this (p, seed)."""
    out = torch.empty((n,), device=device, dtype=F32)
    lib.ggpt_dropout_scale_f32(out.data_ptr(), n, float(p), int(seed), _stream())
    return out


def rmsnorm_fwd(x, w, eps, *, want_rstd=True):
    _check(x, F32, "rmsnorm x", 2)
    _check(w, F32, "rmsnorm w", 1)
    T, d = x.shape
    y = torch.empty((T, d), device=x.device, dtype=BF16)
    rstd = torch.empty((T,), device=x.device, dtype=F32) if want_rstd else None
    lib.ggpt_rmsnorm_fwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), d, _ptr(rstd), T, d, float(eps), _stream())
    return y, rstd


def add_rmsnorm_fwd(x_in, y, w, eps, *, colscale=None, rowscale=None, want_rstd=True, want_h=True):
    """x_out = x_in + rowscale*colscale*y; h = rmsnorm(x_out; w) in bf16.  Returns (x_out, h, rstd)."""
    _check(x_in, F32, "add_rmsnorm x_in", 2)
    _check(y, BF16, "add_rmsnorm y", 2)
    T, d = x_in.shape
    x_out = torch.empty_like(x_in)
    h = torch.empty((T, d), device=x_in.device, dtype=BF16) if want_h else None
    rstd = torch.empty((T,), device=x_in.device, dtype=F32) if want_rstd else None
    lib.ggpt_add_rmsnorm_fwd(x_in.data_ptr(), y.data_ptr(), y.stride(0), _ptr(colscale), _ptr(rowscale), w.data_ptr(),
                             x_out.data_ptr(), _ptr(h), _ptr(rstd), T, d, float(eps), _stream())
    return x_out, h, rstd


def rmsnorm_bwd(dy, x, rstd, w, dresid, dw, *, want_bf16=True):
    """Returns (dx f32 [T,d], dx bf16 or None); accumulates into dw."""
    _check(dy, BF16, "rmsnorm_bwd dy", 2)
    _check(x, F32, "rmsnorm_bwd x", 2)
    T, d = x.shape
    dx = torch.empty((T, d), device=x.device, dtype=F32)
    dxb = torch.empty((T, d), device=x.device, dtype=BF16) if want_bf16 else None
    lib.ggpt_rmsnorm_bwd(dy.data_ptr(), dy.stride(0), x.data_ptr(), rstd.data_ptr(), w.data_ptr(), _ptr(dresid),
                         dx.data_ptr(), _ptr(dxb), dw.data_ptr(), T, d, _stream())
    return dx, dxb


def geglu_bwd(dact, gu):
    _check(dact, BF16, "geglu_bwd dact", 2)
    _check(gu, BF16, "geglu_bwd gu", 2)
    T, I = dact.shape
    dgu = torch.empty_like(gu)
    lib.ggpt_geglu_bwd(dact.data_ptr(), gu.data_ptr(), dgu.data_ptr(), T, I, _stream())
    return dgu


class HeadIndex:
    """Device-side compaction of the labelled rows / entries (modeling_helpers.py:263-301).

    (M, L) = (#rows with a label, #labelled entries) size the head GEMMs and the returned [L, V] logits, so the host needs
    them.  The compaction is launched at the START of forward() and the two counts are copied to pinned host memory right
    behind it; the head reads them ~a whole backbone later, when the copy has long completed — the wait is an event that is
    already signalled, and the device never runs dry because of it (the reference's `hidden_states[mask_m]`,
    modeling_helpers.py:288, drains the stream in the middle of every step)."""

    def __init__(self, counts, sel_rows, ent_src, ent_label, ent_tok):
        self.counts, self.sel_rows, self.ent_src, self.ent_label, self.ent_tok = counts, sel_rows, ent_src, ent_label, ent_tok
        self.M = self.L = None
        self._host = torch.empty((2,), dtype=torch.int32, pin_memory=True)
        self._host.copy_(counts, non_blocking=True)
        self._event = torch.cuda.Event()
        self._event.record()

    def sync_counts(self):
        if self.M is None:
            self._event.synchronize()
            self.M, self.L = (int(v) for v in self._host.tolist())
        return self.M, self.L


def head_compact(labels):
    """labels int64 [T,F]."""
    _check(labels, torch.int64, "labels", 2)
    T, F_ = labels.shape
    dev = labels.device
    scratch = torch.empty((lib.ggpt_head_scratch_ints(T),), device=dev, dtype=torch.int32)
    counts = torch.empty((2,), device=dev, dtype=torch.int32)
    sel_rows = torch.empty((T,), device=dev, dtype=torch.int32)
    ent_src = torch.empty((T * F_,), device=dev, dtype=torch.int32)
    ent_label = torch.empty((T * F_,), device=dev, dtype=torch.int32)
    ent_tok = torch.empty((T * F_,), device=dev, dtype=torch.int32)
    lib.ggpt_head_compact(labels.data_ptr(), T, F_, scratch.data_ptr(), counts.data_ptr(), sel_rows.data_ptr(),
                          ent_src.data_ptr(), ent_label.data_ptr(), ent_tok.data_ptr(), _stream())
    hi = HeadIndex(counts, sel_rows, ent_src, ent_label, ent_tok)
    hi.F = F_
    return hi


def gather_rows(src, idx, n, *, n_ptr=None):
    _check(src, BF16, "gather_rows src", 2)
    d = src.shape[1]
    out = torch.empty((n, d), device=src.device, dtype=BF16)
    lib.ggpt_gather_rows(src.data_ptr(), src.stride(0), _ptr(idx), out.data_ptr(), d, _ptr(n_ptr), n, d, _stream())
    return out


def scatter_rows(src, idx, out, n, *, n_ptr=None):
    _check(src, BF16, "scatter_rows src", 2)
    _check(out, BF16, "scatter_rows out", 2)
    lib.ggpt_scatter_rows(src.data_ptr(), src.stride(0), idx.data_ptr(), out.data_ptr(), out.stride(0), _ptr(n_ptr), n,
                          src.shape[1], _stream())
    return out


def expand_rows(src, idx, n, n_out):
    """bf16 [n_out, d] with out[idx[i]] = src[i] (idx strictly increasing) and zero rows elsewhere, in one pass."""
    _check(src, BF16, "expand_rows src", 2)
    _check(idx, torch.int32, "expand_rows idx", 1)
    d = src.shape[1]
    out = torch.empty((n_out, d), device=src.device, dtype=BF16)
    lib.ggpt_expand_rows(src.data_ptr(), src.stride(0), idx.data_ptr(), int(n), out.data_ptr(), d, int(n_out), d, _stream())
    return out


def ce_fwd(logits, labels, V, wgt=None, *, want_row_loss=False, err_flag=None, focal_gamma=0.0):
    """logits f32 [L, ld>=V]; labels int32 [>=L].  Returns (row_lse, row_loss|None, loss_sum f64[1], wgt_sum f64[1])."""
    _check(logits, F32, "ce logits", 2)
    L = logits.shape[0]
    dev = logits.device
    row_lse = torch.empty((L,), device=dev, dtype=F32)
    row_loss = torch.empty((L,), device=dev, dtype=F32) if want_row_loss else None
    sums = torch.zeros((2,), device=dev, dtype=torch.float64)
    lib.ggpt_ce_fwd(logits.data_ptr(), logits.stride(0), labels.data_ptr(), _ptr(wgt), row_lse.data_ptr(), _ptr(row_loss),
                    sums.data_ptr(), sums.data_ptr() + 8, L, V, float(focal_gamma), _ptr(err_flag), _stream())
    return row_lse, row_loss, sums


def ce_finalize(sums, count_ptr, mode, fixed_denom=1.0):
    out = torch.empty((2,), device=sums.device, dtype=F32)  # [loss, scale]
    lib.ggpt_ce_finalize(sums.data_ptr(), sums.data_ptr() + 8, count_ptr, mode, float(fixed_denom), out.data_ptr(),
                         out.data_ptr() + 4, _stream())
    return out


def ce_bwd(logits, labels, V, row_lse, scale_ptr, gout, wgt=None, focal_gamma=0.0):
    L = logits.shape[0]
    ldd = (V + 7) // 8 * 8
    dlogits = torch.empty((L, ldd), device=logits.device, dtype=BF16)
    lib.ggpt_ce_bwd(logits.data_ptr(), logits.stride(0), labels.data_ptr(), _ptr(wgt), row_lse.data_ptr(), scale_ptr,
                    _ptr(gout), dlogits.data_ptr(), ldd, L, V, float(focal_gamma), _stream())
    return dlogits


# ------------------------------------------------------------------------------------------------
# Fine-tuning head (pooling + score head + task loss)
# ------------------------------------------------------------------------------------------------
def _ptr_array(tensors):
    import ctypes
    arr = (ctypes.c_void_p * len(tensors))(*[(t.data_ptr() if t is not None else None) for t in tensors])
    return arr, ctypes.cast(arr, ctypes.c_void_p)


def _int_array(vals):
    import ctypes
    arr = (ctypes.c_int * len(vals))(*vals)
    return arr, ctypes.cast(arr, ctypes.c_void_p)


class FtHeadState:
    """Buffers shared by ggpt_ft_head_fwd / ggpt_ft_head_bwd for one forward pass."""
    pass


def ft_head_fwd(hidden, in_ids, pad_id, weights, biases, *, act, drop_p=0.0, drop_seed=0, mode=-1, labels_i=None,
                labels_f=None, sample_wgt=None, err_flag=None):
    """hidden bf16 [N*S, d]; in_ids int64 [N,S] (any strides); weights / biases: lists of fp32 tensors (bias entries may
    be None).  Returns FtHeadState with .seq_idx, .pooled bf16 [N,d], .logits f32 [N,C], .loss f32 [2] = [loss, 1/den]."""
    _check(hidden, BF16, "ft_head hidden", 2)
    if in_ids.dtype != torch.int64 or in_ids.dim() != 2 or not in_ids.is_cuda:
        raise RuntimeError("ft_head: in_ids must be a CUDA int64 [N,S] tensor")
    N, S = in_ids.shape
    d = hidden.shape[1]
    if hidden.shape[0] != N * S:
        raise RuntimeError(f"ft_head: hidden has {hidden.shape[0]} rows, expected N*S = {N * S}")
    dims = [d] + [int(w.shape[0]) for w in weights]
    for l, w in enumerate(weights):
        _check(w, F32, f"ft_head weight {l}", 2)
        if w.shape[1] != dims[l] or not w.is_contiguous():
            raise RuntimeError(f"ft_head: weight {l} has shape {tuple(w.shape)}, expected [*, {dims[l]}] contiguous")
    dev = hidden.device
    st = FtHeadState()
    st.N, st.S, st.dims, st.max_dim, st.act, st.mode = N, S, dims, max(dims), int(act), int(mode)
    st.drop_p, st.drop_seed = float(drop_p), int(drop_seed)
    st.weights, st.biases = list(weights), list(biases)
    st.labels_i, st.labels_f, st.sample_wgt = labels_i, labels_f, sample_wgt
    st.seq_idx = torch.empty((N,), device=dev, dtype=torch.int32)
    st.pooled = torch.empty((N, d), device=dev, dtype=BF16)
    st.logits = torch.empty((N, dims[-1]), device=dev, dtype=F32)
    st.pre = torch.empty((len(weights), N, st.max_dim), device=dev, dtype=F32)
    st.row_loss = torch.empty((N,), device=dev, dtype=F32)
    st.row_den = torch.empty((N,), device=dev, dtype=F32)
    st.loss = torch.empty((2,), device=dev, dtype=F32)
    keep_w, pw = _ptr_array(st.weights)
    keep_b, pb = _ptr_array(st.biases)
    keep_d, pd = _int_array(dims)
    lib.ggpt_ft_head_fwd(hidden.data_ptr(), hidden.stride(0), in_ids.data_ptr(), in_ids.stride(0), in_ids.stride(1),
                         int(pad_id), N, S, len(weights), pd, pw, pb, st.act, st.drop_p, st.drop_seed, st.mode,
                         _ptr(labels_i), _ptr(labels_f), _ptr(sample_wgt), st.seq_idx.data_ptr(), st.pooled.data_ptr(),
                         st.logits.data_ptr(), st.pre.data_ptr(), st.max_dim, st.row_loss.data_ptr(), st.row_den.data_ptr(),
                         st.loss.data_ptr(), _ptr(err_flag), _stream())
    return st


def ft_head_bwd(st, gout, dweights, dbiases, dhidden):
    """Adds into dweights / dbiases (lists, entries may be None) and writes the pooled rows' gradient into the pre-zeroed
    dhidden bf16 [N*S, d]."""
    _check(gout, F32, "ft_head gout")
    _check(dhidden, BF16, "ft_head dhidden", 2)
    keep_w, pw = _ptr_array(st.weights)
    keep_dw, pdw = _ptr_array(dweights)
    keep_db, pdb = _ptr_array(dbiases)
    keep_d, pd = _int_array(st.dims)
    lib.ggpt_ft_head_bwd(st.N, st.S, len(st.weights), pd, pw, pdw, pdb, st.act, st.drop_p, st.drop_seed, st.mode,
                         _ptr(st.labels_i), _ptr(st.labels_f), _ptr(st.sample_wgt), st.seq_idx.data_ptr(),
                         st.logits.data_ptr(), st.pre.data_ptr(), st.max_dim, st.loss.data_ptr(), gout.data_ptr(),
                         dhidden.data_ptr(), dhidden.stride(0), _stream())


def ft_intra_fwd(hidden, cls_idx, labels, N, S, C, err_flag=None):
    """token_ce_intra: hidden bf16 [N*S,d], cls_idx int64 [N], labels int64 [N*S] | None.  Returns a state object with
    .logits f32 [N*S,C], .loss f32 [2] = [loss, 1/#labelled], .rn (inverse norms)."""
    _check(hidden, BF16, "ft_intra hidden", 2)
    _check(cls_idx, torch.int64, "ft_intra cls_idx", 1)
    if labels is not None:
        _check(labels, torch.int64, "ft_intra labels", 1)
    T, d = hidden.shape
    dev = hidden.device
    st = FtHeadState()
    st.hidden, st.cls_idx, st.labels, st.N, st.S, st.C, st.d = hidden, cls_idx, labels, N, S, C, d
    st.rn = torch.empty((T,), device=dev, dtype=F32)
    st.logits = torch.empty((T, C), device=dev, dtype=F32)
    st.row_loss = torch.empty((T,), device=dev, dtype=F32)
    st.row_den = torch.empty((T,), device=dev, dtype=F32)
    st.loss = torch.empty((2,), device=dev, dtype=F32)
    lib.ggpt_ft_intra_fwd(hidden.data_ptr(), hidden.stride(0), cls_idx.data_ptr(), _ptr(labels), st.rn.data_ptr(),
                          st.logits.data_ptr(), st.row_loss.data_ptr(), st.row_den.data_ptr(), st.loss.data_ptr(), N, S, C, d,
                          _ptr(err_flag), _stream())
    return st


def ft_intra_bwd(st, gout):
    T, d = st.hidden.shape
    dev = st.hidden.device
    dl = torch.empty((T, st.C), device=dev, dtype=F32)
    g = torch.empty((T, d), device=dev, dtype=F32)
    dh = torch.empty((T, d), device=dev, dtype=BF16)
    lib.ggpt_ft_intra_bwd(st.hidden.data_ptr(), st.hidden.stride(0), st.cls_idx.data_ptr(), st.labels.data_ptr(),
                          st.rn.data_ptr(), st.logits.data_ptr(), st.loss.data_ptr(), gout.data_ptr(), dl.data_ptr(),
                          g.data_ptr(), dh.data_ptr(), d, st.N, st.S, st.C, d, _stream())
    return dh


# ------------------------------------------------------------------------------------------------
# Generation
# ------------------------------------------------------------------------------------------------
def gen_sample(logits, V, *, temperature=0.0, top_k=None, top_p=None, u=None, conf_mode=0, want_probs=False):
    """logits f32 [R, ld>=V].  Returns (conf f32 [R], x0 int64 [R][, probs f32 [R,V]])."""
    _check(logits, F32, "gen_sample logits", 2)
    R = logits.shape[0]
    dev = logits.device
    if u is not None:
        _check(u, F32, "gen_sample u", 1)
        if u.numel() != R:
            raise RuntimeError("gen_sample: one uniform draw per row expected")
    x0 = torch.empty((R,), device=dev, dtype=torch.int64)
    conf = torch.empty((R,), device=dev, dtype=F32)
    probs = torch.empty((R, V), device=dev, dtype=F32) if want_probs else None
    lib.ggpt_gen_sample(logits.data_ptr(), logits.stride(0), R, V, float(temperature), int(top_k or 0),
                        float(top_p) if top_p is not None else 0.0, _ptr(u), int(conf_mode), x0.data_ptr(), conf.data_ptr(),
                        _ptr(probs), V, _stream())
    return (conf, x0, probs) if want_probs else (conf, x0)


def gen_unmask_origin(x, x0, u, p_transfer, mask_token):
    _check(x, torch.int64, "gen_unmask x")
    _check(x0, torch.int64, "gen_unmask x0")
    _check(u, F32, "gen_unmask u")
    if not (x.is_contiguous() and x0.is_contiguous() and u.is_contiguous()) or x0.numel() != x.numel() or u.numel() != x.numel():
        raise RuntimeError("gen_unmask_origin: x, x0, u must be contiguous and equally sized")
    lib.ggpt_gen_unmask_origin(x.data_ptr(), x0.data_ptr(), u.data_ptr(), float(p_transfer), int(mask_token), x.numel(),
                               _stream())
    return x


def gen_unmask_topk(x, x0, conf, k_per_sample, mask_token, *, gumbel_u=None, alg_temp=0.0):
    _check(x, torch.int64, "gen_unmask x", 2)
    _check(x0, torch.int64, "gen_unmask x0", 2)
    _check(conf, F32, "gen_unmask conf", 2)
    _check(k_per_sample, torch.int32, "gen_unmask k_per_sample", 1)
    B, P = x.shape
    if not (x.is_contiguous() and x0.is_contiguous() and conf.is_contiguous()) or x0.shape != x.shape or conf.shape != x.shape:
        raise RuntimeError("gen_unmask_topk: x, x0, conf must be contiguous [B,P]")
    if gumbel_u is not None:
        _check(gumbel_u, F32, "gen_unmask gumbel_u", 2)
    lib.ggpt_gen_unmask_topk(x.data_ptr(), x0.data_ptr(), conf.data_ptr(), _ptr(gumbel_u), float(alg_temp),
                             k_per_sample.data_ptr(), int(mask_token), B, P, _stream())
    return x


def sumsq(g, out):
    lib.ggpt_sumsq(g.data_ptr(), g.numel(), out.data_ptr(), _stream())


def adamw(p, p_bf16, g, m, v, *, lr, betas, eps, weight_decay, step, gnorm_sq=None, max_norm=0.0, grad_scale=1.0):
    lib.ggpt_adamw(p.data_ptr(), _ptr(p_bf16), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), float(lr),
                   float(betas[0]), float(betas[1]), float(eps), float(weight_decay), int(step), _ptr(gnorm_sq),
                   float(max_norm), float(grad_scale), _stream())


def cast_f32_bf16(src, dst):
    lib.ggpt_cast_f32_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream())


def cast_bf16_f32(src, dst):
    lib.ggpt_cast_bf16_f32(src.data_ptr(), dst.data_ptr(), src.numel(), _stream())
