"""Validation-only PRECISE forward (GGPT_PRECISE=1) — never used by bench.py, training or generation.

Why it exists: the fast path stores activations and weights in bf16, which puts its logits 3..6e-3 (relative Frobenius)
from the fp32 reference — the same distance the reference itself shows when run in bf16.  To show that this distance is
storage rounding and not an arithmetic / indexing error of similar size, this mode keeps every activation in fp32 and
sends every GEMM through the SAME tcgen05 kernel (`ggpt_gemm_bf16`) on split-bf16 operands (x = hi + lo, three partial
products folded into one GEMM over K' = 3K, see csrc/precise.cu).  Embedding gather, mask build, label compaction and
cross-entropy are the product kernels (they are fp32 / integer already); norms, RoPE, attention, GeGLU and residual adds
are plain fp32 SIMT kernels.  tests/test_precise_mode_gpu.py holds the result to <= 1e-3 against every fp32 golden.

Reference lines reproduced: HF:375-425 (LlamaModel.forward), HF:303-332 (decoder layer), utils_graphgpt.py:137-166
(LayerScale), modeling_pretrain.py:119-150,213-237, modeling_helpers.py:263-301.
"""
import os

import torch

from . import ops
from .lib import lib

BF16, F32 = torch.bfloat16, torch.float32


def enabled():
    return os.environ.get("GGPT_PRECISE") == "1"


def _s():
    return torch.cuda.current_stream().cuda_stream


def _split3(x, role):
    """fp32 [R,C] -> bf16 [R,3C]: role 0 = A operand (hi|lo|hi), role 1 = B operand (hi|hi|lo)."""
    if x.dtype != F32 or x.dim() != 2 or x.stride(1) != 1:
        raise RuntimeError(f"precise: expected a 2-D fp32 tensor, got {x.dtype} {tuple(x.shape)}")
    R, C = x.shape
    out = torch.empty((R, 3 * C), device=x.device, dtype=BF16)
    lib.ggpt_vp_split3(x.data_ptr(), x.stride(0), out.data_ptr(), 3 * C, R, C, role, _s())
    return out


def linear(x, w):
    """fp32 y[M,N] = x[M,K] @ w[N,K]^T through the tcgen05 GEMM kernel on split-bf16 operands."""
    return ops.gemm(_split3(x, 0), _split3(w, 1), out_dtype=F32)


def rmsnorm(x, w, eps):
    T, d = x.shape
    y = torch.empty_like(x)
    lib.ggpt_vp_rmsnorm_f32(x.data_ptr(), w.data_ptr(), y.data_ptr(), T, d, float(eps), _s())
    return y


def add(x, y, colscale=None, rowscale=None):
    T, d = x.shape
    out = torch.empty_like(x)
    lib.ggpt_vp_add_f32(x.data_ptr(), y.data_ptr(), y.stride(0), ops._ptr(colscale), ops._ptr(rowscale), out.data_ptr(), T, d,
                        _s())
    return out


def gather_rows(src, idx, n):
    d = src.shape[1]
    out = torch.empty((n, d), device=src.device, dtype=F32)
    if n > 0:
        lib.ggpt_vp_gather_rows_f32(src.data_ptr(), src.stride(0), idx.data_ptr(), out.data_ptr(), n, d, _s())
    return out


def _w2(fp, first_name, rows, cols):
    """fp32 [rows, cols] view over adjacent parameters of the flat master buffer (q|k|v, gate|up)."""
    off, _ = fp.offsets[first_name]
    return fp.flat[off:off + rows * cols].view(rows, cols)


def backbone_forward(hot, ids2d, N, S, attention_mask, position_ids, raw=None):
    """ids2d int64 [T,F] -> final-norm hidden states, fp32 [T,d]."""
    fp, cfg = hot.flat, hot.cfg
    dev = ids2d.device
    d, H, I = hot.d, hot.H, hot.I
    T = N * S
    gate = fp.w("stacked_feat_agg.weight") if "stacked_feat_agg.weight" in fp.offsets else None
    long_scale = getattr(cfg, "stack_method", None) == "long" and ids2d.shape[1] > 1
    err = torch.zeros((1,), device=dev, dtype=torch.int32)
    x = ops.embed_fwd(ids2d, fp.w("model.embed_tokens.weight"), gate, long_scale, err)
    pos, cos, sin = hot._rope_inputs(position_ids, N, S, dev)
    mask = ops.attn_mask_build(attention_mask, N, S, cfg.causal_attention, dev)
    if raw is not None:
        raw2d, rlabels, fchk = raw
        if raw2d.dim() != 2:
            raise NotImplementedError("precise mode: [N,S,S,E] edge embeddings have no validation-mode variant")
        E = raw2d.shape[1]
        hr = torch.empty((T, E), device=dev, dtype=F32)
        mask_tok = fp.w("emb_mask_token") if rlabels is not None else None
        lib.ggpt_vp_raw_embed_f32(raw2d.data_ptr(), ops._ptr(rlabels), rlabels.stride(0) if rlabels is not None else 0,
                                  int(fchk), ops._ptr(mask_tok), fp.w("embed_layernorm.weight").data_ptr(), hr.data_ptr(), T,
                                  E, float(hot.eps), _s())
        x = add(x, linear(hr, fp.w("embed_proj.weight")))
    for i in range(hot.L):
        p = f"model.layers.{i}."
        h = rmsnorm(x, fp.w(p + "input_layernorm.weight"), hot.eps)
        qkv = linear(h, _w2(fp, p + "self_attn.q_proj.weight", 3 * d, d))
        lib.ggpt_vp_rope_f32(qkv.data_ptr(), qkv.stride(0), pos.data_ptr(), cos.data_ptr(), sin.data_ptr(), T, 2 * d, _s())
        a = torch.empty((T, d), device=dev, dtype=F32)
        lib.ggpt_vp_attn_f32(qkv.data_ptr(), qkv.stride(0), 0, d, 2 * d, mask.bits.data_ptr(), a.data_ptr(), d, N, S, H, _s())
        y = linear(a, fp.w(p + "self_attn.o_proj.weight"))
        x = add(x, y, fp.w(p + "lambda_1") if hot.layer_scale else None)
        h = rmsnorm(x, fp.w(p + "post_attention_layernorm.weight"), hot.eps)
        gu = linear(h, _w2(fp, p + "mlp.gate_proj.weight", 2 * I, d))
        act = torch.empty((T, I), device=dev, dtype=F32)
        lib.ggpt_vp_geglu_f32(gu.data_ptr(), gu.stride(0), act.data_ptr(), T, I, _s())
        y = linear(act, fp.w(p + "mlp.down_proj.weight"))
        x = add(x, y, fp.w(p + "lambda_2") if hot.layer_scale else None)
    return rmsnorm(x, fp.w("model.norm.weight"), hot.eps)


def head_all_entries(hot, hf):
    """labels=None (inference / generation): logits for every (n,s,f) entry, modeling_helpers.py:284-292."""
    fp = hot.flat
    proj = hf
    if "n_token_proj.weight" in fp.offsets:
        proj = linear(hf, fp.w("n_token_proj.weight")).reshape(-1, hot.d)
    return linear(proj, fp.w("lm_head.weight"))


def labelled_head(hot, hf, labels2d, N, S, ent_wgt_fn, loss_mode, head):
    """Same selection / loss as engine.PretrainHeadFn.forward with fp32 rows and split-operand GEMMs.  Returns
    (loss 0-d fp32, logits fp32 [L,V])."""
    fp = hot.flat
    head_w, proj_w, V = head
    V = hot.V if V is None else V
    hi = ops.head_compact(labels2d)
    d, F_ = hot.d, hi.F
    M, L = hi.sync_counts()
    if M == 0:
        return torch.full((), float("nan"), device=hf.device, dtype=F32), torch.empty((0, V), device=hf.device, dtype=F32)
    hsel = gather_rows(hf, hi.sel_rows, M)
    if proj_w is not None and proj_w in fp.offsets:
        proj = linear(hsel, fp.w(proj_w))
        hl = gather_rows(proj.reshape(M * F_, d), hi.ent_src, L)
    else:
        hl = hsel
    logits = linear(hl, fp.w(head_w))
    wgt = ent_wgt_fn(hi, L) if ent_wgt_fn is not None else None
    focal = float(getattr(hot.cfg, "focal_gamma", 0.0) or 0.0) if wgt is None else 0.0
    _, _, sums = ops.ce_fwd(logits, hi.ent_label, V, wgt, err_flag=None, focal_gamma=focal)
    if loss_mode == "mean":
        ls = ops.ce_finalize(sums, hi.counts.data_ptr() + 4, 0)
    else:
        ls = ops.ce_finalize(sums, 0, 2, float(N * S * hot.cfg.next_n_token))
    return ls[0], logits
