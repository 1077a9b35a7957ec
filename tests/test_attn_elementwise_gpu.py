"""GPU parity of the attention forward kernel and the HBM-bound kernels against plain torch fp32 references of
the same op (the oracle's formulas).  Tolerances are stated per test."""
import math

import numpy as np
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_attention(qkv, am, causal, N, S, H):
    """oracle.attention on bf16-rounded q,k,v (fp32 math); returns out [N*S, H*64], lse [N,H,S] and the keep mask."""
    d = H * 64
    x = qkv.float().view(N, S, 3, H, 64)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    if am is None:
        keep = torch.ones((N, 1, S, S), dtype=torch.bool, device=qkv.device)
    elif am.dim() == 2:
        keep = am[:, None, None, :].bool().expand(N, 1, S, S)
    else:
        keep = am[:, None].bool()
    if causal:
        keep = keep & torch.ones((S, S), dtype=torch.bool, device=qkv.device).tril()[None, None]
    s = (q @ k.transpose(2, 3)) * 0.125
    s = s.masked_fill(~keep, float("-inf"))
    lse = torch.logsumexp(s, -1)
    p = torch.exp(s - lse[..., None])
    p = torch.nan_to_num(p, nan=0.0)
    o = (p @ v).transpose(1, 2).reshape(N * S, d)
    return o, lse, keep


CASES = [
    # N, S, H, mask kind, causal
    (2, 128, 1, "none", False),
    (2, 256, 2, "none", False),
    (3, 40, 2, "pad", False),
    (2, 200, 3, "pad", False),
    (2, 384, 2, "packed", False),
    (2, 1024, 2, "packed", False),
    (2, 300, 2, "pad", True),
    (1, 1024, 12, "none", False),
    (2, 512, 2, "random", False),
    (1, 2048, 2, "pad", True),        # C5: causal NTP at seq 2048
    (1, 4096, 1, "none", False),      # C4: citation2 sequences of 4096
    (1, 4096, 1, "packed", False),
]


@pytest.mark.parametrize("N,S,H,kind,causal", CASES)
def test_attn_fwd(N, S, H, kind, causal):
    from graphgpt_b200 import ops
    g = torch.Generator().manual_seed(S * 7 + H)
    d = H * 64
    qkv = (torch.randn((N * S, 3 * d), generator=g) * 1.5).to(torch.bfloat16).cuda()
    if kind == "none":
        am = None
    elif kind == "pad":
        lens = torch.randint(S // 3, S + 1, (N,), generator=g)
        am = (torch.arange(S)[None, :] < lens[:, None]).long().cuda()
    elif kind == "packed":
        am = torch.zeros((N, S, S), dtype=torch.long)
        for n in range(N):
            o = 0
            while o < S:
                L = int(torch.randint(8, 40, (1,), generator=g))
                L = min(L, S - o)
                am[n, o:o + L, o:o + L] = 1
                o += L
        am = am.cuda()
    else:
        am = (torch.rand((N, S, S), generator=g) < 0.3).long()
        am[:, torch.arange(S), torch.arange(S)] = 1
        am = am.cuda()
    mask = ops.attn_mask_build(am, N, S, causal, qkv.device)
    out, lse = ops.attn_fwd(qkv, mask, H)
    torch.cuda.synchronize()
    ref_o, ref_lse, keep = _ref_attention(qkv, am, causal, N, S, H)
    valid = keep.any(-1)[:, 0]                                    # [N,S] rows with at least one visible key
    vrow = valid.reshape(N * S)
    err = (out.float() - ref_o)[vrow].abs().max().item()
    scale = ref_o[vrow].abs().max().item()
    # P is rounded to bf16 before PV (2^-9 relative per term) and the output is bf16: 1e-2 of the output scale
    assert err <= 1e-2 * scale, f"attn out err {err} scale {scale}"
    lerr = (lse - ref_lse)[valid[:, None, :].expand(N, H, S)].abs().max().item()
    assert lerr <= 2e-3, f"lse err {lerr}"
    # fully masked rows produce exact zeros
    if (~vrow).any():
        assert out[~vrow].abs().max().item() == 0.0


def test_embed_fwd_bwd():
    from graphgpt_b200 import ops
    g = torch.Generator().manual_seed(3)
    T, F_, d, V = 1000, 13, 768, 756
    ids = torch.randint(0, V, (T, F_), generator=g).cuda()
    ids[::7] = 0
    table = torch.randn((V, d), generator=g).cuda()
    gate = torch.randn((F_, d), generator=g).cuda()
    x = ops.embed_fwd(ids, table)
    ref = table[ids].sum(1)
    assert (x - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()     # fp32 sum order only
    xg = ops.embed_fwd(ids, table, gate)
    refg = torch.einsum("tfd,fd->td", table[ids], gate)
    assert (xg - refg).abs().max().item() <= 1e-5 * refg.abs().max().item()
    dx = torch.randn((T, d), generator=g).cuda()
    dtable = torch.zeros_like(table)
    ops.embed_bwd(ids, dx, None, None, dtable, None)
    t2 = table.clone().requires_grad_(True)
    torch.nn.functional.embedding(ids, t2, padding_idx=0).sum(1).backward(dx)
    assert (dtable - t2.grad).abs().max().item() <= 2e-4 * t2.grad.abs().max().item()   # atomic fp32 sum order
    assert dtable[0].abs().max().item() == 0.0
    dtable.zero_()
    dgate = torch.zeros_like(gate)
    ops.embed_bwd(ids, dx, table, gate, dtable, dgate)
    t3 = table.clone().requires_grad_(True)
    g3 = gate.clone().requires_grad_(True)
    torch.einsum("tfd,fd->td", torch.nn.functional.embedding(ids, t3, padding_idx=0), g3).backward(dx)
    assert (dtable - t3.grad).abs().max().item() <= 2e-4 * t3.grad.abs().max().item()
    assert (dgate - g3.grad).abs().max().item() <= 2e-4 * g3.grad.abs().max().item()


@pytest.mark.parametrize("T,d", [(1000, 768), (300, 64), (200, 128), (100, 1024), (1500, 256), (77, 512), (5, 768)])
def test_rmsnorm_fwd_bwd(T, d):
    from graphgpt_b200 import ops
    g = torch.Generator().manual_seed(d)
    x = torch.randn((T, d), generator=g).cuda() * 2
    w = (1 + 0.1 * torch.randn((d,), generator=g)).cuda()
    y, rstd = ops.rmsnorm_fwd(x, w, 1e-6)
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    ref = wr * (xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + 1e-6))
    assert (y.float() - ref).abs().max().item() <= 5e-3 * ref.abs().max().item()        # one bf16 rounding
    dy = torch.randn((T, d), generator=g).to(torch.bfloat16).cuda()
    dres = torch.randn((T, d), generator=g).cuda()
    ref.backward(dy.float())
    dw = torch.zeros_like(w)
    dx, dxb = ops.rmsnorm_bwd(dy, x, rstd, w, dres, dw)
    torch.cuda.synchronize()
    assert (dx - (dres + xr.grad)).abs().max().item() <= 1e-4 * (dres + xr.grad).abs().max().item()
    assert (dw - wr.grad).abs().max().item() <= 1e-4 * wr.grad.abs().max().item()
    assert (dxb.float() - dx).abs().max().item() <= 5e-3 * dx.abs().max().item()


@pytest.mark.parametrize("T,d,scaled", [(1000, 768, False), (300, 64, True), (129, 1024, True)])
def test_add_rmsnorm_fwd(T, d, scaled):
    """Fused residual add (+ LayerScale column scale, DropPath row scale) + RMSNorm: HF:325,331 then HF:59-64."""
    from graphgpt_b200 import ops
    g = torch.Generator().manual_seed(T + d)
    x = torch.randn((T, d), generator=g).cuda() * 2
    y = torch.randn((T, d), generator=g).to(torch.bfloat16).cuda()
    w = (1 + 0.1 * torch.randn((d,), generator=g)).cuda()
    cs = (0.5 + torch.rand((d,), generator=g)).cuda() if scaled else None
    rs = (torch.rand((T,), generator=g) > 0.2).float().cuda() / 0.8 if scaled else None
    x_out, h, rstd = ops.add_rmsnorm_fwd(x, y, w, 1e-6, colscale=cs, rowscale=rs)
    yy = y.float()
    if scaled:
        yy = yy * cs[None, :] * rs[:, None]
    ref_x = x + yy
    ref_rstd = torch.rsqrt(ref_x.pow(2).mean(-1) + 1e-6)
    ref_h = w * (ref_x * ref_rstd[:, None])
    torch.cuda.synchronize()
    assert (x_out - ref_x).abs().max().item() <= 1e-5 * ref_x.abs().max().item()
    assert (rstd - ref_rstd).abs().max().item() <= 1e-5 * ref_rstd.abs().max().item()
    assert (h.float() - ref_h).abs().max().item() <= 5e-3 * ref_h.abs().max().item()        # one bf16 rounding
    x2, h2, r2 = ops.add_rmsnorm_fwd(x, y, w, 1e-6, colscale=cs, rowscale=rs, want_rstd=False, want_h=False)
    assert h2 is None and r2 is None and torch.equal(x2, x_out)


def test_geglu_bwd():
    from graphgpt_b200 import ops
    g = torch.Generator().manual_seed(5)
    T, I = 500, 3072
    gu = torch.randn((T, 2 * I), generator=g).to(torch.bfloat16).cuda()
    dact = torch.randn((T, I), generator=g).to(torch.bfloat16).cuda()
    gg = gu.float().clone().requires_grad_(True)
    (torch.nn.functional.gelu(gg[:, :I]) * gg[:, I:]).backward(dact.float())
    # geglu_bwd multiplies dact with the factors saved by the forward epilogue: gf = [u * gelu'(g) | gelu(g)]
    g_ = gu.float()[:, :I].clone().requires_grad_(True)
    gelu_g = torch.nn.functional.gelu(g_)
    gelu_g.sum().backward()
    gf = torch.cat([gu.float()[:, I:] * g_.grad, gelu_g.detach()], dim=1).to(torch.bfloat16)
    dgu = ops.geglu_bwd(dact, gf)
    assert (dgu.float() - gg.grad).abs().max().item() <= 1e-2 * gg.grad.abs().max().item()


@pytest.mark.parametrize("T,F_", [(1000, 13), (70000, 13), (513, 1), (256, 4)])
def test_head_compact_and_gather(T, F_):
    from graphgpt_b200 import ops
    g = torch.Generator().manual_seed(T)
    labels = torch.where(torch.rand((T, F_), generator=g) < 0.3, torch.randint(0, 700, (T, F_), generator=g), -100)
    labels[T // 3: T // 2] = -100
    labels = labels.cuda()
    hi = ops.head_compact(labels)
    M, L = hi.sync_counts()
    mask = labels != -100
    mask_m = mask.any(-1)
    assert M == int(mask_m.sum()) and L == int(mask.sum())
    assert torch.equal(hi.sel_rows[:M].long(), mask_m.nonzero()[:, 0])
    assert torch.equal(hi.ent_label[:L].long(), labels[mask_m][mask[mask_m]])
    src_ref = mask[mask_m].reshape(-1).nonzero()[:, 0]
    assert torch.equal(hi.ent_src[:L].long(), src_ref)
    assert torch.equal(hi.ent_tok[:L].long(), mask.nonzero()[:, 0])
    h = torch.randn((T, 64), generator=g).to(torch.bfloat16).cuda()
    sel = ops.gather_rows(h, hi.sel_rows, M)
    assert torch.equal(sel, h[mask_m])
    back = torch.zeros_like(h)
    ops.scatter_rows(sel, hi.sel_rows, back, M)
    assert torch.equal(back[mask_m], h[mask_m]) and back[~mask_m].abs().max().item() == 0


@pytest.mark.parametrize("L,V", [(1000, 756), (333, 300), (64, 41244)])
def test_ce_fwd_bwd(L, V):
    from graphgpt_b200 import ops
    g = torch.Generator().manual_seed(V)
    ld = (V + 7) // 8 * 8
    logits = (torch.randn((L, ld), generator=g) * 3).cuda()
    labels = torch.randint(0, V, (L,), generator=g, dtype=torch.int32).cuda()
    wgt = torch.rand((L,), generator=g).cuda()
    lg = logits[:, :V].clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lg, labels.long())
    row_lse, row_loss, sums = ops.ce_fwd(logits, labels, V, want_row_loss=True)
    cnt = torch.tensor([L], dtype=torch.int32).cuda()
    ls = ops.ce_finalize(sums, cnt.data_ptr(), 0)
    assert abs(ls[0].item() - ref.item()) <= 1e-5 * abs(ref.item())
    ref.backward()
    gout = torch.tensor([1.0]).cuda()
    dl = ops.ce_bwd(logits, labels, V, row_lse, ls.data_ptr() + 4, gout)
    assert (dl[:, :V].float() - lg.grad).abs().max().item() <= 5e-3 * lg.grad.abs().max().item()
    assert dl[:, V:].abs().max().item() == 0 if ld > V else True
    # weighted (dLM): sum(ce * w) / fixed
    row_lse, _, sums = ops.ce_fwd(logits, labels, V, wgt)
    ls = ops.ce_finalize(sums, 0, 2, 1234.0)
    refw = (torch.nn.functional.cross_entropy(logits[:, :V], labels.long(), reduction="none") * wgt).sum() / 1234.0
    assert abs(ls[0].item() - refw.item()) <= 1e-5 * abs(refw.item())


def test_adamw_and_sumsq():
    from graphgpt_b200 import ops
    g = torch.Generator().manual_seed(11)
    n = 100003
    p = torch.randn((n,), generator=g).cuda()
    gr = torch.randn((n,), generator=g).cuda() * 3
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-3, betas=(0.9, 0.95), eps=1e-6, weight_decay=0.1)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    pb = torch.empty((n,), dtype=torch.bfloat16).cuda()
    for step in range(1, 4):
        pr.grad = gr.clone()
        torch.nn.utils.clip_grad_norm_([pr], 1.0)
        opt.step()
        ss = torch.zeros((1,), dtype=torch.float64).cuda()
        ops.sumsq(gr, ss)
        assert abs(ss.item() - gr.double().pow(2).sum().item()) <= 1e-6 * ss.item()
        ops.adamw(p, pb, gr, m, v, lr=1e-3, betas=(0.9, 0.95), eps=1e-6, weight_decay=0.1, step=step, gnorm_sq=ss,
                  max_norm=1.0)
        assert (p - pr.detach()).abs().max().item() <= 2e-6
    assert (pb.float() - p).abs().max().item() <= 4e-3 * p.abs().max().item()


BWD_CASES = [
    (2, 128, 1, "none", False, False),
    (2, 256, 2, "none", False, True),
    (3, 40, 2, "pad", False, True),
    (2, 200, 3, "pad", True, True),
    (2, 384, 2, "packed", False, True),
    (1, 1024, 4, "none", False, False),
    (2, 512, 2, "random", False, True),
    (1, 2048, 2, "pad", True, True),       # C5
    (1, 4096, 1, "none", False, True),     # C4
]


@pytest.mark.parametrize("N,S,H,kind,causal,rope", BWD_CASES)
def test_attn_bwd(N, S, H, kind, causal, rope):
    """dq/dk/dv against torch autograd through (RoPE ->) masked softmax attention in fp32."""
    from graphgpt_b200 import ops
    g = torch.Generator().manual_seed(S * 3 + H)
    d = H * 64
    dev = torch.device("cuda")
    pre = (torch.randn((N * S, 3 * d), generator=g) * 1.2).to(torch.bfloat16).float().to(dev).requires_grad_(True)
    if kind == "none":
        am = None
    elif kind == "pad":
        lens = torch.randint(S // 3, S + 1, (N,), generator=g)
        am = (torch.arange(S)[None, :] < lens[:, None]).long().to(dev)
    elif kind == "packed":
        am = torch.zeros((N, S, S), dtype=torch.long)
        for n in range(N):
            o = 0
            while o < S:
                L = min(int(torch.randint(8, 40, (1,), generator=g)), S - o)
                am[n, o:o + L, o:o + L] = 1
                o += L
        am = am.to(dev)
    else:
        am = (torch.rand((N, S, S), generator=g) < 0.3).long()
        am[:, torch.arange(S), torch.arange(S)] = 1
        am = am.to(dev)
    max_pos = 4096
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, 64, 2, dtype=torch.int64).float() / 64))
    freqs = torch.arange(max_pos).float()[:, None] * inv_freq[None, :]
    cos_tab, sin_tab = freqs.cos().to(dev), freqs.sin().to(dev)
    if rope:
        pos = torch.randint(0, max_pos, (N * S,), generator=g, dtype=torch.int32).to(dev)
    else:
        pos = torch.zeros((N * S,), dtype=torch.int32, device=dev)
    # reference: rope(q,k) then attention, all fp32, autograd
    x = pre.view(N * S, 3, H, 64)
    cos = torch.cat([cos_tab, cos_tab], -1)[pos.long()][:, None, :]
    sin = torch.cat([sin_tab, sin_tab], -1)[pos.long()][:, None, :]

    def rot(t):
        return t * cos + torch.cat([-t[..., 32:], t[..., :32]], -1) * sin

    q, k, v = rot(x[:, 0]), rot(x[:, 1]), x[:, 2]
    qkv_rot = torch.stack([q, k, v], 1).reshape(N * S, 3 * d)
    qkv_b = qkv_rot.detach().to(torch.bfloat16)
    qkv_ref = qkv_b.float().requires_grad_(True)       # reference differentiates at the bf16-rounded rotated point
    o_ref, lse_ref, keep = _ref_attention(qkv_ref, am, causal, N, S, H)
    valid = keep.any(-1)[:, 0].reshape(N * S)
    dout = (torch.randn((N * S, d), generator=g)).to(torch.bfloat16).to(dev)
    dout = dout * valid[:, None]                       # padded query rows carry no gradient in the model
    o_ref.backward(dout.float())
    # chain through the rotation: d pre = rot^T(d rotated)
    gr = qkv_ref.grad.view(N * S, 3, H, 64)

    def unrot(t):
        return t * cos + torch.cat([t[..., 32:], -t[..., :32]], -1) * sin

    ref = torch.stack([unrot(gr[:, 0]), unrot(gr[:, 1]), gr[:, 2]], 1).reshape(N * S, 3 * d)

    mask = ops.attn_mask_build(am, N, S, causal, dev)
    out, lse, out_lo = ops.attn_fwd(qkv_b, mask, H, want_lo=True)
    dqkv = ops.attn_bwd(dout, qkv_b, out, lse, mask, H, pos, cos_tab, sin_tab, out_lo=out_lo)
    torch.cuda.synchronize()
    for name, sl in (("dq", slice(0, d)), ("dk", slice(d, 2 * d)), ("dv", slice(2 * d, 3 * d))):
        a, b = dqkv[:, sl].float(), ref[:, sl]
        err = (a - b).abs().max().item()
        scale = b.abs().max().item()
        rel_f = ((a - b).norm() / (b.norm() + 1e-20)).item()
        # P and dS are rounded to bf16 before the gradient MMAs and outputs are bf16
        assert err <= 2e-2 * scale and rel_f <= 1e-2, f"{name}: max err {err} (scale {scale}), relF {rel_f}"


def test_attn_tile_plan_packed_and_dense():
    """Packed block-diagonal masks get row tiles cut on segment boundaries (each query tile sees only itself);
    masks without interior block boundaries keep the uniform 128 grid."""
    from graphgpt_b200 import ops, synth
    b = synth.make_batch(3, 1024, layout="packed", seed=9, return_segments=True)
    am = torch.from_numpy(b["attention_mask"]).cuda()
    m = ops.attn_mask_build(am, 3, 1024, False, am.device)
    torch.cuda.synchronize()
    for n in range(3):
        nt = int(m.n_tiles[n])
        ts = m.tile_start[n, : nt + 1].tolist()
        assert ts[0] == 0 and ts[-1] == 1024 and all(0 < b_ - a_ <= 128 for a_, b_ in zip(ts, ts[1:]))
        bounds = set(np.cumsum([0] + b["segment_lens"][n]).tolist())
        assert all(t in bounds for t in ts), "tile cuts must sit on segment boundaries"
        assert nt <= 11
        cls = m.cls[n, :nt, :nt].cpu()
        assert (cls.diagonal() == 2).all() and (cls - torch.diag(cls.diagonal())).abs().sum() == 0
    # dense / key-padding / causal: uniform grid
    am2 = torch.ones((2, 1024), dtype=torch.long).cuda()
    am2[1, 700:] = 0
    for causal in (False, True):
        m2 = ops.attn_mask_build(am2, 2, 1024, causal, am2.device)
        assert m2.n_tiles.tolist() == [8, 8]
        assert m2.tile_start[0, :9].tolist() == list(range(0, 1025, 128))
        cls = m2.cls[0, :8, :8].cpu()
        if causal:
            assert (cls.triu(1) == 0).all() and (cls.diagonal() == 2).all() and (cls.tril(-1)[cls.tril(-1) > 0] == 1).all()
        else:
            assert (cls == 1).all()
            cls1 = m2.cls[1, :8, :8].cpu()
            assert (cls1[:, 6:] == 0).all() and (cls1[:, :5] == 1).all() and (cls1[:, 5] == 2).all()


@pytest.mark.parametrize("S,H", [(1024, 3), (384, 2), (200, 1)])
def test_attn_diag_path_matches_general_path(S, H, monkeypatch):
    """Packed batches run the persistent isolated-diagonal kernels; they must agree with the general tile-loop
    kernels (forced with GGPT_ATTN_NO_DIAG) to bf16 round-off, forward and backward."""
    from graphgpt_b200 import ops, synth
    N = 3
    b = synth.make_batch(N, S, layout="packed", seed=S)
    am = torch.from_numpy(b["attention_mask"]).cuda()
    g = torch.Generator().manual_seed(S + H)
    d = H * 64
    qkv = (torch.randn((N * S, 3 * d), generator=g) * 1.2).to(torch.bfloat16).cuda()
    dout = torch.randn((N * S, d), generator=g).to(torch.bfloat16).cuda()
    pos = torch.arange(S, dtype=torch.int32).repeat(N).cuda()
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, 64, 2, dtype=torch.int64).float() / 64))
    freqs = torch.arange(S).float()[:, None] * inv_freq[None, :]
    cos_tab, sin_tab = freqs.cos().cuda(), freqs.sin().cuda()
    res = {}
    for mode in ("diag", "general"):
        if mode == "general":
            monkeypatch.setenv("GGPT_ATTN_NO_DIAG", "1")
        else:
            monkeypatch.delenv("GGPT_ATTN_NO_DIAG", raising=False)
        mask = ops.attn_mask_build(am, N, S, False, qkv.device)
        assert mask.use_diag == (mode == "diag")
        if mode == "diag":
            assert int(mask.iso_count[0]) == int(mask.n_tiles.sum()) == int(mask.iso_count[1])      # every tile of a packed batch is isolated
        out, lse, out_lo = ops.attn_fwd(qkv, mask, H, want_lo=True)
        dqkv = ops.attn_bwd(dout, qkv, out, lse, mask, H, pos, cos_tab, sin_tab, out_lo=out_lo)
        torch.cuda.synchronize()
        res[mode] = (out.float(), lse, dqkv.float())
    for a, b_, name, tol in zip(res["diag"], res["general"], ("out", "lse", "dqkv"), (4e-3, 1e-5, 8e-3)):
        err = (a - b_).abs().max().item()
        scale = b_.abs().max().item()
        assert err <= tol * scale, f"{name}: diag vs general max err {err} (scale {scale})"


@pytest.mark.parametrize("kind,S", [("packed", 384), ("none", 256)])
def test_attn_dropout_consistent_forward_backward(kind, S):
    """Attention dropout (HF:217): the kernels regenerate the keep mask from (seed, n, h, q, k).  Recover the mask
    from a forward pass with V = identity-like probes, then check forward output and dq/dk/dv against torch autograd
    of softmax -> (mask / (1-p)) -> PV with that same mask; also check the drop rate and seed sensitivity."""
    from graphgpt_b200 import ops, synth
    N, H, p_drop, seed = 2, 2, 0.25, 123456789
    d = H * 64
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(S)
    if kind == "packed":
        am = torch.from_numpy(synth.make_batch(N, S, layout="packed", seed=5)["attention_mask"]).to(dev)
    else:
        am = None
    mask = ops.attn_mask_build(am, N, S, False, dev)
    # ---- recover the keep mask column block by column block: uniform scores (q = 0) and one-hot V rows
    keep = torch.zeros((N, H, S, S), device=dev)
    visible = torch.ones((N, 1, S, S), dtype=torch.bool, device=dev) if am is None else am[:, None].bool()
    nvis = visible.sum(-1, keepdim=True).float()
    for kb in range(0, S, 64):
        qkv = torch.zeros((N, S, 3, H, 64), device=dev)
        for j in range(min(64, S - kb)):
            qkv[:, kb + j, 2, :, j] = 1.0
        out, _ = ops.attn_fwd(qkv.reshape(N * S, 3 * d).to(torch.bfloat16), mask, H, dropout_p=p_drop, seed=seed)
        o = out.float().view(N, S, H, 64).permute(0, 2, 1, 3)               # [N,H,S,64] = P_drop[:, :, :, kb:kb+64]
        keep[:, :, :, kb:kb + 64] = (o[..., : min(64, S - kb)] * nvis > 0.5).float()
    keep = keep * visible
    rate = 1.0 - keep.sum().item() / visible.expand(N, H, S, S).sum().item()
    assert abs(rate - p_drop) < 0.01, f"drop rate {rate}"
    # ---- forward / backward parity with the recovered mask
    qkv_b = (torch.randn((N * S, 3 * d), generator=g) * 1.2).to(torch.bfloat16).to(dev)
    dout = torch.randn((N * S, d), generator=g).to(torch.bfloat16).to(dev)
    x = qkv_b.float().requires_grad_(True)
    xx = x.view(N, S, 3, H, 64)
    q, k, v = (xx[:, :, i].transpose(1, 2) for i in range(3))
    s_ = (q @ k.transpose(2, 3)) * 0.125
    s_ = s_.masked_fill(~visible, float("-inf"))
    pr = torch.softmax(s_, -1) * keep / (1.0 - p_drop)
    ref_o = (pr @ v).transpose(1, 2).reshape(N * S, d)
    ref_o.backward(dout.float())
    out, lse, out_lo = ops.attn_fwd(qkv_b, mask, H, want_lo=True, dropout_p=p_drop, seed=seed)
    pos = torch.zeros((N * S,), dtype=torch.int32, device=dev)
    cos_tab = torch.ones((1, 32), device=dev)
    sin_tab = torch.zeros((1, 32), device=dev)
    dqkv = ops.attn_bwd(dout, qkv_b, out, lse, mask, H, pos, cos_tab, sin_tab, out_lo=out_lo, dropout_p=p_drop, seed=seed)
    torch.cuda.synchronize()
    assert (out.float() - ref_o).abs().max().item() <= 1.5e-2 * ref_o.abs().max().item()
    for name, sl in (("dq", slice(0, d)), ("dk", slice(d, 2 * d)), ("dv", slice(2 * d, 3 * d))):
        a, b_ = dqkv[:, sl].float(), x.grad[:, sl]
        rel = ((a - b_).norm() / b_.norm()).item()
        assert rel <= 2e-2, f"{name} relF {rel}"
    out2, _ = ops.attn_fwd(qkv_b, mask, H, dropout_p=p_drop, seed=seed + 1)
    assert (out2.float() - out.float()).abs().max().item() > 1e-2          # a different seed gives a different mask
    out3, _ = ops.attn_fwd(qkv_b, mask, H, dropout_p=p_drop, seed=seed)
    assert torch.equal(out3, out)                                          # same seed: bitwise reproducible


def test_smtp_mask_2d_matches_reference_golden():
    """ggpt_smtp_mask_2d (config.smtp_inside) — bit-exact against the reference's outputs for the stored draws."""
    import os
    from graphgpt_b200 import ops
    recs = torch.load(os.path.join(os.path.dirname(__file__), "golden", "aux", "smtp_inside.pt"))
    for r in recs:
        err = torch.zeros((1,), dtype=torch.int32, device="cuda")
        out, lab = ops.smtp_mask_2d(r["ids"].cuda(), r["F"], r["mr"].cuda(), r["u_node"].cuda().contiguous(), r["power"],
                                    err_flag=err)
        assert torch.equal(out.cpu(), r["out_ids"]) and torch.equal(lab.cpu(), r["labels"])
        assert err.item() == 0
    bad = recs[0]["ids"].clone()
    bad[0, 0, recs[0]["F"] + 2] = 10 ** 6                     # node index out of range -> flagged, not a wild read
    err = torch.zeros((1,), dtype=torch.int32, device="cuda")
    ops.smtp_mask_2d(bad.cuda(), recs[0]["F"], recs[0]["mr"].cuda(), recs[0]["u_node"].cuda().contiguous(), 1.0, err_flag=err)
    assert err.item() == 3


def test_smtp_inside_model_forward_matches_oracle():
    """config.smtp_inside: the model masks inside forward() from torch's CUDA generator in the reference's draw order;
    re-creating the draws with the same seed gives the oracle's inputs / labels, and the loss must match."""
    from graphgpt_b200 import GraphGPTConfig, GraphGPTPretrainBase
    from oracle import graphgpt_oracle as oracle
    r = torch.load(os.path.join(os.path.dirname(__file__), "golden", "aux", "smtp_inside.pt"))[0]
    F_ = r["F"]
    cfgd = dict(vocab_size=756, hidden_size=64, intermediate_size=256, num_hidden_layers=2, num_attention_heads=1,
                num_key_value_heads=1, head_dim=64, hidden_act="gelu", max_position_embeddings=256, rms_norm_eps=1e-6,
                rope_theta=10000.0, pad_token_id=0, causal_attention=False, stacked_feat=F_, stack_method="short",
                stacked_feat_agg_method="sum", next_n_token=F_, use_cache=False, smtp_inside=True, smtp_power=1.0)
    sd = oracle.init_state_dict(cfgd, seed=9)
    model = GraphGPTPretrainBase(GraphGPTConfig(**cfgd))
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    ids = r["ids"]
    N, S, _ = ids.shape
    am = (ids[:, :, 0] != 0).long()
    torch.manual_seed(77)
    out = model(input_ids=ids.cuda(), attention_mask=am.cuda())
    torch.manual_seed(77)
    torch.rand((N, 1, 1), device="cuda")
    mr = torch.rand((N, 1, 1), device="cuda").view(-1).cpu()
    u = torch.rand((N, S, F_), device="cuda").cpu()
    ids_m, labels = oracle.smtp_mask_2d(ids[:, :, :F_], ids[:, :, F_ + 2], mr, u, 1.0)
    ref = oracle.pretrain_forward(sd, cfgd, ids_m, am, labels)
    assert out.head1_logits.shape == ref["logits"].shape
    assert abs(out.head1_loss.item() - ref["loss"].item()) <= 1e-3 * abs(ref["loss"].item())
