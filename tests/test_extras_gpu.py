"""GPU parity of the side branches of the hot path: raw-embedding input kernels, element dropout (embed_pdrop /
mlp_pdrop) and their model-level wiring.  Dropout masks are pure functions of (seed, element index), so the test reads
the exact factors through ggpt_dropout_scale_f32 and feeds them to the CPU oracle: training-mode parity is checked
value for value, not statistically (tolerances as in test_model_gpu.py: loss 1e-3, logits 1e-2, gradients 3e-2)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _relf(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-30)).item()


@pytest.mark.parametrize("p", [0.1, 0.5])
def test_dropout_factors_and_inplace_kernel(p):
    from graphgpt_b200 import ops
    n = 8 * 40000 + 5                       # vector body + scalar tail
    sc = ops.dropout_scale(n, p, 1234, "cuda")
    pq = round(p * 65536) / 65536
    vals = torch.unique(sc)
    assert vals.numel() == 2 and vals[0].item() == 0.0 and abs(vals[1].item() - 1 / (1 - pq)) < 1e-6
    keep = (sc > 0).float().mean().item()
    assert abs(keep - (1 - pq)) < 5 * math.sqrt(pq * (1 - pq) / n)
    # neighbours (the two halves of one 32-bit draw) and distant elements are uncorrelated
    k = (sc > 0).float() - (1 - pq)
    for lag in (1, 2, 3, 8, 1024):
        c = (k[:-lag] * k[lag:]).mean().item() / (pq * (1 - pq))
        assert abs(c) < 5 / math.sqrt(n), (lag, c)
    assert torch.equal(sc, ops.dropout_scale(n, p, 1234, "cuda"))
    assert not torch.equal(sc, ops.dropout_scale(n, p, 1235, "cuda"))
    x = torch.randn(n, device="cuda").bfloat16()
    want = (x.float() * sc).bfloat16()
    got = ops.dropout_(x.clone(), p, 1234)
    assert torch.equal(got, want)
    assert torch.equal(ops.dropout_(x.clone(), 0.0, 1), x)


@pytest.mark.parametrize("gated", [False, True])
def test_embed_dropout_fwd_bwd(gated):
    from graphgpt_b200 import ops
    T, F_, d, V, p, seed = 300, 5, 128, 97, 0.25, 99
    g = torch.Generator(device="cuda").manual_seed(3)
    ids = torch.randint(0, V, (T, F_), device="cuda", generator=g)
    table = torch.randn(V, d, device="cuda", generator=g).requires_grad_(True)
    gate = torch.rand(F_, d, device="cuda", generator=g).requires_grad_(True) if gated else None
    sc = ops.dropout_scale(T * F_ * d, p, seed, "cuda").view(T, F_, d)
    e = table[ids] * sc
    ref = (e * gate[None]).sum(1) if gated else e.sum(1)
    out = ops.embed_fwd(ids, table.detach(), gate.detach() if gated else None, drop_p=p, drop_seed=seed)
    assert _relf(out, ref) < 1e-6
    dx = torch.randn(T, d, device="cuda", generator=g)
    ref.backward(dx)
    dtable = torch.zeros(V, d, device="cuda")
    dgate = torch.zeros(F_, d, device="cuda") if gated else None
    ops.embed_bwd(ids, dx, table.detach() if gated else None, gate.detach() if gated else None, dtable, dgate,
                  padding_idx=-1, drop_p=p, drop_seed=seed)
    assert _relf(dtable, table.grad) < 1e-5
    if gated:
        assert _relf(dgate, gate.grad) < 1e-5


@pytest.mark.parametrize("E,with_labels", [(24, True), (16, False), (200, True)])
def test_raw_embed_norm_kernels(E, with_labels):
    from graphgpt_b200 import ops
    T, F_, eps = 777, 13, 1e-6
    g = torch.Generator(device="cuda").manual_seed(5)
    raw = torch.randn(T, E, device="cuda", generator=g) * 2
    w = (1 + 0.1 * torch.randn(E, device="cuda", generator=g)).requires_grad_(True)
    tok = (0.3 * torch.randn(1, 1, E, device="cuda", generator=g)).requires_grad_(True)
    labels = None
    if with_labels:
        labels = torch.randint(0, 50, (T, F_), device="cuda", generator=g)
        labels[torch.rand(T, F_, device="cuda", generator=g) < 0.15] = -100     # ~12 % of rows fully labelled -> swapped
        keep_ref = (labels == -100).any(-1, keepdim=True)
        src = torch.where(keep_ref, raw, tok.view(1, E))
    else:
        src = raw
    rstd_ref = torch.rsqrt(src.pow(2).mean(-1, keepdim=True) + eps)
    h_ref = w * (src * rstd_ref)
    h, rstd, keep = ops.raw_embed_norm_fwd(raw, w.detach(), eps, labels=labels, fchk=F_ if with_labels else 0,
                                           mask_tok=tok.detach() if with_labels else None)
    assert _relf(h, h_ref) < 4e-3                                              # bf16 output rounding
    assert _relf(rstd, rstd_ref.view(-1)) < 1e-6
    if with_labels:
        assert torch.equal(keep.bool(), keep_ref.view(-1)) and 0 < (~keep.bool()).sum().item() < T
    dh = torch.randn(T, E, device="cuda", generator=g).bfloat16()
    h_ref.backward(dh.float())
    dw = torch.zeros(E, device="cuda")
    dtok = torch.zeros(1, 1, E, device="cuda") if with_labels else None
    ops.raw_embed_norm_bwd(dh, raw, keep, tok.detach() if with_labels else None, rstd, w.detach(), dw, dtok)
    assert _relf(dw, w.grad) < 1e-4
    if with_labels:
        assert _relf(dtok, tok.grad) < 1e-4


def _cfg(**kw):
    cfg = dict(vocab_size=756, hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
               num_key_value_heads=2, head_dim=64, hidden_act="gelu", max_position_embeddings=1024, rms_norm_eps=1e-6,
               rope_theta=10000.0, pad_token_id=0, bos_token_id=20, eos_token_id=19, causal_attention=False,
               stacked_feat=13, stack_method="short", stacked_feat_agg_method="sum", next_n_token=13, use_cache=False,
               attention_dropout=0.0)
    cfg.update(kw)
    return cfg


@pytest.mark.parametrize("embed_dim", [0, 16])
def test_training_mode_dropout_matches_oracle_with_the_same_masks(embed_dim):
    """mlp_pdrop + embed_pdrop (+ raw_embed_dropout) in train() mode: loss, logits and every gradient against the
    oracle evaluated with the kernels' own masks (utils_graphgpt.py:69-83, modeling_helpers.py:97-98)."""
    from graphgpt_b200 import GraphGPTConfig, GraphGPTPretrainBase, ops, synth
    from graphgpt_b200.engine import _SEED_ACT, _SEED_EMBED, _SEED_MLP, _SEED_RAW, mix_seed
    from oracle import graphgpt_oracle as oracle
    cfgd = _cfg(mlp_pdrop=0.1, embed_pdrop=0.2, embed_dim=embed_dim)
    b = synth.make_batch(3, 128, layout="packed", seed=31)
    ids, am, labels = (torch.from_numpy(b[k]) for k in ("input_ids", "attention_mask", "labels"))
    N, S, F_ = ids.shape
    d, I, L = cfgd["hidden_size"], cfgd["intermediate_size"], cfgd["num_hidden_layers"]
    sd = oracle.init_state_dict(cfgd, seed=9)
    kw = {}
    if embed_dim:
        g = torch.Generator().manual_seed(2)
        sd["embed_layernorm.weight"] = 1 + 0.1 * torch.randn(embed_dim, generator=g)
        sd["emb_mask_token"] = 0.3 * torch.randn(1, 1, embed_dim, generator=g)
        sd["embed_proj.weight"] = 0.05 * torch.randn(d, embed_dim, generator=g)
        full = (torch.rand(N, S, generator=g) < 0.3) & (am.diagonal(dim1=1, dim2=2) == 1 if am.dim() == 3 else am == 1)
        labels = torch.where(full[:, :, None] & (labels == -100), ids, labels)    # some fully labelled rows -> mask token
        kw["inputs_raw_embeds"] = torch.randn(N, S, embed_dim, generator=g)
    model = GraphGPTPretrainBase(GraphGPTConfig(**cfgd))
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    torch.manual_seed(77)
    out = model(input_ids=ids.cuda(), attention_mask=am.cuda(), labels=labels.cuda(),
                **{k: v.cuda() for k, v in kw.items()})
    out.head1_loss.backward()
    torch.manual_seed(77)
    base = int(torch.randint(0, 2 ** 62, (1,)).item())          # the draw HotPath.backbone_forward made
    T = N * S

    def scale(n, p, stream, shape):
        return ops.dropout_scale(n, p, mix_seed(base, stream), "cuda").view(shape).cpu()

    drop = {"embed": scale(T * F_ * d, 0.2, _SEED_EMBED, (N, S, F_, d)),
            "act": [scale(T * I, 0.1, _SEED_ACT + i, (N, S, I)) for i in range(L)],
            "mlp": [scale(T * d, 0.1, _SEED_MLP + i, (N, S, d)) for i in range(L)]}
    if embed_dim:
        drop["raw"] = scale(T * embed_dim, 0.2, _SEED_RAW, (N, S, embed_dim))
    sd_ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = oracle.pretrain_forward(sd_ref, cfgd, ids, am, labels, drop=drop, **kw)
    ref["loss"].backward()
    e_loss = abs(out.head1_loss.item() - ref["loss"].item()) / ref["loss"].item()
    e_lg = _relf(out.head1_logits, ref["logits"].detach())
    assert e_loss <= 1e-3 and e_lg <= 1e-2, (e_loss, e_lg)
    worst = max((_relf(p.grad, sd_ref[k].grad), k) for k, p in model.named_parameters())
    assert worst[0] <= 3e-2, worst
    # eval mode: every dropout is the identity and matches the plain oracle
    model.eval()
    with torch.no_grad():
        out_e = model(input_ids=ids.cuda(), attention_mask=am.cuda(), labels=labels.cuda(),
                      **{k: v.cuda() for k, v in kw.items()})
    ref_e = oracle.pretrain_forward(sd, cfgd, ids, am, labels, **kw)
    assert abs(out_e.head1_loss.item() - ref_e["loss"].item()) / ref_e["loss"].item() <= 1e-3
    assert abs(out_e.head1_loss.item() - out.head1_loss.item()) > 1e-4       # and training mode really dropped


@pytest.mark.parametrize("T,d,use_lam,use_rs", [(1000, 768, True, True), (333, 128, True, False), (500, 1024, False, True)])
def test_layerscale_bwd_kernel(T, d, use_lam, use_rs):
    from graphgpt_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(T)
    dx = torch.randn(T, d, device="cuda", generator=g)
    lam = (1 + 0.3 * torch.randn(d, device="cuda", generator=g)) if use_lam else None
    rs = ((torch.rand(T, device="cuda", generator=g) < 0.8).float() / 0.8) if use_rs else None
    y = torch.randn(T, d, device="cuda", generator=g)
    x_in = torch.randn(T, d, device="cuda", generator=g)
    scale = (lam[None, :] if use_lam else 1.0) * (rs[:, None] if use_rs else 1.0)
    x_out = x_in + scale * y
    dlam = torch.zeros(d, device="cuda") if use_lam else None
    dy = ops.layerscale_bwd(dx, x_out if use_lam else None, x_in if use_lam else None, lam, rs, dlam)
    want_dy = dx
    if use_lam:
        want_dy = want_dy * lam[None, :]
    if use_rs:
        want_dy = want_dy * rs[:, None]
    assert torch.equal(dy, want_dy.bfloat16())          # same multiplication order as the kernel: (dx * lam) * rs
    if use_lam:
        want = (dx * (rs[:, None] if use_rs else 1.0) * y).sum(0)
        assert _relf(dlam, want) < 2e-4          # (x_out - x_in)/lam recovers rs*y to fp32 round-off


def test_ppa_finetune_training_mode_droppath_layerscale_matches_oracle():
    """C3 (ogbl-ppa fine-tune) training mode: LayerScale (lsi = 1) + DropPath 0.2 with the reference's own draw order
    (torch.rand((N,1,1)) per residual branch, layer by layer) — loss and every gradient against the oracle fed the
    same per-sample masks (utils_graphgpt.py:137-173)."""
    from graphgpt_b200 import GraphGPTConfig, GraphGPTTaskModel, synth
    from oracle import graphgpt_oracle as oracle
    cfgd = _cfg(vocab_size=1200, stacked_feat=4, next_n_token=4, num_hidden_layers=3, num_labels=2,
                problem_type="single_label_classification", pooling_method="last", layer_scale_init_value=1.0,
                path_pdrop=0.4)
    vocab = synth.VocabLayout(vocab_size=1200, scope=512, n_node_attr=2, n_edge_attr=1)
    N = 8
    b = synth.make_batch(N, 64, layout="unpacked", task="ntp", vocab=vocab, seed=51)
    ids, am = torch.from_numpy(b["input_ids"]), torch.from_numpy(b["attention_mask"])
    labels = torch.tensor([1, 0, 1, 1, 0, 0, 1, 0])
    sd = oracle.init_state_dict(cfgd, seed=4, task_head=True)
    g = torch.Generator().manual_seed(1)
    for k in sd:
        if "lambda_" in k:
            sd[k] = sd[k] * (1 + 0.2 * torch.randn(sd[k].shape, generator=g))
    model = GraphGPTTaskModel(GraphGPTConfig(**cfgd))
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    torch.manual_seed(123)
    out = model(input_ids=ids.cuda(), attention_mask=am.cuda(), task_labels=labels.cuda())
    out.task_loss.backward()
    # replay the DropPath draws: linspace(0, p, L) per layer, two draws per layer with p > 0
    torch.manual_seed(123)
    L, p = cfgd["num_hidden_layers"], cfgd["path_pdrop"]
    path = []
    for i in range(L):
        pi = p * i / (L - 1)
        if pi <= 0:
            path.append((None, None))
            continue
        keep = 1 - pi
        path.append(tuple((torch.floor(keep + torch.rand((N, 1, 1), device="cuda")).view(N) / keep).cpu() for _ in range(2)))
    assert any(s is not None and bool((s == 0).any()) for pr in path for s in pr)     # some sample really was dropped
    sd_ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = oracle.task_forward(sd_ref, cfgd, ids, am, task_labels=labels, drop={"path": path})
    ref["loss"].backward()
    assert abs(out.task_loss.item() - ref["loss"].item()) / ref["loss"].item() <= 1e-2
    worst = max((_relf(p_.grad, sd_ref[k].grad), k) for k, p_ in model.named_parameters() if sd_ref[k].grad is not None)
    assert worst[0] <= 3e-2, worst


def test_device_packer_matches_oracle_and_segment_mask_equals_block_diagonal():
    """ggpt_pack_sequences against oracle/packing_oracle.py (bit-exact), and the [N,S] segment-id mask it emits against
    the reference's [N,S,S] block-diagonal mask: the model must produce bit-identical logits and loss with either."""
    import numpy as np
    from graphgpt_b200 import GraphGPTConfig, GraphGPTPretrainBase, ops
    from graphgpt_b200.packing import pack_sequences, plan_greedy
    from oracle import graphgpt_oracle as oracle
    from oracle import packing_oracle as po
    rng = np.random.default_rng(3)
    F_, S = 13, 256
    lengths = rng.integers(4, 40, size=150)
    lengths[7] = 300                                               # one graph longer than a whole sequence
    graphs = [rng.integers(22, 756, size=(L, F_)) for L in lengths]
    rows = torch.from_numpy(np.concatenate(graphs)).cuda()
    cu_rows = np.concatenate([[0], np.cumsum(lengths)])
    seq_graphs, cu_seq = plan_greedy(lengths, S)
    sep = [19] * F_
    out = pack_sequences(rows, cu_rows, seq_graphs, cu_seq, S, sep)
    N = len(cu_seq) - 1
    am3 = np.zeros((N, S, S), np.int64)
    for n in range(N):
        ids, am, seg = po.pack_one([graphs[g] for g in seq_graphs[cu_seq[n]:cu_seq[n + 1]]], sep, S)
        assert np.array_equal(out["input_ids"][n].cpu().numpy(), ids)
        assert np.array_equal(out["attention_mask"][n].cpu().numpy(), seg)
        assert int(out["n_valid"][n]) == int((seg > 0).sum())
        am3[n] = am
    assert torch.equal(out["position_ids"].cpu(), torch.arange(S).expand(N, S))
    # mask bits from segment ids == mask bits from the [N,S,S] tensor
    m_seg = ops.attn_mask_build(out["attention_mask"], N, S, False, rows.device)
    m_3d = ops.attn_mask_build(torch.from_numpy(am3).cuda(), N, S, False, rows.device)
    valid = out["attention_mask"] > 0
    assert torch.equal(m_seg.bits, m_3d.bits) and torch.equal(m_seg.cls, m_3d.cls)     # pad rows included: they see nothing
    assert torch.equal(m_seg.iso_count, m_3d.iso_count)
    # ... and the model agrees bit for bit on every labelled entry
    cfgd = _cfg()
    model = GraphGPTPretrainBase(GraphGPTConfig(**cfgd))
    model.load_state_dict(oracle.init_state_dict(cfgd, seed=2), strict=True)
    model = model.cuda().eval()
    labels = torch.where((torch.rand(N, S, F_, device="cuda") < 0.4) & valid[:, :, None], out["input_ids"],
                         torch.full_like(out["input_ids"], -100))
    ids_in = torch.where(labels != -100, torch.ones_like(labels), out["input_ids"])
    with torch.no_grad():
        a = model(input_ids=ids_in, attention_mask=out["attention_mask"], labels=labels)
        b = model(input_ids=ids_in, attention_mask=torch.from_numpy(am3).cuda(), labels=labels)
    assert torch.equal(a.head1_logits, b.head1_logits) and a.head1_loss.item() == b.head1_loss.item()


@pytest.mark.parametrize("n_out,d,density", [(1000, 768, 0.5), (70000, 64, 0.9), (513, 128, 0.01), (256, 768, 1.0), (300, 64, 0.0)])
def test_expand_rows_equals_memset_plus_scatter(n_out, d, density):
    from graphgpt_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(n_out)
    keep = torch.rand(n_out, device="cuda", generator=g) < density
    idx = keep.nonzero().view(-1).to(torch.int32)
    n = idx.numel()
    src = torch.randn(max(n, 1), d, device="cuda", generator=g).bfloat16()
    want = torch.zeros(n_out, d, device="cuda", dtype=torch.bfloat16)
    if n:
        want[idx.long()] = src[:n]
    got = ops.expand_rows(src, idx if n else torch.zeros(1, device="cuda", dtype=torch.int32), n, n_out)
    assert torch.equal(got, want)


def test_stack_path_rows_matches_python_stacking():
    """ggpt_stack_path_rows (device gather of the stacked token rows, tokenizer.py:1196-1266) against a plain numpy
    restatement, on walks produced by the C++ Euler routine; the rows then feed the device-side packer."""
    import numpy as np
    from graphgpt_b200 import euler
    rng = np.random.default_rng(4)
    An, Ae, base = 3, 2, 22
    graphs, node_attr, edge_attr = [], [], []
    for _ in range(40):
        n = int(rng.integers(1, 25))
        edges = [(j, int(rng.integers(max(0, j - 3), j))) for j in range(1, n)]
        e = euler.dedup_edges(np.asarray(edges, dtype=np.int64).reshape(-1, 2), n)
        graphs.append((n, e))
        node_attr.append(rng.integers(600, 700, size=(n, An)))
        edge_attr.append(rng.integers(700, 750, size=(len(e), Ae)))
    steps, maps = euler.euler_paths(graphs, seed=5, scope=512)
    default_edge = np.asarray([598, 599], dtype=np.int64)
    row_node, row_edge, expect = [], [], []
    n_off = e_off = 0
    for (n, e), st, mp, na, ea in zip(graphs, steps, maps, node_attr, edge_attr):
        rn, re_ = euler.path_rows(st, n_off, e_off)
        row_node.append(rn)
        row_edge.append(re_)
        for node, edge in zip(rn - n_off, np.where(re_ >= 0, re_ - e_off, -1)):
            expect.append([base + int(mp[node])] + na[node].tolist() + (ea[edge].tolist() if edge >= 0 else default_edge.tolist()))
        n_off += n
        e_off += len(e)
    cat = lambda xs, dt: torch.from_numpy(np.concatenate(xs).astype(dt)).cuda()
    rows = euler.stack_rows(cat(row_node, np.int32), cat(row_edge, np.int32), cat(maps, np.int32), cat(node_attr, np.int64),
                            cat([x.reshape(-1, Ae) for x in edge_attr], np.int64), torch.from_numpy(default_edge).cuda(), base)
    assert torch.equal(rows.cpu(), torch.tensor(expect, dtype=torch.int64))
