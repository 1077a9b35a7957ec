"""GPU parity of the un-masking generation kernels (ggpt_gen_sample, ggpt_gen_unmask_origin / _topk) and of the
sample_per_batch / sample_per_example loops (graphgpt_b200.generation) against
  (1) tests/golden/aux/generation.pt = outputs of the reference's own generation_utils.py functions, and
  (2) oracle/generation_oracle.py fed the same logits and the same uniform draws.
Token ids, revealed positions and step counters must agree exactly; probabilities / confidences to 2e-5 relative
(fast-math exp in the kernel)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "aux", "generation.pt")


def test_sample_kernel_matches_reference_filters_and_confidences():
    from graphgpt_b200 import generation as gen
    rec = torch.load(GOLD)
    logits = rec["logits"].cuda()
    for f in rec["filters"]:
        _, _, probs = gen.sample_tokens(logits, 0.0 if f["temperature"] == 0 else f["temperature"], f["top_p"], f["top_k"],
                                        u=torch.zeros(logits.shape[:-1], device="cuda"), want_probs=True)
        assert torch.equal(probs.cpu() > 0, f["probs"] > 0), f
        assert torch.allclose(probs.cpu(), f["probs"], rtol=2e-5, atol=1e-8), f
    for s in rec["sample_tokens"]:
        conf, x0 = gen.sample_tokens(logits, 0.0, s["top_p"], s["top_k"], s["margin"], s["neg_entropy"])
        assert torch.equal(x0.cpu(), s["x0"]), s
        assert torch.allclose(conf.cpu(), s["conf"], rtol=2e-5, atol=1e-7), s


@pytest.mark.parametrize("V", [756, 5000])
def test_sample_kernel_large_rows_and_inverse_cdf_sampling(V):
    """Warp-per-row (V <= 2048) and CTA-per-row kernels against the oracle, including temperature sampling with
    shared uniform draws; then the empirical distribution of the draws."""
    from graphgpt_b200 import generation as gen
    from oracle import generation_oracle as go
    g = torch.Generator().manual_seed(V)
    R = 3000
    logits = torch.randn(R, V, generator=g) * 2.5
    u = torch.rand(R, generator=g)
    for kw in (dict(temperature=0.8), dict(temperature=1.2, top_k=50), dict(temperature=0.7, top_p=0.9, top_k=100)):
        conf, x0, probs = gen.sample_tokens(logits.cuda(), u=u.cuda(), want_probs=True, **kw)
        p_ref = go.filtered_probs(logits, kw["temperature"], kw.get("top_p"), kw.get("top_k"))
        c_ref, x_ref = go.sample_tokens(logits, u=u, **kw)
        # a token whose preceding cumulative mass equals top_p to fp32 round-off may fall on either side (the reference
        # sums a sorted cumsum, the kernel sums the mass above a threshold): allow that on a handful of rows
        rows_ok = ((probs.cpu() > 0) == (p_ref > 0)).all(-1)
        assert rows_ok.float().mean().item() >= (0.998 if kw.get("top_p") else 1.0), (kw, int((~rows_ok).sum()))
        assert torch.allclose(probs.cpu()[rows_ok], p_ref[rows_ok], rtol=2e-5, atol=1e-9)
        agree = (x0.cpu() == x_ref) & rows_ok
        assert agree.float().mean().item() > 0.997            # fp32 vs fp64 prefix sums may flip a draw on a boundary
        assert torch.allclose(conf.cpu()[agree], c_ref[agree], rtol=2e-5, atol=1e-8)
        assert bool((torch.gather(probs.cpu(), 1, x0.cpu()[:, None]) > 0).all())   # never a filtered-out token
    for mode in (dict(margin_confidence=True), dict(neg_entropy=True)):
        conf, x0 = gen.sample_tokens(logits.cuda(), **mode)
        c_ref, x_ref = go.sample_tokens(logits, **mode)
        assert torch.equal(x0.cpu(), x_ref) and torch.allclose(conf.cpu(), c_ref, rtol=5e-5, atol=1e-6)
    # distribution: 20000 draws from one 8-token row
    row = torch.tensor([[2.0, 1.0, 0.5, 0.0, -1.0, -3.0, 1.5, 0.2]]).repeat(20000, 1).cuda()
    _, x0 = gen.sample_tokens(row, temperature=1.0)
    freq = torch.bincount(x0, minlength=8).float().cpu() / 20000
    assert torch.allclose(freq, torch.softmax(row[0].cpu(), -1), atol=0.015)


def test_unmask_kernels_match_reference_traces():
    from graphgpt_b200 import generation as gen
    rec = torch.load(GOLD)
    n = 0
    for U in rec["unmask"]:
        cfg = gen.GenerationConfig(alg=U["alg"], mask_token_id=1)
        ts = U["timesteps"].cuda()
        for st in U["trace"]:
            x = st["x_in"].clone().cuda()
            B, P = x.shape
            lg = st["logits"].cuda().view(B * P, -1)
            if U["alg"] == "origin":
                # feed the reference's own uniform draw instead of a fresh torch.rand
                _, x0 = gen.sample_tokens(lg)
                from graphgpt_b200 import ops
                p = gen._p_transfer(ts, st["i_in"], len(ts) - 1)
                ops.gen_unmask_origin(x, x0.view(B, P), st["u_transfer"].cuda(), float(p), 1)
                i = st["i_in"] + 1
                assert torch.equal(x.cpu(), st["x_out"])
            else:
                x, i = gen._unmask_step(x, lg, ts, st["i_in"], cfg)
                was = st["x_in"] == 1
                assert torch.equal(x.cpu()[was], st["x_out"][was]), (U["alg"], U["steps"], st["i_in"])
                assert torch.equal(x.cpu()[~was], st["x_in"][~was])          # revealed tokens are never re-masked
            assert i == st["i_out"]
            n += 1
    assert n >= 60


def test_unmask_topk_gumbel_and_ties():
    from graphgpt_b200 import ops
    from oracle import generation_oracle as go
    g = torch.Generator().manual_seed(3)
    B, P, V = 7, 500, 40
    x = torch.randint(2, V, (B, P), generator=g)
    x[torch.rand(B, P, generator=g) < 0.6] = 1
    x[0] = 5                                                     # a sample with nothing to reveal
    logits = torch.randn(B, P, V, generator=g)
    ug = torch.rand(B, P, generator=g)
    ts = torch.linspace(1, 1e-3, 6)
    ref, i_ref = go.batch_unmask(x, logits, ts, 1, alg="maskgit_plus", alg_temp=0.7, u_gumbel=ug)
    conf, x0 = go.sample_tokens(logits)
    ntps, k, _ = go.num_transfer((x == 1).sum(1), ts, 1)
    xg = x.clone().cuda()
    ops.gen_unmask_topk(xg, x0.cuda(), conf.cuda(), ntps.cuda(), 1, gumbel_u=ug.cuda(), alg_temp=0.7)
    assert torch.equal(xg.cpu(), ref)
    # exact ties (quantised confidences): lowest positions win, as in the oracle's stable sort
    conf_q = (conf * 4).round() / 4
    xq = x.clone().cuda()
    ops.gen_unmask_topk(xq, x0.cuda(), conf_q.cuda(), ntps.cuda(), 1)
    c = conf_q.clone()
    c[x != 1] = -torch.inf
    order = torch.sort(c, dim=1, descending=True, stable=True)[1]
    want = x.clone()
    for b in range(B):
        sel = order[b, : int(ntps[b])]
        want[b, sel] = x0[b, sel]
    assert torch.equal(xq.cpu(), want)


def _tiny_model():
    from graphgpt_b200 import GraphGPTConfig, GraphGPTPretrainBase
    from oracle import graphgpt_oracle as oracle
    cfgd = dict(vocab_size=120, hidden_size=64, intermediate_size=256, num_hidden_layers=2, num_attention_heads=1,
                num_key_value_heads=1, head_dim=64, hidden_act="gelu", max_position_embeddings=256, rms_norm_eps=1e-6,
                rope_theta=10000.0, pad_token_id=0, causal_attention=False, stacked_feat=4, stack_method="short",
                stacked_feat_agg_method="sum", next_n_token=4, use_cache=False)
    model = GraphGPTPretrainBase(GraphGPTConfig(**cfgd))
    sd = oracle.init_state_dict(cfgd, seed=3)
    sd["lm_head.weight"] *= 25.0        # peaked predictions: confidences of different entries are well separated
    sd["lm_head.weight"][1] = 0.0       # ... and the model rarely predicts the <mask> id itself
    model.load_state_dict(sd, strict=True)
    return model.cuda().eval()


@pytest.mark.parametrize("alg", ["origin", "maskgit_plus", "topk_margin", "entropy"])
def test_sample_per_batch_loop_replays_through_the_oracle(alg):
    """The whole decoding loop: every step's logits are captured and the step is replayed on the CPU oracle with the
    same uniform draws; the token grids must stay identical through to the end."""
    from graphgpt_b200 import generation as gen
    from oracle import generation_oracle as go
    model = _tiny_model()
    g = torch.Generator().manual_seed(11)
    bz, seq, F_ = 3, 24, 4
    ids = torch.randint(22, 120, (bz, seq, F_), generator=g)
    am = torch.ones(bz, seq, dtype=torch.long)
    for b, ln in enumerate((24, 17, 9)):
        ids[b, ln:] = 0
        am[b, ln:] = 0
    masked = (torch.rand(bz, seq, F_, generator=g) < 0.5) & (ids != 0)
    ids[masked] = 1
    cfg = gen.GenerationConfig(alg=alg, steps=6, mask_token_id=1, output_history=True)
    captured = []
    fwd = model.forward

    def spy(**kw):
        out = fwd(**kw)
        captured.append((kw["input_ids"].clone().cpu(), out.head1_logits.clone().cpu()))
        return out

    model.forward = spy
    torch.manual_seed(5)
    x, hist = gen.sample_per_batch(model, cfg, input_ids=ids, attention_mask=am, inputs_raw_embeds=None)
    model.forward = fwd
    assert len(hist) == len(captured) >= 1
    steps = min(int(masked.view(bz, -1).sum(-1).max()), cfg.steps)
    ts = torch.linspace(1, cfg.eps, steps + 1)
    torch.manual_seed(5)
    i = 0
    xr = ids.view(bz, -1).clone()
    for (x_in, lg), h in zip(captured, hist):
        assert torch.equal(x_in.view(bz, -1), xr)
        u = torch.rand(xr.shape, device="cuda").cpu() if alg == "origin" else None
        xr, i = go.batch_unmask(xr, lg.view(bz, seq * F_, -1), ts, i, alg=alg, u_transfer=u)
        assert torch.equal(h.view(bz, -1).cpu(), xr)
    assert torch.equal(x.cpu(), xr)
    assert i == steps
    # every masked entry was generated (an entry still holding id 1 is one where the model predicted <mask> itself)
    last_x0 = captured[-1][1].argmax(-1).view(bz, -1)
    left = x.cpu() == 1
    assert bool((last_x0[left] == 1).all()) and int(left.sum()) <= 2
    assert torch.equal(x.cpu()[~masked.view(bz, -1)], ids.view(bz, -1)[~masked.view(bz, -1)])
    acc = gen.cal_gen_acc_batch(cfg, ids.cuda(), ids.cuda(), x)
    assert acc.shape == (bz,)


def test_sample_per_example_reveals_everything():
    from graphgpt_b200 import generation as gen
    model = _tiny_model()
    g = torch.Generator().manual_seed(12)
    ids = torch.randint(22, 120, (20, 4), generator=g)
    masked = torch.rand(20, 4, generator=g) < 0.4
    ids[masked] = 1
    for alg, alg_temp in (("origin", None), ("maskgit_plus", None), ("entropy", 0.5)):
        cfg = gen.GenerationConfig(alg=alg, alg_temp=alg_temp, steps=8, mask_token_id=1, temperature=0.0)
        x, _ = gen.sample_per_example(model, cfg, input_ids=ids, attention_mask=torch.ones(1, 20, dtype=torch.long))
        assert x.shape == (1, 80) and int((x == 1).sum()) <= 1    # (id 1 survives only where the model predicts it)
        assert torch.equal(x.cpu().view(20, 4)[~masked], ids[~masked])
