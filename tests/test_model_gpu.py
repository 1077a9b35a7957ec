"""End-to-end GPU parity of the drop-in model classes (C-ABI kernels) against
  (1) the committed golden fixtures = outputs of the unmodified reference (fp32, CPU), and
  (2) the CPU oracle on fresh seeded inputs at a larger size.
The kernels feed bf16 operands to the tensor cores (fp32 accumulate, fp32 residual stream / norms / softmax / CE),
so parity with the fp32 reference is bounded by bf16 operand rounding (2^-9 relative per element):
  loss: <= 1e-3 relative (BASELINE north-star tolerance);  logits: relative Frobenius error <= the error of the
  reference's own bf16 run on the same fixture (hidden states of fine-tune fixtures: <= 1e-2);  gradients: <= 3e-2.
Each fixture also records how far the REFERENCE ITSELF lands from its fp32 result when run in bf16
(rec["ref_bf16_error"]); every number is logged next to it in gpurun_out/parity_report.txt and a gradient may
exceed 3e-2 only where the reference's own bf16 run does (tolerance = max(3e-2, 1.5 x reference bf16 error))."""
import glob
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt")))
LOSS_TOL, ACT_TOL, GRAD_TOL = 1e-3, 1e-2, 3e-2
REPORT = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out", "parity_report.txt")


def _relf(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-30)).item()


def _log(msg):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(msg + "\n")
    print(msg)


def _build(rec):
    from graphgpt_b200 import GraphGPTConfig, GraphGPTDoubleHeadsModel, GraphGPTPretrainBase, GraphGPTTaskModel
    cfg = GraphGPTConfig(**rec["config"])
    cls = {"pretrain": GraphGPTPretrainBase, "finetune": GraphGPTTaskModel, "double": GraphGPTDoubleHeadsModel}[rec["kind"]]
    model = cls(cfg)
    missing, unexpected = model.load_state_dict(rec["state_dict"], strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return model.cuda().eval()


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_model_matches_reference_golden(path):
    rec = torch.load(path)
    model = _build(rec)
    inp = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in rec["inputs"].items()}
    name = os.path.basename(path)[:-3]
    out = model(**inp)
    if rec["kind"] == "pretrain":
        lg = out.head1_logits
        assert tuple(lg.shape) == tuple(rec["logits_shape"])
        e_lg = _relf(lg[:: rec["logits_stride"]], rec["logits"])
        loss = out.head1_loss
        ref_bf = rec.get("ref_bf16_error", {}).get("logits_relF")
        _log(f"{name}: logits relF {e_lg:.3e} (reference-in-bf16 relF {ref_bf if ref_bf is not None else float('nan'):.3e})")
        # bf16 storage bounds the fast path (tests/test_parity_depth_gpu.py explains the number, the precise mode of
        # tests/test_precise_mode_gpu.py meets 1e-3 outright): never further from fp32 than the reference's own bf16 run
        assert e_lg <= (max(ref_bf, 2e-3) if ref_bf is not None else ACT_TOL), (e_lg, ref_bf)
    elif rec["kind"] == "double":
        e_tl = _relf(out.task_logits, rec["task_logits"])
        e_t = abs(out.task_loss.item() - rec["task_loss"].item()) / abs(rec["task_loss"].item())
        e_p = abs(out.pretrain_loss.item() - rec["pretrain_loss"].item()) / abs(rec["pretrain_loss"].item())
        N_, S_ = rec["inputs"]["attention_mask"].shape
        assert tuple(out.pretrain_logits.shape) == (N_, S_, rec["config"]["vocab_size"]) and out.hidden_states is None
        _log(f"{name}: task_logits relF {e_tl:.3e} task_loss rel {e_t:.3e} pretrain_loss rel {e_p:.3e}")
        assert e_tl <= 5 * ACT_TOL and e_t <= 1e-2 and e_p <= LOSS_TOL
        loss = out.task_loss + out.pretrain_loss
    else:
        am = rec["inputs"]["attention_mask"].bool()
        if out.task_logits.dim() == 3:       # token-level task: the full [N,S,labels] grid, valid positions compared
            e_tl = _relf(out.task_logits.float().cpu()[am], rec["task_logits"][am])
        else:
            e_tl = _relf(out.task_logits, rec["task_logits"])
        e_th = _relf(out.task_hidden_states, rec["task_hidden"])
        e_h = _relf(out.hidden_states.float().cpu()[am], rec["hidden"][am])
        loss = out.task_loss
        _log(f"{name}: task_logits relF {e_tl:.3e} task_hidden relF {e_th:.3e} hidden relF {e_h:.3e}")
        assert e_th <= ACT_TOL and e_h <= ACT_TOL and e_tl <= 5 * ACT_TOL
    if "loss" in rec:
        e_loss = abs(loss.item() - rec["loss"].item()) / abs(rec["loss"].item())
        _log(f"{name}: loss {loss.item():.6f} ref {rec['loss'].item():.6f} rel {e_loss:.3e} "
             f"(reference-in-bf16 rel {rec.get('ref_bf16_error', {}).get('loss_rel', float('nan')):.3e})")
        assert e_loss <= (LOSS_TOL if rec["kind"] == "pretrain" else 1e-2)
        loss.backward()
        named = dict(model.named_parameters())
        for k, g in rec["grads"].items():
            mine = named[k].grad
            assert mine is not None, k
            e_n = abs(float(mine.double().norm()) - rec["grad_norms"][k]) / (rec["grad_norms"][k] + 1e-30)
            part = mine[:24] if mine.dim() == 2 else mine
            e_g = _relf(part, g)
            ref_bf = rec.get("ref_bf16_error", {}).get("grad_relF", {}).get(k, 0.0)
            _log(f"{name}: grad {k}: norm rel {e_n:.3e} relF {e_g:.3e} (reference-in-bf16 relF {ref_bf:.3e})")
            assert e_g <= max(GRAD_TOL, 1.5 * ref_bf) and e_n <= GRAD_TOL, (k, e_g, e_n, ref_bf)


def test_c2_medium_vs_oracle():
    """4L/256d/F=13/V=756, packed S=256 block-diagonal mask, N=4: forward + backward against the CPU oracle."""
    from graphgpt_b200 import GraphGPTConfig, GraphGPTPretrainBase, synth
    from oracle import graphgpt_oracle as oracle
    cfgd = dict(vocab_size=756, hidden_size=256, intermediate_size=1024, num_hidden_layers=4, num_attention_heads=4,
                num_key_value_heads=4, head_dim=64, hidden_act="gelu", max_position_embeddings=1024, rms_norm_eps=1e-6,
                rope_theta=10000.0, pad_token_id=0, bos_token_id=20, eos_token_id=19, causal_attention=False,
                stacked_feat=13, stack_method="short", stacked_feat_agg_method="sum", next_n_token=13, use_cache=False)
    b = synth.make_batch(4, 256, layout="packed", seed=77)
    sd = oracle.init_state_dict(cfgd, seed=5)
    ids, am, labels = (torch.from_numpy(b[k]) for k in ("input_ids", "attention_mask", "labels"))
    sd_ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = oracle.pretrain_forward(sd_ref, cfgd, ids, am, labels)
    ref["loss"].backward()
    model = GraphGPTPretrainBase(GraphGPTConfig(**cfgd))
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    out = model(input_ids=ids.cuda(), attention_mask=am.cuda(), labels=labels.cuda())
    e_loss = abs(out.head1_loss.item() - ref["loss"].item()) / ref["loss"].item()
    e_lg = _relf(out.head1_logits, ref["logits"].detach())
    _log(f"c2_medium: loss {out.head1_loss.item():.6f} ref {ref['loss'].item():.6f} rel {e_loss:.3e}; logits relF {e_lg:.3e}")
    assert e_loss <= LOSS_TOL and e_lg <= ACT_TOL
    out.head1_loss.backward()
    worst = 0.0
    for k, p in model.named_parameters():
        e = _relf(p.grad, sd_ref[k].grad)
        worst = max(worst, e)
        if e > GRAD_TOL:
            _log(f"c2_medium: grad {k} relF {e:.3e}  <-- above tolerance")
    _log(f"c2_medium: worst grad relF {worst:.3e}")
    assert worst <= GRAD_TOL


def _oracle_parity(cfgd, b, tag, train=True):
    from graphgpt_b200 import GraphGPTConfig, GraphGPTPretrainBase
    from oracle import graphgpt_oracle as oracle
    sd = oracle.init_state_dict(cfgd, seed=5)
    ids, am, labels = (torch.from_numpy(b[k]) for k in ("input_ids", "attention_mask", "labels"))
    sd_ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = oracle.pretrain_forward(sd_ref, cfgd, ids, am, labels)
    ref["loss"].backward()
    model = GraphGPTPretrainBase(GraphGPTConfig(**cfgd))
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    out = model(input_ids=ids.cuda(), attention_mask=am.cuda(), labels=labels.cuda())
    e_loss = abs(out.head1_loss.item() - ref["loss"].item()) / ref["loss"].item()
    e_lg = _relf(out.head1_logits, ref["logits"].detach())
    _log(f"{tag}: loss {out.head1_loss.item():.6f} ref {ref['loss'].item():.6f} rel {e_loss:.3e}; logits relF {e_lg:.3e}")
    assert e_loss <= LOSS_TOL and e_lg <= ACT_TOL
    out.head1_loss.backward()
    # relative Frobenius error per parameter; a gradient that is (numerically) zero in the reference — q/k projections
    # when every query sees a single key, so the softmax is constant — is measured against the largest gradient norm
    floor = 1e-3 * max(float(v.grad.norm()) for v in sd_ref.values() if v.grad is not None)
    worst = max((float((p.grad.double().cpu() - sd_ref[k].grad.double()).norm()) /
                 max(float(sd_ref[k].grad.double().norm()), floor), k) for k, p in model.named_parameters())
    _log(f"{tag}: worst grad relF {worst[0]:.3e} ({worst[1]})")
    assert worst[0] <= GRAD_TOL, worst


def test_c5_shape_causal_ntp_seq2048_vs_oracle():
    """C5 geometry (d=1024, 16 heads, I=4096, causal NTP, seq 2048) at 2 layers, one padded + one full sequence."""
    from graphgpt_b200 import synth
    cfgd = dict(vocab_size=756, hidden_size=1024, intermediate_size=4096, num_hidden_layers=2, num_attention_heads=16,
                num_key_value_heads=16, head_dim=64, hidden_act="gelu", max_position_embeddings=2048, rms_norm_eps=1e-6,
                rope_theta=10000.0, pad_token_id=0, bos_token_id=20, eos_token_id=19, causal_attention=True,
                stacked_feat=13, stack_method="short", stacked_feat_agg_method="sum", next_n_token=13, use_cache=False,
                attention_dropout=0.0)
    b = synth.make_batch(2, 2048, layout="dense", task="ntp", seed=41)
    b["attention_mask"][1, 1500:] = 0
    b["input_ids"][1, 1500:] = 0
    b["labels"][1, 1500:] = -100
    _oracle_parity(cfgd, b, "c5_shape")


def test_c4_shape_seq4096_vs_oracle():
    """C4 geometry: bidirectional sequences of 4096 rows (citation2 sub-graph sampling), F=4, 2L/128d."""
    from graphgpt_b200 import synth
    cfgd = dict(vocab_size=1200, hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                num_key_value_heads=2, head_dim=64, hidden_act="gelu", max_position_embeddings=4096, rms_norm_eps=1e-6,
                rope_theta=10000.0, pad_token_id=0, bos_token_id=20, eos_token_id=19, causal_attention=False,
                stacked_feat=4, stack_method="short", stacked_feat_agg_method="sum", next_n_token=4, use_cache=False,
                attention_dropout=0.0)
    vocab = synth.VocabLayout(vocab_size=1200, scope=512, n_node_attr=2, n_edge_attr=1)
    b = synth.make_batch(2, 4096, layout="dense", vocab=vocab, seed=42)
    b["attention_mask"][0, 3000:] = 0
    b["input_ids"][0, 3000:] = 0
    b["labels"][0, 3000:] = -100
    _oracle_parity(cfgd, b, "c4_shape")


def test_full_size_c2_packing_invariance_and_batch_linearity():
    """BASELINE configs[1] at FULL size (12L/768d/F13/V756, 64 x 1024 packed tokens = one bench step) through
    size-independent properties (the oracle comparison at this depth runs on 2 x 1024 tokens in
    tests/test_parity_depth_gpu.py; 64 x 1024 would take the CPU oracle minutes):
      * packing invariance — attention is block-diagonal and RoPE is relative, so the same graphs fed one per row
        (right-padded [n_seg, 40] grid, 2-D mask) must give the same labelled-entry logits and the same loss as the
        packed [64, 1024] grid with its [N,S,S] mask (different kernels: isolated-tile vs general attention path);
      * batch linearity — gradients of the mean loss over the full batch = label-count-weighted mean of the two
        half-batch gradients (exercises split-K wgrad accumulation and the compaction of labelled rows)."""
    from graphgpt_b200 import GraphGPTConfig, GraphGPTPretrainBase, synth
    import numpy as np
    cfgd = dict(vocab_size=756, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                num_key_value_heads=12, head_dim=64, hidden_act="gelu", max_position_embeddings=1024, rms_norm_eps=1e-6,
                rope_theta=10000.0, pad_token_id=0, bos_token_id=20, eos_token_id=19, causal_attention=False,
                stacked_feat=13, stack_method="short", stacked_feat_agg_method="sum", next_n_token=13, use_cache=False,
                attention_dropout=0.0)
    N, S, F_ = 64, 1024, 13
    b = synth.make_batch(N, S, layout="packed", seed=2024, return_segments=True)
    torch.manual_seed(0)
    model = GraphGPTPretrainBase(GraphGPTConfig(**cfgd)).cuda().train()
    ids, am, labels = (torch.from_numpy(b[k]).cuda() for k in ("input_ids", "attention_mask", "labels"))
    out = model(input_ids=ids, attention_mask=am, labels=labels)
    loss_packed, logits_packed = out.head1_loss.item(), out.head1_logits.float().cpu()
    out.head1_loss.backward()
    g_full = {k: p.grad.clone() for k, p in model.named_parameters()}
    n_full = int((labels != -100).sum())
    assert logits_packed.shape[0] == n_full and np.isfinite(loss_packed)
    # ---- the same graphs, one per row
    segs = [(n, o, L) for n, lens in enumerate(b["segment_lens"]) for o, L in zip(np.cumsum([0] + lens[:-1]), lens)]
    Su = (max(L for _, _, L in segs) + 7) // 8 * 8
    u_ids = np.zeros((len(segs), Su, F_), np.int64)
    u_lab = np.full((len(segs), Su, F_), -100, np.int64)
    u_am = np.zeros((len(segs), Su), np.int64)
    for i, (n, o, L) in enumerate(segs):
        u_ids[i, :L], u_lab[i, :L], u_am[i, :L] = b["input_ids"][n, o:o + L], b["labels"][n, o:o + L], 1
    model.zero_grad(set_to_none=True)
    with torch.no_grad():
        out_u = model(input_ids=torch.from_numpy(u_ids).cuda(), attention_mask=torch.from_numpy(u_am).cuda(),
                      labels=torch.from_numpy(u_lab).cuda())
    e_loss = abs(out_u.head1_loss.item() - loss_packed) / loss_packed
    e_lg = _relf(out_u.head1_logits, logits_packed)
    _log(f"full_size_c2: packed loss {loss_packed:.6f} vs one-graph-per-row {out_u.head1_loss.item():.6f} (rel {e_loss:.2e}); "
         f"logits relF {e_lg:.2e} over {n_full} entries, {len(segs)} graphs")
    assert e_loss <= LOSS_TOL and e_lg <= ACT_TOL
    # ---- batch linearity of the gradients
    acc, cnt = None, 0
    for sl in (slice(0, N // 2), slice(N // 2, N)):
        model.zero_grad(set_to_none=True)
        o2 = model(input_ids=ids[sl], attention_mask=am[sl], labels=labels[sl])
        o2.head1_loss.backward()
        n_part = int((labels[sl] != -100).sum())
        part = {k: p.grad.clone() * n_part for k, p in model.named_parameters()}
        acc = part if acc is None else {k: acc[k] + part[k] for k in acc}
        cnt += n_part
    assert cnt == n_full
    worst = max((_relf(acc[k] / cnt, g_full[k]), k) for k in g_full)
    _log(f"full_size_c2: worst gradient deviation from batch linearity {worst[0]:.2e} ({worst[1]})")
    assert worst[0] <= 2e-2, worst


def test_engine_gradient_accumulation_follows_deepspeed_semantics():
    """gradient_accumulation_steps = 2 (ds_config, conf_utils.py:62-65): backward() scales each micro-batch loss by 1/2 and
    accumulates into the flat gradient buffer, step() is a no-op on the first micro-step and applies AdamW on the
    second; the accumulated gradient equals the mean of the two micro-batch gradients."""
    import copy
    from graphgpt_b200 import GraphGPTConfig, GraphGPTPretrainBase, synth
    from graphgpt_b200.dp import GraphGPTEngine
    cfgd = dict(vocab_size=756, hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                num_key_value_heads=2, head_dim=64, hidden_act="gelu", max_position_embeddings=1024, rms_norm_eps=1e-6,
                rope_theta=10000.0, pad_token_id=0, causal_attention=False, stacked_feat=13, stack_method="short",
                stacked_feat_agg_method="sum", next_n_token=13, use_cache=False, attention_dropout=0.0)
    torch.manual_seed(0)
    model = GraphGPTPretrainBase(GraphGPTConfig(**cfgd)).cuda().train()
    probe = copy.deepcopy(model)
    batches = []
    for seed in (61, 62):
        b = synth.make_batch(2, 256, layout="packed", seed=seed)
        batches.append({k: torch.from_numpy(b[k]).cuda() for k in ("input_ids", "attention_mask", "labels")})
    grads = []
    for b in batches:                                   # reference gradients of each micro-batch, plain autograd surface
        probe.zero_grad(set_to_none=True)
        probe(**b).head1_loss.backward()
        grads.append({k: p.grad.clone() for k, p in probe.named_parameters()})
    engine = GraphGPTEngine(model, lr=1e-3, max_grad_norm=0.0, gradient_accumulation_steps=2)
    w0 = {k: p.detach().clone() for k, p in model.named_parameters()}
    engine.backward(engine(**batches[0]).head1_loss)
    engine.step()
    assert engine.global_steps == 0 and all(torch.equal(p.detach(), w0[k]) for k, p in model.named_parameters())
    engine.backward(engine(**batches[1]).head1_loss)
    acc = {k: p.grad.clone() for k, p in model.named_parameters()}
    worst = max(_relf(acc[k], 0.5 * (grads[0][k] + grads[1][k])) for k in acc)
    assert worst <= 1e-2, worst                          # (bf16 rounding of the 1/2-scaled backward signal)
    engine.step()
    assert engine.global_steps == 1 and engine.micro_steps == 2
    assert any(not torch.equal(p.detach(), w0[k]) for k, p in model.named_parameters())
    assert all(p.grad is None for p in model.parameters())


def _small_cfg(**kw):
    cfg = dict(vocab_size=300, hidden_size=64, intermediate_size=256, num_hidden_layers=2, num_attention_heads=1,
               num_key_value_heads=1, head_dim=64, hidden_act="gelu", max_position_embeddings=256, rms_norm_eps=1e-6,
               rope_theta=10000.0, pad_token_id=0, bos_token_id=20, eos_token_id=19, causal_attention=False,
               stacked_feat=1, stack_method="short", stacked_feat_agg_method="sum", next_n_token=1, use_cache=False,
               attention_dropout=0.0)
    cfg.update(kw)
    return cfg


@pytest.mark.parametrize("N,S,F_", [(1, 8, 1), (3, 37, 1), (2, 130, 4), (5, 1, 13)])
def test_ragged_and_tiny_shapes_vs_oracle(N, S, F_):
    """Edge geometry: a single 8-row sequence, S not a multiple of 8 / 32 / 128, S just above one tile, S = 1; labels
    that are legitimately 0 (the pad id is a valid TARGET, SURVEY §8 semantics 2) and rows with no label at all."""
    import numpy as np
    g = np.random.default_rng(N * 100 + S)
    V = 300
    cfgd = _small_cfg(stacked_feat=F_, next_n_token=F_)
    shape = (N, S, F_) if F_ > 1 else (N, S)
    ids = g.integers(2, V, size=shape).astype(np.int64)
    lens = g.integers(max(1, S // 2), S + 1, size=N)
    am = (np.arange(S)[None, :] < lens[:, None]).astype(np.int64)
    ids[am == 0] = 0
    labels = np.where(g.random(shape) < 0.5, ids, -100)
    labels[g.random(shape) < 0.1] = 0                     # target = pad id: counted, not ignored
    labels[am == 0] = -100
    if (labels != -100).sum() == 0:
        labels.reshape(-1)[0] = 5
    b = {"input_ids": ids, "attention_mask": am, "labels": labels}
    _oracle_parity(cfgd, b, f"ragged_{N}x{S}x{F_}")


def test_batch_without_any_label_gives_nan_loss_and_empty_logits():
    """CrossEntropyLoss over zero targets is NaN in torch (the reference's result); logits are [0, V]."""
    from graphgpt_b200 import GraphGPTConfig, GraphGPTPretrainBase
    model = GraphGPTPretrainBase(GraphGPTConfig(**_small_cfg())).cuda().train()
    ids = torch.randint(2, 300, (2, 16)).cuda()
    out = model(input_ids=ids, attention_mask=torch.ones_like(ids), labels=torch.full_like(ids, -100))
    assert torch.isnan(out.head1_loss) and tuple(out.head1_logits.shape) == (0, 300)
    # ... and its backward still runs, as the reference's does (zero gradients from the loss head; under data parallelism
    # every rank must keep issuing its gradient exchanges — ADVICE r1)
    out.head1_loss.backward()
    assert all(p.grad is not None and float(p.grad.abs().max()) == 0.0 for p in model.parameters())


def test_out_of_range_ids_labels_and_positions_raise_index_error():
    """The reference fails with IndexError / a device-side assert when a token id or a label lies outside the vocabulary
    (nn.Embedding, CrossEntropyLoss).  The kernels flag it on the device; the flag is read back without a mid-step sync,
    so the IndexError surfaces at the next forward / optimizer step, or at once through check_device_errors()."""
    from graphgpt_b200 import GraphGPTConfig, GraphGPTPretrainBase, GraphGPTTaskModel
    model = GraphGPTPretrainBase(GraphGPTConfig(**_small_cfg())).cuda().eval()
    good = torch.randint(2, 300, (2, 16)).cuda()
    am = torch.ones_like(good)
    model(input_ids=good, attention_mask=am, labels=good)
    model.check_device_errors()                                            # clean batch: nothing to report
    bad_ids = good.clone()
    bad_ids[1, 3] = 300                                                    # == vocab_size
    model(input_ids=bad_ids, attention_mask=am, labels=good)
    with pytest.raises(IndexError, match="input_ids"):
        model.check_device_errors()
    bad_lab = good.clone()
    bad_lab[0, 5] = 4000
    model(input_ids=good, attention_mask=am, labels=bad_lab)
    torch.cuda.synchronize()
    with pytest.raises(IndexError, match="labels"):                        # not checked explicitly: the NEXT forward raises
        model(input_ids=good, attention_mask=am, labels=good)
    model(input_ids=good, attention_mask=am, labels=good)                  # the flag was consumed: training can go on
    model.check_device_errors()
    ft = GraphGPTTaskModel(GraphGPTConfig(**_small_cfg(num_labels=2, problem_type="single_label_classification"))).cuda().eval()
    pos = torch.arange(16).repeat(2, 1).cuda()
    pos[0, 2] = 100000                                                     # beyond the rotary table
    ft(input_ids=good, attention_mask=am, position_ids=pos, task_labels=torch.tensor([0, 1]).cuda())
    with pytest.raises(IndexError, match="position_ids"):
        ft.check_device_errors()


def test_frozen_prefix_like_freeze_llama_layers():
    """finetune.freeze (modules_utils.py:45-54): embeddings + the first k layers get requires_grad = False.  Frozen
    parameters must stay bit-identical through engine.step() (no update, no weight decay), backward stops at the first
    trainable layer, and the trainable gradients equal those of the unfrozen model."""
    import copy
    from graphgpt_b200 import GraphGPTConfig, GraphGPTTaskModel, synth
    from graphgpt_b200.dp import GraphGPTEngine
    cfgd = dict(vocab_size=1200, hidden_size=128, intermediate_size=512, num_hidden_layers=3, num_attention_heads=2,
                num_key_value_heads=2, head_dim=64, hidden_act="gelu", max_position_embeddings=1024, rms_norm_eps=1e-6,
                rope_theta=10000.0, pad_token_id=0, causal_attention=False, stacked_feat=4, stack_method="short",
                stacked_feat_agg_method="sum", next_n_token=4, use_cache=False, attention_dropout=0.0, num_labels=2,
                problem_type="single_label_classification", pooling_method="last")
    vocab = synth.VocabLayout(vocab_size=1200, scope=512, n_node_attr=2, n_edge_attr=1)
    b = synth.make_batch(4, 64, layout="unpacked", task="ntp", vocab=vocab, seed=81)
    batch = dict(input_ids=torch.from_numpy(b["input_ids"]).cuda(), attention_mask=torch.from_numpy(b["attention_mask"]).cuda(),
                 task_labels=torch.tensor([1, 0, 0, 1]).cuda())
    torch.manual_seed(0)
    full = GraphGPTTaskModel(GraphGPTConfig(**cfgd)).cuda().train()
    frozen = copy.deepcopy(full)
    for p in frozen.model.embed_tokens.parameters():
        p.requires_grad = False
    for layer in frozen.model.layers[:2]:
        for p in layer.parameters():
            p.requires_grad = False
    full(**batch).task_loss.backward()
    ref_grads = {k: p.grad.clone() for k, p in full.named_parameters()}
    eng = GraphGPTEngine(frozen, lr=1e-2, weight_decay=0.1, max_grad_norm=1.0)
    before = {k: p.detach().clone() for k, p in frozen.named_parameters()}
    eng.backward(eng(**batch).task_loss)
    for k, p in frozen.named_parameters():
        if p.requires_grad:
            assert _relf(p.grad, ref_grads[k]) <= 1e-3, k          # same kernels, same values (split-K order noise only)
        else:
            assert p.grad is None, k
    eng.step()
    for k, p in frozen.named_parameters():
        if p.requires_grad:
            assert not torch.equal(p.detach(), before[k]), k
        else:
            assert torch.equal(p.detach(), before[k]), k


def test_engine_checkpoint_roundtrip(tmp_path):
    """save_checkpoint / load_checkpoint (DeepSpeed-style directory): weights, Adam moments and step counters come back
    exactly, and training continues bit-identically from the restored state."""
    from graphgpt_b200 import GraphGPTConfig, GraphGPTPretrainBase, synth
    from graphgpt_b200.dp import GraphGPTEngine
    cfgd = dict(vocab_size=756, hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                num_key_value_heads=2, head_dim=64, hidden_act="gelu", max_position_embeddings=1024, rms_norm_eps=1e-6,
                rope_theta=10000.0, pad_token_id=0, causal_attention=False, stacked_feat=13, stack_method="short",
                stacked_feat_agg_method="sum", next_n_token=13, use_cache=False, attention_dropout=0.0)
    b = synth.make_batch(2, 256, layout="packed", seed=71)
    batch = {k: torch.from_numpy(b[k]).cuda() for k in ("input_ids", "attention_mask", "labels")}

    def make():
        torch.manual_seed(0)
        return GraphGPTEngine(GraphGPTPretrainBase(GraphGPTConfig(**cfgd)).cuda().train(), lr=1e-3)

    def train_step(e):
        loss = e(**batch).head1_loss
        e.backward(loss)
        e.step()
        return loss.item()

    a = make()
    for _ in range(3):
        train_step(a)
    assert a.save_checkpoint(str(tmp_path), client_state={"epoch": 4})
    assert (tmp_path / "latest").read_text() == "global_step3" and a.device.type == "cuda"
    # <dir>/model.pt: the plain state dict the reference's fine-tune / eval loaders open first (loader_utils.py:176-220)
    plain = torch.load(tmp_path / "model.pt")
    assert all(v.device.type == "cpu" for v in plain.values())
    probe = GraphGPTPretrainBase(GraphGPTConfig(**cfgd))
    assert not any(probe.load_state_dict(plain, strict=True))
    assert all(torch.equal(v, a.module.state_dict()[k].cpu()) for k, v in plain.items())
    cont = [train_step(a) for _ in range(2)]
    fresh = make()
    path, client = fresh.load_checkpoint(str(tmp_path))
    assert path.endswith("global_step3/ggpt_engine_states.pt") and client == {"epoch": 4} and fresh.global_steps == 3
    # the forward pass is deterministic and the restored state is exact: the first loss after the restore is bit-identical,
    # the second within split-K atomics' summation-order noise of the intervening wgrad
    again = [train_step(fresh) for _ in range(2)]
    assert again[0] == cont[0] and abs(again[1] - cont[1]) <= 1e-4 * abs(cont[1])
    assert make().load_checkpoint(str(tmp_path / "nothing_here")) == (None, None)


def test_no_cpu_fallback():
    from graphgpt_b200 import GraphGPTConfig, GraphGPTPretrainBase
    cfg = GraphGPTConfig(vocab_size=300, hidden_size=64, intermediate_size=256, num_hidden_layers=1, num_attention_heads=1,
                         num_key_value_heads=1, hidden_act="gelu", stacked_feat=1, next_n_token=1, causal_attention=False)
    model = GraphGPTPretrainBase(cfg)   # left on the CPU on purpose
    with pytest.raises(RuntimeError):
        model(input_ids=torch.ones((1, 8), dtype=torch.long), attention_mask=torch.ones((1, 8), dtype=torch.long))
