"""Generate the golden fixtures that pin oracle/graphgpt_oracle.py to the REFERENCE's own outputs.

Runs only in the build container (needs /root/reference): imports the unmodified reference model classes through
ref_shim.py, instantiates them on CPU in fp32/eval with seeded weights, feeds seeded synthetic batches and stores
{config, inputs, state_dict, outputs, gradients} as tests/golden/<case>.pt.   Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from ref_shim import load_reference  # noqa: E402

from graphgpt_b200 import synth  # noqa: E402

GRAD_KEYS = [
    "model.embed_tokens.weight", "model.layers.0.self_attn.q_proj.weight", "model.layers.0.self_attn.v_proj.weight",
    "model.layers.0.mlp.gate_proj.weight", "model.layers.1.mlp.down_proj.weight",
    "model.layers.1.input_layernorm.weight", "model.norm.weight", "lm_head.weight", "n_token_proj.weight",
    "score.weight", "stacked_feat_agg.weight", "model.layers.0.lambda_1",
    "embed_proj.weight", "embed_layernorm.weight", "emb_mask_token",
    "score.bias", "score.mlp_modules.0.weight", "score.mlp_modules.1.weight", "score.mlp_modules.1.bias",
]


def base_cfg(**kw):
    cfg = dict(vocab_size=300, hidden_size=64, intermediate_size=256, num_hidden_layers=2, num_attention_heads=1,
               num_key_value_heads=1, head_dim=64, hidden_act="gelu", max_position_embeddings=256, rms_norm_eps=1e-6,
               rope_theta=10000.0, attention_bias=False, mlp_bias=False, tie_word_embeddings=False, use_cache=False,
               pad_token_id=0, bos_token_id=20, eos_token_id=19, causal_attention=False, stacked_feat=1,
               stack_method="short", stacked_feat_agg_method="sum", next_n_token=1, attention_dropout=0.1)
    cfg.update(kw)
    if "num_attention_heads" not in kw:
        cfg["num_attention_heads"] = cfg["num_key_value_heads"] = cfg["hidden_size"] // 64
    return cfg


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


CASES = {}


def case(name):
    def deco(fn):
        CASES[name] = fn
        return fn
    return deco


@case("c1_toy_smtp_2d")
def _c1():
    """C1: toy 2L/64d, F=1, right-padded 2-D mask, bidirectional SMTP."""
    cfg = base_cfg()
    b = synth.make_batch(4, 64, layout="unpacked", vocab=synth.TOY_VOCAB, seed=11)
    return "pretrain", cfg, dict(input_ids=t(b["input_ids"]), attention_mask=t(b["attention_mask"]), labels=t(b["labels"]))


@case("c2_smtp_stacked_2d")
def _c2():
    """C2-shaped: 2L/64d, F=13, V=756, unpacked right-padded."""
    cfg = base_cfg(vocab_size=756, hidden_size=64, intermediate_size=256, stacked_feat=13, next_n_token=13)
    b = synth.make_batch(3, 64, layout="unpacked", seed=12)
    return "pretrain", cfg, dict(input_ids=t(b["input_ids"]), attention_mask=t(b["attention_mask"]), labels=t(b["labels"]))


@case("c2_smtp_stacked_packed3d")
def _c2p():
    """C2 packed: block-diagonal [N,S,S] mask (tokenizer_utils.py:351-355)."""
    cfg = base_cfg(vocab_size=756, hidden_size=128, intermediate_size=512, stacked_feat=13, next_n_token=13)
    b = synth.make_batch(2, 96, layout="packed", seed=13)
    return "pretrain", cfg, dict(input_ids=t(b["input_ids"]), attention_mask=t(b["attention_mask"]), labels=t(b["labels"]))


@case("c5_ntp_causal")
def _c5():
    """C5-shaped: causal NTP, 2-D padding mask combined with causal inside the backbone."""
    cfg = base_cfg(vocab_size=756, hidden_size=64, intermediate_size=256, stacked_feat=13, next_n_token=13,
                   causal_attention=True)
    b = synth.make_batch(3, 64, layout="unpacked", task="ntp", seed=14)
    lab = b["labels"].copy()
    lab[b["attention_mask"] == 0] = -100
    return "pretrain", cfg, dict(input_ids=t(b["input_ids"]), attention_mask=t(b["attention_mask"]), labels=t(lab))


@case("c2_dlm_weighted")
def _dlm():
    """dLM-weighted SMTP loss (sample_wgt given -> per-feat lvl head + sum/(N*S*F), modeling_pretrain.py:229-236)."""
    cfg = base_cfg(vocab_size=756, hidden_size=64, intermediate_size=256, stacked_feat=13, next_n_token=13)
    b = synth.make_batch(3, 64, layout="unpacked", seed=15)
    wgt = torch.tensor([1.7, 0.4, 3.1])
    return "pretrain", cfg, dict(input_ids=t(b["input_ids"]), attention_mask=t(b["attention_mask"]), labels=t(b["labels"]),
                                 sample_wgt=wgt)


@case("c2_gated_agg")
def _gated():
    """gated stacked-feature aggregation (modeling_common.py:128-129)."""
    cfg = base_cfg(vocab_size=756, hidden_size=64, intermediate_size=256, stacked_feat=13, next_n_token=13,
                   stacked_feat_agg_method="gated")
    b = synth.make_batch(2, 48, layout="unpacked", seed=16)
    return "pretrain", cfg, dict(input_ids=t(b["input_ids"]), attention_mask=t(b["attention_mask"]), labels=t(b["labels"]))


@case("c2_focal_loss")
def _focal():
    """FocalLoss head (config.focal_gamma > 0, utils_graphgpt.py:340-377; unweighted branch modeling_pretrain.py:221-228)."""
    cfg = base_cfg(vocab_size=756, hidden_size=64, intermediate_size=256, stacked_feat=13, next_n_token=13, focal_gamma=2.0)
    b = synth.make_batch(2, 48, layout="unpacked", seed=18)
    return "pretrain", cfg, dict(input_ids=t(b["input_ids"]), attention_mask=t(b["attention_mask"]), labels=t(b["labels"]))


@case("c2_infer_all_entries")
def _infer():
    """labels=None: the head runs on every (n,s,f) entry (generation path, modeling_helpers.py:284-292)."""
    cfg = base_cfg(vocab_size=756, hidden_size=64, intermediate_size=256, stacked_feat=13, next_n_token=13)
    b = synth.make_batch(2, 32, layout="unpacked", seed=17)
    return "pretrain", cfg, dict(input_ids=t(b["input_ids"]), attention_mask=t(b["attention_mask"]))


@case("c3_ft_edge_cls")
def _ft():
    """C3-shaped fine-tune: F=4, 2 labels, last-token pooling, position_ids passed (training_utils.py:136-145)."""
    cfg = base_cfg(vocab_size=1200, hidden_size=128, intermediate_size=512, stacked_feat=4, next_n_token=4,
                   num_labels=2, problem_type="single_label_classification", pooling_method="last")
    vocab = synth.VocabLayout(vocab_size=1200, scope=512, n_node_attr=2, n_edge_attr=1)
    b = synth.make_batch(4, 64, layout="unpacked", task="ntp", vocab=vocab, seed=18)
    labels = torch.tensor([1, 0, 1, 1])
    return "finetune", cfg, dict(input_ids=t(b["input_ids"]), attention_mask=t(b["attention_mask"]),
                                 position_ids=t(b["position_ids"]), task_labels=labels)


@case("c3_ft_layerscale")
def _ft_ls():
    """ppa fine-tune variant with LayerScale (lsi=1, examples/edge_lvl/ppa_supervised.sh:22-25), eval mode so
    DropPath is the identity.  Needs the dropout backbone utils_graphgpt.LlamaModel."""
    cfg = base_cfg(vocab_size=1200, hidden_size=64, intermediate_size=256, stacked_feat=4, next_n_token=4,
                   num_labels=2, problem_type="single_label_classification", pooling_method="last",
                   layer_scale_init_value=1.0, path_pdrop=0.2)
    vocab = synth.VocabLayout(vocab_size=1200, scope=512, n_node_attr=2, n_edge_attr=1)
    b = synth.make_batch(3, 64, layout="unpacked", task="ntp", vocab=vocab, seed=19)
    labels = torch.tensor([0, 1, 1])
    return "finetune", cfg, dict(input_ids=t(b["input_ids"]), attention_mask=t(b["attention_mask"]),
                                 position_ids=t(b["position_ids"]), task_labels=labels)


@case("c2_raw_embed_pretrain")
def _raw_pt():
    """Raw-embedding input branch in pre-training (embed_dim = 24): mask-token swap on fully labelled rows,
    embed_layernorm, embed_proj, added to the token-embedding sum (modeling_pretrain.py:119-150)."""
    cfg = base_cfg(vocab_size=756, hidden_size=64, intermediate_size=256, stacked_feat=13, next_n_token=13, embed_dim=24)
    b = synth.make_batch(3, 64, layout="unpacked", seed=21)
    lab = b["labels"].copy()
    g = np.random.default_rng(5)
    full = (g.random(lab.shape[:2]) < 0.3) & (b["attention_mask"] == 1)      # some rows with every feature labelled
    lab[full] = np.where(lab[full] == -100, b["input_ids"][full], lab[full])
    raw = torch.from_numpy(g.standard_normal(lab.shape[:2] + (24,)).astype(np.float32)) * 1.5
    return "pretrain", cfg, dict(input_ids=t(b["input_ids"]), attention_mask=t(b["attention_mask"]), labels=t(lab),
                                 inputs_raw_embeds=raw)


@case("c3_raw_embed_ft")
def _raw_ft():
    """Raw-embedding input branch in fine-tuning (embed_dim = 16; modeling_finetune.py:118-135)."""
    cfg = base_cfg(vocab_size=1200, hidden_size=64, intermediate_size=256, stacked_feat=4, next_n_token=4,
                   num_labels=2, problem_type="single_label_classification", pooling_method="last", embed_dim=16)
    vocab = synth.VocabLayout(vocab_size=1200, scope=512, n_node_attr=2, n_edge_attr=1)
    b = synth.make_batch(3, 48, layout="unpacked", task="ntp", vocab=vocab, seed=22)
    g = np.random.default_rng(6)
    raw = torch.from_numpy(g.standard_normal(b["input_ids"].shape[:2] + (16,)).astype(np.float32))
    return "finetune", cfg, dict(input_ids=t(b["input_ids"]), attention_mask=t(b["attention_mask"]),
                                 task_labels=torch.tensor([1, 0, 0]), inputs_raw_embeds=raw)


@case("c3_ft_token_ce")
def _ft_tok():
    """Node-level fine-tune: loss_type token_ce, score on every position, CE over labelled positions
    (modeling_finetune.py:161-165,195-199)."""
    cfg = base_cfg(vocab_size=1200, hidden_size=64, intermediate_size=256, stacked_feat=4, next_n_token=4,
                   num_labels=7, problem_type="single_label_classification", pooling_method="last", loss_type="token_ce")
    vocab = synth.VocabLayout(vocab_size=1200, scope=512, n_node_attr=2, n_edge_attr=1)
    b = synth.make_batch(3, 48, layout="unpacked", task="ntp", vocab=vocab, seed=23)
    g = np.random.default_rng(8)
    lab = g.integers(0, 7, size=b["attention_mask"].shape).astype(np.int64)
    lab[(b["attention_mask"] == 0) | (g.random(lab.shape) < 0.4)] = -100
    return "finetune", cfg, dict(input_ids=t(b["input_ids"]), attention_mask=t(b["attention_mask"]), task_labels=t(lab))


@case("c3_ft_token_ce_intra")
def _ft_tok_intra():
    """Node-level fine-tune with intra-instance label embeddings: loss_type token_ce_intra, cls_idx [N]
    (modeling_finetune.py:137-165)."""
    cfg = base_cfg(vocab_size=1200, hidden_size=64, intermediate_size=256, stacked_feat=4, next_n_token=4,
                   num_labels=5, problem_type="single_label_classification", pooling_method="last",
                   loss_type="token_ce_intra")
    vocab = synth.VocabLayout(vocab_size=1200, scope=512, n_node_attr=2, n_edge_attr=1)
    b = synth.make_batch(3, 48, layout="unpacked", task="ntp", vocab=vocab, seed=27)
    g = np.random.default_rng(12)
    lab = g.integers(0, 5, size=b["attention_mask"].shape).astype(np.int64)
    lab[(b["attention_mask"] == 0) | (g.random(lab.shape) < 0.4)] = -100
    lens = b["attention_mask"].sum(-1)
    cls_idx = np.array([int(g.integers(0, max(1, l - 5))) for l in lens], dtype=np.int64)
    return "finetune", cfg, dict(input_ids=t(b["input_ids"]), attention_mask=t(b["attention_mask"]), task_labels=t(lab),
                                 cls_idx=t(cls_idx))


@case("c3_raw_edge_embed_ft")
def _raw_edge_ft():
    """Fine-tune with raw EDGE embeddings [N,S,S,E]: embed_layernorm + embed_proj per (n,s,s') row, summed over s'
    (modeling_helpers.py:127-139)."""
    cfg = base_cfg(vocab_size=1200, hidden_size=64, intermediate_size=256, stacked_feat=4, next_n_token=4, embed_dim=16,
                   num_labels=2, problem_type="single_label_classification", pooling_method="last")
    vocab = synth.VocabLayout(vocab_size=1200, scope=512, n_node_attr=2, n_edge_attr=1)
    b = synth.make_batch(3, 24, layout="unpacked", task="ntp", vocab=vocab, seed=28)
    g = np.random.default_rng(13)
    N, S = b["attention_mask"].shape
    raw = g.standard_normal((N, S, S, 16)).astype(np.float32)
    raw[g.random((N, S, S)) < 0.6] = 0.0                       # most (s, s') pairs carry no edge
    am = b["attention_mask"].astype(bool)
    raw[~am] = 0.0
    raw[:, :, :, :][~(am[:, :, None] & am[:, None, :])] = 0.0
    return "finetune", cfg, dict(input_ids=t(b["input_ids"]), attention_mask=t(b["attention_mask"]),
                                 task_labels=torch.tensor([1, 0, 1]), inputs_raw_embeds=torch.from_numpy(raw))


@case("c3_ft_double_heads")
def _ft_double():
    """GraphGPTDoubleHeadsModel with use_aux: edge-level task loss + auxiliary LM loss over pretrain_labels [N,S]
    (modeling_finetune.py:329-423); gradients are those of task_loss + pretrain_loss."""
    cfg = base_cfg(vocab_size=1200, hidden_size=64, intermediate_size=256, stacked_feat=4, next_n_token=4,
                   num_labels=2, problem_type="single_label_classification", pooling_method="last", use_aux=True)
    vocab = synth.VocabLayout(vocab_size=1200, scope=512, n_node_attr=2, n_edge_attr=1)
    b = synth.make_batch(3, 48, layout="unpacked", task="ntp", vocab=vocab, seed=24)
    g = np.random.default_rng(9)
    pl = b["input_ids"][:, :, 0].copy()
    pl[(b["attention_mask"] == 0) | (g.random(pl.shape) < 0.5)] = -100
    return "double", cfg, dict(input_ids=t(b["input_ids"]), attention_mask=t(b["attention_mask"]),
                               task_labels=torch.tensor([1, 1, 0]), pretrain_labels=t(pl))


@case("c2_long_stack")
def _long():
    """stack_method "long": embedding sum rescaled by 1/#non-pad features (modeling_helpers.py:106-110) and the per-feat
    head with per-sample normalised weights + dLM reduction (modeling_helpers.py:327-342, modeling_pretrain.py:229-236)."""
    cfg = base_cfg(vocab_size=756, hidden_size=64, intermediate_size=256, stacked_feat=13, next_n_token=13,
                   stack_method="long")
    b = synth.make_batch(3, 48, layout="unpacked", seed=27)
    ids = b["input_ids"].copy()
    g = np.random.default_rng(12)
    ids[(g.random(ids.shape) < 0.2) & (b["labels"] == -100)] = 0          # pad entries inside valid rows -> nnz varies
    return "pretrain", cfg, dict(input_ids=t(ids), attention_mask=t(b["attention_mask"]), labels=t(b["labels"]))


@case("c2_rope_range")
def _rope_range():
    """rope_range > 0: position_ids rescaled to [0, rope_range) per sample before RoPE (utils_graphgpt.py:574-581)."""
    cfg = base_cfg(vocab_size=756, hidden_size=64, intermediate_size=256, stacked_feat=13, next_n_token=13, rope_range=50)
    b = synth.make_batch(3, 48, layout="unpacked", seed=28)
    return "pretrain", cfg, dict(input_ids=t(b["input_ids"]), attention_mask=t(b["attention_mask"]), labels=t(b["labels"]),
                                 position_ids=t(b["position_ids"]))


def _ft_graph(seed, n=4, s=48):
    vocab = synth.VocabLayout(vocab_size=756, scope=512, n_node_attr=9, n_edge_attr=3)
    b = synth.make_batch(n, s, layout="unpacked", task="ntp", vocab=vocab, seed=seed)
    return dict(input_ids=t(b["input_ids"]), attention_mask=t(b["attention_mask"]))


@case("ft_graph_regression_l1")
def _ft_reg():
    """PCQM4M-v2 fine-tune head: 1 label, regression, L1 loss, score with bias (examples/graph_lvl/pcqm4m_v2_supervised.sh:74-76)."""
    cfg = base_cfg(vocab_size=756, hidden_size=64, intermediate_size=256, stacked_feat=13, next_n_token=13,
                   num_labels=1, problem_type="regression", loss_type="l1", pooling_method="last")
    return "finetune", cfg, dict(_ft_graph(25), task_labels=torch.tensor([0.3, -1.2, 2.5, 0.0]))


@case("ft_cls_sample_wgt")
def _ft_wgt():
    """single-label CE with per-sample weights (modeling_finetune.py:213-227) on a 14-class graph-level head."""
    cfg = base_cfg(vocab_size=756, hidden_size=64, intermediate_size=256, stacked_feat=13, next_n_token=13,
                   num_labels=14, problem_type="single_label_classification", pooling_method="last")
    return "finetune", cfg, dict(_ft_graph(29), task_labels=torch.tensor([3, 0, 13, 7]),
                                 sample_wgt=torch.tensor([0.5, 2.0, 1.0, 0.25]))


@case("ft_graph_regression_mse3")
def _ft_mse():
    """regression with 3 targets and the default (MSE) loss (modeling_finetune.py:181-192)."""
    cfg = base_cfg(vocab_size=756, hidden_size=64, intermediate_size=256, stacked_feat=13, next_n_token=13,
                   num_labels=3, problem_type="regression", pooling_method="last")
    g = np.random.default_rng(13)
    return "finetune", cfg, dict(_ft_graph(30), task_labels=torch.from_numpy(g.standard_normal((4, 3)).astype(np.float32)))


@case("ft_graph_multilabel_bce_mlp")
def _ft_ml():
    """molpcba-style head: 8 labels, multi-label BCE with NaN = unlabelled (molpcba_supervised.sh:74-76), MLP score head
    (config.mlp = [32]; src/utils/modules_utils.py:8-34)."""
    cfg = base_cfg(vocab_size=756, hidden_size=64, intermediate_size=256, stacked_feat=13, next_n_token=13,
                   num_labels=8, problem_type="multi_label_classification", pooling_method="last", mlp=[32])
    g = np.random.default_rng(11)
    lab = (g.random((4, 8)) < 0.4).astype(np.float32)
    lab[g.random((4, 8)) < 0.25] = np.nan
    return "finetune", cfg, dict(_ft_graph(26), task_labels=torch.from_numpy(lab))


def _patch_dropout_backbone():
    """transformers 5.5.0's LlamaModel loop expects decoder layers to return a tensor; the reference's dropout
    layer (utils_graphgpt.py:168-173) returns the 4.53-style tuple.  Unwrap it (SURVEY §8c caveat)."""
    from src.models.graphgpt import utils_graphgpt
    if getattr(utils_graphgpt.LlamaDecoderLayer, "_ggpt_patched", False):
        return
    orig = utils_graphgpt.LlamaDecoderLayer.forward

    def fwd(self, hidden_states, *a, **k):
        k.pop("past_key_values", None)
        out = orig(self, hidden_states, *a, **k)
        return out[0] if isinstance(out, tuple) else out

    utils_graphgpt.LlamaDecoderLayer.forward = fwd
    utils_graphgpt.LlamaDecoderLayer._ggpt_patched = True


def make_smtp_inside():
    """tests/golden/aux/smtp_inside.pt: the reference's prepare_for_2d_smtp_inputs_labels (modeling_helpers.py:399-452)
    run on CPU with a seeded generator; the uniform draws it consumes are re-created with the same seed / order /
    shapes and stored next to its outputs, so the oracle and the kernel can be fed identical randomness."""
    from src.models.graphgpt import modeling_helpers as mh
    recs = []
    for seed, (N, S, F_, power) in enumerate([(3, 40, 13, 1.0), (2, 64, 4, 0.5), (4, 24, 1, 2.0)]):
        g = np.random.default_rng(100 + seed)
        ids = g.integers(1, 700, size=(N, S, F_ + 4)).astype(np.int64)
        n_nodes = g.integers(3, S // 2, size=N)
        for n in range(N):
            ln = int(g.integers(S // 2, S + 1))
            ids[n, :ln, F_ + 2] = g.integers(0, n_nodes[n], size=ln)
            ids[n, ln:, :] = 0                                         # pad rows: ids 0, node index 0
            ids[n, :ln, :F_][g.random((ln, F_)) < 0.1] = 0             # some pad entries inside valid rows
        ids = torch.from_numpy(ids)
        torch.manual_seed(4242 + seed)
        torch.rand((N, 1, 1))
        mr = torch.rand((N, 1, 1))
        u = torch.rand((N, S, F_))
        torch.manual_seed(4242 + seed)
        out_ids, labels = mh.prepare_for_2d_smtp_inputs_labels(ids[:, :, :F_], ids[:, :, F_ + 2], smtp_2d_rate=1, power=power,
                                                               replace_rate=0, vocab=756, global_2d_mask=False)
        recs.append(dict(ids=ids, F=F_, power=power, mr=mr.view(-1).clone(), u_node=u.clone(), seed=4242 + seed,
                         out_ids=out_ids.clone(), labels=labels.clone()))
        print(f"smtp_inside case {seed}: masked {(labels != -100).float().mean():.3f} of entries")
    os.makedirs(os.path.join(HERE, "aux"), exist_ok=True)
    torch.save(recs, os.path.join(HERE, "aux", "smtp_inside.pt"))


def main():
    mp, mf, GraphGPTConfig = load_reference()
    _patch_dropout_backbone()
    only = set(sys.argv[1:])
    if not only or "smtp_inside" in only:
        make_smtp_inside()
    for name, fn in CASES.items():
        if only and name not in only:
            continue
        kind, cfgd, inputs = fn()
        cfg = GraphGPTConfig(**cfgd)
        cfg._attn_implementation = "eager"
        torch.manual_seed(1000 + len(name))
        klass = {"pretrain": mp.GraphGPTPretrainBase, "finetune": mf.GraphGPTTaskModel,
                 "double": mf.GraphGPTDoubleHeadsModel}[kind]
        model = klass(cfg).eval().float()
        with torch.no_grad():  # make norm weights / lambdas non-trivial so their gradients and scaling are exercised
            for n_, p in model.named_parameters():
                if "layernorm" in n_ or n_.endswith("norm.weight"):
                    p.add_(0.1 * torch.randn_like(p))
                if n_ == "emb_mask_token":
                    p.add_(0.5 * torch.randn_like(p))
                if "lambda_" in n_:
                    p.mul_(1.0 + 0.2 * torch.randn_like(p))
        out = model(**inputs)
        rec = {"kind": kind, "config": cfgd, "inputs": inputs,
               "state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()}}
        if kind == "pretrain":
            lg = out.head1_logits.detach().clone()
            rec["logits_shape"] = tuple(lg.shape)
            rec["logits_rowsum"] = lg.double().sum(-1)          # pins every row
            rec["logits_stride"] = 4 if lg.numel() > 100_000 else 1
            rec["logits"] = lg[:: rec["logits_stride"]].clone()  # full rows, every stride-th row
            loss = out.head1_loss
        elif kind == "double":
            rec["task_logits"] = out.task_logits.detach().clone()
            rec["pretrain_logits_rowsum"] = out.pretrain_logits.detach().double().sum(-1)
            rec["task_loss"], rec["pretrain_loss"] = out.task_loss.detach().clone(), out.pretrain_loss.detach().clone()
            loss = out.task_loss + out.pretrain_loss
        else:
            rec["task_logits"] = out.task_logits.detach().clone()
            rec["hidden"] = out.hidden_states.detach().clone()
            rec["task_hidden"] = out.task_hidden_states.detach().clone()
            loss = out.task_loss
        if loss is not None:
            rec["loss"] = loss.detach().clone()
            loss.backward()
            named = dict(model.named_parameters())
            rec["grads"], rec["grad_norms"] = {}, {}
            for k in GRAD_KEYS:
                if k in named and named[k].grad is not None:
                    g = named[k].grad.detach()
                    rec["grad_norms"][k] = float(g.double().norm())     # pins the whole tensor
                    rec["grads"][k] = g[:24].clone() if g.dim() == 2 else g.clone()   # first 24 rows of matrices
        # How far is the REFERENCE ITSELF from its fp32 result when run in bf16 (the precision the reference trains
        # in under DeepSpeed bf16)?  Only the error numbers are stored; they bound what bf16 kernels can achieve.
        import copy
        mb = copy.deepcopy(model).to(torch.bfloat16)
        for p_ in mb.parameters():
            p_.grad = None
        inputs_b = {k: (v.to(torch.bfloat16) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in inputs.items()}
        ob = mb(**inputs_b)
        bf = {}
        if kind == "pretrain":
            lb, lgb = ob.head1_loss, ob.head1_logits.float()
            bf["logits_relF"] = float((lgb - out.head1_logits.detach()).norm() / out.head1_logits.detach().norm())
        elif kind == "double":
            lb = ob.task_loss + ob.pretrain_loss
        else:
            lb = ob.task_loss
            bf["task_hidden_relF"] = float((ob.task_hidden_states.float() - out.task_hidden_states.detach()).norm()
                                           / out.task_hidden_states.detach().norm())
        if loss is not None:
            bf["loss_rel"] = abs(float(lb) - float(loss)) / abs(float(loss))
            lb.backward()
            nb = dict(mb.named_parameters())
            bf["grad_relF"] = {}
            for k in rec["grads"]:
                g32 = named[k].grad.detach()
                g16 = nb[k].grad.detach().float()
                a, b_ = (g16[:24], g32[:24]) if g32.dim() == 2 else (g16, g32)
                bf["grad_relF"][k] = float((a - b_).norm() / (b_.norm() + 1e-30))
        rec["ref_bf16_error"] = bf
        path = os.path.join(HERE, name + ".pt")
        torch.save(rec, path)
        print(f"{name}: loss={None if loss is None else float(loss):.6f}" if loss is not None else f"{name}: no loss",
              {k: tuple(v.shape) for k, v in rec.items() if torch.is_tensor(v)}, f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
