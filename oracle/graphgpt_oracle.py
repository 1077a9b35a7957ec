"""CPU oracle for the GraphGPT transformer hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A dependency-free (torch CPU tensors only; no transformers, no reference import) restatement of the arithmetic the
reference executes per step.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file, and only as the checker / reported baseline.  The product path (graph-gpt_b200/) never
imports it and has no CPU fallback.

Where the algorithm lives: the reference's model classes (ref: /root/reference/src/models/graphgpt/*.py) delegate
the backbone to an UN-VENDORED third-party dependency, `transformers==4.53.3` (requirements.txt:19), module
transformers/models/llama/modeling_llama.py ("HF:" below; line numbers are those of the transformers 5.5.0 copy
installed in the build container, whose Llama arithmetic is identical).  Each function cites what it follows.

PARITY PIN: pinned.  The reference has no tests or golden vectors for this path, so the oracle is
pinned against outputs of the reference ITSELF, run in the build container through tests/golden/ref_shim.py by
tests/golden/make_golden.py; the resulting fixtures tests/golden/*.pt hold inputs, the reference state_dict and
the reference's loss / logits / hidden states / gradients.  tests/test_oracle_golden.py checks this file against
them to fp32 round-off (<= 2e-5 relative).

Numerics: everything fp32, in the reference's op order where order matters for parity:
  RMSNorm: w * (x * rsqrt(mean(x^2) + eps))                                  HF:59-64
  RoPE   : inv_freq = theta^(-2i/64); cos/sin(pos * inv_freq), rotate-half   HF:117-135,138-168
  attn   : softmax_fp32(q k^T / sqrt(64) + additive_mask) v                  HF:199-221
  MLP    : down(gelu_erf(gate(x)) * up(x))                                   HF:182-184
  block  : pre-norm residual, optional LayerScale lambda_1/2                 HF:313-332, utils_graphgpt.py:137-166
"""
import math

import torch
import torch.nn.functional as F

_EPSILON = 1e-7  # modeling_common.py:44

# Two optional departures from plain fp32, both used by tests only to EXPLAIN the distance between the bf16 tensor-core
# kernels and the fp32 reference (tests/test_parity_depth_gpu.py):
#   dtype=torch.bfloat16  — the reference as DeepSpeed-bf16 runs it: parameters, activations and the residual stream in
#                           bf16, RMSNorm / softmax / RoPE tables / CE internally fp32 exactly as HF:62-67,129-135,216.
#   emulate_bf16=True     — fp32 arithmetic with a round-to-bf16 at exactly the points where the sm_100a kernels store a
#                           bf16 tensor (GEMM weights, h = norm(x), q|k|v after RoPE, the un-normalised
#                           softmax numerators P, the attention output, the o_proj / down_proj outputs, act = gelu(g)*u,
#                           n_token_proj's output); the residual stream, statistics and accumulations stay fp32.
_GEMM_WEIGHT_SUFFIXES = ("_proj.weight", "lm_head.weight", "n_token_proj.weight")


def _ste_bf16(t):
    """round-to-nearest-even to bf16, straight-through gradient"""
    return t + (t.detach().bfloat16().to(t.dtype) - t.detach())


def _ident(t):
    return t


# ------------------------------------------------------------------------------------------------
# config helper
# ------------------------------------------------------------------------------------------------
class OracleConfig:
    """The handful of GraphGPTConfig fields the arithmetic depends on (configuration_graphgpt.py:26-110)."""

    def __init__(self, **kw):
        self.vocab_size = kw["vocab_size"]
        self.hidden_size = kw["hidden_size"]
        self.intermediate_size = kw.get("intermediate_size", 4 * self.hidden_size)
        self.num_hidden_layers = kw["num_hidden_layers"]
        self.head_dim = kw.get("head_dim", 64)
        self.num_attention_heads = kw.get("num_attention_heads", self.hidden_size // self.head_dim)
        self.rms_norm_eps = kw.get("rms_norm_eps", 1e-6)
        self.rope_theta = kw.get("rope_theta", 10000.0)
        self.causal_attention = kw.get("causal_attention", False)
        self.stacked_feat = kw.get("stacked_feat", 1)
        self.next_n_token = kw.get("next_n_token", self.stacked_feat)
        self.stack_method = kw.get("stack_method", "short")
        self.stacked_feat_agg_method = kw.get("stacked_feat_agg_method", "sum")
        self.layer_scale_init_value = kw.get("layer_scale_init_value", 0.0)
        self.rope_range = kw.get("rope_range", 0)
        self.pad_token_id = kw.get("pad_token_id", 0)
        self.num_labels = kw.get("num_labels", 2)
        self.problem_type = kw.get("problem_type", None)
        self.loss_type = kw.get("loss_type", None)
        self.focal_gamma = kw.get("focal_gamma", 0.0)

    @classmethod
    def from_any(cls, cfg):
        if isinstance(cfg, cls):
            return cfg
        keys = ["vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "head_dim",
                "num_attention_heads", "rms_norm_eps", "rope_theta", "causal_attention", "stacked_feat",
                "next_n_token", "stack_method", "stacked_feat_agg_method", "layer_scale_init_value", "rope_range",
                "pad_token_id", "num_labels", "problem_type", "loss_type", "focal_gamma"]
        get = (lambda k: cfg.get(k)) if isinstance(cfg, dict) else (lambda k: getattr(cfg, k, None))
        kw = {k: get(k) for k in keys if get(k) is not None}
        if "rope_theta" not in kw:
            rp = get("rope_parameters")
            if rp:
                kw["rope_theta"] = rp["rope_theta"]
        return cls(**kw)


# ------------------------------------------------------------------------------------------------
# backbone pieces
# ------------------------------------------------------------------------------------------------
def rmsnorm(x, w, eps):
    """HF:59-64  LlamaRMSNorm.forward (fp32 throughout here)."""
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(-1, keepdim=True)
    return w * (xf * torch.rsqrt(var + eps)).to(x.dtype)


def rope_cos_sin(position_ids, head_dim, theta, dtype=torch.float32):
    """HF:117-135.  position_ids [N,S] -> cos, sin [N,S,head_dim] (computed in fp32, cast to the activation dtype;
    emb = cat(freqs, freqs))."""
    dev = position_ids.device
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64, device=dev).float() / head_dim))
    freqs = position_ids[:, :, None].float() * inv_freq[None, None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x):
    """HF:138-142"""
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rope(q, k, cos, sin):
    """HF:146-168; q,k [N,H,S,hd], cos/sin [N,S,hd]."""
    cos, sin = cos[:, None], sin[:, None]
    return q * cos + rotate_half(q) * sin, k * cos + rotate_half(k) * sin


def additive_mask(attention_mask, seq_len, causal, dtype=torch.float32, device=None):
    """Additive [N,1,S,S] mask with 0 / finfo.min.
    bidirectional 2-D: transformers _prepare_4d_attention_mask (modeling_helpers.py:41-42)
    bidirectional 3-D: _expand_mask_from_3d_mask (modeling_helpers.py:51-64)
    causal (config.causal_attention): the 2-D mask goes to LlamaModel, which builds causal AND key-padding
    (HF:398-405 create_causal_mask)."""
    neg = torch.finfo(dtype).min
    device = attention_mask.device if attention_mask is not None else device
    if attention_mask is None:
        keep = torch.ones((1, 1, seq_len, seq_len), dtype=torch.bool, device=device)
    elif attention_mask.dim() == 2:
        keep = attention_mask[:, None, None, :].bool().expand(-1, 1, seq_len, -1)
    elif attention_mask.dim() == 3:
        keep = attention_mask[:, None, :, :].bool()
    else:
        raise NotImplementedError(f"attention_mask of shape {tuple(attention_mask.shape)} is not Implemented")
    if causal:
        tri = torch.ones((seq_len, seq_len), dtype=torch.bool, device=device).tril()
        keep = keep & tri[None, None]
    out = torch.zeros(keep.shape, dtype=dtype, device=device)
    return out.masked_fill(~keep, neg)


def attention(q, k, v, mask4d, scaling, rnd=None):
    """HF:199-221 eager_attention_forward, fp32 softmax.  rnd (emulate_bf16): the kernels feed the un-normalised
    numerators exp(s - max) to the PV product in bf16 and divide by the fp32 sum of the UNROUNDED numerators."""
    w = torch.matmul(q, k.transpose(2, 3)) * scaling
    w = w + mask4d
    if rnd is not None:
        e = torch.exp(w - w.max(dim=-1, keepdim=True)[0])
        o = torch.matmul(rnd(e), v) / e.sum(dim=-1, keepdim=True)
        return o.transpose(1, 2).contiguous()
    w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    o = torch.matmul(w, v)
    return o.transpose(1, 2).contiguous()


def decoder_layer(x, sd, prefix, cfg, cos, sin, mask4d, act_scale=None, mlp_scale=None, path_scale=None, rnd=None):
    """HF:313-332 LlamaDecoderLayer.forward (+ LayerScale of utils_graphgpt.py:153-166).  act_scale / mlp_scale are the
    training-mode nn.Dropout factors keep/(1-p) of mlp_act_dropout [N,S,I] and mlp_dropout [N,S,d]
    (utils_graphgpt.py:69-83) supplied by the caller — None = eval mode (identity).  path_scale = (s1, s2): the DropPath
    factors floor(keep + u)/keep per sample ([N]) of the attention and the MLP branch (utils_graphgpt.py:156,166)."""
    N, S, d = x.shape
    H, hd = cfg.num_attention_heads, cfg.head_dim
    r = rnd or _ident                  # emulate_bf16: the kernels' bf16 stores (weights are pre-rounded by the caller)
    h = r(rmsnorm(x, sd[prefix + "input_layernorm.weight"], cfg.rms_norm_eps))
    q = F.linear(h, sd[prefix + "self_attn.q_proj.weight"]).view(N, S, H, hd).transpose(1, 2)   # HF:262
    k = F.linear(h, sd[prefix + "self_attn.k_proj.weight"]).view(N, S, H, hd).transpose(1, 2)
    v = F.linear(h, sd[prefix + "self_attn.v_proj.weight"]).view(N, S, H, hd).transpose(1, 2)
    q, k = apply_rope(q, k, cos, sin)
    q, k, v = r(q), r(k), r(v)
    a = r(attention(q, k, v, mask4d, hd ** -0.5, rnd).reshape(N, S, H * hd))
    a = r(F.linear(a, sd[prefix + "self_attn.o_proj.weight"]))
    if prefix + "lambda_1" in sd:
        a = sd[prefix + "lambda_1"] * a
    if path_scale is not None and path_scale[0] is not None:
        a = a * path_scale[0][:, None, None]
    x = x + a
    h = r(rmsnorm(x, sd[prefix + "post_attention_layernorm.weight"], cfg.rms_norm_eps))
    g = F.linear(h, sd[prefix + "mlp.gate_proj.weight"])
    u = F.linear(h, sd[prefix + "mlp.up_proj.weight"])
    act = r(F.gelu(g) * u)                                             # exact erf GELU, HF:182-184
    if act_scale is not None:
        act = act * act_scale                                          # utils_graphgpt.py:80
    m = r(F.linear(act, sd[prefix + "mlp.down_proj.weight"]))
    if mlp_scale is not None:
        m = m * mlp_scale                                              # utils_graphgpt.py:81
    if prefix + "lambda_2" in sd:
        m = sd[prefix + "lambda_2"] * m
    if path_scale is not None and path_scale[1] is not None:
        m = m * path_scale[1][:, None, None]
    return x + m


def reset_pos_ids(position_ids, cfg):
    """utils_graphgpt.py:574-581"""
    if position_ids is not None and cfg.rope_range > 0:
        mx = position_ids.max(dim=-1, keepdim=True)[0] + 1
        position_ids = position_ids.float() * cfg.rope_range / mx.float()
    return position_ids


def stacked_embed(input_ids, sd, cfg, embed_scale=None):
    """modeling_helpers.py:89-114 + StackedFeatAggregation.forward modeling_common.py:127-135.  embed_scale: the
    embed_dropout factors keep/(1-p) on the gathered rows ([N,S,F,d] / [N,S,d]; :97-98), None = eval."""
    # nn.Embedding(padding_idx=pad_token_id) (HF:361): an ordinary look-up of row 0 in forward, no gradient into it
    emb = F.embedding(input_ids, sd["model.embed_tokens.weight"], padding_idx=cfg.pad_token_id)
    if embed_scale is not None:
        emb = emb * embed_scale
    if input_ids.dim() == 3:
        if cfg.stacked_feat_agg_method == "gated":
            emb = torch.einsum("nsfd,fd->nsd", emb, sd["stacked_feat_agg.weight"])
        else:
            emb = emb.sum(dim=-2)
        in_ = input_ids[:, :, 0]
        if cfg.stack_method == "long":
            nonzero = (input_ids != 0).sum(dim=-1, keepdim=True) + _EPSILON
            emb = emb * torch.clamp(1 / nonzero.to(emb.dtype), max=1)
    else:
        in_ = input_ids
    return emb, in_


def raw_embed_branch(inputs_raw_embeds, sd, cfg, labels=None, smtp_inside=False, raw_scale=None):
    """Raw-embedding input branch (config.embed_dim > 0).
    pre-training (labels given): modeling_pretrain.py:130-148 — rows where ANY examined label is -100 keep their raw
    features, the others take emb_mask_token; then embed_layernorm, raw_embed_dropout, embed_proj.
    fine-tuning (labels None): modeling_helpers.py:127-139 — embed_layernorm, dropout, embed_proj."""
    x = inputs_raw_embeds.float()
    if labels is not None:
        if smtp_inside:
            embed_mask = labels[:, :, 0:1] == -100                                       # :134-135
        else:
            embed_mask = (labels == -100).sum(dim=-1, keepdim=True).to(torch.bool)       # :136-137
        x = embed_mask.float() * x + (~embed_mask).float() * sd["emb_mask_token"]        # :139-143
    x = rmsnorm(x, sd["embed_layernorm.weight"], cfg.rms_norm_eps)
    if raw_scale is not None:
        x = x * raw_scale
    x = F.linear(x, sd["embed_proj.weight"])
    if x.dim() == 4:                          # [N,S,S,d] edge embeddings: summed over the third axis, :135-136
        x = x.sum(dim=-2)
    return x


def backbone(inputs_embeds, attention_mask, position_ids, sd, cfg, collect=None, drop=None, rnd=None):
    """HF:375-425 LlamaModel.forward over `inputs_embeds`; returns the final-norm hidden states [N,S,d].
    drop: optional {"act": [L x [N,S,I]], "mlp": [L x [N,S,d]]} dropout factors (training mode)."""
    N, S, _ = inputs_embeds.shape
    dev, dt = inputs_embeds.device, inputs_embeds.dtype
    if position_ids is None:
        position_ids = torch.arange(S, device=dev)[None, :].expand(N, -1)          # HF:394-397
    cos, sin = rope_cos_sin(position_ids, cfg.head_dim, cfg.rope_theta, dt)
    mask4d = additive_mask(attention_mask, S, cfg.causal_attention, dt, dev)
    x = inputs_embeds
    for i in range(cfg.num_hidden_layers):
        x = decoder_layer(x, sd, f"model.layers.{i}.", cfg, cos, sin, mask4d,
                          None if drop is None or "act" not in drop else drop["act"][i],
                          None if drop is None or "mlp" not in drop else drop["mlp"][i],
                          None if drop is None or "path" not in drop else drop["path"][i], rnd)
        if collect is not None:
            collect.append(x)
    return (rnd or _ident)(rmsnorm(x, sd["model.norm.weight"], cfg.rms_norm_eps))


# ------------------------------------------------------------------------------------------------
# pre-training head (SMTP / NTP)
# ------------------------------------------------------------------------------------------------
def head_select(hidden, labels, sd, cfg, sample_wgt=None, rnd=None):
    """prepare_for_stacked_feat_labels (modeling_helpers.py:362-393).

    short & wgt None -> _prepare_for_stacked_feat_labels_per_mix_lvl (:263-301): rows with any label, n_token_proj,
    then labelled (row, feature) entries in row-major order.
    short & wgt      -> per-feat lvl with per-sample weight (:345-359, 382-392).
    long             -> per-feat lvl with normalised weights (:327-342)."""
    d = hidden.shape[-1]
    F_ = cfg.next_n_token
    r = rnd or _ident
    proj = (lambda t: r(F.linear(t, sd["n_token_proj.weight"]))) if F_ > 1 else (lambda t: t)
    wgt = None
    if labels is None:
        h = proj(hidden.reshape(-1, d)).reshape(-1, d)
        return h, None, None
    if labels.dim() == 2:
        labels = labels[:, :, None]
    if cfg.stack_method == "long" or sample_wgt is not None:
        N, S, _ = hidden.shape
        h = proj(hidden).reshape(N, S, F_, d)
        mask_m = labels != -100
        if cfg.stack_method == "long":
            w = mask_m.float()
            w = w / (w.sum(dim=-1).sum(dim=-1)[:, None, None] + _EPSILON)
        else:
            w = sample_wgt[:, None, None].repeat(1, S, F_)
        return h[mask_m], labels[mask_m], w[mask_m]
    mask = labels != -100
    mask_m = mask.any(dim=-1)
    mask_sel = mask[mask_m]
    h = proj(hidden[mask_m]).reshape(-1, d)
    lab = labels[mask_m][mask_sel]
    h = h[mask_sel.reshape(-1)]
    return h, lab, wgt


def smtp_mask_2d(input_ids, node_idx, mr, u_node, power, mask_token_id=1, label_pad_token_id=-100):
    """prepare_for_2d_smtp_inputs_labels (modeling_helpers.py:399-452) for smtp_2d_rate=1, replace_rate=0,
    global_2d_mask=False (the arguments of modeling_pretrain.py:175-189), with the uniform draws passed in:
    mr [N] = mr_per_sample, u_node [N,S,F] = the mask_per_node draw.  Returns (input_ids, labels)."""
    N = input_ids.shape[0]
    mask_per_node = u_node > (mr.view(N, 1, 1) ** power)                      # :430-433 (sample_mask is all-true)
    bz_idx = torch.arange(N).view(-1, 1)
    mask_per_token = mask_per_node[bz_idx, node_idx] & (input_ids > 0)         # :437-438
    labels = input_ids.clone().masked_fill_(~mask_per_token, label_pad_token_id)
    out = input_ids.clone().masked_fill_(mask_per_token, mask_token_id)
    return out, labels


def focal_loss(logits, labels, gamma):
    """FocalLoss (utils_graphgpt.py:340-377), reduction "mean": -(1 - p_t)^gamma log p_t with p_t detached."""
    logpt = F.log_softmax(logits.float(), dim=-1).gather(1, labels.view(-1, 1)).view(-1)
    pt = logpt.detach().exp()
    return (-1 * (1 - pt) ** gamma * logpt).mean()


def ce_loss(logits, labels, wgt=None, dlm=False, focal_gamma=0.0):
    """_get_ce_loss (modeling_helpers.py:145-177) / _get_dlm_ce_loss (:180-198): CE on logits.float()."""
    if wgt is None:
        if focal_gamma > 0:
            return focal_loss(logits, labels, focal_gamma)
        return F.cross_entropy(logits.float(), labels)
    loss = F.cross_entropy(logits.float(), labels, reduction="none")
    w = wgt.view(-1).float()
    if dlm:
        return (loss * w).sum()
    return (loss * w).sum() / (w.sum() + _EPSILON)


def _prep_sd(sd, dtype, emulate_bf16):
    """fp32 (default) / bf16 copies of the parameters; emulate_bf16 rounds the GEMM weights (the kernels' bf16 compute
    copy) and leaves norms, LayerScale, gates and the embedding table (read in fp32 by embed_fwd) alone."""
    out = {}
    for k, v in sd.items():
        if not v.is_floating_point():
            out[k] = v
            continue
        v = v.to(dtype)
        if emulate_bf16 and (k.endswith(_GEMM_WEIGHT_SUFFIXES)):
            v = _ste_bf16(v)
        out[k] = v
    return out


def pretrain_forward(sd, cfg, input_ids, attention_mask=None, labels=None, sample_wgt=None, position_ids=None,
                     collect=None, inputs_raw_embeds=None, smtp_inside=False, drop=None, dtype=torch.float32,
                     emulate_bf16=False):
    """GraphGPTPretrainBase.forward (modeling_pretrain.py:152-266), generative head only.
    drop: optional training-mode dropout factors {"embed", "raw", "act", "mlp"} (see the helpers above).
    Returns dict(loss, logits, hidden)."""
    cfg = OracleConfig.from_any(cfg)
    sd = _prep_sd(sd, dtype, emulate_bf16)
    rnd = _ste_bf16 if emulate_bf16 else None
    position_ids = reset_pos_ids(position_ids, cfg)
    drop = drop or {}
    emb, _ = stacked_embed(input_ids, sd, cfg, drop.get("embed"))
    if inputs_raw_embeds is not None:
        emb = emb + raw_embed_branch(inputs_raw_embeds, sd, cfg, labels, smtp_inside, drop.get("raw"))   # :149
    hidden = backbone(emb, attention_mask, position_ids, sd, cfg, collect, drop, rnd)
    h, lab, wgt = head_select(hidden, labels, sd, cfg, sample_wgt, rnd)
    logits = F.linear(h, sd["lm_head.weight"])                                        # :218
    loss = None
    if lab is not None:
        if wgt is None:
            loss = ce_loss(logits, lab, focal_gamma=getattr(cfg, "focal_gamma", 0.0) or 0.0)   # :221-228
        else:
            N, S, _ = hidden.shape
            # wgt is not None for stack_method "long" (normalised weights) and for "short" + sample_wgt:
            # both take the dLM branch, modeling_pretrain.py:229-236
            loss = ce_loss(logits, lab, wgt, dlm=True) / (N * S * cfg.next_n_token)
    return {"loss": loss, "logits": logits, "hidden": hidden}


# ------------------------------------------------------------------------------------------------
# fine-tuning head
# ------------------------------------------------------------------------------------------------
def task_forward(sd, cfg, input_ids, attention_mask=None, position_ids=None, task_labels=None, sample_wgt=None,
                 inputs_raw_embeds=None, pretrain_labels=None, drop=None, dtype=torch.float32, emulate_bf16=False,
                 collect=None, cls_idx=None):
    """GraphGPTTaskModel.forward (modeling_finetune.py:236-326): score on all positions, pool at the last non-pad
    index (modeling_helpers.py:78-86), CE / MSE / L1 / BCE loss (modeling_finetune.py:167-234)."""
    cfg = OracleConfig.from_any(cfg)
    sd = _prep_sd(sd, dtype, emulate_bf16)
    position_ids = reset_pos_ids(position_ids, cfg)
    if input_ids.dim() == 3:
        input_ids = input_ids[:, :, : cfg.stacked_feat]
    emb, in_ = stacked_embed(input_ids, sd, cfg)
    if inputs_raw_embeds is not None:
        emb = emb + raw_embed_branch(inputs_raw_embeds, sd, cfg)                          # modeling_finetune.py:130-134
    hidden = backbone(emb, attention_mask, position_ids, sd, cfg, collect, drop, _ste_bf16 if emulate_bf16 else None)
    if "score.weight" in sd:
        logits = F.linear(hidden, sd["score.weight"], sd.get("score.bias"))
    else:       # MLP score head (src/utils/modules_utils.py:8-34): act -> (dropout) -> Linear, for every Linear
        logits, j = hidden, 0
        while f"score.mlp_modules.{j}.weight" in sd:
            logits = F.linear(F.gelu(logits), sd[f"score.mlp_modules.{j}.weight"], sd.get(f"score.mlp_modules.{j}.bias"))
            j += 1
    seq_len = (in_ != cfg.pad_token_id).sum(-1) - 1
    idx = torch.arange(hidden.shape[0], device=hidden.device)
    pooled_logits = logits[idx, seq_len]
    pooled_hidden = hidden[idx, seq_len]
    aux = {}
    if pretrain_labels is not None:
        # GraphGPTDoubleHeadsModel (modeling_finetune.py:402-411): auxiliary LM head over every position
        pl = F.linear(hidden, sd["lm_head.weight"])
        aux = {"pretrain_logits": pl,
               "pretrain_loss": F.cross_entropy(pl.float().view(-1, pl.shape[-1]), pretrain_labels.view(-1))}
    if cfg.loss_type == "token_ce_intra":
        # intra-instance label embeddings (modeling_finetune.py:137-165): cosine logits against the sample's own hidden
        # states at positions cls_idx .. cls_idx + num_labels - 1, temperature 1/20
        hn = F.normalize(hidden, dim=-1)
        N_ = hidden.shape[0]
        idx1 = torch.arange(N_, device=hidden.device).reshape(-1, 1).expand(-1, cfg.num_labels)
        idx2 = torch.arange(cfg.num_labels, device=hidden.device).reshape(1, -1).expand(N_, -1) + cls_idx.reshape(-1, 1)
        logits = torch.matmul(hn, hn[idx1, idx2].transpose(-2, -1)) * 20
        loss = None
        if task_labels is not None:
            loss = F.cross_entropy(logits.view(-1, cfg.num_labels).float(), task_labels.view(-1))
        return {"loss": loss, "task_logits": logits.float(), "hidden": hidden, "task_hidden": pooled_hidden, **aux}
    if cfg.loss_type == "token_ce":
        # token-level task (modeling_finetune.py:161-165,195-199): every position is scored, CE ignores -100
        loss = None
        if task_labels is not None:
            loss = F.cross_entropy(logits.view(-1, cfg.num_labels).float(), task_labels.view(-1))
        return {"loss": loss, "task_logits": logits.float(), "hidden": hidden, "task_hidden": pooled_hidden, **aux}
    loss = None
    if task_labels is not None:
        problem = cfg.problem_type
        if problem is None:
            if cfg.num_labels == 1:
                problem = "regression"
            elif task_labels.dtype in (torch.long, torch.int):
                problem = "single_label_classification"
            else:
                problem = "multi_label_classification"
        if problem == "regression":
            fn = F.l1_loss if cfg.loss_type == "l1" else F.mse_loss
            lab = task_labels.to(pooled_logits.dtype)
            loss = fn(pooled_logits.squeeze(), lab.squeeze()) if cfg.num_labels == 1 else fn(pooled_logits, lab)
        elif problem == "single_label_classification":
            if sample_wgt is None:
                loss = F.cross_entropy(pooled_logits.view(-1, cfg.num_labels).float(), task_labels.view(-1))
            else:
                l = F.cross_entropy(pooled_logits.view(-1, cfg.num_labels).float(), task_labels.view(-1), reduction="none")
                loss = (l * sample_wgt.float().view(-1)).sum() / sample_wgt.float().sum()
        else:
            ok = task_labels == task_labels
            loss = F.binary_cross_entropy_with_logits(pooled_logits[ok], task_labels[ok])
    return {"loss": loss, "task_logits": pooled_logits.float(), "hidden": hidden, "task_hidden": pooled_hidden, **aux}


# ------------------------------------------------------------------------------------------------
# utilities for tests / baselines
# ------------------------------------------------------------------------------------------------
def init_state_dict(cfg, seed=0, task_head=False, gated=False):
    """Random weights with the reference's parameter names/shapes and HF init normal(0, 0.02)."""
    cfg = OracleConfig.from_any(cfg)
    g = torch.Generator().manual_seed(seed)
    d, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size

    def n(*shape):
        return torch.randn(*shape, generator=g) * 0.02

    sd = {"model.embed_tokens.weight": n(V, d)}
    sd["model.embed_tokens.weight"][cfg.pad_token_id].zero_()
    for i in range(cfg.num_hidden_layers):
        p = f"model.layers.{i}."
        for nm in ("q", "k", "v", "o"):
            sd[p + f"self_attn.{nm}_proj.weight"] = n(d, d)
        sd[p + "mlp.gate_proj.weight"] = n(I, d)
        sd[p + "mlp.up_proj.weight"] = n(I, d)
        sd[p + "mlp.down_proj.weight"] = n(d, I)
        sd[p + "input_layernorm.weight"] = 1.0 + 0.1 * torch.randn(d, generator=g)
        sd[p + "post_attention_layernorm.weight"] = 1.0 + 0.1 * torch.randn(d, generator=g)
        if cfg.layer_scale_init_value > 0:
            sd[p + "lambda_1"] = cfg.layer_scale_init_value * torch.ones(d)
            sd[p + "lambda_2"] = cfg.layer_scale_init_value * torch.ones(d)
    sd["model.norm.weight"] = 1.0 + 0.1 * torch.randn(d, generator=g)
    if gated or cfg.stacked_feat_agg_method == "gated":
        sd["stacked_feat_agg.weight"] = torch.rand(cfg.stacked_feat, d, generator=g) * 2 / math.sqrt(d) - 1 / math.sqrt(d)
    if task_head:
        sd["score.weight"] = n(cfg.num_labels, d)
    else:
        sd["lm_head.weight"] = n(V, d)
        if cfg.next_n_token > 1:
            sd["n_token_proj.weight"] = n(d * cfg.next_n_token, d)
    return sd
