"""Drop-in model classes for the reference's `src.models` surface, executing on the B200 hot path.

  GraphGPTConfig            <- src/models/graphgpt/configuration_graphgpt.py:6-203
  DoubleHeadsModelOutput    <- src/models/graphgpt/modeling_common.py:54-99
  GraphGPTPretrainBase      <- src/models/graphgpt/modeling_pretrain.py:57-266   (SMTP + NTP; aliases
                               GraphGPTForMaskedLM / GraphGPTForCausalLM)
  GraphGPTTaskModel         <- src/models/graphgpt/modeling_finetune.py:64-326

The classes keep the reference's constructor (`cls(config)`), forward() keyword surface, returned fields and
state-dict key names/shapes so `TrainingPipeline` and existing checkpoints work unchanged.  The
submodules (`model.layers[i].self_attn.q_proj`, ...) are parameter containers only: all arithmetic runs through
engine.HotPath -> C ABI -> sm_100a kernels, and raises if the CUDA library is unavailable.
"""
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
from torch import nn
from transformers import LlamaConfig
from transformers.utils import ModelOutput

from . import precise
from .engine import PRETRAIN_HEAD, BackboneFn, FtHeadFn, FtIntraFn, HotPath, PretrainHeadFn

# GraphGPT-specific config fields and their defaults (same names as the reference so YAML / checkpoints map 1:1).
_GRAPHGPT_FIELDS = dict(
    pooling_method="last", causal_attention=True, rope_range=0,
    embed_pdrop=0.0, path_pdrop=0.0, mlp_pdrop=0.0, layer_scale_init_value=0.0,
    stacked_feat=1, stack_method=None, stacked_feat_agg_method="sum", pos_agg_method="sum", pos_bins=512,
    embed_dim=0, next_n_token=1, use_generative=True, use_discriminative=False, focal_gamma=0.0, smtp_inside=False,
    cls_token_id=None, mlp=None, dropout=0.0, loss_type=None, num_neg=None, use_aux=False,
    # 3-D position pre-training / denoising heads: carried for config round-trips, not executed by this package
    smtp_power=1.0, pt_problem_type="pos-smtp-line", smtp_3d_power=1.0, smtp_3d_noise_scale=0.2, coord_lvl_mask=True,
    pt_num_bins=1024, pt_num_bins_line=256, pt_num_bins_cube=32, apply_denoise=False, label_smoothing=0.0,
    pt_pos_agg_method="gated", use_pos_proj=False, loss_agg="token-lvl", pt_pos_range="p1p", pt_smtp_2d_rate=0.1,
    smtp_2d_replace_rate=0.0, sep_2d3d_inputs=True, global_2d_mask=False, pt_use_discriminative=False,
    noise_scale=0.35, denoise_wgt=1.0, denoise_schedule_pow=0.0, bi_causal=False, r_2d=4.0, r_3d=0.0, r_both=6.0,
    add_pos_type=True, inputs_transform="token-line", num_bins_line=256, num_bins_cube=32, dn_pos_range="1p",
    dn_use_pos_proj=False, smtp_3d=False, smtp_wgt=1.0, smtp_3d_scheduler_power=0.1, smtp_denoise=True,
    smtp_vocab=256, dn_smtp_2d_rate=0.0, smtp_2d_scheduler_power=0.0,
)


class GraphGPTConfig(LlamaConfig):
    """Flat config bag: LlamaConfig fields + the GraphGPT extras above (configuration_graphgpt.py:26-203)."""

    model_type = "graphgpt"

    def __init__(self, vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                 num_attention_heads=32, hidden_act="silu", max_position_embeddings=2048, initializer_range=0.02,
                 rms_norm_eps=1e-6, use_cache=True, pad_token_id=0, tie_word_embeddings=False, **kwargs):
        for k, default in _GRAPHGPT_FIELDS.items():
            setattr(self, k, kwargs.pop(k, default))
        self.mlp = [] if self.mlp is None else list(self.mlp)
        if self.pooling_method not in {"last", "sum", "mean"}:
            raise ValueError(f"pooling_method={self.pooling_method!r}")
        self.rope_3d = False
        kwargs.pop("rope_scaling", None)   # the reference hard-sets rope_scaling=None (configuration_graphgpt.py:114)
        super().__init__(vocab_size=vocab_size, hidden_size=hidden_size, intermediate_size=intermediate_size,
                         num_hidden_layers=num_hidden_layers, num_attention_heads=num_attention_heads,
                         hidden_act=hidden_act, max_position_embeddings=max_position_embeddings,
                         initializer_range=initializer_range, rms_norm_eps=rms_norm_eps, use_cache=use_cache,
                         pad_token_id=pad_token_id, tie_word_embeddings=tie_word_embeddings, **kwargs)

    @property
    def rope_theta_value(self):
        rp = getattr(self, "rope_parameters", None)
        if rp:
            return rp["rope_theta"]
        return getattr(self, "rope_theta", 10000.0)


def convert_to_legacy_config(model_config):
    """Flatten the reference's structured GraphGPTModelConfig dataclass into GraphGPTConfig
    (configuration_graphgpt.py:210-342).  Accepts any object with the same attribute tree."""
    mc = model_config
    top = ["vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
           "num_key_value_heads", "head_dim", "attention_bias", "mlp_bias", "hidden_act", "max_position_embeddings",
           "initializer_range", "rms_norm_eps", "tie_word_embeddings", "rope_theta", "use_cache", "pad_token_id",
           "bos_token_id", "eos_token_id", "cls_token_id", "causal_attention", "rope_range", "layer_scale_init_value"]
    kw = {k: getattr(mc, k) for k in top if hasattr(mc, k)}
    nested = {
        "dropout_settings": dict(embed_pdrop="embed_dropout", path_pdrop="path_dropout", mlp_pdrop="mlp_dropout",
                                 attention_dropout="attention_dropout"),
        "graph_input": dict(stacked_feat="stacked_feat", stack_method="stack_method",
                            stacked_feat_agg_method="stacked_feat_agg_method", embed_dim="embed_dim"),
        "geometric_input": dict(pos_agg_method="pos_agg_method", pos_bins="pos_bins"),
        "pt_head": dict(next_n_token="next_n_token", use_generative="use_generative",
                        use_discriminative="use_discriminative", focal_gamma="focal_gamma", smtp_inside="smtp_inside"),
        "ft_head": dict(pooling_method="pooling_method", mlp="mlp", dropout="dropout", loss_type="loss_type",
                        num_neg="num_neg", num_labels="num_labels", problem_type="problem_type"),
        # 3-D position pre-training / denoising heads: carried through so that configs round-trip (smtp_power also
        # drives the in-model SMTP masking of GraphGPTPretrainBase)
        "pos_pt_head": dict(smtp_power="smtp_power", pt_problem_type="problem_type", smtp_3d_power="smtp_3d_power",
                            smtp_3d_noise_scale="smtp_3d_noise_scale", coord_lvl_mask="coord_lvl_mask",
                            pt_num_bins="num_bins", pt_num_bins_line="num_bins_line", pt_num_bins_cube="num_bins_cube",
                            apply_denoise="apply_denoise", label_smoothing="label_smoothing",
                            pt_pos_agg_method="pos_agg_method", use_pos_proj="use_pos_proj", loss_agg="loss_agg",
                            pt_pos_range="pos_range", pt_smtp_2d_rate="smtp_2d_rate",
                            smtp_2d_replace_rate="smtp_2d_replace_rate", sep_2d3d_inputs="sep_2d3d_inputs",
                            global_2d_mask="global_2d_mask", pt_use_discriminative="use_discriminative"),
        "denoise_head": dict(noise_scale="noise_scale", denoise_wgt="denoise_wgt",
                             denoise_schedule_pow="denoise_schedule_pow", bi_causal="bi_causal", r_2d="r_2d", r_3d="r_3d",
                             r_both="r_both", add_pos_type="add_pos_type", inputs_transform="inputs_transform",
                             num_bins_line="num_bins_line", num_bins_cube="num_bins_cube", dn_pos_range="pos_range",
                             dn_use_pos_proj="use_pos_proj", smtp_3d="smtp_3d", smtp_wgt="smtp_wgt",
                             smtp_3d_scheduler_power="smtp_3d_scheduler_power", smtp_denoise="smtp_denoise",
                             smtp_vocab="smtp_vocab", dn_smtp_2d_rate="smtp_2d_rate",
                             smtp_2d_scheduler_power="smtp_2d_scheduler_power"),
    }
    for group, fields in nested.items():
        sub = getattr(mc, group, None)
        if sub is None:
            continue
        for dst, src in fields.items():
            if hasattr(sub, src):
                kw[dst] = getattr(sub, src)
    ft = getattr(mc, "ft_head", None)
    if ft is not None and hasattr(ft, "task_ratio"):
        kw["use_aux"] = ft.task_ratio < 1
    kw = {k: v for k, v in kw.items() if v is not None}
    return GraphGPTConfig(**kw)


@dataclass
class DoubleHeadsModelOutput(ModelOutput):
    """Return type of both model classes (field names as modeling_common.py:87-99)."""
    pretrain_loss: Optional[torch.FloatTensor] = None
    task_loss: Optional[torch.FloatTensor] = None
    pretrain_logits: torch.FloatTensor = None
    task_logits: torch.FloatTensor = None
    head1_loss: Optional[torch.FloatTensor] = None
    head2_loss: Optional[torch.FloatTensor] = None
    head1_logits: torch.FloatTensor = None
    head2_logits: torch.FloatTensor = None
    logits: torch.FloatTensor = None
    past_key_values: Optional[Tuple[Tuple[torch.FloatTensor]]] = None
    hidden_states: Optional[Tuple[torch.FloatTensor]] = None
    task_hidden_states: Optional[torch.FloatTensor] = None
    attentions: Optional[Tuple[torch.FloatTensor]] = None


# ------------------------------------------------------------------------------------------------
# parameter containers (names/shapes = the reference's state dict)
# ------------------------------------------------------------------------------------------------
class _Weight(nn.Module):
    """A named weight holder (RMSNorm scale): `.weight` of shape [dim], initialised to ones."""

    def __init__(self, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))


class _Attn(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.q_proj, self.k_proj, self.v_proj, self.o_proj = (nn.Linear(d, d, bias=False) for _ in range(4))


class _MLP(nn.Module):
    def __init__(self, d, inter):
        super().__init__()
        self.gate_proj = nn.Linear(d, inter, bias=False)
        self.up_proj = nn.Linear(d, inter, bias=False)
        self.down_proj = nn.Linear(inter, d, bias=False)


class _DecoderLayer(nn.Module):
    def __init__(self, cfg, drop_prob):
        super().__init__()
        d = cfg.hidden_size
        self.self_attn = _Attn(d)
        self.mlp = _MLP(d, cfg.intermediate_size)
        self.input_layernorm = _Weight(d)
        self.post_attention_layernorm = _Weight(d)
        self.drop_prob = float(drop_prob)
        if cfg.layer_scale_init_value > 0:   # utils_graphgpt.py:95-104
            self.lambda_1 = nn.Parameter(cfg.layer_scale_init_value * torch.ones(d))
            self.lambda_2 = nn.Parameter(cfg.layer_scale_init_value * torch.ones(d))


class _Backbone(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.embed_tokens = nn.Embedding(cfg.vocab_size, cfg.hidden_size, padding_idx=cfg.pad_token_id)
        L = cfg.num_hidden_layers
        # stochastic-depth decay rule, utils_graphgpt.py:184
        dpr = [cfg.path_pdrop * i / max(L - 1, 1) for i in range(L)] if cfg.path_pdrop > 0 else [0.0] * L
        self.layers = nn.ModuleList([_DecoderLayer(cfg, dpr[i]) for i in range(L)])
        self.norm = _Weight(cfg.hidden_size)


class _StackedFeatAgg(nn.Module):
    """modeling_common.py:105-142: `sum` has no parameters, `gated` has weight [F, d]."""

    def __init__(self, cfg):
        super().__init__()
        self.gated = cfg.stacked_feat_agg_method == "gated"
        if self.gated:
            w = torch.empty((cfg.stacked_feat, cfg.hidden_size))
            nn.init.kaiming_uniform_(w, a=5 ** 0.5)
            self.weight = nn.Parameter(w)


class _GraphGPTBase(nn.Module):
    """Shared plumbing: parameter tree, HF-style init, flat layout order, module surface used by the pipeline."""

    supports_gradient_checkpointing = True

    def __init__(self, config: GraphGPTConfig):
        super().__init__()
        self.config = config
        if getattr(config, "embed_dim", 0) > 0 and config.embed_dim % 8 != 0:
            raise NotImplementedError(f"embed_dim={config.embed_dim}: the embed_proj GEMM reads bf16 rows through TMA, which "
                                      "needs 16-byte row pitches (embed_dim % 8 == 0); zero-pad the raw features")
        if getattr(config, "use_discriminative", False):
            raise NotImplementedError("contrastive head (use_discriminative) is out of scope (SURVEY §2 row 16)")
        self.model = _Backbone(config)
        self._hot = None

    def _init_raw_embed(self, with_mask_token):
        """Raw-embedding input branch (config.embed_dim > 0): embed_layernorm [E], embed_proj [d,E] and, for
        pre-training, emb_mask_token [1,1,E]  (modeling_pretrain.py:69-84, modeling_finetune.py:76-86)."""
        cfg = self.config
        if cfg.embed_dim <= 0:
            return
        self.embed_layernorm = _Weight(cfg.embed_dim)
        if with_mask_token:
            self.emb_mask_token = nn.Parameter(torch.empty((1, 1, cfg.embed_dim)).normal_(mean=0.0, std=cfg.initializer_range))
        self.embed_proj = nn.Linear(cfg.embed_dim, cfg.hidden_size, bias=False)

    # ---- init like HF LlamaPreTrainedModel._init_weights: normal(0, initializer_range), padding row zero
    def _init_weights(self):
        std = self.config.initializer_range
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, mean=0.0, std=std)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Embedding):
                nn.init.normal_(m.weight, mean=0.0, std=std)
                if m.padding_idx is not None:
                    with torch.no_grad():
                        m.weight[m.padding_idx].zero_()

    def _flat_param_order(self):
        """Layout of the flat parameter buffer: embedding (+gate) | layer 0 .. L-1 (q,k,v,o,gate,up,down,norms,
        lambdas — q|k|v and gate|up adjacent so the fused GEMMs see one [3d,d] / [2I,d] weight) | final norm | head."""
        named = dict(self.named_parameters())
        order = ["model.embed_tokens.weight"]
        for k in ("stacked_feat_agg.weight", "embed_layernorm.weight", "emb_mask_token", "embed_proj.weight"):
            if k in named:
                order.append(k)
        for i in range(self.config.num_hidden_layers):
            p = f"model.layers.{i}."
            order += [p + "self_attn.q_proj.weight", p + "self_attn.k_proj.weight", p + "self_attn.v_proj.weight",
                      p + "self_attn.o_proj.weight", p + "mlp.gate_proj.weight", p + "mlp.up_proj.weight",
                      p + "mlp.down_proj.weight", p + "input_layernorm.weight", p + "post_attention_layernorm.weight"]
            if p + "lambda_1" in named:
                order += [p + "lambda_1", p + "lambda_2"]
        order.append("model.norm.weight")
        for k in ("n_token_proj.weight", "lm_head.weight"):
            if k in named:
                order.append(k)
        # fine-tuning score head (nn.Linear or the MLP of modules_utils.py:8-34): its few parameters live in the flat
        # buffers too, so the fused AdamW / clipping / gradient exchange cover every parameter of the model
        order += [k for k in named if k.startswith("score.") and k not in order]
        return [(k, named[k]) for k in order]

    @property
    def hot(self) -> HotPath:
        if self._hot is None:
            self._hot = HotPath(self, self.config)
        self._hot.flat.ensure()
        return self._hot

    # ---- surface used by the reference pipeline ---------------------------------------------------
    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def gradient_checkpointing_enable(self, *a, **k):
        """Accepted for API compatibility (pipeline.py:163).  Activations are kept resident instead: 12L/768d at
        64k tokens needs ~27 GB of the 180 GB HBM, so recomputation would only burn FLOPs."""
        self._gradient_checkpointing = True

    def gradient_checkpointing_disable(self):
        self._gradient_checkpointing = False

    def get_input_embeddings(self):
        return self.model.embed_tokens

    def num_parameters(self):
        return sum(p.numel() for p in self.parameters())

    def check_device_errors(self):
        """Blocks until the kernels launched so far have run and raises IndexError if any of them met a token id / label
        outside the vocabulary or a position outside the rotary table (the reference fails with IndexError / a device-side
        assert).  Without this call the same error surfaces at the next forward() or optimizer step."""
        if self._hot is not None:
            torch.cuda.current_stream().synchronize()
            self._hot.check_device_errors(block=True)

    # ---- shared forward pieces ----------------------------------------------------------------------
    def _prep_ids(self, input_ids):
        if input_ids is None:
            raise ValueError("input_ids is required (inputs_embeds is not supported, as in the reference: "
                             "modeling_helpers.py:95)")
        if input_ids.dim() == 3:
            N, S, F_ = input_ids.shape
            if F_ != self.config.stacked_feat:
                raise AssertionError(f"stacked_feat: {self.config.stacked_feat}\nx.shape: {tuple(input_ids.shape)}")
            ids2d = input_ids.reshape(N * S, F_)
            in_ = input_ids[:, :, 0]
        else:
            N, S = input_ids.shape
            ids2d = input_ids.reshape(N * S, 1)
            in_ = input_ids
        return ids2d.contiguous(), in_, N, S

    def _droppath_scales(self, N, S, device):
        """Per-layer (attention-branch, MLP-branch) per-row scales implementing BeitDropPath, training only: each call of
        the reference's drop_path draws torch.rand((N,1,1)) on the device and keeps sample n iff floor(keep + u_n) == 1,
        scaled by 1/keep (utils_graphgpt.py:156,166).  The draws are made here in the reference's order (layer by layer,
        attention branch first), so a seeded run with attention_dropout == 0 drops the same samples as the reference."""
        if not self.training or self.config.path_pdrop <= 0:
            return None
        out = []
        for layer in self.model.layers:
            p = layer.drop_prob
            if p <= 0:
                out.append((None, None))
                continue
            keep = 1.0 - p
            pair = []
            for _ in range(2):
                m = torch.floor(keep + torch.rand((N, 1, 1), dtype=torch.float32, device=device)).view(N) / keep
                pair.append(m.repeat_interleave(S).contiguous())
            out.append(tuple(pair))
        return out

    def _raw_embed_inputs(self, inputs_raw_embeds, N, S, labels, fchk):
        """(raw f32 [T,E], labels int64 [T,F] | None, fchk) for HotPath.backbone_forward, or None when embed_dim == 0."""
        cfg = self.config
        if cfg.embed_dim <= 0:
            return None
        if inputs_raw_embeds is None:
            raise ValueError("config.embed_dim > 0: inputs_raw_embeds [N, S, embed_dim] is required")
        if inputs_raw_embeds.dim() == 4:
            # [N,S,S,E] edge embeddings, summed over the third axis after norm / dropout / projection
            # (modeling_helpers.py:135-136; fine-tuning models only — the pre-training mask-token swap is 3-D)
            if labels is not None or tuple(inputs_raw_embeds.shape) != (N, S, S, cfg.embed_dim):
                raise RuntimeError(f"inputs_raw_embeds shape {tuple(inputs_raw_embeds.shape)}: expected {(N, S, S, cfg.embed_dim)} "
                                   "(fine-tuning) or [N,S,E]")
            raw3 = inputs_raw_embeds.to(device=self.device, dtype=torch.float32).reshape(N * S, S, cfg.embed_dim).contiguous()
            return raw3, None, 0
        if tuple(inputs_raw_embeds.shape) != (N, S, cfg.embed_dim):
            raise RuntimeError(f"inputs_raw_embeds shape {tuple(inputs_raw_embeds.shape)} != {(N, S, cfg.embed_dim)}")
        raw2d = inputs_raw_embeds.to(device=self.device, dtype=torch.float32).reshape(N * S, cfg.embed_dim).contiguous()
        lab2d = None
        if labels is not None:
            lab2d = labels.to(self.device).reshape(N * S, -1).contiguous()
        return raw2d, lab2d, fchk

    def _run_backbone(self, input_ids, attention_mask, position_ids, raw=None):
        ids2d, in_, N, S = self._prep_ids(input_ids)
        hot = self.hot
        dev = self.device
        cfg = self.config
        if ids2d.device != dev:
            ids2d = ids2d.to(dev)
        if attention_mask is not None and attention_mask.device != dev:
            attention_mask = attention_mask.to(dev)
        if precise.enabled():
            # validation-only fp32 forward (split-bf16 operands through the same tcgen05 GEMM); see precise.py
            if self.training or torch.is_grad_enabled():
                raise RuntimeError("GGPT_PRECISE=1 is a validation-only inference mode: call model.eval() under torch.no_grad()")
            return precise.backbone_forward(hot, ids2d, N, S, attention_mask, position_ids, raw), in_, N, S
        dps = self._droppath_scales(N, S, dev)
        attn_drop = float(getattr(cfg, "attention_dropout", 0.0) or 0.0) if self.training else 0.0
        # the reference swaps in the dropout MLP only together with the dropout backbone (modeling_common.py:148-169,
        # utils_graphgpt.py:91-93): mlp_pdrop > 0 always selects it, so the condition reduces to mlp_pdrop itself
        extras = dict(raw=raw, embed_pdrop=float(cfg.embed_pdrop) if self.training else 0.0,
                      mlp_pdrop=float(cfg.mlp_pdrop) if self.training else 0.0)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            params = [p for _, p in hot.flat.order]
            hf = BackboneFn.apply(hot, ids2d, N, S, attention_mask, position_ids, dps, attn_drop, extras, *params)
        else:
            hf = hot.backbone_forward(ids2d, N, S, attention_mask, position_ids, None, dps, attn_drop, **extras)
        return hf, in_, N, S


class GraphGPTPretrainBase(_GraphGPTBase):
    """SMTP (bidirectional) and NTP (causal) pre-training model; the switch is config.causal_attention
    (modeling_pretrain.py:195-196)."""

    def __init__(self, config: GraphGPTConfig):
        super().__init__(config)
        if not config.use_generative:
            raise NotImplementedError("use_generative=False (contrastive-only pre-training) is out of scope")
        self.smtp_inside = bool(config.smtp_inside)     # modeling_pretrain.py:62-63
        self.smtp_power = float(config.smtp_power)
        self.stacked_feat_agg = _StackedFeatAgg(config)
        self._init_raw_embed(with_mask_token=True)
        d = config.hidden_size
        if config.next_n_token > 1:
            self.n_token_proj = nn.Linear(d, d * config.next_n_token, bias=False)
        else:
            self.n_token_proj = nn.Identity()
        self.lm_head = nn.Linear(d, config.vocab_size, bias=False)
        self._init_weights()

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                inputs_raw_embeds=None, labels=None, label_mask=None, sample_wgt=None, use_cache=None,
                output_attentions=None, output_hidden_states=None, return_dict=None, cache_position=None):
        if inputs_embeds is not None:
            raise AssertionError("inputs_embeds must be None (modeling_helpers.py:95)")
        cfg = self.config
        if self.smtp_inside:
            input_ids, labels = self._smtp_inside_inputs(input_ids)
        raw = None
        if cfg.embed_dim > 0:
            if labels is None:
                raise ValueError("config.embed_dim > 0 needs labels: the mask-token swap of the raw embeddings is driven by "
                                 "them (modeling_pretrain.py:134-137)")
            # rows keep their raw features when ANY of the examined labels is -100 (all F of them; only column 0 under
            # smtp_inside), modeling_pretrain.py:134-137
            raw = self._raw_embed_inputs(inputs_raw_embeds, input_ids.shape[0], input_ids.shape[1], labels,
                                         1 if self.smtp_inside else labels.reshape(labels.shape[0], labels.shape[1], -1).shape[-1])
        hi = None
        if labels is not None and not precise.enabled():
            # label compaction first: its two counts travel to the host while the backbone runs (ops.HeadIndex)
            from . import ops
            lab2d = labels.to(self.device).reshape(labels.shape[0] * labels.shape[1], -1).contiguous()
            self.hot                                         # flat buffers / library are validated before the first launch
            hi = ops.head_compact(lab2d)
        hf, in_, N, S = self._run_backbone(input_ids, attention_mask, position_ids, raw)
        hot = self._hot
        dev = hf.device
        loss = None
        if labels is None:
            hot.post_error_flags()
        if labels is None and precise.enabled():
            logits = precise.head_all_entries(hot, hf)
        elif labels is None:
            # inference / generation: every (n,s,f) entry gets logits (modeling_helpers.py:284-292)
            from . import ops
            fp = hot.flat
            if "n_token_proj.weight" in fp.offsets:
                proj = ops.gemm(hf, fp.wb("n_token_proj.weight")).view(-1, hot.d)
            else:
                proj = hf
            logits = ops.gemm(proj, fp.wb("lm_head.weight"), out_dtype=torch.float32)
        else:
            labels = labels.to(dev)
            lab2d = labels.reshape(N * S, -1).contiguous()
            wgt_fn, mode = None, "mean"
            if cfg.stack_method == "long" or sample_wgt is not None:
                mode = "dlm"
                if cfg.stack_method == "long":
                    # normalised per-sample weights, modeling_helpers.py:337-341
                    cnt = (lab2d != -100).view(N, -1).sum(-1).float()
                    per_sample = 1.0 / (cnt + 1e-7)
                else:
                    per_sample = sample_wgt.to(dev).float()

                def wgt_fn(hi, L, per_sample=per_sample, S=S):
                    return per_sample[(hi.ent_tok[:L].long() // S)].contiguous()
            if precise.enabled():
                loss, logits = precise.labelled_head(hot, hf, lab2d, N, S, wgt_fn, mode, PRETRAIN_HEAD)
            else:
                params = [p for _, p in hot.flat.order]
                loss, logits = PretrainHeadFn.apply(hot, hf, hi, N, S, wgt_fn, mode, PRETRAIN_HEAD, *params)
        return DoubleHeadsModelOutput(head1_loss=loss, head1_logits=logits, head2_loss=None, head2_logits=None,
                                      past_key_values=None, hidden_states=None, attentions=None)


    def _smtp_inside_inputs(self, input_ids):
        """modeling_pretrain.py:175-189 -> prepare_for_2d_smtp_inputs_labels (modeling_helpers.py:399-452) with
        smtp_2d_rate=1, replace_rate=0, global_2d_mask=False.  The uniform draws come from torch's generator in the
        reference's order and shapes (so a seeded run consumes the same random stream); the node-level look-up, the
        masking and the label construction run in one kernel (ggpt_smtp_mask_2d)."""
        from . import ops
        F_ = self.config.stacked_feat
        dev = self.device
        ids = input_ids.to(dev).contiguous()
        N, S, _ = ids.shape
        torch.rand((N, 1, 1), dtype=torch.float32, device=dev)            # sample_mask draw: `< smtp_2d_rate` = always true
        mr = torch.rand((N, 1, 1), dtype=torch.float32, device=dev)       # mr_per_sample
        u_node = torch.rand((N, S, F_), dtype=torch.float32, device=dev)  # mask_per_node draw
        torch.randn((N, S, F_), dtype=torch.float32, device=dev)          # _get_gaussian_rnd_tokens (unused: replace_rate = 0)
        torch.rand((N, S, F_), dtype=torch.float32, device=dev)           # replace_mask draw (unused: replace_rate = 0)
        return ops.smtp_mask_2d(ids, F_, mr.view(-1), u_node, self.smtp_power)


# The task description's names for the two pre-training modes; both are the same class.
GraphGPTForMaskedLM = GraphGPTPretrainBase
GraphGPTForCausalLM = GraphGPTPretrainBase


class _ScoreMLP(nn.Module):
    """src/utils/modules_utils.py:8-34 — optional MLP score head.  As in the reference, EVERY Linear (the first included) is
    preceded by the activation and the dropout, and all of them share the `bias` flag."""

    def __init__(self, d, num_labels, mlp, dropout, bias, hidden_act="gelu"):
        super().__init__()
        dims = [d] + list(mlp) + [num_labels]
        self.mlp_modules = nn.ModuleList([nn.Linear(dims[i], dims[i + 1], bias=bias) for i in range(len(dims) - 1)])
        from transformers.activations import ACT2FN
        self.act_fn = ACT2FN[hidden_act]
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        for m in self.mlp_modules:
            x = m(self.dropout(self.act_fn(x)))
        return x


class GraphGPTTaskModel(_GraphGPTBase):
    """Backbone + `score` head with last-valid-token pooling (modeling_finetune.py:64-326).  The score head acts on
    the pooled row only (the reference scores every position and then indexes one row — same values)."""

    def __init__(self, config: GraphGPTConfig):
        super().__init__(config)
        if config.stack_method in {"short", "long"}:
            self.stacked_feat_agg = _StackedFeatAgg(config)
        self._init_raw_embed(with_mask_token=False)
        bias = config.problem_type == "regression"
        self.num_labels = config.num_labels
        if len(config.mlp) > 0:
            self.score = _ScoreMLP(config.hidden_size, self.num_labels, config.mlp, config.dropout, bias, config.hidden_act)
        else:
            self.score = nn.Linear(config.hidden_size, self.num_labels, bias=bias)
        self.pos_weight = None
        self.pooling_method = config.pooling_method
        self._init_weights()

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                inputs_raw_embeds=None, task_labels=None, cls_idx=None, sample_wgt=None, use_cache=None,
                output_attentions=None, output_hidden_states=None, return_dict=None, **kwargs):
        cfg = self.config
        if input_ids is not None and input_ids.dim() == 3:
            input_ids = input_ids[:, :, : cfg.stacked_feat]
        raw = None
        if cfg.embed_dim > 0:
            raw = self._raw_embed_inputs(inputs_raw_embeds, input_ids.shape[0], input_ids.shape[1], None, 0)
        hf, in_, N, S = self._run_backbone(input_ids, attention_mask, position_ids, raw)
        self._hot.post_error_flags()
        dev = hf.device
        hidden = hf.view(N, S, -1)
        if cfg.pad_token_id is None:
            raise AssertionError("pad_token_id must be set")
        if self.pooling_method != "last":
            raise AssertionError(f"{self.pooling_method}!='last'")
        if cfg.loss_type == "token_ce_intra":
            return self._token_ce_intra_outputs(hf, hidden, in_.to(dev), task_labels, cls_idx, N, S)
        self._last_backbone = (hf, N, S)
        if precise.enabled() or cfg.loss_type == "token_ce":
            # validation-only fp32 path / token-level task: pooling by torch indexing ([N]-sized)
            seq_len = (in_.to(dev) != cfg.pad_token_id).sum(-1) - 1                  # modeling_helpers.py:78-86
            pooled_hidden = hidden[torch.arange(N, device=dev), seq_len]             # [N, d]
            if cfg.loss_type == "token_ce":
                return self._token_ce_outputs(hf, hidden, pooled_hidden, task_labels, N, S)
            pooled_logits = self.score(pooled_hidden.float())
            task_loss = None
            if task_labels is not None:
                task_loss = self._task_loss(task_labels.to(dev), pooled_logits, sample_wgt)
            return DoubleHeadsModelOutput(pretrain_loss=None, task_loss=task_loss, pretrain_logits=None,
                                          task_logits=pooled_logits.float(), past_key_values=None, hidden_states=hidden,
                                          task_hidden_states=pooled_hidden, attentions=None)
        task_loss, pooled_logits, pooled_hidden = self._fused_task_head(hf, in_.to(dev), task_labels, sample_wgt)
        return DoubleHeadsModelOutput(pretrain_loss=None, task_loss=task_loss, pretrain_logits=None,
                                      task_logits=pooled_logits, past_key_values=None, hidden_states=hidden,
                                      task_hidden_states=pooled_hidden, attentions=None)

    def _token_ce_intra_outputs(self, hf, hidden, in_, task_labels, cls_idx, N, S):
        """loss_type == "token_ce_intra" (modeling_finetune.py:137-165, 195-199): the normalised hidden states at positions
        cls_idx[n] .. cls_idx[n] + num_labels - 1 act as the sample's label embeddings; logits = 20 * cosine similarity of
        every position with them, CrossEntropyLoss over the labelled positions.  task_logits is [N,S,num_labels]."""
        if precise.enabled():
            raise NotImplementedError("token_ce_intra has no precise-mode variant")
        if cls_idx is None:
            raise ValueError("loss_type token_ce_intra needs cls_idx [N] (first class-token position of every sample)")
        cfg = self.config
        dev = hf.device
        if task_labels is not None:
            if cfg.problem_type is None:
                cfg.problem_type = "single_label_classification"
            if cfg.problem_type != "single_label_classification":
                raise NotImplementedError(f"token_ce_intra with problem_type={cfg.problem_type!r}")
        seq_len = (in_ != cfg.pad_token_id).sum(-1) - 1                              # modeling_helpers.py:78-86
        pooled_hidden = hidden[torch.arange(N, device=dev), seq_len]
        labels = None if task_labels is None else task_labels.to(dev).reshape(-1).to(torch.int64).contiguous()
        params = [p for _, p in self._hot.flat.order]
        loss, logits = FtIntraFn.apply(self._hot, hf, cls_idx.to(dev).reshape(-1).to(torch.int64).contiguous(), labels, N, S,
                                       self.num_labels, *params)
        return DoubleHeadsModelOutput(pretrain_loss=None, task_loss=loss if labels is not None else None,
                                      pretrain_logits=None, task_logits=logits.reshape(N, S, self.num_labels),
                                      past_key_values=None, hidden_states=hidden, task_hidden_states=pooled_hidden,
                                      attentions=None)

    def _score_layers(self):
        """[(weight name, bias name | None)] of the score head in execution order, and whether it is the MLP variant."""
        if isinstance(self.score, _ScoreMLP):
            return [(f"score.mlp_modules.{j}.weight", f"score.mlp_modules.{j}.bias" if m.bias is not None else None)
                    for j, m in enumerate(self.score.mlp_modules)], True
        return [("score.weight", "score.bias" if self.score.bias is not None else None)], False

    def _fused_task_head(self, hf, in_, task_labels, sample_wgt):
        """Pooling + score head + task loss as ONE kernel (ggpt_ft_head_fwd / _bwd); the loss selection mirrors
        calculate_task_loss (modeling_finetune.py:167-234)."""
        cfg = self.config
        hot = self._hot
        dev = hf.device
        layers, is_mlp = self._score_layers()
        if is_mlp and cfg.hidden_act != "gelu":
            raise NotImplementedError(f"score MLP with hidden_act={cfg.hidden_act!r} (GraphGPT uses gelu)")
        if len(layers) > 4:
            raise NotImplementedError("score MLP deeper than 4 Linear layers")
        mode, labels_i, labels_f, wgt = -1, None, None, None
        if task_labels is not None:
            labels = task_labels.to(dev)
            if cfg.problem_type is None:
                if self.num_labels == 1:
                    cfg.problem_type = "regression"
                elif self.num_labels > 1 and labels.dtype in (torch.long, torch.int):
                    cfg.problem_type = "single_label_classification"
                else:
                    cfg.problem_type = "multi_label_classification"
            if cfg.problem_type == "regression":
                mode = 3 if cfg.loss_type == "l1" else 2
                labels_f = labels.to(torch.float32).reshape(-1, self.num_labels).contiguous()
            elif cfg.problem_type == "single_label_classification":
                if cfg.loss_type == "auc":
                    raise NotImplementedError("auc loss is out of scope (SURVEY §2 row 16)")
                labels_i = labels.reshape(-1).to(torch.int64).contiguous()
                if sample_wgt is None:
                    mode = 0
                else:
                    mode, wgt = 1, sample_wgt.to(dev).to(torch.float32).reshape(-1).contiguous()
            else:
                if self.pos_weight is not None:
                    raise NotImplementedError("BCEWithLogitsLoss pos_weight")
                mode = 4
                labels_f = labels.to(torch.float32).reshape(-1, self.num_labels).contiguous()
        drop_p = float(cfg.dropout) if (self.training and is_mlp) else 0.0
        params = [p for _, p in hot.flat.order]
        loss, logits, pooled = FtHeadFn.apply(hot, hf, in_, layers, int(is_mlp), drop_p, mode, labels_i, labels_f, wgt, *params)
        return (loss if mode >= 0 else None), logits, pooled

    def _token_ce_outputs(self, hf, hidden, pooled_hidden, task_labels, N, S):
        """loss_type == "token_ce" (node-level tasks): `score` on every position, CrossEntropyLoss over the labelled
        positions (modeling_finetune.py:161-165, 195-199).  task_logits is the full [N,S,num_labels] grid as in the
        reference; the loss path runs the score GEMM on the labelled rows only (same values, ignore_index = -100)."""
        from . import ops
        hot = self._hot
        if "score.weight" not in hot.flat.offsets:
            raise NotImplementedError("token_ce needs the plain Linear score head without bias (config.mlp == [], "
                                      "problem_type != 'regression')")
        nl = self.num_labels
        with torch.no_grad():
            if precise.enabled():
                logits = precise.linear(hf, hot.flat.w("score.weight"))
            else:
                logits = ops.gemm(hf.detach(), hot.flat.wb("score.weight"), out_dtype=torch.float32)
        task_loss = None
        if task_labels is not None:
            if self.config.problem_type is None:
                self.config.problem_type = "single_label_classification"
            if self.config.problem_type != "single_label_classification":
                raise NotImplementedError(f"token_ce with problem_type={self.config.problem_type!r}")
            lab2d = task_labels.to(hf.device).reshape(N * S, 1).contiguous()
            if precise.enabled():
                task_loss, _ = precise.labelled_head(hot, hf, lab2d, N, S, None, "mean", ("score.weight", None, nl))
            else:
                params = [p for _, p in hot.flat.order]
                task_loss, _ = PretrainHeadFn.apply(hot, hf, lab2d, N, S, None, "mean", ("score.weight", None, nl), *params)
        return DoubleHeadsModelOutput(pretrain_loss=None, task_loss=task_loss, pretrain_logits=None,
                                      task_logits=logits.reshape(N, S, nl), past_key_values=None, hidden_states=hidden,
                                      task_hidden_states=pooled_hidden, attentions=None)

    def _task_loss(self, labels, pooled_logits, sample_wgt):
        """modeling_finetune.py:167-234 on the pooled [N, num_labels] logits (tiny; plain torch)."""
        cfg = self.config
        F = torch.nn.functional
        if cfg.problem_type is None:
            if self.num_labels == 1:
                cfg.problem_type = "regression"
            elif self.num_labels > 1 and labels.dtype in (torch.long, torch.int):
                cfg.problem_type = "single_label_classification"
            else:
                cfg.problem_type = "multi_label_classification"
        if cfg.problem_type == "regression":
            fn = F.l1_loss if cfg.loss_type == "l1" else F.mse_loss
            labels = labels.to(pooled_logits.dtype)
            if self.num_labels == 1:
                return fn(pooled_logits.squeeze(), labels.squeeze())
            return fn(pooled_logits, labels)
        if cfg.problem_type == "single_label_classification":
            if cfg.loss_type == "auc":
                raise NotImplementedError("auc loss is out of scope (SURVEY §2 row 16)")
            lg = pooled_logits.view(-1, self.num_labels).float()
            if sample_wgt is None:
                return F.cross_entropy(lg, labels.view(-1))
            l = F.cross_entropy(lg, labels.view(-1), reduction="none")
            w = sample_wgt.to(l.device).float().view(-1)
            return (l * w).sum() / w.sum()
        ok = labels == labels
        return F.binary_cross_entropy_with_logits(pooled_logits[ok], labels[ok], pos_weight=self.pos_weight)


class GraphGPTDoubleHeadsModel(GraphGPTTaskModel):
    """Task head + auxiliary LM head on one backbone pass (modeling_finetune.py:329-423).  The aux loss is
    CrossEntropyLoss(lm_head(hidden).float().view(-1, V), pretrain_labels.view(-1)): labelled rows are compacted on the
    device and lm_head runs on those rows only.  `pretrain_logits` ([N,S,V], bf16) is materialised in eval mode only —
    at C3 sizes it is a 5 GB tensor no training loop reads."""

    def __init__(self, config: GraphGPTConfig):
        super().__init__(config)
        self.vocab_size = config.vocab_size
        if config.use_aux:
            self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
            nn.init.normal_(self.lm_head.weight, mean=0.0, std=config.initializer_range)
            self._hot = None

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                inputs_raw_embeds=None, pretrain_labels=None, task_labels=None, cls_idx=None, sample_wgt=None,
                use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None):
        res = super().forward(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                              inputs_embeds=inputs_embeds, inputs_raw_embeds=inputs_raw_embeds, task_labels=task_labels,
                              cls_idx=cls_idx, sample_wgt=sample_wgt)
        pretrain_loss, pretrain_logits = None, None
        if self.config.use_aux:
            from . import ops
            hf, N, S = self._last_backbone
            hot = self._hot
            if pretrain_labels is not None:
                lab2d = pretrain_labels.to(hf.device).reshape(N * S, 1).contiguous()
                if precise.enabled():
                    pretrain_loss, _ = precise.labelled_head(hot, hf, lab2d, N, S, None, "mean", ("lm_head.weight", None, None))
                else:
                    params = [p for _, p in hot.flat.order]
                    pretrain_loss, _ = PretrainHeadFn.apply(hot, hf, lab2d, N, S, None, "mean",
                                                            ("lm_head.weight", None, None), *params)
            if not self.training:
                with torch.no_grad():
                    if precise.enabled():
                        pretrain_logits = precise.linear(hf, hot.flat.w("lm_head.weight")).reshape(N, S, -1)
                    else:
                        pretrain_logits = ops.gemm(hf.detach(), hot.flat.wb("lm_head.weight")).reshape(N, S, -1)
        self._last_backbone = None
        return DoubleHeadsModelOutput(pretrain_loss=pretrain_loss, task_loss=res.task_loss, pretrain_logits=pretrain_logits,
                                      task_logits=res.task_logits, past_key_values=None, hidden_states=None, attentions=None)
