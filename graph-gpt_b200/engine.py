"""Host-side orchestration of the GraphGPT hot path: which kernel runs when, over which buffers.

`HotPath` owns three flat device buffers that alias every model parameter — fp32 master weights (the storage of
the nn.Parameters themselves), their bf16 compute copy (tensor-core operands) and the fp32 gradient buffer (the
storage of every `param.grad`) — plus the per-step activation stash.  `BackboneFn` / `PretrainHeadFn` are the
autograd entry points; their backward passes write parameter gradients straight into the flat gradient buffer
(returning None to autograd for parameter inputs) so that data-parallel all-reduce and the fused AdamW work on
contiguous segments.

Reference call sites reproduced (kernel-level citations are in include/ggpt_b200.h):
  LlamaModel.forward HF:375-425, LlamaDecoderLayer.forward HF:313-332 (+ utils_graphgpt.py:137-166 LayerScale /
  DropPath), prepare_for_stacked_feat_labels modeling_helpers.py:362-393, lm_head + CE modeling_pretrain.py:213-237.
"""
import torch

from . import ops

BF16, F32 = torch.bfloat16, torch.float32
# GGPT_DGEGLU_FUSED=0 runs GeGLU backward as the stand-alone multiply kernel after a plain down_proj dgrad GEMM (A/B aid)
_FUSE_DGEGLU = __import__("os").environ.get("GGPT_DGEGLU_FUSED", "1") != "0"


def _align(n, a=128):
    return (n + a - 1) // a * a


_M64 = (1 << 64) - 1


def mix_seed(base, k):
    """splitmix64 of (base + k): well-separated 63-bit seeds for the per-forward dropout streams (attention layer i,
    GeGLU output / MLP output of layer i, embedding, raw embedding), so that no two streams are index-shifted copies."""
    z = (base + 0x9E3779B97F4A7C15 * (k + 1)) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return (z ^ (z >> 31)) >> 1


# stream ids of mix_seed for one forward pass (attention dropout keeps its historical `seed + layer` numbering)
_SEED_ACT, _SEED_MLP, _SEED_EMBED, _SEED_RAW = 0x1000, 0x2000, 0x3000, 0x3001


class FlatParams:
    """Lays every parameter of `module` out in one flat fp32 buffer (each segment 128-element aligned so bf16
    views stay 16-byte aligned for TMA) and keeps `param.data` / `param.grad` aliased to it."""

    def __init__(self, module, order):
        self.module = module
        self.order = order            # list of (name, param) in layout order
        self.offsets = {}
        off = 0
        for name, p in order:
            self.offsets[name] = (off, tuple(p.shape))
            off += _align(p.numel())
        self.numel = off
        self.flat = None
        self.flat_bf16 = None
        self.flat_grad = None
        self._bf16_key = None

    def _views(self, buf, name):
        off, shape = self.offsets[name]
        n = 1
        for s in shape:
            n *= s
        return buf[off:off + n].view(shape)

    def ensure(self):
        """(Re)build the flat buffers when parameters moved (e.g. after .cuda()) and refresh the bf16 copy when any
        parameter changed in place (optimizer step, load_state_dict, manual edit)."""
        p0 = self.order[0][1]
        dev = p0.device
        if dev.type != "cuda":
            raise RuntimeError("graphgpt_b200: the model must live on a CUDA device (sm_100a); there is no CPU path")
        aliased = self.flat is not None and self.flat.device == dev
        if aliased:
            base = self.flat.data_ptr()
            for name, p in self.order:
                if p.dtype != F32 or p.data_ptr() != base + 4 * self.offsets[name][0]:
                    aliased = False
                    break
        if not aliased:
            for name, p in self.order:
                if p.dtype != F32:
                    raise RuntimeError(f"graphgpt_b200: parameter {name} has dtype {p.dtype}; fp32 master weights are "
                                       "required (bf16 compute copies are kept internally)")
            flat = torch.zeros((self.numel,), device=dev, dtype=F32)
            for name, p in self.order:
                v = self._views(flat, name)
                v.copy_(p.data)
                p.data = v
            self.flat = flat
            self.flat_bf16 = torch.empty((self.numel,), device=dev, dtype=BF16)
            self.flat_grad = torch.zeros((self.numel,), device=dev, dtype=F32)
            for name, p in self.order:
                p.grad = None
            self._bf16_key = None
        key = sum(p._version for _, p in self.order)
        if key != self._bf16_key:
            ops.cast_f32_bf16(self.flat, self.flat_bf16)
            self._bf16_key = key

    def invalidate_bf16(self):
        """Force a re-cast of the bf16 compute copy at the next forward.  ensure() notices in-place edits through
        `param._version`; edits made through `param.data` (or raw pointers) bypass the version counter — call this after
        them (GraphGPTEngine does after its parameter broadcast)."""
        self._bf16_key = None

    def mark_bf16_fresh(self):
        """Called by the fused optimizer, which updates fp32 and bf16 copies together through raw pointers."""
        self._bf16_key = sum(p._version for _, p in self.order)

    def w(self, name):
        return self._views(self.flat, name)

    def wb(self, name):
        return self._views(self.flat_bf16, name)

    def g(self, name):
        return self._views(self.flat_grad, name)

    def span(self, first, last):
        """[start, end) element range covering parameters first..last (inclusive) in layout order."""
        a = self.offsets[first][0]
        off, shape = self.offsets[last]
        n = 1
        for s in shape:
            n *= s
        return a, off + _align(n)

    def prepare_grads(self):
        """Before a backward pass: if no parameter has a gradient yet, zero the flat gradient buffer; then alias
        every `param.grad` to its flat segment.  Existing aliased gradients are accumulated into."""
        fresh = all(p.grad is None for _, p in self.order)
        if fresh:
            self.flat_grad.zero_()
        base = self.flat_grad.data_ptr()
        for name, p in self.order:
            if not p.requires_grad:
                continue
            if p.grad is None:
                gv = self.g(name)
                if not fresh:
                    gv.zero_()
                p.grad = gv
            elif p.grad.data_ptr() != base + 4 * self.offsets[name][0]:
                gv = self.g(name)          # a foreign gradient tensor: fold it into the flat buffer
                gv.copy_(p.grad)
                p.grad = gv


class HotPath:
    """Executes the transformer path of one GraphGPT model through the C-ABI kernels."""

    def __init__(self, module, config):
        self.cfg = config
        self.module = module
        d = config.hidden_size
        self.d = d
        self.H = config.num_attention_heads
        self.I = config.intermediate_size
        self.L = config.num_hidden_layers
        self.F = max(1, int(config.stacked_feat)) if getattr(config, "stack_method", None) in ("short", "long") else 1
        self.V = config.vocab_size
        hd = getattr(config, "head_dim", None) or d // self.H
        if hd != 64 or self.H * 64 != d:
            raise NotImplementedError(f"graphgpt_b200 kernels are built for head_dim 64 (got head_dim={hd}, heads={self.H}, "
                                      f"hidden={d}); GraphGPT always derives heads = hidden/64 (modules_utils.py:37-42)")
        kv = getattr(config, "num_key_value_heads", self.H)
        if kv != self.H:
            raise NotImplementedError("grouped-query attention is not used by GraphGPT (num_key_value_heads == heads)")
        if getattr(config, "hidden_act", "gelu") != "gelu":
            raise NotImplementedError(f"hidden_act={config.hidden_act!r}: GraphGPT uses exact GELU (configs/model/base.yaml:21)")
        if getattr(config, "attention_bias", False) or getattr(config, "mlp_bias", False):
            raise NotImplementedError("attention_bias / mlp_bias are not used by GraphGPT (configs/model/base.yaml:18-19)")
        self.layer_scale = getattr(config, "layer_scale_init_value", 0) > 0
        self.eps = config.rms_norm_eps
        self.flat = FlatParams(module, module._flat_param_order())
        self._rope_tab = None
        self._param_groups = None
        self.grad_ready_hook = None   # callable(first_name, last_name) fired as gradient segments complete
        # device-side input validation (the reference raises IndexError / a device-side assert for an id or label outside
        # the vocabulary): kernels set err[0] (input id), err[1] (label), err[2] (position id); the flags are copied to
        # pinned memory behind every forward and examined — without blocking — at the next forward / optimizer step, or
        # blocking through check_device_errors()
        self._err = None
        self._err_host = None
        self._err_event = None

    _ERR_TEXT = ("input_ids holds a token id outside [0, vocab_size)", "labels holds a target outside [0, vocab_size) other "
                 "than -100", "position_ids holds a position outside the rotary table")

    def err_flags(self, device):
        if self._err is None or self._err.device != device:
            self._err = torch.zeros((4,), device=device, dtype=torch.int32)
            self._err_host = torch.zeros((4,), dtype=torch.int32, pin_memory=True)
            self._err_event = None
        return self._err

    def post_error_flags(self):
        """Queue the D2H copy of the flags behind the kernels launched so far."""
        if self._err is not None:
            self._err_host.copy_(self._err, non_blocking=True)
            self._err_event = torch.cuda.Event()
            self._err_event.record()

    def check_device_errors(self, block=True):
        """Raise IndexError if a kernel of an earlier forward saw an out-of-range id / label / position.  block=False only
        looks when the copy has already landed (no host wait)."""
        ev = self._err_event
        if ev is None:
            return
        if not block and not ev.query():
            return
        ev.synchronize()
        self._err_event = None
        flags = self._err_host.tolist()
        if any(flags):
            self._err.zero_()
            msgs = [t for f, t in zip(flags, self._ERR_TEXT) if f]
            raise IndexError("graphgpt_b200: " + "; ".join(msgs) + " (vocab_size = %d)" % self.V)

    # ------------------------------------------------------------------ helpers
    def rope_tables(self, max_pos, device):
        if self._rope_tab is None or self._rope_tab[0].shape[0] < max_pos or self._rope_tab[0].device != device:
            n = max(max_pos, self.cfg.max_position_embeddings)
            theta = self._rope_theta()
            # HF:117-135 in fp32: inv_freq = theta^(-2i/64); freqs = pos * inv_freq
            inv_freq = 1.0 / (theta ** (torch.arange(0, 64, 2, dtype=torch.int64).float() / 64))
            freqs = torch.arange(n, dtype=torch.float32)[:, None] * inv_freq[None, :]
            self._rope_tab = (freqs.cos().to(device).contiguous(), freqs.sin().to(device).contiguous())
        return self._rope_tab

    def _rope_theta(self):
        theta = getattr(self.cfg, "rope_theta", None)      # transformers 4.x attribute; 5.x keeps it in rope_parameters
        if theta is None:
            theta = self.cfg.rope_parameters["rope_theta"]
        return theta

    def _rope_inputs(self, position_ids, N, S, device):
        """Returns (pos int32 [T], cos_tab, sin_tab)."""
        rope_range = getattr(self.cfg, "rope_range", 0)
        if position_ids is None:
            pos = torch.arange(S, device=device, dtype=torch.int32).repeat(N)      # HF:394-397
            cos, sin = self.rope_tables(S, device)
            return pos, cos, sin
        position_ids = position_ids.to(device)
        if rope_range > 0:
            # utils_graphgpt.py:574-581: fractional positions -> per-token cos/sin table, indexed by row
            mx = position_ids.max(dim=-1, keepdim=True)[0] + 1
            fpos = (position_ids.float() * rope_range / mx.float()).reshape(-1)
            inv_freq = 1.0 / (self._rope_theta() ** (torch.arange(0, 64, 2, dtype=torch.int64, device=device).float() / 64))
            freqs = fpos[:, None] * inv_freq[None, :]
            return torch.arange(N * S, device=device, dtype=torch.int32), freqs.cos().contiguous(), freqs.sin().contiguous()
        pos = position_ids.reshape(-1).to(torch.int32)
        max_pos = max(S, self.cfg.max_position_embeddings)
        cos, sin = self.rope_tables(max_pos, device)
        # HF computes cos/sin from the position VALUES and accepts anything; the kernels index a table of cos.shape[0] rows,
        # so out-of-range positions are flagged (IndexError at the next forward / step, check_device_errors()) and clamped
        n_rows = cos.shape[0]
        bad = ((pos < 0) | (pos >= n_rows)).any().to(torch.int32)
        flag = self.err_flags(device)[2:3]
        flag.copy_(torch.maximum(flag, bad.reshape(1)))
        return pos.clamp(0, n_rows - 1).contiguous(), cos, sin

    def _layer_names(self, i):
        p = f"model.layers.{i}."
        return p

    # ------------------------------------------------------------------ backbone
    def backbone_forward(self, ids2d, N, S, attention_mask, position_ids, stash, droppath_scales=None, attn_dropout=0.0,
                         raw=None, embed_pdrop=0.0, mlp_pdrop=0.0):
        """ids2d int64 [T,F].  Returns final-norm hidden states bf16 [T,d]; fills `stash` (a dict) when not None.
        raw = (raw_embeds f32 [T,E], labels int64 [T,F] | None, fchk) feeds the raw-embedding branch (embed_dim > 0);
        embed_pdrop / mlp_pdrop are the training-time element dropouts (0 in eval)."""
        fp = self.flat
        cfg = self.cfg
        dev = ids2d.device
        d, H = self.d, self.H
        gate = fp.w("stacked_feat_agg.weight") if "stacked_feat_agg.weight" in fp.offsets else None
        long_scale = getattr(cfg, "stack_method", None) == "long" and ids2d.shape[1] > 1
        self.check_device_errors(block=False)
        err = self.err_flags(dev)[0:1]
        # dropout (training only): one 62-bit seed per forward drawn from torch's CPU generator (so torch.manual_seed
        # controls it).  Attention (HF:217): layer i uses seed + i; the element dropouts use mix_seed streams.
        # Every mask is a pure function of (seed, index), regenerated by the backward pass — nothing is stored.
        any_drop = attn_dropout > 0 or embed_pdrop > 0 or mlp_pdrop > 0
        drop_seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if any_drop else 0
        x = ops.embed_fwd(ids2d, fp.w("model.embed_tokens.weight"), gate, long_scale, err, drop_p=embed_pdrop,
                          drop_seed=mix_seed(drop_seed, _SEED_EMBED))
        pos, cos, sin = self._rope_inputs(position_ids, N, S, dev)
        mask = ops.attn_mask_build(attention_mask, N, S, cfg.causal_attention, dev)
        keep = stash is not None
        if keep:
            stash.update(ids=ids2d, N=N, S=S, mask=mask, pos=pos, cos=cos, sin=sin, layers=[], err=err,
                         long_scale=long_scale, attn_dropout=attn_dropout, drop_seed=drop_seed,
                         embed_pdrop=embed_pdrop, mlp_pdrop=mlp_pdrop, raw=None)
        # Layer i:  h1 = norm(x) -> qkv(+RoPE) -> attention -> y1 = o_proj -> [x2 = x + y1 ; h2 = norm(x2)] (one fused
        # pass) -> gate|up + GeGLU -> y2 = down_proj -> [x3 = x2 + y2 ; h1' = next layer's input norm / final norm].
        # The residual adds live in the norm kernels: a GEMM epilogue doing fp32 read-modify-write of the residual
        # stream is latency-bound (measured 14 % tensor-pipe on o_proj), a streaming add+norm pass is not.
        w_in0 = fp.w("model.layers.0.input_layernorm.weight")
        if raw is not None:
            # raw-embedding branch (modeling_pretrain.py:119-150 / modeling_helpers.py:127-139): mask-token swap + RMSNorm
            # over E in one kernel, embed_proj on the tensor cores, and the sum with the token embeddings rides in the
            # add+norm pass that produces layer 0's input norm
            raw2d, rlabels, fchk = raw
            mask_tok = fp.w("emb_mask_token") if rlabels is not None else None
            if raw2d.dim() == 3:
                # [N,S,S,E] edge embeddings (fine-tuning): norm + dropout per (n,s,s') row, summed over s' BEFORE the
                # bias-free projection (modeling_helpers.py:127-139)
                hr = ops.raw_embed_norm_sum_fwd(raw2d, fp.w("embed_layernorm.weight"), self.eps, drop_p=embed_pdrop,
                                                drop_seed=mix_seed(drop_seed, _SEED_RAW))
                rstd_r = keep_r = None
            else:
                hr, rstd_r, keep_r = ops.raw_embed_norm_fwd(raw2d, fp.w("embed_layernorm.weight"), self.eps, labels=rlabels,
                                                            fchk=fchk, mask_tok=mask_tok, want_stash=keep)
                ops.dropout_(hr, embed_pdrop, mix_seed(drop_seed, _SEED_RAW))          # raw_embed_dropout
            yr = ops.gemm(hr, fp.wb("embed_proj.weight"))
            x, h1, rstd1 = ops.add_rmsnorm_fwd(x, yr, w_in0, self.eps, want_rstd=keep)
            if keep:
                stash["raw"] = dict(raw=raw2d, hr=hr, rstd=rstd_r, keep=keep_r, mask_tok=mask_tok)
        else:
            h1, rstd1 = ops.rmsnorm_fwd(x, w_in0, self.eps, want_rstd=keep)
        for i in range(self.L):
            p = f"model.layers.{i}."
            # DropPath draws an independent per-sample mask for each of the two residual branches (utils_graphgpt.py:156,166)
            rs1, rs2 = (None, None) if droppath_scales is None else droppath_scales[i]
            qkv = ops.gemm_qkv_rope(h1, self._wqkv(i), pos, cos, sin, 2 * d)
            a, lse, a_lo = ops.attn_fwd(qkv, mask, H, want_lse=keep, want_lo=True, dropout_p=attn_dropout, seed=drop_seed + i)
            y1 = ops.gemm(a, fp.wb(p + "self_attn.o_proj.weight"))
            lam1 = fp.w(p + "lambda_1") if self.layer_scale else None
            x2, h2, rstd2 = ops.add_rmsnorm_fwd(x, y1, fp.w(p + "post_attention_layernorm.weight"), self.eps,
                                                colscale=lam1, rowscale=rs1, want_rstd=keep)
            gu, act = ops.gemm_geglu(h2, self._wgu(i), want_gu=keep)
            ops.dropout_(act, mlp_pdrop, mix_seed(drop_seed, _SEED_ACT + i))          # mlp_act_dropout, utils_graphgpt.py:80
            y2 = ops.gemm(act, fp.wb(p + "mlp.down_proj.weight"))
            ops.dropout_(y2, mlp_pdrop, mix_seed(drop_seed, _SEED_MLP + i))           # mlp_dropout, utils_graphgpt.py:81
            lam2 = fp.w(p + "lambda_2") if self.layer_scale else None
            next_w = fp.w(f"model.layers.{i + 1}.input_layernorm.weight") if i + 1 < self.L else fp.w("model.norm.weight")
            x3, h_next, rstd_next = ops.add_rmsnorm_fwd(x2, y2, next_w, self.eps, colscale=lam2, rowscale=rs2,
                                                        want_rstd=keep)
            if keep:
                stash["layers"].append(dict(x=x, rstd1=rstd1, h1=h1, qkv=qkv, a=a, a_lo=a_lo, lse=lse, x2=x2, rstd2=rstd2, h2=h2,
                                            gu=gu, act=act, x3=x3 if self.layer_scale else None, rs1=rs1, rs2=rs2))
            x, h1, rstd1 = x3, h_next, rstd_next
        hf, rstdf = h1, rstd1
        if keep:
            stash.update(x_final=x, rstdf=rstdf)
        return hf

    def _first_trainable_layer(self):
        """0 when anything in front of the layer stack (embeddings, gate, raw-embedding branch) is trainable; else the
        index of the first layer holding a trainable parameter (L when there is none)."""
        if self._param_groups is None:
            names = [n for n, _ in self.flat.order]
            k0 = names.index("model.layers.0.self_attn.q_proj.weight")
            front = [p for _, p in self.flat.order[:k0]]
            layers = [[p for n, p in self.flat.order if n.startswith(f"model.layers.{i}.")] for i in range(self.L)]
            self._param_groups = (front, layers)
        front, layers = self._param_groups
        if any(p.requires_grad for p in front):
            return 0
        for i, ps in enumerate(layers):
            if any(p.requires_grad for p in ps):
                return i
        return self.L

    def _wqkv(self, i):
        """bf16 [3d, d] view: q_proj, k_proj, v_proj weights are adjacent in the flat layout."""
        fp = self.flat
        off, _ = fp.offsets[f"model.layers.{i}.self_attn.q_proj.weight"]
        return fp.flat_bf16[off:off + 3 * self.d * self.d].view(3 * self.d, self.d)

    def _wgu(self, i):
        fp = self.flat
        off, _ = fp.offsets[f"model.layers.{i}.mlp.gate_proj.weight"]
        return fp.flat_bf16[off:off + 2 * self.I * self.d].view(2 * self.I, self.d)

    def _gqkv(self, i):
        fp = self.flat
        off, _ = fp.offsets[f"model.layers.{i}.self_attn.q_proj.weight"]
        return fp.flat_grad[off:off + 3 * self.d * self.d].view(3 * self.d, self.d)

    def _ggu(self, i):
        fp = self.flat
        off, _ = fp.offsets[f"model.layers.{i}.mlp.gate_proj.weight"]
        return fp.flat_grad[off:off + 2 * self.I * self.d].view(2 * self.I, self.d)

    def backbone_backward(self, dhf, stash):
        """dhf: bf16 [T,d] gradient w.r.t. the final-norm hidden states.  Accumulates all parameter gradients."""
        fp = self.flat
        d, H = self.d, self.H
        wgrad = dict(a_mn_major=True, b_mn_major=True, out_dtype=F32, accumulate=True)
        dx, dxb = ops.rmsnorm_bwd(dhf, stash["x_final"], stash["rstdf"], fp.w("model.norm.weight"), None,
                                  fp.g("model.norm.weight"))
        if self.grad_ready_hook:
            self.grad_ready_hook("model.norm.weight", self.flat.order[-1][0])
        # frozen prefix (freeze_llama_layers, modules_utils.py:45-54: embeddings + the first k layers have requires_grad =
        # False): nothing below the first trainable layer needs a gradient, so backward stops there, as autograd would
        first_trainable = self._first_trainable_layer()
        for i in reversed(range(self.L)):
            if i < first_trainable:
                for j in range(i + 1):
                    stash["layers"][j] = None
                return
            p = f"model.layers.{i}."
            st = stash["layers"][i]
            # ---- MLP block:  x3 = x2 + rs * lam2 * (act @ Wd^T)
            dyb = dxb
            if self.layer_scale or st["rs2"] is not None:
                ls = self.layer_scale
                dyb = ops.layerscale_bwd(dx, st["x3"] if ls else None, st["x2"] if ls else None,
                                         fp.w(p + "lambda_2") if ls else None, st["rs2"], fp.g(p + "lambda_2") if ls else None)
            mlp_p, seed = stash["mlp_pdrop"], stash["drop_seed"]
            ops.dropout_(dyb, mlp_p, mix_seed(seed, _SEED_MLP + i))      # in place: dyb is this block's private bf16 copy
            # down_proj dgrad with the GeGLU backward in its epilogue: the forward stored gf = [u gelu'(g) | gelu(g)], so the
            # epilogue is two multiplies per element and dact never goes to HBM.  With mlp dropout the mask sits between
            # the GEMM and the multiply, so the two steps run separately.
            if mlp_p > 0 or not _FUSE_DGEGLU:
                dact = ops.gemm(dyb, fp.wb(p + "mlp.down_proj.weight"), b_mn_major=True)
                ops.dropout_(dact, mlp_p, mix_seed(seed, _SEED_ACT + i))
                dgu = ops.geglu_bwd(dact, st["gu"])
            else:
                dgu = ops.gemm_dgeglu(dyb, fp.wb(p + "mlp.down_proj.weight"), st["gu"])
            ops.gemm(dyb, st["act"], out=fp.g(p + "mlp.down_proj.weight"), **wgrad)
            ops.gemm(dgu, st["h2"], out=self._ggu(i), **wgrad)
            dh2 = ops.gemm(dgu, self._wgu(i), b_mn_major=True)
            dx2, dx2b = ops.rmsnorm_bwd(dh2, st["x2"], st["rstd2"], fp.w(p + "post_attention_layernorm.weight"), dx,
                                        fp.g(p + "post_attention_layernorm.weight"))
            # ---- attention block:  x2 = x + rs * lam1 * (a @ Wo^T)
            dyb = dx2b
            if self.layer_scale or st["rs1"] is not None:
                ls = self.layer_scale
                dyb = ops.layerscale_bwd(dx2, st["x2"] if ls else None, st["x"] if ls else None,
                                         fp.w(p + "lambda_1") if ls else None, st["rs1"], fp.g(p + "lambda_1") if ls else None)
            da = ops.gemm(dyb, fp.wb(p + "self_attn.o_proj.weight"), b_mn_major=True)
            ops.gemm(dyb, st["a"], out=fp.g(p + "self_attn.o_proj.weight"), **wgrad)
            dqkv = ops.attn_bwd(da, st["qkv"], st["a"], st["lse"], stash["mask"], H, stash["pos"], stash["cos"],
                                stash["sin"], out_lo=st["a_lo"], dropout_p=stash["attn_dropout"], seed=stash["drop_seed"] + i)
            ops.gemm(dqkv, st["h1"], out=self._gqkv(i), **wgrad)
            dh1 = ops.gemm(dqkv, self._wqkv(i), b_mn_major=True)
            dx, dxb = ops.rmsnorm_bwd(dh1, st["x"], st["rstd1"], fp.w(p + "input_layernorm.weight"), dx2,
                                      fp.g(p + "input_layernorm.weight"))
            stash["layers"][i] = None   # free activations as we go
            if self.grad_ready_hook:
                names = [n for n, _ in self.flat.order if n.startswith(p)]
                self.grad_ready_hook(names[0], names[-1])
        gate = fp.w("stacked_feat_agg.weight") if "stacked_feat_agg.weight" in fp.offsets else None
        emb_p = dict(self.flat.order)["model.embed_tokens.weight"]
        pad = self.cfg.pad_token_id if self.cfg.pad_token_id is not None else -1
        embed_p, seed = stash["embed_pdrop"], stash["drop_seed"]
        rw = stash["raw"]
        if rw is not None:
            # x0 = token-embedding sum + embed_proj(hr): dxb is the gradient of both summands
            ops.gemm(dxb, rw["hr"], out=fp.g("embed_proj.weight"), **wgrad)
            dhr = ops.gemm(dxb, fp.wb("embed_proj.weight"), b_mn_major=True)
            if rw["raw"].dim() == 3:
                ops.raw_embed_norm_sum_bwd(dhr, rw["raw"], fp.g("embed_layernorm.weight"), self.eps, drop_p=embed_p,
                                           drop_seed=mix_seed(seed, _SEED_RAW))
            else:
                ops.dropout_(dhr, embed_p, mix_seed(seed, _SEED_RAW))
                ops.raw_embed_norm_bwd(dhr, rw["raw"], rw["keep"], rw["mask_tok"], rw["rstd"], fp.w("embed_layernorm.weight"),
                                       fp.g("embed_layernorm.weight"),
                                       fp.g("emb_mask_token") if rw["mask_tok"] is not None else None)
        if (emb_p.requires_grad and gate is None and not stash["long_scale"] and self.V <= 4096
                and stash["ids"].shape[1] <= 32 and embed_p == 0):
            # small vocabulary: dE = C^T dX on the tensor cores (C = per-token id counts) instead of contended atomics
            cnt = ops.embed_count(stash["ids"], self.V, pad)
            ops.gemm(cnt, dxb, out=fp.g("model.embed_tokens.weight"), **wgrad)
        elif emb_p.requires_grad:
            ops.embed_bwd(stash["ids"], dx, fp.w("model.embed_tokens.weight") if gate is not None else None, gate,
                          fp.g("model.embed_tokens.weight"),
                          fp.g("stacked_feat_agg.weight") if gate is not None else None,
                          padding_idx=self.cfg.pad_token_id if self.cfg.pad_token_id is not None else -1,
                          long_scale=stash["long_scale"], drop_p=embed_p, drop_seed=mix_seed(seed, _SEED_EMBED))
        if self.grad_ready_hook:
            names = [n for n, _ in self.flat.order]
            self.grad_ready_hook(names[0], names[names.index("model.layers.0.self_attn.q_proj.weight") - 1])


# ------------------------------------------------------------------------------------------------
# autograd entry points
# ------------------------------------------------------------------------------------------------
class BackboneFn(torch.autograd.Function):
    """ids -> final-norm hidden states (bf16 [T,d]).  Parameter gradients are written into the flat gradient
    buffer by backward(); autograd only carries the activation gradient."""

    @staticmethod
    def forward(ctx, hot, ids2d, N, S, attention_mask, position_ids, droppath_scales, attn_dropout, extras, *params):
        stash = {}
        hf = hot.backbone_forward(ids2d, N, S, attention_mask, position_ids, stash, droppath_scales, attn_dropout,
                                  **extras)
        ctx.hot, ctx.stash = hot, stash
        return hf

    @staticmethod
    def backward(ctx, dhf):
        hot, stash = ctx.hot, ctx.stash
        if stash is None:
            raise RuntimeError("graphgpt_b200: backward called twice on the same forward (activations were freed)")
        hot.flat.prepare_grads()
        if dhf.dtype != BF16:
            dhf = dhf.to(BF16)
        hot.backbone_backward(dhf.contiguous(), stash)
        ctx.stash = None
        return (None,) * len(ctx.needs_input_grad)


PRETRAIN_HEAD = ("lm_head.weight", "n_token_proj.weight", None)


class PretrainHeadFn(torch.autograd.Function):
    """hidden (bf16 [T,d]) + labels -> (loss, logits[L,V]).  ref: modeling_pretrain.py:213-237.
    `head` = (vocabulary projection, per-feature projection | None, V | None) names the weights: the SMTP / NTP head by
    default; ("score.weight", None, num_labels) is the token-level fine-tuning loss (modeling_finetune.py:195-199) and
    ("lm_head.weight", None, None) the auxiliary LM head of GraphGPTDoubleHeadsModel (:402-411) — the same compaction of
    labelled rows, GEMM on those rows only and fp32 cross-entropy."""

    @staticmethod
    def forward(ctx, hot, hf, labels2d, N, S, ent_wgt_fn, loss_mode, head, *params):
        fp = hot.flat
        head_w, proj_w, V = head
        V = hot.V if V is None else V
        hi = labels2d if isinstance(labels2d, ops.HeadIndex) else ops.head_compact(labels2d)
        d, F_ = hot.d, hi.F
        M, L = hi.sync_counts()
        has_proj = proj_w is not None and proj_w in fp.offsets
        if M == 0:
            # no labelled entry: CrossEntropyLoss over zero targets is NaN in the reference too, and its backward still
            # runs (zero gradients) — under data parallelism every rank must keep issuing its gradient exchanges
            loss = torch.full((), float("nan"), device=hf.device, dtype=F32)
            ctx.hot, ctx.empty, ctx.T = hot, True, hf.shape[0]
            logits = torch.empty((0, V), device=hf.device, dtype=F32)
            ctx.mark_non_differentiable(logits)
            hot.post_error_flags()
            return loss, logits
        hsel = ops.gather_rows(hf, hi.sel_rows, M)
        if has_proj:
            proj = ops.gemm(hsel, fp.wb(proj_w))                                     # [M, F*d]
            hl = ops.gather_rows(proj.view(M * F_, d), hi.ent_src, L)
        else:
            hl = hsel                                                               # F == 1: entries == rows
        logits = ops.gemm(hl, fp.wb(head_w), out_dtype=F32)                        # [L, V] (ld padded to 8)
        wgt = ent_wgt_fn(hi, L) if ent_wgt_fn is not None else None
        # FocalLoss only on the unweighted branch (modeling_pretrain.py:221-236: the dLM loss ignores focal_gamma)
        focal = float(getattr(hot.cfg, "focal_gamma", 0.0) or 0.0) if wgt is None else 0.0
        row_lse, _, sums = ops.ce_fwd(logits, hi.ent_label, V, wgt, err_flag=hot.err_flags(hf.device)[1:2], focal_gamma=focal)
        hot.post_error_flags()
        if loss_mode == "mean":
            ls = ops.ce_finalize(sums, hi.counts.data_ptr() + 4, 0)
        else:                                                                       # dLM: sum / (N*S*F)
            ls = ops.ce_finalize(sums, 0, 2, float(N * S * hot.cfg.next_n_token))
        ctx.hot, ctx.hi, ctx.M, ctx.L, ctx.T = hot, hi, M, L, hf.shape[0]
        ctx.head_w, ctx.proj_w, ctx.V, ctx.F = head_w, proj_w, V, F_
        ctx.saved = (hsel, hl, logits, row_lse, wgt, ls)
        ctx.has_proj, ctx.empty, ctx.focal = has_proj, False, focal
        ctx.mark_non_differentiable(logits)
        return ls[0], logits

    @staticmethod
    def backward(ctx, gloss, _glogits):
        n_in = len(ctx.needs_input_grad)
        if ctx.empty:
            ctx.hot.flat.prepare_grads()
            dhf = torch.zeros((ctx.T, ctx.hot.d), device=gloss.device, dtype=BF16)
            return (None, dhf) + (None,) * (n_in - 2)
        hot, hi, M, L = ctx.hot, ctx.hi, ctx.M, ctx.L
        fp = hot.flat
        d, V = hot.d, ctx.V
        hsel, hl, logits, row_lse, wgt, ls = ctx.saved
        ctx.saved = None
        fp.prepare_grads()
        wgrad = dict(a_mn_major=True, b_mn_major=True, out_dtype=F32, accumulate=True)
        gout = gloss.reshape(1).to(F32).contiguous()
        dlog = ops.ce_bwd(logits, hi.ent_label, V, row_lse, ls.data_ptr() + 4, gout, wgt, focal_gamma=ctx.focal)  # bf16 [L, ld]
        dlv = dlog[:, :V]
        ops.gemm(dlv, hl, out=fp.g(ctx.head_w), **wgrad)
        dhl = ops.gemm(dlv, fp.wb(ctx.head_w), b_mn_major=True)                              # [L, d]
        if ctx.has_proj:
            F_ = ctx.F
            dproj = ops.expand_rows(dhl, hi.ent_src, L, M * F_).view(M, F_ * d)
            ops.gemm(dproj, hsel, out=fp.g(ctx.proj_w), **wgrad)
            dhsel = ops.gemm(dproj, fp.wb(ctx.proj_w), b_mn_major=True)                      # [M, d]
        else:
            dhsel = dhl
        dhf = ops.expand_rows(dhsel, hi.sel_rows, M, ctx.T)
        return (None, dhf) + (None,) * (n_in - 2)


class FtHeadFn(torch.autograd.Function):
    """hidden (bf16 [T,d]) -> (task loss | None, task_logits f32 [N,C], pooled hidden bf16 [N,d]): last-valid-row pooling,
    the `score` head and the task loss in one kernel per direction (csrc/ft_head.cu; ref: modeling_finetune.py:281-296,
    167-234, modeling_helpers.py:78-86, modules_utils.py:8-34).  `layers` = [(weight name, bias name | None), ...] in the
    flat parameter buffer; gradients of the head go straight into the flat gradient buffer."""

    @staticmethod
    def forward(ctx, hot, hf, in_ids, layers, act, drop_p, mode, labels_i, labels_f, sample_wgt, *params):
        fp = hot.flat
        weights = [fp.w(w) for w, _ in layers]
        biases = [fp.w(b) if b is not None else None for _, b in layers]
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if (act and drop_p > 0) else 0
        st = ops.ft_head_fwd(hf, in_ids, hot.cfg.pad_token_id, weights, biases, act=act, drop_p=drop_p, drop_seed=seed,
                             mode=mode, labels_i=labels_i, labels_f=labels_f, sample_wgt=sample_wgt,
                             err_flag=hot.err_flags(hf.device)[1:2])
        hot.post_error_flags()
        ctx.hot, ctx.st, ctx.layers, ctx.T = hot, st, layers, hf.shape[0]
        ctx.mark_non_differentiable(st.logits, st.pooled)
        loss = st.loss[0] if mode >= 0 else torch.zeros((), device=hf.device, dtype=F32)
        return loss, st.logits, st.pooled

    @staticmethod
    def backward(ctx, gloss, _gl, _gp):
        hot, st = ctx.hot, ctx.st
        fp = hot.flat
        fp.prepare_grads()
        named = dict(fp.order)
        dW = [fp.g(w) if named[w].requires_grad else None for w, _ in ctx.layers]
        dB = [fp.g(b) if (b is not None and named[b].requires_grad) else None for _, b in ctx.layers]
        dhf = torch.zeros((ctx.T, hot.d), device=gloss.device, dtype=BF16)
        ops.ft_head_bwd(st, gloss.reshape(1).to(F32).contiguous(), dW, dB, dhf)
        ctx.st = None
        return (None, dhf) + (None,) * (len(ctx.needs_input_grad) - 2)


class FtIntraFn(torch.autograd.Function):
    """loss_type "token_ce_intra" (modeling_finetune.py:137-165): hidden (bf16 [T,d]) -> (loss | 0, logits f32 [T,C]); the
    sample's own hidden states at its class-token positions are the label embeddings (csrc/ft_head.cu)."""

    @staticmethod
    def forward(ctx, hot, hf, cls_idx, labels, N, S, C, *params):
        st = ops.ft_intra_fwd(hf, cls_idx, labels, N, S, C, err_flag=hot.err_flags(hf.device)[1:2])
        hot.post_error_flags()
        ctx.hot, ctx.st = hot, st
        ctx.mark_non_differentiable(st.logits)
        loss = st.loss[0] if labels is not None else torch.zeros((), device=hf.device, dtype=F32)
        return loss, st.logits

    @staticmethod
    def backward(ctx, gloss, _gl):
        ctx.hot.flat.prepare_grads()
        dhf = ops.ft_intra_bwd(ctx.st, gloss.reshape(1).to(F32).contiguous())
        ctx.st = None
        return (None, dhf) + (None,) * (len(ctx.needs_input_grad) - 2)
