"""Un-masking (dLLM-style) generation on the B200 hot path — drop-in for the reference's src/utils/generation_utils.py.

Same names and argument meaning as the reference: `GenerationConfig` (src/conf/generation/generation_configs.py:27-57),
`sample_tokens` (:44-82), `sample_per_batch` (:85-135), `sample_per_example` (:316-430), `cal_gen_acc_batch` (:448-463).
Each decoding step is one forward pass of GraphGPTPretrainBase with labels=None (logits for every (n,s,f) entry)
followed by two kernels: ggpt_gen_sample (temperature / top-p / top-k / softmax / token choice / confidence, the logits
read exactly once) and ggpt_gen_unmask_{origin,topk} (which entries are revealed).  The only torch ops are the
reference's own random draws (torch.rand on the device, so a seeded run consumes the same stream) and the [B]-sized
schedule arithmetic that decides how many entries each sample reveals.

One deliberate difference: for the confidence algorithms the reference's scatter also RE-MASKS already revealed tokens
of samples that hold fewer masked entries than the batch-wide k (:215-227; it writes mask_token_id into the -inf tail
of torch.topk, whose tie order is implementation-defined).  Here such tokens are left untouched; reveals are identical
(tests/test_oracle_golden.py::test_generation_oracle_unmask_matches_reference).
"""
from dataclasses import dataclass
from typing import Optional

import torch

from . import ops


@dataclass
class GenerationConfig:
    alg: str = "origin"                 # origin | maskgit_plus | topk_margin | entropy
    alg_temp: Optional[float] = None
    steps: int = 512
    eps: float = 1e-3
    parallel_gen: bool = False
    temperature: float = 0.0
    top_p: Optional[float] = None
    top_k: Optional[int] = None
    max_length: int = 20
    max_new_tokens: Optional[int] = None
    num_return_sequences: int = 1
    return_dict_in_generate: bool = False
    output_history: bool = False
    mask_token_id: Optional[int] = None
    pad_token_id: Optional[int] = None
    bos_token_id: Optional[int] = None
    eos_token_id: Optional[int] = None


_ALGS = ("origin", "maskgit_plus", "topk_margin", "entropy")


def sample_tokens(logits, temperature=0.0, top_p=None, top_k=None, margin_confidence=False, neg_entropy=False, *,
                  u=None, want_probs=False):
    """logits f32 [..., V] on the device -> (confidence f32 [...], x0 int64 [...]).  temperature > 0 draws the token by
    inverse CDF from `u` (uniform [...]; drawn with torch.rand when None)."""
    lead = logits.shape[:-1]
    V = logits.shape[-1]
    lg = logits.reshape(-1, V)
    if lg.dtype != torch.float32:
        lg = lg.float()
    if temperature > 0 and u is None:
        u = torch.rand(lg.shape[0], device=lg.device, dtype=torch.float32)
    mode = 2 if neg_entropy else (1 if margin_confidence else 0)
    out = ops.gen_sample(lg, V, temperature=temperature, top_k=top_k, top_p=top_p,
                         u=None if u is None else u.reshape(-1), conf_mode=mode, want_probs=want_probs)
    conf, x0 = out[0].view(lead), out[1].view(lead)
    if want_probs:
        return conf, x0, out[2].view(*lead, V)
    return conf, x0


def _p_transfer(timesteps, i, steps):
    """1 - s/t in the schedule's dtype (fp32, as the reference's 0-d tensor arithmetic), 1.0 on the last step."""
    if i < steps - 1:
        return 1 - timesteps[i + 1] / timesteps[i]
    return 1.0


def _unmask_step(x, logits, timesteps, i, cfg):
    """_batch_unmask_without_for_loop (:138-228).  x int64 [B,P] (updated in place), logits f32 [B*P, V]."""
    steps = len(timesteps) - 1
    B, P = x.shape
    if cfg.alg not in _ALGS:
        raise ValueError(f"alg={cfg.alg!r}; expected one of {_ALGS}")
    if cfg.alg == "origin":
        p = _p_transfer(timesteps, i, steps)
        _, x0 = sample_tokens(logits, cfg.temperature, cfg.top_p, cfg.top_k)
        u = torch.rand(x.shape, device=x.device)                              # :165, the reference's draw
        ops.gen_unmask_origin(x, x0.view(B, P), u, float(p), cfg.mask_token_id)
        return x, i + 1
    num_masked = (x == cfg.mask_token_id).sum(dim=1)                          # [B]
    total = int(num_masked.sum().item())
    k, ntps = 0, None
    while k == 0 and total > 0 and i < steps:                                 # :177-186 skip steps that reveal nothing
        ntps = torch.floor(num_masked * _p_transfer(timesteps, i, steps)).int()
        k = int(ntps.max().item())
        i += 1
    if total == 0:
        return x, steps                                                       # nothing left to reveal: stop decoding
    if k == 0:
        return x, i
    conf, x0 = sample_tokens(logits, cfg.temperature, cfg.top_p, cfg.top_k, margin_confidence=(cfg.alg == "topk_margin"),
                             neg_entropy=(cfg.alg == "entropy"))
    gumbel_u = None
    if cfg.alg_temp is not None and cfg.alg_temp > 0:
        gumbel_u = torch.rand((B, P), device=x.device, dtype=torch.float32)   # :201-203 rand_like(confidence)
    ops.gen_unmask_topk(x, x0.view(B, P), conf.view(B, P), ntps, cfg.mask_token_id, gumbel_u=gumbel_u,
                        alg_temp=cfg.alg_temp or 0.0)
    return x, i


@torch.no_grad()
def sample_per_batch(model, cfg: GenerationConfig, *, input_ids, attention_mask, inputs_raw_embeds=None):
    """:85-135.  input_ids int64 [bz, seq, next_n] with cfg.mask_token_id at the entries to generate.
    Returns (x int64 [bz, seq*next_n], histories | None)."""
    model.eval()
    assert input_ids.dim() == 3, "expect [bz, seq, next_n]"
    bz, seq, next_n = input_ids.shape
    device = model.device
    x = input_ids.to(device).clone().view(bz, seq * next_n)
    if attention_mask is not None:
        attention_mask = attention_mask.to(device)
    m = x == cfg.mask_token_id
    steps = min(int(torch.max(m.sum(dim=-1).float()).item()), cfg.steps)
    timesteps = torch.linspace(1, cfg.eps, steps + 1, device=device)
    histories = [] if cfg.output_history else None
    i = 0
    while i < steps:
        logits = model(input_ids=x.view(bz, seq, next_n), attention_mask=attention_mask, labels=None,
                       inputs_raw_embeds=inputs_raw_embeds).head1_logits      # [bz*seq*next_n, V]
        x, i = _unmask_step(x, logits, timesteps, i, cfg)
        if histories is not None:
            histories.append(x.view(bz, seq, next_n).clone())
    return x, histories


@torch.no_grad()
def sample_per_example(model, cfg: GenerationConfig, *, input_ids, attention_mask, inputs_raw_embeds=None):
    """:316-430.  input_ids int64 [seq, next_n] (one example); every step runs (no skipping), the confidence algorithms
    reveal int(n_masked * (1 - s/t)) entries.  Returns (x int64 [1, seq*next_n], histories | None)."""
    model.eval()
    assert input_ids.dim() == 2
    device = model.device
    seq, next_n = input_ids.shape
    x = input_ids.to(device).clone().view(1, seq * next_n)
    if attention_mask is not None:
        attention_mask = attention_mask.to(device)
    steps = min(int((x == cfg.mask_token_id).sum().item()), cfg.steps)
    timesteps = torch.linspace(1, cfg.eps, steps + 1, device=device)
    histories = [] if cfg.output_history else None
    for i in range(steps):
        logits = model(input_ids=x.view(1, seq, next_n), attention_mask=attention_mask, labels=None,
                       inputs_raw_embeds=inputs_raw_embeds).head1_logits
        if cfg.alg == "origin":
            x, _ = _unmask_step(x, logits, timesteps, i, cfg)
        else:
            n_mask = (x == cfg.mask_token_id).sum()
            k = int(n_mask * (1 - timesteps[i + 1] / timesteps[i])) if i < steps - 1 else int(n_mask)
            if k > 0:
                conf, x0 = sample_tokens(logits, cfg.temperature, cfg.top_p, cfg.top_k,
                                         margin_confidence=(cfg.alg == "topk_margin"), neg_entropy=(cfg.alg == "entropy"))
                gumbel_u = None
                if cfg.alg_temp is not None and cfg.alg_temp > 0:
                    # the reference draws torch.multinomial(softmax(conf / alg_temp), k) (:421-427); perturb-and-top-k with
                    # Gumbel noise samples k entries without replacement from the same distribution
                    gumbel_u = torch.rand(x.shape, device=device, dtype=torch.float32)
                ops.gen_unmask_topk(x, x0.view(1, -1), conf.view(1, -1), torch.tensor([k], device=device, dtype=torch.int32),
                                    cfg.mask_token_id, gumbel_u=gumbel_u, alg_temp=cfg.alg_temp or 0.0)
        if histories is not None:
            histories.append(x.view(1, seq, next_n).clone())
    return x, histories


def cal_gen_acc_batch(gen_cfg: GenerationConfig, input_ids, labels, gen_res):
    """:448-463: per-sample accuracy over the generated (initially masked) entries."""
    bz, seq, next_n = input_ids.shape
    gen_res = gen_res.view(bz, seq, next_n)
    mask = input_ids == gen_cfg.mask_token_id
    correct = (gen_res == labels) & mask
    return correct.sum(dim=(1, 2)).float() / mask.sum(dim=(1, 2)).float()
