"""CPU oracle for the un-masking generation step — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see graphgpt_oracle.py for
the rules: only tests/, smoke() and bench.py's baseline legs may import oracle/).

Restates src/utils/generation_utils.py of the reference with every random draw passed in explicitly, so the CUDA
kernels can be compared value for value:
  top_p_logits :22-33, top_k_logits :36-41, sample_tokens :44-82, _batch_unmask_without_for_loop :138-228.

PARITY PIN: pinned.  tests/golden/aux/generation.pt (written by tests/golden/make_golden_generation.py) holds outputs
of the reference's own functions for seeded inputs; tests/test_oracle_golden.py::test_generation_oracle_* compare.
"""
import torch
import torch.nn.functional as F


def top_p_logits(logits, top_p):
    """:22-33.  Sort descending, drop tokens whose preceding cumulative probability exceeds top_p, keep the first."""
    sorted_logits, sorted_indices = torch.sort(logits, descending=True)
    cum = torch.cumsum(F.softmax(sorted_logits, dim=-1), dim=-1)
    remove = cum > top_p
    remove[..., 1:] = remove[..., :-1].clone()
    remove[..., 0] = 0
    mask = torch.zeros_like(logits, dtype=torch.bool).scatter_(-1, sorted_indices, remove)
    return logits.masked_fill(mask, torch.finfo(logits.dtype).min)


def top_k_logits(logits, top_k):
    """:36-41"""
    top_k = min(top_k, logits.size(-1))
    remove = logits < torch.topk(logits, top_k)[0][..., -1, None]
    return logits.masked_fill(remove, torch.finfo(logits.dtype).min)


def filtered_probs(logits, temperature=0.0, top_p=None, top_k=None):
    """The distribution sample_tokens draws from (:52-58)."""
    if temperature > 0:
        logits = logits / temperature
    if top_p is not None and top_p < 1:
        logits = top_p_logits(logits, top_p)
    if top_k is not None:
        logits = top_k_logits(logits, top_k)
    return torch.softmax(logits, dim=-1)


def sample_tokens(logits, temperature=0.0, top_p=None, top_k=None, margin_confidence=False, neg_entropy=False, u=None):
    """:44-82 with the categorical draw replaced by inverse-CDF sampling from the uniform draws `u` (same leading
    shape as logits without the vocabulary axis); u None or temperature == 0 -> arg max."""
    probs = filtered_probs(logits.float(), temperature, top_p, top_k)
    if temperature > 0 and u is not None:
        cdf = torch.cumsum(probs.double(), dim=-1)
        x0 = (cdf <= (u.double() * cdf[..., -1])[..., None]).sum(-1).clamp(max=probs.shape[-1] - 1)
        confidence = torch.gather(probs, -1, x0.unsqueeze(-1)).squeeze(-1)
    else:
        confidence, x0 = probs.max(dim=-1)
    if margin_confidence:
        sp, _ = torch.sort(probs, dim=-1, descending=True)
        confidence = sp[..., 0] - sp[..., 1]
    if neg_entropy:
        confidence = torch.sum(probs * torch.log(probs + 1e-10), dim=-1)
    return confidence, x0


def num_transfer(num_masked_per_sample, timesteps, i):
    """:173-186.  Returns (num_transfer_per_sample int32 [B], k, i_next): skips steps that would reveal nothing."""
    steps = len(timesteps) - 1
    k = 0
    ntps = torch.zeros_like(num_masked_per_sample, dtype=torch.int32)
    total = int(num_masked_per_sample.sum().item())
    while k == 0 and total > 0 and i < steps:
        t, s = timesteps[i], timesteps[i + 1]
        p_transfer = 1 - s / t if i < steps - 1 else 1.0
        ntps = torch.floor(num_masked_per_sample * p_transfer).int()
        k = int(ntps.max().item())
        i += 1
    return ntps, k, i


def batch_unmask(x, logits, timesteps, i, *, alg="origin", alg_temp=None, temperature=0.0, top_p=None, top_k=None,
                 mask_token_id=1, u_transfer=None, u_sample=None, u_gumbel=None):
    """_batch_unmask_without_for_loop (:138-228) on x int64 [B,P], logits f32 [B,P,V].  The reference's torch.rand /
    Categorical / rand_like draws arrive as u_transfer [B,P], u_sample [B,P], u_gumbel [B,P].  Returns (x, i)."""
    steps = len(timesteps) - 1
    x = x.clone()
    mask_index = x == mask_token_id
    if alg == "origin":
        t, s = timesteps[i], timesteps[i + 1]
        p_transfer = 1 - s / t if i < steps - 1 else 1.0
        _, x0 = sample_tokens(logits, temperature, top_p, top_k, u=u_sample)
        x = torch.where(mask_index & (u_transfer < p_transfer), x0, x)
        return x, i + 1
    ntps, k, i = num_transfer(mask_index.sum(dim=1), timesteps, i)
    confidence, x0 = sample_tokens(logits, temperature, top_p, top_k, margin_confidence=(alg == "topk_margin"),
                                   neg_entropy=(alg == "entropy"), u=u_sample)
    confidence = confidence.clone()
    confidence[~mask_index] = -torch.inf
    if alg_temp is not None and alg_temp > 0:
        confidence = confidence / alg_temp - torch.log(-torch.log(u_gumbel + 1e-9) + 1e-9)
    if k > 0:
        # per sample: the ntps[b] highest scores; stable sort = lowest position first on ties
        order = torch.sort(confidence, dim=1, descending=True, stable=True)[1][:, :k]
        upd = torch.gather(x0, 1, order)
        cut = torch.arange(k)[None, :] >= ntps[:, None]
        x.scatter_(1, order, torch.where(cut, torch.gather(x, 1, order), upd))
    return x, i


def sample_per_batch(logits_fn, input_ids, *, alg, steps, eps=1e-3, mask_token_id=1, alg_temp=None, temperature=0.0,
                     top_p=None, top_k=None, draws=None):
    """sample_per_batch (:85-135) over a callable `logits_fn(x [bz, seq, next_n]) -> [bz*seq*next_n, V]`.
    draws: optional callable(name, shape) -> uniform tensor, consulted for "transfer" / "gumbel" draws (None: torch.rand).
    Returns (x [bz, seq*next_n], number of forward passes)."""
    bz, seq, next_n = input_ids.shape
    x = input_ids.clone().view(bz, seq * next_n)
    m = x == mask_token_id
    n_steps = min(int(torch.max(m.sum(dim=-1).float()).item()), steps)
    timesteps = torch.linspace(1, eps, n_steps + 1)
    rnd = draws if draws is not None else (lambda name, shape: torch.rand(shape))
    i, calls = 0, 0
    while i < n_steps:
        logits = logits_fn(x.view(bz, seq, next_n)).view(bz, seq * next_n, -1)
        calls += 1
        if int((x == mask_token_id).sum()) == 0 and alg != "origin":
            break          # the reference would spin here; nothing is left to reveal
        x, i = batch_unmask(x, logits, timesteps, i, alg=alg, alg_temp=alg_temp, temperature=temperature, top_p=top_p,
                            top_k=top_k, mask_token_id=mask_token_id,
                            u_transfer=rnd("transfer", x.shape) if alg == "origin" else None,
                            u_gumbel=rnd("gumbel", x.shape) if (alg_temp is not None and alg_temp > 0) else None)
    return x, calls
