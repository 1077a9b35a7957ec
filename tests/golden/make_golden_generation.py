"""tests/golden/aux/generation.pt: outputs of the REFERENCE's generation helpers (src/utils/generation_utils.py) on
seeded inputs, to pin oracle/generation_oracle.py.  Build container only (needs /root/reference).
Re-run:  python tests/golden/make_golden_generation.py"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from ref_shim import load_reference  # noqa: E402


def main():
    load_reference()
    from types import SimpleNamespace as NS

    from src.utils import generation_utils as gu
    g = torch.Generator().manual_seed(7)
    rec = {"filters": [], "sample_tokens": [], "unmask": []}
    # 1. top-p / top-k filters and the confidence variants of sample_tokens (deterministic: temperature 0 -> arg max)
    logits = torch.randn(6, 40, 97, generator=g) * 3
    for temperature, top_p, top_k in [(0.0, None, None), (0.7, 0.9, None), (1.3, None, 10), (0.5, 0.6, 5), (0.0, 0.3, 200)]:
        lg = logits / temperature if temperature > 0 else logits
        if top_p is not None and top_p < 1:
            lg = gu.top_p_logits(lg, top_p)
        if top_k is not None:
            lg = gu.top_k_logits(lg, top_k)
        rec["filters"].append(dict(temperature=temperature, top_p=top_p, top_k=top_k, probs=torch.softmax(lg, dim=-1)))
    for margin, ent in [(False, False), (True, False), (False, True)]:
        for top_p, top_k in [(None, None), (0.8, 12)]:
            conf, x0 = gu.sample_tokens(logits, temperature=0.0, top_p=top_p, top_k=top_k, margin_confidence=margin,
                                        neg_entropy=ent)
            rec["sample_tokens"].append(dict(top_p=top_p, top_k=top_k, margin=margin, neg_entropy=ent, conf=conf, x0=x0))
    rec["logits"] = logits
    # 2. _batch_unmask_without_for_loop, every algorithm, walked over all steps of a schedule
    B, P, V, mask_id = 5, 60, 33, 1
    x_init = torch.randint(2, V, (B, P), generator=g)
    for b in range(B):
        n_mask = [0, 3, 17, 40, 60][b]
        x_init[b, torch.randperm(P, generator=g)[:n_mask]] = mask_id
    for alg in ("origin", "maskgit_plus", "topk_margin", "entropy"):
        for steps in (4, 12):
            cfg = NS(mask_token_id=mask_id, temperature=0.0, top_p=None, top_k=None, alg=alg, alg_temp=None)
            timesteps = torch.linspace(1, 1e-3, steps + 1)
            x, i, trace = x_init.clone(), 0, []
            while i < steps and (x == mask_id).any():
                lg = torch.randn(B, P, V, generator=g) * 2
                seed = 1000 + len(trace)
                torch.manual_seed(seed)
                x_in, i_in = x.clone(), i
                x, i = gu._batch_unmask_without_for_loop(x.clone(), lg, timesteps, i, cfg)
                torch.manual_seed(seed)
                u = torch.rand(x.shape)            # the draw the "origin" branch made (temperature 0: its only one)
                trace.append(dict(x_in=x_in, i_in=i_in, logits=lg, u_transfer=u, x_out=x.clone(), i_out=i))
            rec["unmask"].append(dict(alg=alg, steps=steps, timesteps=timesteps, trace=trace))
            print(alg, steps, "calls", len(trace), "masked left", int((x == mask_id).sum()))
    os.makedirs(os.path.join(HERE, "aux"), exist_ok=True)
    path = os.path.join(HERE, "aux", "generation.pt")
    torch.save(rec, path)
    print(path, f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
