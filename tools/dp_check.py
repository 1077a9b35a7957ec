"""Data-parallel correctness check (run under torchrun with N >= 2 GPUs):
every rank takes one AdamW step on its own batch through GraphGPTEngine (overlapped NCCL all-reduce); rank 0 then
repeats the step single-process on a fresh copy of the model, feeding all ranks' batches with gradient accumulation
(sum / world), and the resulting parameters must agree."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from graphgpt_b200 import GraphGPTConfig, GraphGPTPretrainBase, ops, synth  # noqa: E402
from graphgpt_b200.dp import GraphGPTEngine  # noqa: E402

CFG = dict(vocab_size=756, hidden_size=256, intermediate_size=1024, num_hidden_layers=3, num_attention_heads=4,
           num_key_value_heads=4, head_dim=64, hidden_act="gelu", max_position_embeddings=1024, rms_norm_eps=1e-6,
           rope_theta=10000.0, pad_token_id=0, bos_token_id=20, eos_token_id=19, causal_attention=False, stacked_feat=13,
           stack_method="short", stacked_feat_agg_method="sum", next_n_token=13, use_cache=False)


def batch(rank, dev):
    b = synth.make_batch(4, 512, layout="packed", seed=100 + rank)
    return {k: torch.from_numpy(b[k]).to(dev) for k in ("input_ids", "attention_mask", "labels")}


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234 + rank)          # different init per rank on purpose: the engine must broadcast rank 0's
    model = GraphGPTPretrainBase(GraphGPTConfig(**CFG)).to(dev).train()
    if rank == 0:
        init = {k: v.detach().clone() for k, v in model.state_dict().items()}
    wire = torch.float32 if "--fp32-wire" in sys.argv else torch.bfloat16
    eng = GraphGPTEngine(model, lr=1e-3, max_grad_norm=1.0, grad_reduce_dtype=wire)
    out = eng(**batch(rank, dev))
    eng.backward(out.head1_loss)
    eng.step()
    grad_dp = eng.flat.flat_grad.clone()          # summed over ranks (the 1/world factor lives in the AdamW kernel)
    flat = eng.flat.flat.clone()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = torch.equal(flat, ref)
    ok = torch.tensor([int(same)], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        # single-process reference: accumulate every rank's batch, scale 1/world inside AdamW
        m2 = GraphGPTPretrainBase(GraphGPTConfig(**CFG))
        m2.load_state_dict(init)
        m2 = m2.to(dev).train()
        e2 = GraphGPTEngine.__new__(GraphGPTEngine)
        hot = m2.hot
        for r in range(world):
            m2(**batch(r, dev)).head1_loss.backward()
        fp = hot.flat
        g_err = ((grad_dp.double() - fp.flat_grad.double()).norm() / fp.flat_grad.double().norm()).item()
        g_tol = 1e-4 if wire == torch.float32 else 4e-3       # bf16 wire: one rounding per rank contribution + bf16 sums
        print(f"dp_check: all-reduced gradient ({'fp32' if wire == torch.float32 else 'bf16'} wire, {world} ranks) vs "
              f"single-process sum of the per-rank gradients: relF {g_err:.3e} (tolerance {g_tol:.0e})")
        assert g_err <= g_tol, g_err
        gn = torch.zeros((1,), device=dev, dtype=torch.float64)
        ops.sumsq(fp.flat_grad, gn)
        m = torch.zeros_like(fp.flat)
        v = torch.zeros_like(fp.flat)
        ops.adamw(fp.flat, fp.flat_bf16, fp.flat_grad, m, v, lr=1e-3, betas=(0.9, 0.95), eps=1e-6, weight_decay=0.1, step=1,
                  gnorm_sq=gn, max_norm=1.0, grad_scale=1.0 / world)
        torch.cuda.synchronize()
        d = (fp.flat - flat).abs()
        diff = d.max().item()
        frac_bad = (d > 2e-5).float().mean().item()
        print(f"dp_check: ranks identical={bool(ok.item())}; max |param_dp - param_single| = {diff:.3e}, "
              f"fraction of elements off by > 2e-5: {frac_bad:.2e} (one Adam step moves a parameter by ~lr = 1e-3; "
              f"near-cancelling gradient sums may flip sign between summation orders)")
        assert ok.item() == 1, "ranks diverged"
        assert frac_bad <= (1e-4 if wire == torch.float32 else 2e-2), "DP step differs from the single-process accumulated step"
        print("dp_check OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
