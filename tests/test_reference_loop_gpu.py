"""The drop-in boundary END TO END: the reference's OWN training-step functions — `batch_training` and
`ft_batch_training` (src/utils/training_utils.py:7-95, 98-205), imported unmodified from baseline/_ref (GPU box) or
/root/reference (build container) — drive this repo's model classes for three optimizer steps on a B200:

  * DeepSpeed branch (`train_stats.use_deepspeed`): `model(**batch)`, `model.backward(loss)`, `model.step()` with
    `model` = GraphGPTEngine(GraphGPTPretrainBase | GraphGPTTaskModel)   — the exact call sequence of :30-45 / :136-159;
  * DDP branch: `torch.autocast(float16)`, `GradScaler.scale(loss).backward()`, `unscale_`, `clip_grad_norm_`,
    `scaler.step(AdamW)` on the bare module (:46-86).
The first-step loss must equal the CPU oracle's (<= 1e-3), and the loss must fall over the three steps on the same batch.
Runs in a subprocess because the reference import installs stub modules for its missing third-party dependencies."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (os.path.isdir("/root/reference/src") or os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "src"))),
                                 reason="needs baseline/_ref (python baseline/make_ref.py in the build container)")]

_SCRIPT = r'''
import sys
ROOT = sys.argv[1]
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/baseline")
import ref_loader
ref_root = ref_loader.load_reference()[0]
import torch
from types import SimpleNamespace as NS
from src.utils import training_utils as tu                      # the reference's own step functions
import graphgpt_b200 as gb
from graphgpt_b200 import synth
from graphgpt_b200.dp import GraphGPTEngine
from oracle import graphgpt_oracle as oracle

dev = torch.device("cuda", 0)
base = dict(hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=2,
            head_dim=64, hidden_act="gelu", max_position_embeddings=1024, rms_norm_eps=1e-6, rope_theta=10000.0,
            pad_token_id=0, bos_token_id=20, eos_token_id=19, causal_attention=False, stack_method="short",
            stacked_feat_agg_method="sum", use_cache=False, attention_dropout=0.0)
train_cfg = NS(optimizer=NS(gradient_accumulation_steps=1, max_grad_norm=1.0), finetune=NS(use_aux=False))


def run(tag, step, stats, first_ref, n=3):
    losses = []
    for i in range(n):
        stats.i = i
        step()
        losses.append(float(stats.loss))
    rel = abs(losses[0] - first_ref) / abs(first_ref)
    print(f"{tag}: losses {[round(l, 5) for l in losses]} oracle first-step loss {first_ref:.5f} rel {rel:.2e}")
    assert rel <= 1e-3, (tag, rel)
    assert losses[-1] < losses[0], (tag, losses)


# ---- pre-training, DeepSpeed branch (training_utils.py:30-45) -------------------------------------------------
cfgd = dict(base, vocab_size=756, stacked_feat=13, next_n_token=13)
sd = oracle.init_state_dict(cfgd, seed=1)
b = synth.make_batch(4, 256, layout="packed", seed=3)
data = {k: torch.from_numpy(b[k]) for k in ("input_ids", "attention_mask", "position_ids", "labels")}
ref_loss = float(oracle.pretrain_forward(sd, cfgd, data["input_ids"], data["attention_mask"], data["labels"])["loss"])
model = gb.GraphGPTPretrainBase(gb.GraphGPTConfig(**cfgd))
model.load_state_dict(sd, strict=True)
model.gradient_checkpointing_enable()                            # pipeline.py:163
engine = GraphGPTEngine(model.to(dev).train(), lr=1e-3, max_grad_norm=1.0)
stats = NS(device=dev, has_embeds_input=False, use_deepspeed=True)
run("batch_training[deepspeed-branch]", lambda: tu.batch_training(data, engine, train_cfg, stats, NS()), stats, ref_loss)
assert stats.main_loss is stats.loss and stats.aux_loss is None and tuple(stats.inputs_shape) == (4, 256, 13)
assert engine.global_steps == 3

# ---- pre-training, DDP branch: autocast + GradScaler + clip + torch AdamW on the bare module (:46-86) -----------
model2 = gb.GraphGPTPretrainBase(gb.GraphGPTConfig(**cfgd))
model2.load_state_dict(sd, strict=True)
model2 = model2.to(dev).train()
opt = torch.optim.AdamW(model2.parameters(), lr=1e-3, betas=(0.9, 0.95), eps=1e-6, weight_decay=0.1)
opt_stats = NS(optimizer=opt, scaler=torch.amp.GradScaler("cuda"), lr_scheduler=torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0))
stats2 = NS(device=dev, has_embeds_input=False, use_deepspeed=False)
run("batch_training[ddp-branch]", lambda: tu.batch_training(data, model2, train_cfg, stats2, opt_stats), stats2, ref_loss)

# ---- fine-tuning, DeepSpeed branch (:136-159) --------------------------------------------------------------------
ft = dict(base, vocab_size=1200, stacked_feat=4, next_n_token=4, num_labels=2, problem_type="single_label_classification",
          pooling_method="last")
vocab = synth.VocabLayout(vocab_size=1200, scope=512, n_node_attr=2, n_edge_attr=1)
fb = synth.make_ft_batch(8, 128, vocab=vocab, seed=5)
fdata = {k: torch.from_numpy(fb[k]) for k in ("input_ids", "attention_mask", "position_ids", "labels", "edge_labels")}
fdata["idx"] = torch.arange(8)
sdf = oracle.init_state_dict(ft, seed=2, task_head=True)
ref_ft = float(oracle.task_forward(sdf, ft, fdata["input_ids"], fdata["attention_mask"], fdata["position_ids"],
                                   fdata["edge_labels"])["loss"])
fmodel = gb.GraphGPTTaskModel(gb.GraphGPTConfig(**ft))
missing, unexpected = fmodel.load_state_dict(sdf, strict=False)
assert not missing and not unexpected, (missing, unexpected)
fengine = GraphGPTEngine(fmodel.to(dev).train(), lr=1e-3, betas=(0.9, 0.99), eps=1e-10, weight_decay=0.0)
fstats = NS(device=dev, has_embeds_input=False, use_deepspeed=True, i=0)
fthead = NS(task_type="edge", problem_type="single_label_classification")
run("ft_batch_training[deepspeed-branch]", lambda: tu.ft_batch_training(fdata, fengine, fthead, train_cfg, fstats, NS()),
    fstats, ref_ft)
assert fstats.aux_loss is None and fstats.main_loss is not None
print("reference loop ok", ref_root)
'''


def test_reference_batch_training_functions_drive_the_b200_classes(tmp_path):
    script = tmp_path / "ref_loop.py"
    script.write_text(_SCRIPT)
    p = subprocess.run([sys.executable, str(script), ROOT], capture_output=True, text=True, timeout=900)
    out = p.stdout + p.stderr
    rep = os.path.join(ROOT, "gpurun_out", "reference_loop_report.txt")
    os.makedirs(os.path.dirname(rep), exist_ok=True)
    with open(rep, "w") as f:
        f.write("\n".join(l for l in p.stdout.splitlines() if "losses" in l or "reference loop ok" in l) + "\n")
    assert p.returncode == 0 and "reference loop ok" in p.stdout, out[-4000:]
