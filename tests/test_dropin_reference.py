"""The INTEGRATION.md edit applied in memory to the REAL reference (build container only: skipped where /root/reference
is absent, e.g. on the GPU box).  `src.models` gets this repo's classes, then the reference's own training modes are
imported and asked for their model class the way `TrainingPipeline._create_model` does (src/training/pipeline.py:150-165)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference checkout")

_SCRIPT = r'''
import sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests/golden")
import ref_shim
ref_shim._Finder.roots = tuple(r for r in ref_shim._Finder.roots if r != "accelerate")   # transformers parses its version
ref_shim.load_reference()
import graphgpt_b200 as gb
import src.models as ref_models
for name in ("GraphGPTPretrainBase", "GraphGPTTaskModel", "GraphGPTDoubleHeadsModel", "GraphGPTConfig", "convert_to_legacy_config"):
    setattr(ref_models, name, getattr(gb, name))            # == the two import lines of INTEGRATION.md section 1
from src.training import finetune_mode, pretrain_mode
from src.conf.model.model_configs import GraphGPTModelConfig
from src.utils import modules_utils
pm = pretrain_mode.PretrainMode.__new__(pretrain_mode.PretrainMode)
fm = finetune_mode.FinetuneMode.__new__(finetune_mode.FinetuneMode)
assert pm.dict_models["graphgpt"] is gb.GraphGPTPretrainBase and fm.dict_models["graphgpt"] is gb.GraphGPTTaskModel
mc = GraphGPTModelConfig()
mc.hidden_size, mc.num_attention_heads, mc.num_key_value_heads, mc.intermediate_size = 64, 1, 1, 256
mc.num_hidden_layers, mc.vocab_size, mc.hidden_act = 3, 300, "gelu"
mc.graph_input.stacked_feat, mc.graph_input.stack_method, mc.pt_head.next_n_token = 13, "short", 13
for mode, key in ((pm, pretrain_mode), (fm, finetune_mode)):
    cfg = key.convert_to_legacy_config(mc)                  # the name the mode module imported from src.models
    model = mode.dict_models["graphgpt"](cfg)               # pipeline.py:158
    model.gradient_checkpointing_enable()                   # pipeline.py:163
    model.config.use_cache = False                          # pipeline.py:164
    modules_utils.freeze_llama_layers(model, 1)             # finetune_mode.py:208-209
    frozen = [n for n, p in model.named_parameters() if not p.requires_grad]
    assert any(n.startswith("model.embed_tokens") for n in frozen) and any(n.startswith("model.layers.0.") for n in frozen)
    assert not any(n.startswith("model.layers.1.") for n in frozen)
    assert model.config.num_params if hasattr(model.config, "num_params") else True
print("dropin ok")
'''


def test_reference_training_modes_resolve_to_the_b200_classes(tmp_path):
    script = tmp_path / "dropin.py"
    script.write_text(_SCRIPT)
    p = subprocess.run([sys.executable, str(script), ROOT], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "dropin ok" in p.stdout, (p.stdout + p.stderr)[-3000:]


_GEN_SCRIPT = r'''
import sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests/golden")
import ref_shim
ref_shim._Finder.roots = tuple(r for r in ref_shim._Finder.roots if r != "accelerate")
ref_shim.load_reference()
import torch
from types import SimpleNamespace as NS
from src.utils import generation_utils as gu
from oracle import generation_oracle as go

bz, seq, F_, V, mask_id = 4, 12, 3, 50, 1
g = torch.Generator().manual_seed(0)
W = torch.randn(V, V, generator=g)                       # a deterministic stand-in "model": logits depend on the tokens
pos = torch.randn(seq * F_, V, generator=g)


def logits_fn(x):
    xf = x.reshape(x.shape[0], -1)
    ctx = torch.nn.functional.one_hot(xf, V).float().mean(1, keepdim=True) @ W      # [bz,1,V]: every entry sees the whole row
    return (W[xf] * 0.3 + ctx + pos[None]).reshape(-1, V)


class Model:
    def eval(self):
        return self

    def __call__(self, input_ids=None, attention_mask=None, labels=None, inputs_raw_embeds=None):
        return NS(head1_logits=logits_fn(input_ids))


ids = torch.randint(2, V, (bz, seq, F_), generator=g)
flat = ids.view(bz, -1)
for b in range(bz):                                      # the same number of masked entries in every sample: the
    flat[b, torch.randperm(seq * F_, generator=g)[:20]] = mask_id   # reference's re-masking quirk never triggers
for alg in ("origin", "maskgit_plus", "topk_margin", "entropy"):
    for steps in (3, 7, 50):
        cfg = NS(eps=1e-3, steps=steps, mask_token_id=mask_id, output_history=False, temperature=0.0, top_p=None, top_k=None,
                 alg=alg, alg_temp=None)
        torch.manual_seed(11)
        ref, _ = gu.sample_per_batch(Model(), cfg, input_ids=ids, attention_mask=None, inputs_raw_embeds=None)
        torch.manual_seed(11)
        mine, calls = go.sample_per_batch(logits_fn, ids, alg=alg, steps=steps, mask_token_id=mask_id)
        assert torch.equal(ref, mine), (alg, steps, int((ref != mine).sum()))
        assert int((mine == mask_id).sum()) <= int((logits_fn(mine.view(bz, seq, F_)).argmax(-1) == mask_id).sum()) + bz * seq * F_
print("generation loop ok")
'''


def test_generation_oracle_loop_matches_reference_sample_per_batch(tmp_path):
    """The whole decoding loop of the reference (generation_utils.sample_per_batch, all four algorithms, three step
    budgets) against oracle/generation_oracle.sample_per_batch on a deterministic stand-in model."""
    script = tmp_path / "gen.py"
    script.write_text(_GEN_SCRIPT)
    p = subprocess.run([sys.executable, str(script), ROOT], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "generation loop ok" in p.stdout, (p.stdout + p.stderr)[-3000:]


_PAD_SCRIPT = r'''
import sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests/golden")
import ref_shim
ref_shim._Finder.roots = tuple(r for r in ref_shim._Finder.roots if r != "accelerate")
ref_shim.load_reference()
import numpy as np, torch
from src.data import tokenizer as rtok
from graphgpt_b200 import synth

for F_vocab, seed in ((synth.PCQM_VOCAB, 3), (synth.TOY_VOCAB, 4)):
    b = synth.make_batch(6, 64, layout="unpacked", vocab=F_vocab, seed=seed)
    feats = []
    for n in range(6):
        L = int(b["attention_mask"][n].sum())
        feats.append({"input_ids": b["input_ids"][n, :L].tolist(), "labels": b["labels"][n, :L].tolist(),
                      "attention_mask": [1] * L, "position_ids": list(range(L))})

    class Fake:
        pad_token_id, label_pad_token_id, padding_side, eos_idx = 0, -100, "right", int(1e8)
        def set_eos_idx(self, x):
            pass
    fake = Fake()
    fake._pad_each_datapoint = lambda feat, pad_to: rtok.GSTTokenizer._pad_each_datapoint(fake, feat, pad_to)
    out = rtok.GSTTokenizer.pad(fake, feats, padding=True, max_length=1024, pad_to_multiple_of=8, return_tensors="pt")
    for k in ("input_ids", "labels", "attention_mask"):
        assert out[k].dtype == torch.int64 and torch.equal(out[k], torch.from_numpy(b[k])), (k, out[k].shape, b[k].shape)
    am = torch.from_numpy(b["attention_mask"]).bool()
    assert torch.equal(out["position_ids"][am], torch.from_numpy(b["position_ids"])[am])
print("collator pad ok")
'''


def test_synthetic_unpacked_batches_equal_the_reference_collator_padding(tmp_path):
    """graphgpt_b200.synth's right-padded batches (keys, dtypes, pad values, multiple-of-8 length) against the reference's
    own tokenizer.pad() (src/data/tokenizer.py:227-357) fed the same ragged samples."""
    script = tmp_path / "pad.py"
    script.write_text(_PAD_SCRIPT)
    p = subprocess.run([sys.executable, str(script), ROOT], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "collator pad ok" in p.stdout, (p.stdout + p.stderr)[-3000:]


_SURFACE_SCRIPT = r'''
import sys, inspect, dataclasses
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests/golden")
from ref_shim import load_reference
mp, mf, RefCfg = load_reference()
import graphgpt_b200 as gb
from src.models.graphgpt.modeling_common import DoubleHeadsModelOutput as RefOut
for name, ref in (("GraphGPTPretrainBase", mp.GraphGPTPretrainBase), ("GraphGPTTaskModel", mf.GraphGPTTaskModel),
                  ("GraphGPTDoubleHeadsModel", mf.GraphGPTDoubleHeadsModel)):
    assert list(inspect.signature(ref.forward).parameters) == list(inspect.signature(getattr(gb, name).forward).parameters), name
assert [f.name for f in dataclasses.fields(RefOut)] == [f.name for f in dataclasses.fields(gb.DoubleHeadsModelOutput)]
a, b = RefCfg().to_dict(), gb.GraphGPTConfig().to_dict()
diff = {k for k in set(a) | set(b) if a.get(k, "<missing>") != b.get(k, "<missing>")} - {"transformers_version", "use_aux"}
assert not diff, diff
print("surface ok")
'''


def test_forward_signatures_output_fields_and_config_defaults_equal_the_reference(tmp_path):
    script = tmp_path / "surface.py"
    script.write_text(_SURFACE_SCRIPT)
    p = subprocess.run([sys.executable, str(script), ROOT], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "surface ok" in p.stdout, (p.stdout + p.stderr)[-3000:]


_CKPT_SCRIPT = r'''
import sys, os, tempfile
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests/golden")
import ref_shim
ref_shim._Finder.roots = tuple(r for r in ref_shim._Finder.roots if r != "accelerate")
ref_shim.load_reference()
import torch
import graphgpt_b200 as gb
from graphgpt_b200.dp import write_model_pt
from src.utils import loader_utils
cfg = dict(vocab_size=300, hidden_size=64, intermediate_size=256, num_hidden_layers=2, num_attention_heads=1,
           num_key_value_heads=1, hidden_act="gelu", stacked_feat=5, next_n_token=5, stack_method="short", causal_attention=False)
torch.manual_seed(0)
pre = gb.GraphGPTPretrainBase(gb.GraphGPTConfig(**cfg))
d = tempfile.mkdtemp()
write_model_pt(pre, d)                                              # what GraphGPTEngine.save_checkpoint(d) leaves in d
torch.manual_seed(1)
ft = gb.GraphGPTTaskModel(gb.GraphGPTConfig(**cfg, num_labels=2, problem_type="single_label_classification"))
before = ft.score.weight.detach().clone()
ft = loader_utils.load_from_ckp_with_try(ft, d, skip_keys=True, strict=False)   # finetune_mode path (loader_utils.py:176-220)
sd_pre, sd_ft = pre.state_dict(), ft.state_dict()
for k in sd_ft:
    if k.startswith("model."):
        assert torch.equal(sd_ft[k], sd_pre[k]), k
assert torch.equal(ft.score.weight, before)                          # head keys are not in a pre-training checkpoint
print("checkpoint handoff ok")
'''


def test_engine_checkpoint_is_readable_by_the_reference_finetune_loader(tmp_path):
    """a checkpoint pre-trained with GraphGPTEngine must be loadable by the UNCHANGED fine-tuning pipeline:
    `load_from_ckp_with_try` (src/utils/loader_utils.py:176-220) opens <ckp>/model.pt first."""
    script = tmp_path / "ckpt.py"
    script.write_text(_CKPT_SCRIPT)
    p = subprocess.run([sys.executable, str(script), ROOT], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "checkpoint handoff ok" in p.stdout, (p.stdout + p.stderr)[-3000:]
