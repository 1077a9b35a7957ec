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
