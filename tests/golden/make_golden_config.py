"""tests/golden/aux/legacy_config.json: the reference's convert_to_legacy_config (configuration_graphgpt.py:210-342) applied
to its own structured GraphGPTModelConfig with several nested fields moved off their defaults.  Build container only.
Re-run:  python tests/golden/make_golden_config.py"""
import dataclasses
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from ref_shim import load_reference  # noqa: E402


def main():
    load_reference()
    from src.conf.model.model_configs import GraphGPTModelConfig
    from src.models.graphgpt import configuration_graphgpt as cg
    mc = GraphGPTModelConfig()
    mc.hidden_size, mc.num_attention_heads, mc.num_hidden_layers, mc.intermediate_size = 768, 12, 12, 3072
    mc.vocab_size, mc.causal_attention, mc.layer_scale_init_value, mc.rope_range = 756, False, 1.0, 0
    mc.dropout_settings.path_dropout, mc.dropout_settings.attention_dropout = 0.15, 0.1
    mc.graph_input.stacked_feat, mc.graph_input.stack_method, mc.pt_head.next_n_token = 13, "short", 13
    mc.pt_head.smtp_inside, mc.pt_head.focal_gamma = True, 2.0
    mc.ft_head.task_ratio, mc.ft_head.num_labels, mc.ft_head.problem_type, mc.ft_head.mlp = 0.5, 7, "regression", [64]
    mc.pos_pt_head.smtp_power, mc.pos_pt_head.pos_range, mc.pos_pt_head.smtp_2d_rate = 0.37, "3p", 0.25
    mc.denoise_head.pos_range, mc.denoise_head.smtp_2d_rate, mc.denoise_head.noise_scale = "2p", 0.3, 0.5
    flat = cg.convert_to_legacy_config(mc).to_dict()
    flat.pop("transformers_version", None)
    os.makedirs(os.path.join(HERE, "aux"), exist_ok=True)
    with open(os.path.join(HERE, "aux", "legacy_config.json"), "w") as f:
        json.dump({"structured": dataclasses.asdict(mc), "flat": flat}, f, indent=1, sort_keys=True)
    print(len(flat), "flat fields")


if __name__ == "__main__":
    main()
