"""graph-gpt_b200 — B200-native GraphGPT transformer hot path (hand-written sm_100a CUDA behind a C ABI).

Import as `graphgpt_b200` (see graphgpt_b200/__init__.py).  Public surface mirrors the reference's
`src.models`: GraphGPTConfig, GraphGPTPretrainBase, GraphGPTTaskModel, DoubleHeadsModelOutput.
"""
__version__ = "0.1.0"


def __getattr__(name):  # lazy: keep `import graphgpt_b200` cheap and torch-free until a class is requested
    if name in ("GraphGPTConfig", "GraphGPTPretrainBase", "GraphGPTTaskModel", "DoubleHeadsModelOutput",
                "GraphGPTForMaskedLM", "GraphGPTForCausalLM", "convert_to_legacy_config", "GraphGPTDoubleHeadsModel"):
        from . import modeling
        return getattr(modeling, name)
    if name in ("GenerationConfig", "sample_per_batch", "sample_per_example", "sample_tokens"):
        from . import generation
        return getattr(generation, name)
    raise AttributeError(name)
