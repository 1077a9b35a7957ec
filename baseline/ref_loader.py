"""Import shim for the UNMODIFIED reference package (`src/`), from /root/reference (build container) or from the
git-ignored copy baseline/_ref/ made by baseline/make_ref.py (GPU box).  Used only by bench.py's reference / cpu_baseline
legs, tests/ and tests/golden/ — never by the product package.

The reference targets transformers 4.53.3 and imports PyG / ogb / rdkit / deepspeed / hydra ... from src/utils/__init__.py;
none of those is on the model path, so missing third-party packages are replaced by permissive stub modules (SURVEY
Appendix A), `is_torch_fx_available` (removed in transformers 5) is restored, and the dropout backbone's decoder layer
(utils_graphgpt.py:168-173 returns a tuple, the 4.53 contract) is unwrapped for transformers 5's layer loop."""
import importlib.machinery
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = ("/root/reference", os.path.join(HERE, "_ref"))
STUB_ROOTS = ("torch_geometric", "ogb", "rdkit", "tensorboardX", "torcheval", "torchmetrics", "deepspeed", "timm", "hydra",
              "omegaconf", "torch_scatter", "torch_sparse", "accelerate", "common_io", "odps")


def reference_root():
    for c in CANDIDATES:
        if os.path.isdir(os.path.join(c, "src", "models")):
            return c
    return None


class _Meta(type):
    def __getattr__(cls, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Meta(n, (), {})

    def __call__(cls, *a, **k):
        return cls

    def __getitem__(cls, k):
        return cls

    def __or__(cls, o):
        return cls

    def __ror__(cls, o):
        return cls


class _Stub(types.ModuleType):
    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return sys.modules.get(f"{self.__name__}.{n}") or _Meta(n, (), {})


class _Finder:
    def __init__(self, roots):
        self.roots = roots

    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, m):
        pass


def load_reference():
    """Returns (root, modeling_pretrain, modeling_finetune, GraphGPTConfig) of the unmodified reference, or raises."""
    root = reference_root()
    if root is None:
        raise RuntimeError("the reference package is neither at /root/reference nor at baseline/_ref (run baseline/make_ref.py "
                           "in the build container)")
    sys.dont_write_bytecode = True
    if root not in sys.path:
        sys.path.insert(0, root)
    import transformers.utils.import_utils as iu

    if not hasattr(iu, "is_torch_fx_available"):
        iu.is_torch_fx_available = lambda: False
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        missing = tuple(r for r in STUB_ROOTS if importlib.util.find_spec(r) is None)
        sys.meta_path.insert(0, _Finder(missing))
    from src.models.graphgpt import modeling_finetune, modeling_pretrain, utils_graphgpt
    from src.models.graphgpt.configuration_graphgpt import GraphGPTConfig

    if not getattr(utils_graphgpt.LlamaDecoderLayer, "_ggpt_unwrapped", False):
        # transformers 5's LlamaModel loop does not unpack the 4.53-style tuple the dropout / LayerScale layer returns
        orig = utils_graphgpt.LlamaDecoderLayer.forward

        def fwd(self, hidden_states, *a, **k):
            k.pop("past_key_values", None)      # 4.53 name is past_key_value; unused without a cache
            out = orig(self, hidden_states, *a, **k)
            return out[0] if isinstance(out, tuple) else out

        utils_graphgpt.LlamaDecoderLayer.forward = fwd
        utils_graphgpt.LlamaDecoderLayer._ggpt_unwrapped = True
    return root, modeling_pretrain, modeling_finetune, GraphGPTConfig
