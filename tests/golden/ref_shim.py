"""Import shim that makes the REFERENCE's own model modules importable in the build container
(Python 3.12, torch 2.11, transformers 5.5.0; the reference pins transformers 4.53.3).

Only tests/golden/make_golden.py uses this; /root/reference does not exist on the GPU box, so nothing under
`-m gpu`, smoke() or bench.py imports it.  Missing third-party packages that src/utils/__init__.py pulls in
transitively (PyG, ogb, rdkit, ...) are replaced by permissive stubs; none of them is on the model path.
"""
import importlib.machinery
import sys
import types

REFERENCE_ROOT = "/root/reference"


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
    roots = ("torch_geometric", "ogb", "rdkit", "tensorboardX", "torcheval", "torchmetrics", "deepspeed", "timm",
             "hydra", "omegaconf", "torch_scatter", "torch_sparse", "accelerate", "common_io", "odps")

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
    """Returns (modeling_pretrain, modeling_finetune, GraphGPTConfig) of the unmodified reference."""
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import transformers.utils.import_utils as iu

    if not hasattr(iu, "is_torch_fx_available"):  # removed in transformers 5; utils_graphgpt.py:42,55
        iu.is_torch_fx_available = lambda: False
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, _Finder())
    from src.models.graphgpt import modeling_finetune, modeling_pretrain
    from src.models.graphgpt.configuration_graphgpt import GraphGPTConfig

    return modeling_pretrain, modeling_finetune, GraphGPTConfig
