"""ctypes binding of libggpt_b200.so — the only way the Python host code reaches the CUDA kernels.

There is deliberately no fallback: if the library is missing or a call fails, a RuntimeError is raised.
Function prototypes are parsed from include/ggpt_b200.h so the header stays the single source of truth.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libggpt_b200.so")
LIB_PATH = os.environ.get("GGPT_LIB_PATH", LIB_PATH)   # profiling aid: an alternative build of the same library
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ggpt_b200.h")

_CTYPES = {
    "int": ctypes.c_int,
    "long long": ctypes.c_longlong,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "uint64_t": ctypes.c_uint64,
    "unsigned long long": ctypes.c_uint64,
}


def parse_header(path=HEADER_PATH):
    """Return {name: (restype, [argtypes], [argnames])} for every function declared in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//.*", "", text)
    protos = {}
    for m in re.finditer(r"(const char\*|long long|int|void)\s+(ggpt_\w+)\s*\(([^)]*)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = {"int": ctypes.c_int, "void": None, "const char*": ctypes.c_char_p, "long long": ctypes.c_longlong}[ret]
        argtypes, argnames = [], []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                    argnames.append(a.split("*")[-1].strip())
                else:
                    parts = a.replace("const ", "").split()
                    ty = " ".join(parts[:-1])
                    argtypes.append(_CTYPES[ty])
                    argnames.append(parts[-1])
        protos[name] = (restype, argtypes, argnames)
    return protos


# functions whose int return is a value, not a status code
_VALUE_FUNCS = {"ggpt_abi_version", "ggpt_attn_mask_words", "ggpt_attn_max_tiles", "ggpt_euler_paths"}
# kernels launched per entry point (default 1; 0 = host-only helper) — feeds bench.py's `gpu_launches`
_LAUNCHES = {"ggpt_last_error": 0, "ggpt_abi_version": 0, "ggpt_device_info": 0, "ggpt_attn_mask_words": 0,
             "ggpt_head_scratch_ints": 0, "ggpt_gemm_split_plan": 0, "ggpt_euler_paths": 0, "ggpt_attn_max_tiles": 0, "ggpt_attn_mask_build": 4, "ggpt_head_compact": 3, "ggpt_attn_fwd": 2, "ggpt_attn_bwd": 4, "ggpt_ft_head_fwd": 2, "ggpt_ft_intra_fwd": 3, "ggpt_ft_intra_bwd": 3}
_GEMM_FUNCS = {"ggpt_gemm_bf16", "ggpt_gemm_bf16_resid", "ggpt_gemm_bf16_geglu", "ggpt_gemm_bf16_qkv_rope",
               "ggpt_gemm_bf16_dgeglu"}


class KernelTimer:
    """CUDA-event timing of C-ABI calls on the launching stream (used by bench.py inside the timed region).
    Collects, per entry point (GEMMs are keyed by operand majors too): calls, device ms, FLOPs."""

    def __init__(self, only=None):
        self.events = []   # (key, start, end, flops)
        self.only = only   # optional set of entry-point names to time (others run untimed: every event pair costs the
                           # stream a few microseconds, which adds up over ~250 calls per training step)

    def summary(self):
        import torch
        torch.cuda.synchronize()
        out = {}
        for key, a, b, fl in self.events:
            d = out.setdefault(key, {"calls": 0, "ms": 0.0, "flops": 0.0})
            d["calls"] += 1
            d["ms"] += a.elapsed_time(b)
            d["flops"] += fl
        return out


class _Lib:
    def __init__(self):
        self._dll = None
        self._protos = None
        self.launch_count = 0
        self.timer = None
        self._cache = {}

    def load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python graph-gpt_b200/build.py` "
                "(there is no CPU / PyTorch fallback for the GraphGPT hot path)"
            )
        dll = ctypes.CDLL(LIB_PATH)
        self._protos = parse_header()
        for name, (restype, argtypes, _) in self._protos.items():
            fn = getattr(dll, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        self._dll = dll
        return dll

    def last_error(self):
        return self.load().ggpt_last_error().decode(errors="replace")

    def call(self, name, *args):
        return self._bound(name)(*args)

    def _bound(self, name):
        """One Python callable per entry point, built on first use: resolves the function pointer, the launch count and
        the status-code convention once, so the per-launch host cost is a dict look-up and the ctypes call."""
        fn = self._cache.get(name)
        if fn is not None:
            return fn
        dll = self.load()
        cfn = getattr(dll, name)
        n_launch = _LAUNCHES.get(name, 1)
        is_status = self._protos[name][0] is ctypes.c_int and name not in _VALUE_FUNCS
        is_gemm = name in _GEMM_FUNCS

        def run(*args):
            self.launch_count += n_launch
            timer = self.timer
            if timer is not None and n_launch > 0 and (timer.only is None or name in timer.only):
                import torch
                key, flops = name, 0.0
                if is_gemm:
                    M, N, K = args[-4], args[-3], args[-2]
                    flops = 2.0 * M * N * K
                    if name == "ggpt_gemm_bf16":
                        key = f"{name}[a_mn={args[2]},b_mn={args[5]}]"
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                rc = cfn(*args)
                b.record()
                timer.events.append((key, a, b, flops))
            else:
                rc = cfn(*args)
            if is_status:
                if rc != 0:
                    raise RuntimeError(f"{name} failed (rc={rc}): {self.last_error()}")
                return None
            return rc

        self._cache[name] = run
        return run

    def __getattr__(self, name):
        if name.startswith("ggpt_"):
            return self._bound(name)
        raise AttributeError(name)


GEMM_FUNCS = frozenset(_GEMM_FUNCS)
lib = _Lib()
