"""Importable alias for the `graph-gpt_b200/` package directory (a hyphen is not a legal module name).

`import graphgpt_b200` executes graph-gpt_b200/__init__.py under this module's name, so
`graphgpt_b200.ops`, `graphgpt_b200.modeling`, ... resolve to the files in `graph-gpt_b200/`.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "graph-gpt_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
