"""Import shim: `computational-graph_b200/` (the product package, named after the reference) is not a
valid Python identifier, so this package aliases it under `computational_graph_b200`."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "computational-graph_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
