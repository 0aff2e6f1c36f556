"""Import alias: the package directory is `graph-neural-additive-networks---gnan_b200/` (not a valid Python identifier);
`import gnan_b200` resolves to it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "graph-neural-additive-networks---gnan_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
