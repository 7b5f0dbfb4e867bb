"""Import shim: the package directory is named `simple-hrnet_b200` (not a valid Python
identifier), so `import simple_hrnet_b200` resolves to it through this alias package."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "simple-hrnet_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
