"""Import alias: the package directory is named ``tf-gnn-samples_b200`` (not an importable name),
so this shim package points its ``__path__`` there and runs that package's ``__init__``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tf-gnn-samples_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f, _real, _os
