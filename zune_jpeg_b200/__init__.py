"""Import shim: the package sources live in ``zune-jpeg_b200/`` (the name the project layout asks for), which
is not a valid Python identifier; this shim makes them importable as ``zune_jpeg_b200``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "zune-jpeg_b200")
__path__.insert(0, _real)  # submodules resolve inside zune-jpeg_b200/

with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
