"""Import shim: the product lives in the directory ``hessian-llm-vision_b200/`` (a name Python
cannot import directly).  This package points its search path there and re-exports it, so
``import hessian_llm_vision_b200 as hlv`` and ``from hessian_llm_vision_b200 import kernels`` work."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "hessian-llm-vision_b200")
if not _os.path.isdir(_real):  # pragma: no cover
    raise ImportError(f"expected the package directory at {_real}")
__path__.insert(0, _real)

with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
