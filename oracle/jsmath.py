"""TEST INFRASTRUCTURE ONLY.  Math.log10 / Math.pow as V8 computes them (include/fa_jsmath.h) for Python."""
import ctypes

from . import build

_lib = ctypes.CDLL(build.ensure_built())
for _n in ("fao_js_log10", "fao_js_log"):
    getattr(_lib, _n).restype = ctypes.c_double
    getattr(_lib, _n).argtypes = [ctypes.c_double]
_lib.fao_js_pow.restype = ctypes.c_double
_lib.fao_js_pow.argtypes = [ctypes.c_double, ctypes.c_double]


def log10(x: float) -> float:
    return _lib.fao_js_log10(float(x))


def log(x: float) -> float:
    return _lib.fao_js_log(float(x))


def pow(x: float, y: float) -> float:  # noqa: A001
    return _lib.fao_js_pow(float(x), float(y))
