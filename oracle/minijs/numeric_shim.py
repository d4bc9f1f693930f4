"""TEST INFRASTRUCTURE ONLY.  The part of numeric@1.2.6 (inner module 5 of the reference bundle, dist/main.js:2@B38281) that
output level 12 uses -- `make_coeffs` @B34527 / polyfit @B33793 call numeric.transpose, dot, inv and uncmin -- as an object
that oracle/minijs can hand to the reference's formant module.

numeric.js builds its element-wise operations with `Function(...)` from string templates and contains regex literals, which
minijs does not implement.  So:
  * every HAND-WRITTEN function level 12 reaches is the reference's own code, cut out of the bundle at run time and executed:
    dim, dotVV, dotMV, dotVM, dotMMsmall, dotMMbig, _getCol, dot, diag, rep, identity, inv, transpose, tensor, gradient, uncmin,
    norm2 -- all the summation orders, the pivoting, the finite-difference steps, the BFGS update and the line search;
  * the GENERATED helpers are natives written from their templates (@B7780 mapreduce / mapreduce2, numeric.pointwise):
    add / sub / mul / div / neg / isFinite work element by element (no order to get wrong), `all` is a conjunction,
    norm2Squared sums xi*xi from the LAST element down (mapreduce2: `for(i=n-1;i!==-1;--i)`), clone copies.
Nothing of the reference is copied into the repo: the sources are read from /root/reference when the fixtures are made."""
from __future__ import annotations

import math

from .interp import Interp, JSArray, JSObject, JSThrow, JSTyped, Native, UNDEF, to_num, truthy
from .parser import matching_end

HAND_WRITTEN = ["dim", "dotVV", "dotMV", "dotVM", "dotMMsmall", "dotMMbig", "_getCol", "dot", "diag", "rep", "identity", "inv",
                "transpose", "tensor", "gradient", "uncmin", "norm2"]


def _cut(src: str, name: str) -> str:
    key = "numeric." + name + "=function"
    a = src.index(key)
    if src.find(key, a + 1) >= 0:
        raise RuntimeError("anchor not unique: " + key)
    start = a + len("numeric." + name + "=")
    brace = src.index("{", start)
    return src[start:matching_end(src, brace)]


def _items(v):
    return v.a if isinstance(v, (JSArray, JSTyped)) else None


def _map1(f, x):
    it = _items(x)
    if it is None:
        return f(to_num(x))
    return JSArray([_map1(f, e) for e in it])


def _map2(f, x, y):
    ix, iy = _items(x), _items(y)
    if ix is None and iy is None:
        return f(to_num(x), to_num(y))
    if ix is not None and iy is not None:
        return JSArray([_map2(f, a, b) for a, b in zip(ix, iy)])
    if ix is not None:
        return JSArray([_map2(f, a, y) for a in ix])
    return JSArray([_map2(f, x, b) for b in iy])


def _div(a, b):
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0:
            return math.nan
        return math.copysign(math.inf, a) * (math.copysign(1.0, b))


def _all(x):
    it = _items(x)
    if it is None:
        return truthy(x)
    return all(_all(e) for e in it)


def _norm2sq(x):
    it = _items(x)
    if it is None:
        v = to_num(x)
        return 0.0 + v * v
    if it and _items(it[0]) is not None:            # matrices: the recursive template, last row first
        acc = 0.0
        for e in reversed(it):
            acc += _norm2sq(e)
        return acc
    acc = 0.0
    for e in reversed(it):                           # mapreduce2: for(i=n-1;i!==-1;--i) accum += xi*xi
        v = to_num(e)
        acc += v * v
    return acc


def _clone(x):
    it = _items(x)
    if it is None:
        return x
    return JSArray([_clone(e) for e in it])


def install(it: Interp, bundle_src: str) -> JSObject:
    """Builds the `numeric` object inside interpreter `it` (also as the global `numeric`, which the cut-out functions name)."""
    num = JSObject()
    it.globals.vars["numeric"] = num
    nat = {
        "add": lambda t, a: _map2(lambda p, q: p + q, a[0], a[1]),
        "sub": lambda t, a: _map2(lambda p, q: p - q, a[0], a[1]),
        "mul": lambda t, a: _map2(lambda p, q: p * q, a[0], a[1]),
        "div": lambda t, a: _map2(_div, a[0], a[1]),
        "neg": lambda t, a: _map1(lambda p: -p, a[0]),
        "isFinite": lambda t, a: _map1(lambda p: not (p != p or p in (math.inf, -math.inf)), a[0]),
        "all": lambda t, a: _all(a[0]),
        "norm2Squared": lambda t, a: _norm2sq(a[0]),
        "clone": lambda t, a: _clone(a[0]),
    }
    for k, f in nat.items():
        num.props[k] = Native(f, k)
    num.props["epsilon"] = 2.220446049250313e-16
    for name in HAND_WRITTEN:
        num.props[name] = it.eval_expression("(" + _cut(bundle_src, name) + ")")
    return num
