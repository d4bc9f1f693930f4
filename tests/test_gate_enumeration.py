"""Math.log10 / Math.pow and the noise gate: is the y -> v map of C() (@B28506) pinned to anything but
our own header?

What is established (tests/golden/make_gate_golden.py, exhaustive over all 4 294 967 295 reachable y -- y is always a positive
integer below 2^32):
  * OFF the integer points of the map (y not a multiple of 200 / 2000 / 20000 in its decade, not a perfect cube) the real value
    is >= 1/y >= 2^-32 relative away from the next integer: include/fa_jsmath.h gives floor(exact) for EVERY such y
    (0 mismatches), and so does any log10 / pow within thousands of ulps -- V8, glibc, a correctly rounded libm: the map is
    engine-independent there.
  * ON the 223 715 integer points the last bit decides between q and q - 1.  The header (restating fdlibm e_log10.c / e_pow.c,
    the algorithm V8's ieee754.cc ports) lands low on 108 350 of them; an ideal, correctly rounded libm lands low on 108 427 and
    disagrees with the header on 3 067 points (1.4 %): THOSE y are genuinely engine-sensitive and are listed in the fixture.
  * One piece of browser evidence ships with the reference: dist/nnmodel/*/model_meta.json stores min / max of feature [3] =
    Math.log10(y) over 74 k / 377 k syllables as full doubles.  Three of the four values equal the header's log10 bit for bit;
    the fourth (y = 121) is one ulp above it -- the correctly rounded value, which fdlibm's e_log10.c misses by 0.55 ulp.  So the
    engine that produced the shipped models was not (only) fdlibm; for y = 121 the difference cannot reach v (10^(t/3) = 4.95).
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest
from conftest import GOLDEN

from oracle import build, oracle

L = oracle._lib
L.fao_js_log10.restype = C.c_double
L.fao_js_log10.argtypes = [C.c_double]
L.fao_js_pow.restype = C.c_double
L.fao_js_pow.argtypes = [C.c_double, C.c_double]


def header_gate_v(y: int) -> int:
    t = L.fao_js_log10(float(y))
    if t > 7:
        return int(L.fao_js_pow(10.0, t - 3) / 20)
    if t > 6:
        return int(L.fao_js_pow(10.0, t - 3) / 2)
    if t > 4:
        return int(L.fao_js_pow(10.0, t - 2) / 2)
    if t > 2:
        return int(L.fao_js_pow(10.0, t / 3))
    if t > 1:
        return int(y / 10)
    return 1


def exact_v(y: int):
    """(exact-arithmetic v, on an integer point of the map?)"""
    if y > 10_000_000:
        return y // 20000, y % 20000 == 0
    if y > 1_000_000:
        return y // 2000, y % 2000 == 0
    if y > 10_000:
        return y // 200, y % 200 == 0
    if y > 100:
        r = round(y ** (1 / 3))
        while r ** 3 > y:
            r -= 1
        while (r + 1) ** 3 <= y:
            r += 1
        return r, r ** 3 == y
    if y > 10:
        return y // 10, False
    return 1, False


@pytest.fixture(scope="module")
def fixture():
    summary = json.load(open(os.path.join(GOLDEN, "gate_enumeration.json")))
    pts = np.load(os.path.join(GOLDEN, "gate_points.npz"))
    return summary, set(pts["low_points"].tolist()), set(pts["cr_differs"].tolist())


def test_summary_of_the_exhaustive_walk(fixture):
    summary, low, differs = fixture
    assert summary["range"] == [1, 1 << 32] and summary["checked"] == (1 << 32) - 1
    assert summary["off_boundary_mismatches"] == 0                   # engine-independent everywhere off the integer points
    assert summary["boundary_points_neither_q_nor_q_minus_1"] == 0
    assert summary["boundary_points"] == 223715 == 17 + 4950 + 4500 + 214248
    assert summary["boundary_points_low"] == len(low) == 108350
    assert summary["correctly_rounded_libm"]["differs_from_header"] == len(differs) == 3067


def test_enumerator_reproduces_the_fixture_on_a_sub_range(fixture):
    """The C enumerator itself, on [1, 2 * 10^7): same low points, no off-boundary mismatch (one second)."""
    _, low, _ = fixture
    exe = build.build_gate_enum()
    d = json.loads(subprocess.check_output([exe, "20000000"], env=dict(os.environ, OMP_NUM_THREADS="4")))
    assert d["off_boundary_mismatches"] == 0 and d["boundary_points"] == 9966
    assert sorted(d["low_points"]) == sorted(y for y in low if y < 20000000)


def test_header_agrees_with_the_fixture_on_sampled_points(fixture):
    _, low, _ = fixture
    rng = np.random.default_rng(7)
    # integer points of every branch: the side is the recorded one
    for y in list(rng.integers(52, 21474, 400) * 200000 // 1000 * 100):      # multiples of 20000 above 10^7
        y = int(y)
        if y <= 10_000_000 or y >= 1 << 32 or y % 20000:
            continue
        q, on = exact_v(y)
        assert on and header_gate_v(y) == (q - 1 if y in low else q)
    for y in [int(k) * 200 for k in rng.integers(51, 5000, 300)] + [int(k) * 2000 for k in rng.integers(501, 5000, 300)] + \
             [r ** 3 for r in range(5, 22)]:
        q, on = exact_v(y)
        assert on and header_gate_v(y) == (q - 1 if y in low else q), y
    # off the integer points: floor of the exact value, always
    for y in [int(v) for v in rng.integers(1, 1 << 32, 20000)] + list(range(1, 3000)):
        q, on = exact_v(y)
        if not on:
            assert header_gate_v(y) == q, y


def test_correctly_rounded_libm_differs_exactly_where_the_fixture_says(fixture):
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 240
    _, low, differs = fixture
    import sys
    sys.path.insert(0, GOLDEN)
    from make_gate_golden import boundary_points, cr_gate_v, exact_q
    pts = list(boundary_points())
    rng = np.random.default_rng(11)
    sample = [pts[i] for i in rng.choice(len(pts), 3000, replace=False)] + [p for p in pts if p[0] in differs][:200]
    for y, br in sample:
        q = exact_q(y, br)
        v_h = q - 1 if y in low else q
        assert (cr_gate_v(y, br, mp) != v_h) == (y in differs), y


def test_browser_evidence_in_the_shipped_models():
    """min / max of feature [3] = Math.log10(y) in dist/nnmodel/{1,4}/cats_emotion/model_meta.json (values copied from there:
    inputs["3"].min / .max), against the header and a correctly rounded log10."""
    shipped = {121: 2.0827853703164503, 295528896: 8.470599951409516, 44: 1.6434526764861874, 227542488: 8.357062502448434}
    same = {y: L.fao_js_log10(float(y)) == t for y, t in shipped.items()}
    assert same == {121: False, 295528896: True, 44: True, 227542488: True}
    # y = 121: the browser's value is the next double above the header's -- and the correctly rounded one
    assert np.nextafter(L.fao_js_log10(121.0), np.inf) == shipped[121]
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    assert all(float(mp.log10(y)) == t for y, t in shipped.items())
    # it cannot reach the gate: t / 3 is far from a power of ten's exponent boundary
    assert header_gate_v(121) == 4 == int(10 ** (shipped[121] / 3))
