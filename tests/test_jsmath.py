"""include/fa_jsmath.h (fdlibm log/log10/pow as V8 ships them) against libm and on exact cases."""
import math
import random
import struct

from oracle import jsmath


def ulps(a: float, b: float) -> int:
    ia, ib = struct.unpack("<q", struct.pack("<d", a))[0], struct.unpack("<q", struct.pack("<d", b))[0]
    return abs(ia - ib)


def test_exact_powers_of_ten():
    for k in range(0, 16):
        assert jsmath.log10(10.0 ** k) == float(k)
        assert jsmath.pow(10.0, float(k)) == 10.0 ** k


def test_within_one_ulp_of_libm():
    rnd = random.Random(7)
    for _ in range(20000):
        y = float(rnd.randrange(1, 1 << 32))
        assert ulps(jsmath.log10(y), math.log10(y)) <= 1
        assert ulps(jsmath.log(y), math.log(y)) <= 1
        e = rnd.uniform(-0.5, 8.5)
        assert ulps(jsmath.pow(10.0, e), math.pow(10.0, e)) <= 1


def test_special_cases():
    assert jsmath.pow(10.0, 0.0) == 1.0
    assert jsmath.pow(10.0, 1.0) == 10.0
    assert jsmath.pow(10.0, 2.0) == 100.0
    assert jsmath.pow(10.0, 0.5) == math.sqrt(10.0)
    assert jsmath.pow(-2.0, 3.0) == -8.0
    assert math.isnan(jsmath.pow(-8.0, 1.0 / 3.0))
    assert jsmath.pow(2.0, -1074.0) == 5e-324
    assert jsmath.log10(0.0) == -math.inf
    assert math.isnan(jsmath.log10(-1.0))
    assert jsmath.log(1.0) == 0.0


def test_noise_gate_threshold_is_an_integer_function():
    # v = parseInt(pow(10, log10(y) - 2) / 2) for 1e4 < y <= 1e6 : close to y / 200, never off by more than 1
    for y in range(10001, 1000000, 997):
        t = jsmath.log10(float(y))
        v = math.trunc(jsmath.pow(10.0, t - 2) / 2)
        assert abs(v - y / 200.0) <= 1.0
