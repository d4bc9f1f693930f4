"""Fixture generator for tests/test_gate_enumeration.py (run in the build container: ~1 minute on 8 cores).

1. oracle/gate/gate_enum.c walks EVERY reachable y in [1, 2^32) through the noise gate's y -> v map (C() @B28506) as
   include/fa_jsmath.h computes it and compares with exact integer arithmetic: off the integer points of the map nothing may
   differ (engine-independent region); on them it records which side (q or q - 1) the fdlibm restatement lands.
2. mpmath re-evaluates the 223 715 integer points with every operation correctly rounded (log10, the subtraction / division
   by 3, pow, the division) -- the answer of an ideal libm -- and the points where that differs from the header are stored too.
Outputs: tests/golden/gate_enumeration.json (summary) and tests/golden/gate_points.npz (low_points, cr_differs)."""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def boundary_points():
    for r in range(5, 22):
        yield r ** 3, 2
    for y in range(10200, 1000001, 200):
        yield y, 4
    for y in range(1002000, 10000001, 2000):
        yield y, 6
    for y in range(10020000, 1 << 32, 20000):
        yield y, 7


def cr_gate_v(y, br, mp):
    """v with every floating-point operation correctly rounded (an ideal libm)."""
    t = float(mp.log10(y))
    if br == 2:
        return int(float(mp.power(10, mp.mpf(t / 3))))
    a, dv = (t - 2, 2) if br == 4 else (t - 3, 2) if br == 6 else (t - 3, 20)
    return int(float(mp.power(10, mp.mpf(a))) / dv)


def exact_q(y, br):
    return round(y ** (1 / 3)) if br == 2 else y // {4: 200, 6: 2000, 7: 20000}[br]


def main():
    import mpmath as mp
    mp.mp.prec = 240
    from oracle import build
    exe = build.build_gate_enum()
    d = json.loads(subprocess.check_output([exe], env=dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))))
    low = np.array(d.pop("low_points"), np.uint32)
    lowset = set(low.tolist())
    differ = []
    n = cr_low = 0
    for y, br in boundary_points():
        n += 1
        q = exact_q(y, br)
        v_cr = cr_gate_v(y, br, mp)
        v_h = q - 1 if y in lowset else q
        cr_low += v_cr == q - 1
        if v_cr != v_h:
            differ.append(y)
    assert n == d["boundary_points"]
    d["correctly_rounded_libm"] = {"boundary_points_low": cr_low, "differs_from_header": len(differ)}
    json.dump(d, open(os.path.join(HERE, "gate_enumeration.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(HERE, "gate_points.npz"), low_points=low, cr_differs=np.array(differ, np.uint32))
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main()
