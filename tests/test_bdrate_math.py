"""The Bjontegaard functions behind the bench line's `bd_rate` (tools/bdrate.py): known answers."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from bdrate import bd_rate, bd_rate_pchip


def test_known_answers():
    q = [34.0, 37.0, 40.0, 43.0, 46.0]
    r = [100.0, 180.0, 330.0, 640.0, 1300.0]
    for f in (bd_rate, bd_rate_pchip):
        assert abs(f(r, q, r, q)) < 1e-9                                         # a curve against itself
        assert abs(f(r, q, [1.5 * x for x in r], q) - 50.0) < 1e-6               # 1.5 x the rate at every quality
        assert abs(f(r, q, [x / 2 for x in r], q) + 50.0) < 1e-6
        assert f(r, q, r, [x + 20 for x in q]) is None                           # no common quality range
        # antisymmetry in the log domain: (1 + a)(1 + b) = 1
        a, b = f(r, q, [1.3 * x for x in r], q), f([1.3 * x for x in r], q, r, q)
        assert abs((1 + a / 100) * (1 + b / 100) - 1) < 1e-9
    # a curve shifted in quality: the same rates buy 1 dB less; on an exactly exponential rate curve (log-rate linear in quality,
    # 0.1 per dB here) that is exp(0.1) - 1 = +10.5 % whatever the interpolation, over the common range
    qq = np.array([30.0, 34.0, 38.0, 42.0, 46.0])
    rr = np.exp(0.1 * qq)
    for f in (bd_rate, bd_rate_pchip):
        assert abs(f(rr, qq, rr, qq - 1.0) - (np.exp(0.1) - 1) * 100) < 1e-6
