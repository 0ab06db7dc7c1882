"""The epilogue's rounding recipe (csrc/pc_common.cuh: pc_round6_micro) restated
with numpy fp64 + an exact-FMA emulation, pinned against CPython's round()."""
import math
from fractions import Fraction

import numpy as np

from conftest import load_golden


def round6_micro_recipe(x):
    """hi = fl(x*1e6); lo = exact(x*1e6) - hi; half-even on hi, lo breaks .5 ties."""
    hi = x * 1e6
    lo = float(Fraction(x) * 1000000 - Fraction(hi))      # what fma(x, 1e6, -hi) returns
    r = float(np.rint(hi))
    if abs(hi - r) == 0.5:
        fl = math.floor(hi)
        if lo > 0:
            r = fl + 1.0
        elif lo < 0:
            r = float(fl)
    return int(r)


def test_recipe_equals_python_round_on_golden():
    for case in load_golden("round6.json"):
        x = float.fromhex(case["x"])
        assert round6_micro_recipe(x) == case["micro"] == int(round(round(x, 6) * 1e6))


def test_recipe_on_random_ratios_and_complements():
    rng = np.random.default_rng(0)
    for _ in range(20000):
        den = int(rng.integers(1, 100000))
        num = int(rng.integers(0, den + 1))
        for x in (num / den, 1.0 - num / den, (2.0 * num) / (den + num)):
            assert round6_micro_recipe(x) == int(round(round(x, 6) * 1e6))


def test_micro_units_reproduce_the_rounded_float_exactly():
    # host side: n / 1e6 is the double round(x, 6) returns; inversion is exact in integers
    rng = np.random.default_rng(1)
    for _ in range(20000):
        x = float(rng.random())
        n = int(round(round(x, 6) * 1e6))
        assert n / 1e6 == round(x, 6)
        assert round(1.0 - n / 1e6, 6) == (1000000 - n) / 1e6
