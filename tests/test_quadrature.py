"""Quadrature rules of the product (lcu_quad_rule) and of the oracle against
float32 copies of the reference's own tables packed as its quad_rule() does
(tests/golden/quad_rules.npz, made by tools/make_golden.py)."""
import os

import numpy as np
import pytest

import lensed_b200 as L
from oracle import pyoracle as O

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "quad_rules.npz"))
RULES = ["point", "sub2", "sub4", "gm75", "g3k7", "g5k11", "g7k15"]
SIZES = dict(point=1, sub2=4, sub4=16, gm75=17, g3k7=49, g5k11=121, g7k15=225)


@pytest.mark.parametrize("rule", RULES)
@pytest.mark.parametrize("tag,sx,sy", [("", 1.0, 1.0), ("_scaled", 0.75, -1.25)])
def test_rule_matches_reference_tables_bit_for_bit(rule, tag, sx, sy):
    for name, (qq, ww) in (("product", L.quad_rule(rule, sx, sy)), ("oracle", O.quad_rule(rule, sx, sy))):
        assert qq.shape == (SIZES[rule], 2)
        assert np.array_equal(qq.view(np.uint32), GOLD[f"{rule}{tag}_qq"].view(np.uint32)), f"{name} {rule} abscissae"
        assert np.array_equal(ww.view(np.uint32), GOLD[f"{rule}{tag}_ww"].view(np.uint32)), f"{name} {rule} weights"


@pytest.mark.parametrize("rule", RULES)
def test_rule_properties(rule):
    qq, ww = L.quad_rule(rule)
    assert abs(ww[:, 0].astype(np.float64).sum() - 1) < 1e-6          # weights sum to one
    assert abs(ww[:, 1].astype(np.float64).sum()) < 1e-6              # error weights to zero
    assert np.abs(qq).max() < 0.5                                      # inside the pixel
    # the polynomial rules integrate a quadratic over the unit pixel exactly
    if rule in ("gm75", "g3k7", "g5k11", "g7k15"):
        x, y = qq[:, 0].astype(np.float64), qq[:, 1].astype(np.float64)
        assert abs((ww[:, 0]*(x*x + 0.5*x*y + y)).sum() - 1/12) < 1e-6
