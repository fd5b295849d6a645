"""Dubins word + length pinned on the reference's OWN code: next to the un-vendored pydubins its planner imports
(requirements.txt:14), the reference repository carries a pure-Python Dubins planner
(path_planner/utils/dubins_path.py: planning_from_origin tries LSL, RSR, LSR, RSL, RLR, LRL and keeps the first
minimum).  tests/golden/dubins_ref_golden.npz holds its answers on 2000 seeded cases (oracle/gen_golden.py dubins_ref);
oracle/dubins_port.py -- the restatement of dubins.c that stands in for pydubins wherever the reference's planner code
is run for goldens -- must choose the same word and the same length."""
import math
import os
import sys
import types

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import dubins_port as DP                      # noqa: E402
from oracle import gen_golden, ref_loader                 # noqa: E402

GOLD = os.path.join(HERE, "golden", "dubins_ref_golden.npz")
LEN_TOL = 1e-9          # metres, on lengths of 2 .. 80 m (the two codes order their float64 operations differently)


def test_cases_are_the_seeded_ones():
    g = np.load(GOLD)
    assert np.array_equal(g["cases"], gen_golden.dubins_ref_cases())
    assert set(np.unique(g["word"]).tolist()) == set(range(6))          # every word occurs


def test_port_equals_reference_dubins_path_golden():
    g = np.load(GOLD)
    for c, w, length in zip(g["cases"], g["word"], g["length"]):
        p = DP.shortest_path(tuple(c[:3]), tuple(c[3:6]), float(c[6]))
        assert p.type == int(w), (c, DP.WORD_NAMES[p.type], DP.WORD_NAMES[int(w)])
        assert abs(p.path_length() - length) <= LEN_TOL
        # and the port's own end pose is the goal (sampling the whole length)
        q = p.sample(p.path_length())
        assert math.hypot(q[0] - c[3], q[1] - c[4]) < 1e-6


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")
def test_live_reference_dubins_path_equals_golden():
    sys.modules.setdefault("draw", types.ModuleType("draw"))           # plotting helper of dubins_path.py
    mod = ref_loader.load("dubins_path")
    g = np.load(GOLD)
    word, length = gen_golden.dubins_ref_outputs(mod, g["cases"][:400])
    assert np.array_equal(word, g["word"][:400])
    assert np.array_equal(length, g["length"][:400])
