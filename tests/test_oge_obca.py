"""OBCA obstacle extraction (``path_planner/OGE_OBCA.py``): the mirror
``headland_trajectory_planning_b200.OGE_OBCA.orchard_environment_OBCA`` against goldens made by running the REFERENCE's
own class (``oracle/gen_golden.py oge``: ``create_boundary_polygons`` / ``get_obstacle_tree_rows`` /
``get_tree_row_obstacles`` / ``get_obstacles_for_OBCA`` of ``test/obca.ipynb`` cell 12 on six orchards), and live against
the reference when the checkout is present.  Host-only: no GPU needed."""
import os

import numpy as np
import pytest

from oracle import gen_golden as GG
from oracle import ref_loader

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oge_golden.npz")


def _mirror_outputs(case):
    from headland_trajectory_planning_b200.OGE_OBCA import orchard_environment_OBCA
    return GG.oge_outputs(orchard_environment_OBCA, case)


def _same(got, want, tol=0.0):
    assert len(got) == len(want)
    for g, w in zip(got, want):
        g, w = np.asarray(g, dtype=float), np.asarray(w, dtype=float)
        assert g.shape == w.shape
        if tol:
            np.testing.assert_allclose(g, w, rtol=0, atol=tol)
        else:
            assert np.array_equal(g, w)


def test_mirror_equals_reference_golden():
    z = np.load(GOLDEN, allow_pickle=False)
    n_cases = int(z["n_cases"])
    assert n_cases == len(GG.oge_cases())
    for c, case in enumerate(GG.oge_cases()):
        got = _mirror_outputs(case)
        for name, polys in got.items():
            want = GG.unpack_polys(z[f"c{c}_{name}_v"], z[f"c{c}_{name}_n"])
            _same(polys, want)


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout absent")
def test_mirror_equals_live_reference():
    ref = ref_loader.load_oge_obca()
    for case in GG.oge_cases():
        want = GG.oge_outputs(ref.orchard_environment_OBCA, case)
        got = _mirror_outputs(case)
        assert got.keys() == want.keys()
        for name in want:
            _same(got[name], want[name])


def test_rdp_restatements_agree():
    """The product's iterative RDP and the oracle's recursive one (two independent restatements of the ``rdp`` package,
    which is not installable: parity with it unpinned) on random polylines, straight runs, repeated and closed points."""
    from headland_trajectory_planning_b200.OGE_OBCA import rdp as rdp_iter
    from oracle.rdp_port import rdp as rdp_rec
    rng = np.random.default_rng(3)
    for k in range(200):
        n = int(rng.integers(2, 40))
        pts = np.cumsum(rng.normal(size=(n, 2)), axis=0)
        if k % 5 == 0:
            pts[:, 0] = np.linspace(0, 5, n)
            pts[:, 1] = 0.3 * pts[:, 0] + (rng.normal(size=n) * 0.01 if k % 10 else 0.0)
        if k % 7 == 0 and n > 3:
            pts[-1] = pts[0]
        eps = float(rng.choice([0.0, 0.05, 0.15, 1.0]))
        assert np.array_equal(rdp_iter(pts, eps), rdp_rec(pts, eps))
