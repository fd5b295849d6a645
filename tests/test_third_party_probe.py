"""Run-time probe for the REAL third-party packages the reference uses on this path.

shapely (GEOS) and heapdict are not installable in the build container (no wheel, no network), so the oracle restates
their semantics (``oracle/geometry.py``, ``oracle/heapdict_port.py``) and the tests above pin those restatements
by exact-rational brute force and by the reference's notebook outputs.  Wherever the packages DO import -- a
developer machine, a future image -- these tests compare the restatements with the real thing on 10^4 random cases
and the parity claim pins itself.  They are skipped (not failed) when the packages are absent."""
import math

import numpy as np
import pytest

import hl_helpers as H
from oracle import geometry as geo
from oracle import planner as OP


def test_geometry_predicates_equal_shapely():
    shapely = pytest.importorskip("shapely")
    from shapely.geometry import LineString, Point, Polygon
    from shapely.ops import unary_union
    rng = np.random.default_rng(0)
    rows = H.canonical_rows(l_std=1.0)
    env = OP.OrchardGeometryEnvironment(rows, [(-3.0, 6.0)], tree_width=0.3, headland_width=6.0)
    ext = geo.body_extent(0.55, 3.0, 1.48)
    poses = np.concatenate([H.random_poses(rng, 5000), H.headland_poses(rng, 5000, rows)])
    corners = geo.rect_corners(poses, ext)
    obstacles = [Polygon(q) for q in env.obs_poly_list]
    field = Polygon(env.field_poly)
    # the polygons the reference itself builds (orchard_geometry_environment.py:277-286, :350)
    for r, q in zip(env.map_tree_rows, env.tree_polys):
        ref = LineString(r).buffer(env.tree_width / 2, cap_style=2)
        assert Polygon(q).symmetric_difference(ref).area < 1e-12
    sq = Point(-3.0, 6.0).buffer(0.3, cap_style=3)
    assert Polygon(geo.point_square_buffer(-3.0, 6.0, 0.3)).symmetric_difference(sq).area < 1e-12
    hit = np.zeros(len(poses), dtype=bool)
    for q in env.obs_poly_list:
        hit |= geo.rects_hit_convex(poses, ext, geo.ccw(q), corners)
    inside = geo.rects_inside_polygon(poses, ext, geo.ccw(env.field_poly), corners)
    for p in range(len(poses)):
        body = Polygon(corners[p])
        assert hit[p] == any(o.intersects(body) for o in obstacles), poses[p]
        assert inside[p] == field.contains(body), poses[p]
    # lane = union of GEOS round-cap buffers (reference_line_heuristic.py:65-67, 81)
    way = np.array([[0.0, 3.75], [-4.5, 5.0], [-4.06, 7.5], [-3.62, 10.0], [-1.5, 11.0]])
    lane = geo.Lane(way)
    caps = [LineString([a, b]).buffer(6, cap_style=1, join_style=3) for a, b in zip(way[:-1], way[1:])]
    for mine, theirs in zip(lane.polys, caps):
        ring = np.asarray(theirs.exterior.coords)[:-1]
        assert len(ring) == len(mine) == 66
        # same vertex set (GEOS may start the ring elsewhere / wind the other way)
        d = np.abs(mine[:, None, :] - ring[None, :, :]).sum(axis=2).min(axis=1)
        assert d.max() < 1e-9
    union = unary_union(caps)
    lp = H.random_poses(rng, 4000, box=(-12.0, 6.0, -4.0, 18.0))
    lc = geo.rect_corners(lp, ext)
    got = lane.rects_inside(lp, ext, lc)
    for p in range(len(lp)):
        assert got[p] == union.contains(Polygon(lc[p])), lp[p]
    for x, y in rng.uniform(-12, 6, (2000, 2)):
        want = -1
        for i, c in enumerate(caps):
            if c.contains(Point(x, y)):
                want = i
        assert lane.search_segment(x, y) == want


def test_heapdict_port_equals_heapdict():
    heapdict = pytest.importorskip("heapdict")
    from oracle.heapdict_port import HeapDict
    rng = np.random.default_rng(1)
    for _ in range(300):
        a, b = heapdict.heapdict(), HeapDict()
        keys = list(range(int(rng.integers(1, 40))))
        for step in range(int(rng.integers(1, 120))):
            op = rng.random()
            if op < 0.65 or len(a) == 0:
                k = keys[rng.integers(len(keys))]
                pr = float(rng.integers(0, 6))                 # few distinct priorities: ties are the norm
                a[k] = pr
                b[k] = pr
            else:
                assert a.popitem() == b.popitem()
        while len(a):
            assert a.popitem() == b.popitem()
        assert len(b) == 0


def test_reference_modules_run_unmodified_when_dependencies_exist():
    """With shapely + heapdict (+ the reference checkout) present, the UNMODIFIED reference classes give the oracle's
    booleans and node sequence on a config-5 scenario."""
    pytest.importorskip("shapely")
    pytest.importorskip("heapdict")
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference checkout absent")
    import sys
    import os
    for sub in ("path_planner", os.path.join("path_planner", "utils")):
        sys.path.insert(0, os.path.join(ref_loader.REFERENCE_ROOT, sub))
    try:
        from car_model import CarModel
        from orchard_geometry_environment import OrchardGeometryEnvironment
    except Exception as exc:                                    # e.g. matplotlib / dubins missing
        pytest.skip(f"reference modules do not import: {exc!r}")
    rows = H.canonical_rows(l_std=0.5)
    r_env = OrchardGeometryEnvironment(rows, [], tree_width=0.3, headland_width=6.0)
    r_car = CarModel(max_steer=0.55, axle_to_front=3, axle_to_back=0.55, width=1.48)
    o_env = OP.OrchardGeometryEnvironment(rows, [], tree_width=0.3, headland_width=6.0)
    o_car = OP.CarModel(max_steer=0.55, axle_to_front=3, axle_to_back=0.55, width=1.48)
    rng = np.random.default_rng(2)
    for pose in H.headland_poses(rng, 2000, rows):
        p = pose[None, :]
        assert r_env.check_path_feasibility(r_car, p) == o_env.check_path_feasibility(o_car, p), pose
