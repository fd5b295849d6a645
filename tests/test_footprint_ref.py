"""WHERE the footprint rectangles are, pinned on the reference's OWN car_model.py (get_car_poly, get_aux_shapely_polys,
get_path_poly run unmodified; shapely reduced to a Polygon that keeps its vertices and a unary_union that returns the
list): body rectangle at every pose, implement rectangles at every SECOND pose (car_model.py:58), vertex order, the
feature -> rectangle rule.  The oracle's `rect_corners` (and through it the float64 predicates of the device, which
use the same expression order) must give the same vertices -- to 2 ulp, because the reference rotates with a 2 x 2
`np.dot` whose BLAS kernel may fuse the multiply-add (the reference itself is only reproducible to that).  What GEOS
does with the rectangles (intersects / contains) is not touched by this test and stays unpinned.
Golden: tests/golden/footprint_ref_golden.npz (written by running this file)."""
import math
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import geometry as geo                                          # noqa: E402
from oracle import planner as OP                                            # noqa: E402
from oracle import ref_loader                                               # noqa: E402

GOLD = os.path.join(HERE, "golden", "footprint_ref_golden.npz")
CARS = [dict(max_steer=0.55, axle_to_back=0.55, width=1.48),                                     # notebook tractor
        dict(max_steer=0.55, axle_to_front=3, axle_to_back=0.55, width=1.48),                    # planner car of test/obca.ipynb
        dict(max_steer=0.55, axle_to_back=0.55, width=1.48, with_aux=True,
             aux_poly_features=[[[-1.84, 0.5], 1.0, 1.1]]),                                       # mower (test/obca.ipynb cell 5)
        dict(max_steer=0.5, axle_to_front=4.2, axle_to_back=0.7, width=1.6, with_aux=True,
             aux_poly_features=[[[-2.5, 0.9], 1.8, 1.2], [[4.3, 0.4], 0.8, 0.6]])]


def paths():
    rng = np.random.default_rng(21)
    return [np.column_stack([rng.uniform(-40, 40, n), rng.uniform(-40, 40, n), rng.uniform(-2 * math.pi, 2 * math.pi, n)])
            for n in (1, 2, 7, 64, 129)]


def reference_vertices(mod):
    """[car][path] -> (body [P,4,2], aux [n_aux][ceil(P/2),4,2]) from the reference's polygons (closing vertex dropped)."""
    out = []
    for kw in CARS:
        car = mod.CarModel(**kw)
        for path in paths():
            body, aux = car.get_path_poly(path)
            out.append(np.array([np.asarray(p.exterior.coords)[:4] for p in body]).reshape(-1))
            for polys in aux:
                out.append(np.array([np.asarray(p.exterior.coords)[:4] for p in polys]).reshape(-1))
    return np.concatenate(out)


def oracle_vertices():
    out = []
    for kw in CARS:
        car = OP.CarModel(**kw)
        for path in paths():
            out.append(geo.rect_corners(path, car.body_ext).reshape(-1))
            for ext in car.aux_exts:
                # the reference lists an implement rectangle as (x, y), (x + w, y), (x + w, y - h), (x, y - h): the
                # oracle's corner order (x0,y1), (x0,y0), (x1,y0), (x1,y1) rotated -- compare as vertex SETS per rectangle
                out.append(geo.rect_corners(path[::2], ext).reshape(-1))
    return np.concatenate(out)


def _same_rectangles(a, b, ulps=2):
    a, b = a.reshape(-1, 4, 2), b.reshape(-1, 4, 2)
    assert a.shape == b.shape
    tol = ulps * np.spacing(np.maximum(np.abs(a).max(), 64.0))
    for ra, rb in zip(a, b):
        # match every vertex of ra to one of rb (order may be rotated / reversed between the two codes)
        d = np.abs(ra[:, None, :] - rb[None, :, :]).max(axis=2)
        assert (d.min(axis=1) <= tol).all() and (d.min(axis=0) <= tol).all(), (ra, rb)


def test_oracle_footprints_equal_reference_golden():
    _same_rectangles(oracle_vertices(), np.load(GOLD)["vertices"])


def test_body_rectangle_vertex_order_is_the_reference_s():
    """For the body the ORDER matters too (the device's separating-axis code walks the edges in it)."""
    g = np.load(GOLD)["vertices"]
    car = OP.CarModel(**CARS[0])
    p = paths()[0]
    want = g[:8].reshape(4, 2)
    got = geo.rect_corners(p, car.body_ext).reshape(4, 2)
    assert np.allclose(got, want, rtol=0, atol=2 * np.spacing(64.0))


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")
def test_live_reference_footprints_equal_golden():
    assert np.array_equal(reference_vertices(ref_loader.load_car_model()), np.load(GOLD)["vertices"])


if __name__ == "__main__":
    v = reference_vertices(ref_loader.load_car_model())
    np.savez_compressed(GOLD, vertices=v)
    print("footprint golden:", v.shape)
