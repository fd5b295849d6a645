"""CPU: host-side logic of the product package -- the C-ABI library loads and exports
every symbol the header declares, struct layouts match the header, host geometry equals
the oracle's, scenarios are deterministic, compute calls fail loudly without a GPU."""
import ctypes
import math
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import hl_helpers as H
from oracle import geometry as geo
from oracle import planner as OP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "headland_b200.h")


def test_library_exports_every_declared_symbol(built_library):
    from headland_trajectory_planning_b200 import _lib
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(hl_[a-z0-9_]+)\s*\(", text))
    declared -= {"hl_ctx", "hl_env_batch"}
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = ctypes.CDLL(built_library)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.hl_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header(built_library):
    from headland_trajectory_planning_b200 import _lib
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "headland_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(HlEnvHost), sizeof(HlSearchParams), sizeof(HlScenario), sizeof(HlPlanResult), sizeof(HlRsWord));
  printf("%zu %zu %zu %zu\n", offsetof(HlPlanResult, path_offset), offsetof(HlPlanResult, n_exact), offsetof(HlRsWord, nlen), offsetof(HlSearchParams, max_nodes));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe], text=True).split()
    sizes = [int(v) for v in out]
    assert sizes[0] == ctypes.sizeof(_lib.HlEnvHost)
    assert sizes[1] == ctypes.sizeof(_lib.HlSearchParams)
    assert sizes[2] == _lib.SCENARIO_DTYPE.itemsize
    assert sizes[3] == _lib.RESULT_DTYPE.itemsize
    assert sizes[4] == _lib.RSWORD_DTYPE.itemsize
    assert sizes[5] == _lib.RESULT_DTYPE.fields["path_offset"][1]
    assert sizes[6] == _lib.RESULT_DTYPE.fields["n_exact"][1]
    assert sizes[7] == _lib.RSWORD_DTYPE.fields["nlen"][1]
    assert sizes[8] == _lib.HlSearchParams.max_nodes.offset


def test_compute_fails_loudly_without_gpu(built_library):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from headland_trajectory_planning_b200 import HeadlandError
    rows = H.canonical_rows()
    _, (g_env, g_car, _) = H.make_pair(rows)
    with pytest.raises(HeadlandError):
        g_env.check_path_feasibility(g_car, np.array([[0.0, 3.75, 0.0]]))
    from headland_trajectory_planning_b200.utils import reeds_shepp
    with pytest.raises(HeadlandError):
        reeds_shepp.calc_all_paths(0, 0, 0, 3, 4, 1.0, 0.3)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "headland_trajectory_planning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py") or f.endswith(".cu") or f.endswith(".cuh"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "from oracle" not in text and "import oracle" not in text, f


def test_host_geometry_equals_oracle():
    rows = H.canonical_rows(l_std=1.0)
    way = np.array([[0.4, 3.75], [-4.0, 5.0], [-3.6, 7.5], [-1.5, 8.75]])
    (o_env, o_car, o_h), (g_env, g_car, g_h) = H.make_pair(rows, obstacles=[(-3.0, 6.0)], waypoints=way,
                                                            goal=[-1.5, 8.75, 0.0], aux=H.MOWER_AUX, axle_to_front=2.85)
    assert np.array_equal(g_env.field_ring(), o_env.field_poly)
    assert np.array_equal(g_env.obstacle_quads(), np.array(o_env.obs_poly_list))
    assert np.array_equal(g_h.seg_polys, np.array(o_h.lane.polys))
    assert g_h.seg_polys.shape[1] == 66
    assert np.array_equal(g_h.guided_path, o_h.guided_path)
    assert np.array_equal(g_h.search_lengths, o_h.search_lengths)
    assert np.allclose(g_h.crit_xy, o_h.lane.critical, atol=1e-12) and len(g_h.crit_xy) > 0
    assert tuple(g_car.body_ext) == tuple(o_car.body_ext)
    assert np.array_equal(g_car.aux_exts, np.array(o_car.aux_exts))
    assert g_car.curvature == o_car.curvature == math.tan(0.55) / 1.9
    # capsule polygon: every vertex within the radius, caps inscribed at 6*cos(pi/64)
    seg = way[:2]
    poly = g_h.seg_polys[0]
    d = np.array([OP.geo.Lane(way).__class__._seg_dist(OP.geo.Lane(way), 0, np.array([p[0]]), np.array([p[1]]))[0] for p in poly])
    assert d.max() <= 6.0 + 1e-9 and d.min() >= 6.0 * math.cos(math.pi / 64) - 1e-6


def test_search_params_table():
    from headland_trajectory_planning_b200.car_model import CarModel
    from headland_trajectory_planning_b200.hybrid_a_star_search import make_search_params
    car = CarModel(max_steer=0.55, axle_to_front=3, axle_to_back=0.55, width=1.48)
    p, steers = make_search_params(car, "King", plan_resolution=0.2, max_nodes=400)
    o = OP.HybridAStarSearch([0, 0, 0], [1, 1, 0], None, OP.CarModel(max_steer=0.55, axle_to_front=3, axle_to_back=0.55, width=1.48),
                             None, motion_type="King", plan_resolution=0.2)
    assert p.n_prims == 14 and np.array_equal(steers, o.motion_steers)
    for i in range(14):
        st, d = o.motion_steers[i]
        assert p.prim_yaw_step[i] == d * 0.2 / 1.9 * math.tan(st)
        assert p.prim_curv[i] == np.tan(st) / 1.9
    # Pawn (hybrid_a_star_search.py:331-341): 8 forward-only steers, the last one beyond -MAX_STEER
    pp, psteers = make_search_params(car, "Pawn", plan_resolution=0.1)
    op = OP.HybridAStarSearch([0, 0, 0], [1, 1, 0], None, OP.CarModel(max_steer=0.55, axle_to_front=3, axle_to_back=0.55, width=1.48),
                              None, motion_type="Pawn", plan_resolution=0.1)
    assert pp.n_prims == 8 and pp.motion_type == 1 and np.array_equal(psteers, op.motion_steers)
    assert (psteers[:, 1] == 1).all() and abs(psteers[-1, 0] + 0.67173048) < 1e-8
    with pytest.raises(ValueError):
        make_search_params(car, "Rook")


def test_scenarios_are_deterministic_plain_data():
    from headland_trajectory_planning_b200 import scenarios as SC
    a, b = SC.scenario_spec(17), SC.scenario_spec(17)
    assert np.array_equal(a["rows"], b["rows"]) and np.array_equal(a["start"], b["start"])
    assert all(np.array_equal(x, y) for x, y in zip(a["ypark_candidates"], b["ypark_candidates"]))
    assert len(a["ypark_candidates"]) == len(SC.YPARK_CANDIDATES) == 16
    feas = [False] * 16
    feas[5] = True
    s1, s2 = SC.finalize(a, feas), SC.finalize(b, feas)
    assert s1["ypark_pick"] == 5 and np.array_equal(s1["waypoints"], s2["waypoints"])
    assert np.array_equal(s1["goal"][:2], a["ypark_candidates"][5][0][:2])
    none = SC.finalize(a, [False] * 16)
    assert none["ypark_pick"] == -1 and np.allclose(none["goal"][:2], a["end"][:2])
    specs = [SC.scenario_spec(i) for i in range(40)]
    assert len({round(s["headland_width"], 6) for s in specs}) == 40
    assert {s["side"] for s in specs} == {1, 2}


def test_shard_indices_interleave():
    from headland_trajectory_planning_b200 import sweep
    all_idx = sorted(i for r in range(8) for i in sweep.shard_indices(4096, r, 8))
    assert all_idx == list(range(4096))
    assert sweep.shard_indices(10, 1, 4) == [1, 5, 9]


@pytest.mark.timeout(60)
def test_failing_ctx_create_raises_instead_of_deadlocking(built_library, monkeypatch):
    """hl_ctx_create fails in this container (no CUDA device).  get_ctx used to call check() -> last_error() ->
    load_library() while holding the non-reentrant module lock: the failure hung forever instead of raising."""
    import torch
    from headland_trajectory_planning_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("needs a box without a CUDA device")
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    with pytest.raises(_lib.HeadlandError, match="hl_ctx_create"):
        _lib.get_ctx(None)
    assert 0 not in _lib._ctxs
