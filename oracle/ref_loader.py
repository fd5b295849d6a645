"""Import the reference's OWN importable modules from ``/root/reference`` --
build-container only (the GPU box has no ``/root/reference``).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Used by ``oracle/gen_golden.py``
and by CPU tests that are skipped when the checkout is absent.  Nothing under
``-m gpu``, ``smoke()`` or ``bench.py`` calls this.

Importable as-is: ``path_planner/utils/reeds_shepp.py``
(needs a ``matplotlib.pyplot`` stub: imported at ``:3`` but unused),
``path_planner/utils/a_star_utils.py``, ``path_planner/utils/path_utils.py``.
Everything else needs shapely / heapdict / dubins, which are not installable.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("HL_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "path_planner", "utils", "reeds_shepp.py"))


def _stub_matplotlib():
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
            import matplotlib.pyplot  # noqa: F401
            return
        except Exception:
            pass
        m = types.ModuleType("matplotlib")
        p = types.ModuleType("matplotlib.pyplot")
        m.pyplot = p
        sys.modules["matplotlib"] = m
        sys.modules["matplotlib.pyplot"] = p


def load(name):
    """Load ``path_planner/utils/<name>.py`` under a private module name."""
    _stub_matplotlib()
    path = os.path.join(REFERENCE_ROOT, "path_planner", "utils", name + ".py")
    spec = importlib.util.spec_from_file_location("_hl_reference_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Anything:
    """Attribute sink standing in for classes / functions of an absent third-party package."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise RuntimeError("stub of an absent third-party package was called")

    def __getattr__(self, name):
        return _Anything()


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything


_ABSENT = ("shapely", "heapdict", "dubins", "skspatial", "matplotlib", "casadi", "cv2", "pypoman", "rdp")


def load_planner(name, heapdict_port=True, dubins_port=False, extra=None):
    """Import ``path_planner/<name>.py`` of the reference FOR REAL, with the absent third-party packages
    (shapely, heapdict, dubins, skspatial, ...) replaced by inert stubs.  Everything that does not touch those
    packages -- the Y-type parking sweep, the motion-path rollout, the odom transform
    (``headland_path_planning.py:382-527``) -- then runs as the reference wrote it; geometry objects are passed
    in duck-typed (the oracle's environment / car model).  With ``heapdict_port`` the ``heapdict`` import
    resolves to ``oracle.heapdict_port`` so that the reference's own search loop
    (``hybrid_a_star_search.py``) runs too.  ``sys.modules`` / ``sys.path`` are restored."""
    import importlib
    import importlib.abc
    import importlib.machinery

    class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        def find_spec(self, fullname, path=None, target=None):
            if fullname.split(".")[0] in _ABSENT:
                return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
            return None

        def create_module(self, spec):
            return _StubModule(spec.name)

        def exec_module(self, module):
            pass

    saved_mods = dict(sys.modules)
    saved_path = list(sys.path)
    finder = _Finder()
    # flat stubs an earlier load() left behind (e.g. a non-package ``matplotlib``) would shadow the finder
    stash = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k.split(".")[0] in _ABSENT and getattr(sys.modules[k], "__file__", None) is None}
    try:
        sys.meta_path.insert(0, finder)
        if heapdict_port:
            from . import heapdict_port as _hp
            hd = types.ModuleType("heapdict")
            hd.heapdict = _hp.HeapDict
            sys.modules["heapdict"] = hd
        if dubins_port:                        # ``import dubins`` resolves to the restated dubins.c (oracle.dubins_port)
            from . import dubins_port as _dp
            sys.modules["dubins"] = _dp
        for k, v in (extra or {}).items():     # functional stand-ins a caller provides (load_oge_obca)
            sys.modules[k] = v
        sys.path.insert(0, os.path.join(REFERENCE_ROOT, "path_planner", "utils"))   # notebooks put both on sys.path
        sys.path.insert(0, os.path.join(REFERENCE_ROOT, "path_planner"))
        return importlib.import_module(name)
    finally:
        sys.meta_path.remove(finder)
        for k in list(sys.modules):
            if k not in saved_mods:
                del sys.modules[k]
        sys.modules.update(stash)
        sys.path[:] = saved_path


class _Ring:
    def __init__(self, pts):
        self.coords = [tuple(p) for p in pts]
        self.xy = ([p[0] for p in pts], [p[1] for p in pts])


class _MiniPolygon:
    """What ``OGE_OBCA.py`` and the constructor of ``orchard_geometry_environment.py`` read of a shapely Polygon:
    the closed exterior ring in the order given (shapely keeps the caller's orientation) and an empty ``interiors``."""

    def __init__(self, pts=()):
        import numpy as np
        pts = np.asarray(list(pts), dtype=float).reshape(-1, 2)
        if len(pts) and not (pts[0] == pts[-1]).all():
            pts = np.vstack([pts, pts[:1]])
        self.exterior = _Ring(pts)
        self.interiors = []


class _MiniGeom:
    def __init__(self, *a, **k):
        pass

    def buffer(self, *a, **k):
        return _MiniGeom()


def load_oge_obca():
    """``path_planner/OGE_OBCA.py`` of the reference FOR REAL (class ``orchard_environment_OBCA`` on top of the
    reference's own ``OrchardGeometryEnvironment``): shapely is replaced by the few attributes these two files read
    (``Polygon(pts).exterior.coords``; buffers and the STRtree are built and never queried by the obstacle extraction),
    ``rdp`` by the independent recursive restatement ``oracle.rdp_port`` (the package is not installable: parity with
    it unpinned), pypoman / matplotlib by inert stubs; numpy and scipy's ConvexHull are the real ones."""
    from . import rdp_port
    geom = types.ModuleType("shapely.geometry")
    geom.Polygon = _MiniPolygon
    geom.Point = _MiniGeom
    geom.LineString = _MiniGeom
    geom.MultiPolygon = _MiniGeom
    strtree = types.ModuleType("shapely.strtree")
    strtree.STRtree = _MiniGeom
    shp = _StubModule("shapely")
    shp.geometry = geom
    shp.strtree = strtree
    rdp_mod = types.ModuleType("rdp")
    rdp_mod.rdp = rdp_port.rdp
    return load_planner("OGE_OBCA", extra={"shapely": shp, "shapely.geometry": geom, "shapely.strtree": strtree, "rdp": rdp_mod})


def load_reference_line_heuristic():
    """``path_planner/reference_line_heuristic.py`` of the reference for real, for the parts that are numpy only: the
    guide polyline (``get_guide_line``, :50-82), the per-segment search lengths without obstacle polygons
    (``create_segment_lengths``, :84-96 -- the call sites pass none) and ``calculate_state_cost`` (:131-158).
    ``LineString(...).buffer(...)``, ``unary_union`` and ``STRtree`` are built and never queried by these, so they are
    inert stand-ins; the lane predicates (``get_search_length``, ``check_path_feasibility``) are NOT exercised
    (shapely: parity unpinned)."""
    geom = types.ModuleType("shapely.geometry")
    geom.Polygon = _MiniPolygon
    geom.Point = _MiniGeom
    geom.LineString = _MiniGeom
    geom.MultiPolygon = _MiniGeom
    strtree = types.ModuleType("shapely.strtree")
    strtree.STRtree = _MiniGeom
    ops_mod = types.ModuleType("shapely.ops")
    ops_mod.unary_union = lambda geoms: _MiniGeom()
    shp = _StubModule("shapely")
    shp.geometry, shp.strtree, shp.ops = geom, strtree, ops_mod
    return load_planner("reference_line_heuristic", extra={"shapely": shp, "shapely.geometry": geom,
                                                           "shapely.strtree": strtree, "shapely.ops": ops_mod})


def load_car_model():
    """``path_planner/car_model.py`` of the reference for real, for the part of the footprint that is numpy only: WHERE
    the body rectangle and the implement rectangles of a path are (``get_car_poly`` :75-143, ``get_aux_shapely_polys``
    :146-162, ``get_path_poly`` :39-73 -- rotation by a 2 x 2 ``np.dot``, implements on every second pose).  ``Polygon``
    keeps the vertices it is given and ``unary_union`` returns the list of polygons instead of dissolving them, so the
    caller sees every rectangle; what GEOS then DOES with them (intersects / contains) stays unpinned."""
    geom = types.ModuleType("shapely.geometry")
    geom.Polygon = _MiniPolygon
    geom.Point = _MiniGeom
    ops_mod = types.ModuleType("shapely.ops")
    ops_mod.unary_union = lambda polys: list(polys)
    shp = _StubModule("shapely")
    shp.geometry, shp.ops = geom, ops_mod
    return load_planner("car_model", extra={"shapely": shp, "shapely.geometry": geom, "shapely.ops": ops_mod})


def load_obca_util():
    """``obca_py/util.py`` of the reference for real (its ``car_model_obca`` import -- casadi -- is replaced by a stub;
    ``cubic_spline`` resolves to ``path_planner/utils/cubic_spline.py``, scipy is installed)."""
    saved_mods = dict(sys.modules)
    saved_path = list(sys.path)
    try:
        stub = types.ModuleType("car_model_obca")
        stub.CarModel = object
        sys.modules["car_model_obca"] = stub
        sys.path.insert(0, os.path.join(REFERENCE_ROOT, "path_planner", "utils"))
        path = os.path.join(REFERENCE_ROOT, "obca_py", "util.py")
        spec = importlib.util.spec_from_file_location("_hl_reference_obca_util", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        for k in list(sys.modules):
            if k not in saved_mods:
                del sys.modules[k]
        sys.path[:] = saved_path
