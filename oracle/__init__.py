"""CPU oracle for the warm-start search path -- TEST INFRASTRUCTURE ONLY.

This package is a float64 CPU restatement of the reference's warm-start
search (Hybrid A* primitive expansion, Reeds-Shepp analytic shots, footprint
collision checking, guide-line heuristic, grid distance field).  Every function
cites the reference file:line it follows (paths relative to the reference
checkout).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker / CPU baseline.  The product package
``headland_trajectory_planning_b200`` never imports it and has no CPU fallback.

Pinning status (see DESIGN.md "Oracle"):
  * Reeds-Shepp (``rs_port``) and the grid distance field
    (``distance_field``): PINNED -- checked bit-for-bit against the reference's
    own importable modules (``path_planner/utils/reeds_shepp.py``,
    ``path_planner/utils/a_star_utils.py``) by ``oracle/gen_golden.py``; the
    vectors are committed under ``tests/golden/``.
  * ``path_utils`` helpers: PINNED the same way.
  * The search loop itself (``planner.HybridAStarSearch``: pop order, tolerance arrival,
    analytic shot and its cost queue, primitive rollout, g-costs, merge rules, path
    reconstruction) and the Y-type parking sweep (``planner.search_y_type_parking_path``):
    PINNED on the reference's OWN code -- ``oracle/ref_loader.load_planner`` imports
    ``path_planner/hybrid_a_star_search.py`` / ``headland_path_planning.py`` unmodified
    (shapely / dubins / skspatial replaced by inert stubs, ``heapdict`` resolved to
    ``oracle/heapdict_port``) and runs them on the oracle's duck-typed environment, car and
    heuristic objects; 192 config-5 scenarios (counters 1 .. 401) and 108 sweeps agree bit for
    bit / to 1 ulp (``tests/golden/astar_ref_golden.npz``, ``ypark_golden.npz``,
    ``tests/test_ypark_golden.py::test_live_reference_sweep_and_search_loop``).
  * What sits on shapely/GEOS, heapdict or pydubins in the reference
    (footprint predicates, lane containment / guide-line buffers, the heapdict package's
    tie order): **parity unpinned** -- those third-party packages are not installable here, so the
    restatement below defines parity; the notebook golden values that do exist
    (tree-row seed, poses, printed polygons, curvature, ``counter of nodes: 1``)
    are checked in ``tests/test_oracle_golden.py``.
"""
