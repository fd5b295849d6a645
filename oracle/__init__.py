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
  * Everything that sits on shapely/GEOS, heapdict or pydubins in the reference
    (footprint predicates, lane containment, open-list tie order): **parity
    unpinned** -- those third-party packages are not installable here, so the
    restatement below defines parity; the notebook golden values that do exist
    (tree-row seed, poses, printed polygons, curvature, ``counter of nodes: 1``)
    are checked in ``tests/test_oracle_golden.py``.
"""
