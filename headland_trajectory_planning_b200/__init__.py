"""headland_trajectory_planning_b200 -- B200 (sm_100a) warm-start search for the OBCA
headland planner: Hybrid A* primitive expansion, Reeds-Shepp analytic shots,
footprint collision checking and search heuristics as hand-written CUDA behind a
C ABI (``include/headland_b200.h``), with host-side mirrors of the reference's
Python classes so the reference's callers keep working unchanged.

Mirror modules (same names as the reference's ``path_planner`` modules):
``car_model``, ``orchard_geometry_environment``, ``reference_line_heuristic``,
``hybrid_a_star_search``, ``reeds_shepp``, ``a_star_utils``, ``path_utils``,
``map_utils``.  There is no CPU fallback: importing is cheap, but any compute call
raises ``HeadlandError`` when the CUDA library or a GPU is missing.
"""
from ._lib import HeadlandError, library_path, have_gpu  # noqa: F401

__version__ = "0.1.0"
