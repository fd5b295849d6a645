"""Host-side mirrors of the reference's ``path_planner/utils`` helpers that sit on the warm-start path
(``path_utils``, ``map_utils``, ``reeds_shepp``, ``a_star_utils``, ``occupancy_grid_utils``): same function names
and return conventions, geometry built with numpy, every compute call forwarded to the CUDA library."""
