"""Mirror of ``path_planner/utils/a_star_utils.py``'s public entry point backed by the
tiled-wavefront kernel ``hl_distance_field``."""
import numpy as np

from .. import ops

BLOCK_COST = 100


def holonomic_motion_commands():
    return [[-1, 0], [-1, 1], [0, 1], [1, 1], [1, 0], [1, -1], [0, -1], [-1, -1]]


def forward_holonomic_motion_commands():
    return [[-1, 0], [0, 1], [-1, 1], [1, 1], [1, 0]]


def holonomic_costs_with_obstacles(goal_index, obstacles, motion_type="King"):
    """a_star_utils.py:75-142.  Grids whose border is not fully occupied would trigger the
    reference's index wrap-around (:54-61); those raise instead of returning different numbers."""
    out, _ = ops.distance_field(np.asarray(obstacles), goal_index, motion_type)
    return out.cpu().numpy()
