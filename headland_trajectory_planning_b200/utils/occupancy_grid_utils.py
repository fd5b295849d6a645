"""Mirror of the grid container of ``path_planner/utils/occupancy_grid_utils.py`` without
its ROS / cv2 imports (the reference module cannot be imported outside ROS, :2-8): same
grid layout -- 1-cell free padding, obstacle = cells equal to the map maximum, cell
centres at ``idx*res + res/2`` (:70-101) -- plus the bit-packed device copy the footprint
kernel reads."""
import numpy as np

from .. import ops


class GridMapFeatures:
    def __init__(self):
        self.obstacles_boolean = np.array([])
        self.obstacles_xs = np.array([])
        self.obstacles_ys = np.array([])
        self.obstacle_field_map = None
        self.resolution = None
        self.device_bits = None


def shrink_grid_map(occupancy_map, old_resolution, new_resolution):
    """occupancy_grid_utils.py:43-61 with nearest-neighbour index selection in numpy."""
    if new_resolution < old_resolution:
        return occupancy_map
    w, h = occupancy_map.shape
    rw, rh = int(w * (old_resolution / new_resolution)), int(h * (old_resolution / new_resolution))
    ii = np.minimum((np.arange(rw) * (w / rw)).astype(int), w - 1)
    jj = np.minimum((np.arange(rh) * (h / rh)).astype(int), h - 1)
    return occupancy_map[np.ix_(ii, jj)]


def get_grid_map_features(occupancy_map, resolution, map_2d_position=(0.0, 0.0), upload=True):
    """occupancy_grid_utils.py:70-101."""
    obstacle_value = np.max(occupancy_map)
    padded = np.pad(occupancy_map, 1, mode="constant", constant_values=0)      # padder=0: free border (:77)
    idx = np.where(padded == obstacle_value)
    f = GridMapFeatures()
    f.obstacles_xs = idx[0] * resolution + resolution / 2.0
    f.obstacles_ys = idx[1] * resolution + resolution / 2.0
    f.obstacles_boolean = np.zeros(padded.shape, dtype=bool)
    f.obstacles_boolean[idx] = True
    f.obstacle_field_map = padded
    f.resolution = resolution
    if upload:
        f.device_bits = ops.grid_pack(f.obstacles_boolean)
    return f


def check_poses_on_grid(features, car_model, path):
    """Grid footprint check (the reference's commented-out
    ``check_path_feasibility_with_grid_map``): uint8 CUDA tensor, 1 = body meets an occupied cell."""
    return ops.grid_footprint_check(features.device_bits, features.obstacles_boolean.shape, features.resolution,
                                    np.asarray(path, dtype=np.float64)[:, :3], car_model.body_ext)
