"""Mirror of the two ``path_planner/utils/map_utils.py`` helpers the warm-start path and
its scenario generator use: the synthetic orchard (``create_tree_rows``, :45-61) and the
row leave / enter base poses (``get_base_pose``, :228-271)."""
import numpy as np

NEAR_SIDE = 1
FAR_SIDE = 2
LEAVE_POSE = 1
ENTER_POSE = 2


def create_tree_rows(row_num, row_width, row_lengths, slope_angle=0, l_std=0.0):
    rows = []
    delta_x = row_width * np.tan(slope_angle)
    for i in range(row_num):
        length = row_lengths[i] if isinstance(row_lengths, (list, np.ndarray)) else row_lengths
        x = delta_x * i + np.random.uniform(-l_std, l_std)
        y = row_width * i
        rows.append(np.array([[x, y], [x + length, y]]))
    return np.array(rows)


def get_base_pose(row_id, map_tree_rows, min_offset, side=NEAR_SIDE, pose_type=LEAVE_POSE):
    near = [(map_tree_rows[row_id, 0, 0] + map_tree_rows[row_id + 1, 0, 0]) / 2,
            (map_tree_rows[row_id, 0, 1] + map_tree_rows[row_id + 1, 0, 1]) / 2]
    far = [(map_tree_rows[row_id, 1, 0] + map_tree_rows[row_id + 1, 1, 0]) / 2,
           (map_tree_rows[row_id, 1, 1] + map_tree_rows[row_id + 1, 1, 1]) / 2]
    row_yaw = np.arctan2(far[1] - near[1], far[0] - near[0])
    if pose_type == LEAVE_POSE:
        pose_yaw = row_yaw if side == FAR_SIDE else row_yaw + np.pi
        extend_dir = 1
    else:
        pose_yaw = row_yaw + np.pi if side == FAR_SIDE else row_yaw
        extend_dir = -1
    end = near if side == NEAR_SIDE else far
    pos = end + np.array([np.cos(pose_yaw), np.sin(pose_yaw)]) * min_offset * extend_dir
    return np.array([pos[0], pos[1], pose_yaw])
