"""Mirror of ``path_planner/reference_line_heuristic.py``: the guide polyline, the
6 m lane (as capsule polygons) and the per-segment search lengths are built on the
host exactly like the reference builds them (numpy ``linspace`` / ``hypot`` /
``cumsum``); the per-pose queries (lane containment, search length, state cost) run
on the GPU -- inside the search kernel, or through ``hl_collision_check`` for the
stand-alone ``check_path_feasibility``."""
import math

import numpy as np

from . import ops
from .env_batch import EnvBatch, EnvRecord
from .geometry_host import capsule_polygon, lane_critical_points


class ReferenceLineHeuristic(object):
    ACCEPT_PATH_DEVIATION = 2
    DRIVE_ROW_OFFSET = 5.0
    LARGE_SEARCH_LENGTH = 1.0

    def __init__(self, waypoints, goal_pose, car_model, obstacle_polys=[], default_search_length=1.5):
        self.default_search_length = default_search_length
        self.goal_pose = goal_pose
        self.car_model = car_model
        self.way_points = np.asarray(waypoints, dtype=np.float64)
        self.guided_path = self.get_guide_line(self.way_points)
        self.seg_xy = np.stack([self.way_points[:-1], self.way_points[1:]], axis=1)       # [S,2,2]
        self.seg_polys = np.array([capsule_polygon(a, b) for a, b in self.seg_xy])        # [S,66,2]
        self.crit_xy = lane_critical_points(self.seg_xy.reshape(-1, 4), self.seg_polys)
        if len(obstacle_polys) != 0:
            raise NotImplementedError("obstacle-dependent search lengths: no reference call site passes obstacles")
        self.search_lengths = self.create_segment_lengths(len(self.seg_xy))
        self._upload = None

    @staticmethod
    def get_guide_line(waypoints):
        """reference_line_heuristic.py:50-82 (polyline part)."""
        step = 0.1
        xs_all, ys_all, yaws_all = np.array([]), np.array([]), np.array([])
        for i in range(1, len(waypoints)):
            x_end, x_start = waypoints[i, 0], waypoints[i - 1, 0]
            y_end, y_start = waypoints[i, 1], waypoints[i - 1, 1]
            num = int(np.hypot(x_end - x_start, y_end - y_start) / step)
            xs = np.linspace(x_start, x_end, num)
            xs_all = np.append(xs_all, xs)
            ys_all = np.append(ys_all, np.linspace(y_start, y_end, num))
            yaws_all = np.append(yaws_all, np.ones_like(xs) * math.atan2(y_end - y_start, x_end - x_start))
        ss = np.zeros_like(xs_all)
        ss[1:] = np.cumsum(np.hypot(np.diff(xs_all), np.diff(ys_all)))
        return np.ascontiguousarray(np.array([xs_all, ys_all, yaws_all, ss]).T)

    def create_segment_lengths(self, n):
        """reference_line_heuristic.py:84-96 (obstacle-free branch)."""
        lengths = np.ones(n) * self.default_search_length
        if n > 4:
            lengths[2:n - 1] = self.LARGE_SEARCH_LENGTH
        return lengths

    def _env_batch(self, car_model):
        if self._upload is None or self._upload[0] != car_model.footprint_key():
            rec = EnvRecord(np.zeros((0, 4, 2)), np.zeros((0, 2)), car_model.body_ext, car_model.aux_exts,
                            seg_xy=self.seg_xy, seg_polys=self.seg_polys, seg_len=self.search_lengths,
                            crit_xy=self.crit_xy, guide=self.guided_path,
                            default_search_length=self.default_search_length)
            self._upload = (car_model.footprint_key(), EnvBatch([rec]))
        return self._upload[1]

    def check_path_feasibility(self, car_model, path):
        """reference_line_heuristic.py:105-118: swept body inside the lane."""
        path = np.asarray(path, dtype=np.float64)
        bad = ops.collision_check(self._env_batch(car_model), path[:, :3], flags=ops.CHECK_LANE)
        return not bool(bad.any().item())
