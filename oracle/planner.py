"""Float64 CPU restatement of the reference's warm-start search stack.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  The classes keep the
reference's names and call signatures so parity tests read like calls into the
reference; bodies follow the cited lines with shapely replaced by
``oracle.geometry`` and heapdict by ``oracle.heapdict_port`` (**parity unpinned**
at those two third-party boundaries; Reeds-Shepp is pinned, ``oracle.rs_port``).

  CarModel                     path_planner/car_model.py:9-37,75-162
  OrchardGeometryEnvironment   path_planner/orchard_geometry_environment.py:15-32,277-353,423-472
  ReferenceLineHeuristic       path_planner/reference_line_heuristic.py:13-158
  HybridAStarSearch            path_planner/hybrid_a_star_search.py:26-607
  create_tree_rows/get_base_pose  path_planner/utils/map_utils.py:45-61,228-271
"""
import math

import numpy as np

from . import geometry as geo
from . import rs_port as rs_curves
from .heapdict_port import HeapDict


def angle_wrap(angles):
    """path_planner/utils/path_utils.py:26-29."""
    return (angles + math.pi) % (2 * math.pi) - math.pi


def calculate_path_length(xs, ys):
    """path_planner/utils/path_utils.py:5-12."""
    ds = np.hypot(np.diff(xs), np.diff(ys))
    return np.cumsum(ds)[-1]


# ------------------------------------------------------------------ map utils
NEAR_SIDE = 1
FAR_SIDE = 2
LEAVE_POSE = 1
ENTER_POSE = 2


def create_tree_rows(row_num, row_width, row_lengths, slope_angle=0, l_std=0.0):
    """path_planner/utils/map_utils.py:45-61 (consumes np.random)."""
    tree_rows = []
    delta_x = row_width * np.tan(slope_angle)
    for i in range(row_num):
        if isinstance(row_lengths, (list, np.ndarray)):
            row_length = row_lengths[i]
        else:
            row_length = row_lengths
        y = row_width * i
        x = delta_x * i
        x += np.random.uniform(-l_std, l_std)
        tree_rows.append(np.array([[x, y], [x + row_length, y]]))
    return np.array(tree_rows)


def get_base_pose(row_id, map_tree_rows, min_offset, side=NEAR_SIDE, pose_type=LEAVE_POSE):
    """path_planner/utils/map_utils.py:228-271."""
    near_side_end = [
        (map_tree_rows[row_id, 0, 0] + map_tree_rows[row_id + 1, 0, 0]) / 2,
        (map_tree_rows[row_id, 0, 1] + map_tree_rows[row_id + 1, 0, 1]) / 2,
    ]
    far_side_end = [
        (map_tree_rows[row_id, 1, 0] + map_tree_rows[row_id + 1, 1, 0]) / 2,
        (map_tree_rows[row_id, 1, 1] + map_tree_rows[row_id + 1, 1, 1]) / 2,
    ]
    row_yaw = np.arctan2(far_side_end[1] - near_side_end[1], far_side_end[0] - near_side_end[0])
    if pose_type == LEAVE_POSE:
        pose_yaw = row_yaw if side == FAR_SIDE else row_yaw + np.pi
        extend_dir = 1
    else:
        pose_yaw = row_yaw + np.pi if side == FAR_SIDE else row_yaw
        extend_dir = -1
    end = near_side_end if side == NEAR_SIDE else far_side_end
    pose_position = end + np.array([np.cos(pose_yaw), np.sin(pose_yaw)]) * min_offset * extend_dir
    return np.array([pose_position[0], pose_position[1], pose_yaw])


# ------------------------------------------------------------------ car model
class _Bounds:
    def __init__(self, ext):
        self.bounds = (ext[0], ext[2], ext[1], ext[3])


class CarModel:
    def __init__(self, max_steer=0.55, wheel_base=1.9, axle_to_front=2.85, axle_to_back=0.5,
                 width=1.48, head_out=0.542, head_side=0.44, body_vertices=[],
                 aux_poly_features=[], with_aux=False):
        self.MAX_STEER = max_steer
        self.WHEEL_BASE = wheel_base
        self.AXLE_TO_FRONT = axle_to_front
        self.AXLE_TO_BACK = axle_to_back
        self.WIDTH = width
        self.curvature = math.tan(self.MAX_STEER) / self.WHEEL_BASE      # car_model.py:34
        self.with_aux = with_aux
        self.body_ext = geo.body_extent(axle_to_back, axle_to_front, width)
        self.aux_exts = [geo.aux_extent(f) for f in aux_poly_features] if with_aux else []
        # what the orchestration code reads from the shapely polygons: .bounds = (minx, miny, maxx, maxy)
        self.car_poly = _Bounds(self.body_ext)
        self.aux_polys = [_Bounds(e) for e in self.aux_exts]

    def calculate_motion_path_new(self, init_pose, motion_dir, steer_dir, turning_radius, delta_yaw, step_size=0.1):
        """car_model.py:236-269."""
        turning_radius = max(1.0 / self.curvature, turning_radius)
        steer_angle = math.atan(self.WHEEL_BASE / turning_radius) * steer_dir
        arc_length = abs(delta_yaw * turning_radius)
        num_steps = int(arc_length / step_size)
        actual_step_size = arc_length / num_steps
        yaw_step = motion_dir * actual_step_size / self.WHEEL_BASE * math.tan(steer_angle)
        init_x, init_y = init_pose[0], init_pose[1]
        init_yaw = angle_wrap(init_pose[-1])
        yaws = angle_wrap(np.linspace(init_yaw, init_yaw + yaw_step * num_steps, num_steps + 1))
        xs = init_x + turning_radius * (np.sin(yaws) - np.sin(init_yaw)) * steer_dir
        ys = init_y - turning_radius * (np.cos(yaws) - np.cos(init_yaw)) * steer_dir
        path = np.vstack([init_pose, np.vstack([xs, ys, yaws]).T])
        curvature = math.tan(steer_angle) / self.WHEEL_BASE if abs(steer_angle) > 0.00001 else 0
        return np.hstack((path, np.ones((len(path), 1)) * curvature, np.ones((len(path), 1)) * motion_dir))

    def get_turn_radius(self, max_steer_angle=None):
        if max_steer_angle is None:
            return 1 / self.curvature
        return self.WHEEL_BASE / math.tan(max_steer_angle)


# ---------------------------------------------------------------- environment
class OrchardGeometryEnvironment:
    NEAR_SIDE = 1
    FAR_SIDE = -1

    def __init__(self, map_tree_rows, obstacles, contour_points=[], tree_width=0.2,
                 headland_width=7, obstacle_dim=0.3):
        self.map_tree_rows = np.asarray(map_tree_rows, dtype=np.float64)
        self.tree_width = tree_width
        self.headland_width = headland_width
        self.tree_polys = [geo.line_flat_buffer(r[0], r[1], tree_width / 2) for r in self.map_tree_rows]
        self.obstacle_polys = [geo.point_square_buffer(o[0], o[1], obstacle_dim) for o in obstacles]
        if len(contour_points) == 0:
            self.field_poly = self.get_map_exterior_pts(headland_width)
        else:
            self.field_poly = np.asarray(contour_points, dtype=np.float64)
        if len(self.field_poly) > 1 and (self.field_poly[0] == self.field_poly[-1]).all():
            self.field_poly = self.field_poly[:-1]
        # orchard_geometry_environment.py:31 (STRtree is built once; update_tree_width leaves it stale)
        self.obs_poly_list = self.obstacle_polys + self.tree_polys

    def get_headland_angle(self, side):
        """orchard_geometry_environment.py:463-472."""
        side_idx = 0 if side == self.NEAR_SIDE else 1
        xs = self.map_tree_rows[:, side_idx, 0]
        if np.std(xs) < 0.01:
            return np.pi / 2
        ys = self.map_tree_rows[:, side_idx, 1]
        k = np.polyfit(xs, ys, deg=1)[0]
        return math.atan(k)

    def get_map_exterior_pts(self, headland_width):
        """orchard_geometry_environment.py:288-334."""
        rows = self.map_tree_rows
        row_width = np.mean(np.diff(rows[:, 0, 1]))
        near_angle = self.get_headland_angle(self.NEAR_SIDE)
        far_angle = self.get_headland_angle(self.FAR_SIDE)
        delta_x_near = abs(headland_width / math.sin(near_angle))
        delta_x_far = abs(headland_width / math.sin(far_angle))
        near = np.array([r[0] for r in rows])
        near[:, 0] -= delta_x_near
        up = np.argmax(near[:, 1])
        near[up, 1] += row_width
        delta_x = 0 if np.abs(np.sin(near_angle)) < 1e-5 else row_width / np.tan(near_angle)
        near[up, 0] += delta_x
        low = np.argmin(near[:, 1])
        near[low, 1] -= row_width
        near[low, 0] -= delta_x
        far = np.array([r[1] for r in rows][::-1])
        far[:, 0] += delta_x_far
        up = np.argmax(far[:, 1])
        delta_x = 0 if np.abs(np.sin(far_angle)) < 1e-5 else row_width / np.tan(far_angle)
        far[up, 1] += row_width
        far[up, 0] += delta_x
        low = np.argmin(far[:, 1])
        far[low, 1] -= row_width
        far[low, 0] -= delta_x
        return np.concatenate((near, far))

    def pose_flags(self, car_model, path, boundary_check=True, aux_check=False):
        """Per-pose infeasibility flags (P,) -- the per-pose decomposition of
        ``check_path_feasibility``; aux rectangles only count at poses 0,2,4,..
        (car_model.py:58)."""
        path = np.asarray(path, dtype=np.float64)
        poses = path[:, :3]
        bad = self._part_flags(poses, car_model.body_ext, boundary_check)
        if aux_check:
            for ext in car_model.aux_exts:
                sub = self._part_flags(poses[0::2], ext, boundary_check)
                bad[0::2] |= sub
        return bad

    def _part_flags(self, poses, ext, boundary_check):
        corners = geo.rect_corners(poses, ext)
        bad = np.zeros(len(poses), dtype=bool)
        for poly in self.obs_poly_list:
            bad |= geo.rects_hit_convex(poses, ext, poly, corners)
        if boundary_check:
            bad |= ~geo.rects_inside_polygon(poses, ext, self.field_poly, corners)
        return bad

    def check_path_feasibility(self, car_model, path, boundary_check=True, aux_check=False):
        """orchard_geometry_environment.py:423-458."""
        return not self.pose_flags(car_model, path, boundary_check, aux_check).any()

    def get_min_distance_to_boundary(self, car_model, path, with_aux=True):
        """orchard_geometry_environment.py:393-412 (the union's vertex set is oracle.geometry.union_boundary_vertices)."""
        path = np.asarray(path, dtype=np.float64)
        pts = [geo.union_boundary_vertices(path[:, :3], car_model.body_ext)]
        if with_aux:
            for ext in car_model.aux_exts:
                pts.append(geo.union_boundary_vertices(path[0:len(path):2, :3], ext))
        d = geo.signed_distance_to_ring(np.concatenate(pts), geo.ccw(self.field_poly))
        return float(d.min())

    def get_row_ids_between_start_and_end(self, start_pose, end_pose):
        """orchard_geometry_environment.py:93-127."""
        near_xs = self.map_tree_rows[:, 0, 0]
        near_ys = self.map_tree_rows[:, 0, 1]
        far_xs = self.map_tree_rows[:, 1, 0]
        far_ys = self.map_tree_rows[:, 1, 1]
        row_id = np.argmin(np.abs(start_pose[1] - near_ys))
        is_near = abs(start_pose[0] - near_xs[row_id]) < abs(start_pose[0] - far_xs[row_id])
        ys = np.copy(near_ys) if is_near else np.copy(far_ys)
        xs = np.copy(near_xs) if is_near else np.copy(far_xs)
        sy, ey = start_pose[1], end_pose[1]
        if sy > ey:
            idx = np.where((ys > ey) & (ys < sy))[0]
        else:
            idx = np.where((ys > sy) & (ys < ey))[0]
        return xs, ys, idx

    def check_side_of_a_point(self, point):
        """orchard_geometry_environment.py:49-64 (consumes np.random, Appendix-A quirk 9)."""
        row_centers = np.mean(self.map_tree_rows[:, :, :], axis=1)
        epsilon = np.random.uniform(-0.5, 0.5, size=(len(row_centers),))
        k, b = np.polyfit(row_centers[:, 0] + epsilon, row_centers[:, 1], deg=1)
        origin_sign = np.sign(0 * k + b - 0)
        judge = np.sign(point[0] * k + b - point[1])
        return self.NEAR_SIDE if origin_sign == judge else self.FAR_SIDE

    def get_topology_waypoints(self, start_pose, end_pose, drive_row_offset):
        """orchard_geometry_environment.py:199-248."""
        xs, ys, idx = self.get_row_ids_between_start_and_end(start_pose, end_pose)
        iys = ys[idx]
        ixs = xs[idx]
        order = np.argsort(np.abs(iys - start_pose[1]))
        iys = iys[order]
        ixs = ixs[order]
        side = self.check_side_of_a_point(start_pose[:2])
        offset = -drive_row_offset if side == self.NEAR_SIDE else drive_row_offset
        contour = np.vstack((ixs + offset, iys)).T
        return np.vstack((start_pose[:2], contour, end_pose[:2]))


# ------------------------------------------------------------------ heuristic
class ReferenceLineHeuristic:
    ACCEPT_PATH_DEVIATION = 2
    DRIVE_ROW_OFFSET = 5.0
    LARGE_SEARCH_LENGTH = 1.0

    def __init__(self, waypoints, goal_pose, car_model, obstacle_polys=[], default_search_length=1.5):
        self.default_search_length = default_search_length
        self.goal_pose = goal_pose
        self.car_model = car_model
        self.way_points = np.asarray(waypoints, dtype=np.float64)
        self.guided_path = self.get_guide_line(self.way_points)
        self.lane = geo.Lane(self.way_points)
        n = len(self.way_points) - 1
        lengths = np.ones(n) * default_search_length             # reference_line_heuristic.py:84-96
        if n > 4:
            assert len(obstacle_polys) == 0, "call sites never pass obstacles (SURVEY 8a-11)"
            lengths[2:n - 1] = self.LARGE_SEARCH_LENGTH
        self.search_lengths = lengths

    @staticmethod
    def get_guide_line(waypoints):
        """reference_line_heuristic.py:50-82 (guide polyline only)."""
        step = 0.1
        way_xs, way_ys, way_yaws = np.array([]), np.array([]), np.array([])
        for i in range(1, len(waypoints)):
            x_end, x_start = waypoints[i, 0], waypoints[i - 1, 0]
            y_end, y_start = waypoints[i, 1], waypoints[i - 1, 1]
            dist = np.hypot(x_end - x_start, y_end - y_start)
            num = int(dist / step)
            xs = np.linspace(x_start, x_end, num)
            ys = np.linspace(y_start, y_end, num)
            way_xs = np.append(way_xs, xs)
            way_ys = np.append(way_ys, ys)
            yaw = math.atan2(y_end - y_start, x_end - x_start)
            way_yaws = np.append(way_yaws, np.ones_like(xs) * yaw)
        delta_ss = np.hypot(np.diff(way_xs), np.diff(way_ys))
        way_ss = np.zeros_like(way_xs)
        way_ss[1:] = np.cumsum(delta_ss)
        return np.array([way_xs, way_ys, way_yaws, way_ss]).T

    def check_path_feasibility(self, car_model, path):
        """reference_line_heuristic.py:105-118 (body only)."""
        path = np.asarray(path, dtype=np.float64)
        return bool(self.lane.rects_inside(path[:, :3], car_model.body_ext).all())

    def pose_flags(self, car_model, path):
        path = np.asarray(path, dtype=np.float64)
        return ~self.lane.rects_inside(path[:, :3], car_model.body_ext)

    def get_search_length(self, pose):
        """reference_line_heuristic.py:120-129: LAST containing segment wins."""
        i = self.lane.search_segment(pose[0], pose[1])
        return self.default_search_length if i < 0 else self.search_lengths[i]

    def calculate_state_cost(self, pose):
        """reference_line_heuristic.py:131-158."""
        dists = np.hypot(self.guided_path[:, 0] - pose[0], self.guided_path[:, 1] - pose[1])
        match_idx = np.argmin(dists)
        match_pose = self.guided_path[match_idx]
        distance_to_path = dists[match_idx] * 100
        yaw_difference = abs(angle_wrap(match_pose[2] - pose[2]))
        if distance_to_path > self.ACCEPT_PATH_DEVIATION:
            distance_to_path = 100
        dist_to_goal = self.guided_path[-1, -1] - self.guided_path[match_idx, -1]
        return distance_to_path + yaw_difference * 0.2 + dist_to_goal * 5


# --------------------------------------------------------------- hybrid A star
class Node:
    """hybrid_a_star_search.py:13-23."""
    __slots__ = ("grid_index", "traj", "curvature", "cost", "parent_index", "direction")

    def __init__(self, grid_index, traj, curvature, cost, direction, parent_index):
        self.grid_index = grid_index
        self.traj = traj
        self.curvature = curvature
        self.cost = cost
        self.parent_index = parent_index
        self.direction = direction

    def get_hybrid_index(self):
        return tuple([self.grid_index[0], self.grid_index[1], self.grid_index[2]])


class HybridAStarSearch:
    STEER_COST = 1
    DELTA_STEER_COST = 5
    DEVIATION_COST = 1
    DISTANCE_COST = 1
    DIRECTION_CHANGE_COST = 1000
    REVERSE_COST = 5000
    HYBRID_COST = 50
    MIN_LENGTH_TO_GOAL = 1000

    def __init__(self, start_pose, goal_pose, config_environment, car_model, search_heuristic,
                 motion_type="Pawn", yaw_resolution=math.radians(10), plan_resolution=0.1):
        self.plan_resolution = plan_resolution
        self.yaw_resolution = yaw_resolution
        self.config_env = config_environment
        self.car_model = car_model
        self.search_heuristic = search_heuristic
        self.motion_type = motion_type
        if motion_type == "King":
            self.motion_steers = self._get_motion_steers_reeds_shepp()
        elif motion_type == "Pawn":
            # Dubins goal extension: pydubins is un-vendored -> oracle.dubins_port restates dubins.c (parity unpinned)
            self.motion_steers = self._get_motion_steers_dubins()
        else:
            raise ValueError(f"unknown motion_type {motion_type!r}")
        self.start_node = self.init_node(start_pose)
        self.goal_node = self.init_node(goal_pose)
        # audit trail for parity tests (not in the reference)
        self.expanded = []          # popped keys in pop order
        self.stats = {"primitive_poses": 0, "rs_poses": 0, "rs_words": 0, "pushes": 0}

    def calculate_node_index(self, x, y, yaw):
        return (round(x / self.plan_resolution), round(y / self.plan_resolution),
                round(yaw / self.yaw_resolution))

    def init_node(self, pose):
        x, y, yaw = pose[0], pose[1], pose[2]
        idx = self.calculate_node_index(x, y, yaw)
        return Node(idx, [[x, y, yaw]], [0], 0, [1], idx)

    def _get_motion_steers_dubins(self):
        """hybrid_a_star_search.py:331-341: 8 steers 0.55 .. -0.67173, all forward."""
        steer_ranges = np.arange(self.car_model.MAX_STEER, -(self.car_model.MAX_STEER + self.yaw_resolution),
                                 -self.yaw_resolution)
        return np.vstack((steer_ranges, np.ones_like(steer_ranges))).T

    def get_dubins_path(self, start_x, start_y, start_yaw, goal_x, goal_y, goal_yaw, curvature):
        """hybrid_a_star_search.py:289-304: shortest Dubins path sampled at plan_resolution, the goal pose appended,
        then a cubic-spline course (scipy not-a-knot) resampled at plan_resolution -> rows (x, y, yaw, k)."""
        from . import dubins_port as dubins
        from .obca_util import calc_spline_course
        path = dubins.shortest_path([start_x, start_y, start_yaw], [goal_x, goal_y, goal_yaw], 1.0 / curvature)
        dubins_path, _ = path.sample_many(self.plan_resolution)
        dubins_path = np.vstack([np.array(dubins_path), np.array([[goal_x, goal_y, goal_yaw]])])
        xs, ys, yaws, ks, _ = calc_spline_course(dubins_path[:, 0], dubins_path[:, 1], ds=self.plan_resolution)
        return np.array([xs, ys, yaws, ks]).T

    def calculate_dubins_path_cost(self, current_node, path):
        """hybrid_a_star_search.py:162-182.  [Q] ``path[:, -1]`` is the CURVATURE column of the 4-column array the
        caller passes, so the "steer cost" is the wrapped spread of the curvatures; the method returns a tuple and
        the caller stores that tuple as the goal node's cost (:203, :219-226) -- harmless, the search ends there."""
        cost = current_node.cost
        path_length = calculate_path_length(path[:, 0], path[:, 1])
        cost += path_length * self.DISTANCE_COST
        delta_yaw = angle_wrap(np.max(path[:, -1]) - np.min(path[:, -1]))
        cost += delta_yaw * self.STEER_COST
        return cost, path_length

    def _get_goal_extension_with_dubins_path(self, current_node):
        """hybrid_a_star_search.py:184-230."""
        sx, sy, syaw = current_node.traj[-1][0], current_node.traj[-1][1], current_node.traj[-1][2]
        gx, gy, gyaw = self.goal_node.traj[-1][0], self.goal_node.traj[-1][1], self.goal_node.traj[-1][2]
        dubin_path = self.get_dubins_path(sx, sy, syaw, gx, gy, gyaw, self.car_model.curvature)
        cost = self.calculate_dubins_path_cost(current_node, dubin_path)
        traj = np.copy(dubin_path[:, :3])
        traj[:, -1] = angle_wrap(traj[:, -1])
        path_length = calculate_path_length(dubin_path[:, 0], dubin_path[:, 1])
        ks = list(dubin_path[:, 3])
        self.stats["rs_words"] += 1
        self.stats["rs_poses"] += len(traj)
        if not self.check_collision(traj) and path_length < self.MIN_LENGTH_TO_GOAL:
            return Node(self.goal_node.get_hybrid_index(), traj, ks, cost, np.ones_like(ks).tolist(),
                        current_node.get_hybrid_index())
        return None

    def get_goal_extension_path(self, current_node):
        """hybrid_a_star_search.py:456-462."""
        if self.motion_type == "King":
            return self._get_goal_extension_with_reeds_shepp_path(current_node)
        return self._get_goal_extension_with_dubins_path(current_node)

    def _get_motion_steers_reeds_shepp(self):
        """hybrid_a_star_search.py:343-354."""
        steer_ranges = np.arange(self.car_model.MAX_STEER,
                                 -(self.car_model.MAX_STEER + self.yaw_resolution / 2.0),
                                 -self.yaw_resolution / 2.0)
        directions = np.ones_like(steer_ranges)
        directions[1:len(directions):2] = -1
        return np.vstack((steer_ranges, directions)).T

    def calculate_reeds_shepp_path_cost(self, current_node, path):
        """hybrid_a_star_search.py:129-160 (quirks kept: ``len(np.where(..))`` is the
        tuple length 1; 'L' compares with "WB" so left arcs count as steer 0)."""
        cost = current_node.cost
        path_lengths = np.array(path.lengths)
        idxs = np.where(path_lengths < 0)[0]
        other_move_cost = len(path_lengths) - len(idxs)
        cost += self.REVERSE_COST * len(idxs) + other_move_cost
        direction_changes = np.array(path_lengths[:-1]) * np.array(path_lengths[1:])
        idxs = np.where(direction_changes < 0)
        cost += len(idxs) * self.DIRECTION_CHANGE_COST
        path_types = np.array(path.ctypes)
        idxs = np.where(path_types != "S")
        cost += self.car_model.MAX_STEER * self.STEER_COST * len(idxs)
        steers = np.zeros(len(path_types))
        steers[np.where(path_types == "R")[0]] = -self.car_model.MAX_STEER
        steers[np.where(path_types == "WB")[0]] = self.car_model.MAX_STEER
        cost += np.sum(np.abs(np.diff(steers)))
        return cost

    def rs_candidates_in_pop_order(self, current_node):
        """The heapdict pop order of the Reeds-Shepp candidates
        (hybrid_a_star_search.py:249-271)."""
        sx, sy, syaw = current_node.traj[-1][0], current_node.traj[-1][1], current_node.traj[-1][2]
        gx, gy, gyaw = self.goal_node.traj[-1][0], self.goal_node.traj[-1][1], self.goal_node.traj[-1][2]
        paths = rs_curves.calc_all_paths(sx, sy, syaw, gx, gy, gyaw,
                                         self.car_model.curvature, self.plan_resolution)
        q = HeapDict()
        for i, p in enumerate(paths):
            q[i] = self.calculate_reeds_shepp_path_cost(current_node, p)
        order = []
        while len(q) != 0:
            i, c = q.popitem()
            order.append((paths[i], c))
        return order

    def _get_goal_extension_with_reeds_shepp_path(self, current_node):
        """hybrid_a_star_search.py:232-287."""
        for path, path_cost in self.rs_candidates_in_pop_order(current_node):
            traj = np.array([path.x, path.y, path.yaw]).T
            self.stats["rs_words"] += 1
            self.stats["rs_poses"] += len(traj)
            if not self.check_collision(traj) and path.L < self.MIN_LENGTH_TO_GOAL:
                return Node(self.goal_node.get_hybrid_index(), traj, path.cs, path_cost,
                            path.directions, current_node.get_hybrid_index())
        return None

    def simulated_path_cost(self, current_node, traj, motion_command):
        """hybrid_a_star_search.py:306-329."""
        cost = current_node.cost
        cost += calculate_path_length(traj[:, 0], traj[:, 1])
        if motion_command[1] == -1:
            cost += self.REVERSE_COST
        cost += motion_command[0] * self.STEER_COST
        steer_angle = math.atan(current_node.curvature[0] * self.car_model.WHEEL_BASE)
        cost += abs(motion_command[0] - steer_angle) * self.DELTA_STEER_COST
        if current_node.direction[0] != motion_command[1]:
            cost += self.DIRECTION_CHANGE_COST
        return cost

    def rollout(self, end_pose, motion_command):
        """Kinematic part of ``kinematic_simulation_node`` (hybrid_a_star_search.py:366-394)."""
        steer_angle = motion_command[0]
        speed_direction = motion_command[1]
        search_length = self.search_heuristic.get_search_length(end_pose)
        num_steps = round(search_length / self.plan_resolution)
        yaw_step = (speed_direction * self.plan_resolution / self.car_model.WHEEL_BASE
                    * math.tan(steer_angle))
        init_yaw = angle_wrap(end_pose[2] + yaw_step)
        yaws = np.linspace(init_yaw, init_yaw + yaw_step * (num_steps + 1), num_steps + 2)
        yaws = angle_wrap(yaws)
        xs = self.plan_resolution * np.cos(yaws[:-1]) * speed_direction
        xs = end_pose[0] + np.cumsum(xs)
        ys = self.plan_resolution * np.sin(yaws[:-1]) * speed_direction
        ys = end_pose[1] + np.cumsum(ys)
        traj = np.vstack([xs, ys, yaws[1:]]).T
        grid_index = self.calculate_node_index(traj[-1][0], traj[-1][1], traj[-1][2])
        return traj, grid_index

    def kinematic_simulation_node(self, current_node, motion_command):
        """hybrid_a_star_search.py:357-410."""
        traj, grid_index = self.rollout(current_node.traj[-1], motion_command)
        self.stats["primitive_poses"] += len(traj)
        if self.check_collision(traj):
            return None
        cost = self.simulated_path_cost(current_node, traj, motion_command)
        curvature = np.tan(motion_command[0]) / self.car_model.WHEEL_BASE
        return Node(grid_index, traj, [curvature] * len(traj), cost,
                    [motion_command[1]] * len(traj), current_node.get_hybrid_index())

    def check_collision(self, path):
        """hybrid_a_star_search.py:412-427."""
        feasible = self.config_env.check_path_feasibility(self.car_model, path)
        in_range = self.search_heuristic.check_path_feasibility(self.car_model, path)
        return (not feasible) or (not in_range)

    def get_heuristic_cost(self, pose):
        return self.search_heuristic.calculate_state_cost(pose)

    def check_the_arrival(self, goal_extension_node, current_node):
        """hybrid_a_star_search.py:464-495."""
        goal_node = goal_extension_node
        x_dist = np.abs(current_node.traj[-1][0] - self.goal_node.traj[0][0])
        y_dist = np.abs(current_node.traj[-1][1] - self.goal_node.traj[0][1])
        yaw_diff = np.abs(angle_wrap(current_node.traj[-1][2] - self.goal_node.traj[0][2]))
        if x_dist < self.plan_resolution and y_dist < self.plan_resolution and yaw_diff < self.yaw_resolution:
            goal_node = current_node
            goal_node.grid_index = self.goal_node.grid_index
        return goal_node

    def get_path_from_expanded_nodes(self, closed_set):
        """hybrid_a_star_search.py:429-454."""
        start_idx = self.start_node.get_hybrid_index()
        cur_idx = self.goal_node.parent_index
        if cur_idx not in closed_set:
            return [], [], [], [], []
        cur = closed_set[cur_idx]
        xs, ys, yaws, dirs, ks = [], [], [], [], []
        while cur_idx != start_idx:
            a, b, c = zip(*cur.traj)
            xs += a[::-1]
            ys += b[::-1]
            yaws += c[::-1]
            dirs += list(cur.direction)[::-1]
            ks += list(cur.curvature)[::-1]
            cur_idx = cur.parent_index
            cur = closed_set[cur_idx]
        return xs[::-1], ys[::-1], yaws[::-1], ks[::-1], dirs[::-1]

    def hybrid_a_star_search(self, plt=None, max_nodes=2000):
        """hybrid_a_star_search.py:497-607."""
        open_set = {self.start_node.get_hybrid_index(): self.start_node}
        closed_set = {}
        cost_queue = HeapDict()
        cost_queue[self.start_node.get_hybrid_index()] = max(
            self.start_node.cost, self.HYBRID_COST * self.get_heuristic_cost(self.start_node.traj[-1]))
        counter = 0
        self.expanded = []
        self.status = "none"
        if self.check_collision(self.start_node.traj) or self.check_collision(self.goal_node.traj):
            self.status = "start_goal_blocked"
            return [], [], [], [], [], 0
        while True:
            if counter > max_nodes:
                self.status = "max_nodes"
                break
            counter += 1
            if not open_set:
                self.status = "open_empty"
                break
            current_idx, _ = cost_queue.popitem()
            current_node = open_set.pop(current_idx)
            closed_set[current_idx] = current_node
            self.expanded.append(current_idx)
            goal_ext = self.get_goal_extension_path(current_node)
            goal_node = self.check_the_arrival(goal_ext, current_node)
            if goal_node is not None:
                closed_set[goal_node.get_hybrid_index()] = goal_node
                self.status = "ok"
                break
            for i in range(len(self.motion_steers)):
                sim = self.kinematic_simulation_node(current_node, self.motion_steers[i])
                if not sim:
                    continue
                idx = sim.get_hybrid_index()
                if idx not in closed_set:
                    if idx not in open_set:
                        open_set[idx] = sim
                        cost_queue[idx] = max(sim.cost, self.HYBRID_COST * self.get_heuristic_cost(sim.traj[-1]))
                        self.stats["pushes"] += 1
                    elif sim.cost < open_set[idx].cost:
                        open_set[idx] = sim
                        cost_queue[idx] = max(sim.cost, self.HYBRID_COST * self.get_heuristic_cost(sim.traj[-1]))
                        self.stats["pushes"] += 1
        x, y, yaw, ks, dirs = self.get_path_from_expanded_nodes(closed_set)
        return (x, y, yaw, dirs, ks, counter)


# ---------------------------------------------------------------------------------------------------------
# the Y-type parking parameter sweep that runs right before every Hybrid A*
# call and produces its goal pose (path_planner/headland_path_planning.py:352-527).  Pinned against the
# reference's OWN functions run in the build container (oracle/ref_loader.load_planner, tests/golden/
# ypark_golden.npz): only ``check_path_feasibility`` is the restated geometry.

def get_backward_steer_dir_for_y_type_parking(start_pose, end_pose):
    """headland_path_planning.py:360-368."""
    if end_pose[1] - start_pose[1] > 0:
        return np.sign(1 * math.cos(start_pose[2]))
    return np.sign(-1 * math.cos(start_pose[2]))


def calculate_motion_path(init_pose, motion_command, search_length, wheel_base, step):
    """headland_path_planning.py:455-485: [P+1, 5] rows (x, y, yaw, curvature, direction)."""
    steer_angle, speed_direction = motion_command[0], motion_command[1]
    num_steps = round(search_length / step)
    yaw_step = speed_direction * step / wheel_base * math.tan(steer_angle)
    init_yaw = angle_wrap(init_pose[-1] + yaw_step)
    yaws = angle_wrap(np.linspace(init_yaw, init_yaw + yaw_step * num_steps, num_steps + 1))
    xs = init_pose[0] + np.cumsum(step * np.cos(yaws[:-1]) * speed_direction)
    ys = init_pose[1] + np.cumsum(step * np.sin(yaws[:-1]) * speed_direction)
    path = np.vstack([init_pose, np.vstack([xs, ys, yaws[1:]]).T])
    curvature = math.tan(steer_angle) / wheel_base if abs(steer_angle) > 0.00001 else 0
    return np.hstack((path, np.ones((len(path), 1)) * curvature, np.ones((len(path), 1)) * speed_direction))


def get_y_type_parking_path(car_model, backward_length, backward_steer, forward_length, forward_steer, step):
    """headland_path_planning.py:488-516: planned inversely from the end pose, in the base-link frame."""
    back_path = calculate_motion_path([0, 0, 0], [backward_steer, -1], backward_length, car_model.WHEEL_BASE, step)
    forward_path = calculate_motion_path(back_path[-1, :3], [forward_steer, 1], forward_length,
                                         car_model.WHEEL_BASE, step)
    back_path[:, -1] = 1
    back_path = back_path[::-1]
    forward_path[:, -1] = -1
    forward_path = forward_path[::-1]
    return np.vstack([forward_path, back_path])


def get_path_in_odom(end_pose, path):
    """headland_path_planning.py:519-527 with utils/transformation.py:7-61 and
    navigation_utils.py:196-203 restated for a planar pose: the yaw offset is atan2(sin, cos) of the end
    yaw (rotationMatrixToEulerAngles), positions go through the homogeneous 4x4 product (row sums in the
    order x*R00 + y*R01 + 0*R02 + 1*tx; BLAS may fuse/reorder, hence the 1e-12 tolerance of the pin)."""
    c, s = math.cos(end_pose[2]), math.sin(end_pose[2])
    out = np.copy(path)
    out[:, 2] += math.atan2(s, c)
    x, y = path[:, 0], path[:, 1]
    out[:, 0] = c * x + (-s) * y + end_pose[0]
    out[:, 1] = s * x + c * y + end_pose[1]
    return out


def y_park_candidates(backward_steer_dir, forward_steer_dir, max_steer_backward=0.4, max_steer_forward=0.45,
                      max_backward_distance=3.5, max_forward_distance=2.0, min_forward_distance=1.4,
                      min_backward_distance=0.7, min_steer_backward=0.3, min_steer_forward=0.3):
    """The 4-deep candidate enumeration of search_y_type_parking_path (:405-420) in loop order:
    rows (backward_length, forward_length, steer_backward, steer_forward), unsigned steers."""
    steer_backwards = list(np.arange(min_steer_backward, max_steer_backward + 0.1, 0.1))
    if np.max(steer_backwards) < max_steer_backward:
        steer_backwards.append(max_steer_backward)
    steer_forwards = list(np.arange(min_steer_forward, max_steer_forward + 0.1, 0.1))
    if np.max(steer_forwards) < max_steer_forward:
        steer_forwards.append(max_steer_forward)
    out = []
    for backward_length in np.arange(max_backward_distance, min_backward_distance, -0.1):
        for forward_length in np.arange(max_forward_distance, min_forward_distance, -0.1):
            for steer_backward in steer_backwards:
                for steer_forward in steer_forwards:
                    out.append((backward_length, forward_length, steer_backward, steer_forward))
    return np.array(out, dtype=np.float64).reshape(-1, 4)


def search_y_type_parking_path(car_model, config_env, end_pose, backward_steer_dir, forward_steer_dir,
                               max_steer_backward=0.4, max_steer_forward=0.45, max_backward_distance=3.5,
                               max_forward_distance=2.0, min_forward_distance=1.4, min_backward_distance=0.7,
                               min_steer_backward=0.3, min_steer_forward=0.3, step_size=0.1, debug=False):
    """headland_path_planning.py:382-451: the first feasible candidate in loop order wins."""
    if not config_env.check_path_feasibility(car_model, np.array([end_pose])):
        return [], []                      # :400-402: a tuple whatever `debug` says
    cands = y_park_candidates(backward_steer_dir, forward_steer_dir, max_steer_backward, max_steer_forward,
                              max_backward_distance, max_forward_distance, min_forward_distance,
                              min_backward_distance, min_steer_backward, min_steer_forward)
    for bl, fl, sb, sf in cands:
        local = get_y_type_parking_path(car_model, bl, sb * backward_steer_dir, fl, sf * forward_steer_dir, step_size)
        path = get_path_in_odom(end_pose, local)
        if config_env.check_path_feasibility(car_model, path):
            return (path, [bl, fl, sb, sf]) if debug else path
    return ([], []) if debug else []


# ---------------------------------------------------------------------------------------------------------
# SURVEY.md section 8(f) rank 2 (first part): the offset-pose sweep of safety_forward_path_plan.py:248-283.
# Pinned against the reference's OWN get_offset_pose + CarModel.calculate_motion_path
# (oracle/ref_loader.load_planner, tests/golden/offset_golden.npz).

def car_calculate_motion_path(car, init_pose, motion_command, delta_yaw, step):
    """CarModel.calculate_motion_path, car_model.py:202-234 (note the linspace end yaw_step*(num_steps+1))."""
    steer_angle, speed_direction = motion_command[0], motion_command[1]
    search_length = delta_yaw / car.curvature
    num_steps = round(search_length / step)
    yaw_step = speed_direction * step / car.WHEEL_BASE * math.tan(steer_angle)
    init_yaw = angle_wrap(init_pose[-1] + yaw_step)
    yaws = angle_wrap(np.linspace(init_yaw, init_yaw + yaw_step * (num_steps + 1), num_steps + 1))
    xs = init_pose[0] + np.cumsum(step * np.cos(yaws[:-1]) * speed_direction)
    ys = init_pose[1] + np.cumsum(step * np.sin(yaws[:-1]) * speed_direction)
    path = np.vstack([init_pose, np.vstack([xs, ys, yaws[1:]]).T])
    curvature = math.tan(steer_angle) / car.WHEEL_BASE if abs(steer_angle) > 0.00001 else 0
    return np.hstack((path, np.ones((len(path), 1)) * curvature, np.ones((len(path), 1)) * speed_direction))


def get_offset_pose(init_pose, pose_type, turn_out_dir, car, config_env, steer_angle=0.55,
                    delta_yaw=math.radians(45), max_offset=5, accuracy=0.1):
    """safety_forward_path_plan.py:248-283."""
    init_x, init_y, init_yaw = init_pose[0], init_pose[1], init_pose[2]
    motion_dir = -1 if pose_type == ENTER_POSE else 1
    offset_dir = -1 if pose_type == ENTER_POSE else 1
    for dist in np.arange(0, max_offset + accuracy, accuracy):
        x = init_x + dist * np.cos(init_yaw) * offset_dir
        y = init_y + dist * np.sin(init_yaw) * offset_dir
        pose = np.array([x, y, init_yaw])
        path = car_calculate_motion_path(car, pose, [steer_angle * turn_out_dir, motion_dir], delta_yaw, accuracy)
        if config_env.check_path_feasibility(car, path, boundary_check=False):
            break
    return dist, pose, path
