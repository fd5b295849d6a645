"""Mirror of ``path_planner/car_model.py``: same constructor, attributes and
kinematic helpers; the footprint sweep (``get_path_poly`` + shapely union) is not
materialised on the host -- footprints are evaluated per pose on the GPU from the
rectangle extents kept here."""
import math

import numpy as np

from .geometry_host import Poly
from .utils.path_utils import angle_wrap


class CarModel:
    def __init__(self, max_steer=0.55, wheel_base=1.9, axle_to_front=2.85, axle_to_back=0.5,
                 width=1.48, head_out=0.542, head_side=0.44, body_vertices=[],
                 aux_poly_features=[], with_aux=False):
        self.MAX_STEER = max_steer
        self.WHEEL_BASE = wheel_base
        self.aux_polys = []
        self.AXLE_TO_FRONT = axle_to_front
        self.AXLE_TO_BACK = axle_to_back
        self.WIDTH = width
        self.HEAD_OUT = head_out
        self.HEAD_SIDE = head_side
        self.curvature = math.tan(self.MAX_STEER) / self.WHEEL_BASE      # car_model.py:34
        self.with_aux = with_aux
        self.body_vertices = body_vertices
        self.get_car_poly(aux_poly_features)

    # -- footprint description (car_model.py:75-162) ---------------------------
    def get_car_poly(self, aux_poly_features):
        b, f, hw = self.AXLE_TO_BACK, self.AXLE_TO_FRONT, self.WIDTH / 2
        self.car_poly = Poly([[-b, hw], [-b, -hw], [f, -hw], [f, hw], [-b, hw]])
        self.body_ext = np.array([-b, f, -hw, hw], dtype=np.float64)
        self.aux_exts = np.zeros((0, 4), dtype=np.float64)
        if self.with_aux:
            self.aux_polys = self.get_aux_shapely_polys(aux_poly_features)
            exts = []
            for feat in aux_poly_features:
                x, y, h, w = feat[0][0], feat[0][1], feat[1], feat[2]
                exts.append([x, x + w, y - h, y])
            self.aux_exts = np.array(exts, dtype=np.float64).reshape(-1, 4)

    def get_aux_shapely_polys(self, aux_polys):
        out = []
        for feat in aux_polys:
            x, y, h, w = feat[0][0], feat[0][1], feat[1], feat[2]
            out.append(Poly([[x, y], [x + w, y], [x + w, y - h], [x, y - h]]))
        return out

    def footprint_key(self):
        return (tuple(self.body_ext), tuple(map(tuple, self.aux_exts)))

    def get_car_poly_in_odom(self, odom_x, odom_y, odom_yaw):
        """car_model.py:178-200."""
        rot = np.array([[math.cos(odom_yaw), -math.sin(odom_yaw)],
                        [math.sin(odom_yaw), math.cos(odom_yaw)]])

        def move(poly):
            pts = np.array(poly.exterior.xy)
            pts = np.dot(rot, pts) + np.array([[odom_x], [odom_y]])
            return Poly(pts.T)

        return move(self.car_poly), [move(p) for p in self.aux_polys]

    def draw_car(self, plt, x, y, yaw, color="black", alpha=0.1):
        car_poly, aux_polys = self.get_car_poly_in_odom(x, y, yaw)
        for poly in [car_poly] + aux_polys:
            pts = np.asarray(poly.exterior.xy)
            plt.plot(pts[0, :], pts[1, :], color, alpha=alpha)
            plt.fill(*pts, color="orange", alpha=alpha)
        return car_poly, aux_polys

    # -- kinematic helpers used by the orchestration code ------------------------
    def calculate_motion_path(self, init_pose, motion_command, delta_yaw, step):
        """car_model.py:202-234."""
        steer_angle, speed_direction = motion_command[0], motion_command[1]
        search_length = delta_yaw / self.curvature
        num_steps = round(search_length / step)
        yaw_step = speed_direction * step / self.WHEEL_BASE * math.tan(steer_angle)
        init_yaw = angle_wrap(init_pose[-1] + yaw_step)
        yaws = angle_wrap(np.linspace(init_yaw, init_yaw + yaw_step * (num_steps + 1), num_steps + 1))
        xs = init_pose[0] + np.cumsum(step * np.cos(yaws[:-1]) * speed_direction)
        ys = init_pose[1] + np.cumsum(step * np.sin(yaws[:-1]) * speed_direction)
        path = np.vstack([init_pose, np.vstack([xs, ys, yaws[1:]]).T])
        curvature = math.tan(steer_angle) / self.WHEEL_BASE if abs(steer_angle) > 0.00001 else 0
        ks = np.ones((len(path), 1)) * curvature
        dirs = np.ones((len(path), 1)) * speed_direction
        return np.hstack((path, ks, dirs))

    def calculate_motion_path_new(self, init_pose, motion_dir, steer_dir, turning_radius, delta_yaw,
                                  step_size=0.1):
        """car_model.py:236-269."""
        turning_radius = max(1.0 / self.curvature, turning_radius)
        steer_angle = math.atan(self.WHEEL_BASE / turning_radius) * steer_dir
        arc_length = abs(delta_yaw * turning_radius)
        num_steps = int(arc_length / step_size)
        actual = arc_length / num_steps
        yaw_step = motion_dir * actual / self.WHEEL_BASE * math.tan(steer_angle)
        init_yaw = angle_wrap(init_pose[-1])
        yaws = angle_wrap(np.linspace(init_yaw, init_yaw + yaw_step * num_steps, num_steps + 1))
        xs = init_pose[0] + turning_radius * (np.sin(yaws) - np.sin(init_yaw)) * steer_dir
        ys = init_pose[1] - turning_radius * (np.cos(yaws) - np.cos(init_yaw)) * steer_dir
        path = np.vstack([init_pose, np.vstack([xs, ys, yaws]).T])
        curvature = math.tan(steer_angle) / self.WHEEL_BASE if abs(steer_angle) > 0.00001 else 0
        ks = np.ones((len(path), 1)) * curvature
        dirs = np.ones((len(path), 1)) * motion_dir
        return np.hstack((path, ks, dirs))

    def get_turn_radius(self, max_steer_angle=None):
        if max_steer_angle is None:
            return 1 / self.curvature
        return self.WHEEL_BASE / math.tan(max_steer_angle)

    def get_steer_angle(self, curvature):
        return min(math.atan(self.WHEEL_BASE * curvature), self.MAX_STEER)
