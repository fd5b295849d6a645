"""Mirror of ``path_planner/OGE_OBCA.py``: the convex obstacle polygons the OBCA solve is given
(``test/obca.ipynb`` cell 12: ``create_boundary_polygons`` -> ``get_obstacle_tree_rows`` -> ``get_obstacles_for_OBCA``).

Host-side polytope bookkeeping: a few dozen vertices per headland turn, no data-parallel part,
so it stays numpy -- on top of the mirrored ``OrchardGeometryEnvironment`` (whose collision methods run on the GPU),
without shapely / rdp / pypoman:

* the field polygon is the mirror's vertex ring (``field_range_poly.exterior.coords`` in the reference);
* ``rdp`` (Ramer-Douglas-Peucker, ``rdp>=0.8`` in ``requirements.txt:11``, not installable here) is restated below:
  point-to-LINE distances, first maximum, split while the maximum exceeds epsilon -- parity with the package unpinned;
* the convex hull is scipy's (``scipy.spatial.ConvexHull``, like the reference).

Same class name, constructor, method names, argument meaning and return conventions (lists of ``(k, 2)`` vertex arrays).
"""
import math

import numpy as np
from scipy.spatial import ConvexHull

from .orchard_geometry_environment import OrchardGeometryEnvironment


# ---- module-level helpers (OGE_OBCA.py:11-42) ---------------------------------------------------------------------
def shortest_distance(x1, y1, a, b, c):
    """Distance of the points (x1, y1) to the line a x + b y + c = 0."""
    return np.abs(a * x1 + b * y1 + c) / math.sqrt(a * a + b * b)


def point_along_centerline(A, B, d):
    """The point at signed distance d from the midpoint of AB along the normal (B_y - A_y, -(B_x - A_x)) / |AB|."""
    mid = (A + B) / 2.0
    length = np.linalg.norm(B - A)
    normal = np.array([(B[1] - A[1]) / length, -(B[0] - A[0]) / length])
    return mid + normal * d


def point_side_of_line(A, B, C):
    """Sign of the cross product AB x AC."""
    return np.sign((B[0] - A[0]) * (C[1] - A[1]) - (B[1] - A[1]) * (C[0] - A[0]))


def points_along_rectangles(A, B, d):
    """The two far corners (C above B, D above A) of the rectangle of height d erected on AB."""
    top_mid = point_along_centerline(A, B, d)
    half = (A - B) / 2
    return top_mid + half, top_mid - half


def _line_distances(points, start, end):
    if (start == end).all():
        return np.linalg.norm(points - start, axis=1)
    seg = end - start
    rel = start - points
    return np.abs(seg[0] * rel[:, 1] - seg[1] * rel[:, 0]) / np.linalg.norm(seg)


def rdp(points, epsilon=0.0):
    """Ramer-Douglas-Peucker simplification of an open polyline (the ``rdp`` package's result: distances to the LINE
    through the two end points of a span, the first farthest point splits a span whose maximum exceeds epsilon)."""
    pts = np.asarray(points, dtype=np.float64)
    keep = np.ones(len(pts), dtype=bool)
    spans = [(0, len(pts) - 1)]
    while spans:
        a, b = spans.pop()
        if b - a < 2:
            continue
        d = _line_distances(pts[a + 1:b], pts[a], pts[b])
        j = int(np.argmax(d))
        if d[j] > epsilon:
            spans.append((a, a + 1 + j))
            spans.append((a + 1 + j, b))
        else:
            keep[a + 1:b] = False
    return pts[keep]


def _row_box(row, margin, half_width):
    """Rectangle around one tree row: `margin` beyond both row ends in x, +-half_width in y (near-low, near-high, far-high,
    far-low)."""
    near, far = row[0], row[1]
    return np.array([[near[0] - margin, near[1] - half_width], [near[0] - margin, near[1] + half_width],
                     [far[0] + margin, far[1] + half_width], [far[0] + margin, far[1] - half_width]])


class orchard_environment_OBCA(OrchardGeometryEnvironment):
    MIN_ROW_WIDTH = 0.5
    SAFETY_BOUND = 0.2

    def __init__(self, map_tree_rows, obstacles, contour_points=[], tree_width=0.5, headland_width=7, obstacle_dim=0.3):
        super().__init__(map_tree_rows, obstacles, contour_points=contour_points, tree_width=tree_width,
                         headland_width=headland_width, obstacle_dim=obstacle_dim)
        self.row_width = np.abs(np.mean(np.diff(self.map_tree_rows[:, 0, 1])))

    # ---- orchard_geometry_environment.py:66-95 ---------------------------------------------------------------------
    def create_headland_countour_lines(self, field_range_poly):
        """Vertices of the field polygon on the near / far side of the (jittered, like the reference) fit line through
        the row centres.  ``field_range_poly``: anything with ``.exterior.coords`` (closed ring) or a vertex array."""
        ring = np.array(field_range_poly.exterior.coords)[:-1] if hasattr(field_range_poly, "exterior") \
            else np.asarray(field_range_poly, dtype=np.float64)
        centres = np.mean(self.map_tree_rows[:, :, :], axis=1)
        jitter = np.random.uniform(-0.5, 0.5, size=(len(centres),))
        self._center_line_coeff = np.polyfit(centres[:, 0] + jitter, centres[:, 1], deg=1)
        k, b = self._center_line_coeff[0], self._center_line_coeff[1]
        self._origin_sign = np.sign(0 * k + b - 0)                   # the map origin is on the near side
        side = np.sign(ring[:, 0] * k + b - ring[:, 1])
        return ring[side == self._origin_sign], ring[side != self._origin_sign]

    # ---- OGE_OBCA.py:171-262 -----------------------------------------------------------------------------------------
    def cover_side_points(self, contour_points, side, width=2):
        """Quadrilaterals of depth `width` behind one side's boundary polyline: one quad when the polyline is (nearly)
        straight, else one rectangle per polyline segment."""
        pts = contour_points
        shift = -width if side == self.NEAR_SIDE else width
        if np.std(pts[:, 0]) < 1e-3 or len(pts) == 2:           # vertical fit line, or just two points
            top, bottom = pts[np.argmax(pts[:, 1])], pts[np.argmin(pts[:, 1])]
            return [np.array([top + [shift, 0.0], top, bottom, bottom + [shift, 0.0]])]
        k, b = np.polyfit(pts[:, 1], pts[:, 0], deg=1)              # x = k y + b
        signed = pts[:, 0] - k * pts[:, 1] - b
        inward = np.where(signed >= 0 if side == self.NEAR_SIDE else signed <= 0)[0]
        dist = shortest_distance(pts[inward, 0], pts[inward, 1], 1, -k, -b)
        if dist.mean() < 0.1:                                       # close to a line: one quad through the farthest point
            far_pt = pts[inward[np.argmax(dist)]]
            b_max = far_pt[0] - k * far_pt[1]
            y_hi, y_lo = np.max(pts[:, 1]), np.min(pts[:, 1])
            x_hi, x_lo = y_hi * k + b_max, y_lo * k + b_max
            return [np.array([[x_hi + shift, y_hi], [x_hi, y_hi], [x_lo, y_lo], [x_lo + shift, y_lo]])]
        depth = (-width if side == self.NEAR_SIDE else width) * np.sign(pts[1][1] - pts[0][1])
        quads = []
        for a, c in zip(pts[:-1], pts[1:]):
            p1, p2 = points_along_rectangles(a, c, depth)
            quads.append(np.array([a, c, p1, p2]))
        return quads

    # ---- OGE_OBCA.py:306-373 -----------------------------------------------------------------------------------------
    def create_boundary_polygons(self):
        """(near, far, low, up): lists of convex polygons fencing the field on its four sides."""
        near_pts, far_pts = self.create_headland_countour_lines(self.field_range_poly)
        eps = 0.15
        near = self.cover_side_points(rdp(near_pts, eps), self.NEAR_SIDE)
        far = self.cover_side_points(rdp(far_pts, eps), self.FAR_SIDE)

        def bar(row_index, dy):                                     # a 1 m thick bar one row width beyond the outermost row
            a, c = np.copy(self.map_tree_rows[row_index, 0, :]), np.copy(self.map_tree_rows[row_index, 1, :])
            a[1] += dy * self.row_width; c[1] += dy * self.row_width
            a[0] -= 8; c[0] += 8
            return np.vstack([a, a + [0.0, dy * 1.0], c + [0.0, dy * 1.0], c])
        up = bar(np.argmax(self.map_tree_rows[:, 0, 1]), +1)
        low = bar(np.argmin(self.map_tree_rows[:, 0, 1]), -1)
        return near, far, [low], [up]

    # ---- OGE_OBCA.py:375-409 -----------------------------------------------------------------------------------------
    def polygon_to_convex_sets(self, coords):
        """The convex hull of the points as a one-element list (hull vertices in scipy's counter-clockwise order)."""
        pts = np.array(coords)
        hull = ConvexHull(pts)
        return [np.array([pts[v] for v in hull.vertices])]

    def _rows_between(self, start_pose, end_pose):
        ys = self.map_tree_rows[:, 0, 1]
        lo, hi = min(start_pose[1], end_pose[1]), max(start_pose[1], end_pose[1])
        return np.sort(np.where((ys > lo) & (ys < hi))[0])

    # ---- OGE_OBCA.py:411-475 -----------------------------------------------------------------------------------------
    def get_tree_row_obstacles(self, start_pose, end_pose):
        """The rows next to the block crossed by the turn as boxes (rounded to 7 decimals) + the block itself as ONE
        convex polygon."""
        between = self._rows_between(start_pose, end_pose)
        lo, hi = between[0], between[-1]
        rows = self.map_tree_rows
        low_edge = np.copy(rows[lo]); low_edge[:, 1] -= self.tree_width / 2.0
        up_edge = np.copy(rows[hi]); up_edge[:, 1] += self.tree_width / 2.0
        block = np.concatenate([low_edge, rows[between, 1, :], up_edge, rows[between[::-1], 0, :]])
        out = [np.round(_row_box(rows[i], self.SAFETY_BOUND, self.tree_width / 2.0), 7)
               for i in (min(hi + 1, len(rows) - 1), max(lo - 1, 0))]
        return out + self.polygon_to_convex_sets(block)

    # ---- OGE_OBCA.py:477-591 -----------------------------------------------------------------------------------------
    def get_obstacle_tree_rows(self, start_pose, end_pose):
        """Boxes around the tree rows relevant to a turn from start_pose to end_pose."""
        between = self._rows_between(start_pose, end_pose)
        lo, hi = between[0], between[-1]
        rows = self.map_tree_rows
        if lo == hi:                                                # one row crossed: rows hi-2 .. hi+1, rounded
            first, last = max(hi - 2, 0), min(hi + 2, len(rows) - 1)
            return [np.round(_row_box(rows[i], self.SAFETY_BOUND, self.tree_width / 2.0), 7) for i in range(first, last)]
        first, last = max(lo - 2, 0), min(hi + 3, len(rows) - 1)
        return [_row_box(rows[i], self.SAFETY_BOUND, self.tree_width / 2.0) for i in range(first, last)]

    # ---- OGE_OBCA.py:593-677 -----------------------------------------------------------------------------------------
    def get_obstacles_for_OBCA(self, boundary_polys, row_polys, start_pose, end_pose, side, width=2, buffer_distance=1):
        """Obstacle list of one turn: the side boundary next to the turn (its rectangles near the turn merged into convex
        chains), the low or up bar, and the row boxes."""
        side_polys = boundary_polys[0 if side == self.NEAR_SIDE else 1]
        out = []
        if len(side_polys) <= 1:
            out.append(*side_polys)
        else:
            y_hi, y_lo = max(start_pose[1], end_pose[1]), min(start_pose[1], end_pose[1])
            near_turn = [p for p in side_polys
                         if np.max(p[:2, 1]) > y_lo - buffer_distance and np.min(p[:2, 1]) < y_hi + buffer_distance]
            chain = np.vstack([np.array([p[0, :] for p in near_turn]), near_turn[-1][1, :]])
            y_dir = np.sign(chain[1][1] - chain[0][1])
            depth = (-width if side == self.NEAR_SIDE else width) * y_dir
            turn = (1 if side == self.NEAR_SIDE else -1) * y_dir

            def close(piece):                                       # erect the rectangle behind first -> last vertex
                p1, p2 = points_along_rectangles(piece[0], piece[-1], depth)
                return np.vstack([piece, p1, p2])
            piece = np.array([chain[0], chain[1]])
            end = 1
            while end < len(chain) - 1:
                if point_side_of_line(chain[end - 1], chain[end], chain[end + 1]) == turn:   # still convex: extend
                    piece = np.vstack([piece, chain[end + 1]])
                    end += 1
                else:                                               # close the piece, start a new one at this vertex
                    out.append(close(piece))
                    piece = np.array([chain[end], chain[end + 1]])
                    end += 1
            out.append(close(piece))
        out += [p for p in boundary_polys[2 if start_pose[1] > end_pose[1] else 3]]
        return out + row_polys
