"""Drop-in for the ``dubins`` package the reference imports (``requirements.txt:14``, pydubins: a Cython wrapper of
Andrew Walker's dubins.c; un-vendored and un-pinned, so parity is unpinned): the two calls the reference makes --

    path = dubins.shortest_path(q0, q1, turning_radius)         hybrid_a_star_search.py:294, navigation_utils.py:211
    configurations, distances = path.sample_many(step_size)     :295, :212

-- answered by the K9 kernels (``hl_dubins_count`` / ``hl_dubins_knots``); batches go through
``ops.dubins_course_batch`` directly."""
import numpy as np

from . import ops

WORDS = ("LSL", "LSR", "RSL", "RSR", "RLR", "LRL")


class DubinsPath:
    def __init__(self, q0, q1, rho):
        self._pair = np.array([[q0[0], q0[1], q0[2], q1[0], q1[1], q1[2]]], dtype=np.float64)
        self._rho = float(rho)
        _, _, word, length = ops.dubins_course_batch(self._pair, self._rho, step=max(self._rho, 1.0), ds=1.0)
        if int(word[0]) < 0:
            raise RuntimeError("path did not work out")
        self._word, self._length = int(word[0]), float(length[0])

    def path_length(self):
        return self._length

    def path_type(self):
        return self._word

    def sample_many(self, step_size):
        """(configurations, distances): t = 0, step, 2*step, ... < length, t accumulated by repeated addition."""
        _, _, _, _, samples, slot_off = ops.dubins_course_batch(self._pair, self._rho, step=float(step_size),
                                                                ds=float(step_size), want_samples=True)
        n = int(slot_off[1]) - 1
        qs = [tuple(r) for r in samples[:n].cpu().numpy().tolist()]
        ts, x = [], 0.0
        for _ in range(n):
            ts.append(x)
            x += step_size
        return qs, ts


def shortest_path(q0, q1, rho):
    return DubinsPath(q0, q1, rho)
