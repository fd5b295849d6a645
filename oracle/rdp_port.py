"""Ramer-Douglas-Peucker polyline simplification as the ``rdp`` package (>= 0.8, ``requirements.txt:11`` of the
reference, not installable here) publishes it: recursive form, distance of a point to the LINE through the end points
of the span (to the point itself when they coincide), the FIRST farthest point splits a span whose maximum exceeds
epsilon.  TEST INFRASTRUCTURE (see ``oracle/__init__.py``): stands in for ``rdp.rdp`` when the reference's
``OGE_OBCA.py`` is run for goldens, and cross-checks the iterative restatement of the product package.
parity unpinned (no copy of the package to compare with)."""
import numpy as np


def _dist(point, start, end):
    if np.all(np.equal(start, end)):
        return np.linalg.norm(point - start)
    return np.divide(np.abs(np.linalg.norm(np.cross(end - start, start - point))), np.linalg.norm(end - start))


def rdp(M, epsilon=0.0):
    M = np.asarray(M, dtype=float)
    dmax, index = 0.0, -1
    for i in range(1, M.shape[0]):
        d = _dist(M[i], M[0], M[-1])
        if d > dmax:
            index, dmax = i, d
    if dmax > epsilon:
        left = rdp(M[:index + 1], epsilon)
        right = rdp(M[index:], epsilon)
        return np.vstack((left[:-1], right))
    return np.vstack((M[0], M[-1]))
