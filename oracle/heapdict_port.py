"""Restatement of ``heapdict`` 1.0.1 (third-party, not vendored by the reference;
``requirements.txt:9``), the priority queue behind both the Hybrid A* open list
(``path_planner/hybrid_a_star_search.py:504-510,542,584-596``) and the
Reeds-Shepp candidate queue (``:265-271``).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  parity unpinned: heapdict is
not installable here, so the published algorithm is restated: a binary min-heap
of ``[priority, key, pos]`` wrappers plus a dict ``key -> wrapper``.

Tie behaviour that the device open list must reproduce:
  * ``_decrease_key`` stops only when ``parent < child`` (strict), so a new
    entry with priority EQUAL to its parent moves above it;
  * ``_min_heapify`` uses strict ``<`` and looks at the left child first;
  * re-keying an existing key first bubbles the old wrapper to the root by
    unconditional swaps and pops it, then appends the new wrapper.
"""


class HeapDict:
    def __init__(self):
        self.heap = []
        self.d = {}

    def __len__(self):
        return len(self.d)

    def __contains__(self, key):
        return key in self.d

    def __getitem__(self, key):
        return self.d[key][0]

    def __setitem__(self, key, value):
        if key in self.d:
            del self[key]
        wrapper = [value, key, len(self)]
        self.d[key] = wrapper
        self.heap.append(wrapper)
        self._decrease_key(len(self.heap) - 1)

    def _min_heapify(self, i):
        n = len(self.heap)
        h = self.heap
        while True:
            l = (i << 1) + 1
            r = (i + 1) << 1
            if l < n and h[l][0] < h[i][0]:
                low = l
            else:
                low = i
            if r < n and h[r][0] < h[low][0]:
                low = r
            if low == i:
                break
            self._swap(i, low)
            i = low

    def _decrease_key(self, i):
        while i:
            parent = (i - 1) >> 1
            if self.heap[parent][0] < self.heap[i][0]:
                break
            self._swap(i, parent)
            i = parent

    def _swap(self, i, j):
        h = self.heap
        h[i], h[j] = h[j], h[i]
        h[i][2] = i
        h[j][2] = j

    def __delitem__(self, key):
        wrapper = self.d[key]
        while wrapper[2]:
            parentpos = (wrapper[2] - 1) >> 1
            parent = self.heap[parentpos]
            self._swap(wrapper[2], parent[2])
        self.popitem()

    def popitem(self):
        wrapper = self.heap[0]
        if len(self.heap) == 1:
            self.heap.pop()
        else:
            self.heap[0] = self.heap.pop()
            self.heap[0][2] = 0
            self._min_heapify(0)
        del self.d[wrapper[1]]
        return wrapper[1], wrapper[0]

    def peekitem(self):
        return (self.heap[0][1], self.heap[0][0])
