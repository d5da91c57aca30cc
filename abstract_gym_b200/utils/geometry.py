"""Point / Line / Square with the reference's names and fields (utils/geometry.py:1-46).

Plain holders.  `Line.compute_line_function` returns the (a, b, c) of a*x + b*y + c = 0 with the
reference's special cases for vertical / horizontal segments (utils/geometry.py:14-32); the
arithmetic runs in the CUDA extension (ag_segment_square), like every other part of the hot path.
"""


class Point:
    __slots__ = ("x", "y")

    def __init__(self, x, y):
        self.x, self.y = x, y

    def __repr__(self):
        return "Point(%r, %r)" % (self.x, self.y)


class Square:
    """Axis-aligned square given by its bottom-left and upper-right corners (utils/geometry.py:35-46)."""
    __slots__ = ("min_x", "min_y", "max_x", "max_y")

    def __init__(self, pbl, pur):
        self.min_x, self.min_y = pbl.x, pbl.y
        self.max_x, self.max_y = pur.x, pur.y

    def __repr__(self):
        return "Square(%r, %r, %r, %r)" % (self.min_x, self.min_y, self.max_x, self.max_y)


class Line:
    """Segment p0 -> p1 (utils/geometry.py:8-12)."""
    __slots__ = ("p0", "p1")

    def __init__(self, p0, p1):
        self.p0, self.p1 = p0, p1

    def compute_line_function(self):
        from .collision_checker import segment_square_arrays
        import numpy as np
        seg = np.array([[self.p0.x, self.p0.y, self.p1.x, self.p1.y]], dtype=np.float64)
        out = segment_square_arrays(seg, np.zeros((1, 4)), want_abc=True)
        a, b, c = (float(v) for v in out["abc"][0])
        return a, b, c
