"""CollisionChecker with the reference's interface (utils/collision_checker.py:7-91), evaluated by
the CUDA extension, plus the array form used by tests and tools."""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from .._device import ptr, require_cuda, stream_ptr
from .geometry import Line, Point, Square  # noqa: F401  (re-exported like the reference module)


def segment_square_arrays(seg, sq, section_eps=1e-10, want_abc=False, want_corner_values=False, device=None):
    """seg [n,4] (p0x,p0y,p1x,p1y), sq [n,4] (min_x,min_y,max_x,max_y) -> dict(hit, abc, v, axis_aligned).

    Accepts numpy arrays or CUDA tensors; returns numpy arrays."""
    dev = require_cuda(device)
    lib = _lib.load()
    seg_t = torch.as_tensor(seg, dtype=torch.float64, device=dev).reshape(-1, 4).contiguous()
    sq_t = torch.as_tensor(sq, dtype=torch.float64, device=dev).reshape(-1, 4).contiguous()
    n = seg_t.shape[0]
    if sq_t.shape[0] != n:
        raise ValueError("seg and sq must have the same length")
    hit = torch.zeros(n, dtype=torch.uint8, device=dev)
    abc = torch.zeros(n, 3, dtype=torch.float64, device=dev) if want_abc else None
    v = torch.zeros(n, 4, dtype=torch.float64, device=dev) if want_corner_values else None
    axis = torch.zeros(1, dtype=torch.int64, device=dev)
    _lib.check(lib.ag_segment_square(ptr(seg_t), ptr(sq_t), section_eps, ptr(hit), ptr(abc), ptr(v), ptr(axis), n,
                                     stream_ptr(dev)), "ag_segment_square")
    return dict(hit=hit.cpu().numpy().astype(bool), abc=None if abc is None else abc.cpu().numpy(),
                v=None if v is None else v.cpu().numpy(), axis_aligned=int(axis.item()))


class CollisionChecker:
    """Check whether a line segment collides with a square region (same semantics as the
    reference, including "a segment wholly inside the square is not a collision")."""

    def __init__(self, line, square):
        self.l = line
        self.s = square
        self.a, self.b, self.c = self.l.compute_line_function()

    def _arrays(self):
        seg = np.array([[self.l.p0.x, self.l.p0.y, self.l.p1.x, self.l.p1.y]], dtype=np.float64)
        sq = np.array([[self.s.min_x, self.s.min_y, self.s.max_x, self.s.max_y]], dtype=np.float64)
        return seg, sq

    def compute_corner_line_value(self):
        """signs of a*x + b*y + c at the four corners (utils/collision_checker.py:23-32)"""
        seg, sq = self._arrays()
        return np.sign(segment_square_arrays(seg, sq, want_corner_values=True)["v"][0])

    def collision_check(self):
        """utils/collision_checker.py:34-46.  Where the reference raises AttributeError (an
        axis-aligned segment whose line crosses the square, :59-68) this returns the
        interval-overlap result that code intends."""
        seg, sq = self._arrays()
        return bool(segment_square_arrays(seg, sq)["hit"][0])

    def check_sections(self):
        """Only meaningful after the corner-sign test passed; kept for interface parity."""
        return self.collision_check()

    def take_x(self, p):
        return p.x

    def compute_lambda(self, x):
        return (x - self.l.p0.x) / (self.l.p1.x - self.l.p0.x)
