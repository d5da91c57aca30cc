"""Occupancy grids: the reference's `OccupancyGrid` (environment/occupancy_grid.py:7-93) plus the
device-side form the kernels read (bit-packed rows + float64 corner tables) and a batched
many-grids variant.

Frame convention, reproduced from environment/occupancy_grid.py:28,59-67: S cells per side, each
of side E/(S-1); cell (row r, col c) has bottom-left corner ((c*E)/(S-1) - E/2, -((r*E)/(S-1) - E/2));
row 0 is the TOP row; the grid is not centred.
"""
import ctypes as C

import os

import numpy as np
import torch

from .. import _lib
from .._device import ptr, require_cuda, stream_ptr
from ..utils.geometry import Point, Square


def _wants_transposed(S, stride_words):
    """A transposed copy of the bits lets the FAST engine walk shallow links by columns.  It pays when both copies
    fit the kernels' shared-memory staging window (32 KB; e.g. the 256x256 per-batch maps: 2 x 8 KB); for a grid
    that stays in global memory (1024x1024: 2 x 128 KB) the second array only thrashes L1 (measured: 261 -> 292 ms)."""
    spad = (S + 1) & ~1
    form = os.environ.get("AG_GRID_FORM", "")
    if form:
        return form == "transposed"
    return 32 < S < 64 and 16 + 2 * stride_words * 4 + spad * 16 <= 32768


def _wants_hier(S):
    """The two-level form (8x8 tiles + summary bitmap, ag_grid.hier) for every map beyond the obstacle-list class:
    the FAST engine then walks S/8 summary lines per link instead of S (measured on configs 4 / 5: DESIGN.md)."""
    form = os.environ.get("AG_GRID_FORM", "")
    if form:
        return form == "hier"
    return S >= 64


class DeviceGrid:
    """What the kernels read: bits[n_grids][stride_words] uint32, min_x[S], min_y[S] float64."""

    def __init__(self, bits, S, environment_size, n_grids=1, envs_per_grid=1 << 62, max_occupied=None, bits_t=None,
                 hier=None):
        lib = _lib.load()
        self.device = bits.device
        self.S = int(S)
        self.environment_size = float(environment_size)
        self.words_per_row = int(lib.ag_grid_words_per_row(self.S))
        self.stride_words = int(lib.ag_grid_stride_words(self.S))
        self.n_grids = int(n_grids)
        self.envs_per_grid = int(envs_per_grid)
        assert bits.dtype == torch.int32 and bits.numel() == self.n_grids * self.stride_words
        self.bits = bits
        self.bits_t = bits_t          # transposed copy (column-major lines) for the minor-axis traversal, or None
        self.max_occupied = -1 if max_occupied is None else int(max_occupied)
        self.hier = hier              # two-level form (tiles + summary bitmap), uint8 [n_grids * hier_bytes], or None
        if hier is None and _wants_hier(self.S):
            hb = int(lib.ag_grid_hier_bytes(self.S))
            self.hier = torch.empty(self.n_grids * hb, dtype=torch.uint8, device=self.device)
            _lib.check(lib.ag_grid_pack_hier(ptr(self.bits), self.S, self.n_grids, self.stride_words, ptr(self.hier),
                                             stream_ptr(self.device)), "ag_grid_pack_hier")
        spad = (self.S + 1) & ~1
        mx = np.zeros(spad, dtype=np.float64)
        my = np.zeros(spad, dtype=np.float64)
        side = C.c_double()
        _lib.check(lib.ag_grid_tables_host(self.S, self.environment_size, mx.ctypes.data_as(C.c_void_p),
                                           my.ctypes.data_as(C.c_void_p), C.byref(side)), "ag_grid_tables_host")
        self.side = side.value
        self.min_x_host, self.min_y_host = mx[:self.S].copy(), my[:self.S].copy()
        self.min_x = torch.from_numpy(mx).to(self.device)
        self.min_y = torch.from_numpy(my).to(self.device)

    def c_struct(self, envs_per_grid=None) -> _lib.Grid:
        g = _lib.Grid()
        g.bits, g.min_x, g.min_y = self.bits.data_ptr(), self.min_x.data_ptr(), self.min_y.data_ptr()
        g.side, g.env_size = self.side, self.environment_size
        g.S, g.words_per_row, g.n_grids = self.S, self.words_per_row, self.n_grids
        g.max_occupied = self.max_occupied
        g.grid_stride_words = self.stride_words
        g.envs_per_grid = self.envs_per_grid if envs_per_grid is None else int(envs_per_grid)
        g.bits_t = None if self.bits_t is None else self.bits_t.data_ptr()
        g.hier = None if self.hier is None else self.hier.data_ptr()
        return g

    @classmethod
    def from_host_matrix(cls, occ, environment_size=1.6, device=None):
        dev = require_cuda(device)
        lib = _lib.load()
        occ8 = np.ascontiguousarray(np.asarray(occ) != 0, dtype=np.uint8)
        if occ8.ndim != 2:
            raise ValueError("occupancy matrix must be 2-D")
        words = np.zeros(int(lib.ag_grid_stride_words(occ8.shape[0])), dtype=np.uint32)
        _lib.check(lib.ag_grid_pack_host(occ8.ctypes.data_as(C.c_void_p), occ8.shape[0], occ8.shape[1],
                                         words.ctypes.data_as(C.c_void_p)), "ag_grid_pack_host")
        bits = torch.from_numpy(words.view(np.int32)).to(dev)
        bits_t = None
        if _wants_transposed(occ8.shape[0], len(words)) and occ8.shape[0] == occ8.shape[1]:
            words_t = np.zeros_like(words)
            _lib.check(lib.ag_grid_pack_host(np.ascontiguousarray(occ8.T).ctypes.data_as(C.c_void_p), occ8.shape[0],
                                             occ8.shape[1], words_t.ctypes.data_as(C.c_void_p)), "ag_grid_pack_host")
            bits_t = torch.from_numpy(words_t.view(np.int32)).to(dev)
        return cls(bits, occ8.shape[0], environment_size, max_occupied=int(occ8.sum()), bits_t=bits_t)

    @classmethod
    def from_device_matrices(cls, occ, environment_size=1.6, envs_per_grid=1 << 62):
        """occ: [G,S,S] CUDA tensor (any integer/bool dtype) -> packed on the device (kernel K5)."""
        if occ.dim() == 2:
            occ = occ.unsqueeze(0)
        if occ.dim() != 3 or occ.shape[1] != occ.shape[2]:
            raise ValueError("expected [G,S,S] occupancy matrices")
        dev = require_cuda(occ.device)
        lib = _lib.load()
        G, S = occ.shape[0], occ.shape[1]
        occ8 = (occ != 0).to(torch.uint8).contiguous()
        stride = int(lib.ag_grid_stride_words(S))
        bits = torch.zeros(G * stride, dtype=torch.int32, device=dev)
        _lib.check(lib.ag_grid_pack(ptr(occ8), S, G, ptr(bits), stride, stream_ptr(dev)), "ag_grid_pack")
        bits_t = None
        if _wants_transposed(S, stride):
            occ8_t = occ8.transpose(1, 2).contiguous()
            bits_t = torch.zeros(G * stride, dtype=torch.int32, device=dev)
            _lib.check(lib.ag_grid_pack(ptr(occ8_t), S, G, ptr(bits_t), stride, stream_ptr(dev)), "ag_grid_pack")
        return cls(bits, S, environment_size, n_grids=G, envs_per_grid=envs_per_grid, bits_t=bits_t,
                   max_occupied=int(occ8.sum(dim=(1, 2), dtype=torch.int64).max().item()))

    def unpack(self, indices=None):
        """[G,S,S] uint8 numpy (inverse of the packing; for tests and tools); `indices`: only these grids"""
        bits = self.bits.view(self.n_grids, self.stride_words)
        if indices is not None:
            bits = bits[torch.as_tensor(list(indices), device=self.device, dtype=torch.long)]
        w = bits.cpu().numpy().view(np.uint32).reshape(-1, self.stride_words)
        w = w[:, :self.S * self.words_per_row].reshape(-1, self.S, self.words_per_row)
        b = ((w[..., None] >> np.arange(32, dtype=np.uint32)) & 1).astype(np.uint8)
        return b.reshape(w.shape[0], self.S, self.words_per_row * 32)[:, :, :self.S]


class OccupancyGrid:
    """Same constructor, attributes and methods as the reference class.

    occ                  : the size x size matrix (non-zero = obstacle)
    occ_coordinate       : [M,2] bottom-left corners in the robot frame (after transform_frame)
    obstacle_list        : M `Square`s, in the reference's order
    obstacle_side_length : environment_size / (size - 1)

    Differences, both counted in DESIGN.md: a grid with no obstacle is accepted (the reference
    raises ValueError at occupancy_grid.py:64); `manual_cells` lets a caller state the hard-coded
    example map explicitly."""

    MANUAL_CELLS = ((5, 6), (5, 7), (2, 3))   # (row, col) of occupancy_grid.py:45-47

    def __init__(self, size=9, random_obstacle=True, obstacle_probability=0.1, environment_size=1.6):
        self.size = size
        self.environment_size = environment_size
        self.obstacle_side_length = environment_size / (size - 1)
        self.obstacle_list = []
        self._device_grids = {}
        if random_obstacle:
            # same draw as the reference (one np.random.random((size,size)) from the global stream)
            self.occ = np.where(np.random.random((size, size)) < obstacle_probability, 1, 0)
            rows, cols = np.nonzero(self.occ)
            self.occ_coordinate = [[c, r] for r, c in zip(rows, cols)]
        else:
            self.occ = np.zeros((size, size))
            self.occ_coordinate = []
            for r, c in self.MANUAL_CELLS:      # list order of occupancy_grid.py:48-50
                self.occ[r][c] = 1
                self.occ_coordinate.append([c, r])
        self.transform_frame()

    def transform_frame(self):
        """matrix (col,row) -> robot frame bottom-left corners, then Squares (occupancy_grid.py:54-68)"""
        self._device_grids = {}
        self.obstacle_list = []
        idx = np.array(self.occ_coordinate, dtype=np.int64).reshape(-1, 2)
        if idx.shape[0] == 0:
            self.occ_coordinate = np.zeros((0, 2))
            return
        xy = idx * self.environment_size / (self.size - 1) - self.environment_size / 2.0
        xy *= [1, -1]
        self.occ_coordinate = xy
        s = self.obstacle_side_length
        self.obstacle_list = [Square(Point(x, y), Point(x + s, y + s)) for x, y in xy]

    def get_occupancy_grid(self):
        return self.occ, self.occ_coordinate, self.obstacle_list, self.obstacle_side_length

    def load_from_matrix(self, matrix, environment_size=1.6):
        """occupancy_grid.py:73-93"""
        matrix = np.asarray(matrix)
        if matrix.ndim != 2 or matrix.shape[0] != matrix.shape[1]:
            print("The matrix is not square.")
            return
        self.size = matrix.shape[0]
        self.occ = np.copy(matrix)
        self.environment_size = environment_size
        self.obstacle_side_length = environment_size / (self.size - 1)
        rows, cols = np.nonzero(self.occ)
        self.occ_coordinate = [[c, r] for r, c in zip(rows, cols)]
        self.transform_frame()

    # ---- device form -------------------------------------------------------------------------
    def device_grid(self, device=None) -> DeviceGrid:
        dev = require_cuda(device)
        key = str(dev)
        if key not in self._device_grids:
            self._device_grids[key] = DeviceGrid.from_host_matrix(self.occ, self.environment_size, dev)
        return self._device_grids[key]


class BatchedOccupancyGrid:
    """G distinct S x S grids resident on the device; env with global id e uses grid
    (e // envs_per_grid) % G (BASELINE config 5: per-env-batch heterogeneous maps)."""

    def __init__(self, occ, envs_per_grid, environment_size=1.6, device=None):
        dev = require_cuda(device if device is not None else (occ.device if torch.is_tensor(occ) else None))
        occ_t = torch.as_tensor(occ, device=dev)
        self._grid = DeviceGrid.from_device_matrices(occ_t, environment_size, envs_per_grid)
        self.size = self._grid.S
        self.environment_size = environment_size
        self.obstacle_side_length = self._grid.side
        self.envs_per_grid = envs_per_grid
        self.n_grids = self._grid.n_grids

    @classmethod
    def random(cls, n_grids, size, obstacle_probability, envs_per_grid, environment_size=1.6, device=None,
               generator=None, clear_base_cells=0):
        """i.i.d. Bernoulli(obstacle_probability) maps (occupancy_grid.py:35-36 for every grid).  clear_base_cells = k > 0
        frees the (2k) x (2k) cells around the arm's base at the origin: a map whose base cell is occupied has no free
        pose at all (the reference's random_valid_pose would never return, scene_0.py:179)."""
        dev = require_cuda(device)
        occ = torch.rand(n_grids, size, size, device=dev, generator=generator) < obstacle_probability
        k = int(clear_base_cells)
        if k > 0:
            # the origin is the corner shared by cells (row, col) = (S/2 - 1 .. S/2 + 1 ...): x = c*side - E/2 = 0 at
            # c = (S-1)/2, y = E/2 - r*side = 0 at r = (S-1)/2 (occupancy_grid.py:59-67)
            mid = (size - 1) / 2.0
            lo, hi = int(np.floor(mid)) - k + 1, int(np.ceil(mid)) + k
            occ[:, max(lo, 0):hi + 1, max(lo - 1, 0):hi] = False
        return cls(occ, envs_per_grid, environment_size, dev)

    @classmethod
    def clustered(cls, n_grids, size, n_blobs, blob_radius_cells, envs_per_grid, clear_radius=0.12, environment_size=1.6,
                  device=None, generator=None):
        """Clustered obstacles for high-resolution maps (i.i.d. dense cells leave no free pose at
        1024x1024): `n_blobs` discs of `blob_radius_cells` cells per grid at uniform random centres, with a disc of
        `clear_radius` metres around the arm's base kept free so that free poses exist.  Set-up code (torch on the
        device), not part of the step path."""
        dev = require_cuda(device)
        S = int(size)
        side = environment_size / (S - 1)
        idx = torch.arange(S, device=dev, dtype=torch.float32)
        cx = (torch.rand(n_grids, n_blobs, device=dev, generator=generator) * S).view(n_grids, n_blobs, 1, 1)
        cy = (torch.rand(n_grids, n_blobs, device=dev, generator=generator) * S).view(n_grids, n_blobs, 1, 1)
        rr = (idx.view(1, 1, S, 1) - cy) ** 2 + (idx.view(1, 1, 1, S) - cx) ** 2        # [G, B, row, col]
        occ = (rr <= float(blob_radius_cells) ** 2).any(dim=1)
        # cell centres in metres (occupancy_grid.py:59-67: col c -> x = c*side - E/2, row r -> y = E/2 - r*side, + side/2)
        xc = idx * side - environment_size / 2 + side / 2
        yc = environment_size / 2 - idx * side + side / 2
        clear = (yc.view(S, 1) ** 2 + xc.view(1, S) ** 2) <= (clear_radius + side) ** 2
        occ &= ~clear.view(1, S, S)
        return cls(occ, envs_per_grid, environment_size, dev)

    def device_grid(self, device=None) -> DeviceGrid:
        return self._grid
