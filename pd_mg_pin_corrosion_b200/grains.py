"""GrainStructure (src/grains.h:6-13) for the Python host mirror: ctypes view of the C++ host
generator in host/grains.cpp (same algorithm and libstdc++ RNG calls as the reference's
GrainStructure::generate, so the flags are reproduced bit for bit)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .config import Config, PdConfig

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB = os.path.join(os.path.dirname(HERE), "host", "libpdhost.so")
_lib = None


def _load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB):
            raise RuntimeError(f"{HOST_LIB} is missing: run `make -C host` (or __graft_entry__.build())")
        from . import lib as _l
        _l.load()   # libpdhost.so links libpdgpu.so (grid extents)
        L = C.CDLL(HOST_LIB)
        L.pdhost_generate_grains.restype = C.c_int
        L.pdhost_generate_grains.argtypes = [C.POINTER(PdConfig), C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.POINTER(C.c_int)]
        L.pdhost_generate_grains_device.restype = C.c_int
        L.pdhost_generate_grains_device.argtypes = [C.c_void_p] + L.pdhost_generate_grains.argtypes
        _lib = L
    return _lib


class GrainStructure:
    def __init__(self):
        self.grain_id = self.is_grain_boundary = self.is_precipitate = None
        self.n_grains = 0

    def generate(self, node_type: np.ndarray, cfg: Config, dim: int, seed: int = 42, grid=None) -> "GrainStructure":
        """grid = a built solver.Grid: the Voronoi / grain-boundary / cluster passes run on its device
        (pdgpu_grains_voronoi, pdgpu_grains_grow_precip; collective for slab grids), the random draws stay
        on the host in the reference's order. Same arrays either way."""
        L = _load()
        nt = np.ascontiguousarray(node_type, np.uint8)
        N = nt.size
        self.grain_id = np.full(N, -1, np.int32)
        self.is_grain_boundary = np.zeros(N, np.uint8)
        self.is_precipitate = np.zeros(N, np.uint8)
        s = cfg.to_struct()
        n = C.c_int()
        args = (C.byref(s), cfg.grain_size_mean, cfg.precip_fraction, cfg.gb_width_cells,
                cfg.precip_cluster_cells, dim, nt.ctypes.data_as(C.c_void_p), seed,
                self.grain_id.ctypes.data_as(C.c_void_p), self.is_grain_boundary.ctypes.data_as(C.c_void_p),
                self.is_precipitate.ctypes.data_as(C.c_void_p), C.byref(n))
        rc = L.pdhost_generate_grains(*args) if grid is None else L.pdhost_generate_grains_device(grid.ctx, *args)
        if rc != 0:
            from . import lib as _l
            raise RuntimeError("grain generation failed: " + _l.load().pdgpu_last_error().decode(errors="replace"))
        self.n_grains = n.value
        return self
