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
        _lib = L
    return _lib


class GrainStructure:
    def __init__(self):
        self.grain_id = self.is_grain_boundary = self.is_precipitate = None
        self.n_grains = 0

    def generate(self, node_type: np.ndarray, cfg: Config, dim: int, seed: int = 42) -> "GrainStructure":
        L = _load()
        nt = np.ascontiguousarray(node_type, np.uint8)
        N = nt.size
        self.grain_id = np.full(N, -1, np.int32)
        self.is_grain_boundary = np.zeros(N, np.uint8)
        self.is_precipitate = np.zeros(N, np.uint8)
        s = cfg.to_struct()
        n = C.c_int()
        rc = L.pdhost_generate_grains(C.byref(s), cfg.grain_size_mean, cfg.precip_fraction, cfg.gb_width_cells,
                                      cfg.precip_cluster_cells, dim, nt.ctypes.data_as(C.c_void_p), seed,
                                      self.grain_id.ctypes.data_as(C.c_void_p),
                                      self.is_grain_boundary.ctypes.data_as(C.c_void_p),
                                      self.is_precipitate.ctypes.data_as(C.c_void_p), C.byref(n))
        if rc != 0:
            raise RuntimeError("pdhost_generate_grains failed")
        self.n_grains = n.value
        return self
