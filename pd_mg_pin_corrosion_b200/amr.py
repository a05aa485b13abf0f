"""Two-level AMR grid: Python mirror of the reference surface it replaces --
Grid::build_amr / build_neighbors_celllist / update_fictitious (src/grid.h:54-56, src/grid.cpp:352-842) and the
explicit solvers / BCs on that grid (src/pd_ns.cpp, src/pd_ard.cpp, src/boundary.cpp AMR branches).  2D, like the
reference.  The grid build is host code inside libpdgpu.so (no device needed); everything else runs on the B200.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _l
from .config import Config

FICTITIOUS = 6

_GEOM = {
    "pos": (np.float64, lambda i: (i.N_total, 2)), "node_type": (np.uint8, lambda i: (i.N_total,)),
    "dx_local": (np.float64, lambda i: (i.N_total,)), "delta_local": (np.float64, lambda i: (i.N_total,)),
    "grid_level": (np.int32, lambda i: (i.N_total,)), "fict_offset": (np.int32, lambda i: (i.N_total + 1,)),
    "fict_source": (np.int32, lambda i: (i.n_fict_entries,)), "fict_weight": (np.float64, lambda i: (i.n_fict_entries,)),
    "nbr_offset": (np.int32, lambda i: (i.N_total + 1,)), "nbr_index": (np.int32, lambda i: (i.nnz,)),
    "nbr_dist": (np.float64, lambda i: (i.nnz,)), "nbr_evec": (np.float64, lambda i: (i.nnz, 2)),
    "nbr_vol": (np.float64, lambda i: (i.nnz,)), "wall_mirror": (np.int32, lambda i: (i.N_total,)),
}
_FIELDS = {
    "rho": (np.float64, 1), "vel": (np.float64, 2), "pressure": (np.float64, 1), "C": (np.float64, 1),
    "rho_new": (np.float64, 1), "vel_new": (np.float64, 2), "C_new": (np.float64, 1),
    "phase": (np.uint8, 1), "is_gb": (np.uint8, 1), "is_precip": (np.uint8, 1), "node_type": (np.uint8, 1),
}


class AmrGrid:
    """Grid with use_amr = 1 (src/main.cpp:151-154)."""

    def __init__(self, cfg: Config):
        self.cfg = cfg
        self.ctx = C.c_void_p()
        s = cfg.to_struct()
        _l.check(_l.load().pdamr_create(C.byref(s), int(cfg.amr_ratio), float(cfg.amr_buffer), C.byref(self.ctx)))
        self.info = _l.PdAmrInfo()

    def _refresh(self):
        _l.check(_l.load().pdamr_info(self.ctx, C.byref(self.info)))
        self.N_total = int(self.info.N_total)

    def build_amr(self) -> None:
        _l.check(_l.load().pdamr_build(self.ctx))
        self._refresh()

    def build_neighbors_celllist(self) -> None:
        _l.check(_l.load().pdamr_build_neighbors(self.ctx))
        self._refresh()

    def get(self, name: str) -> np.ndarray:
        """geometry arrays by the reference's member names (+ wall_mirror)"""
        dt, shp = _GEOM[name]
        out = np.zeros(shp(self.info), dt)
        if out.size:
            _l.check(_l.load().pdamr_get(self.ctx, name.encode(), out.ctypes.data_as(C.c_void_p)))
        return out

    # ---- device -------------------------------------------------------------------------------
    def device_init(self, device: int = 0) -> None:
        _l.check(_l.load().pdamr_device_init(self.ctx, device))

    def set_field(self, name: str, value) -> None:
        dt, comps = _FIELDS[name]
        a = np.ascontiguousarray(value, dt)
        assert a.size == self.N_total * comps, (name, a.shape)
        _l.check(_l.load().pdamr_field_set(self.ctx, name.encode(), a.ctypes.data_as(C.c_void_p)))

    def get_field(self, name: str) -> np.ndarray:
        dt, comps = _FIELDS[name]
        out = np.zeros((self.N_total, comps) if comps > 1 else (self.N_total,), dt)
        _l.check(_l.load().pdamr_field_get(self.ctx, name.encode(), out.ctypes.data_as(C.c_void_p)))
        return out

    def update_fictitious(self) -> None:
        _l.check(_l.load().pdamr_update_fictitious(self.ctx))

    def inlet_bc(self): _l.check(_l.load().pdamr_bc(self.ctx, 0))
    def outlet_bc(self): _l.check(_l.load().pdamr_bc(self.ctx, 1))
    def wall_bc(self): _l.check(_l.load().pdamr_bc(self.ctx, 2))
    def solid_bc(self): _l.check(_l.load().pdamr_bc(self.ctx, 3))
    def wall_conc_bc(self): _l.check(_l.load().pdamr_bc(self.ctx, 4))
    def wall_bc_new(self): _l.check(_l.load().pdamr_bc(self.ctx, 5))

    def ns_compute_dt(self) -> float:
        dt = C.c_double()
        _l.check(_l.load().pdamr_ns_compute_dt(self.ctx, C.byref(dt)))
        return dt.value

    def ns_step(self, dt: float) -> None:
        _l.check(_l.load().pdamr_ns_step(self.ctx, dt))

    def ns_iterate(self, iters: int, dt: float) -> None:
        _l.check(_l.load().pdamr_ns_iterate(self.ctx, iters, dt))

    def ns_solve_steady(self, verbose: bool = False) -> _l.PdSteadyResult:
        r = _l.PdSteadyResult()
        _l.check(_l.load().pdamr_ns_solve_steady(self.ctx, C.byref(r), 1 if verbose else 0))
        return r

    def ard_set_volume_loss(self, vl: float) -> None:
        _l.check(_l.load().pdamr_ard_set_volume_loss(self.ctx, vl))

    def ard_compute_dt(self) -> float:
        dt = C.c_double()
        _l.check(_l.load().pdamr_ard_compute_dt(self.ctx, C.byref(dt)))
        return dt.value

    def ard_step(self, dt: float) -> None:
        _l.check(_l.load().pdamr_ard_step(self.ctx, dt))

    def ard_iterate(self, steps: int, dt: float) -> None:
        _l.check(_l.load().pdamr_ard_iterate(self.ctx, steps, dt))

    def phase_change(self) -> int:
        n = C.c_int()
        _l.check(_l.load().pdamr_phase_change(self.ctx, C.byref(n)))
        return n.value

    def close(self) -> None:
        if self.ctx:
            _l.load().pdamr_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
