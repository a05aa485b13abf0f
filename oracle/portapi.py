"""ctypes view of oracle/libpdoracle.so (pd_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

PDO_FLUID, PDO_SOLID, PDO_WALL, PDO_INLET, PDO_OUTLET, PDO_OUTSIDE = range(6)
HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libpdoracle.so")


class PdoConfig(C.Structure):
    """ctypes image of `struct PdoConfig` (oracle/pd_oracle.h)."""
    _fields_ = (
        [(n, C.c_double) for n in (
            "dx", "R_wire", "L_wire", "R_tube", "L_upstream", "L_downstream",
            "rho_f", "mu_f", "gamma_eos", "c0", "eta_density", "Q_flow", "rho_m",
            "D_liquid", "D_grain", "D_gb", "D_precip",
            "C_solid_init", "C_liquid_init", "C_thresh", "C_sat", "alpha_art_diff",
            "corrosion_decay_l", "cfl_factor", "cfl_factor_corr", "flow_conv_tol", "T_final",
            "delta", "U_in")]
        + [(n, C.c_int) for n in (
            "m_ratio", "flow_max_iters", "corrosion_steps_per_check", "output_every_flow",
            "output_every_corr", "channel_flow_corrections", "use_implicit", "reserved")]
    )

    @classmethod
    def from_obj(cls, obj) -> "PdoConfig":
        """Copy same-named attributes (or dict keys) of a host Config into the POD."""
        s = cls()
        for name, _ in cls._fields_:
            if name == "reserved":
                continue
            v = obj[name] if isinstance(obj, dict) else getattr(obj, name)
            setattr(s, name, int(v) if dict(cls._fields_)[name] is C.c_int else float(v))
        return s


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(
            os.path.join(HERE, "pd_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, dp, u8p, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_int)
        cp = C.POINTER(PdoConfig)
        L.pdo_grid_build.restype = vp
        L.pdo_grid_build.argtypes = [cp, C.c_int]
        L.pdo_grid_free.argtypes = [vp]
        L.pdo_grid_info.argtypes = [vp, C.POINTER(C.c_longlong), dp]
        for n, rt in (("pdo_grid_types", u8p), ("pdo_grid_off_d", ip), ("pdo_grid_off_dist", dp),
                      ("pdo_grid_off_evec", dp), ("pdo_grid_off_vol", dp), ("pdo_grid_wall_mirror", ip)):
            getattr(L, n).restype = rt
            getattr(L, n).argtypes = [vp]
        L.pdo_csr_offsets.argtypes = [vp, C.POINTER(C.c_longlong)]
        L.pdo_csr_fill.argtypes = [vp, C.POINTER(C.c_longlong), ip, dp, dp, dp]
        L.pdo_wall_mirror_build.argtypes = [vp, cp]
        L.pdo_inlet_bc.argtypes = [vp, cp, dp, dp, dp]
        L.pdo_outlet_bc.argtypes = [vp, cp, dp, dp, dp]
        L.pdo_wall_bc.argtypes = [vp, cp, dp, dp]
        L.pdo_wall_conc_bc.argtypes = [vp, dp]
        L.pdo_solid_bc.argtypes = [vp, dp]
        L.pdo_smooth_conc.argtypes = [vp, cp, dp]
        L.pdo_max_fluid_speed.restype = C.c_double
        L.pdo_max_fluid_speed.argtypes = [vp, dp]
        L.pdo_ns_compute_dt.restype = C.c_double
        L.pdo_ns_compute_dt.argtypes = [vp, cp, dp]
        L.pdo_ns_step.argtypes = [vp, cp, C.c_double, dp, dp, dp, dp, dp]
        L.pdo_ns_residual.argtypes = [vp, dp, dp, dp, dp]
        L.pdo_ard_compute_dt.restype = C.c_double
        L.pdo_ard_compute_dt.argtypes = [vp, cp, dp]
        L.pdo_ard_step.argtypes = [vp, cp, C.c_double, C.c_double, dp, dp, u8p, u8p, dp]
        L.pdo_phase_change.restype = C.c_int
        L.pdo_phase_change.argtypes = [vp, cp, u8p, dp, dp, dp, ip]
        L.pdo_ns_iterate.restype = C.c_double
        L.pdo_ns_iterate.argtypes = [vp, cp, C.c_int, C.c_double, dp, dp, dp, dp, dp, dp]
        L.pdo_ard_iterate.restype = C.c_double
        L.pdo_ard_iterate.argtypes = [vp, cp, C.c_int, C.c_double, C.c_double, dp, dp, dp, dp, u8p, u8p]
        L.pdo_set_threads.argtypes = [C.c_int]
        L.pdo_write_vti.argtypes = [vp, C.c_char_p, dp, dp, dp, dp, u8p, vp, vp, u8p, u8p]
        L.pdo_write_vti.restype = C.c_int
        _lib = L
    return _lib


def _dp(a): return a.ctypes.data_as(C.POINTER(C.c_double))
def _u8(a): return a.ctypes.data_as(C.POINTER(C.c_uint8))
def _ip(a): return a.ctypes.data_as(C.POINTER(C.c_int))


def _view(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    addr = C.addressof(ptr.contents)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dtype)


class PortSim:
    """State holder over the plain-C oracle; arrays are numpy, layout as the reference
    (vel is [N, dim])."""

    def __init__(self, dim: int, cfg, threads: int = 4):
        self.L = lib()
        self.L.pdo_set_threads(threads)
        self.dim = dim
        self.cfg = PdoConfig.from_obj(cfg)
        self.g = C.c_void_p(self.L.pdo_grid_build(C.byref(self.cfg), dim))
        info = (C.c_longlong * 7)()
        org = (C.c_double * 3)()
        self.L.pdo_grid_info(self.g, info, org)
        _, self.Nx, self.Ny, self.Nz, self.m, self.n_off, self.N = [int(x) for x in info]
        self.origin = tuple(org)
        self.node_type = _view(self.L.pdo_grid_types(self.g), self.N, np.uint8)
        self.off_d = _view(self.L.pdo_grid_off_d(self.g), 3 * self.n_off, np.int32).reshape(-1, 3)
        self.off_dist = _view(self.L.pdo_grid_off_dist(self.g), self.n_off, np.float64)
        self.off_evec = _view(self.L.pdo_grid_off_evec(self.g), dim * self.n_off, np.float64).reshape(-1, dim)
        self.off_vol = _view(self.L.pdo_grid_off_vol(self.g), self.n_off, np.float64)
        self.wall_mirror = _view(self.L.pdo_grid_wall_mirror(self.g), self.N, np.int32)
        N = self.N
        self.rho = np.zeros(N); self.vel = np.zeros((N, dim)); self.pressure = np.zeros(N)
        self.C = np.zeros(N); self.rho_new = np.zeros(N); self.vel_new = np.zeros((N, dim))
        self.C_new = np.zeros(N)
        self.phase = np.ones(N, np.uint8); self.is_gb = np.zeros(N, np.uint8)
        self.is_precip = np.zeros(N, np.uint8)
        self.volume_loss = 0.0

    def close(self):
        if self.g:
            self.L.pdo_grid_free(self.g)
            self.g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def csr(self):
        off = np.zeros(self.N + 1, np.int64)
        self.L.pdo_csr_offsets(self.g, off.ctypes.data_as(C.POINTER(C.c_longlong)))
        nnz = int(off[-1])
        idx = np.zeros(nnz, np.int32); dist = np.zeros(nnz); evec = np.zeros((nnz, self.dim))
        vol = np.zeros(nnz)
        self.L.pdo_csr_fill(self.g, off.ctypes.data_as(C.POINTER(C.c_longlong)), _ip(idx), _dp(dist),
                            _dp(evec), _dp(vol))
        return off, idx, dist, evec, vol

    # RefSim-compatible accessors
    def get(self, name):
        return np.array(getattr(self, name), copy=True)

    def set(self, name, value):
        getattr(self, name)[...] = value

    def load_state(self, src):
        """Copy rho/vel/C/(_new)/phase/is_gb/is_precip from an object exposing .get(name)."""
        for n in ("rho", "vel", "pressure", "C", "rho_new", "vel_new", "C_new", "phase", "is_gb", "is_precip"):
            getattr(self, n)[...] = src.get(n)

    c = property(lambda self: C.byref(self.cfg))

    def inlet_bc(self): self.L.pdo_inlet_bc(self.g, self.c, _dp(self.rho), _dp(self.vel), _dp(self.C))
    def outlet_bc(self): self.L.pdo_outlet_bc(self.g, self.c, _dp(self.rho), _dp(self.vel), _dp(self.C))
    def wall_bc(self): self.L.pdo_wall_bc(self.g, self.c, _dp(self.rho), _dp(self.vel))
    def wall_bc_new(self): self.L.pdo_wall_bc(self.g, self.c, _dp(self.rho_new), _dp(self.vel_new))
    def wall_conc_bc(self): self.L.pdo_wall_conc_bc(self.g, _dp(self.C))
    def solid_bc(self): self.L.pdo_solid_bc(self.g, _dp(self.vel))
    def smooth_conc(self): self.L.pdo_smooth_conc(self.g, self.c, _dp(self.C))
    def ns_compute_dt(self): return self.L.pdo_ns_compute_dt(self.g, self.c, _dp(self.vel))
    def ard_compute_dt(self): return self.L.pdo_ard_compute_dt(self.g, self.c, _dp(self.vel))

    def ns_step(self, dt):
        self.L.pdo_ns_step(self.g, self.c, dt, _dp(self.rho), _dp(self.vel), _dp(self.pressure),
                           _dp(self.rho_new), _dp(self.vel_new))

    def ns_residual(self):
        out = np.zeros(6)
        self.L.pdo_ns_residual(self.g, _dp(self.vel), _dp(self.vel_new), _dp(self.rho_new), _dp(out))
        return out

    def swap_flow(self):
        self.rho, self.rho_new = self.rho_new, self.rho
        self.vel, self.vel_new = self.vel_new, self.vel

    def swap_C(self):
        self.C, self.C_new = self.C_new, self.C

    def ns_iterate(self, n, dt):
        return self.L.pdo_ns_iterate(self.g, self.c, n, dt, _dp(self.rho), _dp(self.vel), _dp(self.pressure),
                                     _dp(self.C), _dp(self.rho_new), _dp(self.vel_new))

    def ard_step(self, dt):
        self.L.pdo_ard_step(self.g, self.c, dt, self.volume_loss, _dp(self.C), _dp(self.vel),
                            _u8(self.is_gb), _u8(self.is_precip), _dp(self.C_new))

    def ard_iterate(self, n, dt):
        return self.L.pdo_ard_iterate(self.g, self.c, n, dt, self.volume_loss, _dp(self.rho), _dp(self.vel),
                                      _dp(self.C), _dp(self.C_new), _u8(self.is_gb), _u8(self.is_precip))

    def phase_change(self) -> int:
        d = np.zeros(self.N, np.int32)
        n = self.L.pdo_phase_change(self.g, self.c, _u8(self.phase), _dp(self.rho), _dp(self.vel),
                                    _dp(self.C), _ip(d))
        self.last_dissolved = d[:n].copy()
        return n

    def write_vti(self, path: str, grain_id=None, D_map=None) -> float:
        """VTKWriter::write restated (printf "%g"); returns seconds"""
        import time
        gid = None if grain_id is None else np.ascontiguousarray(grain_id, np.int32)
        dm = None if D_map is None else np.ascontiguousarray(D_map, np.float64)
        t0 = time.perf_counter()
        rc = self.L.pdo_write_vti(self.g, path.encode(), _dp(self.rho), _dp(self.vel), _dp(self.pressure),
                                  _dp(self.C), _u8(self.phase),
                                  gid.ctypes.data_as(C.c_void_p) if gid is not None else None,
                                  dm.ctypes.data_as(C.c_void_p) if dm is not None else None,
                                  _u8(self.is_gb), _u8(self.is_precip))
        if rc:
            raise OSError(f"cannot write {path}")
        return time.perf_counter() - t0

    def rebuild_tables(self):
        """after editing node_type in place (hand-built geometries)"""
        self.L.pdo_wall_mirror_build(self.g, self.c)

    def rebuild_neighbors(self):
        pass  # stencil-implicit: nothing to rebuild (wall-mirror table is refreshed by pdo_phase_change)

    def ard_set_volume_loss(self, v):
        self.volume_loss = v

    def init_fields(self, is_gb=None, is_precip=None):
        """initialize_fields (src/main.cpp:9-127) restated with numpy."""
        nt, dim, c = self.node_type, self.dim, self.cfg
        N = self.N
        idx = np.arange(N)
        i = idx % self.Nx
        # pos = origin + i*dx evaluated as one fma in the reference build; the difference is
        # <= 1 ulp in a velocity profile value and only enters fields at the 1e-16 level.
        px = self.origin[0] + i * c.dx
        R2 = c.R_tube * c.R_tube
        if dim == 2:
            rr = np.minimum(px * px / R2, 1.0)
            vax = 1.5 * c.U_in * (1.0 - rr)
        else:
            j = (idx % (self.Nx * self.Ny)) // self.Nx
            py = self.origin[1] + j * c.dx
            rr = np.minimum((px * px + py * py) / R2, 1.0)
            vax = 2.0 * c.U_in * (1.0 - rr)
        self.rho[:] = np.where(nt == PDO_OUTSIDE, 0.0, c.rho_f)
        self.vel[:] = 0.0
        prof = (nt == PDO_FLUID) | (nt == PDO_INLET)
        self.vel[prof, dim - 1] = vax[prof]
        self.C[:] = 0.0
        self.C[nt == PDO_SOLID] = c.C_solid_init
        self.C[(nt == PDO_FLUID) | (nt == PDO_INLET) | (nt == PDO_OUTLET)] = c.C_liquid_init
        self.phase[:] = 1
        self.phase[nt == PDO_SOLID] = 0
        self.is_gb[:] = 0 if is_gb is None else is_gb
        self.is_precip[:] = 0 if is_precip is None else is_precip
        self.rho_new[:] = self.rho
        self.vel_new[:] = self.vel
        self.C_new[:] = self.C
