"""ctypes view of oracle/_ref/libpdref{2,3}d.so -- TEST INFRASTRUCTURE ONLY.

The .so is the unmodified reference hot path (compiled by oracle/Makefile from
/root/reference/src, see oracle/ref_shim.cpp).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the
product (pd_mg_pin_corrosion_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
CONFIG_DIR = os.path.join(REPO, "configs")

CONFIG_FIELDS = [
    "dx", "m_ratio", "R_wire", "L_wire", "R_tube", "L_upstream", "L_downstream", "rho_f",
    "mu_f", "gamma_eos", "c0", "eta_density", "Q_flow", "D_liquid", "D_grain", "D_gb",
    "D_precip", "C_solid_init", "C_liquid_init", "C_thresh", "C_sat", "alpha_art_diff",
    "corrosion_decay_l", "cfl_factor", "cfl_factor_corr", "delta", "U_in", "flow_max_iters",
    "flow_conv_tol", "T_final", "corrosion_steps_per_check", "output_every_corr",
    "use_implicit", "channel_flow_corrections", "precip_fraction", "grain_size_mean",
    "gb_width_cells", "precip_cluster_cells", "output_every_flow",
]


CONFIG_EXTRA_FIELDS = ["implicit_dt_fraction", "implicit_dt_max", "implicit_output_every", "diagnostic_every", "newton_tol",
                       "newton_max_iter", "use_amr", "amr_ratio", "amr_buffer", "dx_coarse", "delta_coarse", "rho_m"]


def parse_config_with_reference(dim: int, path: str) -> dict:
    """Config::load + compute_derived of the reference on the file as it stands (no overrides appended): every
    numeric member and output_dir."""
    lib = _lib(dim)
    lib.ref_get_config_extra.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.ref_config_output_dir.argtypes = [C.c_void_p]
    lib.ref_config_output_dir.restype = C.c_char_p
    h = C.c_void_p(lib.ref_create(path.encode()))
    a = (C.c_double * len(CONFIG_FIELDS))()
    b = (C.c_double * len(CONFIG_EXTRA_FIELDS))()
    lib.ref_get_config(h, a)
    lib.ref_get_config_extra(h, b)
    out = dict(zip(CONFIG_FIELDS, list(a)))
    out.update(zip(CONFIG_EXTRA_FIELDS, list(b)))
    out["output_dir"] = lib.ref_config_output_dir(h).decode()
    lib.ref_destroy(h)
    return out


def ref_lib_path(dim: int, implicit: bool = False) -> str:
    """implicit=True: the build that also holds the reference's src/pd_ard_implicit.cpp, compiled against the
    Eigen work-alike oracle/eigen_min/ (see oracle/ref_shim.cpp)."""
    return os.path.join(HERE, "_ref", f"libpdref{'imp' if implicit else ''}{dim}d.so")


def have_ref(dim: int, implicit: bool = False) -> bool:
    return os.path.exists(ref_lib_path(dim, implicit))


_LIBS: dict[tuple, C.CDLL] = {}


def _lib(dim: int, implicit: bool = False) -> C.CDLL:
    if (dim, implicit) in _LIBS:
        return _LIBS[(dim, implicit)]
    lib = C.CDLL(ref_lib_path(dim, implicit))
    vp = C.c_void_p
    if implicit:
        assert lib.ref_has_implicit() == 1
        for name in ("ref_imp_init", "ref_imp_assemble"):
            getattr(lib, name).restype = None
            getattr(lib, name).argtypes = [vp]
        lib.ref_imp_set_volume_loss.restype = None
        lib.ref_imp_set_volume_loss.argtypes = [vp, C.c_double]
        lib.ref_imp_compute_adaptive_dt.restype = C.c_double
        lib.ref_imp_compute_adaptive_dt.argtypes = [vp]
        lib.ref_imp_step.restype = C.c_int
        lib.ref_imp_step.argtypes = [vp, C.c_double]
        lib.ref_imp_phase_change.restype = C.c_int
        lib.ref_imp_phase_change.argtypes = [vp]
        lib.ref_imp_last_info.argtypes = [C.POINTER(C.c_longlong), C.POINTER(C.c_double)]
        lib.ref_imp_last_system.argtypes = [vp] * 5
    lib.ref_create.restype = vp
    lib.ref_create.argtypes = [C.c_char_p]
    lib.ref_ptr.restype = vp
    lib.ref_ptr.argtypes = [vp, C.c_char_p]
    for name in ("ref_ns_compute_dt", "ref_ard_compute_dt"):
        getattr(lib, name).restype = C.c_double
        getattr(lib, name).argtypes = [vp]
    for name in ("ref_time_ns_iterate", "ref_time_ard_iterate"):
        getattr(lib, name).restype = C.c_double
        getattr(lib, name).argtypes = [vp, C.c_int, C.c_double]
    for name in ("ref_ns_iterate", "ref_ard_iterate", "ref_ns_iterate_amr"):
        getattr(lib, name).restype = None
        getattr(lib, name).argtypes = [vp, C.c_int, C.c_double]
    for name in ("ref_ns_step", "ref_ard_step", "ref_ard_set_volume_loss"):
        getattr(lib, name).restype = None
        getattr(lib, name).argtypes = [vp, C.c_double]
    for name in ("ref_destroy", "ref_grid_build", "ref_build_neighbors", "ref_generate_grains",
                 "ref_fields_init", "ref_apply_inlet_bc", "ref_apply_outlet_bc",
                 "ref_apply_wall_bc", "ref_apply_wall_bc_new",
                 "ref_apply_wall_concentration_bc", "ref_apply_solid_surface_bc",
                 "ref_smooth_boundary_concentration",
                 "ref_update_node_types", "ref_ns_init", "ref_swap_flow", "ref_ard_init",
                 "ref_swap_C", "ref_grid_build_amr", "ref_build_neighbors_celllist", "ref_update_fictitious"):
        getattr(lib, name).restype = None
        getattr(lib, name).argtypes = [vp]
    for name in ("ref_ns_solve_steady", "ref_ard_phase_change", "ref_n_grains"):
        getattr(lib, name).restype = C.c_int
        getattr(lib, name).argtypes = [vp]
    lib.ref_fict_entries.restype = C.c_longlong
    lib.ref_fict_entries.argtypes = [vp]
    lib.ref_get_config.argtypes = [vp, C.POINTER(C.c_double)]
    lib.ref_get_dims.argtypes = [vp, C.POINTER(C.c_longlong)]
    lib.ref_get_origin.argtypes = [vp, C.POINTER(C.c_double)]
    lib.ref_write_vti.argtypes = [vp, C.c_char_p]
    lib.ref_write_vti.restype = C.c_double
    lib.ref_main.argtypes = [C.c_char_p]
    lib.ref_main.restype = C.c_int
    lib.ref_write_vtu.argtypes = [vp, C.c_char_p]
    lib.ref_write_vtu.restype = None
    lib.ref_set_threads.argtypes = [C.c_int]
    _LIBS[(dim, implicit)] = lib
    return lib


def write_cfg(base: str | None, overrides: dict | None = None, path: str | None = None) -> str:
    """Write `base` (a file under configs/, or None) + `overrides` to a cfg file.

    Later keys win in the reference parser (src/config.cpp:26-93 is a sequential scan),
    so overrides are simply appended.
    """
    text = ""
    if base is not None:
        src = base if os.path.isabs(base) else os.path.join(CONFIG_DIR, base)
        with open(src) as f:
            text = f.read()
    for k, v in (overrides or {}).items():
        text += f"{k} = {v!r}\n" if isinstance(v, float) else f"{k} = {v}\n"
    if path is None:
        fd, path = tempfile.mkstemp(suffix=".cfg", prefix="pdcfg_")
        os.close(fd)
    with open(path, "w") as f:
        f.write(text)
    return path


class RefSim:
    """One reference simulation state (Config + Grid + Fields + solvers)."""

    def __init__(self, dim: int, base: str | None = "params.cfg", overrides: dict | None = None,
                 threads: int = 4, build: bool = True, fields: bool = True, implicit: bool = False):
        self.dim = dim
        self.lib = _lib(dim, implicit)
        self.lib.ref_set_threads(threads)
        ov = {"use_implicit": 0}
        ov.update(overrides or {})
        self.cfg_path = write_cfg(base, ov)
        self.h = C.c_void_p(self.lib.ref_create(self.cfg_path.encode()))
        buf = (C.c_double * len(CONFIG_FIELDS))()
        self.lib.ref_get_config(self.h, buf)
        self.cfg = dict(zip(CONFIG_FIELDS, list(buf)))
        self.N = 0
        self.amr = False                  # the last `use_amr = ...` line wins, as in the reference parser
        with open(self.cfg_path) as f:
            for line in f:
                k, _, v = line.partition("=")
                if k.strip() == "use_amr":
                    self.amr = bool(int(v.split("#")[0].strip()))
        if build:
            if self.amr:      # src/main.cpp:151-154
                self.lib.ref_grid_build_amr(self.h)
                self.lib.ref_build_neighbors_celllist(self.h)
            else:
                self.lib.ref_grid_build(self.h)
                self.lib.ref_build_neighbors(self.h)
            self._dims()
            if fields:
                self.lib.ref_generate_grains(self.h)
                self.lib.ref_fields_init(self.h)
                self.lib.ref_ns_init(self.h)
                self.lib.ref_ard_init(self.h)

    def _dims(self):
        d = (C.c_longlong * 5)()
        self.lib.ref_get_dims(self.h, d)
        self.Nx, self.Ny, self.Nz, self.N, self.nnz = [int(x) for x in d]
        o = (C.c_double * 3)()
        self.lib.ref_get_origin(self.h, o)
        self.origin = tuple(o)

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None
        try:
            os.unlink(self.cfg_path)
        except OSError:
            pass

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- array views (no copy; invalidated by swaps/rebuilds: re-fetch after those) --
    _SPEC = {
        "pos": (np.float64, "N,DIM"), "node_type": (np.uint8, "N"),
        "nbr_offset": (np.int32, "N+1"), "nbr_index": (np.int32, "NNZ"),
        "nbr_dist": (np.float64, "NNZ"), "nbr_evec": (np.float64, "NNZ,DIM"),
        "nbr_vol": (np.float64, "NNZ"), "rho": (np.float64, "N"), "vel": (np.float64, "N,DIM"),
        "pressure": (np.float64, "N"), "C": (np.float64, "N"), "D_map": (np.float64, "N"),
        "phase": (np.uint8, "N"), "grain_id": (np.int32, "N"), "is_gb": (np.uint8, "N"),
        "is_precip": (np.uint8, "N"), "rho_new": (np.float64, "N"),
        "vel_new": (np.float64, "N,DIM"), "C_new": (np.float64, "N"),
        "dx_local": (np.float64, "N"), "delta_local": (np.float64, "N"), "grid_level": (np.int32, "N"),
        "fict_offset": (np.int32, "N+1"), "fict_source": (np.int32, "NF"), "fict_weight": (np.float64, "NF"),
    }

    def arr(self, name: str) -> np.ndarray:
        dt, shp = self._SPEC[name]
        self._dims()
        dims = {"N": self.N, "N+1": self.N + 1, "NNZ": self.nnz, "DIM": self.dim,
                "NF": int(self.lib.ref_fict_entries(self.h))}
        shape = tuple(dims[s] for s in shp.split(","))
        ptr = self.lib.ref_ptr(self.h, name.encode())
        if not ptr or 0 in shape:
            return np.zeros(shape, dt)
        n = int(np.prod(shape))
        buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dt).reshape(shape)

    def get(self, name: str) -> np.ndarray:
        return self.arr(name).copy()

    def set(self, name: str, value) -> None:
        self.arr(name)[...] = value

    # -- operators --------------------------------------------------------------
    def inlet_bc(self): self.lib.ref_apply_inlet_bc(self.h)
    def outlet_bc(self): self.lib.ref_apply_outlet_bc(self.h)
    def wall_bc(self): self.lib.ref_apply_wall_bc(self.h)
    def wall_bc_new(self): self.lib.ref_apply_wall_bc_new(self.h)
    def wall_conc_bc(self): self.lib.ref_apply_wall_concentration_bc(self.h)
    def solid_bc(self): self.lib.ref_apply_solid_surface_bc(self.h)
    def smooth_conc(self): self.lib.ref_smooth_boundary_concentration(self.h)
    def ns_compute_dt(self) -> float: return self.lib.ref_ns_compute_dt(self.h)
    def ns_step(self, dt): self.lib.ref_ns_step(self.h, dt)
    def ns_iterate(self, n, dt): (self.lib.ref_ns_iterate_amr if self.amr else self.lib.ref_ns_iterate)(self.h, n, dt)
    def update_fictitious(self): self.lib.ref_update_fictitious(self.h)
    def ns_solve_steady(self) -> int: return self.lib.ref_ns_solve_steady(self.h)
    def swap_flow(self): self.lib.ref_swap_flow(self.h)
    def ard_set_volume_loss(self, v): self.lib.ref_ard_set_volume_loss(self.h, v)
    def ard_compute_dt(self) -> float: return self.lib.ref_ard_compute_dt(self.h)
    def ard_step(self, dt): self.lib.ref_ard_step(self.h, dt)
    def ard_iterate(self, n, dt): self.lib.ref_ard_iterate(self.h, n, dt)
    def swap_C(self): self.lib.ref_swap_C(self.h)
    def phase_change(self) -> int: return self.lib.ref_ard_phase_change(self.h)
    def rebuild_neighbors(self):
        self.lib.ref_update_node_types(self.h)
        self.lib.ref_build_neighbors(self.h)
    def rebuild_tables(self):
        """after editing node_type in place: Grid::build_neighbors filters OUTSIDE only -> nothing to do"""

    def write_vti(self, path: str) -> float:
        """VTKWriter::write of the current state; returns the seconds it took"""
        return self.lib.ref_write_vti(self.h, path.encode())

    def write_vtu(self, path: str) -> None:
        """VTKWriter::write_vtu of the current state (AMR clouds)"""
        self.lib.ref_write_vtu(self.h, path.encode())

    # -- implicit branch (implicit=True builds only) ------------------------------
    def imp_init(self): self.lib.ref_imp_init(self.h)
    def imp_set_volume_loss(self, v): self.lib.ref_imp_set_volume_loss(self.h, v)
    def imp_assemble(self): self.lib.ref_imp_assemble(self.h)
    def imp_compute_adaptive_dt(self) -> float: return self.lib.ref_imp_compute_adaptive_dt(self.h)
    def imp_step(self, dt) -> int: return self.lib.ref_imp_step(self.h, dt)
    def imp_phase_change(self) -> int: return self.lib.ref_imp_phase_change(self.h)

    def imp_last_system(self):
        """(A as scipy CSR, b, x before the clamp, iterations, relative residual) of the last imp_step:
        the system the reference's own PD_ARD_ImplicitSolver::step built (src/pd_ard_implicit.cpp:380-409)."""
        import scipy.sparse as sp
        info = (C.c_longlong * 3)()
        err = C.c_double()
        self.lib.ref_imp_last_info(info, C.byref(err))
        n, nnz, iters = int(info[0]), int(info[1]), int(info[2])
        ptr, col = np.zeros(n + 1, np.int64), np.zeros(nnz, np.int32)
        val, b, x = np.zeros(nnz), np.zeros(n), np.zeros(n)
        self.lib.ref_imp_last_system(*[a.ctypes.data_as(C.c_void_p) for a in (ptr, col, val, b, x)])
        return sp.csr_matrix((val, col, ptr), shape=(n, n)), b, x, iters, err.value

    def time_ns(self, n, dt) -> float: return self.lib.ref_time_ns_iterate(self.h, n, dt)
    def time_ard(self, n, dt) -> float: return self.lib.ref_time_ard_iterate(self.h, n, dt)


def run_reference_main(dim: int, cfg_path: str, implicit: bool = False) -> int:
    return _lib(dim, implicit).ref_main(cfg_path.encode())
