"""ctypes binding of libpdgpu.so (include/pdgpu.h). No fallback: a missing library or a
missing CUDA device raises."""
from __future__ import annotations

import ctypes as C
import os
import re

from .config import PdConfig

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpdgpu.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "pdgpu.h")

# Field ids (include/pdgpu.h)
F_RHO, F_VEL, F_PRESSURE, F_C, F_RHO_NEW, F_VEL_NEW, F_C_NEW, F_PHASE, F_IS_GB, F_IS_PRECIP, F_NODE_TYPE = range(11)
FLUID, SOLID_MG, WALL, INLET, OUTLET, OUTSIDE = range(6)


class PdGridInfo(C.Structure):
    _fields_ = [("dim", C.c_int), ("Nx", C.c_int), ("Ny", C.c_int), ("Nz", C.c_int),
                ("m", C.c_int), ("n_off", C.c_int), ("reach", C.c_int),
                ("a0", C.c_int), ("a1", C.c_int),
                ("N_total", C.c_longlong), ("plane", C.c_longlong), ("counts", C.c_longlong * 6),
                ("ns_bonds", C.c_longlong), ("ard_bonds", C.c_longlong), ("nnz", C.c_longlong),
                ("origin", C.c_double * 3)]


class PdResidual(C.Structure):
    _fields_ = [("num", C.c_double), ("den", C.c_double), ("v_max", C.c_double),
                ("rho_min", C.c_double), ("rho_max", C.c_double), ("has_nan", C.c_int), ("pad", C.c_int)]


class PdSteadyResult(C.Structure):
    _fields_ = [("iters", C.c_int), ("status", C.c_int), ("eps", C.c_double), ("dt", C.c_double),
                ("v_max", C.c_double), ("rho_min", C.c_double), ("rho_max", C.c_double),
                ("poiseuille_l2", C.c_double), ("poiseuille_nodes", C.c_int), ("pad", C.c_int)]


class PdAmrInfo(C.Structure):
    _fields_ = [("N_total", C.c_longlong), ("n_fine", C.c_longlong), ("n_coarse", C.c_longlong),
                ("n_fict", C.c_longlong), ("nnz", C.c_longlong), ("n_fict_entries", C.c_longlong),
                ("counts", C.c_longlong * 7), ("origin", C.c_double * 2), ("dx_coarse", C.c_double),
                ("delta_coarse", C.c_double)]


class PdDiag(C.Structure):
    _fields_ = [("solid_count", C.c_longlong), ("v_max", C.c_double), ("C_max_fluid", C.c_double)]


class PdLinSolveInfo(C.Structure):
    _fields_ = [("iters", C.c_int), ("converged", C.c_int), ("rel_res", C.c_double), ("pad", C.c_int)]


class PdGpuError(RuntimeError):
    pass


def declared_symbols(header: str = HEADER_PATH) -> list[str]:
    """Every function include/pdgpu.h declares (used by the CPU-side export test)."""
    text = open(header).read()
    return sorted(set(re.findall(r"\b(pd(?:gpu|amr)_[a-z0-9_]+)\s*\(", text)))


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PdGpuError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, ip, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
    cfgp = C.POINTER(PdConfig)
    L.pdgpu_last_error.restype = C.c_char_p
    sig = {
        "pdgpu_device_count": [ip],
        "pdgpu_grid_extents": [cfgp, C.c_int, ip, ip, ip, dp],
        "pdgpu_partition": [C.c_int, C.c_int, C.c_int, ip, ip],
        "pdgpu_partition_balanced": [cfgp, C.c_int, C.c_int, C.c_int, ip, ip],
        "pdgpu_slab_layout_range": [C.c_int, C.c_int, C.c_longlong, C.c_int, C.POINTER(C.c_longlong)],
        "pdgpu_slab_layout": [C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)],
        "pdgpu_stencil": [cfgp, C.c_int, ip, vp, vp, vp, vp],
        "pdgpu_create": [cfgp, C.c_int, C.c_int, C.POINTER(vp)],
        "pdgpu_create_slab": [cfgp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)],
        "pdgpu_destroy": [vp], "pdgpu_sync": [vp],
        "pdgpu_grid_build": [vp], "pdgpu_grid_set_types": [vp, vp],
        "pdgpu_grid_info": [vp, C.POINTER(PdGridInfo)],
        "pdgpu_grid_build_neighbors": [vp, C.POINTER(C.c_longlong)],
        "pdgpu_grid_download_csr": [vp, vp, vp, vp, vp, vp],
        "pdgpu_grid_free_neighbors": [vp], "pdgpu_grid_download_wall_mirror": [vp, vp],
        "pdgpu_fields_upload": [vp, C.c_int, vp], "pdgpu_fields_download": [vp, C.c_int, vp],
        "pdgpu_fields_init": [vp, vp, vp], "pdgpu_swap_flow": [vp], "pdgpu_swap_C": [vp],
        "pdgpu_gather": [vp, C.c_int, vp, C.c_longlong, vp],
        "pdgpu_bc_inlet": [vp], "pdgpu_bc_outlet": [vp], "pdgpu_bc_wall": [vp], "pdgpu_bc_wall_new": [vp],
        "pdgpu_bc_wall_conc": [vp], "pdgpu_bc_solid": [vp], "pdgpu_bc_smooth_conc": [vp],
        "pdgpu_ns_compute_dt": [vp, dp], "pdgpu_ns_step": [vp, C.c_double],
        "pdgpu_ns_iterate": [vp, C.c_int, C.c_double], "pdgpu_ns_residual": [vp, C.POINTER(PdResidual)],
        "pdgpu_ns_solve_steady": [vp, C.POINTER(PdSteadyResult), C.c_int],
        "pdgpu_ard_set_volume_loss": [vp, C.c_double], "pdgpu_ard_compute_dt": [vp, dp],
        "pdgpu_ard_step": [vp, C.c_double], "pdgpu_ard_iterate": [vp, C.c_int, C.c_double],
        "pdgpu_step_iterate": [vp, C.c_int, C.c_double, C.c_double],
        "pdgpu_phase_change": [vp, ip, vp, C.c_int], "pdgpu_diag": [vp, C.POINTER(PdDiag)],
        "pdgpu_comm_get_uid": [vp], "pdgpu_comm_init": [vp, vp, C.c_int, C.c_int],
        "pdgpu_halo_exchange": [vp, C.c_int],
        "pdgpu_comm_allreduce": [vp, dp, C.c_int, C.c_int],
        "pdgpu_solid_below_thresh": [vp, ip],
        "pdgpu_implicit_assemble": [vp], "pdgpu_implicit_compute_dt": [vp, C.c_double, C.c_double, dp],
        "pdgpu_implicit_step": [vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(PdLinSolveInfo)],
        "pdgpu_implicit_matvec": [vp, C.c_double, vp, vp], "pdgpu_implicit_rhs": [vp, C.c_double, vp],
        "pdgpu_grains_voronoi": [vp, vp, C.c_int, C.c_int, vp, vp],
        "pdgpu_grains_grow_precip": [vp, vp, vp, C.c_int, vp],
        "pdgpu_fields_download_all": [vp, C.c_int, vp],
        "pdgpu_timer_start": [vp], "pdgpu_timer_stop": [vp, C.POINTER(C.c_float)],
        "pdgpu_launch_count": [vp, C.POINTER(C.c_longlong), C.c_int],
        "pdgpu_set_option": [vp, C.c_char_p, C.c_int], "pdgpu_flush_l2": [vp],
        "pdgpu_time_kernel": [vp, C.c_int, C.c_int, C.POINTER(C.c_float)],
        "pdgpu_fp64_peak": [vp, dp], "pdgpu_fp64_peak3": [vp, dp],
        "pdgpu_step_host": [vp, C.c_double, C.c_double, vp, vp, vp, C.c_int],
        "pdgpu_step_host_chunks": [vp, C.c_int, ip, C.c_char_p, C.c_int],
        "pdgpu_step_host_trace": [vp, dp, C.c_int, ip],
        "pdgpu_vti_write": [vp, C.c_char_p, vp, vp, C.POINTER(C.c_longlong), C.POINTER(C.c_float)],
        "pdgpu_format_g": [vp, vp, C.c_longlong, vp],
        "pdgpu_checkpoint_save": [vp, C.c_char_p, C.POINTER(C.c_longlong)], "pdgpu_checkpoint_load": [vp, C.c_char_p],
        "pdgpu_host_register": [vp, C.c_size_t], "pdgpu_host_unregister": [vp],
        "pdamr_create": [cfgp, C.c_int, C.c_double, C.POINTER(vp)], "pdamr_build": [vp], "pdamr_build_neighbors": [vp],
        "pdamr_info": [vp, C.POINTER(PdAmrInfo)], "pdamr_get": [vp, C.c_char_p, vp],
        "pdamr_device_init": [vp, C.c_int], "pdamr_field_set": [vp, C.c_char_p, vp], "pdamr_field_get": [vp, C.c_char_p, vp],
        "pdamr_update_fictitious": [vp], "pdamr_bc": [vp, C.c_int],
        "pdamr_ns_compute_dt": [vp, dp], "pdamr_ns_step": [vp, C.c_double], "pdamr_ns_iterate": [vp, C.c_int, C.c_double],
        "pdamr_ns_solve_steady": [vp, C.POINTER(PdSteadyResult), C.c_int],
        "pdamr_ard_set_volume_loss": [vp, C.c_double], "pdamr_ard_compute_dt": [vp, dp],
        "pdamr_ard_step": [vp, C.c_double], "pdamr_ard_iterate": [vp, C.c_int, C.c_double],
        "pdamr_phase_change": [vp, ip], "pdamr_destroy": [vp],
        "pdamr_implicit_assemble": [vp], "pdamr_implicit_matvec": [vp, C.c_double, vp, vp],
        "pdamr_implicit_rhs": [vp, C.c_double, vp], "pdamr_implicit_compute_dt": [vp, C.c_double, C.c_double, dp],
        "pdamr_implicit_step": [vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(PdLinSolveInfo)],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    L.pdgpu_version.restype = C.c_int
    L.pdgpu_comm_uid_bytes.restype = C.c_int
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise PdGpuError(load().pdgpu_last_error().decode(errors="replace"))
