"""Host-side mirror of the reference's operator surface over libpdgpu.so.

Same names, argument order and meaning as the C++ the reference's main.cpp/coupling.cpp
call:

    Grid.build(cfg) / Grid.build_neighbors()                 src/grid.h:52-53
    Fields.allocate / swap_buffers / swap_all_buffers        src/fields.h:28-58
    initialize_fields(fields, grid, grains, cfg)             src/main.cpp:9-127
    apply_inlet_bc ... apply_solid_surface_bc                src/boundary.h:6-13
    PD_NS_Solver.{init, compute_dt, step, solve_steady}      src/pd_ns.h:9-17
    PD_ARD_Solver.{init, set_volume_loss, compute_dt, step, apply_phase_change}  src/pd_ard.h:9-20
    CoupledSolver.run(grid, fields, cfg)                     src/coupling.cpp:82-302 (explicit branch)

The difference: arrays live in HBM.  `Fields` members are device-backed -- reading
`fields.rho` downloads, assigning uploads -- and the solvers run on the device arrays
without host round trips.  Everything raises if the CUDA library or a GPU is missing.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import lib as _l
from .config import Config


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Grid:
    """Grid (src/grid.h:19-68) with the arrays on device."""

    def __init__(self, dim: int, device: int = 0, rank: int = 0, nranks: int = 1):
        self.dim, self.device, self.rank, self.nranks = dim, device, rank, nranks
        self.ctx = C.c_void_p()
        self.cfg: Config | None = None
        self.info = _l.PdGridInfo()
        self._nnz = -1

    # -- Grid::build (src/grid.cpp:29-155) ------------------------------------------
    def build(self, cfg: Config) -> None:
        L = _l.load()
        cfg.check_supported()
        self.cfg = cfg
        s = cfg.to_struct()
        if self.ctx:
            _l.check(L.pdgpu_destroy(self.ctx))
            self.ctx = C.c_void_p()
        if self.nranks == 1:
            _l.check(L.pdgpu_create(C.byref(s), self.dim, self.device, C.byref(self.ctx)))
        else:
            _l.check(L.pdgpu_create_slab(C.byref(s), self.dim, self.device, self.rank, self.nranks,
                                         C.byref(self.ctx)))
        _l.check(L.pdgpu_grid_build(self.ctx))
        self._refresh()

    def _refresh(self) -> None:
        _l.check(_l.load().pdgpu_grid_info(self.ctx, C.byref(self.info)))
        i = self.info
        self.Nx, self.Ny, self.Nz, self.N_total = i.Nx, i.Ny, i.Nz, i.N_total
        self.dx, self.delta, self.m = self.cfg.dx, self.cfg.delta, self.cfg.m_ratio
        self.origin_x, self.origin_y, self.origin_z = i.origin[0], i.origin[1], i.origin[2]
        self.a0, self.a1, self.plane = i.a0, i.a1, i.plane

    # -- z-slab runs: one process per GPU (the reference has no distributed layer) --
    def comm_init(self, uid: bytes) -> None:
        """Join the NCCL communicator of the slab run; `uid` comes from rank 0's comm_uid()."""
        _l.check(_l.load().pdgpu_comm_init(self.ctx, uid, self.rank, self.nranks))
        self._refresh()

    @staticmethod
    def comm_uid() -> bytes:
        L = _l.load()
        buf = (C.c_ubyte * L.pdgpu_comm_uid_bytes())()
        _l.check(L.pdgpu_comm_get_uid(buf))
        return bytes(buf)

    def comm_init_torch(self) -> None:
        """comm_init with the id broadcast over an initialised torch.distributed process group."""
        import torch
        import torch.distributed as dist
        nb = _l.load().pdgpu_comm_uid_bytes()
        dev = torch.device("cuda", self.device) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.zeros(nb, dtype=torch.uint8, device=dev)
        if self.rank == 0:
            t = torch.tensor(list(self.comm_uid()), dtype=torch.uint8, device=dev)
        dist.broadcast(t, 0)
        self.comm_init(bytes(t.cpu().tolist()))

    def allreduce(self, vals, op: str = "sum") -> np.ndarray:
        """Host scalars combined over the ranks (no-op for one rank)."""
        a = np.ascontiguousarray(np.atleast_1d(vals), np.float64).copy()
        _l.check(_l.load().pdgpu_comm_allreduce(self.ctx, a.ctypes.data_as(C.POINTER(C.c_double)), a.size,
                                                {"sum": 0, "max": 1, "min": 2}[op]))
        return a

    @property
    def node_type_all(self) -> np.ndarray:
        """Node types of the WHOLE grid on every rank (collective for slab contexts)."""
        out = np.full(self.N_total, _l.OUTSIDE, np.uint8)
        _l.check(_l.load().pdgpu_fields_download_all(self.ctx, _l.F_NODE_TYPE, _ptr(out)))
        return out

    def set_node_types(self, node_type: np.ndarray) -> None:
        """Hand-built geometry (tests/test_implicit.cpp:737-772 style)."""
        nt = np.ascontiguousarray(node_type, np.uint8)
        assert nt.size == self.N_total
        _l.check(_l.load().pdgpu_grid_set_types(self.ctx, _ptr(nt)))
        self._refresh()

    # -- Grid::build_neighbors (src/grid.cpp:157-294) -------------------------------
    def build_neighbors(self) -> int:
        nnz = C.c_longlong()
        _l.check(_l.load().pdgpu_grid_build_neighbors(self.ctx, C.byref(nnz)))
        self._nnz = nnz.value
        return self._nnz

    def free_neighbors(self) -> None:
        _l.check(_l.load().pdgpu_grid_free_neighbors(self.ctx))
        self._nnz = -1

    def csr(self):
        """(nbr_offset[int64], nbr_index, nbr_dist, nbr_evec[nnz,dim], nbr_vol) of the owned rows."""
        if self._nnz < 0:
            self.build_neighbors()
        n_own = (self.a1 - self.a0) * self.plane
        off = np.zeros(n_own + 1, np.int64)
        idx = np.zeros(self._nnz, np.int32)
        dist = np.zeros(self._nnz)
        evec = np.zeros((self._nnz, self.dim))
        vol = np.zeros(self._nnz)
        _l.check(_l.load().pdgpu_grid_download_csr(self.ctx, _ptr(off), _ptr(idx), _ptr(dist), _ptr(evec), _ptr(vol)))
        return off, idx, dist, evec, vol

    @property
    def node_type(self) -> np.ndarray:
        out = np.full(self.N_total, _l.OUTSIDE, np.uint8)
        _l.check(_l.load().pdgpu_fields_download(self.ctx, _l.F_NODE_TYPE, _ptr(out)))
        return out

    @property
    def wall_mirror(self) -> np.ndarray:
        out = np.full(self.N_total, -1, np.int32)
        _l.check(_l.load().pdgpu_grid_download_wall_mirror(self.ctx, _ptr(out)))
        return out

    def stencil(self):
        L = _l.load()
        s = self.cfg.to_struct()
        n = C.c_int()
        _l.check(L.pdgpu_stencil(C.byref(s), self.dim, C.byref(n), None, None, None, None))
        d = np.zeros((n.value, 3), np.int32)
        dist = np.zeros(n.value)
        evec = np.zeros((n.value, self.dim))
        vol = np.zeros(n.value)
        _l.check(L.pdgpu_stencil(C.byref(s), self.dim, C.byref(n), _ptr(d), _ptr(dist), _ptr(evec), _ptr(vol)))
        return d, dist, evec, vol

    def idx(self, i: int, j: int, k: int = 0) -> int:
        return j * self.Nx + i if self.dim == 2 else k * self.Nx * self.Ny + j * self.Nx + i

    def sync(self) -> None:
        _l.check(_l.load().pdgpu_sync(self.ctx))

    def set_option(self, name: str, value: int) -> None:
        _l.check(_l.load().pdgpu_set_option(self.ctx, name.encode(), value))

    def launch_count(self, reset: bool = False) -> int:
        """kernels (and graph nodes) this context has launched since the last reset"""
        n = C.c_longlong()
        _l.check(_l.load().pdgpu_launch_count(self.ctx, C.byref(n), 1 if reset else 0))
        return n.value

    def close(self) -> None:
        if self.ctx:
            _l.load().pdgpu_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_FIELD_IDS = {"rho": _l.F_RHO, "vel": _l.F_VEL, "pressure": _l.F_PRESSURE, "C": _l.F_C,
              "rho_new": _l.F_RHO_NEW, "vel_new": _l.F_VEL_NEW, "C_new": _l.F_C_NEW, "phase": _l.F_PHASE,
              "is_gb": _l.F_IS_GB, "is_precip": _l.F_IS_PRECIP}
_U8 = {"phase", "is_gb", "is_precip"}


class Fields:
    """Fields (src/fields.h:7-59), device-backed. `D_map` and `grain_id` are never read by the
    explicit solvers and stay host-side numpy arrays."""

    def __init__(self):
        object.__setattr__(self, "_grid", None)
        object.__setattr__(self, "D_map", None)
        object.__setattr__(self, "grain_id", None)

    def allocate(self, N: int, grid: Grid | None = None) -> None:
        if grid is not None:
            object.__setattr__(self, "_grid", grid)
        object.__setattr__(self, "D_map", np.zeros(N))
        object.__setattr__(self, "grain_id", np.full(N, -1, np.int32))

    def bind(self, grid: Grid) -> None:
        object.__setattr__(self, "_grid", grid)

    def get(self, name: str) -> np.ndarray:
        g = self._grid
        shape = (g.N_total, g.dim) if name in ("vel", "vel_new") else (g.N_total,)
        out = np.zeros(shape, np.uint8 if name in _U8 else np.float64)
        _l.check(_l.load().pdgpu_fields_download(g.ctx, _FIELD_IDS[name], _ptr(out)))
        return out

    def get_all(self, name: str) -> np.ndarray:
        """The whole global array on every rank (collective for slab contexts)."""
        g = self._grid
        shape = (g.N_total, g.dim) if name in ("vel", "vel_new") else (g.N_total,)
        out = np.zeros(shape, np.uint8 if name in _U8 else np.float64)
        _l.check(_l.load().pdgpu_fields_download_all(g.ctx, _FIELD_IDS[name], _ptr(out)))
        return out

    def set(self, name: str, value) -> None:
        g = self._grid
        shape = (g.N_total, g.dim) if name in ("vel", "vel_new") else (g.N_total,)
        a = np.ascontiguousarray(np.broadcast_to(value, shape), np.uint8 if name in _U8 else np.float64)
        _l.check(_l.load().pdgpu_fields_upload(g.ctx, _FIELD_IDS[name], _ptr(a)))

    def __getattr__(self, name):
        if name in _FIELD_IDS:
            return self.get(name)
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in _FIELD_IDS:
            self.set(name, value)
        else:
            object.__setattr__(self, name, value)

    def swap_buffers(self) -> None:
        _l.check(_l.load().pdgpu_swap_flow(self._grid.ctx))

    def swap_C(self) -> None:
        _l.check(_l.load().pdgpu_swap_C(self._grid.ctx))

    def swap_all_buffers(self) -> None:
        self.swap_buffers()
        self.swap_C()

    def gather(self, name: str, idx: np.ndarray) -> np.ndarray:
        idx = np.ascontiguousarray(idx, np.int32)
        out = np.zeros(idx.size)
        _l.check(_l.load().pdgpu_gather(self._grid.ctx, _FIELD_IDS[name], _ptr(idx), idx.size, _ptr(out)))
        return out


def initialize_fields(fields: Fields, grid: Grid, grains, cfg: Config) -> None:
    """initialize_fields (src/main.cpp:9-127). `grains` exposes is_grain_boundary / is_precipitate
    / grain_id arrays [N_total] (or is None: no grain structure)."""
    fields.bind(grid)
    gb = pr = None
    if grains is not None:
        gb = np.ascontiguousarray(grains.is_grain_boundary, np.uint8)
        pr = np.ascontiguousarray(grains.is_precipitate, np.uint8)
        fields.grain_id = np.asarray(grains.grain_id, np.int32).copy()
    _l.check(_l.load().pdgpu_fields_init(grid.ctx, _ptr(gb) if gb is not None else None,
                                         _ptr(pr) if pr is not None else None))
    if fields.D_map is not None:   # host-only output field
        nt = grid.node_type_all if grid.nranks > 1 else grid.node_type
        D = np.zeros(grid.N_total)
        D[(nt == _l.FLUID) | (nt == _l.INLET) | (nt == _l.OUTLET)] = cfg.D_liquid
        if gb is not None:
            s = nt == _l.SOLID_MG
            D[s] = np.where(gb[s] != 0, cfg.D_gb, np.where(pr[s] != 0, cfg.D_precip, cfg.D_grain))
        fields.D_map = D


# ---- boundary operators (src/boundary.h:6-13) ---------------------------------------------
def apply_inlet_bc(f: Fields, g: Grid, cfg: Config) -> None: _l.check(_l.load().pdgpu_bc_inlet(g.ctx))
def apply_outlet_bc(f: Fields, g: Grid, cfg: Config) -> None: _l.check(_l.load().pdgpu_bc_outlet(g.ctx))
def apply_wall_bc(f: Fields, g: Grid, cfg: Config) -> None: _l.check(_l.load().pdgpu_bc_wall(g.ctx))
def apply_wall_bc_new(f: Fields, g: Grid, cfg: Config) -> None: _l.check(_l.load().pdgpu_bc_wall_new(g.ctx))
def apply_wall_concentration_bc(f: Fields, g: Grid, cfg: Config) -> None: _l.check(_l.load().pdgpu_bc_wall_conc(g.ctx))
def apply_solid_surface_bc(f: Fields, g: Grid) -> None: _l.check(_l.load().pdgpu_bc_solid(g.ctx))
def smooth_boundary_concentration(f: Fields, g: Grid, cfg: Config) -> None: _l.check(_l.load().pdgpu_bc_smooth_conc(g.ctx))


def update_node_types_after_dissolution(g: Grid, f: Fields) -> None:
    """Already a no-op in the reference (src/boundary.cpp:395-402): apply_phase_change set the types."""


def step_host(grid: Grid, dt_ns: float, dt_ard: float, rho: np.ndarray, vel: np.ndarray, Cc: np.ndarray,
              n_chunks: int = 32) -> None:
    """One coupling-loop pass (NS loop body, src/pd_ns.cpp:196-205,325; ARD loop body,
    src/coupling.cpp:232-240) on HOST arrays, in place: the call a driver that keeps the
    reference's Fields vectors in host memory makes. rho [N], vel [N,dim], Cc [N], float64,
    C-contiguous; pinned memory lets uploads, kernels and downloads overlap."""
    for a in (rho, vel, Cc):
        if a.dtype != np.float64 or not a.flags.c_contiguous:
            raise ValueError("step_host: float64 C-contiguous arrays required")
    if rho.shape != (grid.N_total,) or Cc.shape != (grid.N_total,) or vel.shape != (grid.N_total, grid.dim):
        raise ValueError("step_host: arrays must be global [N_total] / [N_total, dim]")
    _l.check(_l.load().pdgpu_step_host(grid.ctx, dt_ns, dt_ard, _ptr(rho), _ptr(vel), _ptr(Cc), n_chunks))


def step_host_chunks(grid: Grid, n_chunks: int) -> tuple[int, str]:
    """(chunks pdgpu_step_host uses on this geometry, reason when it fell back to one)"""
    n = C.c_int()
    why = C.create_string_buffer(160)
    _l.check(_l.load().pdgpu_step_host_chunks(grid.ctx, n_chunks, C.byref(n), why, 160))
    return n.value, why.value.decode()


class PD_NS_Solver:
    """PD_NS_Solver (src/pd_ns.h:9-17)."""

    def init(self, grid: Grid, cfg: Config) -> None:
        self.last = None

    def compute_dt(self, fields: Fields, grid: Grid, cfg: Config) -> float:
        dt = C.c_double()
        _l.check(_l.load().pdgpu_ns_compute_dt(grid.ctx, C.byref(dt)))
        return dt.value

    def step(self, fields: Fields, grid: Grid, cfg: Config, dt: float) -> None:
        _l.check(_l.load().pdgpu_ns_step(grid.ctx, dt))

    def iterate(self, fields: Fields, grid: Grid, cfg: Config, iters: int, dt: float) -> None:
        """`iters` loop bodies of solve_steady without the convergence block (device resident)."""
        _l.check(_l.load().pdgpu_ns_iterate(grid.ctx, iters, dt))

    def residual(self, grid: Grid) -> _l.PdResidual:
        r = _l.PdResidual()
        _l.check(_l.load().pdgpu_ns_residual(grid.ctx, C.byref(r)))
        return r

    def solve_steady(self, fields: Fields, grid: Grid, cfg: Config, verbose: bool = True) -> int:
        r = _l.PdSteadyResult()
        _l.check(_l.load().pdgpu_ns_solve_steady(grid.ctx, C.byref(r), 1 if verbose else 0))
        self.last = r
        return r.iters


class PD_ARD_Solver:
    """PD_ARD_Solver (src/pd_ard.h:9-20), explicit."""

    def init(self, grid: Grid, cfg: Config) -> None:
        self.volume_loss_fraction = 0.0

    def set_volume_loss(self, vl: float, grid: Grid | None = None) -> None:
        self.volume_loss_fraction = vl
        if grid is not None:
            _l.check(_l.load().pdgpu_ard_set_volume_loss(grid.ctx, vl))

    def compute_dt(self, fields: Fields, grid: Grid, cfg: Config) -> float:
        dt = C.c_double()
        _l.check(_l.load().pdgpu_ard_compute_dt(grid.ctx, C.byref(dt)))
        return dt.value

    def step(self, fields: Fields, grid: Grid, cfg: Config, dt: float) -> None:
        _l.check(_l.load().pdgpu_ard_set_volume_loss(grid.ctx, self.volume_loss_fraction))
        _l.check(_l.load().pdgpu_ard_step(grid.ctx, dt))

    def iterate(self, fields: Fields, grid: Grid, cfg: Config, steps: int, dt: float) -> None:
        _l.check(_l.load().pdgpu_ard_set_volume_loss(grid.ctx, self.volume_loss_fraction))
        _l.check(_l.load().pdgpu_ard_iterate(grid.ctx, steps, dt))

    def apply_phase_change(self, fields: Fields, grid: Grid, cfg: Config) -> int:
        n = C.c_int()
        cap = max(int(grid.info.counts[_l.SOLID_MG]), 1)
        out = np.zeros(cap, np.int32)
        _l.check(_l.load().pdgpu_phase_change(grid.ctx, C.byref(n), _ptr(out), cap))
        self.last_dissolved = out[:n.value].copy()
        if n.value and fields.D_map is not None:
            fields.D_map[self.last_dissolved] = cfg.D_liquid   # src/pd_ard.cpp:202
        grid._refresh()
        return n.value


class PD_ARD_ImplicitSolver:
    """PD_ARD_ImplicitSolver (src/pd_ard_implicit.h:12-40): backward Euler with the bond operator applied
    matrix-free on the device and a restarted GMRES. The reference solves the same system with
    Eigen's GMRES + IncompleteLUT (tolerance 1e-10, restart 50, 200 iterations)."""

    def __init__(self, tol: float = 1e-10, restart: int = 50, max_iters: int = 2000, precond: int = 2):
        self.tol, self.restart, self.max_iters, self.precond = tol, restart, max_iters, precond
        self.volume_loss_fraction = 0.0
        self.last = None

    def init(self, grid: Grid, cfg: Config) -> None:
        pass

    def set_volume_loss(self, vl: float, grid: Grid | None = None) -> None:
        self.volume_loss_fraction = vl
        if grid is not None:
            _l.check(_l.load().pdgpu_ard_set_volume_loss(grid.ctx, vl))

    def assemble(self, fields: Fields, grid: Grid, cfg: Config) -> None:
        _l.check(_l.load().pdgpu_ard_set_volume_loss(grid.ctx, self.volume_loss_fraction))
        _l.check(_l.load().pdgpu_implicit_assemble(grid.ctx))

    def compute_adaptive_dt(self, fields: Fields, grid: Grid, cfg: Config) -> float:
        dt = C.c_double()
        _l.check(_l.load().pdgpu_implicit_compute_dt(grid.ctx, cfg.implicit_dt_fraction, cfg.implicit_dt_max, C.byref(dt)))
        return dt.value

    def step(self, fields: Fields, grid: Grid, cfg: Config, dt: float) -> int:
        info = _l.PdLinSolveInfo()
        _l.check(_l.load().pdgpu_implicit_step(grid.ctx, dt, self.tol, self.restart, self.max_iters, self.precond,
                                               C.byref(info)))
        self.last = info
        return 1

    def matvec(self, grid: Grid, dt: float, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float64)
        y = np.zeros_like(x)
        _l.check(_l.load().pdgpu_implicit_matvec(grid.ctx, dt, _ptr(x), _ptr(y)))
        return y

    def rhs(self, grid: Grid, dt: float) -> np.ndarray:
        b = np.zeros(grid.N_total)
        _l.check(_l.load().pdgpu_implicit_rhs(grid.ctx, dt, _ptr(b)))
        return b

    def apply_phase_change(self, fields: Fields, grid: Grid, cfg: Config) -> int:
        return PD_ARD_Solver.apply_phase_change(self, fields, grid, cfg)   # same rule (src/pd_ard_implicit.cpp:540-561)


def diagnostics(grid: Grid) -> _l.PdDiag:
    d = _l.PdDiag()
    _l.check(_l.load().pdgpu_diag(grid.ctx, C.byref(d)))
    return d


class VTKWriter:
    """VTKWriter (src/vtk_writer.h): write() is pdgpu_vti_write -- the snapshot text is formatted on the
    device, byte-identical to the reference's file; the PVD collection is rewritten after every entry."""

    def __init__(self):
        self.pvd_entries: list[tuple[float, str]] = []
        self.pvd_path = ""

    def write(self, filename: str, grid: Grid, fields: "Fields", cfg: Config) -> None:
        gid = None if fields.grain_id is None else np.ascontiguousarray(fields.grain_id, np.int32)
        dmap = None if fields.D_map is None else np.ascontiguousarray(fields.D_map, np.float64)
        _l.check(_l.load().pdgpu_vti_write(grid.ctx, filename.encode(), None if gid is None else _ptr(gid),
                                           None if dmap is None else _ptr(dmap), None, None))

    def set_pvd_path(self, path: str) -> None:
        self.pvd_path = path

    def add_timestep(self, time: float, vti_file: str) -> None:
        self.pvd_entries.append((time, vti_file))
        if self.pvd_path:
            self.write_pvd(self.pvd_path)

    def write_pvd(self, filename: str) -> None:   # src/vtk_writer.cpp:157-186
        pvd_dir = filename[:filename.rfind("/") + 1] if "/" in filename else ""
        with open(filename, "w") as out:
            out.write('<?xml version="1.0"?>\n<VTKFile type="Collection" version="1.0" byte_order="LittleEndian">\n'
                      "  <Collection>\n")
            for t, f in self.pvd_entries:
                rel = f[len(pvd_dir):] if pvd_dir and f.startswith(pvd_dir) else f
                out.write(f'    <DataSet timestep="{t:.6e}" file="{rel}"/>\n')
            out.write("  </Collection>\n</VTKFile>\n")


class CoupledSolver:
    """CoupledSolver::run, explicit branch (src/coupling.cpp:82-302): diagnostics.csv / mass_loss.csv
    and (write_vti = True) the state_/flow_/corr_/final_ VTI series + PVD files as the reference writes
    them."""

    def __init__(self):
        self.write_vti = False
        self.writer, self.flow_writer, self.frame_count = VTKWriter(), VTKWriter(), 0
        self.flow_solver = PD_NS_Solver()
        self.ard_solver = PD_ARD_Solver()
        self.ard_implicit_solver = PD_ARD_ImplicitSolver()
        self.total_implicit_steps = 0
        self.initial_solid_indices = np.zeros(0, np.int32)
        self.total_dissolved = 0
        self.dissolved_since_flow = 0
        self.log = print

    # ordered host sum: (1 - sum/n) is a catastrophic cancellation, so the
    # gathered values are added sequentially in index order like src/coupling.cpp:32-38.
    def _solid_C_sum(self, fields: Fields) -> float:
        vals = fields.gather("C", self.initial_solid_indices)
        s = 0.0
        for v in vals.tolist():
            s += v
        return s

    @staticmethod
    def _any_solid_below(grid: Grid, fields: Fields, cfg: Config) -> bool:
        n = C.c_int()
        _l.check(_l.load().pdgpu_solid_below_thresh(grid.ctx, C.byref(n)))
        return n.value > 0

    def write_diagnostics(self, grid: Grid, fields: Fields, t_corr: float, cfg: Config) -> None:
        d = diagnostics(grid)                      # reductions over all ranks
        n0 = len(self.initial_solid_indices)
        loss = (1.0 - self._solid_C_sum(fields) / (n0 + 1e-30)) * 100.0   # collective gather, same sum on every rank
        if loss < 0.0:
            loss = 0.0
        if grid.rank != 0:
            return
        self.log(f"  t={t_corr:.1f} s ({t_corr / 3600.0:.2f} h)  pin_mass_loss={loss:.2f}%  solid={d.solid_count}"
                 f"  v_max={d.v_max:.3e}  C_max_fluid={d.C_max_fluid:.4f}")
        with open(os.path.join(cfg.output_dir, "diagnostics.csv"), "a") as f:
            f.write(f"{t_corr:.6e},{t_corr / 3600.0:.6e},{loss:.6e},{d.solid_count},{d.v_max:.6e},"
                    f"{d.C_max_fluid:.6e}\n")
        with open(os.path.join(cfg.output_dir, "mass_loss.csv"), "a") as f:
            f.write(f"{t_corr / 3600.0:.6f},{loss:.6f}\n")

    def _snapshot(self, grid, fields, cfg, prefix: str, t: float, series: VTKWriter, count: bool = True) -> None:
        if not self.write_vti:
            return
        fname = f"{cfg.output_dir}/{prefix}_{self.frame_count:06d}_t{t:.1f}s.vti"   # src/coupling.cpp:10-18
        series.write(fname, grid, fields, cfg)
        series.add_timestep(t, fname)
        if count:
            self.frame_count += 1

    def run(self, grid: Grid, fields: Fields, cfg: Config) -> float:
        """One rank per GPU when grid.nranks > 1 (z-slabs, comm_init done by the caller): every rank runs
        this loop, the reductions inside the library span all ranks, rank 0 writes the files."""
        cfg.check_supported()
        root = grid.rank == 0
        if grid.nranks > 1:
            if self.write_vti:
                raise ValueError("CoupledSolver: VTI snapshots are written by single-GPU runs only")
            if not root:
                self.log = lambda *a, **k: None
        if root:
            os.makedirs(cfg.output_dir, exist_ok=True)
            self.writer.set_pvd_path(cfg.output_dir + "/simulation.pvd")
            self.flow_writer.set_pvd_path(cfg.output_dir + "/flow.pvd")
            with open(os.path.join(cfg.output_dir, "diagnostics.csv"), "w") as f:
                f.write("time_s,time_h,pin_mass_loss_pct,solid_nodes,v_max,C_max_fluid\n")
            with open(os.path.join(cfg.output_dir, "mass_loss.csv"), "w") as f:
                f.write("time_h,pin_mass_loss_pct\n")
        self.initial_solid_indices = np.nonzero(grid.node_type_all == _l.SOLID_MG)[0].astype(np.int32)
        n0 = len(self.initial_solid_indices)
        self.log(f"Initial solid nodes: {n0}")
        self.flow_solver.init(grid, cfg)
        self.ard_solver.init(grid, cfg)
        implicit = bool(cfg.use_implicit)
        if implicit and grid.nranks > 1:
            raise ValueError("CoupledSolver: the implicit branch runs on single-GPU grids only")
        self.log("Using IMPLICIT ARD solver (matrix-free GMRES)" if implicit else "Using EXPLICIT ARD solver")
        self._snapshot(grid, fields, cfg, "state", 0.0, self.writer)
        t_corr, cycle, need_flow_solve = 0.0, 0, True
        self.dissolved_since_flow = 0
        while t_corr < cfg.T_final:
            cycle += 1
            self.log(f"\n=== Coupling cycle {cycle}, t={t_corr:.1f} s ({t_corr / 3600.0:.2f} h) ===")
            if need_flow_solve:
                self.flow_solver.solve_steady(fields, grid, cfg, verbose=self.log is print and root)
                self.dissolved_since_flow = 0
                need_flow_solve = False
                self._snapshot(grid, fields, cfg, "flow", t_corr, self.flow_writer)
            vol_loss = 1.0 - self._solid_C_sum(fields) / (n0 + 1e-30)
            if implicit:   # src/coupling.cpp:154-216
                imp = self.ard_implicit_solver
                imp.set_volume_loss(max(vol_loss, 0.0), grid)
                imp.assemble(fields, grid, cfg)                     # once per coupling cycle
                implicit_step, t_cycle_start, dissolution = 0, t_corr, False
                while implicit_step < cfg.corrosion_steps_per_check and t_corr < cfg.T_final and not dissolution:
                    dt_impl = imp.compute_adaptive_dt(fields, grid, cfg)
                    apply_inlet_bc(fields, grid, cfg)
                    apply_outlet_bc(fields, grid, cfg)
                    apply_wall_concentration_bc(fields, grid, cfg)
                    imp.step(fields, grid, cfg, dt_impl)
                    smooth_boundary_concentration(fields, grid, cfg)
                    t_corr += dt_impl
                    implicit_step += 1
                    self.total_implicit_steps += 1
                    if self.total_implicit_steps % cfg.diagnostic_every == 0:
                        self.write_diagnostics(grid, fields, t_corr, cfg)
                    if self.total_implicit_steps % cfg.implicit_output_every == 0:
                        self._snapshot(grid, fields, cfg, "corr", t_corr, self.writer)
                    # any solid node below the threshold ends the cycle (src/coupling.cpp:206-211)
                    dissolution = self._any_solid_below(grid, fields, cfg)
                self.log(f"  Implicit cycle: {implicit_step} steps, t={t_cycle_start:.2f} to {t_corr:.2f} s "
                         f"({t_corr / 3600.0:.4f} h); last solve {imp.last.iters} iterations, "
                         f"|res| = {imp.last.rel_res:.2e}")
                n_dissolved = imp.apply_phase_change(fields, grid, cfg)
                n_dissolved = int(grid.allreduce([n_dissolved], "sum")[0])            # all ranks take the same branch
                self.total_dissolved += n_dissolved
                self.dissolved_since_flow += n_dissolved
                if n_dissolved > 0:
                    self.log(f"  Phase change: {n_dissolved} nodes dissolved (total: {self.total_dissolved})")
                    need_flow_solve = True
                else:
                    self.log("  No phase changes this cycle")
                if diagnostics(grid).solid_count == 0:
                    self.log(f"\n=== All solid nodes dissolved at t={t_corr:.1f} s ===")
                    break
                continue
            self.ard_solver.set_volume_loss(max(vol_loss, 0.0), grid)
            dt_corr = self.ard_solver.compute_dt(fields, grid, cfg)
            self.log(f"  Corrosion dt = {dt_corr:.4e} s")
            step, n_steps = 0, cfg.corrosion_steps_per_check
            while step < n_steps:
                # run up to the next output point in one device-resident call
                chunk = min(n_steps - step, cfg.output_every_corr - step % cfg.output_every_corr)
                # the reference stops the cycle as soon as t_corr >= T_final (coupling.cpp:251)
                done = 0
                t_probe = t_corr
                while done < chunk:
                    t_probe += dt_corr
                    done += 1
                    if t_probe >= cfg.T_final:
                        break
                self.ard_solver.iterate(fields, grid, cfg, done, dt_corr)
                for _ in range(done):
                    t_corr += dt_corr
                step += done
                if step % cfg.output_every_corr == 0:
                    self._snapshot(grid, fields, cfg, "corr", t_corr, self.writer)
                    self.write_diagnostics(grid, fields, t_corr, cfg)
                if t_corr >= cfg.T_final:
                    break
            n_dissolved = self.ard_solver.apply_phase_change(fields, grid, cfg)   # this rank's slab
            n_dissolved = int(grid.allreduce([n_dissolved], "sum")[0])            # all ranks take the same branch
            self.total_dissolved += n_dissolved
            self.dissolved_since_flow += n_dissolved
            if n_dissolved > 0:
                self.log(f"  Phase change: {n_dissolved} nodes dissolved (total: {self.total_dissolved})")
                need_flow_solve = True
            else:
                self.log("  No phase changes this cycle")
            if diagnostics(grid).solid_count == 0:
                self.log(f"\n=== All solid nodes dissolved at t={t_corr:.1f} s ===")
                break
        self._snapshot(grid, fields, cfg, "final", t_corr, self.writer, count=False)
        self.log("\n=== Simulation complete ===")
        return t_corr
