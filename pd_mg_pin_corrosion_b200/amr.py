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

    # ---- implicit branch (PD_ARD_ImplicitSolver with use_amr, src/pd_ard_implicit.cpp) ----------
    def smooth_conc(self): _l.check(_l.load().pdamr_bc(self.ctx, 6))

    def implicit_assemble(self) -> None:
        _l.check(_l.load().pdamr_implicit_assemble(self.ctx))

    def implicit_matvec(self, dt: float, x) -> np.ndarray:
        """y = (I - dt M) x with the fictitious-node coupling rows; one entry per node"""
        x = np.ascontiguousarray(x, np.float64)
        assert x.size == self.N_total
        y = np.zeros(self.N_total)
        _l.check(_l.load().pdamr_implicit_matvec(self.ctx, dt, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p)))
        return y

    def implicit_rhs(self, dt: float) -> np.ndarray:
        b = np.zeros(self.N_total)
        _l.check(_l.load().pdamr_implicit_rhs(self.ctx, dt, b.ctypes.data_as(C.c_void_p)))
        return b

    def implicit_compute_dt(self, dt_fraction: float | None = None, dt_max: float | None = None) -> float:
        dt = C.c_double()
        _l.check(_l.load().pdamr_implicit_compute_dt(
            self.ctx, self.cfg.implicit_dt_fraction if dt_fraction is None else dt_fraction,
            self.cfg.implicit_dt_max if dt_max is None else dt_max, C.byref(dt)))
        return dt.value

    def implicit_step(self, dt: float, tol: float = 1e-10, restart: int = 50, max_iters: int = 200, precond: int = 1):
        """one backward-Euler step on the current C buffer; returns PdLinSolveInfo (iters, converged, rel_res)"""
        info = _l.PdLinSolveInfo()
        _l.check(_l.load().pdamr_implicit_step(self.ctx, dt, tol, restart, max_iters, precond, C.byref(info)))
        return info

    def close(self) -> None:
        if self.ctx:
            _l.load().pdamr_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def generate_grains(grid: AmrGrid, seed: int = 42):
    """GrainStructure::generate (src/grains.cpp:9-179) on the cloud: the reference's generator works on grid.pos and
    the CSR, so it runs unchanged with use_amr = 1; same passes and libstdc++ RNG calls here (host/grains.cpp).
    Returns (grain_id, is_grain_boundary, is_precipitate, n_grains)."""
    from . import grains as _g
    L = _g._load()
    if not hasattr(L, "_cloud_bound"):
        L.pdhost_generate_grains_cloud.restype = C.c_int
        L.pdhost_generate_grains_cloud.argtypes = [C.c_void_p,
                                                   C.c_double, C.c_double, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5 + \
                                                  [C.c_int] + [C.c_void_p] * 3 + [C.POINTER(C.c_int)]
        L._cloud_bound = True
    cfg = grid.cfg
    N = grid.N_total
    pos, nt = grid.get("pos"), grid.get("node_type")
    off, idx, dist = grid.get("nbr_offset"), grid.get("nbr_index"), grid.get("nbr_dist")
    gid = np.full(N, -1, np.int32)
    gb = np.zeros(N, np.uint8)
    pr = np.zeros(N, np.uint8)
    n = C.c_int()
    s = cfg.to_struct()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = L.pdhost_generate_grains_cloud(C.byref(s), cfg.grain_size_mean, cfg.precip_fraction, int(cfg.gb_width_cells),
                                        int(cfg.precip_cluster_cells), N, p(pos), p(nt), p(off), p(idx), p(dist), seed,
                                        p(gid), p(gb), p(pr), C.byref(n))
    if rc != 0:
        raise RuntimeError("grain generation on the AMR cloud failed")
    return gid, gb, pr, n.value


def _host_lib():
    from . import grains as _g
    L = _g._load()
    if not getattr(L, "_vtu_ready", False):
        L.pdhost_write_vtu.restype = C.c_int
        L.pdhost_write_vtu.argtypes = [C.c_char_p, C.c_int] + [C.c_void_p] * 12
        L.pdhost_init_dmap.restype = None
        L.pdhost_init_dmap.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double,
                                       C.c_double, C.c_void_p]
        L._vtu_ready = True
    return L


def init_dmap(cfg: Config, node_type, is_gb, is_precip) -> np.ndarray:
    """Fields::D_map as initialize_fields sets it (src/main.cpp:19-112); host-side, only the writer reads it"""
    nt, gb, pr = (np.ascontiguousarray(a, np.uint8) for a in (node_type, is_gb, is_precip))
    out = np.zeros(nt.size)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    _host_lib().pdhost_init_dmap(nt.size, p(nt), p(gb), p(pr), cfg.D_liquid, cfg.D_grain, cfg.D_gb, cfg.D_precip, p(out))
    return out


def write_vtu_arrays(path: str, pos, node_type, vel, pressure, Cc, phase, grid_level, dx_local, grain_id, D_map, is_gb,
                     is_precip) -> None:
    """VTKWriter::write_vtu (src/vtk_writer.cpp:199-346) from host arrays: byte-identical text (host/vtu.cpp)"""
    arrs = [np.ascontiguousarray(pos, np.float64), np.ascontiguousarray(node_type, np.uint8),
            np.ascontiguousarray(vel, np.float64), np.ascontiguousarray(pressure, np.float64),
            np.ascontiguousarray(Cc, np.float64), np.ascontiguousarray(phase, np.uint8),
            None if grid_level is None else np.ascontiguousarray(grid_level, np.int32),
            None if dx_local is None else np.ascontiguousarray(dx_local, np.float64),
            np.ascontiguousarray(grain_id, np.int32), np.ascontiguousarray(D_map, np.float64),
            np.ascontiguousarray(is_gb, np.uint8), None if is_precip is None else np.ascontiguousarray(is_precip, np.uint8)]
    N = arrs[1].size
    ptrs = [None if a is None else a.ctypes.data_as(C.c_void_p) for a in arrs]
    if _host_lib().pdhost_write_vtu(path.encode(), N, *ptrs) != 0:
        raise OSError(f"cannot write {path}")


def write_vtu(path: str, grid: "AmrGrid", grain_id, D_map) -> None:
    """snapshot of the cloud's current device state"""
    write_vtu_arrays(path, grid.get("pos"), grid.get_field("node_type"), grid.get_field("vel"), grid.get_field("pressure"),
                     grid.get_field("C"), grid.get_field("phase"), grid.get("grid_level"), grid.get("dx_local"), grain_id,
                     D_map, grid.get_field("is_gb"), grid.get_field("is_precip"))


def initialize_fields(grid: AmrGrid, is_gb, is_precip) -> None:
    """initialize_fields (src/main.cpp:9-126) on the AMR cloud: Poiseuille profile on FLUID / INLET nodes,
    C = C_solid_init on the wire, FICTITIOUS nodes at rest; new buffers = current buffers."""
    cfg = grid.cfg
    nt = grid.get("node_type")
    x = grid.get("pos")[:, 0]
    N = grid.N_total
    rr = np.minimum((x * x) / (cfg.R_tube * cfg.R_tube), 1.0)
    v_ax = 1.5 * cfg.U_in * (1.0 - rr)
    rho = np.full(N, cfg.rho_f)
    rho[nt == 5] = 0.0
    vel = np.zeros((N, 2))
    flow = (nt == 0) | (nt == 3)
    vel[flow, 1] = v_ax[flow]
    Cc = np.zeros(N)
    Cc[nt == 0] = cfg.C_liquid_init
    Cc[nt == 3] = cfg.C_liquid_init
    Cc[nt == 4] = cfg.C_liquid_init
    Cc[nt == 1] = cfg.C_solid_init
    phase = np.ones(N, np.uint8)
    phase[nt == 1] = 0
    for name, val in (("rho", rho), ("rho_new", rho), ("vel", vel), ("vel_new", vel), ("C", Cc), ("C_new", Cc),
                      ("phase", phase), ("is_gb", np.asarray(is_gb, np.uint8)),
                      ("is_precip", np.asarray(is_precip, np.uint8))):
        grid.set_field(name, val)


class AmrCoupledSolver:
    """CoupledSolver::run with use_amr = 1, explicit or implicit ARD branch (src/coupling.cpp:82-302): flow solve when the
    geometry changed + IDW refresh of the FICTITIOUS nodes (:138-139), corrosion sub-steps with the frozen flow,
    phase change, diagnostics rows (:20-68) every output_every_corr steps.  Returns the rows; writes
    <output_dir>/diagnostics.csv when `out_dir` is given, and with `grain_id` (snapshots need the host-side
    grain ids) the state_/flow_/corr_/final_ VTU series + simulation.pvd / flow.pvd the reference writes
    (:117-121,142-147,242-246,292-296)."""

    def __init__(self, log=None):
        self.log = log or (lambda *a, **k: None)
        self.rows: list[list[float]] = []
        self.frame_count = 0
        self.total_implicit_steps = 0
        self.tol, self.restart, self.max_iters = 1e-10, 50, 200      # src/pd_ard_implicit.cpp:400-402
        self.last = None

    def _snapshot(self, grid, out_dir, prefix, t, series, count=True):
        fname = f"{out_dir}/{prefix}_{self.frame_count:06d}_t{t:.1f}s.vtu"      # make_filename (:10-18)
        write_vtu(fname, grid, self._grain_id, self._D_map)
        series.add_timestep(t, fname)
        if count:
            self.frame_count += 1

    def _diag(self, grid: AmrGrid, t: float, solid0: np.ndarray) -> None:
        nt = grid.get_field("node_type")
        Cc = grid.get_field("C")
        s = 0.0
        for v in Cc[solid0].tolist():          # ordered sum (src/coupling.cpp:32-38)
            s += v
        loss = max((1.0 - s / (len(solid0) + 1e-30)) * 100.0, 0.0)
        fl = nt == 0
        v = grid.get_field("vel")[fl]
        vmax = float(np.sqrt((v * v).sum(1)).max()) if fl.any() else 0.0
        cmax = float(max(Cc[fl].max(), 0.0)) if fl.any() else 0.0
        self.rows.append([t, t / 3600.0, loss, float((nt == 1).sum()), vmax, cmax])

    def _implicit_cycle(self, grid, cycle: int, t_corr: float, solid0, snaps) -> float:
        """phase 2, implicit (src/coupling.cpp:154-216): assemble once, then adaptive backward-Euler steps until
        corrosion_steps_per_check, T_final or the first solid node below C_thresh"""
        cfg = grid.cfg
        grid.implicit_assemble()
        step, dissolved = 0, False
        while step < cfg.corrosion_steps_per_check and t_corr < cfg.T_final and not dissolved:
            dt_impl = grid.implicit_compute_dt()
            grid.inlet_bc(); grid.outlet_bc(); grid.wall_conc_bc()
            self.last = grid.implicit_step(dt_impl, tol=self.tol, restart=self.restart, max_iters=self.max_iters)
            grid.smooth_conc()
            grid.update_fictitious()
            t_corr += dt_impl
            step += 1
            self.total_implicit_steps += 1
            if self.total_implicit_steps % int(cfg.diagnostic_every) == 0:
                self._diag(grid, t_corr, solid0)
            if snaps and self.total_implicit_steps % int(cfg.implicit_output_every) == 0:
                self._snapshot(grid, snaps[0], "corr", t_corr, snaps[1])
            dissolved = bool(((grid.get_field("node_type") == 1) & (grid.get_field("C") < cfg.C_thresh)).any())
        self.log(f"cycle {cycle}: {step} implicit steps to t = {t_corr:.4e} s; last solve {self.last.iters} "
                 f"iterations, |res| = {self.last.rel_res:.2e}")
        return t_corr

    def _explicit_cycle(self, grid, cycle: int, t_corr: float, solid0, snaps) -> float:
        """phase 2, explicit (src/coupling.cpp:217-253): device-resident batches of sub-steps between output points"""
        cfg = grid.cfg
        dtc = grid.ard_compute_dt()
        step, every = 0, int(cfg.output_every_corr)
        while step < cfg.corrosion_steps_per_check:
            n = min(every - step % every, cfg.corrosion_steps_per_check - step)
            done = 0
            while done < n:                      # t_corr advances per step (:237) and ends the cycle (:248)
                t_corr += dtc
                done += 1
                if t_corr >= cfg.T_final:
                    break
            grid.ard_iterate(done, dtc)
            step += done
            if done == n and step % every == 0:
                if snaps:
                    self._snapshot(grid, snaps[0], "corr", t_corr, snaps[1])
                self._diag(grid, t_corr, solid0)
            if t_corr >= cfg.T_final:
                break
        return t_corr

    def run(self, grid: AmrGrid, out_dir: str | None = None, grain_id=None) -> list[list[float]]:
        cfg = grid.cfg
        solid0 = np.nonzero(grid.get_field("node_type") == 1)[0]
        n0 = len(solid0)
        t_corr, need_flow, cycle = 0.0, True, 0
        snap = out_dir is not None and grain_id is not None
        writer = flow_writer = None
        if snap:
            import os
            from .solver import VTKWriter
            os.makedirs(out_dir, exist_ok=True)
            self._grain_id = grain_id
            self._D_map = init_dmap(cfg, grid.get_field("node_type"), grid.get_field("is_gb"), grid.get_field("is_precip"))
            writer, flow_writer = VTKWriter(), VTKWriter()
            writer.set_pvd_path(out_dir + "/simulation.pvd")
            flow_writer.set_pvd_path(out_dir + "/flow.pvd")
            self._snapshot(grid, out_dir, "state", 0.0, writer)
        while t_corr < cfg.T_final:
            cycle += 1
            if need_flow:
                r = grid.ns_solve_steady()
                grid.update_fictitious()
                self.log(f"cycle {cycle}: flow solve {r.iters} iterations, eps {r.eps:.3e}")
                need_flow = False
                if snap:
                    self._snapshot(grid, out_dir, "flow", t_corr, flow_writer)
            s = 0.0
            for v in grid.get_field("C")[solid0].tolist():
                s += v
            grid.ard_set_volume_loss(max(1.0 - s / (n0 + 1e-30), 0.0))
            snaps = (out_dir, writer) if snap else None
            t_corr = (self._implicit_cycle if cfg.use_implicit else self._explicit_cycle)(grid, cycle, t_corr, solid0, snaps)
            before = grid.get_field("node_type") if snap else None
            n_diss = grid.phase_change()
            if n_diss > 0:
                need_flow = True
                if snap:                         # apply_phase_change: D_map = D_liquid (src/pd_ard.cpp:202)
                    self._D_map[(before == 1) & (grid.get_field("node_type") == 0)] = cfg.D_liquid
            self.log(f"cycle {cycle}: t = {t_corr:.4e} s, {n_diss} nodes dissolved")
            if (grid.get_field("node_type") == 1).sum() == 0:
                break
        if snap:
            self._snapshot(grid, out_dir, "final", t_corr, writer, count=False)
        if out_dir is not None:
            import os
            os.makedirs(out_dir, exist_ok=True)
            with open(os.path.join(out_dir, "diagnostics.csv"), "w") as f:
                f.write("time_s,time_h,pin_mass_loss_pct,solid_nodes,v_max,C_max_fluid\n")
                for row in self.rows:
                    f.write(f"{row[0]:.6e},{row[1]:.6e},{row[2]:.6e},{int(row[3])},{row[4]:.6e},{row[5]:.6e}\n")
        return self.rows
