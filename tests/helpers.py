"""Shared test helpers: configs, oracle construction, parity metrics."""
from __future__ import annotations

import os

import numpy as np

from oracle import refapi
from oracle.portapi import PortSim
from pd_mg_pin_corrosion_b200.config import Config

CONFIG_DIR = refapi.CONFIG_DIR

# name -> (dim, cfg file, overrides)
CASES = {
    "2d_default": (2, "params.cfg", {}),
    "2d_poiseuille": (2, "params_poiseuille.cfg", {}),
    "3d_default": (3, "params.cfg", {}),
    # small 3D tube (fast everywhere): same physics, shorter/narrower
    "3d_small": (3, "params.cfg", {"R_tube": 60e-6, "R_wire": 20e-6, "L_wire": 60e-6, "L_upstream": 40e-6,
                                   "L_downstream": 40e-6, "grain_size_mean": 20e-6}),
    # off-lattice geometry: radii/lengths that are not multiples of dx
    "2d_offgrid": (2, "params.cfg", {"dx": 4.3e-6, "R_tube": 101e-6, "R_wire": 33e-6, "L_wire": 207e-6,
                                     "L_upstream": 111e-6, "L_downstream": 93e-6}),
    "3d_offgrid": (3, "params.cfg", {"dx": 4.3e-6, "R_tube": 51e-6, "R_wire": 17e-6, "L_wire": 47e-6,
                                     "L_upstream": 31e-6, "L_downstream": 29e-6, "grain_size_mean": 15e-6}),
    # phase change fires within a few hundred explicit steps
    "2d_dissolve": (2, "params.cfg", {"D_grain": 5e-11, "D_gb": 5e-9, "C_thresh": 0.999,
                                      "corrosion_steps_per_check": 50, "flow_max_iters": 300, "T_final": 6e-4,
                                      "output_every_corr": 10}),
    # the same in 3D on the small tube: whole-run parity with phase change + table rebuilds in 3D
    "3d_dissolve": (3, "params.cfg", {"R_tube": 60e-6, "R_wire": 20e-6, "L_wire": 60e-6, "L_upstream": 40e-6,
                                      "L_downstream": 40e-6, "grain_size_mean": 20e-6, "D_grain": 5e-11, "D_gb": 5e-9,
                                      "Q_flow": 1.667e-10, "C_thresh": 0.99999, "corrosion_steps_per_check": 40,
                                      "flow_max_iters": 400, "T_final": 6e-3, "output_every_corr": 10}),
}


def load_cfg(case: str, extra: dict | None = None) -> tuple[int, Config, dict]:
    dim, base, ov = CASES[case]
    ov = dict(ov)
    ov["use_implicit"] = 0
    ov.update(extra or {})
    cfg = Config.load(os.path.join(CONFIG_DIR, base) if base else None, ov, quiet=True)
    return dim, cfg, ov


_REF_CACHE: dict = {}


def make_ref(case: str, extra: dict | None = None, threads: int = 4):
    """The compiled reference (oracle/_ref) if it was built, else the plain-C port.

    Instances of the plain geometry cases are cached per session (the 3D reference CSR
    takes ~20 s to build); their fields are re-initialised on every request."""
    dim, base, ov = CASES[case]
    ov = dict(ov)
    ov.update(extra or {})
    cacheable = not extra and case != "2d_dissolve"
    if cacheable and case in _REF_CACHE:
        r = _REF_CACHE[case]
        if isinstance(r, refapi.RefSim):
            r.lib.ref_fields_init(r.h)
        else:
            r.init_fields(r.is_gb.copy(), r.is_precip.copy())
        return r
    if refapi.have_ref(dim):
        r = refapi.RefSim(dim, base, ov, threads=threads)
    else:
        r = make_port(case, extra, threads)
        r.init_fields()
    if cacheable:
        _REF_CACHE[case] = r
    return r


def make_port(case: str, extra: dict | None = None, threads: int = 4, state_from=None) -> PortSim:
    dim, cfg, _ = load_cfg(case, extra)
    p = PortSim(dim, cfg, threads=threads)
    if state_from is not None:
        p.load_state(state_from)
    return p


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """max_i |a_i - b_i| / max_i |b_i|  (SURVEY.md 7.3 field tolerance)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = float(np.max(np.abs(a - b))) if a.size else 0.0
    m = float(np.max(np.abs(b))) if b.size else 0.0
    return d / m if m > 0 else d


def perturbed_state(ref, seed: int = 0):
    """Deterministic non-trivial state written into the reference object: Poiseuille + noise."""
    rng = np.random.default_rng(seed)
    nt = ref.get("node_type") if hasattr(ref, "get") else ref.node_type
    N, dim = nt.size, ref.dim
    inside = (nt != 5)
    rho = ref.get("rho") * (1.0 + 1e-4 * rng.standard_normal(N)) * inside
    vel = (ref.get("vel") * (1.0 + 1e-2 * rng.standard_normal((N, dim)))
           + 1e-3 * rng.standard_normal((N, dim))) * inside[:, None]
    C = np.abs(ref.get("C") + 0.02 * rng.standard_normal(N)) * inside
    ref.set("rho", rho)
    ref.set("vel", vel)
    ref.set("C", C)
    ref.set("rho_new", rho)
    ref.set("vel_new", vel)
    ref.set("C_new", C)
    return rho, vel, C


def coupled_run(sim, cfg, log=None):
    """CoupledSolver::run, explicit branch (src/coupling.cpp:82-302), restated over any object
    with the RefSim/PortSim operator surface. Returns the diagnostics.csv rows (already
    rounded through the reference's `%.6e` formatting)."""
    nt0 = sim.get("node_type")
    solid0 = np.nonzero(nt0 == 1)[0]
    n0 = len(solid0)
    rows = []

    def solid_sum():
        s = 0.0
        for v in sim.get("C")[solid0].tolist():   # ordered sum, src/coupling.cpp:32-38
            s += v
        return s

    def diagnostics(t):
        nt = sim.get("node_type")
        loss = (1.0 - solid_sum() / (n0 + 1e-30)) * 100.0
        loss = max(loss, 0.0)
        fl = nt == 0
        v = sim.get("vel")[fl]
        vmax = float(np.sqrt((v * v).sum(1)).max()) if fl.any() else 0.0
        cmax = float(max(sim.get("C")[fl].max(), 0.0)) if fl.any() else 0.0
        vals = [t, t / 3600.0, loss, float((nt == 1).sum()), vmax, cmax]
        rows.append([float(f"{x:.6e}") for x in vals])

    def solve_steady():   # src/pd_ns.cpp:182-372
        dt = sim.ns_compute_dt()
        it = 1
        while it <= cfg.flow_max_iters:
            sim.inlet_bc(); sim.outlet_bc(); sim.wall_bc(); sim.solid_bc()
            sim.ns_step(dt)
            sim.wall_bc_new()
            if it <= 10 or it % 100 == 0:
                nt = sim.get("node_type")
                fl = nt == 0
                v, vn, rn = sim.get("vel")[fl], sim.get("vel_new")[fl], sim.get("rho_new")[fl]
                if np.isnan(vn[:, 0]).any() or np.isnan(rn).any():
                    break
                num, den = float(((vn - v) ** 2).sum()), float((v ** 2).sum())
                eps = np.sqrt(num / den) if den > 1e-30 else np.sqrt(num)
                if np.sqrt((vn ** 2).sum(1)).max() > 100.0 * cfg.U_in:
                    break
                if eps < cfg.flow_conv_tol and it > 100:
                    break
            sim.swap_flow()
            if it % 200 == 0:
                dt = sim.ns_compute_dt()
            it += 1
        return it

    t_corr, need_flow = 0.0, True
    while t_corr < cfg.T_final:
        if need_flow:
            solve_steady()
            need_flow = False
        vl = max(1.0 - solid_sum() / (n0 + 1e-30), 0.0)
        sim.ard_set_volume_loss(vl)
        dtc = sim.ard_compute_dt()
        for step in range(1, cfg.corrosion_steps_per_check + 1):
            sim.inlet_bc(); sim.outlet_bc(); sim.wall_conc_bc()
            sim.ard_step(dtc)
            sim.swap_C()
            t_corr += dtc
            if step % cfg.output_every_corr == 0:
                diagnostics(t_corr)
            if t_corr >= cfg.T_final:
                break
        n = sim.phase_change()
        if n > 0:
            sim.rebuild_neighbors()
            need_flow = True
        if (sim.get("node_type") == 1).sum() == 0:
            break
    return rows


def coupled_run_implicit(sim, cfg, is_gb, is_precip, log=None):
    """CoupledSolver::run, IMPLICIT branch (src/coupling.cpp:154-216) over the plain-C port (flow solve,
    BCs, smoother, phase change) and the numpy/scipy restatement of PD_ARD_ImplicitSolver with the exact
    sparse solve in place of Eigen's GMRES. Returns the diagnostics.csv rows."""
    from oracle.implicit_oracle import ImplicitOracle
    nt0 = sim.get("node_type")
    solid0 = np.nonzero(nt0 == 1)[0]
    n0 = len(solid0)
    rows = []

    def solid_sum():
        s = 0.0
        for v in sim.get("C")[solid0].tolist():
            s += v
        return s

    def diagnostics(t):
        nt = sim.get("node_type")
        loss = max((1.0 - solid_sum() / (n0 + 1e-30)) * 100.0, 0.0)
        fl = nt == 0
        v = sim.get("vel")[fl]
        vmax = float(np.sqrt((v * v).sum(1)).max()) if fl.any() else 0.0
        cmax = float(max(sim.get("C")[fl].max(), 0.0)) if fl.any() else 0.0
        rows.append([float(f"{x:.6e}") for x in [t, t / 3600.0, loss, float((nt == 1).sum()), vmax, cmax]])

    def solve_steady():
        dt = sim.ns_compute_dt()
        it = 1
        while it <= cfg.flow_max_iters:
            sim.inlet_bc(); sim.outlet_bc(); sim.wall_bc(); sim.solid_bc()
            sim.ns_step(dt)
            sim.wall_bc_new()
            if it <= 10 or it % 100 == 0:
                fl = sim.get("node_type") == 0
                v, vn = sim.get("vel")[fl], sim.get("vel_new")[fl]
                num, den = float(((vn - v) ** 2).sum()), float((v ** 2).sum())
                eps = np.sqrt(num / den) if den > 1e-30 else np.sqrt(num)
                if eps < cfg.flow_conv_tol and it > 100:
                    break
            sim.swap_flow()
            if it % 200 == 0:
                dt = sim.ns_compute_dt()
            it += 1

    t_corr, need_flow, total_steps = 0.0, True, 0
    while t_corr < cfg.T_final:
        if need_flow:
            solve_steady()
            need_flow = False
        orc = ImplicitOracle(sim.dim, sim.Nx, sim.Ny, sim.Nz, sim.get("node_type"), sim.off_d, sim.off_dist,
                             sim.off_evec, sim.off_vol, cfg)
        orc.volume_loss = max(1.0 - solid_sum() / (n0 + 1e-30), 0.0)
        orc.assemble(sim.get("C"), sim.get("vel"), is_gb, is_precip)
        step, dissolved = 0, False
        while step < cfg.corrosion_steps_per_check and t_corr < cfg.T_final and not dissolved:
            dt = orc.adaptive_dt(sim.get("C"), cfg.implicit_dt_fraction, cfg.implicit_dt_max)
            sim.inlet_bc(); sim.outlet_bc(); sim.wall_conc_bc()
            sim.set("C", orc.step(sim.get("C"), dt))
            sim.smooth_conc()
            t_corr += dt
            step += 1
            total_steps += 1
            if total_steps % cfg.diagnostic_every == 0:
                diagnostics(t_corr)
            nt = sim.get("node_type")
            dissolved = bool(((nt == 1) & (sim.get("C") < cfg.C_thresh)).any())
        n = sim.phase_change()
        if n > 0:
            sim.rebuild_neighbors()
            need_flow = True
        if (sim.get("node_type") == 1).sum() == 0:
            break
    return rows


def port_coupled_run(case: str, is_gb, is_precip):
    dim, cfg, _ = load_cfg(case)
    p = PortSim(dim, cfg, threads=4)
    p.init_fields(is_gb, is_precip)
    return coupled_run(p, cfg)


def synthetic_state(N: int, dim: int) -> dict:
    """Deterministic, platform-independent field values (integer hashing + exact IEEE ops only) that
    exercise every branch of the "%g" formatter: magnitudes 1e-20..1e+20, signs, zeros, exact ties,
    values below the 1e-300 flush threshold, NaN/Inf. Used for the VTI golden hash."""
    i = np.arange(N, dtype=np.int64)
    pw = np.array([float(f"1e{k}") for k in range(-20, 21)])

    def field(salt: int) -> np.ndarray:
        h = (i * 2654435761 + salt * 40503) % 1000003
        v = (h.astype(np.float64) / 1000003.0) * pw[(i + salt) % 41]
        v[(i + salt) % 7 == 0] *= -1.0
        v[(i + salt) % 53 == 0] = 0.0
        v[(i + salt) % 101 == 0] = ((i[(i + salt) % 101 == 0] % 900000) + 100000).astype(np.float64) + 0.5   # ties
        v[(i + salt) % 211 == 0] = 1e-305
        return v

    vel = np.stack([field(1 + d) for d in range(dim)], axis=1)
    C = field(5)
    C[i % 997 == 0] = np.nan
    C[i % 991 == 0] = np.inf
    # rho / rho_f in {1/2, 1, 2}: the EOS power is exact in every libm, so `pressure` cannot differ
    rho = np.array([500.0, 1000.0, 2000.0])[i % 3]
    return {"rho": rho, "vel": np.ascontiguousarray(vel), "C": C,
            "phase": (i % 2).astype(np.uint8), "is_gb": (i % 5 == 0).astype(np.uint8),
            "is_precip": (i % 11 == 0).astype(np.uint8), "grain_id": ((i * 7919) % 60 - 1).astype(np.int32),
            "D_map": np.abs(field(6)) * 1e-9}
