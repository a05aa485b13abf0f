"""Implicit ARD branch on the device (src/pd_ard_implicit.cpp) against the numpy / scipy
restatement oracle/implicit_oracle.py (anchored on the reference's own tests/test_implicit.cpp Test 1 in
tests/test_implicit_oracle.py).  Operator, right-hand side and adaptive step: 1e-12.  Linear solve: against the
exact sparse solution of the same system, to the solver tolerance (the reference's own Eigen solve is
"unpinned": Eigen is absent from the reference tree)."""
import numpy as np
import pytest

import helpers as H
from oracle.implicit_oracle import ImplicitOracle

pytestmark = pytest.mark.gpu


def _device_and_oracle(case, extra, ns_iters, seed):
    """device context + oracle on the same state: flow after `ns_iters` NS loop bodies, perturbed C"""
    from pd_mg_pin_corrosion_b200 import solver as S
    from pd_mg_pin_corrosion_b200.grains import GrainStructure
    dim, cfg, _ = H.load_cfg(case, extra)
    grid = S.Grid(dim)
    grid.build(cfg)
    nt = grid.node_type
    grains = GrainStructure().generate(nt, cfg, dim)
    fields = S.Fields()
    fields.allocate(grid.N_total, grid)
    S.initialize_fields(fields, grid, grains, cfg)
    if ns_iters:
        ns = S.PD_NS_Solver()
        ns.init(grid, cfg)
        ns.iterate(fields, grid, cfg, ns_iters, ns.compute_dt(fields, grid, cfg))
    rng = np.random.default_rng(seed)
    C = fields.get("C")
    C = np.abs(C + 0.05 * rng.standard_normal(C.size)) * (nt != 5)
    C[(nt == 0) & (np.arange(C.size) % 41 == 0)] = 0.95            # some saturated fluid: salt layer on a few solids
    fields.set("C", C)
    d, dist, evec, vol = grid.stencil()
    orc = ImplicitOracle(dim, grid.Nx, grid.Ny, grid.Nz, nt, d, dist, evec, vol, cfg)
    imp = S.PD_ARD_ImplicitSolver()
    imp.set_volume_loss(0.013, grid)
    orc.volume_loss = 0.013
    imp.assemble(fields, grid, cfg)
    orc.assemble(C, fields.get("vel"), grains.is_grain_boundary, grains.is_precipitate)
    return S, cfg, grid, fields, imp, orc, C, nt


@pytest.mark.parametrize("case,extra,iters", [("2d_dissolve", {"corrosion_decay_l": 0.1}, 300), ("3d_small", None, 40),
                                              ("2d_offgrid", None, 100)])
def test_operator_rhs_and_adaptive_dt(case, extra, iters):
    S, cfg, grid, fields, imp, orc, C, nt = _device_and_oracle(case, extra, iters, seed=3)
    rng = np.random.default_rng(4)
    unknown = (nt == 0) | (nt == 1)
    for dt in (1e-3, 0.7):
        x = rng.standard_normal(grid.N_total) * unknown
        y = imp.matvec(grid, dt, x)
        A, b = orc.system(C, dt)
        want = np.zeros(grid.N_total)
        want[orc.l2g] = A @ x[orc.l2g]
        assert H.rel_err(y, want) <= 1e-12, (dt, "matvec")
        bw = np.zeros(grid.N_total)
        bw[orc.l2g] = b
        assert H.rel_err(imp.rhs(grid, dt), bw) <= 1e-12, (dt, "rhs")
    for frac, dmax in ((0.5, 60.0), (0.5, 1e-3)):
        cfg.implicit_dt_fraction, cfg.implicit_dt_max = frac, dmax
        got, want = imp.compute_adaptive_dt(fields, grid, cfg), orc.adaptive_dt(C, frac, dmax)
        assert abs(got - want) <= 1e-10 * want, (frac, dmax, got, want)


@pytest.mark.parametrize("precond", [1, 2])
def test_reference_test1_pure_diffusion_on_device(precond):
    """tests/test_implicit.cpp Test 1 on the device: every step against the exact sparse solve, and the
    reference's assertions at the end (L2 <= 0.05 against the analytical Gaussian, mass change <= 1 %)."""
    from test_implicit_oracle import TEST_CFG, gaussian, l2
    from pd_mg_pin_corrosion_b200 import solver as S
    from pd_mg_pin_corrosion_b200.config import Config
    cfg = Config.load(None, dict(TEST_CFG), quiet=True)
    grid = S.Grid(2)
    grid.build(cfg)
    nt = grid.node_type
    fields = S.Fields()
    fields.allocate(grid.N_total, grid)
    S.initialize_fields(fields, grid, None, cfg)
    fields.set("vel", np.zeros((grid.N_total, 2)))
    i = np.arange(grid.N_total) % grid.Nx
    j = np.arange(grid.N_total) // grid.Nx
    x, y = grid.origin_x + i * cfg.dx, grid.origin_y + j * cfg.dx
    sigma, D, t_end, dt = 30e-6, 1.0e-9, 0.5, 0.01
    C0 = gaussian(x, y, nt, sigma)
    fields.set("C", C0)
    d, dist, evec, vol = grid.stencil()
    orc = ImplicitOracle(2, grid.Nx, grid.Ny, grid.Nz, nt, d, dist, evec, vol, cfg)
    zeros = np.zeros(grid.N_total, np.uint8)
    orc.assemble(C0, np.zeros((grid.N_total, 2)), zeros, zeros)
    imp = S.PD_ARD_ImplicitSolver(tol=1e-12, restart=50, max_iters=2000, precond=precond)
    imp.assemble(fields, grid, cfg)
    C_ref, t, worst = C0.copy(), 0.0, 0.0
    for _ in range(50):
        imp.step(fields, grid, cfg, dt)
        assert imp.last.converged and imp.last.rel_res <= 1e-12, (imp.last.iters, imp.last.rel_res)
        C_ref = orc.step(C_ref, dt)
        worst = max(worst, H.rel_err(fields.get("C"), C_ref))
        t += dt
    assert worst <= 1e-9, worst
    C = fields.get("C")
    assert l2(C, gaussian(x, y, nt, sigma, D, t_end), nt) <= 0.05
    assert abs(C[nt == 0].sum() - C0[nt == 0].sum()) / C0[nt == 0].sum() * 100.0 <= 1.0


@pytest.mark.parametrize("case,extra,iters,dt", [("2d_dissolve", None, 400, 0.05), ("2d_dissolve", None, 400, 5.0),
                                                 ("3d_small", None, 60, 0.05)])
def test_step_with_flow_matches_exact_solve(case, extra, iters, dt):
    """advection-dominated system (flow around the wire): GMRES with the axial-sweep preconditioner against
    the exact sparse solution, including the clamp to [0, C_solid_init]"""
    S, cfg, grid, fields, imp, orc, C, nt = _device_and_oracle(case, extra, iters, seed=9)
    imp.tol, imp.max_iters, imp.precond = 1e-11, 4000, 2
    want = orc.step(C, dt)
    imp.step(fields, grid, cfg, dt)
    assert imp.last.converged, (imp.last.iters, imp.last.rel_res)
    assert H.rel_err(fields.get("C"), want) <= 1e-8, (imp.last.iters, imp.last.rel_res)


def test_whole_implicit_coupled_run(tmp_path):
    """solver.CoupledSolver.run with use_implicit = 1 (src/coupling.cpp:154-216: assemble per cycle, adaptive
    dt, BCs, implicit step, smoother, diagnostics every step, cycle ends at the first solid below C_thresh,
    phase change, flow re-solve) against the CPU restatement of the same loop (plain-C port + exact sparse
    solve): solid counts exact, every numeric diagnostics column within 1e-6."""
    from oracle.portapi import PortSim
    from pd_mg_pin_corrosion_b200 import solver as S
    from pd_mg_pin_corrosion_b200.grains import GrainStructure
    extra = {"use_implicit": 1, "D_grain": 5e-11, "D_gb": 5e-9, "C_thresh": 0.999, "corrosion_steps_per_check": 6,
             "flow_max_iters": 300, "T_final": 1.2e-3, "implicit_dt_max": 0.004, "implicit_dt_fraction": 0.5,
             "diagnostic_every": 1, "output_dir": str(tmp_path / "out")}
    dim, cfg, _ = H.load_cfg("2d_default", extra)
    cfg.use_implicit = 1
    grid = S.Grid(dim)
    grid.build(cfg)
    grains = GrainStructure().generate(grid.node_type, cfg, dim)
    fields = S.Fields()
    fields.allocate(grid.N_total, grid)
    S.initialize_fields(fields, grid, grains, cfg)
    cs = S.CoupledSolver()
    cs.log = lambda *a, **k: None
    cs.ard_implicit_solver = S.PD_ARD_ImplicitSolver(tol=1e-12, restart=50, max_iters=4000, precond=2)
    cs.run(grid, fields, cfg)
    got = np.loadtxt(str(tmp_path / "out" / "diagnostics.csv"), delimiter=",", skiprows=1, ndmin=2)
    port = PortSim(dim, cfg, threads=4)
    port.init_fields(grains.is_grain_boundary, grains.is_precipitate)
    want = np.array(H.coupled_run_implicit(port, cfg, grains.is_grain_boundary, grains.is_precipitate))
    assert got.shape == want.shape and got.shape[0] >= 5, (got.shape, want.shape)
    assert np.array_equal(got[:, 3], want[:, 3])
    assert cs.total_dissolved > 0
    for col in (0, 1, 2, 4, 5):
        rel = np.abs(got[:, col] - want[:, col]) / np.maximum(np.abs(want[:, col]), 1e-300)
        assert rel.max() <= 1e-6, (col, float(rel.max()))
