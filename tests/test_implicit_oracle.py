"""CPU tests of the implicit-branch restatement (oracle/implicit_oracle.py).

The reference's implicit solver cannot be built here (Eigen 3.4.0 is fetched at configure time), so the
restatement is anchored on the reference's OWN test for that branch: tests/test_implicit.cpp Test 1 (pure PD
diffusion of a Gaussian pulse, :167-327) with its set-up (:25-61, :98-126) and its hard assertions (:307-322):
L2 error against the analytical Gaussian <= 0.05 at the finest step, mass change <= 1 %, a convergence rate
> 0.4.  Structural checks: M has zero row sums on rows without INLET/OUTLET neighbours (conservation) and
non-negative off-diagonals (the per-bond upwind stabilisation, :272-282)."""
import numpy as np
import pytest

import helpers as H
from oracle.implicit_oracle import ImplicitOracle
from oracle.portapi import PortSim
from pd_mg_pin_corrosion_b200.config import Config

TEST_CFG = dict(dx=5.0e-6, m_ratio=3, R_wire=0.0, L_wire=0.0, R_tube=200.0e-6, L_upstream=300.0e-6,
                L_downstream=300.0e-6, rho_f=1000.0, mu_f=1.0e-3, c0=5.0, eta_density=0.1, gamma_eos=7.0, Q_flow=0.0,
                rho_m=1738.0, D_liquid=1.0e-9, D_grain=0.0, D_gb=0.0, C_solid_init=1.0, C_liquid_init=0.0,
                C_thresh=0.2, C_sat=10.0, alpha_art_diff=0.0, gb_width_cells=0, cfl_factor=0.25,
                cfl_factor_corr=0.25, use_implicit=1, implicit_dt_max=60.0, implicit_dt_fraction=0.5)


def make_case(extra=None):
    ov = dict(TEST_CFG)
    ov.update(extra or {})
    cfg = Config.load(None, ov, quiet=True)      # make_test_config (tests/test_implicit.cpp:25-61)
    port = PortSim(2, cfg, threads=2)
    orc = ImplicitOracle(2, port.Nx, port.Ny, port.Nz, port.node_type, port.off_d, port.off_dist, port.off_evec,
                         port.off_vol, cfg)
    i = np.arange(port.N) % port.Nx
    j = np.arange(port.N) // port.Nx
    x = port.origin[0] + i * cfg.dx
    y = port.origin[1] + j * cfg.dx
    return cfg, port, orc, x, y


def gaussian(x, y, nt, sigma, D=0.0, t=0.0):
    s2 = sigma * sigma
    s2t = s2 + 2.0 * D * t                                           # gaussian_exact_2d (:117-126)
    return np.where(nt == 0, (s2 / s2t) * np.exp(-(x * x + y * y) / (2.0 * s2t)), 0.0)


def l2(C, ref, nt):                                                  # compute_L2_error (:129-139)
    m = nt == 0
    return float(np.sqrt(((C[m] - ref[m]) ** 2).sum() / ((ref[m] ** 2).sum() + 1e-30)))


def test_operator_structure():
    cfg, port, orc, x, y = make_case()
    nt = port.node_type
    C0 = gaussian(x, y, nt, 30e-6)
    zeros = np.zeros(port.N, np.uint8)
    M = orc.assemble(C0, np.zeros((port.N, 2)), zeros, zeros)
    off = M - __import__("scipy.sparse", fromlist=["diags"]).diags(M.diagonal())
    assert off.data.min() >= 0.0
    rows_with_bc = np.zeros(orc.l2g.size, bool)
    rows_with_bc[orc.bc_k] = True
    rs = np.asarray(M.sum(axis=1)).ravel()
    scale = np.abs(M.diagonal()).max()
    assert np.abs(rs[~rows_with_bc]).max() <= 1e-12 * scale
    assert (rs[rows_with_bc] < 0).all()                              # the BC bonds sit on the diagonal only


def test_reference_test1_pure_diffusion_thresholds():
    cfg, port, orc, x, y = make_case()
    nt = port.node_type
    sigma, D, t_end = 30e-6, 1.0e-9, 0.5
    C0 = gaussian(x, y, nt, sigma)
    exact = gaussian(x, y, nt, sigma, D, t_end)
    zeros = np.zeros(port.N, np.uint8)
    orc.assemble(C0, np.zeros((port.N, 2)), zeros, zeros)
    mass0 = C0[nt == 0].sum()
    errs, dts = [], [0.01, 0.05, 0.1]
    import scipy.sparse.linalg as spla
    for dt in dts:
        C, t, lu = C0.copy(), 0.0, {}
        while t < t_end - 1e-12:
            h = min(dt, t_end - t)
            A, b = orc.system(C, h)                  # orc.step with the factorisation of A(h) reused across the steps
            if h not in lu:
                lu[h] = spla.splu(A.tocsc())
            C = C.copy()
            C[orc.l2g] = np.clip(lu[h].solve(b), 0.0, cfg.C_solid_init)
            t += h
        errs.append(l2(C, exact, nt))
        if dt == dts[0]:
            assert abs(C[nt == 0].sum() - mass0) / mass0 * 100.0 <= 1.0          # :312-315
    assert errs[0] <= 0.05, errs                                                  # :308-311
    rates = [np.log(errs[k + 1] / errs[k]) / np.log(dts[k + 1] / dts[k]) for k in range(len(dts) - 1)]
    assert max(rates) > 0.4, rates                                                # :316-321
