"""Implicit ARD branch on the two-level AMR cloud (pdamr_implicit_*, csrc/amr_implicit.cuh; SURVEY 8f-2 x 8f-4)
against the reference's OWN compiled code: oracle/_ref/libpdrefimp2d.so = the unmodified src/pd_ard_implicit.cpp
(+ grid / boundary / coupling) built against the Eigen work-alike oracle/eigen_min/ (see
tests/test_reference_implicit.py).  System matrix (through matvec on random vectors), right-hand side and adaptive
step 1e-12; one step to the solve tolerance; the boundary smoother exactly; the whole implicit coupled run of the
reference's main() with use_amr = 1 row for row (1e-6) with its VTU series."""
import os

import numpy as np
import pytest

import helpers as H
from oracle import refapi
from pd_mg_pin_corrosion_b200.config import Config
from test_amr import AMR_CASES, IMPLICIT_AMR_RUN, assert_same_snapshots

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refapi.have_ref(2, implicit=True), reason="oracle/_ref/libpdrefimp2d.so not built")]


def _pair(case, extra, ns_iters):
    from pd_mg_pin_corrosion_b200 import amr as A
    base, ov = AMR_CASES[case]
    ov = dict(ov, use_implicit=1, **extra)
    ref = refapi.RefSim(2, base, ov, threads=1, build=True, fields=True, implicit=True)
    ref.ns_iterate(ns_iters, ref.ns_compute_dt())
    cfg = Config.load(os.path.join(H.CONFIG_DIR, base), ov, quiet=True)
    g = A.AmrGrid(cfg)
    g.build_amr(); g.build_neighbors_celllist(); g.device_init(0)
    nt = ref.get("node_type")
    rng = np.random.default_rng(5)
    Cc = ref.get("C")
    Cc = np.abs(Cc + 0.05 * rng.standard_normal(Cc.size)) * (nt != 5)
    Cc[(nt == 0) & (np.arange(Cc.size) % 41 == 0)] = 0.95            # saturated fluid: salt layer on a few solids
    ref.set("C", Cc)
    for n in ("rho", "vel", "C", "rho_new", "vel_new", "C_new", "phase", "is_gb", "is_precip"):
        g.set_field(n, ref.get(n))
    ref.imp_init()
    ref.imp_set_volume_loss(0.013)
    ref.imp_assemble()
    g.ard_set_volume_loss(0.013)
    g.implicit_assemble()
    return ref, g, cfg, nt, Cc


@pytest.mark.parametrize("case", ["amr_default", "amr_ratio2"])
def test_amr_implicit_operator_rhs_dt_step(case):
    ref, g, cfg, nt, Cc = _pair(case, {"D_grain": 5e-11, "D_gb": 5e-9, "corrosion_decay_l": 0.1, "C_thresh": 0.999,
                                            "implicit_dt_max": 1e-4}, 150)
    unk = np.nonzero((nt == 0) | (nt == 1) | (nt == 6))[0]
    assert (nt == 6).sum() > 0
    t = ref.imp_compute_adaptive_dt()                                 # flux-limited: dt_max = 1e-4 s in this config
    assert 0.01 * cfg.implicit_dt_max < t < cfg.implicit_dt_max * cfg.implicit_dt_fraction
    assert g.implicit_compute_dt() == pytest.approx(t, rel=1e-12)     # src/pd_ard_implicit.cpp:438-487
    assert g.implicit_compute_dt(0.5, 1e-6) == pytest.approx(5e-7, rel=1e-12)      # capped
    assert g.implicit_compute_dt(0.5, 600.0) == pytest.approx(6.0, rel=1e-12)      # floor: 1 % of dt_max
    rng = np.random.default_rng(6)
    for dt in (1e-3, 0.7, 30.0):
        ref.set("C", Cc)
        assert ref.imp_step(dt) == 1
        A, b, x, _, err = ref.imp_last_system()
        assert A.shape[0] == unk.size and err <= 1e-9
        xv = np.zeros(nt.size)
        xv[unk] = rng.standard_normal(unk.size)
        y = g.implicit_matvec(dt, xv)
        want = A @ xv[unk]
        assert np.abs(y[unk] - want).max() <= 1e-12 * np.abs(want).max(), (dt, "matvec")      # :104-346, :384-388, :500-531
        mask = np.ones(nt.size, bool)
        mask[unk] = False
        assert not y[mask].any()
        g.set_field("C", Cc)
        bd = g.implicit_rhs(dt)
        assert np.abs(bd[unk] - b).max() <= 1e-12 * np.abs(b).max(), (dt, "rhs")              # :352-362, :391, :519-530
        info = g.implicit_step(dt, tol=1e-12, restart=50, max_iters=2000)
        assert info.converged, (dt, info.iters, info.rel_res)
        assert H.rel_err(g.get_field("C"), ref.get("C")) <= 1e-8, (dt, info.iters, info.rel_res)   # solve + clamp, :409-427
    g.close()
    ref.close()


def test_amr_sweep_preconditioner_iteration_count():
    """the axial sweep keeps GMRES within the reference's iteration budget (200) at the production step size"""
    ref, g, cfg, nt, Cc = _pair("amr_default", {"D_grain": 5e-11, "D_gb": 5e-9}, 400)
    g.set_field("C", ref.get("C"))
    info = g.implicit_step(30.0, tol=1e-10, restart=50, max_iters=200, precond=1)
    assert info.converged and info.iters <= 60, (info.iters, info.rel_res)
    g.close()
    ref.close()


def test_amr_boundary_smoother_exact():
    """smooth_boundary_concentration on the cloud: in place in ascending node order (single-threaded reference)"""
    ref, g, cfg, nt, Cc = _pair("amr_offgrid", {}, 20)
    ref.lib.ref_set_threads(1)
    rng = np.random.default_rng(8)
    Cr = rng.random(nt.size) * (nt != 5)
    ref.set("C", Cr)
    g.set_field("C", Cr)
    ref.smooth_conc()
    g.smooth_conc()
    want, got = ref.get("C"), g.get_field("C")
    assert (want != Cr).sum() > 50
    assert np.array_equal(got != Cr, want != Cr)
    assert np.abs(got - want).max() <= 4e-16
    g.close()
    ref.close()


def test_amr_implicit_whole_run_matches_reference_main(tmp_path):
    """CoupledSolver::run with use_amr = 1 and use_implicit = 1 through the reference's own main() against
    amr.AmrCoupledSolver on the device: diagnostics rows within 1e-6, identical solid counts, same VTU series."""
    from pd_mg_pin_corrosion_b200 import amr as A
    base, ov = AMR_CASES["amr_ratio2"]
    ov = dict(ov, **IMPLICIT_AMR_RUN, output_dir=str(tmp_path / "ref"))
    cfg_path = refapi.write_cfg(base, ov, str(tmp_path / "amr.cfg"))
    refapi._lib(2, True).ref_set_threads(1)
    assert refapi.run_reference_main(2, cfg_path, implicit=True) == 0
    gold = np.loadtxt(tmp_path / "ref" / "diagnostics.csv", delimiter=",", skiprows=1, ndmin=2)
    assert gold.shape[0] >= 6 and gold[-1, 3] < gold[0, 3], "the run must dissolve nodes"
    cfg = Config.load(os.path.join(H.CONFIG_DIR, base), ov, quiet=True)
    g = A.AmrGrid(cfg)
    g.build_amr(); g.build_neighbors_celllist(); g.device_init(0)
    gid, gb, pr, _ = A.generate_grains(g)
    A.initialize_fields(g, gb, pr)
    cs = A.AmrCoupledSolver()
    cs.tol, cs.max_iters = 1e-12, 2000
    rows = np.array(cs.run(g, str(tmp_path / "gpu"), grain_id=gid))
    assert rows.shape == gold.shape
    assert np.array_equal(rows[:, 3], gold[:, 3])
    for col in (0, 1, 2, 4, 5):
        rel = np.abs(rows[:, col] - gold[:, col]) / np.maximum(np.abs(gold[:, col]), 1e-300)
        assert rel.max() <= 1e-6, (col, float(rel.max()))
    assert_same_snapshots(tmp_path / "ref", tmp_path / "gpu")
    g.close()


def test_amr_implicit_host_driver_matches_reference_main(tmp_path):
    """host/pd_corrosion_gpu with use_amr = 1 and use_implicit = 1 (host/amr_run.cpp over pdamr_implicit_*) against the
    reference's own main(): diagnostics rows within 1e-6, identical solid counts, same VTU series."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "host", "pd_corrosion_gpu")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(root, "host")])
    base, ov = AMR_CASES["amr_default"]
    outs = {}
    for who in ("ref", "gpu"):
        o = dict(ov, **dict(IMPLICIT_AMR_RUN, T_final=6e-4), output_dir=str(tmp_path / who))
        cfg_path = refapi.write_cfg(base, o, str(tmp_path / f"{who}.cfg"))
        if who == "ref":
            refapi._lib(2, True).ref_set_threads(1)
            assert refapi.run_reference_main(2, cfg_path, implicit=True) == 0
        else:
            r = subprocess.run([exe, cfg_path, "--dim", "2"], capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs[who] = np.loadtxt(tmp_path / who / "diagnostics.csv", delimiter=",", skiprows=1, ndmin=2)
    gold, got = outs["ref"], outs["gpu"]
    assert gold.shape == got.shape and gold.shape[0] >= 4
    assert np.array_equal(gold[:, 3], got[:, 3]) and gold[-1, 3] < gold[0, 3]
    for col in (0, 1, 2, 4, 5):
        rel = np.abs(got[:, col] - gold[:, col]) / np.maximum(np.abs(gold[:, col]), 1e-300)
        assert rel.max() <= 1e-6, (col, float(rel.max()))
    assert_same_snapshots(tmp_path / "ref", tmp_path / "gpu")
