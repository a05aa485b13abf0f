"""Two-level AMR grid.  CPU part: the host-side grid build of libpdgpu.so against the UNMODIFIED
reference (oracle/_ref: Grid::build_amr / build_neighbors_celllist, src/grid.cpp:352-796) -- every array bit for
bit.  GPU part: update_fictitious, the BCs, the explicit NS / ARD loop bodies, solve_steady and the phase change
on that grid against the reference, 1e-12 of the field maximum (same summation order: CSR order)."""
import os

import numpy as np
import pytest

from oracle import refapi
from pd_mg_pin_corrosion_b200.config import Config

import helpers as H

needs_ref = pytest.mark.skipif(not refapi.have_ref(2), reason="oracle/_ref not built")
TOL = 1e-12

AMR_CASES = {
    "amr_default": ("params.cfg", {"use_amr": 1, "amr_ratio": 3, "amr_buffer": 50e-6}),
    "amr_ratio2": ("params.cfg", {"use_amr": 1, "amr_ratio": 2, "amr_buffer": 40e-6, "dx": 4.0e-6}),
    "amr_shipped": ("params_amr.cfg", {}),
    "amr_offgrid": ("params.cfg", {"use_amr": 1, "amr_ratio": 3, "amr_buffer": 47e-6, "dx": 4.3e-6, "R_tube": 151e-6,
                                   "R_wire": 33e-6, "L_wire": 207e-6, "L_upstream": 260e-6, "L_downstream": 310e-6}),
}
GEOM = ["pos", "node_type", "dx_local", "delta_local", "grid_level", "fict_offset", "fict_source", "fict_weight",
        "nbr_offset", "nbr_index", "nbr_dist", "nbr_evec", "nbr_vol"]


def _vtu_arrays(path):
    import re
    out = {}
    for m in re.finditer(r'<DataArray type="\w+"(?: Name="(\w+)")?[^>]*>\n(.*?)</DataArray>', open(path).read(), re.S):
        out[m.group(1) or "points"] = np.array(m.group(2).split(), float)
    return out


def assert_same_snapshots(ref_dir, got_dir):
    """the VTU series + PVD collections of a whole run against the reference's own files: same file names, the
    initial state byte for byte, integer arrays and geometry exactly, fields to the 6 printed digits"""
    import glob
    names = sorted(os.path.basename(f) for f in glob.glob(os.path.join(str(ref_dir), "*.vtu")))
    assert len(names) >= 4 and names == sorted(os.path.basename(f) for f in glob.glob(os.path.join(str(got_dir), "*.vtu")))
    first = [n for n in names if n.startswith("state_000000")][0]
    assert open(os.path.join(str(ref_dir), first), "rb").read() == open(os.path.join(str(got_dir), first), "rb").read()
    for n in names:
        a, b = _vtu_arrays(os.path.join(str(ref_dir), n)), _vtu_arrays(os.path.join(str(got_dir), n))
        assert list(a) == list(b), n
        for k in a:
            assert a[k].shape == b[k].shape, (n, k)
            if k in ("velocity", "pressure", "concentration", "D_map"):
                assert np.abs(a[k] - b[k]).max() <= 2e-5 * max(np.abs(a[k]).max(), 1e-300), (n, k)
            else:
                assert np.array_equal(a[k], b[k]), (n, k)
    for pvd in ("simulation.pvd", "flow.pvd"):
        ra, rb = open(os.path.join(str(ref_dir), pvd)).read().split(), open(os.path.join(str(got_dir), pvd)).read().split()
        assert [t for t in ra if t.startswith("file=")] == [t for t in rb if t.startswith("file=")], pvd
        ta = np.array([float(t.split('"')[1]) for t in ra if t.startswith("timestep=")])
        tb = np.array([float(t.split('"')[1]) for t in rb if t.startswith("timestep=")])
        assert ta.shape == tb.shape and np.allclose(ta, tb, rtol=1e-6, atol=0.0), pvd


def both(case, fields=False, extra=None):
    from pd_mg_pin_corrosion_b200.amr import AmrGrid
    base, ov = AMR_CASES[case]
    ov = dict(ov, use_implicit=0, **(extra or {}))
    ref = refapi.RefSim(2, base, ov, threads=1, build=True, fields=fields)
    cfg = Config.load(os.path.join(H.CONFIG_DIR, base), ov, quiet=True)
    g = AmrGrid(cfg)
    g.build_amr()
    g.build_neighbors_celllist()
    return ref, cfg, g


@needs_ref
@pytest.mark.parametrize("case", list(AMR_CASES))
def test_amr_grid_build_bit_exact(case):
    ref, cfg, g = both(case)
    i = g.info
    assert (i.N_total, i.nnz, i.n_fict_entries) == (ref.N, ref.nnz, ref.lib.ref_fict_entries(ref.h))
    assert list(i.counts) == np.bincount(ref.get("node_type"), minlength=7).tolist()
    assert i.n_fict > 0 and i.n_fine > 0 and i.n_coarse > 0
    for name in GEOM:
        a, b = g.get(name), ref.get(name)
        assert a.shape == b.shape, name
        assert np.array_equal(a, b), (case, name, int((a != b).sum()))
    assert (ref.origin[0], ref.origin[1]) == (i.origin[0], i.origin[1])


@needs_ref
@pytest.mark.parametrize("case", ["amr_default", "amr_offgrid"])
def test_vtu_writer_bytes_match_reference(case, tmp_path):
    """host/vtu.cpp (pdhost_write_vtu; the snapshot writer of both AMR drivers) against the reference's own
    VTKWriter::write_vtu (src/vtk_writer.cpp:199-346) on the same arrays: byte-identical files, including the values
    safe_val() flushes (NaN, inf, |v| < 1e-300), WALL velocities written as 0, negative zero and OUTSIDE filtering.
    Also D_map of initialize_fields (pdhost_init_dmap vs src/main.cpp:19-112)."""
    from pd_mg_pin_corrosion_b200 import amr as A
    ref, cfg, g = both(case, fields=True)
    nt = ref.get("node_type")
    assert np.array_equal(A.init_dmap(cfg, nt, ref.get("is_gb"), ref.get("is_precip")), ref.get("D_map"))
    rng = np.random.default_rng(11)
    N = ref.N
    for name, comps in (("vel", 2), ("pressure", 1), ("C", 1), ("D_map", 1)):
        a = rng.standard_normal((N, comps) if comps > 1 else N) * 10.0 ** rng.integers(-12, 9, (N, comps) if comps > 1 else N)
        flat = a.reshape(-1)
        flat[::17] = 0.0
        flat[3::101] = -0.0
        flat[5::103] = np.nan
        flat[7::107] = np.inf
        flat[11::109] = -np.inf
        flat[13::113] = 4.9e-324
        flat[19::127] = -9.9e-301
        flat[23::131] = 1e-300
        flat[29::137] = 123456.5
        flat[31::139] = 0.0001234565
        flat[37::149] = 1e22
        ref.set(name, a)
    ref.set("phase", rng.integers(0, 2, N).astype(np.uint8))
    ref.set("grain_id", rng.integers(-1, 500, N).astype(np.int32))
    want, got = str(tmp_path / "ref.vtu"), str(tmp_path / "got.vtu")
    ref.write_vtu(want)
    A.write_vtu_arrays(got, ref.get("pos"), nt, ref.get("vel"), ref.get("pressure"), ref.get("C"), ref.get("phase"),
                       ref.get("grid_level"), ref.get("dx_local"), ref.get("grain_id"), ref.get("D_map"), ref.get("is_gb"),
                       ref.get("is_precip"))
    a, b = open(want, "rb").read(), open(got, "rb").read()
    assert len(a) > 100000 and a == b
    ref.close()


class _ReferenceBackedCloud:
    """The AmrGrid surface amr.AmrCoupledSolver drives, served by the compiled reference instead of the device: lets
    the HOST logic of the coupled loop (cycle structure, batching between diagnostics rows, snapshot schedule, frame
    numbering, PVD files, D_map patching, CSV formatting) be checked on CPU against the reference's own main()."""

    def __init__(self, ref, cfg):
        self.ref, self.cfg = ref, cfg
        if cfg.use_implicit:
            ref.imp_init()

    # implicit branch (libpdrefimp2d.so: the reference's own src/pd_ard_implicit.cpp)
    def implicit_assemble(self): self.ref.imp_assemble()
    def implicit_compute_dt(self): return self.ref.imp_compute_adaptive_dt()
    def inlet_bc(self): self.ref.inlet_bc()
    def outlet_bc(self): self.ref.outlet_bc()
    def wall_conc_bc(self): self.ref.wall_conc_bc()
    def smooth_conc(self): self.ref.smooth_conc()

    def implicit_step(self, dt, tol=1e-10, restart=50, max_iters=200):
        from types import SimpleNamespace
        self.ref.imp_step(dt)
        return SimpleNamespace(iters=0, rel_res=0.0)

    def get(self, name): return self.ref.get(name)
    def get_field(self, name): return self.ref.get(name)

    def ns_solve_steady(self):
        from types import SimpleNamespace
        return SimpleNamespace(iters=self.ref.ns_solve_steady(), eps=0.0)

    def update_fictitious(self): self.ref.update_fictitious()
    def ard_set_volume_loss(self, v):
        self.ref.ard_set_volume_loss(v)
        if self.cfg.use_implicit:
            self.ref.imp_set_volume_loss(v)
    def ard_compute_dt(self): return self.ref.ard_compute_dt()
    def ard_iterate(self, n, dt): self.ref.ard_iterate(n, dt)

    def phase_change(self):
        n = self.ref.phase_change()
        if n > 0:                                   # src/coupling.cpp:262-268
            self.ref.lib.ref_update_node_types(self.ref.h)
            self.ref.lib.ref_build_neighbors_celllist(self.ref.h)
        return n


@needs_ref
def test_amr_coupled_loop_host_logic_matches_reference_main(tmp_path):
    """amr.AmrCoupledSolver.run (the Python driver of the AMR coupled loop) over reference-served operators against
    the reference's own main() with use_amr = 1: every output file byte for byte -- 20+ VTU snapshots with their frame
    numbers and times, simulation.pvd, flow.pvd, diagnostics.csv."""
    import glob
    from pd_mg_pin_corrosion_b200 import amr as A
    base, ov = AMR_CASES["amr_ratio2"]
    ov = dict(ov, use_implicit=0, D_grain=5e-11, D_gb=5e-9, C_thresh=0.999, corrosion_steps_per_check=40,
              flow_max_iters=120, T_final=3.2e-4, output_every_corr=10)
    cfg_path = refapi.write_cfg(base, dict(ov, output_dir=str(tmp_path / "ref")), str(tmp_path / "amr.cfg"))
    refapi._lib(2).ref_set_threads(1)
    assert refapi.run_reference_main(2, cfg_path) == 0
    ref = refapi.RefSim(2, base, ov, threads=1, build=True, fields=True)
    cfg = Config.load(os.path.join(H.CONFIG_DIR, base), ov, quiet=True)
    cs = A.AmrCoupledSolver()
    cs.run(_ReferenceBackedCloud(ref, cfg), str(tmp_path / "got"), grain_id=ref.get("grain_id"))
    names = sorted(os.path.basename(f) for f in glob.glob(str(tmp_path / "ref" / "*")))
    assert sorted(os.path.basename(f) for f in glob.glob(str(tmp_path / "got" / "*"))) == [n for n in names if n != "mass_loss.csv"]
    assert sum(n.endswith(".vtu") for n in names) >= 12 and cs.rows[-1][3] < cs.rows[0][3]
    for n in names:
        if n == "mass_loss.csv":
            continue
        a, b = open(tmp_path / "ref" / n, "rb").read(), open(tmp_path / "got" / n, "rb").read()
        if n.endswith(".pvd"):                       # the reference stores the path it was given; same relative names
            a, b = a.replace(str(tmp_path / "ref").encode(), b""), b.replace(str(tmp_path / "got").encode(), b"")
        assert a == b, n
    ref.close()


IMPLICIT_AMR_RUN = dict(use_implicit=1, D_grain=5e-11, D_gb=5e-9, C_thresh=0.999, corrosion_steps_per_check=6, flow_max_iters=150,
                        T_final=8e-4, implicit_dt_max=0.004, implicit_dt_fraction=0.5, diagnostic_every=1, implicit_output_every=3)


@pytest.mark.skipif(not refapi.have_ref(2, implicit=True), reason="oracle/_ref/libpdrefimp2d.so not built")
def test_amr_implicit_coupled_loop_host_logic_matches_reference_main(tmp_path):
    """the IMPLICIT branch of amr.AmrCoupledSolver.run (src/coupling.cpp:154-216 with use_amr = 1: assemble per cycle,
    adaptive dt, BCs, step, smoother, IDW refresh, diagnostics / snapshot cadence, cycle end at the first solid below
    C_thresh) over reference-served operators against the reference's own main(): every output file byte for byte."""
    import glob
    from pd_mg_pin_corrosion_b200 import amr as A
    base, ov = AMR_CASES["amr_ratio2"]
    ov = dict(ov, **dict(IMPLICIT_AMR_RUN, T_final=4.5e-4))
    cfg_path = refapi.write_cfg(base, dict(ov, output_dir=str(tmp_path / "ref")), str(tmp_path / "amr.cfg"))
    refapi._lib(2, True).ref_set_threads(1)          # one thread: the reference's in-place smoother is order dependent
    assert refapi.run_reference_main(2, cfg_path, implicit=True) == 0
    ref = refapi.RefSim(2, base, ov, threads=1, build=True, fields=True, implicit=True)
    cfg = Config.load(os.path.join(H.CONFIG_DIR, base), ov, quiet=True)
    assert cfg.use_implicit == 1
    cs = A.AmrCoupledSolver()
    cs.run(_ReferenceBackedCloud(ref, cfg), str(tmp_path / "got"), grain_id=ref.get("grain_id"))
    names = sorted(os.path.basename(f) for f in glob.glob(str(tmp_path / "ref" / "*")))
    assert sorted(os.path.basename(f) for f in glob.glob(str(tmp_path / "got" / "*"))) == [n for n in names if n != "mass_loss.csv"]
    assert len(cs.rows) >= 6 and cs.rows[-1][3] < cs.rows[0][3] and any(n.startswith("corr_") for n in names)
    for n in names:
        if n == "mass_loss.csv":
            continue
        a, b = open(tmp_path / "ref" / n, "rb").read(), open(tmp_path / "got" / n, "rb").read()
        if n.endswith(".pvd"):
            a, b = a.replace(str(tmp_path / "ref").encode(), b""), b.replace(str(tmp_path / "got").encode(), b"")
        assert a == b, n
    ref.close()


@needs_ref
def test_amr_wall_mirror_table():
    """index trick: rho[i] = i + 0.25, apply_wall_bc on the reference, read the WALL nodes"""
    ref, cfg, g = both("amr_default", fields=True)
    N = ref.N
    ref.set("rho", np.arange(N) + 0.25)
    ref.wall_bc()
    nt = ref.get("node_type")
    got = g.get("wall_mirror")
    rho = ref.get("rho")
    walls = np.nonzero(nt == 2)[0]
    assert walls.size > 0 and np.all(got[nt != 2] == -2)
    for n in walls:
        m = got[n]
        expect = (m + 0.25) if m >= 0 else cfg.rho_f
        assert rho[n] == expect, (n, m, rho[n])


@needs_ref
@pytest.mark.parametrize("case,extra", [("amr_default", None), ("amr_shipped", None),
                                        ("amr_ratio2", {"gb_width_cells": 1, "precip_cluster_cells": 2})])
def test_amr_grain_generation_bit_exact(case, extra):
    """GrainStructure::generate on the cloud (positions + CSR, same RNG draws) against the reference"""
    from pd_mg_pin_corrosion_b200 import amr as A
    ref, cfg, g = both(case, fields=True, extra=extra)
    gid, gb, pr, n = A.generate_grains(g)
    assert n == ref.lib.ref_n_grains(ref.h) and n > 1
    assert np.array_equal(gid, ref.get("grain_id"))
    assert np.array_equal(gb, ref.get("is_gb")) and gb.sum() > 0
    assert np.array_equal(pr, ref.get("is_precip")) and pr.sum() > 0


def gpu_pair(case, extra=None):
    ref, cfg, g = both(case, fields=True, extra=extra)
    g.device_init(0)
    for n in ("rho", "vel", "C", "rho_new", "vel_new", "C_new", "phase", "is_gb", "is_precip"):
        g.set_field(n, ref.get(n))
    return ref, cfg, g


def assert_close(g, ref, names, tol=TOL):
    for n in names:
        a, b = g.get_field(n), ref.get(n)
        e = H.rel_err(a, b)
        assert e <= tol, f"{n}: rel err {e:.3e}"


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("case", ["amr_default", "amr_ratio2", "amr_offgrid"])
def test_amr_operators_match_reference(case):
    ref, cfg, g = gpu_pair(case)
    H.perturbed_state(ref, seed=3)
    for n in ("rho", "vel", "C", "rho_new", "vel_new", "C_new"):
        g.set_field(n, ref.get(n))
    for op in ("inlet_bc", "outlet_bc", "wall_bc", "solid_bc", "wall_conc_bc"):
        getattr(ref, op)(); getattr(g, op)()
        assert_close(g, ref, ("rho", "vel", "C"))
    ref.update_fictitious(); g.update_fictitious()
    assert_close(g, ref, ("rho", "vel", "C", "pressure"))
    dt = ref.ns_compute_dt()
    assert abs(g.ns_compute_dt() - dt) <= 1e-15 * dt
    ref.ns_step(dt); g.ns_step(dt)
    assert_close(g, ref, ("rho_new", "vel_new", "pressure"))
    ref.wall_bc_new(); g.wall_bc_new()
    assert_close(g, ref, ("rho_new", "vel_new"))
    dtc = ref.ard_compute_dt()
    assert abs(g.ard_compute_dt() - dtc) <= 1e-15 * dtc
    ref.ard_step(dtc); g.ard_step(dtc)
    assert_close(g, ref, ("C_new",))
    g.close()


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("case", ["amr_default", "amr_ratio2"])
def test_amr_loop_bodies_match_reference(case):
    ref, cfg, g = gpu_pair(case)
    dt = ref.ns_compute_dt()
    ref.ns_iterate(40, dt); g.ns_iterate(40, dt)
    assert_close(g, ref, ("rho", "vel", "C", "pressure"))
    dtc = ref.ard_compute_dt()
    ref.ard_iterate(25, dtc); g.ard_iterate(25, dtc)
    assert_close(g, ref, ("rho", "vel", "C"))
    g.close()


@pytest.mark.gpu
@needs_ref
def test_amr_outlet_level_list_kernel():
    """the general outlet kernel (levels over global memory; taken when the OUTLET nodes do not fit one CTA), forced
    through the environment in a child process: same loop-body results"""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import test_amr as T\n"
            "ref, cfg, g = T.gpu_pair('amr_default')\n"
            "dt = ref.ns_compute_dt(); ref.ns_iterate(12, dt); g.ns_iterate(12, dt)\n"
            "T.assert_close(g, ref, ('rho', 'vel', 'C'))\n"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PDGPU_AMR_OUTLET_LIST="1"), capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]


@pytest.mark.gpu
@needs_ref
def test_amr_solve_steady_and_phase_change():
    extra = {"flow_max_iters": 300, "D_grain": 5e-11, "D_gb": 5e-9, "C_thresh": 0.999}
    ref, cfg, g = gpu_pair("amr_default", extra)
    it_ref = ref.ns_solve_steady()
    r = g.ns_solve_steady()
    assert r.iters == it_ref
    assert_close(g, ref, ("rho", "vel", "rho_new", "vel_new"), tol=1e-11)
    ref.update_fictitious(); g.update_fictitious()          # src/coupling.cpp:139
    dtc = ref.ard_compute_dt()
    total = 0
    for _ in range(6):
        ref.ard_iterate(40, dtc); g.ard_iterate(40, dtc)
        n_ref, n = ref.phase_change(), g.phase_change()
        assert n == n_ref
        total += n
        if n_ref:
            ref.lib.ref_update_node_types(ref.h)
            ref.lib.ref_build_neighbors_celllist(ref.h)
        assert np.array_equal(g.get_field("node_type"), ref.get("node_type"))
        assert np.array_equal(g.get_field("phase"), ref.get("phase"))
        assert_close(g, ref, ("rho", "vel", "C"), tol=1e-11)
    assert total > 0, "the synthetic diffusivities must dissolve nodes"
    g.close()


@pytest.mark.gpu
@needs_ref
def test_amr_whole_coupled_run_matches_reference_main(tmp_path):
    """CoupledSolver::run with use_amr = 1 (explicit branch) through the reference's own main() -- grid build,
    grains, flow solve, IDW refresh, corrosion cycles with dissolution and flow re-solves -- against
    amr.AmrCoupledSolver on the device: diagnostics.csv rows within 1e-6, identical solid counts."""
    from pd_mg_pin_corrosion_b200 import amr as A
    base, ov = AMR_CASES["amr_default"]
    ov = dict(ov, use_implicit=0, D_grain=5e-11, D_gb=5e-9, C_thresh=0.999, corrosion_steps_per_check=50,
              flow_max_iters=300, T_final=6e-4, output_every_corr=10, output_dir=str(tmp_path / "ref"))
    cfg_path = refapi.write_cfg(base, ov, str(tmp_path / "amr.cfg"))
    refapi._lib(2).ref_set_threads(1)
    assert refapi.run_reference_main(2, cfg_path) == 0
    gold = np.loadtxt(tmp_path / "ref" / "diagnostics.csv", delimiter=",", skiprows=1, ndmin=2)
    assert gold.shape[0] >= 5 and gold[-1, 3] < gold[0, 3] + 1, "the run must dissolve nodes"
    ref = refapi.RefSim(2, base, ov, threads=1, build=True, fields=True)      # same grains (seed 42)
    cfg = Config.load(os.path.join(H.CONFIG_DIR, base), ov, quiet=True)
    g = A.AmrGrid(cfg)
    g.build_amr(); g.build_neighbors_celllist(); g.device_init(0)
    gid, gb, pr, _ = A.generate_grains(g)               # standalone: own grains (bit-exact, tested above)
    A.initialize_fields(g, gb, pr)
    for n in ("rho", "vel", "C", "phase", "is_gb", "is_precip"):
        assert np.array_equal(g.get_field(n), ref.get(n)), n        # initialize_fields itself
    rows = np.array(A.AmrCoupledSolver().run(g, str(tmp_path / "gpu"), grain_id=gid))
    assert rows.shape == gold.shape
    assert np.array_equal(rows[:, 3], gold[:, 3])
    for col in (0, 1, 2, 4, 5):
        rel = np.abs(rows[:, col] - gold[:, col]) / np.maximum(np.abs(gold[:, col]), 1e-300)
        assert rel.max() <= 1e-6, (col, float(rel.max()))
    got = np.loadtxt(tmp_path / "gpu" / "diagnostics.csv", delimiter=",", skiprows=1, ndmin=2)
    assert got.shape == gold.shape
    assert_same_snapshots(tmp_path / "ref", tmp_path / "gpu")          # VTU series + PVD files (src/coupling.cpp:117-296)
    g.close()


@pytest.mark.gpu
@needs_ref
def test_amr_host_driver_matches_reference_main(tmp_path):
    """host/pd_corrosion_gpu with use_amr = 1 (C++ driver over pdamr_*, host/amr_run.cpp) against the reference's own
    main(): same number of diagnostics rows, identical solid counts, values within 1e-6."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "host", "pd_corrosion_gpu")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(root, "host")])
    base, ov = AMR_CASES["amr_ratio2"]
    outs = {}
    for who in ("ref", "gpu"):
        o = dict(ov, use_implicit=0, D_grain=5e-11, D_gb=5e-9, C_thresh=0.999, corrosion_steps_per_check=40,
                 flow_max_iters=250, T_final=5e-4, output_every_corr=10, output_dir=str(tmp_path / who))
        cfg_path = refapi.write_cfg(base, o, str(tmp_path / f"{who}.cfg"))
        if who == "ref":
            refapi._lib(2).ref_set_threads(1)
            assert refapi.run_reference_main(2, cfg_path) == 0
        else:
            r = subprocess.run([exe, cfg_path, "--dim", "2"], capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs[who] = np.loadtxt(tmp_path / who / "diagnostics.csv", delimiter=",", skiprows=1, ndmin=2)
    gold, got = outs["ref"], outs["gpu"]
    assert_same_snapshots(tmp_path / "ref", tmp_path / "gpu")
    assert gold.shape == got.shape and gold.shape[0] >= 4
    assert np.array_equal(gold[:, 3], got[:, 3]) and gold[-1, 3] < gold[0, 3]
    for col in (0, 1, 2, 4, 5):
        rel = np.abs(got[:, col] - gold[:, col]) / np.maximum(np.abs(gold[:, col]), 1e-300)
        assert rel.max() <= 1e-6, (col, float(rel.max()))
