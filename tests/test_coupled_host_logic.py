"""CPU tests of the HOST logic of solver.CoupledSolver.run (the Python mirror of CoupledSolver::run,
src/coupling.cpp:82-302) without a GPU: the loop runs over operators served by the compiled reference
(oracle/_ref), so everything the driver itself decides -- cycle structure, device-resident batches between output
points, T_final cut-off inside a cycle, flow re-solve triggers, ordered volume-loss sum, diagnostics cadence of the
implicit branch, cycle end at the first solid below C_thresh, CSV formatting -- must reproduce diagnostics.csv and
mass_loss.csv of the reference's own main() byte for byte.  (The operators themselves are checked on the device in
the -m gpu tests.)"""
import os
from types import SimpleNamespace

import numpy as np
import pytest

import helpers as H
from oracle import refapi
from pd_mg_pin_corrosion_b200 import solver as S
from pd_mg_pin_corrosion_b200.config import Config


class _Grid:
    rank, nranks = 0, 1

    def __init__(self, ref):
        self.ref = ref

    @property
    def node_type_all(self): return self.ref.get("node_type")
    def allreduce(self, vals, op="sum"): return np.atleast_1d(np.asarray(vals, float))


class _Fields:
    def __init__(self, ref): self.ref = ref
    def gather(self, name, idx): return self.ref.get(name)[np.asarray(idx, np.int64)]


class _Flow:
    def __init__(self, ref): self.ref = ref
    def init(self, grid, cfg): self.ref.lib.ref_ns_init(self.ref.h)
    def solve_steady(self, fields, grid, cfg, verbose=False): return self.ref.ns_solve_steady()


class _Ard:
    def __init__(self, ref): self.ref = ref
    def init(self, grid, cfg): self.ref.lib.ref_ard_init(self.ref.h)
    def set_volume_loss(self, v, grid): self.ref.ard_set_volume_loss(v)
    def compute_dt(self, fields, grid, cfg): return self.ref.ard_compute_dt()
    def iterate(self, fields, grid, cfg, n, dt): self.ref.ard_iterate(n, dt)

    def apply_phase_change(self, fields, grid, cfg):
        n = self.ref.phase_change()
        if n > 0:
            self.ref.rebuild_neighbors()             # src/coupling.cpp:262-268
        return n


class _Implicit(_Ard):
    last = SimpleNamespace(iters=0, rel_res=0.0)
    def set_volume_loss(self, v, grid): self.ref.imp_set_volume_loss(v)
    def assemble(self, fields, grid, cfg): self.ref.imp_assemble()
    def compute_adaptive_dt(self, fields, grid, cfg): return self.ref.imp_compute_adaptive_dt()
    def step(self, fields, grid, cfg, dt): return self.ref.imp_step(dt)

    def apply_phase_change(self, fields, grid, cfg):
        n = self.ref.imp_phase_change()
        if n > 0:
            self.ref.rebuild_neighbors()
        return n


def _run_both(monkeypatch, tmp_path, dim, base, ov, implicit):
    cfg_path = refapi.write_cfg(base, dict(ov, output_dir=str(tmp_path / "ref")), str(tmp_path / "run.cfg"))
    refapi._lib(dim, implicit).ref_set_threads(4)
    assert refapi.run_reference_main(dim, cfg_path, implicit=implicit) == 0
    ref = refapi.RefSim(dim, base, ov, threads=4, implicit=implicit)
    if implicit:
        ref.imp_init()
    cfg = Config.load(os.path.join(H.CONFIG_DIR, base), dict(ov, output_dir=str(tmp_path / "got")), quiet=True)

    def diag(grid):
        nt, v, Cc = ref.get("node_type"), ref.get("vel"), ref.get("C")
        fl = nt == 0
        return SimpleNamespace(solid_count=int((nt == 1).sum()),
                               v_max=float(np.sqrt((v[fl] ** 2).sum(1)).max()) if fl.any() else 0.0,
                               C_max_fluid=float(max(Cc[fl].max(), 0.0)) if fl.any() else 0.0)

    monkeypatch.setattr(S, "diagnostics", diag)
    monkeypatch.setattr(S, "apply_inlet_bc", lambda f, g, c: ref.inlet_bc())
    monkeypatch.setattr(S, "apply_outlet_bc", lambda f, g, c: ref.outlet_bc())
    monkeypatch.setattr(S, "apply_wall_concentration_bc", lambda f, g, c: ref.wall_conc_bc())
    monkeypatch.setattr(S, "smooth_boundary_concentration", lambda f, g, c: ref.smooth_conc())
    monkeypatch.setattr(S.CoupledSolver, "_any_solid_below", staticmethod(
        lambda grid, fields, c: bool(((ref.get("node_type") == 1) & (ref.get("C") < c.C_thresh)).any())))
    cs = S.CoupledSolver()
    cs.log = lambda *a, **k: None
    cs.flow_solver, cs.ard_solver, cs.ard_implicit_solver = _Flow(ref), _Ard(ref), _Implicit(ref)
    cs.run(_Grid(ref), _Fields(ref), cfg)
    for name in ("diagnostics.csv", "mass_loss.csv"):
        a, b = open(tmp_path / "ref" / name, "rb").read(), open(tmp_path / "got" / name, "rb").read()
        assert len(a.splitlines()) >= 5, name
        assert a == b, name
    assert cs.total_dissolved > 0
    ref.close()


@pytest.mark.skipif(not refapi.have_ref(2), reason="oracle/_ref not built")
def test_explicit_loop_host_logic_matches_reference_main(monkeypatch, tmp_path):
    dim, base, ov = H.CASES["2d_dissolve"]
    ov = dict(ov, use_implicit=0, flow_max_iters=200)
    _run_both(monkeypatch, tmp_path, dim, base, ov, implicit=False)


@pytest.mark.skipif(not refapi.have_ref(2, implicit=True), reason="oracle/_ref/libpdrefimp2d.so not built")
def test_implicit_loop_host_logic_matches_reference_main(monkeypatch, tmp_path):
    dim, base, cov = H.CASES["2d_default"]
    ov = dict(cov, use_implicit=1, D_grain=5e-11, D_gb=5e-9, C_thresh=0.999, corrosion_steps_per_check=6, flow_max_iters=200,
              T_final=5e-4, implicit_dt_max=0.004, implicit_dt_fraction=0.5, diagnostic_every=2, implicit_output_every=1000000)
    _run_both(monkeypatch, tmp_path, dim, base, ov, implicit=True)
