"""Binary checkpoint / resume: a run continued from a checkpoint in a NEW context
is bit-identical to the uninterrupted run, including dissolved nodes and buffer parities."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
from test_gpu_parity import gpu_side

pytestmark = pytest.mark.gpu


def _advance(S, cfg, grid, fields, ns, ard, dt, n_ns, n_ard, dissolve):
    ns.iterate(fields, grid, cfg, n_ns, dt)
    ard.iterate(fields, grid, cfg, n_ard, ard.compute_dt(fields, grid, cfg))
    return ard.apply_phase_change(fields, grid, cfg) if dissolve else 0


@pytest.mark.parametrize("case,n1,n2,dissolve", [("3d_small", (7, 3), (6, 4), False), ("2d_dissolve", (300, 150), (40, 31), True)])
def test_resume_is_bit_identical(case, n1, n2, dissolve, tmp_path):
    from pd_mg_pin_corrosion_b200 import lib as L_
    L = L_.load()
    ref = H.make_ref(case)
    dt = ref.ns_compute_dt()
    path = str(tmp_path / "state.pdck").encode()

    S, cfg, grid, fields = gpu_side(case, ref=ref, upload=False)
    ns, ard = S.PD_NS_Solver(), S.PD_ARD_Solver()
    ns.init(grid, cfg); ard.init(grid, cfg)
    nd = _advance(S, cfg, grid, fields, ns, ard, dt, n1[0], n1[1], dissolve)
    if dissolve:
        assert nd > 0
    nbytes = C.c_longlong()
    L_.check(L.pdgpu_checkpoint_save(grid.ctx, path, C.byref(nbytes)))
    assert nbytes.value > grid.N_total * 8 * 2 * (2 + grid.dim)
    _advance(S, cfg, grid, fields, ns, ard, dt, n2[0], n2[1], dissolve)
    want = {n: fields.get(n) for n in ("rho", "vel", "C", "pressure", "phase")}
    want["node_type"] = grid.node_type.copy()
    grid.close()

    S, cfg, grid2, fields2 = gpu_side(case, ref=ref, upload=False)     # fresh context, fresh (initial) fields
    ns2, ard2 = S.PD_NS_Solver(), S.PD_ARD_Solver()
    ns2.init(grid2, cfg); ard2.init(grid2, cfg)
    L_.check(L.pdgpu_checkpoint_load(grid2.ctx, path))
    grid2._refresh()
    _advance(S, cfg, grid2, fields2, ns2, ard2, dt, n2[0], n2[1], dissolve)
    for n in ("rho", "vel", "C", "pressure", "phase"):
        assert np.array_equal(fields2.get(n), want[n], equal_nan=True), n
    assert np.array_equal(grid2.node_type, want["node_type"])
    grid2.close()


def test_checkpoint_rejects_other_configuration(tmp_path):
    from pd_mg_pin_corrosion_b200 import lib as L_
    L = L_.load()
    path = str(tmp_path / "a.pdck").encode()
    S, cfg, grid, fields = gpu_side("2d_default", ref=None)
    L_.check(L.pdgpu_checkpoint_save(grid.ctx, path, None))
    grid.close()
    S, cfg, grid, fields = gpu_side("2d_poiseuille", ref=None)
    assert L.pdgpu_checkpoint_load(grid.ctx, path) != 0
    assert b"another configuration" in L.pdgpu_last_error()
    (tmp_path / "junk").write_bytes(b"not a checkpoint")
    assert L.pdgpu_checkpoint_load(grid.ctx, str(tmp_path / "junk").encode()) != 0
    grid.close()


def test_host_driver_resume_reproduces_the_tail_of_the_run(tmp_path):
    """host/pd_corrosion_gpu --checkpoint/--resume: the rows a resumed run appends to diagnostics.csv are,
    character for character, the rows the uninterrupted run wrote after that cycle."""
    import os
    import subprocess
    from oracle import refapi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "host", "pd_corrosion_gpu")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(root, "host")])
    dim, base, ov = H.CASES["2d_dissolve"]
    ck = str(tmp_path / "ck")
    cfg_a = refapi.write_cfg(base, dict(ov, use_implicit=0, output_dir=str(tmp_path / "a")), str(tmp_path / "a.cfg"))
    r = subprocess.run([exe, cfg_a, "--dim", "2", "--no-vti", "--checkpoint", ck, "--checkpoint-every", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows_a = (tmp_path / "a" / "diagnostics.csv").read_text().splitlines()[1:]
    per_cycle = ov["corrosion_steps_per_check"] // ov["output_every_corr"]
    assert len(rows_a) >= 3 * per_cycle and os.path.exists(ck + "_c0002.pdck")
    cfg_b = refapi.write_cfg(base, dict(ov, use_implicit=0, output_dir=str(tmp_path / "b")), str(tmp_path / "b.cfg"))
    r = subprocess.run([exe, cfg_b, "--dim", "2", "--no-vti", "--resume", ck + "_c0002"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows_b = (tmp_path / "b" / "diagnostics.csv").read_text().splitlines()
    assert rows_b == rows_a[2 * per_cycle:]
