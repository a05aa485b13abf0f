"""CPU tests of the oracle itself (no GPU): the plain-C restatement (oracle/pd_oracle.c)
against (a) the golden vectors generated from the unmodified reference
(tests/golden/make_golden.py) and (b) oracle/_ref where it is present.

This is what "pins" the oracle: every function of pd_oracle.c is checked against the
reference's own output before any GPU parity claim rests on it.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import helpers as H
from oracle import refapi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLD_CASES = ["2d_default", "2d_poiseuille", "2d_offgrid", "3d_small", "3d_offgrid"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_gold(case):
    z = np.load(os.path.join(GOLD, f"steps_{case}.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


@pytest.mark.parametrize("case", GOLD_CASES)
def test_port_matches_golden(case):
    z, meta = load_gold(case)
    p = H.make_port(case)
    Nx, Ny, Nz, N, nnz = [int(v) for v in z["dims"]]
    assert (p.Nx, p.Ny, p.Nz, p.N) == (Nx, Ny, Nz, N)
    assert np.array_equal(p.origin, z["origin"])
    # bit-exact: node classification, CSR neighbour list, wall-mirror table
    assert np.array_equal(p.node_type, z["node_type"])
    off, idx, dist, evec, vol = p.csr()
    assert int(off[-1]) == nnz
    got = {"nbr_offset": off.astype(np.int32), "nbr_index": idx, "nbr_dist": dist, "nbr_evec": evec, "nbr_vol": vol}
    for name, want in meta["csr_sha"].items():
        assert sha(got[name]) == want, name
    assert np.array_equal(p.wall_mirror, z["wall_mirror"])
    # fields after 20 NS loop bodies + 10 ARD loop bodies from initialize_fields
    is_gb = np.unpackbits(z["is_gb"])[:N]
    is_pr = np.unpackbits(z["is_precip"])[:N]
    p.init_fields(is_gb, is_pr)
    dt = p.ns_compute_dt()
    assert dt == meta["dt_ns"]
    p.ns_iterate(meta["ns_iters"], dt)
    dtc = p.ard_compute_dt()
    assert abs(dtc - meta["dt_ard"]) <= 1e-13 * dtc
    p.ard_iterate(meta["ard_steps"], dtc)
    s = meta["stride"]
    for name in ("rho", "vel", "C"):
        e = H.rel_err(getattr(p, name)[::s], z[name])
        assert e <= 1e-12, f"{name}: {e:.2e}"


def test_reference_known_answers():
    """Grid facts measured on the compiled reference."""
    z, _ = load_gold("2d_default")
    assert [int(v) for v in z["dims"]] == [67, 287, 1, 19229, 673896]
    assert np.bincount(z["node_type"], minlength=6).tolist() == [15520, 1280, 2009, 180, 240, 0]
    assert int(np.unpackbits(z["is_gb"])[:19229].sum()) == 583
    assert int(np.unpackbits(z["is_precip"])[:19229].sum()) == 34
    z, _ = load_gold("2d_poiseuille")
    assert [int(v) for v in z["dims"]] == [87, 127, 1, 11049, 386696]
    assert np.bincount(z["node_type"], minlength=6).tolist() == [9720, 0, 762, 243, 324, 0]


@pytest.mark.skipif(not refapi.have_ref(2), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("case", ["2d_default", "2d_poiseuille", "2d_offgrid", "3d_small"])
def test_port_vs_compiled_reference(case):
    """Operator by operator: plain-C port vs the unmodified reference objects."""
    r = H.make_ref(case)
    H.perturbed_state(r, seed=7)
    p = H.make_port(case, state_from=r)
    assert np.array_equal(p.node_type, r.get("node_type"))
    for op in ("inlet_bc", "outlet_bc", "wall_bc", "solid_bc", "wall_conc_bc"):
        getattr(r, op)()
        getattr(p, op)()
        for f in ("rho", "vel", "C"):
            assert H.rel_err(getattr(p, f), r.get(f)) <= 1e-15, (op, f)
    # smooth_boundary_concentration (implicit-branch BC): in place and order dependent;
    # the port's sequential sweep must equal the reference (4 OpenMP threads) bit for bit, twice over
    for _ in range(2):
        r.smooth_conc()
        p.smooth_conc()
        assert np.array_equal(p.C, r.get("C")), "smooth_conc"
    dt = r.ns_compute_dt()
    assert p.ns_compute_dt() == dt
    r.ns_step(dt)
    p.ns_step(dt)
    for f in ("rho_new", "vel_new", "pressure"):
        assert H.rel_err(getattr(p, f), r.get(f)) <= 1e-14, f
    dtc = r.ard_compute_dt()
    assert abs(p.ard_compute_dt() - dtc) <= 1e-15 * dtc
    C = r.get("C")
    nt = r.get("node_type")
    C[(nt == 0) & (np.arange(C.size) % 37 == 0)] = 0.95   # exercise the salt layer
    r.set("C", C)
    p.set("C", C)
    r.ard_step(dtc)
    p.ard_step(dtc)
    assert H.rel_err(p.C_new, r.get("C_new")) <= 1e-14


@pytest.mark.skipif(not refapi.have_ref(2), reason="oracle/_ref not built")
def test_port_phase_change_vs_reference():
    case = "2d_dissolve"
    r = H.make_ref(case)
    p = H.make_port(case, state_from=r)
    dt = r.ns_compute_dt()
    for cycle in range(2):
        r.ns_iterate(300, dt)
        p.ns_iterate(300, dt)
        dtc = r.ard_compute_dt()
        r.ard_iterate(50, dtc)
        p.ard_iterate(50, dtc)
        before = r.get("node_type")
        n = r.phase_change()
        dissolved = np.nonzero(before != r.get("node_type"))[0]
        assert p.phase_change() == n
        assert np.array_equal(p.last_dissolved, dissolved)
        r.rebuild_neighbors()
        assert H.rel_err(p.C, r.get("C")) <= 1e-11
    assert (p.node_type == 1).sum() < 1280


@pytest.mark.parametrize("case,geom", [("2d_dissolve", "2d_default"), ("3d_dissolve", "3d_small")])
def test_port_coupled_run_matches_reference_diagnostics(case, geom):
    """Whole explicit coupling run (src/coupling.cpp:82-302) driven over the plain-C oracle:
    every numeric column of diagnostics.csv within 1e-6 relative of the reference's own run
    (tests/golden/diagnostics_<case>.csv, written by the reference's main())."""
    gold = np.loadtxt(os.path.join(GOLD, f"diagnostics_{case}.csv"), delimiter=",", skiprows=1)
    z, _ = load_gold(geom)   # same geometry / grains
    n = int(z["dims"][3])
    rows = H.port_coupled_run(case, np.unpackbits(z["is_gb"])[:n], np.unpackbits(z["is_precip"])[:n])
    rows = np.array(rows)
    assert rows.shape == gold.shape
    assert np.array_equal(rows[:, 3], gold[:, 3])                     # solid_nodes exact
    for col in (0, 1, 2, 4, 5):
        rel = np.abs(rows[:, col] - gold[:, col]) / np.maximum(np.abs(gold[:, col]), 1e-300)
        assert rel.max() <= 1e-6, (col, rel.max())


# ---- VTI snapshot writer ------------------------------------------------------
def _synthetic_port(case):
    dim, _, _ = H.CASES[case]
    p = H.make_port(case)
    st = H.synthetic_state(p.N, dim)
    p.init_fields()
    for n in ("rho", "vel", "C", "phase", "is_gb", "is_precip"):
        getattr(p, n)[...] = st[n]
    p.ns_step(p.ns_compute_dt())          # pressure := EOS(rho), as the reference's member (src/pd_ns.cpp:84)
    return p, st


@pytest.mark.parametrize("case", ["2d_poiseuille", "3d_small"])
def test_port_vti_matches_golden_hash(case, tmp_path):
    """The printf("%g") restatement of VTKWriter::write reproduces the sha256 of the file the
    reference's own writer produced for the synthetic state (tests/golden/vti.json)."""
    gold = json.load(open(os.path.join(GOLD, "vti.json")))[case]
    p, st = _synthetic_port(case)
    path = str(tmp_path / "port.vti")
    p.write_vti(path, st["grain_id"], st["D_map"])
    data = open(path, "rb").read()
    assert len(data) == gold["bytes"]
    assert hashlib.sha256(data).hexdigest() == gold["sha256"]


@pytest.mark.skipif(not refapi.have_ref(2), reason="oracle/_ref not built")
def test_port_vti_vs_reference_writer(tmp_path):
    ref = H.make_ref("2d_default")
    ref.ns_iterate(40, ref.ns_compute_dt())
    ref.ard_iterate(5, ref.ard_compute_dt())
    ref.write_vti(str(tmp_path / "ref.vti"))
    port = H.make_port("2d_default", state_from=ref)
    port.write_vti(str(tmp_path / "port.vti"), ref.get("grain_id"), ref.get("D_map"))
    assert (tmp_path / "ref.vti").read_bytes() == (tmp_path / "port.vti").read_bytes()
