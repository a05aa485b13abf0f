"""Generate the golden fixtures from the UNMODIFIED reference (oracle/_ref, built from
/root/reference/src by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py            # step fixtures (seconds)
    python tests/golden/make_golden.py --steady   # + steady-state known answers (minutes)

Outputs (committed): tests/golden/steps_<case>.npz, tests/golden/steady.json,
tests/golden/diagnostics_2d_dissolve.csv
"""
import hashlib
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H  # noqa: E402
from oracle import refapi  # noqa: E402

STRIDE = 5
NS_ITERS, ARD_STEPS = 20, 10


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def step_fixture(case: str) -> None:
    dim, base, ov = H.CASES[case]
    r = refapi.RefSim(dim, base, ov, threads=4)
    out = {"dims": np.array([r.Nx, r.Ny, r.Nz, r.N, r.nnz]), "node_type": r.get("node_type"),
           "is_gb": np.packbits(r.get("is_gb")), "is_precip": np.packbits(r.get("is_precip")),
           "origin": np.array(r.origin)}
    meta = {"csr_sha": {n: sha(r.get(n)) for n in ("nbr_offset", "nbr_index", "nbr_dist", "nbr_evec", "nbr_vol")}}
    # wall-mirror table by the index trick
    N = r.N
    r.set("rho", np.arange(N) + 0.25)
    r.wall_bc()
    rr = r.get("rho")
    nt = out["node_type"]
    mir = np.full(N, -1, np.int32)
    w = nt == 2
    mir[w] = np.where(np.modf(rr[w])[0] == 0.25, (rr[w] - 0.25).astype(np.int64), -1)
    out["wall_mirror"] = mir
    r.lib.ref_fields_init(r.h)
    dt = r.ns_compute_dt()
    r.ns_iterate(NS_ITERS, dt)
    dtc = r.ard_compute_dt()
    r.ard_iterate(ARD_STEPS, dtc)
    meta.update({"dt_ns": dt, "dt_ard": dtc, "ns_iters": NS_ITERS, "ard_steps": ARD_STEPS, "stride": STRIDE,
                 "c0": r.cfg["c0"], "U_in": r.cfg["U_in"]})
    for n in ("rho", "vel", "C"):
        out[n] = r.get(n)[::STRIDE].copy()
        meta[n + "_sha"] = sha(r.get(n))
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, f"steps_{case}.npz"), **out)
    print(case, "N", r.N, "nnz", r.nnz, "dt", dt, dtc)


def vti_fixture() -> None:
    """sha256 of the reference's own VTKWriter::write output on a synthetic state"""
    res = {}
    for case in ("2d_poiseuille", "3d_small"):
        dim, base, ov = H.CASES[case]
        r = refapi.RefSim(dim, base, ov, threads=4)
        st = H.synthetic_state(r.N, dim)
        for n, v in st.items():
            r.set(n, v)
        r.ns_step(r.ns_compute_dt())       # `pressure` member := EOS(rho) (src/pd_ns.cpp:84)
        path = os.path.join(tempfile.mkdtemp(), "g.vti")
        r.write_vti(path)
        data = open(path, "rb").read()
        res[case] = {"sha256": hashlib.sha256(data).hexdigest(), "bytes": len(data),
                     "pressure_sha": sha(r.get("pressure"))}
        os.unlink(path)
        print(case, "vti bytes", len(data))
    json.dump(res, open(os.path.join(HERE, "vti.json"), "w"), indent=1)


def steady_fixture() -> None:
    res = {}
    for case in ("2d_poiseuille", "2d_default"):
        dim, base, ov = H.CASES[case]
        r = refapi.RefSim(dim, base, ov, threads=4)
        iters = r.ns_solve_steady()
        v = r.get("vel")
        nt = r.get("node_type")
        res[case] = {"iters": int(iters), "vmax_fluid": float(np.sqrt((v[nt == 0] ** 2).sum(1)).max()),
                     "vel_sha": sha(v), "vel_sample": v[::97].tolist(), "rho_sample": r.get("rho")[::97].tolist()}
        print(case, "iters", iters)
    json.dump(res, open(os.path.join(HERE, "steady.json"), "w"), indent=1)


def diagnostics_fixture(case: str = "2d_dissolve") -> None:
    """Whole run of the reference's own main() on the dissolving synthetic config."""
    dim, base, ov = H.CASES[case]
    tmp = tempfile.mkdtemp(prefix="pdgold_")
    ov = dict(ov, use_implicit=0, output_dir=os.path.join(tmp, "out"))
    cfg_path = refapi.write_cfg(base, ov)
    rc = refapi.run_reference_main(dim, cfg_path)
    assert rc == 0
    shutil.copy(os.path.join(tmp, "out", "diagnostics.csv"), os.path.join(HERE, f"diagnostics_{case}.csv"))
    shutil.rmtree(tmp)
    os.unlink(cfg_path)


IMPLICIT_RUN = {"use_implicit": 1, "D_grain": 5e-11, "D_gb": 5e-9, "C_thresh": 0.999, "corrosion_steps_per_check": 6,
                "flow_max_iters": 300, "T_final": 1.2e-3, "implicit_dt_max": 0.004, "implicit_dt_fraction": 0.5,
                "diagnostic_every": 1, "implicit_output_every": 1000000}


def implicit_diagnostics_fixture() -> None:
    """Whole run of the reference's own main() with use_implicit = 1 (src/coupling.cpp:154-216) on 2D params.cfg with
    the overrides above (those of tests/test_gpu_implicit.py::test_whole_implicit_coupled_run).  The library is
    oracle/_ref/libpdrefimp2d.so: the unmodified src/pd_ard_implicit.cpp compiled against the Eigen work-alike
    oracle/eigen_min/ (the solve meets the reference's 1e-10 tolerance; the rows are printed with 7 digits)."""
    dim, base, ov = H.CASES["2d_default"]
    tmp = tempfile.mkdtemp(prefix="pdgold_")
    cfg_path = refapi.write_cfg(base, dict(ov, **IMPLICIT_RUN, output_dir=os.path.join(tmp, "out")))
    assert refapi.run_reference_main(dim, cfg_path, implicit=True) == 0
    shutil.copy(os.path.join(tmp, "out", "diagnostics.csv"), os.path.join(HERE, "diagnostics_2d_implicit.csv"))
    shutil.rmtree(tmp)
    os.unlink(cfg_path)


if __name__ == "__main__":
    if "--implicit" in sys.argv:
        implicit_diagnostics_fixture()
        sys.exit(0)
    for c in ("2d_default", "2d_poiseuille", "2d_offgrid", "3d_small", "3d_offgrid"):
        step_fixture(c)
    diagnostics_fixture()
    diagnostics_fixture("3d_dissolve")
    vti_fixture()
    if "--steady" in sys.argv:
        steady_fixture()
