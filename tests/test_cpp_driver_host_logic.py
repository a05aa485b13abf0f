"""CPU test of the HOST logic of the C++ driver on the AMR cloud (host/amr_run.cpp inside host/pd_corrosion_gpu): the
binary runs against tests/fake_pdgpu -- a stand-in libpdgpu.so whose pdamr_* entry points are served by the compiled
reference (oracle/_ref/libpdrefimp2d.so) -- so everything the driver itself does (grain generation on the cloud,
initialize_fields, cycle structure, batches between output points, diagnostics / snapshot cadence, frame numbers, PVD
and CSV writing, the VTU writer, D_map bookkeeping) must reproduce EVERY output file of the reference's own main() byte
for byte, in the explicit and in the implicit branch.  The operators are checked on the device in the -m gpu tests."""
import glob
import os
import subprocess

import pytest

from oracle import refapi
from test_amr import AMR_CASES, IMPLICIT_AMR_RUN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "host", "pd_corrosion_gpu")
pytestmark = pytest.mark.skipif(not (refapi.have_ref(2, implicit=True) and os.path.exists(EXE)),
                                reason="needs oracle/_ref/libpdrefimp2d.so and host/pd_corrosion_gpu (build())")


@pytest.fixture(scope="module")
def fake_lib_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("fake_pdgpu")
    src = os.path.join(ROOT, "tests", "fake_pdgpu")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-o", str(d / "libpdgpu.so"),
                           os.path.join(src, "fake_pdgpu.cpp"), os.path.join(src, "fake_stubs.cpp"), "-ldl"])
    return str(d)


@pytest.mark.parametrize("branch", ["explicit", "implicit"])
def test_cpp_amr_driver_reproduces_reference_main(branch, fake_lib_dir, tmp_path):
    base, ov = AMR_CASES["amr_ratio2"]
    if branch == "implicit":
        ov = dict(ov, **dict(IMPLICIT_AMR_RUN, T_final=4.5e-4))
    else:
        ov = dict(ov, use_implicit=0, D_grain=5e-11, D_gb=5e-9, C_thresh=0.999, corrosion_steps_per_check=40, flow_max_iters=120,
                  T_final=3.2e-4, output_every_corr=10)
    cfg_ref = refapi.write_cfg(base, dict(ov, output_dir=str(tmp_path / "ref")), str(tmp_path / "ref.cfg"))
    refapi._lib(2, True).ref_set_threads(1)               # the implicit build serves both runs: same code, one thread
    assert refapi.run_reference_main(2, cfg_ref, implicit=True) == 0
    cfg_got = refapi.write_cfg(base, dict(ov, output_dir=str(tmp_path / "got")), str(tmp_path / "got.cfg"))
    env = dict(os.environ, LD_LIBRARY_PATH=fake_lib_dir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""),
               PD_FAKE_REF_LIB=refapi.ref_lib_path(2, True), OMP_NUM_THREADS="1")
    r = subprocess.run([EXE, cfg_got, "--dim", "2"], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    names = sorted(os.path.basename(f) for f in glob.glob(str(tmp_path / "ref" / "*")))
    assert names == sorted(os.path.basename(f) for f in glob.glob(str(tmp_path / "got" / "*")))
    assert sum(n.endswith(".vtu") for n in names) >= 8 and {"diagnostics.csv", "mass_loss.csv", "simulation.pvd", "flow.pvd"} <= set(names)
    for n in names:
        a, b = open(tmp_path / "ref" / n, "rb").read(), open(tmp_path / "got" / n, "rb").read()
        if n.endswith(".pvd"):
            a, b = a.replace(str(tmp_path / "ref").encode(), b""), b.replace(str(tmp_path / "got").encode(), b"")
        assert a == b, n
    assert len(open(tmp_path / "got" / "diagnostics.csv").read().splitlines()) >= 6


@pytest.mark.parametrize("branch", ["explicit", "implicit"])
def test_cpp_lattice_driver_reproduces_reference_main(branch, fake_lib_dir, tmp_path):
    """host/main.cpp + host/coupling.cpp (2D lattice, one rank, host grain generation): diagnostics.csv, mass_loss.csv,
    the VTI series (written by the reference's writer from the reference-served state with the DRIVER's grain_id / D_map
    bookkeeping) and both PVD collections against the reference's own main()."""
    import helpers as H
    dim, base, cov = H.CASES["2d_default"]
    ov = dict(cov, D_grain=5e-11, D_gb=5e-9, C_thresh=0.999, flow_max_iters=200)
    if branch == "implicit":
        ov.update(use_implicit=1, corrosion_steps_per_check=6, T_final=5e-4, implicit_dt_max=0.004, implicit_dt_fraction=0.5,
                  diagnostic_every=2, implicit_output_every=3)
    else:
        ov.update(use_implicit=0, corrosion_steps_per_check=50, T_final=2.6e-4, output_every_corr=10)
    cfg_ref = refapi.write_cfg(base, dict(ov, output_dir=str(tmp_path / "ref")), str(tmp_path / "ref.cfg"))
    refapi._lib(2, True).ref_set_threads(1)
    assert refapi.run_reference_main(2, cfg_ref, implicit=True) == 0
    cfg_got = refapi.write_cfg(base, dict(ov, output_dir=str(tmp_path / "got")), str(tmp_path / "got.cfg"))
    env = dict(os.environ, LD_LIBRARY_PATH=fake_lib_dir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""),
               PD_FAKE_REF_LIB=refapi.ref_lib_path(2, True), PD_FAKE_CFG=cfg_got, OMP_NUM_THREADS="1")
    r = subprocess.run([EXE, cfg_got, "--dim", "2", "--host-grains"], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    names = sorted(os.path.basename(f) for f in glob.glob(str(tmp_path / "ref" / "*")))
    assert names == sorted(os.path.basename(f) for f in glob.glob(str(tmp_path / "got" / "*")))
    assert sum(n.endswith(".vti") for n in names) >= 5 and {"diagnostics.csv", "mass_loss.csv", "simulation.pvd", "flow.pvd"} <= set(names)
    for n in names:
        a, b = open(tmp_path / "ref" / n, "rb").read(), open(tmp_path / "got" / n, "rb").read()
        if n.endswith(".pvd"):
            a, b = a.replace(str(tmp_path / "ref").encode(), b""), b.replace(str(tmp_path / "got").encode(), b"")
        assert a == b, n
    assert len(open(tmp_path / "got" / "diagnostics.csv").read().splitlines()) >= 5
