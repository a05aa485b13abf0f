"""device-side VTI formatting (csrc/vti.cu) against the reference's own
VTKWriter::write (oracle/_ref) / its plain-C restatement -- byte for byte."""
import ctypes as C
import filecmp
import math
import os
import struct

import numpy as np
import pytest

import helpers as H
from test_gpu_parity import gpu_side

pytestmark = pytest.mark.gpu


def _printf_g(v: float) -> str:
    if math.isnan(v) or math.isinf(v):          # safe_val, src/vtk_writer.cpp:8-14
        v = 0.0
    if v != 0.0 and abs(v) < 1e-300:
        v = 0.0
    return "%g" % v


def _device_g(vals: np.ndarray) -> list:
    from pd_mg_pin_corrosion_b200 import lib as L_, solver as S
    dim, cfg, _ = H.load_cfg("2d_default")
    grid = S.Grid(dim)
    grid.build(cfg)
    cells = np.zeros((vals.size, 16), np.uint8)
    L_.check(L_.load().pdgpu_format_g(grid.ctx, vals.ctypes.data_as(C.c_void_p), vals.size,
                                      cells.ctypes.data_as(C.c_void_p)))
    grid.close()
    return [bytes(r).split(b"\0", 1)[0].decode() for r in cells]


def test_format_g_special_values():
    specials = [0.0, -0.0, 1.0, -1.0, 0.5, 100000.0, 999999.0, 999999.5, 999999.4999999999, 1000000.0, 1000005.0,
                100000.5, 123456.5, 123455.5, 1e5, 1e6, 1e-4, 1e-5, 9.9999949999e-5, 9.999995e-5, 0.0001, 0.00012345675,
                1234565.0, 12345650.0, 1.2345650e10, 2.5, 0.3, 1 / 3, 2 / 3, 1e22, 1e23, 1e-300, 0.99e-300, 5e-324,
                1.7976931348623157e308, 2.2250738585072014e-308, float("nan"), float("inf"), -float("inf"),
                1e100, 1e-100, 9.999995e99, 9.9999949999999e99, 1.0000005e-7, 123456789012345678.0,
                0.1, 0.2, 1000.0, 999.9999999, 1000.0000001, 5.8958e+00, 3.5336e-08, 4.4e-16, 1e-25]
    specials += [float(2 ** k) for k in range(-60, 61, 7)] + [float(10 ** k) for k in range(0, 23)]
    specials += [(2 * n + 1) * 5.0 for n in (100000, 123456, 499999)]            # exact ties through an inexact 1e-1
    specials += [(2 * n + 1) * 50.0 for n in (100000, 123456, 499999)]           # 1e-2
    specials += [(2 * n + 1) * 5.0 ** 10 * 2.0 ** 9 for n in (100000, 123457)]   # 1e-10
    vals = np.array(specials, np.float64)
    got = _device_g(vals)
    exp = [_printf_g(float(v)) for v in vals]
    bad = [(repr(float(v)), g, e) for v, g, e in zip(vals, got, exp) if g != e]
    assert not bad, bad[:10]


def test_format_g_random_bit_patterns():
    rng = np.random.default_rng(7)
    n = 2_000_000
    bits = rng.integers(0, 2 ** 64, n, dtype=np.uint64)
    vals = bits.view(np.float64).copy()
    # a second population near the decimal half-way points of 6-digit numbers
    base = rng.integers(100000, 1000000, n // 4).astype(np.float64) + 0.5
    scale = 10.0 ** rng.integers(-12, 12, n // 4)
    near = base * scale * (1.0 + rng.integers(-3, 4, n // 4) * 2.0 ** -52)
    vals = np.concatenate([vals, near, -near])
    got = _device_g(vals)
    bad = []
    for v, g in zip(vals.tolist(), got):
        e = _printf_g(v)
        if g != e:
            bad.append((struct.pack(">d", v).hex(), g, e))
            if len(bad) > 10:
                break
    assert not bad, bad


@pytest.mark.parametrize("case,iters,steps", [("2d_default", 60, 10), ("3d_small", 20, 6), ("2d_dissolve", 300, 120)])
def test_vti_file_is_byte_identical(case, iters, steps, tmp_path):
    ref = H.make_ref(case)
    dt = ref.ns_compute_dt()
    ref.ns_iterate(iters, dt)
    ref.ard_iterate(steps, ref.ard_compute_dt())
    if case == "2d_dissolve":
        ref.phase_change()
    ref.ns_step(dt)          # refreshes the reference's `pressure` member from the current rho (src/pd_ns.cpp:84)
    S, cfg, grid, fields = gpu_side(case, ref=ref, upload=False)
    if case == "2d_dissolve":
        grid.set_node_types(ref.get("node_type"))
    for n in ("rho", "vel", "C", "phase"):
        fields.set(n, ref.get(n))
    if hasattr(ref, "write_vti") and hasattr(ref, "lib"):       # compiled reference
        ref.write_vti(str(tmp_path / "ref.vti"))
        gid, dmap = ref.get("grain_id"), ref.get("D_map")
    else:                                                       # plain-C restatement
        gid = np.full(ref.N, -1, np.int32)
        dmap = np.zeros(ref.N)
        ref.write_vti(str(tmp_path / "ref.vti"), gid, dmap)
    from pd_mg_pin_corrosion_b200 import lib as L_
    nbytes, ms = C.c_longlong(), C.c_float()
    L_.check(L_.load().pdgpu_vti_write(grid.ctx, str(tmp_path / "gpu.vti").encode(),
                                       np.ascontiguousarray(gid, np.int32).ctypes.data_as(C.c_void_p),
                                       np.ascontiguousarray(dmap, np.float64).ctypes.data_as(C.c_void_p),
                                       C.byref(nbytes), C.byref(ms)))
    grid.close()
    a, b = (tmp_path / "ref.vti").read_bytes(), (tmp_path / "gpu.vti").read_bytes()
    if a != b:
        la, lb = a.split(b"\n"), b.split(b"\n")
        diff = [(i, x, y) for i, (x, y) in enumerate(zip(la, lb)) if x != y][:8]
        raise AssertionError(f"{len(la)} vs {len(lb)} lines; first differences: {diff}")
    assert nbytes.value > 0


@pytest.mark.parametrize("case", ["2d_poiseuille", "3d_small"])
def test_vti_golden_hash(case, tmp_path):
    """Device writer on the synthetic state against the sha256 of the reference writer's file
    (tests/golden/vti.json, generated in the build container by make_golden.py)."""
    import hashlib
    import json
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vti.json")))[case]
    S, cfg, grid, fields = gpu_side(case, ref=None)
    st = H.synthetic_state(grid.N_total, grid.dim)
    from pd_mg_pin_corrosion_b200 import lib as L_
    L = L_.load()
    L_.check(L.pdgpu_fields_init(grid.ctx, st["is_gb"].ctypes.data_as(C.c_void_p), st["is_precip"].ctypes.data_as(C.c_void_p)))
    for n in ("rho", "vel", "C", "phase"):
        fields.set(n, st[n])
    path = str(tmp_path / "gpu.vti")
    L_.check(L.pdgpu_vti_write(grid.ctx, path.encode(), st["grain_id"].ctypes.data_as(C.c_void_p),
                               st["D_map"].ctypes.data_as(C.c_void_p), None, None))
    grid.close()
    data = open(path, "rb").read()
    assert len(data) == gold["bytes"]
    assert hashlib.sha256(data).hexdigest() == gold["sha256"]
