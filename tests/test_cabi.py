"""The C-ABI library: loads, exports every symbol include/pdgpu.h declares, host-only helpers
agree with the oracle bit for bit, and -- without a GPU -- fails loudly instead of falling
back to anything."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
from pd_mg_pin_corrosion_b200 import lib as L_
from pd_mg_pin_corrosion_b200.config import PdConfig


def test_library_exports_every_declared_symbol():
    L = L_.load()
    names = L_.declared_symbols()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.pdgpu_version() == 100
    assert sum(n.startswith("pdamr_") for n in names) >= 24          # the AMR context incl. pdamr_implicit_*


def test_host_library_exports_declared_symbols():
    """libpdhost.so (host/): every extern "C" function host/grains.h and host/vtu.h declare"""
    import os
    import re
    from pd_mg_pin_corrosion_b200 import grains as G
    H_ = G._load()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = set(re.findall(r'extern "C" \w+ (pdhost_[a-z0-9_]+)\s*\(', open(os.path.join(root, "host", "grains.h")).read()))
    names |= set(re.findall(r"\b(pdhost_[a-z0-9_]+)\s*\(", open(os.path.join(root, "host", "vtu.h")).read()))
    assert {"pdhost_generate_grains", "pdhost_generate_grains_cloud", "pdhost_write_vtu", "pdhost_init_dmap"} <= names
    missing = [n for n in sorted(names) if not hasattr(H_, n)]
    assert not missing, missing


def test_struct_layout_matches_header():
    # 29 doubles + 8 ints, no padding surprises
    assert C.sizeof(PdConfig) == 29 * 8 + 8 * 4
    assert C.sizeof(L_.PdGridInfo) == 9 * 4 + 4 + 8 * 2 + 6 * 8 + 3 * 8 + 3 * 8


@pytest.mark.parametrize("case", ["2d_default", "2d_poiseuille", "2d_offgrid", "3d_small", "3d_offgrid", "3d_default"])
def test_grid_extents_and_stencil_match_oracle(case):
    L = L_.load()
    dim, cfg, _ = H.load_cfg(case)
    s = cfg.to_struct()
    nx, ny, nz = C.c_int(), C.c_int(), C.c_int()
    org = (C.c_double * 3)()
    L_.check(L.pdgpu_grid_extents(C.byref(s), dim, C.byref(nx), C.byref(ny), C.byref(nz), org))
    if case == "3d_default":
        assert (nx.value, ny.value, nz.value) == (67, 67, 287)
        return
    p = H.make_port(case)
    assert (nx.value, ny.value, nz.value) == (p.Nx, p.Ny, p.Nz)
    assert tuple(org) == p.origin
    n = C.c_int()
    L_.check(L.pdgpu_stencil(C.byref(s), dim, C.byref(n), None, None, None, None))
    assert n.value == p.n_off == (36 if dim == 2 else 178)
    d = np.zeros((n.value, 3), np.int32); dist = np.zeros(n.value); ev = np.zeros((n.value, dim)); vol = np.zeros(n.value)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    L_.check(L.pdgpu_stencil(C.byref(s), dim, C.byref(n), vp(d), vp(dist), vp(ev), vp(vol)))
    assert np.array_equal(d, p.off_d)
    assert dist.tobytes() == p.off_dist.tobytes()
    assert ev.tobytes() == p.off_evec.tobytes()
    assert vol.tobytes() == p.off_vol.tobytes()


def test_partition_tiles_the_axis():
    L = L_.load()
    for n_axial, nranks in [(707, 1), (707, 2), (707, 8), (287, 4), (2807, 8), (64, 8)]:
        prev = 0
        for r in range(nranks):
            a0, a1 = C.c_int(), C.c_int()
            L_.check(L.pdgpu_partition(n_axial, nranks, r, C.byref(a0), C.byref(a1)))
            assert a0.value == prev and a1.value > a0.value
            assert abs((a1.value - a0.value) - n_axial / nranks) < 1.0
            prev = a1.value
        assert prev == n_axial
    a0, a1 = C.c_int(), C.c_int()
    assert L.pdgpu_partition(3, 8, 0, C.byref(a0), C.byref(a1)) != 0
    assert b"pdgpu_partition" in L.pdgpu_last_error()


def test_no_gpu_means_loud_failure_not_fallback():
    L = L_.load()
    n = C.c_int(-1)
    rc = L.pdgpu_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    _, cfg, _ = H.load_cfg("2d_default")
    s = cfg.to_struct()
    ctx = C.c_void_p()
    assert L.pdgpu_create(C.byref(s), 2, 0, C.byref(ctx)) != 0
    assert not ctx.value
    msg = L.pdgpu_last_error().decode()
    assert "no CUDA device" in msg and "no CPU fallback" in msg
    from pd_mg_pin_corrosion_b200 import solver as S
    with pytest.raises(L_.PdGpuError):
        S.Grid(2).build(cfg)
    # null-context calls are errors too, never silent no-ops
    assert L.pdgpu_ns_step(None, 1e-8) != 0


def test_balanced_slab_partition():
    """pdgpu_partition_balanced (what pdgpu_create_slab uses): contiguous cover of the axial planes, equal COST per
    rank with the wire planes surcharged, deterministic, slabs thick enough for the halo; surcharge 0 = equal counts."""
    import os
    import subprocess
    import sys
    from pd_mg_pin_corrosion_b200.config import Config
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    L = L_.load()
    for base, ov in (("params_fine.cfg", {}), ("params_fine.cfg", {"L_wire": 3200e-6, "L_upstream": 4000e-6, "L_downstream": 4000e-6}),
                     ("params.cfg", {})):
        cfg = Config.load(os.path.join(H.CONFIG_DIR, base), dict(ov, use_implicit=0), quiet=True)
        s = cfg.to_struct()
        Nx, Ny, Nz = C.c_int(), C.c_int(), C.c_int()
        org = (C.c_double * 3)()
        L_.check(L.pdgpu_grid_extents(C.byref(s), 3, C.byref(Nx), C.byref(Ny), C.byref(Nz), org))
        z = org[2] + np.arange(Nz.value) * cfg.dx
        cost = 1.0 + 0.048 * ((z >= -cfg.m_ratio * cfg.dx) & (z <= cfg.L_wire + cfg.m_ratio * cfg.dx))
        for nranks in (2, 4, 8):
            a0, a1 = C.c_int(), C.c_int()
            prev, loads, counts = 0, [], []
            for r in range(nranks):
                L_.check(L.pdgpu_partition_balanced(C.byref(s), 3, nranks, r, C.byref(a0), C.byref(a1)))
                assert a0.value == prev and a1.value - a0.value >= 2 * cfg.m_ratio + 2
                prev = a1.value
                loads.append(cost[a0.value:a1.value].sum()); counts.append(a1.value - a0.value)
            assert prev == Nz.value
            assert max(loads) - min(loads) <= 2.1, (base, nranks, loads)           # within two planes of equal cost
            if nranks >= 4:
                assert max(counts) > min(counts) + 1, counts                        # wire slabs are thinner
    code = ("import ctypes as C, os, sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from pd_mg_pin_corrosion_b200 import lib as L_; from pd_mg_pin_corrosion_b200.config import Config\n"
            "L = L_.load(); cfg = Config.load(%r, {'use_implicit': 0}, quiet=True); s = cfg.to_struct()\n"
            "a0, a1, b0, b1 = C.c_int(), C.c_int(), C.c_int(), C.c_int()\n"
            "for r in range(8):\n"
            "    L_.check(L.pdgpu_partition_balanced(C.byref(s), 3, 8, r, C.byref(a0), C.byref(a1)))\n"
            "    L_.check(L.pdgpu_partition(707, 8, r, C.byref(b0), C.byref(b1)))\n"
            "    assert (a0.value, a1.value) == (b0.value, b1.value)\n"
            % (ROOT, os.path.join(ROOT, "tests"), os.path.join(H.CONFIG_DIR, "params_fine.cfg")))
    env = dict(os.environ, PDGPU_SLAB_PIN_COST="0")
    assert subprocess.run([sys.executable, "-c", code], env=env).returncode == 0
