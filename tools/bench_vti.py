#!/usr/bin/env python
"""SURVEY.md 8(f)-1 measurement: VTI snapshot of the current state -- device-side formatting
(pdgpu_vti_write) next to the reference's own VTKWriter::write (oracle/_ref, or its plain-C
restatement) on the same state.  Prints one JSON line.

    python tools/bench_vti.py [--fine]      # 3D params.cfg (1.29 M nodes) / params_fine (17.4 M nodes)
"""
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pd_mg_pin_corrosion_b200 import lib as L_, solver as S   # noqa: E402
from pd_mg_pin_corrosion_b200.config import Config            # noqa: E402

fine = "--fine" in sys.argv
cfg_name = "params_fine.cfg" if fine else "params.cfg"
cfg = Config.load(os.path.join(ROOT, "configs", cfg_name), {"use_implicit": 0}, quiet=True)
L = L_.load()
grid = S.Grid(3)
grid.build(cfg)
fields = S.Fields(); fields.bind(grid)
L_.check(L.pdgpu_fields_init(grid.ctx, None, None))
ns, ard = S.PD_NS_Solver(), S.PD_ARD_Solver()
ns.init(grid, cfg); ard.init(grid, cfg)
dt = ns.compute_dt(fields, grid, cfg)
L_.check(L.pdgpu_ns_iterate(grid.ctx, 20, dt))
L_.check(L.pdgpu_ard_iterate(grid.ctx, 10, ard.compute_dt(fields, grid, cfg)))
N = grid.N_total
tmp = tempfile.mkdtemp(prefix="pdvti_")
path = os.path.join(tmp, "gpu.vti")
nbytes, ms = C.c_longlong(), C.c_float()
walls, fmts = [], []
for rep in range(3):
    t0 = time.perf_counter()
    L_.check(L.pdgpu_vti_write(grid.ctx, path.encode(), None, None, C.byref(nbytes), C.byref(ms)))
    walls.append(time.perf_counter() - t0)
    fmts.append(ms.value)
size = os.path.getsize(path)
fmt_ms = min(fmts)
# bytes the formatting kernels move: fields in (7 doubles + 5 bytes + grain id), slots out + in, text out
alg_bytes = N * (7 * 8 + 4 + 4) + nbytes.value
res = {"metric": "vti_snapshot_text_bytes_per_s", "workload": f"3D {cfg_name} ({grid.Nx}x{grid.Ny}x{grid.Nz})",
       "nodes": int(N), "file_bytes": int(size), "device_format_ms": fmt_ms,
       "device_format_gbs": alg_bytes / (fmt_ms * 1e-3) / 1e9, "algorithmic_bytes": int(alg_bytes),
       "call_wall_s": min(walls), "call_text_mb_per_s": size / min(walls) / 1e6,
       "what": "pdgpu_vti_write: format on device, chunked D2H, 3 pwrite threads, to " + tmp}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    res["hbm_peak_gbs"] = peaks.get("hbm_gbs")
    res["frac_of_hbm"] = res["device_format_gbs"] / float(peaks.get("hbm_gbs"))
except Exception:
    pass
# CPU baseline: the reference writer on the same state (bounded: only for the 1.29 M node case)
if not fine:
    try:
        from oracle import refapi
        import helpers as H
        have = refapi.have_ref(3)
        r = refapi.RefSim(3, "params.cfg", {}, threads=os.cpu_count() or 1) if have else H.make_port("3d_default")
        kind = "reference" if have else "port"
        if not have:
            r.init_fields()
        # same grains, same state; both sides take one NS step so that `pressure` = EOS(current rho)
        L_.check(L.pdgpu_fields_init(grid.ctx, r.get("is_gb").ctypes.data_as(C.c_void_p),
                                     r.get("is_precip").ctypes.data_as(C.c_void_p)))
        L_.check(L.pdgpu_ns_iterate(grid.ctx, 20, dt))
        L_.check(L.pdgpu_ard_iterate(grid.ctx, 10, ard.compute_dt(fields, grid, cfg)))
        for n in ("rho", "vel", "C"):
            r.set(n, fields.get(n)) if have else getattr(r, n).__setitem__(Ellipsis, fields.get(n))
        L_.check(L.pdgpu_ns_step(grid.ctx, dt))
        r.ns_step(dt)
        gid = r.get("grain_id") if have else np.full(N, -1, np.int32)
        dmap = r.get("D_map") if have else np.zeros(N)
        L_.check(L.pdgpu_vti_write(grid.ctx, path.encode(), gid.ctypes.data_as(C.c_void_p),
                                   dmap.ctypes.data_as(C.c_void_p), None, None))
        rp = os.path.join(tmp, "ref.vti")
        t = r.write_vti(rp) if have else r.write_vti(rp, gid, dmap)
        res["cpu_baseline"] = {"kind": kind, "seconds": t, "text_mb_per_s": os.path.getsize(rp) / t / 1e6, "cores": 1,
                               "sample": "same state, whole file"}
        a, b = open(rp, "rb").read(), open(path, "rb").read()
        res["file_identical_to_cpu_baseline"] = bool(a == b)
        res["arrays_identical"] = [int(x == y) for x, y in zip(a.split(b"</DataArray>"), b.split(b"</DataArray>"))]
    except Exception as e:   # noqa: BLE001
        res["cpu_baseline"] = {"error": repr(e)}
print(json.dumps(res))
