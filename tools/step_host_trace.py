"""Timeline of pdgpu_step_host on the bench workload (diagnostics): per chunk upload/compute/download
completion times.  usage: python tools/step_host_trace.py [n_chunks] [--small]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pd_mg_pin_corrosion_b200 import lib as L_, solver as S   # noqa: E402
from pd_mg_pin_corrosion_b200.config import Config            # noqa: E402

n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 16
small = "--small" in sys.argv
cfg = Config.load(os.path.join(ROOT, "configs", "params.cfg" if small else "params_fine.cfg"), {"use_implicit": 0}, quiet=True)
L = L_.load()
grid = S.Grid(3)
grid.build(cfg)
fields = S.Fields(); fields.bind(grid)
L_.check(L.pdgpu_fields_init(grid.ctx, None, None))
ns, ard = S.PD_NS_Solver(), S.PD_ARD_Solver()
ns.init(grid, cfg); ard.init(grid, cfg)
dt = ns.compute_dt(fields, grid, cfg); dtc = ard.compute_dt(fields, grid, cfg)
N = grid.N_total
host = {}
for name, shape in (("rho", (N,)), ("vel", (N, 3)), ("C", (N,))):
    t = torch.empty(shape, dtype=torch.float64, pin_memory=True)
    host[name] = t.numpy(); host["_" + name] = t
    L_.check(L.pdgpu_fields_download(grid.ctx, S._FIELD_IDS[name], host[name].ctypes.data_as(C.c_void_p)))
import time
for _ in range(5):
    t0 = time.perf_counter()
    S.step_host(grid, dt, dtc, host["rho"], host["vel"], host["C"], n_chunks)
    print(f"wall {1e3 * (time.perf_counter() - t0):.3f} ms")
buf = (C.c_double * (3 * 256))()
n = C.c_int()
L_.check(L.pdgpu_step_host_trace(grid.ctx, buf, 3 * 256, C.byref(n)))
print("chunks", n.value)
print(" k   up_done  cmp_done down_done  [ms]")
for k in range(n.value):
    print(f"{k:2d} {buf[3*k]:9.3f} {buf[3*k+1]:9.3f} {buf[3*k+2]:9.3f}")
