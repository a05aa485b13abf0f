"""2D solve_steady timing (BASELINE configs 1-2): loop body per iteration under option sets, whole
solve_steady wall time.  usage: python tools/time_2d.py [--fine | --cfg=params_poiseuille.cfg] [--solve] "ns2d=1" "ns2d=0" ..."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pd_mg_pin_corrosion_b200 import lib as L_, solver as S   # noqa: E402
from pd_mg_pin_corrosion_b200.config import Config            # noqa: E402

fine = "--fine" in sys.argv
cfg_name = "params_fine.cfg" if fine else "params.cfg"
for a in sys.argv[1:]:
    if a.startswith("--cfg="):
        cfg_name = a[6:]
sets = [a for a in sys.argv[1:] if not a.startswith("--")] or ["graph=1"]
cfg = Config.load(os.path.join(ROOT, "configs", cfg_name), {}, quiet=True)
L = L_.load()
grid = S.Grid(2)
grid.build(cfg)
print(f"{cfg_name}: 2D lattice {grid.Nx} x {grid.Ny} = {grid.N_total} nodes", flush=True)
fields = S.Fields(); fields.bind(grid)
L_.check(L.pdgpu_fields_init(grid.ctx, None, None))
ns = S.PD_NS_Solver(); ns.init(grid, cfg)
dt = ns.compute_dt(fields, grid, cfg)
ms = C.c_float()
iters = int(os.environ.get("ITERS", "2000"))
for st in sets:
    for kv in st.split(","):
        k, v = kv.split("=")
        grid.set_option(k, int(v))
    L_.check(L.pdgpu_ns_iterate(grid.ctx, int(os.environ.get("WARM", "4000")), dt))   # long enough for the SM clock to ramp up
    L_.check(L.pdgpu_timer_start(grid.ctx))
    L_.check(L.pdgpu_ns_iterate(grid.ctx, iters, dt))
    L_.check(L.pdgpu_timer_stop(grid.ctx, C.byref(ms)))
    print(f"{st:30s} {1e3 * ms.value / iters:8.2f} us per NS iteration", flush=True)
    # ARD loop body (src/coupling.cpp:232-240) on the same lattice
    ard = S.PD_ARD_Solver(); ard.init(grid, cfg)
    dtc = ard.compute_dt(fields, grid, cfg)
    L_.check(L.pdgpu_ard_iterate(grid.ctx, 500, dtc))
    L_.check(L.pdgpu_timer_start(grid.ctx))
    L_.check(L.pdgpu_ard_iterate(grid.ctx, iters, dtc))
    L_.check(L.pdgpu_timer_stop(grid.ctx, C.byref(ms)))
    print(f"{st:30s} {1e3 * ms.value / iters:8.2f} us per ARD loop body", flush=True)
    if "--solve" in sys.argv:
        L_.check(L.pdgpu_fields_init(grid.ctx, None, None))
        t0 = time.perf_counter()
        r = ns.solve_steady(fields, grid, cfg, verbose=False)
        t1 = time.perf_counter()
        print(f"{st:30s} solve_steady: {r} in {t1 - t0:.3f} s", flush=True)
        L_.check(L.pdgpu_fields_init(grid.ctx, None, None))
