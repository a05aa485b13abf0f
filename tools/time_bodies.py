"""Times the NS and ARD loop bodies separately (CUDA events around pdgpu_*_iterate) on the bench
workload under option sets.  usage: python tools/time_bodies.py "overlap=1" "overlap=0" ..."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pd_mg_pin_corrosion_b200 import lib as L_, solver as S   # noqa: E402
from pd_mg_pin_corrosion_b200.config import Config            # noqa: E402

small = "--small" in sys.argv
big = "--big" in sys.argv   # params_fine cross-section at dx = 1 um (307 x 307), short tube
sets = [a for a in sys.argv[1:] if not a.startswith("--")] or ["overlap=1"]
ov = {"use_implicit": 0}
if big:
    ov.update({"dx": 1.0e-6, "L_wire": 60e-6, "L_upstream": 40e-6, "L_downstream": 40e-6})
cfg = Config.load(os.path.join(ROOT, "configs", "params.cfg" if small else "params_fine.cfg"), ov, quiet=True)
L = L_.load()
grid = S.Grid(3)
grid.build(cfg)
fields = S.Fields(); fields.bind(grid)
L_.check(L.pdgpu_fields_init(grid.ctx, None, None))
ns, ard = S.PD_NS_Solver(), S.PD_ARD_Solver()
ns.init(grid, cfg); ard.init(grid, cfg)
dt = ns.compute_dt(fields, grid, cfg); dtc = ard.compute_dt(fields, grid, cfg)
ms = C.c_float()
for st in sets:
    for kv in st.split(","):
        k, v = kv.split("=")
        grid.set_option(k, int(v))
    L_.check(L.pdgpu_ns_iterate(grid.ctx, 4, dt)); L_.check(L.pdgpu_ard_iterate(grid.ctx, 4, dtc))
    out = []
    for which in (0, 1):
        L_.check(L.pdgpu_timer_start(grid.ctx))
        if which == 0:
            L_.check(L.pdgpu_ns_iterate(grid.ctx, 20, dt))
        else:
            L_.check(L.pdgpu_ard_iterate(grid.ctx, 20, dtc))
        L_.check(L.pdgpu_timer_stop(grid.ctx, C.byref(ms)))
        out.append(ms.value / 20)
    print(f"{st:40s} ns body {out[0]:.3f} ms  ard body {out[1]:.3f} ms  sum {out[0]+out[1]:.3f}", flush=True)
# outlet BC alone (pre-pass + both sweeps)
for ok, g in ((3, 8), (3, 4), (3, 2), (2, 0)):
    grid.set_option("outlet_kernel", ok)
    if g:
        grid.set_option("outlet_rows_g", g)
    L_.check(L.pdgpu_bc_outlet(grid.ctx))
    L_.check(L.pdgpu_timer_start(grid.ctx))
    for _ in range(10):
        L_.check(L.pdgpu_bc_outlet(grid.ctx))
    L_.check(L.pdgpu_timer_stop(grid.ctx, C.byref(ms)))
    print(f"outlet_kernel={ok} rows_g={g}: apply_outlet_bc {ms.value / 10:.3f} ms")
for name, fn in (("inlet", L.pdgpu_bc_inlet), ("wall", L.pdgpu_bc_wall), ("solid", L.pdgpu_bc_solid)):
    L_.check(fn(grid.ctx))
    L_.check(L.pdgpu_timer_start(grid.ctx))
    for _ in range(10):
        L_.check(fn(grid.ctx))
    L_.check(L.pdgpu_timer_stop(grid.ctx, C.byref(ms)))
    print(f"apply_{name}_bc {ms.value / 10:.3f} ms (incl. one host sync per call)")
