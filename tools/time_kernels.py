"""Times the two bond kernels alone (CUDA events, pdgpu_time_kernel) on the bench workload under
option sets.  usage: python tools/time_kernels.py "ns_kernel=1" "ns_kernel=2" ... [--small]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pd_mg_pin_corrosion_b200 import lib as L_, solver as S   # noqa: E402
from pd_mg_pin_corrosion_b200.config import Config            # noqa: E402

small = "--small" in sys.argv
sets = [a for a in sys.argv[1:] if not a.startswith("--")] or ["ns_kernel=1"]
cfg = Config.load(os.path.join(ROOT, "configs", "params.cfg" if small else "params_fine.cfg"), {"use_implicit": 0}, quiet=True)
L = L_.load()
grid = S.Grid(3)
grid.build(cfg)
fields = S.Fields(); fields.bind(grid)
L_.check(L.pdgpu_fields_init(grid.ctx, None, None))
ns, ard = S.PD_NS_Solver(), S.PD_ARD_Solver()
ns.init(grid, cfg); ard.init(grid, cfg)
dt = ns.compute_dt(fields, grid, cfg); dtc = ard.compute_dt(fields, grid, cfg)
L_.check(L.pdgpu_ns_iterate(grid.ctx, 2, dt)); L_.check(L.pdgpu_ard_iterate(grid.ctx, 2, dtc))
for st in sets:
    for kv in st.split(","):
        k, v = kv.split("=")
        grid.set_option(k, int(v))
    a, b = C.c_float(), C.c_float()
    L_.check(L.pdgpu_time_kernel(grid.ctx, 0, 10, C.byref(a)))
    L_.check(L.pdgpu_time_kernel(grid.ctx, 1, 10, C.byref(b)))
    print(f"{st:40s} ns {a.value:.3f} ms  ard {b.value:.3f} ms", flush=True)
