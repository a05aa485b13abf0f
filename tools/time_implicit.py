"""Implicit ARD step (csrc/implicit.cu) on a 3D lattice after a short flow relaxation: assembly, adaptive dt, GMRES
iterations and time per step.  usage: python tools/time_implicit.py [--fine] [ns_iters]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pd_mg_pin_corrosion_b200 import lib as L_, solver as S   # noqa: E402
from pd_mg_pin_corrosion_b200.config import Config            # noqa: E402

fine = "--fine" in sys.argv
nums = [a for a in sys.argv[1:] if a.isdigit()]
ns_iters = int(nums[0]) if nums else 200
cfg = Config.load(os.path.join(ROOT, "configs", "params_fine.cfg" if fine else "params.cfg"), {}, quiet=True)
L = L_.load()
grid = S.Grid(3)
grid.build(cfg)
fields = S.Fields(); fields.bind(grid)
L_.check(L.pdgpu_fields_init(grid.ctx, None, None))
ns = S.PD_NS_Solver(); ns.init(grid, cfg)
dt = ns.compute_dt(fields, grid, cfg)
ns.iterate(fields, grid, cfg, ns_iters, dt)
imp = S.PD_ARD_ImplicitSolver()
imp.init(grid, cfg)
print(f"3D lattice {grid.Nx} x {grid.Ny} x {grid.Nz} = {grid.N_total} nodes, {ns_iters} NS iterations of flow", flush=True)
t0 = time.perf_counter(); imp.assemble(fields, grid, cfg); grid.sync(); t1 = time.perf_counter()
dti = imp.compute_adaptive_dt(fields, grid, cfg); t2 = time.perf_counter()
print(f"assemble {1e3 * (t1 - t0):.1f} ms, adaptive dt {dti:.3f} s in {1e3 * (t2 - t1):.1f} ms", flush=True)
for k in range(3):
    t0 = time.perf_counter()
    imp.step(fields, grid, cfg, dti)
    grid.sync()
    t1 = time.perf_counter()
    print(f"implicit step {k}: GMRES {imp.last.iters} iterations, |res| {imp.last.rel_res:.2e}, converged {imp.last.converged}, "
          f"{1e3 * (t1 - t0):.1f} ms", flush=True)
