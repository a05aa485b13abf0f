"""AMR cloud (shipped params_amr.cfg): device time per NS loop body / ARD loop body.
usage: python tools/time_amr.py   (CPU side: tests/time_amr_reference.py)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pd_mg_pin_corrosion_b200 import amr as A            # noqa: E402
from pd_mg_pin_corrosion_b200.config import Config       # noqa: E402

cfg = Config.load(os.path.join(ROOT, "configs", "params_amr.cfg"), {"use_implicit": 0}, quiet=True)
t0 = time.perf_counter()
g = A.AmrGrid(cfg)
g.build_amr()
t1 = time.perf_counter()
g.build_neighbors_celllist()
t2 = time.perf_counter()
i = g.info
print(f"AMR cloud: {i.N_total} nodes (fine {i.n_fine}, coarse {i.n_coarse}, fictitious {i.n_fict}), {i.nnz} CSR entries; "
      f"build_amr {t1 - t0:.3f} s, build_neighbors_celllist {t2 - t1:.3f} s (host)")
g.device_init(0)
nt = g.get("node_type")
A.initialize_fields(g, np.zeros(i.N_total, np.uint8), np.zeros(i.N_total, np.uint8))
dt = g.ns_compute_dt()
g.ns_iterate(200, dt)
t0 = time.perf_counter(); g.ns_iterate(2000, dt); t_ns = (time.perf_counter() - t0) / 2000
dtc = g.ard_compute_dt()
g.ard_iterate(100, dtc)
t0 = time.perf_counter(); g.ard_iterate(1000, dtc); t_ard = (time.perf_counter() - t0) / 1000
print(f"device: NS loop body {1e6 * t_ns:.1f} us, ARD loop body {1e6 * t_ard:.1f} us per iteration")
print("(the reference's own loops on the host cores: python tests/time_amr_reference.py)")

# implicit branch on the cloud (pdamr_implicit_*): assemble + a few backward-Euler steps at the shipped step sizes
if "--implicit" in sys.argv:
    gid, gb, pr, _ = A.generate_grains(g)
    A.initialize_fields(g, gb, pr)
    g.ns_iterate(3000, dt)
    g.update_fictitious()
    g.ard_set_volume_loss(0.0)
    t0 = time.perf_counter(); g.implicit_assemble(); t_asm = time.perf_counter() - t0
    print(f"implicit: assemble {1e3 * t_asm:.2f} ms (salt flags + {i.nnz} bond weights)")
    for dt_impl in (0.6, 30.0, 30.0):
        g.inlet_bc(); g.outlet_bc(); g.wall_conc_bc()
        for pc in (1, 0):
            C0 = g.get_field("C")
            t0 = time.perf_counter()
            info = g.implicit_step(dt_impl, tol=1e-10, restart=50, max_iters=400, precond=pc)
            t_step = time.perf_counter() - t0
            print(f"implicit step dt = {dt_impl:g} s, precond {'axial sweep' if pc else 'none'}: {info.iters} GMRES iterations, "
                  f"|res| = {info.rel_res:.2e}, converged {info.converged}, {1e3 * t_step:.1f} ms")
            if pc == 1:
                C1 = g.get_field("C")
                g.set_field("C", C0)
        g.set_field("C", C1)
        t0 = time.perf_counter(); g.smooth_conc(); g.update_fictitious(); t_sm = time.perf_counter() - t0
        print(f"  smoother + IDW refresh {1e6 * t_sm:.0f} us")
