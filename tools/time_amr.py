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
